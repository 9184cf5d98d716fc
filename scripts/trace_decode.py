"""Per-CTA timeline of the paged decode kernel at config 2 (diagnostics; uses atoma_debug_decode_trace)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import _lib, ops  # noqa: E402


def main():
    dev = "cuda:0"
    HK = int(os.environ.get("ATOMA_MB_HK", "8"))   # kv heads on this rank (8 / tensor-parallel degree)
    B, H, D, page, MB = 32, 4 * HK, 128, 16, 256
    NB = B * MB + 64
    g = torch.Generator(device=dev).manual_seed(0)
    caches = [torch.randn((2, NB, page, HK, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
              for _ in range(3)]
    q = torch.randn((B, 1, H, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    bt = torch.from_numpy(np.random.default_rng(0).permutation(NB)[: B * MB].astype(np.int32).reshape(B, MB)).to(dev)
    sl = torch.full((B,), 4096, dtype=torch.int32, device=dev)
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    trace = torch.zeros((sms, 8), dtype=torch.int64, device=dev)
    for i in range(6):
        ops.flash_attn_kv_cache_full(q, caches[i % 3][0], caches[i % 3][1], None, D ** -0.5, bt, sl, True)
    torch.cuda.synchronize()
    L.atoma_debug_decode_trace(trace.data_ptr())
    ops.flash_attn_kv_cache_full(q, caches[0][0], caches[0][1], None, D ** -0.5, bt, sl, True)
    torch.cuda.synchronize()
    L.atoma_debug_decode_trace(None)
    t = trace.cpu().numpy()
    t0 = t[:, 1].min()
    start, seg, end = (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    print("CTAs", len(t), "span us", end.max())
    for name, x in (("start", start), ("last segment done", seg), ("end", end)):
        print(f"{name:18s} min {x.min():7.2f} p10 {np.percentile(x, 10):7.2f} median {np.median(x):7.2f} "
              f"p90 {np.percentile(x, 90):7.2f} max {x.max():7.2f}")
    for name, col in (("published", 4), ("arrival counted", 5), ("reduce start", 6), ("reduce end", 7)):
        m = t[:, col] > 0
        x = (t[m, col] - t0) / 1e3
        d = (t[m, col] - t[m, 2]) / 1e3
        print(f"{name:18s} n={m.sum():3d} median {np.median(x):7.2f} max {x.max():7.2f}   after 'segment done': median {np.median(d):6.2f} max {d.max():6.2f}")
    order = np.argsort(end)
    print("slowest CTAs (cta, smid, start, seg, end):")
    for i in order[-8:]:
        print(int(i), int(t[i, 0]), f"{start[i]:.2f} {seg[i]:.2f} {end[i]:.2f}")
    print("fastest CTAs:")
    for i in order[:8]:
        print(int(i), int(t[i, 0]), f"{start[i]:.2f} {seg[i]:.2f} {end[i]:.2f}")
    # by sm id parity / groups
    dur = end - start
    smid = t[:, 0]
    for lo in range(0, 148, 16):
        m = (smid >= lo) & (smid < lo + 16)
        if m.any():
            print(f"smid {lo:3d}-{lo+15:3d}: mean duration {dur[m].mean():7.2f} us (n={m.sum()})")


if __name__ == "__main__":
    main()
