"""Per-CTA timeline of the projection GEMM (diagnostics; uses atoma_debug_gemm_trace)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import _lib, ops  # noqa: E402


def main():
    dev = "cuda:0"
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for (T, N, K) in [(32, 6144, 4096), (32, 4096, 4096), (1, 3072, 2048)]:
        ws = [torch.randn((N, K), device=dev).to(torch.bfloat16) for _ in range(4)]
        x = torch.randn((T, K), device=dev).to(torch.bfloat16)
        for i in range(8):
            ops.gemm_bf16_tn(x, ws[i % 4])
        torch.cuda.synchronize()
        trace = torch.zeros((sms, 4), dtype=torch.int64, device=dev)
        L.atoma_debug_gemm_trace(trace.data_ptr())
        ops.gemm_bf16_tn(x, ws[0])
        torch.cuda.synchronize()
        L.atoma_debug_gemm_trace(None)
        t = trace.cpu().numpy()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        print(f"T={T} N={N} K={K}: {len(t)} CTAs")
        for name, col in (("start", 0), ("first tile landed", 1), ("last accumulator done", 2), ("end", 3)):
            x_ = (t[:, col] - t0) / 1e3
            print(f"  {name:22s} min {x_.min():6.2f} median {np.median(x_):6.2f} max {x_.max():6.2f} us")


if __name__ == "__main__":
    main()
