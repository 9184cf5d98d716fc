"""Kernel-level timing of causal prefill attention (BASELINE.json config 3 geometry): ours
(tcgen05 kernel) vs the reference's kernels (oracle/_ref), CUDA events."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import ops  # noqa: E402


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    dev = "cuda:0"
    H, HK, D = 32, 8, 128
    res = []
    for lens in ([8192], [4096, 4096], [2048] * 4, [1024] * 8):
        T = sum(lens)
        q = torch.randn((T, H, D), device=dev).to(torch.bfloat16)
        k = torch.randn((T, HK, D), device=dev).to(torch.bfloat16)
        v = torch.randn((T, HK, D), device=dev).to(torch.bfloat16)
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
        scale = D ** -0.5
        flops = sum(4 * (n * (n + 1) / 2) * D * H for n in lens)
        t = time_it(lambda: ops.flash_attn_varlen(q, k, v, cu, cu, max(lens), max(lens), scale, True))
        r = {"lens": lens, "flops": flops, "ours_ms": t * 1e3, "ours_TFLOPs": flops / t / 1e12, "path": ops.last_mha_path()}
        sys.path.insert(0, str(ROOT / "tests"))
        from oracle.ref_lib import RefLib
        if RefLib.available() and "--no-ref" not in sys.argv:
            ref = RefLib()
            sms = torch.cuda.get_device_properties(0).multi_processor_count
            plan = ops.flash_attn_varlen_plan(q, k, v, cu, cu, max(lens), max(lens), scale, None, 0, num_sms=sms)
            out = torch.empty_like(q)
            lse = torch.zeros(int(plan.lse_elems), dtype=torch.float32, device=dev)
            t2 = time_it(lambda: ref.run_mha_plan(plan, out, lse, sync=False))
            r.update({"ref_ms": t2 * 1e3, "ref_TFLOPs": flops / t2 / 1e12})
        print(json.dumps(r), flush=True)
        res.append(r)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "microbench_prefill.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
