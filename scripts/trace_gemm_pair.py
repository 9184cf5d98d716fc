"""Timeline of consecutive projection GEMMs inside ONE CUDA graph (each launch traced into its own
buffer): where does the time between 'kernel i ends' and 'kernel i+1 streams' go?"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import _lib, ops  # noqa: E402


def run(T, N, K, stages, pdl, nk=6):
    dev = "cuda:0"
    L = _lib.lib()
    L.atoma_debug_gemm_stages(stages)
    ops.set_launch_mode(pdl=pdl, const_weights=pdl)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ws = [torch.randn((N, K), device=dev).to(torch.bfloat16) for _ in range(16)]
    x = torch.randn((T, K), device=dev).to(torch.bfloat16)
    y = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
    traces = [torch.zeros((sms, 4), dtype=torch.int64, device=dev) for _ in range(nk)]
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(16):
            ops.gemm_bf16_tn(x, ws[i], out=y)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(nk):
            L.atoma_debug_gemm_trace(traces[i].data_ptr())
            ops.gemm_bf16_tn(x, ws[i], out=y)
        L.atoma_debug_gemm_trace(None)
        for i in range(nk, 16):
            ops.gemm_bf16_tn(x, ws[i], out=y)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    per = a.elapsed_time(b) * 1e3 / (5 * 16)
    ts = [t.cpu().numpy() for t in traces]
    t0 = ts[1][:, 0].min()
    print(f"T={T} N={N} K={K} stages={stages or 'max'} pdl={pdl}: {per:.2f} us per GEMM in the graph")
    for i in range(1, nk):
        t = ts[i]
        f = lambda col, fn: (fn(t[:, col]) - t0) / 1e3
        print(f"  k{i}: start {f(0, np.min):6.2f}..{f(0, np.max):6.2f}  first tile {f(1, np.min):6.2f}/{f(1, np.median):6.2f}  "
              f"last acc {f(2, np.median):6.2f}/{f(2, np.max):6.2f}  end {f(3, np.min):6.2f}..{f(3, np.max):6.2f}")
    L.atoma_debug_gemm_stages(0)
    ops.set_launch_mode(pdl=False, const_weights=False)


if __name__ == "__main__":
    for shape in [(32, 6144, 4096), (32, 4096, 4096)]:
        for stages, pdl in [(0, False), (0, True), (5, False), (5, True)]:
            run(*shape, stages, pdl)
