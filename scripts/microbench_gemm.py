"""Kernel-level timing of the projection GEMMs at decode batch 32 (ours vs cuBLAS), CUDA events,
weights rotated over distinct buffers so every call streams from HBM."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import ops  # noqa: E402


def main():
    dev = "cuda:0"
    out = []
    shapes = [(32, 6144, 4096), (32, 4096, 4096), (1, 3072, 2048), (16, 1280, 8192), (16, 8192, 1024),
              (8192, 6144, 4096), (8192, 4096, 4096), (2048, 6144, 4096)]
    if "--decode-only" in sys.argv:
        shapes = [s_ for s_ in shapes if s_[0] <= 64]
    for (T, N, K) in shapes:
        nrot = 16 if T <= 256 else 2
        ws = [torch.randn((N, K), device=dev).to(torch.bfloat16) for _ in range(nrot)]
        x = torch.randn((T, K), device=dev).to(torch.bfloat16)
        y = torch.empty((T, N), device=dev, dtype=torch.bfloat16)

        def time_it(fn, reps=5):
            """CUDA graph of one pass over the rotating weights (no host launch overhead in the timing)."""
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(nrot):
                    fn(i)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(nrot):
                    fn(i)
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e3 / (reps * nrot)
        t_ours = time_it(lambda i: ops.gemm_bf16_tn(x, ws[i % nrot], out=y))
        t_cluster = None
        t_cublas = time_it(lambda i: torch.matmul(x, ws[i % nrot].t(), out=y))
        byts = 2 * (N * K + T * K + T * N)
        flops = 2.0 * T * N * K
        out.append({"T": T, "N": N, "K": K, "ours_us": t_ours, "ours_cluster_us": t_cluster, "cublas_us": t_cublas,
                    "ours_GBps": byts / t_ours / 1e3, "cublas_GBps": byts / t_cublas / 1e3,
                    "ours_TFLOPs": flops / t_ours / 1e6, "cublas_TFLOPs": flops / t_cublas / 1e6})
        print(json.dumps(out[-1]), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "microbench_gemm.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
