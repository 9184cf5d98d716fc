"""Decode-time GEMM (csrc/gemm_skinny.cu) at the shapes of a Llama decode step: quick numerics check
against fp32 torch, then kernel timing vs cuBLAS (torch.matmul) — CUDA-graph loops over rotating
weights so every launch streams W from HBM, CUDA events, programmatic dependent launch on for ours
as in the step.  Writes gpurun_out/microbench_skinny.json.

  python scripts/microbench_skinny.py [--quick] [--sweep]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
BF = torch.bfloat16


def time_graph(fn, n, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * n)


def check():
    """max |err| relative to the row scale, vs fp32 matmul of the same bf16 inputs."""
    torch.manual_seed(0)
    res = {}
    for (T, N, K) in [(32, 6144, 4096), (32, 4096, 4096), (1, 3072, 2048), (7, 1000, 512), (64, 4096, 14336),
                      (100, 4096, 1024), (128, 2048, 2048), (32, 128256, 4096)]:
        x = torch.randn((T, K), device=DEV).to(BF)
        w = (torch.randn((N, K), device=DEV) * K ** -0.5).to(BF)
        y = ops.gemm_tn(x, w)
        ref = x.float() @ w.float().t()
        err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
        res[f"plain_{T}x{N}x{K}"] = err
        print(f"plain T={T} N={N} K={K}: rel err {err:.3e}", flush=True)
    # fused qkv + rope + cache vs unfused (bit-exact expected)
    T, H, HK, D, K = 32, 32, 8, 128, 4096
    x = torch.randn((T, K), device=DEV).to(BF)
    w = (torch.randn(((H + 2 * HK) * D, K), device=DEV) * K ** -0.5).to(BF)
    pos = torch.randint(0, 4096, (T,), device=DEV, dtype=torch.int64)
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    ang = torch.arange(4096 + 16, dtype=torch.float32)[:, None] * inv[None, :]
    cos, sin = ang.cos().to(BF).to(DEV), ang.sin().to(BF).to(DEV)
    NB, page = 64, 16
    kv1 = torch.zeros((2, NB, page, HK, D), device=DEV, dtype=BF)
    kv2 = torch.zeros_like(kv1)
    slots = torch.randperm(NB * page, device=DEV)[:T].to(torch.int64)
    slots[3] = -1
    fused = ops.qkv_rope_cache(x, w, H, HK, D, pos, cos, sin, kv1[0], kv1[1], slots)
    qkv = ops.gemm_tn(x, w)
    nq, nkv = H * D, HK * D
    q = qkv[:, :nq].unflatten(1, (H, D))
    k = qkv[:, nq:nq + nkv].unflatten(1, (HK, D))
    v = qkv[:, nq + nkv:].unflatten(1, (HK, D))
    ops.rope_reshape_and_cache(q, k, v, kv2[0], kv2[1], pos, cos, sin, slots)
    torch.cuda.synchronize()
    res["rope_fused_equals_unfused"] = bool(torch.equal(fused, qkv))
    res["rope_cache_equal"] = bool(torch.equal(kv1, kv2))
    print("fused qkv == unfused:", res["rope_fused_equals_unfused"], " caches equal:", res["rope_cache_equal"],
          " max diff", (fused.float() - qkv.float()).abs().max().item(), flush=True)
    # silu
    T, I, K = 32, 14336, 4096
    x = torch.randn((T, K), device=DEV).to(BF)
    w = (torch.randn((2 * I, K), device=DEV) * K ** -0.5).to(BF)
    y = ops.gate_up_silu(x, w)
    gu = (x.float() @ w.float().t()).to(BF).float()
    ref = torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    res["silu_rel_err"] = err
    print(f"gate_up_silu rel err {err:.3e}", flush=True)
    return res


def bench(shapes, label_extra=""):
    out = []
    for (name, T, N, K, kind) in shapes:
        nbytes = N * K * 2
        nrot = max(2, min(16, int(2.5e9 // nbytes)))   # > 2x L2 of distinct weights
        ws = [(torch.randn((N, K), device=DEV) * K ** -0.5).to(BF) for _ in range(nrot)]
        x = torch.randn((T, K), device=DEV).to(BF)
        if kind == "silu":
            y = torch.empty((T, N // 2), device=DEV, dtype=BF)
            ours = lambda i: ops.gate_up_silu(x, ws[i % nrot], out=y)
            y2 = torch.empty((T, N), device=DEV, dtype=BF)
            lib = lambda i: torch.matmul(x, ws[i % nrot].t(), out=y2)
        else:
            y = torch.empty((T, N), device=DEV, dtype=BF)
            ours = lambda i: ops.gemm_tn(x, ws[i % nrot], out=y)
            lib = lambda i: torch.matmul(x, ws[i % nrot].t(), out=y)
        n = nrot * 2
        t_ours = time_graph(ours, n)
        t_lib = time_graph(lib, n)
        byts = 2 * (N * K + T * K + T * N)
        rec = {"shape": name, "T": T, "N": N, "K": K, "ours_us": round(t_ours, 2), "cublas_us": round(t_lib, 2),
               "ours_GBps": round(byts / t_ours / 1e3, 1), "cublas_GBps": round(byts / t_lib / 1e3, 1), "note": label_extra}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del ws
        torch.cuda.empty_cache()
    return out


def main():
    L = _lib.lib()
    res = {"check": check()}
    ops.set_launch_mode(pdl=True, const_weights=True)
    shapes = [("qkv_8b", 32, 6144, 4096, "plain"), ("o_8b", 32, 4096, 4096, "plain"),
              ("gate_up_8b", 32, 28672, 4096, "silu"), ("down_8b", 32, 4096, 14336, "plain"),
              ("lm_head", 32, 128256, 4096, "plain"),
              ("qkv_1b_b1", 1, 3072, 2048, "plain"), ("qkv_70b_tp8", 16, 1280, 8192, "plain"),
              ("o_70b_tp8", 16, 8192, 1024, "plain"), ("qkv_8b_b128", 128, 6144, 4096, "plain"),
              ("qkv_8b_tp8", 32, 768, 4096, "plain"), ("o_8b_tp8", 32, 4096, 512, "plain"),
              ("gate_up_8b_tp8", 32, 3584, 4096, "silu"), ("down_8b_tp8", 32, 4096, 1792, "plain")]
    if "--quick" in sys.argv:
        shapes = shapes[:4]
    res["pdl"] = bench(shapes, "pdl+const_weights")
    if "--sweep" in sys.argv:
        sw = []
        for (np_, kbs, st) in [(0, 2, 0), (0, 4, 0), (0, 8, 0), (0, 16, 0), (0, 0, 2), (64, 0, 0)]:
            L.atoma_debug_skinny_tuning(np_, kbs, st)
            sw += bench([s for s in shapes[:4] if s[4] == "plain"], f"np={np_} kbs={kbs} stages={st}")
        L.atoma_debug_skinny_tuning(0, 0, 0)
        res["sweep"] = sw
    if "--budgets" in sys.argv:
        # ring budget (KiB): <= 100 = half footprint (two CTAs per SM, cross-kernel overlap under PDL)
        bd = []
        for kib in (100, 80, 64):
            L.atoma_debug_skinny_budget(kib)
            bd += bench(shapes[:5], f"budget={kib}KiB")
        L.atoma_debug_skinny_budget(0)
        # how much can ONE SM pull?  force wide tiles so that only 64 / 32 SMs stream
        for np_ in (64, 128):
            L.atoma_debug_skinny_tuning(np_, 0, 0)
            bd += bench([shapes[1], shapes[3]], f"np={np_} (tiles on {4096 // np_} SMs only)")
        L.atoma_debug_skinny_tuning(0, 0, 0)
        res["budgets"] = bd
    if "--half-ring" in sys.argv:
        # Is a 100 KiB ring enough in steady state?  Without PDL consecutive launches do not overlap, so the block
        # scheduler places one CTA per SM (no packing of two CTAs of one launch on an SM): the clean comparison.
        hr = []
        ops.set_launch_mode()
        for kib in (0, 100):
            L.atoma_debug_skinny_budget(kib)
            hr += bench(shapes[:5], f"no pdl, budget={kib}KiB")
        for kbs in (1, 2, 4):
            L.atoma_debug_skinny_tuning(0, kbs, 0)
            hr += bench(shapes[:4], f"no pdl, budget=100KiB kbs={kbs}")
        L.atoma_debug_skinny_tuning(0, 0, 0)
        L.atoma_debug_skinny_budget(0)
        ops.set_launch_mode(pdl=True, const_weights=True)
        res["half_ring"] = hr
    if "--experiments" in sys.argv:
        ex = []
        core = [s for s in shapes[:4] if s[4] == "plain"]
        for (skip, w_ops, kbs) in [(3, 1, 0), (1, 1, 0), (2, 1, 0), (0, 1, 4), (0, 1, 2), (0, 1, 1), (3, 1, 1)]:
            L.atoma_debug_skinny_mode(skip, w_ops)
            L.atoma_debug_skinny_tuning(0, kbs, 0)
            ex += bench(core, f"skip={skip} w_ops={w_ops} kbs={kbs}")
        L.atoma_debug_skinny_mode(0, 1)
        L.atoma_debug_skinny_tuning(0, 0, 0)
        res["experiments"] = ex
        # per-CTA timeline of single launches (no PDL overlap): start -> first stage landed -> accumulator done -> end
        ops.set_launch_mode()
        for (name, T, N, K, kind) in core:
            w = (torch.randn((N, K), device=DEV) * K ** -0.5).to(BF)
            x = torch.randn((T, K), device=DEV).to(BF)
            tr = torch.zeros((148 * 8,), dtype=torch.int64, device=DEV)
            ops.gemm_tn(x, w)
            big = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
            big.zero_()   # flush L2
            torch.cuda.synchronize()
            L.atoma_debug_skinny_trace(tr.data_ptr())
            ops.gemm_tn(x, w)
            torch.cuda.synchronize()
            L.atoma_debug_skinny_trace(None)
            t = tr.view(148, 8).cpu()
            t = t[t[:, 3] > 0]
            t0 = t[:, 0].min()
            d = (t - t0).float() / 1e3
            rec = {"shape": name, "ctas": int(t.shape[0]),
                   "start_us": [round(d[:, 0].min().item(), 2), round(d[:, 0].max().item(), 2)],
                   "first_stage_us": [round(d[:, 1].min().item(), 2), round(d[:, 1].median().item(), 2), round(d[:, 1].max().item(), 2)],
                   "acc_done_us": [round(d[:, 2].min().item(), 2), round(d[:, 2].median().item(), 2), round(d[:, 2].max().item(), 2)],
                   "end_us": [round(d[:, 3].min().item(), 2), round(d[:, 3].median().item(), 2), round(d[:, 3].max().item(), 2)]}
            print("trace", json.dumps(rec), flush=True)
            ex.append(rec)
        ops.set_launch_mode(pdl=True, const_weights=True)
    ops.set_launch_mode()
    res["no_pdl"] = bench(shapes[:4], "no pdl")
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "microbench_skinny.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
