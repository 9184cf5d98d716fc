"""Does programmatic dependent launch overlap the o_proj GEMM with the tail of the decode kernel?
Per-CTA globaltimer traces of both kernels, launched back to back."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import _lib, ops  # noqa: E402


def main():
    dev = "cuda:0"
    B, H, HK, D, page, MB = 32, 32, 8, 128, 16, 256
    NB = B * MB + 64
    g = torch.Generator(device=dev).manual_seed(0)
    kv = torch.randn((2, NB, page, HK, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q = torch.randn((B, 1, H, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    bt = torch.from_numpy(np.random.default_rng(0).permutation(NB)[: B * MB].astype(np.int32).reshape(B, MB)).to(dev)
    sl = torch.full((B,), 4096, dtype=torch.int32, device=dev)
    wo = (torch.randn((4096, 4096), device=dev) / 64).to(torch.bfloat16)
    wq = (torch.randn((6144, 4096), device=dev) / 64).to(torch.bfloat16)
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for mode in ((False, False), (True, False), (True, True)):
        ops.set_launch_mode(*mode)
        for _ in range(3):
            o = ops.flash_attn_kv_cache_full(q, kv[0], kv[1], None, D ** -0.5, bt, sl, True)
            x = ops.gemm_bf16_tn(o.reshape(B, H * D), wo)
            y = ops.gemm_bf16_tn(x, wq)
        torch.cuda.synchronize()
        td = torch.zeros((sms, 8), dtype=torch.int64, device=dev)
        tg = torch.zeros((sms, 4), dtype=torch.int64, device=dev)
        o_buf = torch.empty_like(q)
        x_buf = torch.empty((B, 4096), device=dev, dtype=torch.bfloat16)
        plan = ops.flash_attn_kv_cache_plan(q, kv[0], kv[1], D ** -0.5, block_table=bt, seqlens_k=sl, window_size_right=0)
        import ctypes as C
        stream = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        L.atoma_debug_decode_trace(td.data_ptr())
        L.atoma_debug_gemm_trace(tg.data_ptr())
        L.atoma_mha_execute(C.byref(plan), o_buf.data_ptr(), None, None, None, 0, stream)
        L.atoma_gemm_bf16_tn(o_buf.data_ptr(), wo.data_ptr(), x_buf.data_ptr(), B, 4096, 4096, 4096, 4096, 4096, stream)
        torch.cuda.synchronize()
        L.atoma_debug_decode_trace(None)
        L.atoma_debug_gemm_trace(None)
        d, gm = td.cpu().numpy(), tg.cpu().numpy()
        t0 = d[:, 1].min()
        dend = (d[:, 3] - t0) / 1e3
        gs, gfirst, gacc, gend = [(gm[:, i] - t0) / 1e3 for i in range(4)]
        print(f"pdl={mode[0]} const_w={mode[1]}: decode end median {np.median(dend):.2f} max {dend.max():.2f} | "
              f"gemm start min {gs.min():.2f} median {np.median(gs):.2f} | first tile median {np.median(gfirst):.2f} | "
              f"last acc median {np.median(gacc):.2f} max {gacc.max():.2f} | end max {gend.max():.2f}")
    ops.set_launch_mode(False, False)


if __name__ == "__main__":
    main()


def graph_chain():
    """Per-iteration time of the captured chain decode -> o_proj -> qkv (x8, rotating caches) with and
    without PDL, ours vs cuBLAS GEMMs."""
    dev = "cuda:0"
    B, H, HK, D, page, MB = 32, 32, 8, 128, 16, 256
    NB = B * MB + 64
    n = 8
    kvs = [torch.randn((2, NB, page, HK, D), device=dev, dtype=torch.bfloat16) for _ in range(n)]
    wos = [(torch.randn((4096, 4096), device=dev) / 64).to(torch.bfloat16) for _ in range(n)]
    wqs = [(torch.randn((6144, 4096), device=dev) / 64).to(torch.bfloat16) for _ in range(n)]
    bt = torch.from_numpy(np.random.default_rng(0).permutation(NB)[: B * MB].astype(np.int32).reshape(B, MB)).to(dev)
    sl = torch.full((B,), 4096, dtype=torch.int32, device=dev)
    x0 = torch.randn((B, 4096), device=dev).to(torch.bfloat16)

    def chain(gemm):
        x = x0
        for i in range(n):
            qkv = gemm(x, wqs[i])
            q = qkv[:, :4096].unflatten(1, (H, D)).unsqueeze(1)
            o = ops.flash_attn_kv_cache_full(q, kvs[i][0], kvs[i][1], None, D ** -0.5, bt, sl, True)
            x = gemm(o.reshape(B, H * D), wos[i])
        return x

    for name, gemm in (("atoma", lambda a, w: ops.gemm_bf16_tn(a, w)), ("cublas", lambda a, w: torch.matmul(a, w.t()))):
        for mode in ((False, False), (True, False), (True, True)):
            ops.set_launch_mode(*mode)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                chain(gemm)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                chain(gemm)
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            print(f"graph chain gemm={name} pdl={mode[0]} const_w={mode[1]}: {a.elapsed_time(b) * 1e3 / (5 * n):.2f} us per layer-chain")
    ops.set_launch_mode(False, False)


if __name__ == "__main__":
    graph_chain()


def eager_chain():
    """Same chain launched eagerly through pre-planned direct C-ABI calls (no Python op layer)."""
    import ctypes as C
    dev = "cuda:0"
    B, H, HK, D, page, MB = 32, 32, 8, 128, 16, 256
    NB = B * MB + 64
    n = 8
    L = _lib.lib()
    kvs = [torch.randn((2, NB, page, HK, D), device=dev, dtype=torch.bfloat16) for _ in range(n)]
    wos = [(torch.randn((4096, 4096), device=dev) / 64).to(torch.bfloat16) for _ in range(n)]
    wqs = [(torch.randn((6144, 4096), device=dev) / 64).to(torch.bfloat16) for _ in range(n)]
    bt = torch.from_numpy(np.random.default_rng(0).permutation(NB)[: B * MB].astype(np.int32).reshape(B, MB)).to(dev)
    sl = torch.full((B,), 4096, dtype=torch.int32, device=dev)
    xs = [torch.randn((B, 4096), device=dev).to(torch.bfloat16) for _ in range(2)]
    qkv = torch.empty((B, 6144), device=dev, dtype=torch.bfloat16)
    o = torch.empty((B, 1, H, D), device=dev, dtype=torch.bfloat16)
    q = qkv[:, :4096].unflatten(1, (H, D)).unsqueeze(1)
    plans = [ops.flash_attn_kv_cache_plan(q, kvs[i][0], kvs[i][1], D ** -0.5, block_table=bt, seqlens_k=sl,
                                          window_size_right=0) for i in range(n)]
    stream = torch.cuda.current_stream().cuda_stream

    def chain():
        for i in range(n):
            L.atoma_gemm_bf16_tn(xs[i % 2].data_ptr(), wqs[i].data_ptr(), qkv.data_ptr(), B, 6144, 4096, 4096, 4096, 6144, stream)
            L.atoma_mha_execute(C.byref(plans[i]), o.data_ptr(), None, None, None, 0, stream)
            L.atoma_gemm_bf16_tn(o.data_ptr(), wos[i].data_ptr(), xs[(i + 1) % 2].data_ptr(), B, 4096, 4096, 4096, 4096, 4096, stream)

    for mode in ((False, False), (True, False), (True, True)):
        ops.set_launch_mode(*mode)
        for _ in range(3):
            chain()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            chain()
        b.record()
        torch.cuda.synchronize()
        print(f"eager chain (direct C ABI) pdl={mode[0]} const_w={mode[1]}: {a.elapsed_time(b) * 1e3 / (5 * n):.2f} us per layer-chain")
    ops.set_launch_mode(False, False)


if __name__ == "__main__":
    eager_chain()
