"""Kernel-level timing of the paged decode attention at BASELINE.json config 2 (per layer):
ours vs the reference kernels (oracle/_ref), CUDA events, L2 defeated by rotating caches."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from atoma_infer_b200 import ops  # noqa: E402


def main():
    dev = "cuda:0"
    HK = int(os.environ.get("ATOMA_MB_HK", "8"))   # kv heads on this rank (8 / tensor-parallel degree)
    B, H, D, page, MB = 32, 4 * HK, 128, 16, 256
    ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    NB = B * MB + 64
    nrot = 4
    g = torch.Generator(device=dev).manual_seed(0)
    caches = [torch.randn((2, NB, page, HK, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
              for _ in range(nrot)]
    q = torch.randn((B, 1, H, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    bt = torch.from_numpy(np.random.default_rng(0).permutation(NB)[: B * MB].astype(np.int32).reshape(B, MB)).to(dev)
    sl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    scale = 1.0 / np.sqrt(D)
    bytes_alg = 2 * B * ctx * HK * D * 2 + 2 * B * H * D * 2 + B * MB * 4
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    res = {"ctx": ctx, "kv_heads": HK, "bytes": bytes_alg, "sms": sms}

    def time_it(fn, iters=40, warm=8):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (a, b) in enumerate(evs):
            a.record()
            fn(i)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3

    plans = [ops.flash_attn_kv_cache_plan(q, c[0], c[1], scale, block_table=bt, seqlens_k=sl, window_size_right=0)
             for c in caches]
    import ctypes as C
    from atoma_infer_b200 import _lib
    L = _lib.lib()
    out = torch.empty_like(q)
    stream = torch.cuda.current_stream().cuda_stream

    def ours(i):
        L.atoma_mha_execute(C.byref(plans[i % nrot]), out.data_ptr(), None, None, None, 0, stream)

    if os.environ.get("ATOMA_MB_NO_MERGE"):
        L.atoma_debug_decode_merge(0)
        res["merge"] = False
    med, best = time_it(ours)
    res["ours_us_median"] = med * 1e6
    res["ours_us_best"] = best * 1e6
    res["ours_GBps"] = bytes_alg / med / 1e9
    print(json.dumps(res), flush=True)

    if "--only-ours" in sys.argv:
        return
    # read-only / copy bandwidth probes on the same box (context for the roofline fraction)
    big = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
    big.zero_()
    med, _ = time_it(lambda i: big.view(torch.int32).sum(), iters=10, warm=3)
    res["torch_sum_read_GBps"] = big.numel() / med / 1e9
    dst = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    med, _ = time_it(lambda i: dst.copy_(big[: 1 << 30]), iters=10, warm=3)
    res["torch_copy_rw_GBps"] = 2 * dst.numel() / med / 1e9
    del big, dst
    print(json.dumps(res), flush=True)

    from helpers import RefLib
    if RefLib.available() and "--no-ref" not in sys.argv:
        ref = RefLib()
        lse = torch.zeros(B * H, dtype=torch.float32, device=dev)
        out2 = torch.empty_like(q)

        def theirs(i):
            a = plans[i % nrot].args
            ref.lib.run_mha(
                a.q_ptr, a.k_ptr, a.v_ptr, out2.data_ptr(), lse.data_ptr(), a.alibi_slopes_ptr,
                a.cu_seqlens_q_ptr, a.cu_seqlens_k_ptr, a.is_seqlens_k_cumulative,
                a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.o_batch_stride, a.alibi_slopes_batch_stride,
                a.q_row_stride, a.k_row_stride, a.v_row_stride, a.o_row_stride,
                a.q_head_stride, a.k_head_stride, a.v_head_stride, a.o_head_stride,
                a.num_splits, a.b, a.h, a.h_k, a.d, a.d_rounded, a.softmax_scale, a.scale_softmax_log2,
                a.block_table, a.block_table_batch_stride, a.page_block_size, a.seqused_k,
                a.seqlen_q, a.seqlen_k, a.seqlen_q_rounded, a.seqlen_k_rounded, a.is_bf16, a.is_causal,
                a.window_size_left, a.window_size_right, a.softcap, a.unpadded_lse, a.force_split_kernel, None, None)

        # the reference launches on the legacy default stream; torch events are on torch's stream
        # (blocking stream semantics keep them ordered)
        med, best = time_it(theirs)
        res["ref_us_median"] = med * 1e6
        res["ref_GBps"] = bytes_alg / med / 1e9
        torch.cuda.synchronize()
        ours(0)
        theirs(0)
        torch.cuda.synchronize()
        res["max_abs_diff_vs_ref"] = float((out.float() - out2.float()).abs().max())
    print(json.dumps(res), flush=True)
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"microbench_decode_{ctx}_hk{HK}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
