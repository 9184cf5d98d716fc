"""BASELINE.json config 5 (KV-swap stress) + the cache-write op at prefill size: achieved bandwidth of
reshape_and_cache_flash / copy_blocks / swap_blocks, ours vs the reference kernels (oracle/_ref) and
vs the reference's per-block memcpy loop for swaps (csrc/src/cache_manager.rs:18-128).  Llama-3.1-8B
cache geometry: block = 16 tokens x 8 kv heads x 128 = 32 KiB per K or V."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import ops  # noqa: E402
from atoma_infer_b200 import _lib  # noqa: E402
from oracle.ref_lib import RefLib  # noqa: E402

LIB = _lib.lib()

DEV = "cuda:0"
L, HK, D, BS = 32, 8, 128, 16


def ev_time(fn, iters=10, warm=3):
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
    res = {}
    ref = RefLib() if RefLib.available() else None
    stream = torch.cuda.current_stream().cuda_stream
    # 64 sequences x 4096 tokens = 16384 blocks (+ spare) per layer; 0.3 x 180 GB would allow ~100k
    NB = 64 * 256 + 2048
    caches = [torch.randn((2, NB, BS, HK, D), device=DEV, dtype=torch.bfloat16) for _ in range(L)]
    block_bytes = BS * HK * D * 2
    rng = np.random.default_rng(0)

    # ---- copy_blocks: P random disjoint pairs across all 32 layers ----
    for P in (1, 64, 1024):
        perm = rng.permutation(NB)[: 2 * P]
        pairs = torch.from_numpy(np.stack([perm[:P], perm[P:]], 1).astype(np.int64)).to(DEV)
        byts = L * P * 2 * block_bytes * 2   # read + written
        kp = torch.tensor([c[0].data_ptr() for c in caches], dtype=torch.int64, device=DEV)
        vp = torch.tensor([c[1].data_ptr() for c in caches], dtype=torch.int64, device=DEV)
        # C ABI with device pointer arrays already uploaded (what the kernel costs) ...
        t = ev_time(lambda: LIB.copy_blocks_bf16(kp.data_ptr(), vp.data_ptr(), pairs.data_ptr(), L, P, BS * HK * D, stream))
        # ... and the operator (validation + pointer upload, as csrc::copy_blocks does per call)
        t_op = ev_time(lambda: ops.copy_blocks([c[0] for c in caches], [c[1] for c in caches], pairs))
        r = {"pairs": P, "bytes_rw": byts, "ours_us": t * 1e6, "ours_GBps": byts / t / 1e9, "ours_op_us": t_op * 1e6}
        if ref is not None:
            t2 = ev_time(lambda: ref.lib.copy_blocks_bf16(kp.data_ptr(), vp.data_ptr(), pairs.data_ptr(), L, P,
                                                          BS * HK * D, stream))
            r.update({"ref_us": t2 * 1e6, "ref_GBps": byts / t2 / 1e9})
        res[f"copy_blocks_P{P}"] = r
        print("copy_blocks", json.dumps(r), flush=True)

    # ---- reshape_and_cache_flash: decode step (32 tokens) and an 8192-token prefill ----
    for T in (32, 8192):
        key = torch.randn((T, HK, D), device=DEV, dtype=torch.bfloat16)
        val = torch.randn((T, HK, D), device=DEV, dtype=torch.bfloat16)
        slots = torch.from_numpy(rng.permutation(NB * BS)[:T].astype(np.int64)).to(DEV)
        byts = 2 * T * HK * D * 2 * 2
        c = caches[0]
        t = ev_time(lambda: LIB.reshape_and_cache_flash(key.data_ptr(), val.data_ptr(), c[0].data_ptr(), c[1].data_ptr(),
                                                        slots.data_ptr(), BS * HK * D, T, HK, D, BS, HK * D, HK * D, 1,
                                                        stream), iters=20)
        t_op = ev_time(lambda: ops.reshape_and_cache_flash(key, val, c[0], c[1], slots), iters=20)
        r = {"tokens": T, "bytes_rw": byts, "ours_us": t * 1e6, "ours_GBps": byts / t / 1e9, "ours_op_us": t_op * 1e6}
        if ref is not None:
            t2 = ev_time(lambda: ref.lib.reshape_and_cache_flash(key.data_ptr(), val.data_ptr(), c[0].data_ptr(),
                                                                 c[1].data_ptr(), slots.data_ptr(), BS * HK * D, T, HK,
                                                                 D, BS, HK * D, HK * D, 1, stream), iters=20)
            r.update({"ref_us": t2 * 1e6, "ref_GBps": byts / t2 / 1e9})
        res[f"reshape_T{T}"] = r
        print("reshape_and_cache_flash", json.dumps(r), flush=True)

    # ---- the host link's own peak: ONE contiguous 512 MiB cudaMemcpyAsync each way, pinned memory (the denominator
    # of the swap numbers below; PCIe Gen5 x16 nominal 63 GB/s per direction) ----
    big_d = torch.empty(512 << 20, dtype=torch.uint8, device=DEV)
    big_h = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
    link = {}
    for name, (dst, src) in (("h2d", (big_d, big_h)), ("d2h", (big_h, big_d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        link[name + "_GBps"] = 3 * big_d.numel() / (time.perf_counter() - t0) / 1e9
    res["host_link_peak"] = link
    print("host_link_peak", json.dumps(link), flush=True)
    del big_d, big_h

    # ---- swap_blocks: whole sequences (256 blocks) of one layer's K, GPU <-> pinned host ----
    for nseq in (1, 8, 64):
        nblk = nseq * 256
        host = torch.empty((nblk + 8, BS, HK, D), dtype=torch.bfloat16).pin_memory()
        src_blocks = rng.permutation(NB)[:nblk]
        out_map = {int(s): i for i, s in enumerate(src_blocks)}     # gpu block -> host block i
        in_map = {i: int(s) for i, s in enumerate(src_blocks)}
        byts = nblk * block_bytes
        c = caches[1][0]

        def swap_out():
            ops.swap_blocks(c, host, out_map)

        def swap_in():
            ops.swap_blocks(host, c, in_map)
        for name, fn in (("out", swap_out), ("in", swap_in)):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t = (time.perf_counter() - t0) / 3
            r = {"blocks": nblk, "bytes": byts, "ours_ms": t * 1e3, "ours_GBps": byts / t / 1e9,
                 "frac_of_host_link_peak": byts / t / 1e9 / link["d2h_GBps" if name == "out" else "h2d_GBps"]}
            # the reference: one memcpy per block, issued from the host (cache_manager.rs:18-128, ops.rs:158-216)
            pairs = list((out_map if name == "out" else in_map).items())
            s_t, d_t = (c, host) if name == "out" else (host, c)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for sblk, dblk in pairs:
                d_t[dblk].copy_(s_t[sblk], non_blocking=True)
            torch.cuda.synchronize()
            t2 = time.perf_counter() - t0
            r.update({"per_block_memcpy_ms": t2 * 1e3, "per_block_memcpy_GBps": byts / t2 / 1e9})
            res[f"swap_{name}_{nseq}seq"] = r
            print(f"swap_{name}", json.dumps(r), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bench_cache_ops.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
