"""The tensor-parallel all-reduce fused into the row-parallel GEMM (SURVEY §8 rows e, f4) — exercised on ONE GPU.

The protocol is pure memory semantics: rank r's `atoma_gemm_tn_push` stores its partial tile into `recv[slot][r]` of EVERY
rank's region, and rank r's `atoma_allreduce_recv_add_rmsnorm` polls its own region.  Nothing in it needs the regions to live
on different devices, so a W-rank all-reduce can be replayed on one device: W regions allocated locally, the W push GEMMs
issued one after the other (each as "rank r" with its own shard of x and w), then the W receive kernels.  This covers what
the real multi-GPU runs cannot on a one-GPU box — the 32-column block layout of the receive buffers for every tile width the
planner picks (the 8B and the 70B shapes: 32- and 64-column tiles), the epoch / slot-parity bookkeeping over repeated
all-reduces, world sizes 2, 4 and 8 — bit for bit against the rank-order fp32 sum rounded once
(models/src/multi_gpu.rs:141-179 computes the same sum with ncclAllReduce).  The NVLink transport itself is covered by
tests/test_tp_step_gpu.py on multi-GPU boxes."""
import ctypes as C
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu

BF, F16 = torch.bfloat16, torch.float16


@pytest.mark.parametrize("world,T,N,K,dtype", [(2, 32, 4096, 2048, BF), (8, 32, 4096, 512, BF), (8, 16, 8192, 1024, BF),
                                               (4, 24, 1024, 512, BF), (4, 7, 2048, 448, F16), (8, 128, 1024, 64, BF),
                                               (2, 1, 8192, 3584, BF)])
def test_push_and_receive_sum_the_ranks_in_order_bit_for_bit(world, T, N, K, dtype):
    from atoma_infer_b200 import _lib, ops
    L = _lib.lib()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(world * 1000 + T + N)
    xs = [torch.randn((T, K), generator=g).to(dtype).to(dev) for _ in range(world)]
    ws = [(torch.randn((N, K), generator=g) * K ** -0.5).to(dtype).to(dev) for _ in range(world)]
    resid = torch.randn((T, N), generator=g).to(dtype).to(dev)
    nw = (1 + 0.1 * torch.randn((N,), generator=g)).to(dtype).to(dev)
    slot_bytes = max(T * N * 2, 32768)
    nbytes = L.atoma_p2p_region_bytes(slot_bytes)
    mem = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]     # header (epoch words) zeroed
    regions = (C.c_void_p * 8)(*[m.data_ptr() for m in mem])
    stream = lambda: torch.cuda.current_stream().cuda_stream
    dt = 1 if dtype == BF else 0

    acc = torch.zeros((T, N), dtype=torch.float32, device=dev)
    for r in range(world):
        acc = acc + ops.gemm_tn(xs[r], ws[r]).float()          # rank order, fp32, each addend rounded to the io type
    want_h, want_n = ops.add_rmsnorm(acc.to(dtype), resid, nw, 1e-5)

    def all_reduce(outs):
        for r in range(world):                                  # send side: every "rank" pushes into every region
            L.atoma_gemm_tn_push(xs[r].data_ptr(), ws[r].data_ptr(), T, N, K, K, K, regions, r, world, slot_bytes, dt, stream())
            assert L.atoma_last_status() == 0, L.atoma_last_error().decode()
        for r in range(world):                                  # receive side: each "rank" reads its own region
            h, n = outs[r]
            L.atoma_allreduce_recv_add_rmsnorm(regions, r, world, slot_bytes, resid.data_ptr(), h.data_ptr(), nw.data_ptr(),
                                               n.data_ptr(), T, N, N, N, N, 1e-5, dt, stream())
            assert L.atoma_last_status() == 0, L.atoma_last_error().decode()

    outs = [(torch.empty_like(resid), torch.empty_like(resid)) for _ in range(world)]
    for it in range(3):                                         # odd and even epochs: both receive slots
        for h, n in outs:
            h.zero_(); n.zero_()
        all_reduce(outs)
        torch.cuda.synchronize()
        for r, (h, n) in enumerate(outs):
            assert torch.equal(h, want_h), f"all-reduce {it}, rank {r}: residual stream differs"
            assert torch.equal(n, want_n), f"all-reduce {it}, rank {r}: normed output differs"
    for m in mem:                                               # no rank gave up waiting
        assert int(m[8:12].view(torch.int32).item()) == 0      # header word 2: the sticky error flag
    # with programmatic dependent launch, replayed from a graph
    ops.set_launch_mode(pdl=True, const_weights=True)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            all_reduce(outs)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            all_reduce(outs)
        for rep in range(2):
            for h, n in outs:
                h.zero_(); n.zero_()
            gr.replay()
            torch.cuda.synchronize()
            for r, (h, n) in enumerate(outs):
                assert torch.equal(h, want_h) and torch.equal(n, want_n), f"graph replay {rep}, rank {r}"
    finally:
        ops.set_launch_mode()
