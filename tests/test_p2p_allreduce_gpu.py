"""Peer-to-peer all-reduce over NVLink (csrc/allreduce_p2p.cu ≙ the reference's AllReduce op over
NCCL, models/src/multi_gpu.rs:141-179).  Needs 2 GPUs (`gpurun --gpus 2`); skipped on a one-GPU
box.  Oracle: the fp32 sum of the ranks' addends in rank order, rounded once — the kernel must match
it BIT FOR BIT on every rank, eagerly and replayed from a CUDA graph, for many consecutive
all-reduces on alternating slots."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _addends(world, shape, dtype, it):
    g = torch.Generator(device="cpu").manual_seed(100 + it)
    return [torch.randn(shape, generator=g).to(dtype) for _ in range(world)]


def _want(parts, dtype):
    acc = torch.zeros_like(parts[0], dtype=torch.float32)
    for p in parts:           # rank order, fp32
        acc = acc + p.float()
    return acc.to(dtype)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from atoma_infer_b200.p2p import P2PComm, P2PError
    shape = (32, 4096)
    comm = P2PComm.from_torch_distributed(dev, 32 * 4096 * 4)
    from atoma_infer_b200 import _lib
    bad = []
    k = 0
    for algo in (1, 2, 0):      # one-shot, two-shot, library's choice
        _lib.lib().atoma_debug_allreduce_algo(algo)
        for dtype in (torch.bfloat16, torch.float16, torch.float32):
            out = torch.empty(shape, dtype=dtype, device=dev)
            for it in range(4):     # eager, alternating slots
                parts = _addends(world, shape, dtype, it)
                comm.slot(k % 2, shape, dtype).copy_(parts[rank].to(dev))
                comm.all_reduce(k % 2, out)
                k += 1
                if not torch.equal(out.cpu(), _want(parts, dtype)):
                    bad.append(("eager", algo, str(dtype), it))
        # ragged size: 1000 elements (2000 bytes = 125 vectors, not a multiple of the world size)
        parts = _addends(world, (1000,), torch.bfloat16, 70 + algo)
        o2 = torch.empty((1000,), dtype=torch.bfloat16, device=dev)
        comm.slot(k % 2, (1000,), torch.bfloat16).copy_(parts[rank].to(dev))
        comm.all_reduce(k % 2, o2)
        k += 1
        if not torch.equal(o2.cpu(), _want(parts, torch.bfloat16)):
            bad.append(("ragged", algo))
    # ragged size: 1000 elements (2000 bytes, a multiple of 16)
    parts = _addends(world, (1000,), torch.bfloat16, 77)
    o2 = torch.empty((1000,), dtype=torch.bfloat16, device=dev)
    comm.slot(k % 2, (1000,), torch.bfloat16).copy_(parts[rank].to(dev))
    comm.all_reduce(k % 2, o2)
    k += 1
    if not torch.equal(o2.cpu(), _want(parts, torch.bfloat16)):
        bad.append(("ragged",))
    # the same staging slot twice in a row: the receive buffer is chosen on the device (epoch parity)
    for rep in range(3):
        parts = _addends(world, (1000,), torch.bfloat16, 90 + rep)
        comm.slot(0, (1000,), torch.bfloat16).copy_(parts[rank].to(dev))
        comm.all_reduce(0, o2)
        if not torch.equal(o2.cpu(), _want(parts, torch.bfloat16)):
            bad.append(("same staging slot", rep))
    # CUDA graph: 7 (copy-in, all-reduce) pairs — an ODD count — replayed 3 times, once per algorithm
    dtype = torch.bfloat16
    NG = 7
    srcs = [p[rank].to(dev) for p in (_addends(world, shape, dtype, 200 + i) for i in range(NG))]
    wants = [_want(_addends(world, shape, dtype, 200 + i), dtype) for i in range(NG)]
    outs = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(NG)]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    torch.cuda.synchronize()
    for algo in (1, 2):
        _lib.lib().atoma_debug_allreduce_algo(algo)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(NG):
                comm.slot((k + i) % 2, shape, dtype).copy_(srcs[i])
                comm.all_reduce((k + i) % 2, outs[i])
        for rep in range(3):
            for o in outs:
                o.zero_()
            g.replay()
            torch.cuda.synchronize()
            for i in range(NG):
                if not torch.equal(outs[i].cpu(), wants[i]):
                    bad.append(("graph", algo, rep, i))
        del g
    _lib.lib().atoma_debug_allreduce_algo(0)
    comm.check()
    torch.save(bad, os.path.join(out_dir, f"r{rank}.pt"))
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_p2p_all_reduce_bit_exact_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    world = 8 if n >= 8 else 4 if n >= 4 else 2   # every world size the bench runs at, up to what the box has
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(tmp_path / f"r{r}.pt") == [], f"rank {r} mismatches"


def test_argument_validation_without_gpu():
    """Host-side checks of the C entry (no launch): bad world / rank / size set a status."""
    import ctypes as C
    from atoma_infer_b200 import _lib
    L = _lib.lib()
    regions = (C.c_void_p * 2)(0x1000, 0x2000)
    L.atoma_allreduce_sum_p2p(regions, 0, 3, 1 << 18, 0, C.c_void_p(0x3000), 1024, 1, None)
    assert L.atoma_last_status() != 0 and b"world must be 2, 4 or 8" in L.atoma_last_error()
    L.atoma_allreduce_sum_p2p(regions, 0, 2, 1 << 18, 0, C.c_void_p(0x3000), 1001, 1, None)
    assert L.atoma_last_status() != 0 and b"multiple of 16 bytes" in L.atoma_last_error()
    # header + 2 staging slots + recv[2 slots][8 source ranks] of 2 x slot bytes (data + flag per 8-byte unit)
    assert L.atoma_p2p_region_bytes(1 << 18) == 4096 + 2 * (1 << 18) + 2 * 8 * 2 * (1 << 18)
    assert L.atoma_last_status() == 0   # a successful call clears the thread's status
