"""Direct NCCL binding (atoma_infer_b200/nccl.py ≙ the reference's cudarc `Comm`,
models/src/multi_gpu.rs:141-179): on CPU only that the library loads, exports what the binding
declares and validates its arguments; the collective itself is exercised by the 2-GPU test below
(skipped on a one-GPU box) and by `bench.py --gpus N`."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from atoma_infer_b200 import nccl  # noqa: E402


def test_library_loads_and_exports():
    L = nccl.lib()
    for name in ("ncclGetUniqueId", "ncclCommInitRank", "ncclCommDestroy", "ncclAllReduce", "ncclMemAlloc",
                 "ncclMemFree", "ncclGetErrorString", "ncclGetVersion"):
        assert hasattr(L, name), name
    assert nccl.version() >= 22000   # the reference's cudarc 0.17 binds NCCL 2.x


def test_argument_validation():
    with pytest.raises(nccl.NcclError, match="128 bytes"):
        nccl.Comm.from_rank(0, 2, b"short")
    with pytest.raises(nccl.NcclError, match="outside world"):
        nccl.Comm.from_rank(2, 2, bytes(128))
    c = nccl.Comm(handle=None, rank=0, world=1)
    with pytest.raises(nccl.NcclError, match="contiguous CUDA tensor"):
        c.all_reduce_(torch.zeros(4))
    c.destroy()  # no handle: a no-op


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # bootstrap only
    comm = nccl.init_from_torch_distributed(dev)
    buf = comm.symmetric_empty((32, 4096), torch.bfloat16, dev)
    g = torch.Generator(device="cpu").manual_seed(7)
    parts = [torch.randn((32, 4096), generator=g).to(torch.bfloat16) for _ in range(world)]
    want = parts[0].float()
    for p in parts[1:]:
        want = (want + p.float())
    # eager, on a side stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        buf.copy_(parts[rank].to(dev))
        comm.all_reduce_(buf)
    s.synchronize()
    eager = buf.float().cpu()
    # captured: copy-in + all-reduce replayed as one CUDA graph
    src = parts[rank].to(dev)
    with torch.cuda.stream(s):
        buf.copy_(src)
        comm.all_reduce_(buf)
    s.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        buf.copy_(src)
        comm.all_reduce_(buf)
    buf.zero_()
    gr.replay()
    torch.cuda.synchronize()
    replayed = buf.float().cpu()
    torch.save({"eager": eager, "replayed": replayed, "want": want}, os.path.join(out_dir, f"r{rank}.pt"))
    del gr
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_all_reduce_two_gpus_eager_and_captured(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = torch.load(tmp_path / f"r{r}.pt")
        # two bf16 addends: the sum is rounded once to bf16 whatever the algorithm
        want = d["want"].to(torch.bfloat16).float()
        assert torch.equal(d["eager"], want)
        assert torch.equal(d["replayed"], want)
