"""2-GPU diagnostic: numerics and latency of the direct NCCL all-reduce (eager, graph, symmetric on/off)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from atoma_infer_b200 import nccl  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    comm = nccl.init_from_torch_distributed(dev)
    g = torch.Generator(device="cpu").manual_seed(7)
    parts = [torch.randn((32, 4096), generator=g).to(torch.bfloat16) for _ in range(world)]
    exact = sum(p.double() for p in parts)
    want = exact.float().to(torch.bfloat16).float()
    for label, buf in (("symmetric", comm.symmetric_empty((32, 4096), torch.bfloat16, dev)),
                       ("plain", torch.empty((32, 4096), dtype=torch.bfloat16, device=dev))):
        buf.copy_(parts[rank].to(dev))
        comm.all_reduce_(buf)
        torch.cuda.synchronize()
        got = buf.float().cpu()
        nbad = int((got != want).sum())
        err = (got.double() - exact).abs()
        werr = (want.double() - exact).abs()
        # timing: graph of 64 all-reduces
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            comm.all_reduce_(buf)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(64):
                comm.all_reduce_(buf)
        gr.replay(); torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gr.replay()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (5 * 64)
        print(f"[r{rank}] {label}: windows={comm.symmetric_windows()} mismatches vs RN(sum)={nbad}/{got.numel()} "
              f"max|got-exact|={err.max():.4g} max|RN-exact|={werr.max():.4g} got_is_rounder={bool((err <= werr + 1e-12).all())} "
              f"all-reduce 256KiB in graph: {us:.2f} us", flush=True)
        del gr
    comm.destroy()
    # ---- this library's one-shot all-reduce over peer memory ----
    from atoma_infer_b200.p2p import P2PComm
    from atoma_infer_b200 import ops
    p2p = P2PComm.from_torch_distributed(dev, 32 * 4096 * 2)
    out = torch.empty((32, 4096), dtype=torch.bfloat16, device=dev)
    for i in range(2):
        p2p.slot(i, (32, 4096), torch.bfloat16).copy_(parts[rank].to(dev))
    from atoma_infer_b200 import _lib
    for algo, pdl in ((1, False), (1, True), (2, False), (2, True)):
        _lib.lib().atoma_debug_allreduce_algo(algo)
        ops.set_launch_mode(pdl=pdl, const_weights=pdl)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(2):
                p2p.all_reduce(i, out)
        s.synchronize()
        got = out.float().cpu()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(64):
                p2p.all_reduce(i % 2, out)
        gr.replay(); torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gr.replay()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (5 * 64)
        print(f"[r{rank}] p2p {'one' if algo == 1 else 'two'}-shot (pdl={pdl}): mismatches vs RN(fp32 sum)={int((got != want).sum())}/{got.numel()} "
              f"all-reduce 256KiB in graph: {us:.2f} us", flush=True)
        del gr
    ops.set_launch_mode(pdl=False, const_weights=False)
    p2p.check()
    p2p.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
