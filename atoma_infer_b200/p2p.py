"""Peer-to-peer all-reduce communicator over NVLink: the host side of
`atoma_allreduce_sum_p2p` (csrc/allreduce_p2p.cu, include/atoma_flash.h).

Serves the same call as the reference's `AllReduce` custom op (models/src/multi_gpu.rs:141-179:
`comm.all_reduce(src, &mut dst, &ReduceOp::Sum)` on a contiguous f16/bf16/f32 tensor) with one
kernel over mapped peer memory instead of ncclAllReduce: out-of-place (`dst` is a fresh tensor in
the reference too), stream-ordered, graph-capturable, bit-identical on every rank.

Bootstrap mirrors NCCL's: what the unique id is there, the 64-byte CUDA IPC handle of each rank's
region is here — exchanged once over an existing torch.distributed group (any backend).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_DTYPES = {torch.float16: 0, torch.bfloat16: 1, torch.float32: 2}


class P2PError(RuntimeError):
    pass


class _DeviceBuffer:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 3, "strides": None}


class P2PComm:
    """One region per rank, mapped by every peer.  `slot(i, shape, dtype)` is where the producer
    (the o_proj GEMM) writes this rank's addend; `all_reduce(i, out)` sums slot i of every rank into
    `out`.  i names the local staging slot only; which receive buffer an all-reduce uses is decided on the
    device from the parity of its epoch, so any call order (and any number of all-reduces per captured
    graph) is safe."""

    def __init__(self, rank: int, world: int, device: torch.device, slot_bytes: int, local: int, regions):
        self.rank, self.world, self.device, self.slot_bytes = rank, world, device, slot_bytes
        self._local = local
        self._regions = regions                      # ctypes array of world void*

    @classmethod
    def from_torch_distributed(cls, device, slot_bytes: int) -> "P2PComm":
        import torch.distributed as dist
        L = _lib.lib()
        rank, world = dist.get_rank(), dist.get_world_size()
        if world not in (2, 4, 8):
            raise P2PError(f"peer-to-peer all-reduce supports 2, 4 or 8 ranks, not {world}")
        torch.cuda.set_device(device)
        nbytes = L.atoma_p2p_region_bytes(slot_bytes)
        local = C.c_void_p()
        if L.atoma_p2p_alloc(C.byref(local), nbytes) != 0:
            raise P2PError(L.atoma_last_error().decode())
        handle = C.create_string_buffer(64)
        if L.atoma_p2p_export(local, handle) != 0:
            raise P2PError(L.atoma_last_error().decode())
        handles = [None] * world
        dist.all_gather_object(handles, (rank, bytes(handle.raw)))
        regions = (C.c_void_p * world)()
        for r, h in handles:
            if r == rank:
                regions[r] = local.value
            else:
                p = C.c_void_p()
                if L.atoma_p2p_import(C.create_string_buffer(h, 64), C.byref(p)) != 0:
                    raise P2PError(f"rank {rank} cannot map rank {r}'s region: {L.atoma_last_error().decode()}")
                regions[r] = p.value
        torch.cuda.synchronize(device)
        dist.barrier()   # every region is zeroed and mapped before anybody signals
        return cls(rank, world, device, slot_bytes, local.value, regions)

    def slot(self, i: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if nbytes > self.slot_bytes:
            raise P2PError(f"{nbytes} bytes do not fit the {self.slot_bytes}-byte slot")
        ptr = _lib.lib().atoma_p2p_slot(self._local, self.slot_bytes, i)
        raw = torch.as_tensor(_DeviceBuffer(ptr, self.slot_bytes), device=self.device)
        return raw[:nbytes].view(dtype).reshape(tuple(shape))

    def all_reduce(self, i: int, out: torch.Tensor, stream: int | None = None) -> torch.Tensor:
        """out = sum over ranks of slot i (rank order, fp32 accumulation, one rounding)."""
        if out.dtype not in _DTYPES or not out.is_cuda or not out.is_contiguous():
            raise P2PError("all_reduce needs a contiguous f16/bf16/f32 CUDA tensor (multi_gpu.rs:146-178)")
        if stream is None:
            stream = torch.cuda.current_stream(out.device).cuda_stream
        L = _lib.lib()
        L.atoma_allreduce_sum_p2p(self._regions, self.rank, self.world, self.slot_bytes, i, out.data_ptr(),
                                  out.numel(), _DTYPES[out.dtype], stream)
        if L.atoma_last_status() != 0:
            raise P2PError(L.atoma_last_error().decode())
        return out

    def check(self) -> None:
        """Synchronises the current stream; raises if a peer ever failed to arrive."""
        rc = _lib.lib().atoma_p2p_error(self._local, torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise P2PError("a peer did not arrive at an all-reduce within 30 s" if rc == 1 else f"CUDA error {rc}")

    def destroy(self) -> None:
        if self._local is None:
            return
        import torch.distributed as dist
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier()   # nobody unmaps while a peer may still read
        for r in range(self.world):
            if r != self.rank and self._regions[r]:
                L.atoma_p2p_close(self._regions[r])
        L.atoma_p2p_free(self._local)
        self._local = None
