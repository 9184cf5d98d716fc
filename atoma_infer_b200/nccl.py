"""Direct NCCL binding for the tensor-parallel all-reduce after o_proj.

The reference talks to NCCL itself, through cudarc (`Comm::from_rank(device, rank, world, id)` in
backends/vllm/src/model_executor.rs and `comm.all_reduce(src, &mut dst, &ReduceOp::Sum)` in
models/src/multi_gpu.rs:141-179) — not through a framework's process group.  This module is the same
thing over ctypes: one communicator per GPU process, `ncclAllReduce` enqueued on the CALLER's stream.
That matters on this path: a collective enqueued on the caller's stream is captured into the decode
step's CUDA graph like any other kernel, so the tensor-parallel step replays as ONE graph (64
collectives inside), where `torch.distributed.all_reduce` has to be launched eagerly between
per-layer graphs.

The collective itself stays NCCL (north star: "keeps the repo's existing NCCL allreduce ... reported
but not re-implemented").  Buffers may be allocated with `ncclMemAlloc` and registered as symmetric
windows (`ncclCommWindowRegister`, NCCL >= 2.27) so NCCL can pick its symmetric kernels over NVLink.
Registration is OFF by default (ATOMA_NCCL_SYMMETRIC=1 switches it on): measured on 2 x B200 with NCCL
2.28.9 (scripts/nccl_diag.py) the symmetric kernel is slower for this 256 KiB payload (25.0 vs 21.6 us)
and its bf16 sum is not the correctly rounded a + b (19 % of elements off by one ulp), where the
default ring/tree path returns RN(a + b) exactly.

No CPU fallback: every call raises `NcclError` when the library reports an error.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

UNIQUE_ID_BYTES = 128
NCCL_SUM = 0
NCCL_F16, NCCL_F32, NCCL_BF16 = 6, 7, 9
NCCL_WIN_COLL_SYMMETRIC = 0x01


class NcclError(RuntimeError):
    pass


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * UNIQUE_ID_BYTES)]


_LIB = None


def _find_library() -> str:
    env = os.environ.get("ATOMA_NCCL_LIB")
    if env:
        return env
    try:  # the copy torch itself links (so both see one set of NCCL globals)
        import nvidia.nccl as pkg
        base = Path(list(pkg.__path__)[0]) / "lib" / "libnccl.so.2"
        if base.exists():
            return str(base)
    except Exception:
        pass
    return "libnccl.so.2"


def lib():
    """Loads libnccl and declares the entry points used here."""
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(_find_library(), mode=C.RTLD_GLOBAL)
    L.ncclGetErrorString.restype = C.c_char_p
    L.ncclGetErrorString.argtypes = [C.c_int]
    L.ncclGetVersion.argtypes = [C.POINTER(C.c_int)]
    L.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    L.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    L.ncclCommDestroy.argtypes = [C.c_void_p]
    L.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ncclMemAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.ncclMemFree.argtypes = [C.c_void_p]
    if hasattr(L, "ncclCommWindowRegister"):
        L.ncclCommWindowRegister.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int]
        L.ncclCommWindowDeregister.argtypes = [C.c_void_p, C.c_void_p]
    for name in ("ncclGetVersion", "ncclGetUniqueId", "ncclCommInitRank", "ncclCommDestroy", "ncclAllReduce",
                 "ncclMemAlloc", "ncclMemFree"):
        getattr(L, name).restype = C.c_int
    _LIB = L
    return L


def _check(code: int, what: str) -> None:
    if code != 0:
        raise NcclError(f"{what}: {lib().ncclGetErrorString(code).decode()} (ncclResult_t {code})")


def version() -> int:
    v = C.c_int(0)
    _check(lib().ncclGetVersion(C.byref(v)), "ncclGetVersion")
    return v.value


def get_unique_id() -> bytes:
    """cudarc `Id::new()` — rank 0 creates it and hands it to the other ranks out of band."""
    uid = UniqueId()
    _check(lib().ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
    return bytes(C.string_at(C.addressof(uid), UNIQUE_ID_BYTES))


class _DeviceBuffer:
    """Exposes a raw device allocation to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


class Comm:
    """One NCCL communicator of this process' GPU (cudarc `Comm::from_rank`)."""

    def __init__(self, handle: int, rank: int, world: int):
        self.handle, self.rank, self.world = handle, rank, world
        self._allocs: list[tuple[int, int | None]] = []   # (ptr, window)

    @classmethod
    def from_rank(cls, rank: int, world: int, unique_id: bytes) -> "Comm":
        """The current CUDA device must already be the rank's device (cudaSetDevice)."""
        if len(unique_id) != UNIQUE_ID_BYTES:
            raise NcclError(f"unique id must be {UNIQUE_ID_BYTES} bytes, got {len(unique_id)}")
        if not 0 <= rank < world:
            raise NcclError(f"rank {rank} outside world of {world}")
        uid = UniqueId()
        C.memmove(C.addressof(uid), unique_id, UNIQUE_ID_BYTES)
        h = C.c_void_p()
        _check(lib().ncclCommInitRank(C.byref(h), world, uid, rank), "ncclCommInitRank")
        return cls(h.value, rank, world)

    @staticmethod
    def _dtype(t) -> int:
        import torch
        try:
            return {torch.bfloat16: NCCL_BF16, torch.float16: NCCL_F16, torch.float32: NCCL_F32}[t.dtype]
        except KeyError:
            raise NcclError(f"all_reduce: unsupported dtype {t.dtype} (the reference reduces f16/bf16/f32, "
                            f"multi_gpu.rs:150-178)") from None

    def all_reduce_(self, t, stream: int | None = None):
        """In-place sum over ranks of a contiguous device tensor, enqueued on `stream` (a raw
        cudaStream_t; default: torch's current stream).  multi_gpu.rs:141-179."""
        import torch
        if not t.is_cuda or not t.is_contiguous():
            raise NcclError("all_reduce needs a contiguous CUDA tensor (multi_gpu.rs:146-149: 'input has to be contiguous')")
        if stream is None:
            stream = torch.cuda.current_stream(t.device).cuda_stream
        _check(lib().ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), self._dtype(t), NCCL_SUM,
                                   self.handle, C.c_void_p(stream)), "ncclAllReduce")
        return t

    def symmetric_empty(self, shape, dtype, device):
        """A tensor over `ncclMemAlloc` memory, registered as a symmetric window when
        ATOMA_NCCL_SYMMETRIC=1 and the library supports it (collective call: every rank must make
        it, with the same size)."""
        import torch
        item = torch.empty((), dtype=dtype).element_size()
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = max(n * item, 16)
        ptr = C.c_void_p()
        _check(lib().ncclMemAlloc(C.byref(ptr), nbytes), "ncclMemAlloc")
        win = None
        if hasattr(lib(), "ncclCommWindowRegister") and os.environ.get("ATOMA_NCCL_SYMMETRIC", "0") == "1":
            w = C.c_void_p()
            rc = lib().ncclCommWindowRegister(self.handle, ptr, nbytes, C.byref(w), NCCL_WIN_COLL_SYMMETRIC)
            win = w.value if rc == 0 else None
        self._allocs.append((ptr.value, win))
        raw = torch.as_tensor(_DeviceBuffer(ptr.value, (nbytes,), "|u1"), device=device)
        return raw[: n * item].view(dtype).reshape(tuple(shape))

    def symmetric_windows(self) -> int:
        return sum(1 for _, w in self._allocs if w is not None)

    def destroy(self) -> None:
        if self.handle is None:
            return
        for ptr, win in self._allocs:
            if win is not None:
                lib().ncclCommWindowDeregister(self.handle, win)
            lib().ncclMemFree(ptr)
        self._allocs.clear()
        lib().ncclCommDestroy(self.handle)
        self.handle = None


def init_from_torch_distributed(device) -> Comm:
    """Bootstrap over an already initialised torch.distributed group (any backend): rank 0 makes
    the unique id, everybody receives it, then each rank joins with its own GPU current."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    torch.cuda.set_device(device)
    return Comm.from_rank(rank, world, box[0])
