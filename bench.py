#!/usr/bin/env python
"""Benchmark of the attention hot path at BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config 2 of BASELINE.json): one DECODE STEP of Llama-3.1-8B's attention path, bf16,
batch 32, every sequence at context 4096, block_size 16 — for each of the 32 layers

    qkv = x . Wqkv^T            (models/src/llama.rs:269-271; one fused-N GEMM)
    RoPE on q, k                (llama.rs:218-251, 300-301)        } ours: ONE fused kernel
    reshape_and_cache_flash     (models/src/flash_attention.rs:361) } (atoma_rope_reshape_and_cache)
    paged decode attention      (flash_attention.rs:455 -> flash_attn_kv_cache_full -> run_mha)
    x = attn . Wo^T             (llama.rs:311)  [+ NCCL all-reduce under tensor parallelism]

with a distinct 541 MB KV cache per layer (17.3 GB in total: every layer's K/V comes from HBM, no
L2 flush needed).  MLP, norms and lm_head are outside SURVEY §8's scope table and not timed;
`config.workload` says so.  Metric: decode tokens/s = batch / step time, whole job.

`--impl ours` runs libflashattention (this repo's sm_100a kernels) through the public operator API
(atoma_infer_b200.ops); `--impl reference` runs the REFERENCE's own kernels (oracle/_ref, compiled
from /root/reference/csrc/kernels for sm_100) with the argument tuples its Rust wrappers derive,
launched eagerly from one host thread on the legacy default stream as the reference does (RoPE with
candle's kernel count — 2 index_select + 2 rope + 4 transpose/contiguous copies per layer — as 8
element-wise torch kernels in bf16).  The
reference has no CPU implementation of this path (cpu_fwd bails, csrc/src/lib.rs:1862-1872): its
"CPU-host-timed" number is host wall-clock around the GPU step, which is what that arm reports.
Dense GEMMs are cuBLAS (torch.matmul) in the reference arm, as in the reference (candle -> cuBLAS).

N > 1: tensor parallel over heads exactly like the reference (models/src/llama_nccl.rs:164-171):
each rank holds H/N query heads, Hkv/N kv heads, its slice of every KV cache and of Wqkv / Wo, and
all-reduces the o_proj output with NCCL.  Total work is fixed -> "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

# Llama-3.1-8B attention geometry
L, HIDDEN, H, HK, D, PAGE = 32, 4096, 32, 8, 128, 16
BATCH, CTX = 32, 4096
MB = CTX // PAGE  # block-table width
METRIC = "decode tokens/sec Llama-3.1-8B bf16 b32 ctx4k (attention path: qkv_proj + RoPE + cache write + paged decode attention + o_proj, 32 layers)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=L, help="debug only; the reported config needs 32")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--tp-comm", default="p2p", choices=["p2p", "nccl", "torch"],
                    help="N>1, the all-reduce after o_proj: 'p2p' = this library's one-shot kernel over NVLink peer memory "
                         "(csrc/allreduce_p2p.cu; falls back to 'nccl' if peer mapping fails); 'nccl' = direct ncclAllReduce on "
                         "the step's stream (atoma_infer_b200/nccl.py); both are captured with the step in one CUDA graph; "
                         "'torch' = torch.distributed.all_reduce launched eagerly between per-layer graphs")
    ap.add_argument("--gemm", default="auto", choices=["auto", "atoma", "cublas"])
    ap.add_argument("--ar-algo", type=int, default=0, help="diagnostics: 1 = one-shot, 2 = two-shot peer-to-peer all-reduce (0 = library's choice)")
    ap.add_argument("--emulate-tp", type=int, default=0,
                    help="diagnostics, one GPU: run ONE rank's shard of a tensor-parallel degree-N step (no all-reduce); "
                         "the JSON line is marked and is not a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="progress markers on stderr")
    ap.add_argument("--no-early-kv", action="store_true", help="do not let the decode kernel stream older pages ahead of its dependency")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch between our kernels")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode (20 ms period), started before the warm-up so that the process is up when
    the timed region begins; only samples whose timestamp falls inside [t_start, t_end] (the timed
    regions) are used."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.proc, self.path, self.dev = None, None, device_index
        if shutil.which("nvidia-smi"):
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(device_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self, t_start: float, t_end: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        inside, everything = [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rec = (float(f[1]), float(f[2]), [n for n, v in zip(names, f[5:9]) if v.lower() == "active"])
            except ValueError:
                continue
            everything.append(rec)
            if t_start - 0.02 <= ts <= t_end + 0.02:
                inside.append(rec)
        os.unlink(self.path)
        used, where = (inside, "timed regions") if inside else (everything[-10:], "warm-up + timed regions (no sample fell inside the timed window)")
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(r[0] for r in used)
        reasons = sorted({n for r in used for n in r[2]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in used), "reasons": reasons,
                "samples": len(used), "window": where}


# ------------------------------------------------------------------------------------------------
# state
# ------------------------------------------------------------------------------------------------
class State:
    """Per-rank synthetic model slice + KV caches + step metadata, all seeded."""

    def __init__(self, torch, rank: int, world: int, layers: int, dev):
        self.torch, self.rank, self.world, self.layers, self.dev = torch, rank, world, layers, dev
        from atoma_infer_b200.tp import HeadShard
        self.shard = HeadShard(rank, world, H, HK, D)   # llama_nccl.rs:164-171
        self.h, self.hk = self.shard.heads, self.shard.kv_heads
        self.nb = BATCH * MB + 64
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        bf = torch.bfloat16
        nq, nkv = self.h * D, self.hk * D
        self.wqkv = [(torch.randn((nq + 2 * nkv, HIDDEN), generator=g, device=dev, dtype=torch.float32)
                      * HIDDEN ** -0.5).to(bf) for _ in range(layers)]
        self.wo = [(torch.randn((HIDDEN, nq), generator=g, device=dev, dtype=torch.float32)
                    * (H * D) ** -0.5).to(bf) for _ in range(layers)]
        self.kv = []
        for _ in range(layers):
            c = torch.empty((2, self.nb, PAGE, self.hk, D), device=dev, dtype=bf)
            c.normal_(generator=g)
            self.kv.append(c)
        # metadata is identical on every rank (same scheduler output broadcast, model_executor.rs:531-542)
        gm = torch.Generator(device="cpu").manual_seed(99)
        perm = torch.randperm(self.nb, generator=gm)[: BATCH * MB].to(torch.int32).reshape(BATCH, MB)
        lens = torch.full((BATCH,), CTX, dtype=torch.int32)
        last = (lens.long() - 1)
        slots = perm[torch.arange(BATCH), (last // PAGE)].long() * PAGE + last % PAGE  # worker.rs:392-401
        x = (torch.randn((BATCH, HIDDEN), generator=gm, dtype=torch.float32)).to(bf)
        # rotary tables as the reference builds them (models/src/llama.rs:145-199), positions = len - 1
        inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
        ang = torch.arange(CTX + PAGE, dtype=torch.float32)[:, None] * inv[None, :]
        self.cos, self.sin = ang.cos().to(bf).to(dev), ang.sin().to(bf).to(dev)
        # The step's inputs.  Hidden states: a pinned host tensor and its device image.  Metadata
        # (block tables, sequence lengths, slot mapping, positions): built from the scheduler's view of
        # the 32 sequences by the C++ restatement of ModelWorker::prepare_input_tensors into ONE pinned
        # buffer and shipped with ONE copy (atoma_infer_b200/worker.py); the device
        # tensors below are views of that buffer's device image.
        from atoma_infer_b200.worker import InputStager, SequenceInput
        import numpy as np
        self.stager = InputStager(dev, PAGE)
        rngt = np.random.default_rng(5)
        self.seqs = [SequenceInput(rngt.integers(0, 128256, size=CTX).astype(np.uint32), 0, 1,
                                   perm[b].numpy().astype(np.uint32), False) for b in range(BATCH)]
        self.marshalled = self.stager.marshal(self.seqs)
        mi = self.stager.prepare_marshalled(self.marshalled)
        md = mi.attention_metadata
        self.meta_bytes = int(mi.layout.total_bytes)
        self.host = {"x": x.pin_memory()}
        self.devt = {"x": x.to(dev), "bt": md.decoding_metadata.block_tables,
                     "lens": md.decoding_metadata.sequence_lengths, "slots": md.slot_mapping,
                     "pos": mi.input_positions[0]}
        torch.cuda.synchronize()
        assert torch.equal(self.devt["slots"].cpu(), slots) and torch.equal(self.devt["pos"].cpu(), last)
        assert torch.equal(self.devt["bt"].cpu(), perm) and torch.equal(self.devt["lens"].cpu(), lens)
        self.out = torch.empty((BATCH, HIDDEN), device=dev, dtype=bf)
        self.out_host = torch.empty((BATCH, HIDDEN), dtype=bf).pin_memory()
        self.scale = D ** -0.5

    def h2d_bytes(self):
        return self.host["x"].numel() * self.host["x"].element_size() + self.meta_bytes

    def d2h_bytes(self):
        return self.out_host.numel() * self.out_host.element_size()

    def upload(self):
        """Per step: hidden states (one copy) + the step's metadata, rebuilt on the host and shipped
        with one copy."""
        self.devt["x"].copy_(self.host["x"], non_blocking=True)
        lay = self.stager.refresh(self.marshalled)
        assert lay.total_bytes == self.meta_bytes   # same shapes -> the views taken at start-up stay valid

    def download(self):
        self.out_host.copy_(self.out, non_blocking=True)


def make_step(st: State, impl: str, gemm: str, dist, comm=None, p2p=None):
    """Returns step() -> None that runs all layers on the current stream and leaves the result in
    st.out.  The all-reduce after o_proj (multi_gpu.rs:141-179) is, in this order of preference:
    `p2p` — o_proj writes into the rank's peer-mapped slot and atoma_allreduce_sum_p2p sums the slots
    of all ranks into a local buffer; `comm` (direct NCCL communicator) — ncclAllReduce in place on
    ping-pong buffers; both on the step's stream (capturable); else torch.distributed."""
    torch = st.torch
    nq, nkv = st.h * D, st.hk * D
    d = st.devt

    if impl == "ours":
        from atoma_infer_b200 import ops

        def mm(a, w, out=None):
            if gemm == "atoma":
                return ops.gemm_bf16_tn(a, w, out)
            return torch.matmul(a, w.t(), out=out)

        def attn(q, kc, vc, layer):
            return ops.flash_attn_kv_cache_full(q, kc, vc, None, st.scale, d["bt"], d["lens"], True)

        def rope_and_cache_write(q, k, v, kc, vc):
            ops.rope_reshape_and_cache(q, k, v, kc, vc, d["pos"], st.cos, st.sin, d["slots"])
            return q, k
    else:
        from atoma_infer_b200 import ops  # host planner only: derives the reference's run_mha arguments
        from oracle.ref_lib import RefLib
        ref = RefLib()
        sms = torch.cuda.get_device_properties(st.dev).multi_processor_count
        stream0 = 0  # legacy default stream, like the reference (flash_api.cu:157)

        def mm(a, w, out=None):
            return torch.matmul(a, w.t(), out=out)

        def attn(q, kc, vc, layer):
            plan = ops.flash_attn_kv_cache_plan(q, kc, vc, st.scale, block_table=d["bt"], seqlens_k=d["lens"],
                                                window_size_right=0, num_sms=sms)
            return ref.execute(plan, q, sync=False)

        # candle runs RoPE as 2 index_select + 2 rope kernels + 4 transpose/contiguous copies per layer
        # (models/src/llama.rs:238-304).  Same kernel count here, in the model dtype:
        # rope(x) = x * [cos, cos] + roll(x, d/2) * [-sin, sin]  ->  index_select x2 + (roll, mul, addcmul) x2.
        cos_full = torch.cat([st.cos, st.cos], dim=-1)
        sin_signed = torch.cat([-st.sin, st.sin], dim=-1)

        def rope_torch(x, cos, sin):
            return torch.addcmul(x * cos, torch.roll(x, D // 2, dims=-1), sin)

        def rope_and_cache_write(q, k, v, kc, vc):
            cos = cos_full.index_select(0, d["pos"])[:, None, :]
            sin = sin_signed.index_select(0, d["pos"])[:, None, :]
            q, k = rope_torch(q, cos, sin), rope_torch(k, cos, sin)
            ref.lib.reshape_and_cache_flash(k.data_ptr(), v.data_ptr(), kc.data_ptr(), vc.data_ptr(),
                                            d["slots"].data_ptr(), kc.stride(0), BATCH, st.hk, D, PAGE,
                                            k.stride(0), v.stride(0), 1, stream0)
            return q, k

    def layer(l, x, out=None):
        """One layer of the path on this rank: x [B, hidden] -> partial (pre-all-reduce) o_proj output."""
        qkv = mm(x, st.wqkv[l])                                   # [B, (h + 2 hk) D]
        q = qkv[:, :nq].unflatten(1, (st.h, D))                   # [B, h, D] (strided view)
        k = qkv[:, nq:nq + nkv].unflatten(1, (st.hk, D))          # [B, hk, D]
        v = qkv[:, nq + nkv:].unflatten(1, (st.hk, D))
        kc, vc = st.kv[l][0], st.kv[l][1]
        q, k = rope_and_cache_write(q, k, v, kc, vc)
        o = attn(q.unsqueeze(1), kc, vc, l)                       # [B, 1, h, D]
        return mm(o.reshape(BATCH, nq), st.wo[l], out)            # [B, hidden]

    def step():
        x = d["x"]
        for l in range(st.layers):
            if p2p is not None:
                layer(l, x, out=st.tp_slots[l % 2])
                x = p2p.all_reduce(l % 2, st.tp_out[l % 2])
            elif comm is not None:
                x = layer(l, x, out=st.tp_bufs[l % 2])
                comm.all_reduce_(x)                                   # multi_gpu.rs:141-179, on this stream
            else:
                x = layer(l, x)
                if st.world > 1:
                    dist.all_reduce(x)                                # multi_gpu.rs:141-179
        st.out.copy_(x)

    step.layer = layer
    return step


# ------------------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle ("port") on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_baseline(seconds_budget: float = 20.0):
    """One host thread: numpy's BLAS pool is pinned to 1 for the duration, so that `cores: 1` is what
    actually ran (the oracle's attention loops are single-threaded anyway)."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            return _cpu_baseline(seconds_budget, "1 (BLAS pool pinned to one thread)")
    except ImportError:
        import os as _os
        return _cpu_baseline(seconds_budget, f"numpy BLAS default pool (up to {_os.cpu_count()} threads) + 1 for the oracle's loops")


def _cpu_baseline(seconds_budget: float, threads_note: str):
    import numpy as np
    from oracle import attention_oracle as orc
    rng = np.random.default_rng(0)
    b, mb = 2, MB
    nb = b * mb
    f = lambda *s: orc.round_to_io(rng.standard_normal(s, dtype=np.float32), True)
    kc, vc = f(nb, PAGE, HK, D), f(nb, PAGE, HK, D)
    x = f(b, HIDDEN)
    wqkv, wo = f((H + 2 * HK) * D, HIDDEN) * HIDDEN ** -0.5, f(HIDDEN, H * D) * (H * D) ** -0.5
    bt = rng.permutation(nb).astype(np.int32).reshape(b, mb)
    lens = np.full(b, CTX)
    slots = bt[np.arange(b), (lens - 1) // PAGE].astype(np.int64) * PAGE + (lens - 1) % PAGE
    kcb, vcb = orc.from_f32(kc, True).ravel(), orc.from_f32(vc, True).ravel()   # flat storage bits
    cos, sin = orc.rope_tables(D, CTX + PAGE, is_bf16=True)
    t0 = time.perf_counter()
    n = 0
    while True:  # one "layer" of the path for b sequences
        qkv = orc.round_to_io((x @ wqkv.T).astype(np.float32), True)
        nq, nkv = H * D, HK * D
        q3 = orc.round_to_io(orc.rope(qkv[:, :nq].reshape(b, H, D), lens - 1, cos, sin).astype(np.float32), True)
        k3 = orc.round_to_io(orc.rope(qkv[:, nq:nq + nkv].reshape(b, HK, D), lens - 1, cos, sin).astype(np.float32), True)
        q = q3.reshape(b, 1, H, D)
        orc.reshape_and_cache_flash(orc.from_f32(k3, True).ravel(), orc.from_f32(qkv, True).ravel()[nq + nkv:],
                                    kcb, vcb, slots, PAGE * nkv, b, HK, D, PAGE, nkv, qkv.shape[1])
        o, _ = orc.attention_paged_decode(q, orc.to_f32(kcb, True).reshape(kc.shape), orc.to_f32(vcb, True).reshape(vc.shape),
                                          bt, lens, D ** -0.5, True, precision="reference")
        x2 = orc.round_to_io((o.reshape(b, nq) @ wo.T).astype(np.float32), True)
        n += 1
        if time.perf_counter() - t0 > seconds_budget or n >= 8:
            break
    dt = (time.perf_counter() - t0) / n            # seconds per layer for b sequences
    tok_s = b / (dt * L)
    del x2
    return {"value": tok_s, "unit": "tokens/s", "cores": 1 if threads_note.startswith("1 ") else (os.cpu_count() or 1),
            "kind": "port",
            "sample": f"numpy oracle (oracle/attention_oracle.py), {n} layer pass(es) of the same path for {b} of "
                      f"the 32 sequences at ctx {CTX}; scaled to 32 layers; host threads: {threads_note}"}


def ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the newest
    committed `ncu --set full` capture under profiles/ (taken with scripts/gpu_profile.sh on the same
    config-2 inputs); None when no capture is present.  Evidence, not a timing."""
    import csv
    best = None
    for f in sorted((ROOT / "profiles").glob("r*_ncu_full_raw.csv")):
        try:
            rows = list(csv.reader(open(f)))
            h = rows[0]
            kn, rd, wr = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = [float(r[rd]) * unit[rows[1][rd]] + float(r[wr]) * unit[rows[1][wr]]
                    for r in rows[2:] if kernel_substr in r[kn]]
            if vals:
                best = {"bytes": sum(vals) / len(vals), "file": f"profiles/{f.name}"}
        except Exception:
            continue
    return best


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference" and rank != 0:
        return  # rank 0 alone runs the reference arm
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner when
    # NCCL_DEBUG is set in the environment): keep the real stdout aside and point fd 1 at stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def mark(msg):
        if args.verbose:
            print(f"[bench r{rank}] {time.strftime('%H:%M:%S')} {msg}", file=sys.stderr, flush=True)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU or PyTorch fallback for the product path)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    tp = world if args.impl == "ours" else 1
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if HK % tp != 0:
        raise SystemExit(f"tensor parallel degree {tp} must divide the {HK} kv heads")

    from atoma_infer_b200 import _lib, ops
    _lib.lib()  # fail loudly if libflashattention.so is missing
    if args.impl == "reference":
        from oracle.ref_lib import RefLib
        if not RefLib.available():
            raise SystemExit("oracle/_ref/libatoma_ref.so is missing: run __graft_entry__.build() where /root/reference exists")

    gemm = args.gemm
    if gemm == "auto":
        gemm = "cublas"
        if args.impl == "ours" and os.environ.get("ATOMA_BENCH_GEMM", "cublas") == "atoma":
            # use the hand-written GEMM once it is built (the placeholder reports "unsupported")
            try:
                a = torch.zeros((BATCH, 64), device=dev, dtype=torch.bfloat16)
                w = torch.zeros((128, 64), device=dev, dtype=torch.bfloat16)
                ops.gemm_bf16_tn(a, w)
                gemm = "atoma"
            except Exception:
                gemm = "cublas"

    if args.impl == "ours":
        # weights are immutable during a decode step: the GEMM may stream them ahead of its dependency
        # ... and a decode step writes only the slot of the token being decoded (each sequence's final
        # page): the decode kernel may stream the older pages while the cache write is still running
        ops.set_launch_mode(pdl=not args.no_pdl, const_weights=not args.no_pdl,
                            stable_cache=not (args.no_pdl or args.no_early_kv))
    mark("init done, building state")
    if args.emulate_tp:
        if world != 1 or args.impl != "ours":
            raise SystemExit("--emulate-tp is a single-process diagnostic of our arm")
        tp = args.emulate_tp
    st = State(torch, rank if args.impl == "ours" else 0, tp, args.layers, dev)
    if args.emulate_tp:
        st.world = 1   # no collective in the step
    mark("state built")
    comm = p2p = None
    tp_comm = args.tp_comm
    if dist is not None and tp_comm == "p2p":
        from atoma_infer_b200.p2p import P2PComm
        ok = torch.ones(1, device=dev)
        try:
            if args.layers % 2:
                raise RuntimeError("slots alternate per layer: the layer count must be even")
            p2p = P2PComm.from_torch_distributed(dev, BATCH * HIDDEN * 2)
        except Exception as e:
            print(f"[bench r{rank}] peer-to-peer all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
        if ok.item() == 0:
            if p2p is not None:
                p2p.destroy()
            p2p, tp_comm = None, "nccl"
        else:
            _lib.lib().atoma_debug_allreduce_algo(args.ar_algo)
            st.tp_slots = [p2p.slot(i, (BATCH, HIDDEN), torch.bfloat16) for i in range(2)]
            st.tp_out = [torch.empty((BATCH, HIDDEN), device=dev, dtype=torch.bfloat16) for _ in range(2)]
            mark("peer-to-peer all-reduce regions mapped")
    if dist is not None and tp_comm == "nccl":
        from atoma_infer_b200 import nccl
        comm = nccl.init_from_torch_distributed(dev)
        st.tp_bufs = [comm.symmetric_empty((BATCH, HIDDEN), torch.bfloat16, dev) for _ in range(2)]
        mark(f"direct NCCL communicator up (NCCL {nccl.version()}, {comm.symmetric_windows()} symmetric windows)")
    step = make_step(st, args.impl, gemm, dist, comm, p2p)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- one eager step: allocates library workspaces, warms cuBLAS; then capture ----
    if dist is not None:
        # library initialisation (cuBLAS handles, kernel loading, workspaces) before the first collective,
        # so that the ranks reach it together
        step.layer(0, st.devt["x"])
        torch.cuda.synchronize()
        dist.barrier()
    ops.reset_launch_count()
    step()
    torch.cuda.synchronize()
    mark("first eager step done")
    launches_per_step = ops.launch_count() if args.impl == "ours" else 0
    # torch.distributed's NCCL collectives are launched eagerly (not graph-captured); the direct
    # communicator's are captured with the rest of the step
    use_graph = args.impl == "ours" and not args.no_graph and (world == 1 or comm is not None or p2p is not None)
    graph = None
    if use_graph:
        try:
            # The library keeps one workspace per (device, stream) and refuses to allocate during
            # capture: run one eager step on the capture stream first.
            side = torch.cuda.Stream()
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                step()
            stream.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover - depends on the box
            print(f"[bench] CUDA graph capture failed ({e}); launching eagerly", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    run = graph.replay if graph is not None else step
    layer_graphs = None
    if args.impl == "ours" and world > 1 and graph is None and not args.no_graph:
        # Tensor parallel: torch's NCCL collectives are launched eagerly (capturing them hangs with this
        # torch/NCCL pair), everything between two all-reduces is one CUDA graph per layer.
        try:
            bufs = [torch.empty_like(st.out), torch.empty_like(st.out)]
            side = torch.cuda.Stream()
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                for l in range(st.layers):
                    step.layer(l, bufs[l % 2], out=bufs[(l + 1) % 2])
            stream.wait_stream(side)
            torch.cuda.synchronize()
            layer_graphs = []
            for l in range(st.layers):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    step.layer(l, bufs[l % 2], out=bufs[(l + 1) % 2])
                layer_graphs.append(g)
            torch.cuda.synchronize()

            def run():
                bufs[0].copy_(st.devt["x"])
                for l, g in enumerate(layer_graphs):
                    g.replay()
                    dist.all_reduce(bufs[(l + 1) % 2])
                st.out.copy_(bufs[st.layers % 2])
        except Exception as e:  # pragma: no cover
            print(f"[bench] per-layer graph capture failed ({e}); launching eagerly", file=sys.stderr)
            layer_graphs = None
            run = step
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # Untimed: W warm-up steps, then a fixed number of extra steps (the same on every rank: the step
    # holds collectives) that keep the GPU under the step's load while nvidia-smi comes up.
    for _ in range(max(args.warmup, 3) + (100 if args.impl == "ours" else 20)):
        run()
    barrier()
    mark("warm-up done")

    # ---- timed region 1: device-resident inputs, K steps, CUDA events ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_clock0 = time.time()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_dev = e0.elapsed_time(e1)
    mark("timed region 1 done")

    # ---- timed region 2 (e2e): host buffers -> H2D -> step -> D2H, every step ----
    for _ in range(2):
        st.upload(); run(); st.download()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        st.upload()
        run()
        st.download()
        stream.synchronize()  # the step's result is on the host before the next step starts
    e3.record()
    barrier()
    t_e2e_wall = time.perf_counter() - t0
    ms_e2e = max(e2.elapsed_time(e3), t_e2e_wall * 1e3)  # host wall-clock includes the syncs
    clocks = sampler.stop(t_clock0, time.time()) if sampler else None
    mark("timed region 2 done")

    # ---- dominant kernel alone: paged decode attention, one launch per layer (distinct caches) ----
    nq = st.h * D
    qs = torch.randn((BATCH, 1, st.h, D), device=dev, dtype=torch.float32).to(torch.bfloat16)
    if args.impl == "ours":
        def attn_only(l):
            return ops.flash_attn_kv_cache_full(qs, st.kv[l][0], st.kv[l][1], None, st.scale, st.devt["bt"],
                                                st.devt["lens"], True)
    else:
        from oracle.ref_lib import RefLib
        ref = RefLib()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count

        def attn_only(l):
            plan = ops.flash_attn_kv_cache_plan(qs, st.kv[l][0], st.kv[l][1], st.scale, block_table=st.devt["bt"],
                                                seqlens_k=st.devt["lens"], window_size_right=0, num_sms=sms)
            return ref.execute(plan, qs, sync=False)
    def attn_pass():
        for l in range(st.layers):
            attn_only(l)

    attn_pass()
    torch.cuda.synchronize()
    attn_run = attn_pass
    if args.impl == "ours" and not args.no_graph:
        # same launch mechanism as the timed step: replay a captured pass (no host launch overhead)
        try:
            side2 = torch.cuda.Stream()
            side2.wait_stream(stream)
            with torch.cuda.stream(side2):
                attn_pass()
            stream.wait_stream(side2)
            torch.cuda.synchronize()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=side2):
                attn_pass()
            attn_run = g2.replay
        except Exception as e:  # pragma: no cover
            print(f"[bench] attention-only graph capture failed ({e})", file=sys.stderr)
            torch.cuda.synchronize()
    attn_run()
    torch.cuda.synchronize()
    reps = 3
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        attn_run()
    k1.record()
    torch.cuda.synchronize()
    us_kernel = k0.elapsed_time(k1) * 1e3 / (reps * st.layers)

    # ---- max over ranks ----
    times = torch.tensor([ms_dev, ms_e2e, us_kernel], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, us_kernel = [float(x) for x in times.tolist()]

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
        # ALGORITHMIC bytes of one decode-attention launch: K and V of every cached token
        # once + q and o + the block table, per rank.
        kv_bytes = 2 * BATCH * CTX * st.hk * D * 2
        alg_bytes = kv_bytes + 2 * BATCH * st.h * D * 2 + BATCH * MB * 4
        achieved = alg_bytes / (us_kernel * 1e-6) / 1e9
        ms_step = ms_dev / args.steps
        ms_step_e2e = ms_e2e / args.steps
        line = {
            "metric": METRIC,
            "value": BATCH / (ms_step * 1e-3),
            "unit": "tokens/s",
            "n_gpus": world if args.impl == "ours" else 1,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "bf16",
            "data": "synthetic (seeded N(0,1) activations / KV caches, random-init weights, random block tables)",
            "config": {
                "workload": f"BASELINE.json configs[1]: Llama-3.1-8B bf16 batch=32 ctx=4096 decode-only, block_size=16; "
                            f"attention path of one decode step ({st.layers} layers x [fused qkv GEMM, RoPE + reshape_and_cache_flash, "
                            f"paged decode attention, o_proj GEMM]); MLP/norms/lm_head not in SURVEY §8 scope and not timed",
                "batch": BATCH, "ctx": CTX, "layers": st.layers, "heads": H, "kv_heads": HK, "head_dim": D,
                "block_size": PAGE, "parallelism": f"tp{tp}",
                "l2": "inputs larger than L2: 541 MB of KV per layer, 17.3 GB per step, distinct per layer",
                "launch": ("cuda graph replay" + (" (NCCL all-reduces captured in the graph)" if comm is not None else
                                                  " (peer-to-peer all-reduce kernels captured in the graph)" if p2p is not None else "")
                           if graph is not None else
                           "one cuda graph per layer + eager NCCL all-reduce" if layer_graphs is not None else
                           "eager, one host thread") +
                          (", programmatic dependent launch" if args.impl == "ours" and not args.no_pdl else "") +
                          (", early KV streaming" if args.impl == "ours" and not (args.no_pdl or args.no_early_kv) else ""),
                "all_reduce": (None if tp == 1 else
                               "atoma_allreduce_sum_p2p (one-shot kernel over NVLink peer memory)" if p2p is not None else
                               "ncclAllReduce (direct binding)" if comm is not None else "torch.distributed.all_reduce (NCCL)"),
                "gemm": "atoma tcgen05 kernel" if gemm == "atoma" else "cuBLAS via torch.matmul (library GEMM, as in the reference)",
                "wall_ms_per_step": t_wall * 1e3 / args.steps,
            },
            "e2e": {"value": BATCH / (ms_step_e2e * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": st.h2d_bytes(), "d2h_bytes_per_step": st.d2h_bytes(),
                    "ms_per_step": ms_step_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "decode_paged_kernel" if args.impl == "ours" else "flash_fwd_splitkv_kernel",
                         "achieved": achieved, "peak": peak_hbm, "unit": "GB/s", "frac": achieved / peak_hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "us_per_launch": us_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                         "frac_of_8TBps": achieved / 8000.0, "traffic": None},
        }
        if args.impl == "ours" and tp == 1 and st.layers == L:
            t = ncu_traffic("decode_paged_kernel")
            if t:
                line["roofline"]["traffic"] = t["bytes"]
                line["roofline"]["traffic_source"] = t["file"] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, same inputs)"
        if args.emulate_tp:
            line["diagnostic"] = f"one rank's shard of a tp{tp} step on one GPU, no all-reduce: NOT a bench value"
        if args.impl == "reference":
            line["impl"] = "reference"
            line["value"] = line["e2e"]["value"]   # the reference's number is host-timed by definition
            line["ms_per_step"] = ms_step_e2e
            line["e2e"] = {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
            line["config"]["device_ms_per_step"] = ms_step
            line["cpu_baseline"] = {
                "value": line["value"], "unit": "tokens/s", "cores": 1, "kind": "reference",
                "sample": "the reference's own CUDA kernels (oracle/_ref, built from /root/reference/csrc/kernels for sm_100) "
                          "driven from ONE host thread with host wall-clock timing, as the reference times itself "
                          "(model_executor.rs:331-347); the reference has no CPU implementation of this path"}
        elif not args.no_cpu_baseline and world == 1 and not args.emulate_tp:
            line["cpu_baseline"] = cpu_baseline()   # rank 0, N = 1 only: the CPU port does not change with N
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
        run = graph = None   # the graph holds the captured collectives: drop it before the communicator
        if p2p is not None:
            p2p.check()
            p2p.destroy()
        if comm is not None:
            comm.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
