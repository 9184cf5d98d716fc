#!/usr/bin/env python
"""Benchmark of the decode hot path at BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4]

Workload (default, config 2 = BASELINE.json configs[1]): one DECODE STEP of Llama-3.1-8B, bf16,
batch 32, every sequence at context 4096, block_size 16 — the whole model, as `Llama::forward` runs it
(models/src/llama.rs:456-478) plus greedy sampling (backends/vllm/src/model_executor.rs:227-262):

    wte gather; 32 x [ RMSNorm; q/k/v projections; RoPE; reshape_and_cache_flash; paged decode attention;
    o_proj (+all-reduce); residual; RMSNorm; gate/up projections; SiLU*up; down_proj (+all-reduce);
    residual ]; final RMSNorm; lm_head; sampling

with a distinct 541 MB KV cache per layer and 15 GB of weights: 32.2 GB of HBM traffic per step, far
larger than L2 (no flush needed).  Metric: decode tokens/s = batch / step time, whole job.

`--impl ours`: the product — `atoma_llama_decode_step` (atoma_infer_b200.llama.DecodeStep): 6 hand-written
sm_100a kernels per layer (fused qkv+RoPE+cache-write GEMM, paged decode attention, o_proj GEMM, fused
add+RMSNorm x2, gate_up+SwiGLU and down_proj chained in one kernel), no library kernel in the step, replayed from
one CUDA graph; the step's inputs come from `InputStager` (one host-to-device copy per step).

`--impl reference`: the REFERENCE's own kernels (oracle/_ref, compiled from /root/reference/csrc/kernels
for sm_100: flash_fwd_splitkv_kernel via run_mha with the argument tuple its Rust wrapper derives,
reshape_and_cache_flash) + cuBLAS (torch.matmul; candle -> cuBLAS in the reference) for the seven
projections of a layer and lm_head + element-wise torch kernels standing in for candle's (rms_norm, silu,
mul, add; RoPE as candle runs it: 2 index_select + 2 rope + 4 transpose/contiguous copies), launched
eagerly from ONE host thread on the legacy default stream, logits copied to the host per sequence and
sampled there, host wall-clock timing — how the reference runs and times itself
(model_executor.rs:331-347).  The reference has no CPU implementation of this path (cpu_fwd bails,
csrc/src/lib.rs:1862-1872).  This arm loads nothing of this repo's product (no libflashattention.so).

N > 1: tensor parallel exactly like the reference (models/src/llama_nccl.rs:153-202, multi_gpu.rs:8-57):
heads, kv heads, caches and MLP columns sharded, row-parallel o_proj / down_proj all-reduced; embedding,
norms and lm_head replicated.  Ours: the all-reduce is fused into the GEMM epilogue (push over NVLink peer
memory) and the following add+RMSNorm kernel (receive side).  Reference arm: N ranks, ncclAllReduce through
torch.distributed, eager.  Total work is fixed -> "scaling": "strong".

Other configs (`--config`, not the headline): 1 = Llama-3.2-1B single-stream decode (launch-overhead bound);
4 = Llama-3.1-70B TP=8 batch 16 (needs --gpus 8); 3 = the continuous-batching mix through
FlashAttention.forward (8192-token prefill + decode tokens in one flattened batch, attention operator only).
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

PAGE = 16
METRIC = "decode tokens/sec Llama-3.1-8B bf16 b32 ctx4k"
CONFIGS = {
    1: dict(model="1b", batch=1, ctx=544, tp=1, metric="decode tokens/sec Llama-3.2-1B bf16 b1 single-prompt greedy decode",
            name="BASELINE.json configs[0]: Llama-3.2-1B bf16 single-prompt greedy decode, 1 device, block_size=16 "
                 "(one decode step at context 544 = 32 prompt + 512 generated tokens)"),
    2: dict(model="8b", batch=32, ctx=4096, tp=None, metric=METRIC,
            name="BASELINE.json configs[1]: Llama-3.1-8B bf16 batch=32 ctx=4096 decode-only, block_size=16"),
    4: dict(model="70b", batch=16, ctx=4096, tp=8, metric="decode tokens/sec Llama-3.1-70B bf16 TP=8 b16 ctx4k",
            name="BASELINE.json configs[3]: Llama-3.1-70B bf16 TP=8 over NVLink, batch=16 ctx=4096"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs: 2 (default) is the headline; 1, 3, 4 are the other measured configurations")
    ap.add_argument("--layers", type=int, default=0, help="debug only; the reported config needs the model's own layer count")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch between our kernels")
    ap.add_argument("--emulate-tp", type=int, default=0,
                    help="diagnostics, one GPU: run ONE rank's shard of a tensor-parallel degree-N step (no all-reduce); "
                         "the JSON line is marked and is not a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tp-parity", action="store_true", help="N > 1: skip the tp=1 recomputation on rank 0")
    ap.add_argument("--verbose", action="store_true", help="progress markers on stderr")
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
# synthetic model + step inputs (seeded; identical for both arms and, sharded, for every rank)
# ------------------------------------------------------------------------------------------------
def model_config(name: str):
    from atoma_infer_b200.llama import LlamaConfig
    return {"8b": LlamaConfig.llama31_8b, "1b": LlamaConfig.llama32_1b, "70b": LlamaConfig.llama31_70b}[name]()


class Workload:
    """Weights (this rank's shard of the seeded full model), KV caches (this rank's kv heads of the seeded
    full caches) and the scheduler's view of the batch.  Pure torch + the host-side sharding helpers: both
    arms build it the same way."""

    def __init__(self, torch, cfg, batch, ctx, rank, world, layers, dev):
        from atoma_infer_b200.llama import LlamaShard
        self.torch, self.cfg, self.batch, self.ctx, self.rank, self.world, self.dev = torch, cfg, batch, ctx, rank, world, dev
        self.mb = (ctx + PAGE - 1) // PAGE
        self.nb = batch * self.mb + 64
        self.shard = LlamaShard.random_init(cfg, rank, world, dtype=torch.bfloat16, device=dev, seed=1234,
                                            num_layers=layers, max_pos=ctx + PAGE)
        hk, D = cfg.num_key_value_heads, cfg.head_dim
        hs = self.shard.shard.kv_range
        self.kv = []
        g = torch.Generator(device=dev)
        for l in range(layers):
            # the full [2, NB, page, Hkv, D] cache of layer l from its own seed, of which this rank keeps its heads
            # (worker.rs:584-591); generated head by head so that the transient stays small
            c = torch.empty((2, self.nb, PAGE, hs.stop - hs.start, D), device=dev, dtype=torch.bfloat16)
            for i, h in enumerate(range(hs.start, hs.stop)):
                g.manual_seed(1000003 * (l + 1) + h)
                c[:, :, :, i].normal_(generator=g)
            self.kv.append(c)
        gm = torch.Generator(device="cpu").manual_seed(99)
        self.perm = torch.randperm(self.nb, generator=gm)[: batch * self.mb].to(torch.int32).reshape(batch, self.mb)
        self.lens = torch.full((batch,), ctx, dtype=torch.int32)
        self.pos = self.lens.long() - 1
        self.slots = self.perm[torch.arange(batch), (self.pos // PAGE)].long() * PAGE + self.pos % PAGE   # worker.rs:392-401
        self.tokens = torch.randint(0, cfg.vocab_size, (batch,), generator=gm, dtype=torch.int32)

    def kv_bytes_per_layer(self):
        c = self.kv[0]
        return 2 * self.batch * self.ctx * c.shape[3] * c.shape[4] * 2


# ------------------------------------------------------------------------------------------------
# our arm: the product's step executor
# ------------------------------------------------------------------------------------------------
class OursArm:
    def __init__(self, torch, wl: Workload, p2p, no_comm: bool):
        from atoma_infer_b200.llama import DecodeStep
        from atoma_infer_b200.worker import InputStager, SequenceInput
        import numpy as np
        self.torch, self.wl = torch, wl
        # The step's inputs (tokens, positions, slot mapping, block tables, sequence lengths): built from the
        # scheduler's view of the sequences by the C++ restatement of ModelWorker::prepare_input_tensors into
        # ONE pinned buffer and shipped with ONE copy (atoma_infer_b200/worker.py, row f2); the device tensors
        # are views of that buffer's device image.
        self.stager = InputStager(wl.dev, PAGE)
        rng = np.random.default_rng(5)
        seqs = []
        for b in range(wl.batch):
            toks = rng.integers(0, wl.cfg.vocab_size, size=wl.ctx).astype(np.uint32)
            toks[-1] = int(wl.tokens[b])
            seqs.append(SequenceInput(toks, 0, 1, wl.perm[b].numpy().astype(np.uint32), False))
        self.marshalled = self.stager.marshal(seqs)
        mi = self.stager.prepare_marshalled(self.marshalled)
        md = mi.attention_metadata
        self.meta_bytes = int(mi.layout.total_bytes)
        self.tokens_dev = mi.input_tokens_tensor[0]
        self.bt, self.lens = md.decoding_metadata.block_tables, md.decoding_metadata.sequence_lengths
        self.slots, self.pos = md.slot_mapping, mi.input_positions[0]
        torch.cuda.synchronize()
        assert torch.equal(self.slots.cpu(), wl.slots) and torch.equal(self.pos.cpu(), wl.pos)
        assert torch.equal(self.bt.cpu(), wl.perm) and torch.equal(self.lens.cpu(), wl.lens)
        assert torch.equal(self.tokens_dev.cpu(), wl.tokens)
        self.step = DecodeStep(wl.shard, wl.kv, max_batch=wl.batch, p2p=p2p, no_comm=no_comm)
        self.step.bind(self.tokens_dev, self.pos, self.slots, self.bt, self.lens)
        self.out_host = torch.empty((wl.batch,), dtype=torch.int32).pin_memory()
        self.val_host = torch.empty((wl.batch,), dtype=torch.float32).pin_memory()

    def h2d_bytes(self):
        return self.meta_bytes

    def d2h_bytes(self):
        return self.out_host.numel() * 4 + self.val_host.numel() * 4

    def upload(self):
        lay = self.stager.refresh(self.marshalled)
        assert lay.total_bytes == self.meta_bytes   # same shapes -> the views taken at start-up stay valid

    def download(self):
        self.out_host.copy_(self.step.next_tokens[: self.wl.batch], non_blocking=True)
        self.val_host.copy_(self.step.next_logits[: self.wl.batch], non_blocking=True)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's kernels + cuBLAS + element-wise torch, eager, one host thread
# ------------------------------------------------------------------------------------------------
class ReferenceArm:
    def __init__(self, torch, wl: Workload, dist):
        from oracle.ref_plan import RefKernels, kv_cache_plan
        self.torch, self.wl, self.dist = torch, wl, dist
        self.ref = RefKernels()
        dev, cfg = wl.dev, wl.cfg
        sh = wl.shard.shard
        self.h, self.hk, self.D = sh.heads, sh.kv_heads, cfg.head_dim
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        q_shape = [wl.batch, 1, self.h, self.D]
        kc = wl.kv[0][0]
        self.plan = kv_cache_plan(q_shape, [self.h * self.D, self.h * self.D, self.D, 1], list(kc.shape), list(kc.stride()),
                                  list(wl.kv[0][1].stride()), [wl.batch, wl.mb], self.D ** -0.5, True, sms, causal=True)
        # host-side step inputs (pinned) and their device tensors: one upload per tensor per step, as
        # prepare_input_tensors does (worker.rs:405-450)
        self.host = {"tokens": wl.tokens.pin_memory(), "pos": wl.pos.pin_memory(), "slots": wl.slots.pin_memory(),
                     "lens": wl.lens.pin_memory(), "bt": wl.perm.contiguous().pin_memory()}
        self.d = {k: v.to(dev) for k, v in self.host.items()}
        self.cos_full = torch.cat([wl.shard.cos, wl.shard.cos], dim=-1)
        self.sin_signed = torch.cat([-wl.shard.sin, wl.shard.sin], dim=-1)
        self.next_tokens = [0] * wl.batch

    def h2d_bytes(self):
        return 0

    def d2h_bytes(self):
        return 0

    def upload(self):
        for k, v in self.host.items():
            self.d[k].copy_(v, non_blocking=True)

    def rope(self, x, cos, sin):
        # rope(x) = x * [cos, cos] + roll(x, D/2) * [-sin, sin]: 3 element-wise kernels per tensor + the two
        # index_selects = candle's kernel count for RoPE (llama.rs:238-304) in the model dtype
        return self.torch.addcmul(x * cos, self.torch.roll(x, self.D // 2, dims=-1), sin)

    def step(self):
        torch, wl, F = self.torch, self.wl, self.torch.nn.functional
        cfg, d, B = wl.cfg, self.d, wl.batch
        nq, nkv = self.h * self.D, self.hk * self.D
        eps = cfg.rms_norm_eps
        x = wl.shard.embed.index_select(0, d["tokens"])                       # llama.rs:470
        for l, ly in enumerate(wl.shard.layers):
            kc, vc = wl.kv[l][0], wl.kv[l][1]
            xn = F.rms_norm(x, (cfg.hidden_size,), ly["rms_attn"], eps)       # llama.rs:399
            wqkv = ly["wqkv"]
            q = torch.matmul(xn, wqkv[:nq].t())                               # three Linear calls (llama.rs:269-271)
            k = torch.matmul(xn, wqkv[nq:nq + nkv].t())
            v = torch.matmul(xn, wqkv[nq + nkv:].t())
            cos = self.cos_full.index_select(0, d["pos"])[:, None, :]
            sin = self.sin_signed.index_select(0, d["pos"])[:, None, :]
            q = self.rope(q.view(B, self.h, self.D), cos, sin)
            k = self.rope(k.view(B, self.hk, self.D), cos, sin)
            v = v.view(B, self.hk, self.D)
            self.ref.lib.reshape_and_cache_flash(k.data_ptr(), v.data_ptr(), kc.data_ptr(), vc.data_ptr(),
                                                 d["slots"].data_ptr(), kc.stride(0), B, self.hk, self.D, PAGE,
                                                 k.stride(0), v.stride(0), 1, 0)          # flash_attention.rs:361
            # what the Rust wrapper allocates per call (csrc/src/lib.rs:1726-1729, 1785-1800)
            o = torch.empty((B, 1, self.h, self.D), dtype=x.dtype, device=x.device)
            lse = torch.zeros(self.plan.lse_elems, dtype=torch.float32, device=x.device)
            lacc = oacc = None
            if self.plan.num_splits > 1:
                lacc = torch.zeros(self.plan.lseaccum_elems, dtype=torch.float32, device=x.device)
                oacc = torch.zeros(self.plan.oaccum_elems, dtype=torch.float32, device=x.device)
            self.ref.run_kv_cache(self.plan, q.data_ptr(), kc.data_ptr(), vc.data_ptr(), o.data_ptr(), lse.data_ptr(),
                                  d["bt"].data_ptr(), d["lens"].data_ptr(),
                                  lacc.data_ptr() if lacc is not None else None,
                                  oacc.data_ptr() if oacc is not None else None)     # flash_attention.rs:455
            a = torch.matmul(o.view(B, nq), ly["wo"].t())                     # llama.rs:311
            if self.dist is not None:
                self.dist.all_reduce(a)                                       # multi_gpu.rs:141-179
            x = a + x                                                         # llama.rs:401-404
            xn = F.rms_norm(x, (cfg.hidden_size,), ly["rms_mlp"], eps)
            wgu = ly["w_gate_up"]
            inter = wgu.shape[0] // 2
            act = F.silu(torch.matmul(xn, wgu[:inter].t())) * torch.matmul(xn, wgu[inter:].t())   # llama.rs:364
            m = torch.matmul(act, ly["w_down"].t())                           # llama.rs:365
            if self.dist is not None:
                self.dist.all_reduce(m)
            x = m + x
        xf = F.rms_norm(x, (cfg.hidden_size,), wl.shard.norm_f, eps)          # llama.rs:474
        logits = torch.matmul(xf, wl.shard.lm_head.t()).float()               # llama.rs:476-477
        # model_executor.rs:227-262: per sequence, the logits row goes to the host (to_vec1) and is sampled there
        for b in range(B):
            row = logits[b].cpu().numpy()
            self.next_tokens[b] = int(len(row) - 1 - row[::-1].argmax())      # max_by: the LAST maximum
        return logits


# ------------------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle ("port") on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, batch, ctx, seconds_budget: float = 20.0):
    """One host thread: numpy's BLAS pool is pinned to 1 for the duration, so that `cores: 1` is what
    actually ran (the oracle's attention loops are single-threaded anyway)."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            return _cpu_baseline(cfg, batch, ctx, seconds_budget, "1 (BLAS pool pinned to one thread)")
    except ImportError:
        return _cpu_baseline(cfg, batch, ctx, seconds_budget, f"numpy BLAS default pool (up to {os.cpu_count()} threads)")


def _cpu_baseline(cfg, batch, ctx, seconds_budget: float, threads_note: str):
    import numpy as np
    from oracle import attention_oracle as orc
    from oracle import llama_oracle as lo
    rng = np.random.default_rng(0)
    H, HK, D, hid, inter = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.hidden_size, cfg.intermediate_size
    b = min(2, batch)
    mb = (ctx + PAGE - 1) // PAGE
    nb = b * mb
    f = lambda *s: orc.round_to_io(rng.standard_normal(s, dtype=np.float32), True)
    kv = [np.stack([f(nb, PAGE, HK, D), f(nb, PAGE, HK, D)])]
    lw = lo.LayerWeights(f(H * D, hid) * hid ** -0.5, f(HK * D, hid) * hid ** -0.5, f(HK * D, hid) * hid ** -0.5,
                         f(hid, H * D) * hid ** -0.5, f(inter, hid) * hid ** -0.5, f(inter, hid) * hid ** -0.5,
                         f(hid, inter) * inter ** -0.5, np.ones(hid, np.float32), np.ones(hid, np.float32))
    x = f(b, hid)
    bt = rng.permutation(nb).astype(np.int32).reshape(b, mb)
    lens = np.full(b, ctx)
    slots = bt[np.arange(b), (lens - 1) // PAGE].astype(np.int64) * PAGE + (lens - 1) % PAGE
    cos, sin = orc.rope_tables(D, ctx + PAGE, is_bf16=True)
    t0 = time.perf_counter()
    n = 0
    while True:   # one layer of the step for b sequences (the final norm of decode_step is negligible next to it)
        lo.decode_step(x, [lw], kv, np.ones(hid, np.float32), None, cos, sin, lens - 1, slots, bt, lens, H, HK, D,
                       cfg.rms_norm_eps, True)
        n += 1
        if time.perf_counter() - t0 > seconds_budget or n >= 8:
            break
    dt = (time.perf_counter() - t0) / n            # seconds per layer for b sequences
    tok_s = b / (dt * cfg.num_hidden_layers)
    return {"value": tok_s, "unit": "tokens/s", "cores": 1 if threads_note.startswith("1 ") else (os.cpu_count() or 1),
            "kind": "port",
            "sample": f"numpy oracle (oracle/llama_oracle.py + oracle/attention_oracle.py), {n} pass(es) of one full layer "
                      f"(norms, q/k/v/o, RoPE, cache write, paged attention, SwiGLU MLP) for {b} of the {batch} sequences at "
                      f"ctx {ctx}; scaled to {cfg.num_hidden_layers} layers (lm_head not included); host threads: {threads_note}"}


def ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the newest
    committed `ncu --set full` capture under profiles/ (taken on the same config-2 inputs); None when no
    capture is present.  Evidence, not a timing."""
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


def tp_roofline(ms_step: float, hbm_bytes_per_rank: int, peak_hbm_gbs: float, tp: int, payload_bytes: int,
                all_reduces: int, fused_push: bool, nvlink_gbs: float = 770.0):
    """north_star: "1/2/4/8-GPU TP scaling reported as fraction of NVLink-bound roofline".  The target time of a
    tensor-parallel step is the slower of (this rank's HBM bytes / measured HBM peak) and (the bytes that MUST
    cross NVLink / link bandwidth) — B200_PROFILING.md's rule for a fused compute+collective, with its measured
    peer-copy figure of 770 GB/s per direction per GPU (900 nominal).  An all-reduce of P bytes over N ranks must
    move 2 (N-1)/N P out of (and into) every rank (reduce-scatter + all-gather, the bandwidth-optimal schedule);
    `protocol` is what this repo's one-shot push actually sends: (N-1) P payload bytes in 8-byte {4 data, epoch}
    units = 2 (N-1) P on the wire — latency-optimal, not bandwidth-optimal, at a 256 KiB payload."""
    hbm_ms = hbm_bytes_per_rank / (peak_hbm_gbs * 1e9) * 1e3
    must = all_reduces * 2.0 * (tp - 1) / tp * payload_bytes
    sent = all_reduces * 2.0 * (tp - 1) * payload_bytes if fused_push else must
    nv_ms, nv_sent_ms = must / (nvlink_gbs * 1e9) * 1e3, sent / (nvlink_gbs * 1e9) * 1e3
    target = max(hbm_ms, nv_ms)
    return {"bound": "hbm" if hbm_ms >= nv_ms else "nvlink", "hbm_ms": hbm_ms, "nvlink_ms": nv_ms,
            "nvlink_ms_as_sent": nv_sent_ms, "nvlink_bytes_per_rank_per_step": int(must),
            "nvlink_bytes_per_rank_per_step_as_sent": int(sent), "target_ms": target, "frac": target / ms_step,
            "frac_serial": (hbm_ms + nv_ms) / ms_step,
            "nvlink_gbs": nvlink_gbs, "hbm_gbs": peak_hbm_gbs,
            "note": "target = max(HBM time, NVLink time), i.e. perfect overlap; frac_serial = (HBM + NVLink) / measured, "
                    "no overlap; peaks: MEASURED_PEAKS.json hbm_gbs and B200_PROFILING.md's measured peer copy "
                    "(770 GB/s per direction per GPU), both 'of measured'"}


def tp_parity(torch, args, cfg, batch, ctx, layers, rank, world, dev, arm, dist):
    """N > 1: the tensor-parallel step's result against a tp=1 recomputation of the SAME step on rank 0
    (full weights and caches rebuilt from the same seeds).  Two numbers: after ONE layer without the head
    (two all-reduces: the bit budget is one bf16 rounding per all-reduce plus the regrouping of the fp32 partial
    sums) the largest difference in bf16 ulps taken at each row's rms; after the whole step the norm-wise relative difference of the
    logits and the fraction of sequences whose greedy token agrees."""
    from atoma_infer_b200.llama import DecodeStep
    res = {}
    for tag, nl, head in (("one_layer", 1, False), ("full_step", layers, True)):
        sub = arm.wl.shard.prefix(nl, with_head=head)
        st = DecodeStep(sub, arm.wl.kv[:nl], max_batch=batch, p2p=arm.step.p2p, want_logits=head, want_hidden=not head)
        st.bind(arm.tokens_dev, arm.pos, arm.slots, arm.bt, arm.lens)
        st.launch()
        torch.cuda.synchronize()
        dist.barrier()
        got = (st.logits if head else st.hidden_out)[:batch].float()
        tok = st.next_tokens[:batch].clone() if head else None
        if head and getattr(arm.step, "vocab_parallel_head", False):
            # the timed step samples through the vocab-parallel head; `st` has the replicated head: same tokens, same logits
            arm.step.run()
            torch.cuda.synchronize()
            res["vocab_parallel_tokens_identical"] = bool(torch.equal(arm.step.next_tokens[:batch], tok) and
                                                          torch.equal(arm.step.next_logits[:batch], st.next_logits[:batch]))
        if rank == 0:
            w1 = Workload(torch, cfg, batch, ctx, 0, 1, nl, dev)
            s1 = DecodeStep(w1.shard.prefix(nl, with_head=head), w1.kv, max_batch=batch, want_logits=head, want_hidden=not head)
            s1.bind(arm.tokens_dev, arm.pos, arm.slots, arm.bt, arm.lens)
            s1.launch()
            torch.cuda.synchronize()
            want = (s1.logits if head else s1.hidden_out)[:batch].float()
            diff = (got - want).abs()
            rel = (diff.norm() / want.norm()).item()
            if head:
                res["logits_rel_err"] = rel
                res["greedy_agreement"] = (tok == s1.next_tokens[:batch]).float().mean().item()
                # Where the greedy tokens differ the choice was a near-tie (a random-init model has a flat head): how far
                # below the tp=1 maximum the tp=1 logit of the token the TP step chose lies, in units of the row's logit
                # standard deviation (worst sequence).
                chosen = want.gather(1, tok.long().unsqueeze(1)).squeeze(1)
                res["greedy_choice_gap_sigma_max"] = ((want.max(dim=-1).values - chosen) / want.std(dim=-1)).max().item()
            else:
                ulp = torch.exp2(torch.floor(torch.log2(want.pow(2).mean(dim=-1, keepdim=True).sqrt().clamp_min(2.0 ** -126))) - 7)   # bf16 ulp at the row's rms
                res["one_layer_max_ulp"] = (diff / ulp).max().item()
                res["one_layer_rel_err"] = rel
            del w1, s1
            torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        res["ok"] = bool(res["one_layer_max_ulp"] <= 8.0 and res["one_layer_rel_err"] <= 6e-3 and res["logits_rel_err"] <= 5e-2 and
                         res.get("vocab_parallel_tokens_identical", True))
        res["how"] = ("rank 0 rebuilds the tp=1 model and caches from the same seeds and reruns the step; one_layer = 1 layer, "
                      "no head (2 all-reduces), full_step = the timed configuration")
    return res


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
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
    if args.config == 3:
        import bench_mixed
        return bench_mixed.run(args, torch, dev, json_fd)
    conf = CONFIGS[args.config]
    cfg = model_config(conf["model"])
    batch, ctx = conf["batch"], conf["ctx"]
    layers = args.layers or cfg.num_hidden_layers
    dist = None
    tp = world
    if conf["tp"] is not None and world != conf["tp"] and not args.emulate_tp:
        raise SystemExit(f"config {args.config} runs at tensor-parallel degree {conf['tp']}: launch with --gpus {conf['tp']} "
                         f"(or --emulate-tp {conf['tp']} for one rank's shard on one GPU)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.emulate_tp:
        if world != 1 or args.impl != "ours":
            raise SystemExit("--emulate-tp is a single-process diagnostic of our arm")
        tp = args.emulate_tp
    if cfg.num_key_value_heads % tp != 0:
        raise SystemExit(f"tensor parallel degree {tp} must divide the {cfg.num_key_value_heads} kv heads")

    p2p = None
    if args.impl == "ours":
        from atoma_infer_b200 import _lib, ops
        _lib.lib()  # fail loudly if libflashattention.so is missing
        # weights are immutable during a decode step: the GEMMs may stream them ahead of their dependency; the
        # "only the final page is being written" promise of early KV streaming is made per call by the executor
        ops.set_launch_mode(pdl=not args.no_pdl, const_weights=not args.no_pdl)
    else:
        from oracle.ref_plan import RefKernels
        if not RefKernels.available():
            raise SystemExit("oracle/_ref/libatoma_ref.so is missing: run __graft_entry__.build() where /root/reference exists")
    mark("init done, building the workload")
    wl = Workload(torch, cfg, batch, ctx, rank if not args.emulate_tp else 0, tp, layers, dev)
    mark(f"workload built: {wl.shard.weight_bytes() / 1e9:.2f} GB of weights, {len(wl.kv)} caches of {wl.kv[0].numel() * 2 / 1e6:.0f} MB")
    if args.impl == "ours":
        if dist is not None:
            from atoma_infer_b200.p2p import P2PComm
            p2p = P2PComm.from_torch_distributed(dev, batch * cfg.hidden_size * 2)   # fails loudly without peer access
            mark("peer-to-peer regions mapped")
        arm = OursArm(torch, wl, p2p, no_comm=bool(args.emulate_tp))
    else:
        arm = ReferenceArm(torch, wl, dist)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- first step eagerly (library workspaces, cuBLAS handles), then the graph ----
    launches_per_step = 0
    graph_note = "eager, one host thread"
    if args.impl == "ours":
        launches_per_step = arm.step.launches_per_step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        if not args.no_graph:
            arm.step.capture()
            graph_note = "cuda graph replay" + (" (all-reduce fused into the row-parallel GEMM epilogues + the add/RMSNorm kernels)" if p2p else "")
        run = arm.step.run
    else:
        arm.step()
        torch.cuda.synchronize()
        run = arm.step
    mark("first step done")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # Untimed: W warm-up steps, then a fixed number of extra steps (the same on every rank: the step
    # holds collectives) that keep the GPU under the step's load while nvidia-smi comes up.
    for _ in range(max(args.warmup, 3) + (60 if args.impl == "ours" else 10)):
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

    # ---- timed region 2 (e2e): host buffers -> H2D -> step -> D2H of the sampled tokens, every step ----
    if args.impl == "ours":
        for _ in range(2):
            arm.upload(); run(); arm.download()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        arm.upload()
        run()
        if args.impl == "ours":
            arm.download()
        stream.synchronize()  # the step's result is on the host before the next step starts
    e3.record()
    barrier()
    t_e2e_wall = time.perf_counter() - t0
    ms_e2e = max(e2.elapsed_time(e3), t_e2e_wall * 1e3)  # host wall-clock includes the syncs
    clocks = sampler.stop(t_clock0, time.time()) if sampler else None
    mark("timed region 2 done")
    if args.impl == "ours":
        arm.step.check()   # loud if a bounded in-kernel wait (grid barrier, all-reduce peer) ever expired during the runs

    # ---- dominant kernel alone: paged decode attention, one launch per layer (distinct caches) ----
    sh = wl.shard.shard
    h_r, hk_r, D = sh.heads, sh.kv_heads, cfg.head_dim
    qs = torch.randn((batch, 1, h_r, D), device=dev, dtype=torch.float32).to(torch.bfloat16)
    bt_dev, lens_dev = (arm.bt, arm.lens) if args.impl == "ours" else (arm.d["bt"], arm.d["lens"])
    if args.impl == "ours":
        from atoma_infer_b200 import ops

        def attn_only(l):
            return ops.flash_attn_kv_cache_full(qs, wl.kv[l][0], wl.kv[l][1], None, D ** -0.5, bt_dev, lens_dev, True)
    else:
        o_buf = torch.empty_like(qs)
        lse_buf = torch.zeros(arm.plan.lse_elems, dtype=torch.float32, device=dev)
        acc = (torch.zeros(max(arm.plan.lseaccum_elems, 1), dtype=torch.float32, device=dev),
               torch.zeros(max(arm.plan.oaccum_elems, 1), dtype=torch.float32, device=dev))

        def attn_only(l):
            arm.ref.run_kv_cache(arm.plan, qs.data_ptr(), wl.kv[l][0].data_ptr(), wl.kv[l][1].data_ptr(), o_buf.data_ptr(),
                                 lse_buf.data_ptr(), bt_dev.data_ptr(), lens_dev.data_ptr(),
                                 acc[0].data_ptr() if arm.plan.num_splits > 1 else None,
                                 acc[1].data_ptr() if arm.plan.num_splits > 1 else None)

    def attn_pass():
        for l in range(layers):
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
    us_kernel = k0.elapsed_time(k1) * 1e3 / (reps * layers)

    # ---- tensor-parallel parity: the step's result vs a tp=1 recomputation on rank 0 ----
    parity = None
    if dist is not None and args.impl == "ours" and not args.no_tp_parity:
        mark("tp parity check")
        parity = tp_parity(torch, args, cfg, batch, ctx, layers, rank, world, dev, arm, dist)

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
        alg_bytes = wl.kv_bytes_per_layer() + 2 * batch * h_r * D * 2 + batch * wl.mb * 4
        achieved = alg_bytes / (us_kernel * 1e-6) / 1e9
        ms_step = ms_dev / args.steps
        ms_step_e2e = ms_e2e / args.steps
        step_bytes = wl.shard.weight_bytes() + layers * wl.kv_bytes_per_layer()   # per rank: all weights + all KV once
        vocab_parallel = args.impl == "ours" and getattr(arm.step, "vocab_parallel_head", False)
        if vocab_parallel:   # this rank streams only its [vocab / tp, hidden] slice of lm_head
            lm = wl.shard.lm_head
            step_bytes -= lm.numel() * lm.element_size() * (tp - 1) // tp
        if wl.shard.embed is not None and wl.shard.embed is not wl.shard.lm_head:
            pass   # the embedding table is gathered (batch rows), not streamed
        line = {
            "metric": conf["metric"],
            "value": batch / (ms_step * 1e-3),
            "unit": "tokens/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "bf16",
            "data": "synthetic (seeded random-init weights of the named architecture, N(0,1) KV caches, random token ids and block tables)",
            "config": {
                "workload": f"{conf['name']}; the WHOLE decode step ({layers} layers: RMSNorm, q/k/v projections, RoPE, "
                            f"reshape_and_cache_flash, paged decode attention, o_proj, residual, RMSNorm, gate/up, SiLU*up, down_proj, "
                            f"residual; final norm, lm_head, sampling)",
                "batch": batch, "ctx": ctx, "layers": layers, "heads": cfg.num_attention_heads,
                "kv_heads": cfg.num_key_value_heads, "head_dim": D, "hidden": cfg.hidden_size,
                "intermediate": cfg.intermediate_size, "vocab": cfg.vocab_size, "block_size": PAGE, "parallelism": f"tp{tp}",
                "l2": f"inputs larger than L2: {step_bytes / 1e9:.1f} GB of weights + KV per rank per step, every byte distinct",
                "launch": graph_note + (", programmatic dependent launch, weight prefetch, early KV streaming (per call)"
                                        if args.impl == "ours" and not args.no_pdl else ""),
                "all_reduce": (None if tp == 1 or args.emulate_tp else
                               "fused: row-parallel GEMM epilogue pushes its bf16 partial tile to every rank over NVLink peer memory; "
                               "the add+RMSNorm kernel sums the addends in rank order" if args.impl == "ours" else
                               "torch.distributed.all_reduce (ncclAllReduce), eager"),
                "gemm": ("atoma tcgen05 kernels (gemm_skinny.cu: fused qkv+RoPE+cache write, o_proj, fused gate_up+SwiGLU, "
                         "down_proj, lm_head)" if args.impl == "ours" else
                         "cuBLAS via torch.matmul (as in the reference: candle Linear -> cuBLAS), q/k/v and gate/up as separate calls"),
                "sampling": ("greedy on the device (argmax kernel), 8 bytes per sequence to the host" +
                             ("; vocab-parallel head: each rank multiplies by its 1/tp slice of lm_head and the ranks exchange their "
                              "argmax candidates over NVLink peer memory (same tokens as the replicated head: tp_parity."
                              "vocab_parallel_tokens_identical)" if vocab_parallel else "")) if args.impl == "ours" else
                            "logits row to the host per sequence + host argmax (model_executor.rs:227-262)",
                "wall_ms_per_step": t_wall * 1e3 / args.steps,
                "hbm_bytes_per_step_per_rank": step_bytes,
                "step_hbm_roofline_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak_hbm,
            },
            "e2e": {"value": batch / (ms_step_e2e * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": arm.h2d_bytes(), "d2h_bytes_per_step": arm.d2h_bytes(),
                    "ms_per_step": ms_step_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "decode_paged_kernel" if args.impl == "ours" else "flash_fwd_splitkv_kernel",
                         "achieved": achieved, "peak": peak_hbm, "unit": "GB/s", "frac": achieved / peak_hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "us_per_launch": us_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                         "frac_of_8TBps": achieved / 8000.0, "traffic": None},
        }
        if args.impl == "ours" and tp == 1 and args.config == 2:
            t = ncu_traffic("decode_paged_kernel")
            if t:
                line["roofline"]["traffic"] = t["bytes"]
                line["roofline"]["traffic_source"] = t["file"] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, same inputs)"
        if tp > 1 and not args.emulate_tp:
            line["tp_roofline"] = tp_roofline(ms_step_e2e if args.impl == "reference" else ms_step, step_bytes, peak_hbm, tp, batch * cfg.hidden_size * 2, 2 * layers,
                                              fused_push=args.impl == "ours")
        if parity is not None:
            line["tp_parity"] = parity
        if args.emulate_tp:
            line["diagnostic"] = f"one rank's shard of a tp{tp} step on one GPU, no all-reduce: NOT a bench value"
        if args.layers and args.layers != cfg.num_hidden_layers:
            line["diagnostic"] = f"{args.layers} of {cfg.num_hidden_layers} layers: NOT a bench value"
        if args.impl == "reference":
            line["impl"] = "reference"
            line["value"] = line["e2e"]["value"]   # the reference's number is host-timed by definition
            line["ms_per_step"] = ms_step_e2e
            line["e2e"] = {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
            line["config"]["device_ms_per_step"] = ms_step
            line["gpu_launches"] = 0
            line["cpu_baseline"] = {
                "value": line["value"], "unit": "tokens/s", "cores": 1, "kind": "reference",
                "sample": "the reference's own CUDA kernels (oracle/_ref, built from /root/reference/csrc/kernels for sm_100) + cuBLAS "
                          "+ element-wise torch kernels for candle's, driven from ONE host thread per GPU with host wall-clock timing, "
                          "as the reference times itself (model_executor.rs:331-347); the reference has no CPU implementation of "
                          "this path"}
        elif not args.no_cpu_baseline and world == 1 and not args.emulate_tp:
            line["cpu_baseline"] = cpu_baseline(cfg, batch, ctx)   # rank 0, N = 1 only: the CPU port does not change with N
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
        if args.impl == "ours":
            arm.step.release()   # the graph holds kernels that poll peer memory: drop it before unmapping
            if p2p is not None:
                p2p.check()
                p2p.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
