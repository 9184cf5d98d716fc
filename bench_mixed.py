"""bench.py --config 3: BASELINE.json configs[2], the continuous-batching mix, through the attention
operator `FlashAttention.forward` (models/src/flash_attention.rs:322-469): ONE flattened batch per layer
holding an 8192-token prompt chunk (prefill tokens first, flash_attention.rs:156-174) followed by 32
decode tokens of sequences that are 1..2048 tokens past their own 8k prompt (context 8193..10240) —
cache write for all 8224 tokens, causal varlen prefill attention on tokens [0, 8192), paged decode
attention on tokens [8192, 8224), 32 layers with distinct caches.  Not the headline (config 2 is); one
JSON line in the same format.  Attention operator only: q/k/v arrive as the projections would leave them."""
from __future__ import annotations

import json
import os
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PAGE, L, H, HK, D = 16, 32, 32, 8, 128
P_TOKENS, DN = 8192, 32


def run(args, torch, dev, json_fd):
    import numpy as np
    from atoma_infer_b200 import _lib, ops
    from atoma_infer_b200.flash_attention import FlashAttention, FlashAttentionMetadata
    if args.impl != "ours":
        os.write(json_fd, (json.dumps({"impl": "reference", "unavailable": "config 3 is measured for the product only; the "
                                       "reference arm covers the headline config 2 (and 1, 4)"}) + "\n").encode())
        return
    _lib.lib()
    ops.set_launch_mode(pdl=not args.no_pdl, const_weights=not args.no_pdl)
    layers = args.layers or L
    rng = np.random.default_rng(3)
    dec_ctx = rng.integers(P_TOKENS + 1, P_TOKENS + 2049, size=DN)
    seq_lens = np.concatenate([[P_TOKENS], dec_ctx]).astype(np.int32)
    blocks_per = (seq_lens + PAGE - 1) // PAGE
    mb = int(blocks_per.max())
    nb = int(blocks_per.sum()) + 8
    perm = rng.permutation(nb)
    bt = np.zeros((1 + DN, mb), np.int32)
    off = 0
    for i, n in enumerate(blocks_per):
        bt[i, :n] = perm[off:off + n]
        off += n
    # slot mapping: every prompt token, then the token being decoded of each sequence (worker.rs:392-401)
    slots = [bt[0, t // PAGE] * PAGE + t % PAGE for t in range(P_TOKENS)]
    slots += [bt[1 + i, (c - 1) // PAGE] * PAGE + (c - 1) % PAGE for i, c in enumerate(dec_ctx)]
    T = P_TOKENS + DN
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.randn((T, H, D), generator=g, device=dev, dtype=torch.float32).to(bf)
    k = torch.randn((T, HK, D), generator=g, device=dev, dtype=torch.float32).to(bf)
    v = torch.randn((T, HK, D), generator=g, device=dev, dtype=torch.float32).to(bf)
    kv = []
    for _ in range(layers):
        c = torch.empty((2, nb, PAGE, HK, D), device=dev, dtype=bf)
        c.normal_(generator=g)
        kv.append(c)
    to = lambda a, dt: torch.from_numpy(np.asarray(a)).to(dt).to(dev)
    md = FlashAttentionMetadata.new(
        context_lengths=to(np.concatenate([[0], dec_ctx - 1]), torch.int32), slot_mapping=to(slots, torch.int64),
        query_start_locations=to([0, P_TOKENS], torch.int32), num_prefill_tokens=P_TOKENS, num_decoding_tokens=DN,
        max_query_length=P_TOKENS, max_decoding_sequence_length=int(dec_ctx.max()), max_prefill_sequence_length=P_TOKENS,
        num_prefill_sequences=1, sequence_start_locations=to([0, P_TOKENS], torch.int32),
        sequence_lengths=to(seq_lens, torch.int32), block_table=to(bt, torch.int32), prefix_caching=False)
    md.decoding_metadata.block_tables = to(bt[1:], torch.int32)
    attn = FlashAttention(H, HK, D, D ** -0.5, device=dev)

    def step():
        out = None
        for l in range(layers):
            out = attn.forward(q, k, v, kv[l], md)
        return out

    ops.reset_launch_count()
    out = step()
    torch.cuda.synchronize()
    launches = ops.launch_count()
    assert torch.isfinite(out.float()).all()
    run_fn = step
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            step()
        run_fn = gr.replay
    import bench
    sampler = bench.ClockSampler(dev.index or 0)
    for _ in range(max(args.warmup, 3) + 10):
        run_fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_clock0 = time.time()
    e0.record()
    for _ in range(args.steps):
        run_fn()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    # e2e: q/k/v and the step's metadata from pinned host memory every step, the attention output back
    host = {n: t.cpu().pin_memory() for n, t in (("q", q), ("k", k), ("v", v))}
    out_host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        q.copy_(host["q"], non_blocking=True); k.copy_(host["k"], non_blocking=True); v.copy_(host["v"], non_blocking=True)
        run_fn()
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop(t_clock0, time.time())
    # the dominant kernel alone: causal prefill attention over the 8192-token chunk
    cu = to([0, P_TOKENS], torch.int32)
    qp, kp, vp = q[:P_TOKENS], k[:P_TOKENS], v[:P_TOKENS]
    for _ in range(3):
        ops.flash_attn_varlen(qp, kp, vp, cu, cu, P_TOKENS, P_TOKENS, D ** -0.5, True)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(10):
        ops.flash_attn_varlen(qp, kp, vp, cu, cu, P_TOKENS, P_TOKENS, D ** -0.5, True)
    k1.record()
    torch.cuda.synchronize()
    us_prefill = k0.elapsed_time(k1) * 1e3 / 10
    flops = 4.0 * (P_TOKENS * (P_TOKENS + 1) / 2) * D * H
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    achieved = flops / (us_prefill * 1e-6) / 1e12
    line = {
        "metric": "tokens/sec through the attention operator, Llama-3.1-8B bf16 prefill 8k + 32 decode tokens (continuous-batching mix)",
        "value": T / (ms_step * 1e-3), "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (seeded N(0,1) q/k/v and KV caches, random block tables)",
        "config": {"workload": "BASELINE.json configs[2]: Llama-3.1-8B bf16 prefill 8k -> decode 2k continuous-batching mix, 1xB200: "
                               f"{layers} layers x FlashAttention.forward on one flattened batch (8192 prefill tokens + 32 decode tokens at "
                               "context 8193..10240): cache write, causal varlen prefill attention, paged decode attention; attention "
                               "operator only (not the headline config)",
                   "prefill_tokens": P_TOKENS, "decode_tokens": DN, "layers": layers, "heads": H, "kv_heads": HK, "head_dim": D,
                   "block_size": PAGE, "l2": "a distinct KV cache per layer, 45 GB in total",
                   "launch": ("eager" if args.no_graph else "cuda graph replay") + ("" if args.no_pdl else ", programmatic dependent launch")},
        "e2e": {"value": T / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": sum(t.numel() * 2 for t in host.values()), "d2h_bytes_per_step": out_host.numel() * 2},
        "gpu_launches": launches * args.steps, "launches_per_step": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "prefill_attn_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf, "us_per_launch": us_prefill, "algorithmic_flops_per_launch": flops,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, of measured)" if peaks else "fallback 1590 (of fallback)",
                     "traffic": None},
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
