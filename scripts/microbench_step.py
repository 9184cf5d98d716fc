"""A/B timing of the decode step (config 2 shapes, a few layers, no lm_head by default) under library tuning
knobs: each variant re-captures the step's CUDA graph with the knob set and times graph replays with CUDA
events.  Diagnostics for DESIGN.md, not a bench value.

  python scripts/microbench_step.py [--layers 8] [--head] [--variants default,b100,b96,...]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from atoma_infer_b200 import _lib, ops  # noqa: E402
from atoma_infer_b200.llama import DecodeStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--head", action="store_true")
    ap.add_argument("--tp", type=int, default=1, help="one rank's shard of a tp-N step, no all-reduce")
    ap.add_argument("--variants", default="default,b100,b80,default")
    ap.add_argument("--out", default="gpurun_out/microbench_step.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    cfg = bench.model_config("8b")
    wl = bench.Workload(torch, cfg, args.batch, args.ctx, 0, args.tp, args.layers, dev)
    shard = wl.shard.prefix(args.layers, with_head=args.head)
    tokens, pos, slots = wl.tokens.to(dev), wl.pos.to(dev), wl.slots.to(dev)
    bt, lens = wl.perm.contiguous().to(dev), wl.lens.to(dev)
    res = []
    first_out = None
    for v in args.variants.split(","):
        ops.set_launch_mode(pdl=True, const_weights=True)
        L.atoma_debug_skinny_budget(0)
        L.atoma_debug_skinny_tuning(0, 0, 0)
        L.atoma_debug_step_prefetch(0, 32)
        L.atoma_debug_step_fusion(0)
        L.atoma_debug_skinny_mode(0, 1)
        if v == "dbuf":             # two accumulator buffers even for single-tile CTAs (the round-2 default before)
            L.atoma_debug_skinny_mode(8, 1)
        elif v.startswith("late"):    # late<N>: GEMMs release their dependents N stages before their last weight load
            L.atoma_debug_skinny_mode(int(v[4:]) << 8, 1)
        elif v.startswith("fuse"):
            L.atoma_debug_step_fusion(int(v[4:]))
        elif v.startswith("pf"):      # pf<mask>[_<cap MiB>]
            mk, _, cap = v[2:].partition("_")
            L.atoma_debug_step_prefetch(int(mk), int(cap) if cap else 32)
        elif v.startswith("b"):
            L.atoma_debug_skinny_budget(int(v[1:]))
        elif v == "nopdl":
            ops.set_launch_mode()
        elif v.startswith("kbs"):
            L.atoma_debug_skinny_tuning(0, int(v[3:]), 0)
        st = DecodeStep(shard, wl.kv[: args.layers], max_batch=args.batch, no_comm=args.tp > 1, want_hidden=not args.head)
        st.bind(tokens, pos, slots, bt, lens)
        st.capture()
        for _ in range(5):
            st.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            st.run()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        out = (st.logits if args.head and st.logits is not None else st.hidden_out)
        out = out.clone() if out is not None else st.next_tokens.clone()
        if first_out is None:
            first_out = out
        rec = {"variant": v, "layers": args.layers, "us_per_step": round(us, 1), "us_per_layer": round(us / args.layers, 2),
               "bit_identical_to_first": bool(torch.equal(out, first_out)), "finite": bool(torch.isfinite(out.float()).all())}
        print(json.dumps(rec), flush=True)
        res.append(rec)
        st.release()
        del st
    L.atoma_debug_skinny_budget(0)
    L.atoma_debug_skinny_tuning(0, 0, 0)
    L.atoma_debug_step_prefetch(0, 32)
    L.atoma_debug_step_fusion(0)
    L.atoma_debug_skinny_mode(0, 1)
    ops.set_launch_mode()
    (ROOT / args.out).parent.mkdir(exist_ok=True)
    (ROOT / args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
