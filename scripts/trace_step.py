"""Timeline of the decode step (config 2 shapes, a few layers): every kernel of the captured step writes per-CTA
globaltimer stamps (atoma_debug_step_trace); the graph is replayed and the stamps of the LAST replay are turned
into one line per kernel: when its first CTA started, when its weights / KV began to land, when its last byte
landed, when it ended — relative to the previous kernel — i.e. where HBM is idle at kernel boundaries.

  python scripts/trace_step.py [--layers 4] [--tp 1]   ->  gpurun_out/trace_step.json + a table on stdout
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from atoma_infer_b200 import _lib, ops  # noqa: E402
from atoma_infer_b200.llama import DecodeStep  # noqa: E402

STRIDE = 1280


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--fusion", type=int, default=0, help="atoma_debug_step_fusion flags (2 = MLP as two kernels)")
    ap.add_argument("--out", default="gpurun_out/trace_step.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    cfg = bench.model_config("8b")
    wl = bench.Workload(torch, cfg, args.batch, args.ctx, 0, args.tp, args.layers, dev)
    shard = wl.shard.prefix(args.layers, with_head=False)
    tokens, pos, slots = wl.tokens.to(dev), wl.pos.to(dev), wl.slots.to(dev)
    bt, lens = wl.perm.contiguous().to(dev), wl.lens.to(dev)
    ops.set_launch_mode(pdl=True, const_weights=True)
    L.atoma_debug_step_fusion(args.fusion)
    st = DecodeStep(shard, wl.kv[: args.layers], max_batch=args.batch, no_comm=args.tp > 1, want_hidden=True)
    st.bind(tokens, pos, slots, bt, lens)
    st.launch()                      # library workspaces
    torch.cuda.synchronize()
    per_step = st.launches_per_step()
    torch.cuda.synchronize()
    nk = 2 * per_step + 8
    buf = torch.zeros((nk * STRIDE,), dtype=torch.int64, device=dev)
    L.atoma_debug_step_trace(buf.data_ptr(), nk)
    st.capture()                     # one eager step (slices 0..per_step-1) + the captured one (the next per_step)
    names = C.create_string_buffer(1 << 16)
    L.atoma_debug_step_trace_names(names, 1 << 16)
    L.atoma_debug_step_trace(None, 0)
    labels = names.value.decode().strip().split("\n")
    for _ in range(5):
        st.run()
    torch.cuda.synchronize()
    buf.zero_()
    torch.cuda.synchronize()
    st.run()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(nk, STRIDE // 8, 8)
    rows = []
    # traced kernels of the captured step: the second half of the labels
    idx = [i for i in range(len(labels))][len(labels) // 2:]
    t0 = None
    last_end_on_sm, last_kernel_name = {}, None
    prev_end = None
    prev_last_land = None
    for i in idx:
        k = t[i]
        name = labels[i]
        if name.startswith("gemm") or name.startswith("chain"):
            live = k[:, 3] > 0
            start, end = k[live, 0], k[live, 3]
            first_land, last_land = k[live, 1], k[live, 7]
            dep = k[live, 6]
            first_issue = k[live, 0]
        elif name.startswith("decode"):
            live = k[:, 3] > 0
            start, end = k[live, 1], k[live, 3]
            first_land = last_land = dep = first_issue = None
            stream_done = k[live, 2]
        else:
            live = k[:, 3] > 0
            start, end = k[live, 1], k[live, 3]
            first_land = last_land = first_issue = None
            dep = k[live, 2]
        if not live.any():
            continue
        if t0 is None:
            t0 = start.min()
        us = lambda v: round(float(v - t0) / 1e3, 2)
        rec = {"kernel": name, "ctas": int(live.sum()), "first_cta_start": us(start.min()), "last_cta_start": us(start.max()),
               "first_cta_end": us(end.min()), "end": us(end.max()), "dur": round(float(end.max() - start.min()) / 1e3, 2)}
        if dep is not None:
            rec["dependency_resolved"] = us(np.median(dep))
        if first_land is not None:
            rec["first_stage_landed"] = us(np.median(first_land))
            rec["last_stage_landed"] = [us(last_land.min()), us(np.median(last_land)), us(last_land.max())]
            if name.startswith("gemm"):
                dd = (k[live, 5] - k[live, 3]) / 1e3
                rec["tmem_dealloc_us"] = [round(float(dd.min()), 2), round(float(np.median(dd)), 2), round(float(dd.max()), 2)]
        if name.startswith("decode"):
            rec["streaming_done"] = [us(stream_done.min()), us(np.median(stream_done)), us(stream_done.max())]
        if prev_end is not None:
            rec["start_minus_prev_end"] = round(rec["first_cta_start"] - prev_end, 2)
        prev_end = rec["end"]
        # SM hand-over: for every SM, this kernel's CTA start minus the end of the previous traced CTA on that SM
        sm = (k[live, 4] if name.startswith("gemm") or name == "chain gate_up+silu" else k[live, 0] if name.startswith("decode") else None)
        if sm is not None:
            d = []
            for s_, st_, en_ in zip(sm.tolist(), start.tolist(), end.tolist()):
                if s_ in last_end_on_sm:
                    d.append((st_ - last_end_on_sm[s_][0]) / 1e3)
            if d:
                d = np.asarray(d)
                rec["handover_us"] = {"after": last_kernel_name, "n": int(d.size), "min": round(float(d.min()), 2),
                                      "median": round(float(np.median(d)), 2), "max": round(float(d.max()), 2)}
            for s_, en_ in zip(sm.tolist(), end.tolist()):
                last_end_on_sm[s_] = (en_, name)
            last_kernel_name = name
        rows.append(rec)
    (ROOT / args.out).parent.mkdir(exist_ok=True)
    (ROOT / args.out).write_text(json.dumps(rows, indent=1))
    for r in rows:
        print(json.dumps(r))
    ops.set_launch_mode()


if __name__ == "__main__":
    main()
