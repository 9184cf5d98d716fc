"""Tensor-parallel decode step on real GPUs (SURVEY §8 rows e, f4): the row-parallel GEMMs push their
partial tiles into every rank's receive buffers from the epilogue (atoma_gemm_tn_push) and the add+RMSNorm
kernel sums them (atoma_allreduce_recv_add_rmsnorm) — against

  * the fp32 rank-order sum of the ranks' own partial projections, rounded once: BIT FOR BIT, on every rank,
    eagerly and replayed from a CUDA graph (odd and even numbers of all-reduces per graph: the receive buffer
    is chosen by the parity of a device-resident epoch, not by the host);
  * the tp=1 step of the same seeded model: after one layer within 8 bf16 ulps taken at the row's rms (one rounding per
    all-reduce + regrouped fp32 partial sums, then a norm and an MLP), after the whole step norm-wise.

Needs >= 2 GPUs (`gpurun --gpus N`); world = min(8, device_count).  Skipped on a one-GPU box."""
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


def _setup(T, cfg, rank, world, dev, layers):
    from atoma_infer_b200.llama import LlamaShard
    shard = LlamaShard.random_init(cfg, rank, world, dtype=torch.bfloat16, device=dev, seed=3, gen_device="cpu",
                                   num_layers=layers)
    page, MB = 16, 5
    NB = T * MB + 2
    g = torch.Generator().manual_seed(11)
    hk, D = cfg.num_key_value_heads, cfg.head_dim
    full = [torch.randn((2, NB, page, hk, D), generator=g).to(torch.bfloat16) for _ in range(layers)]
    hs = shard.shard.kv_range
    kv = [c[:, :, :, hs].contiguous().to(dev) for c in full]
    bt = torch.randperm(NB, generator=g)[: T * MB].to(torch.int32).reshape(T, MB).to(dev)
    lens = torch.randint(1, MB * page + 1, (T,), generator=g, dtype=torch.int32).to(dev)
    pos = lens.long() - 1
    slots = bt[torch.arange(T, device=dev), (pos // page)].long() * page + pos % page
    tokens = torch.randint(0, cfg.vocab_size, (T,), generator=g, dtype=torch.int32).to(dev)
    return shard, kv, bt, lens, pos, slots, tokens


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from atoma_infer_b200 import _lib, ops
    from atoma_infer_b200.llama import DecodeStep, LlamaConfig
    from atoma_infer_b200.p2p import P2PComm
    L = _lib.lib()
    bad = []
    T, hidden = 24, 1024
    comm = P2PComm.from_torch_distributed(dev, T * hidden * 2)
    regions = C.cast(comm._regions, C.POINTER(C.c_void_p))
    stream = lambda: torch.cuda.current_stream().cuda_stream

    # ---- (1) the fused all-reduce alone: push GEMM + receive/add/norm vs the rank-order fp32 sum ----
    K = 512
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn((T, K), generator=g).to(torch.bfloat16) for _ in range(world)]
    ws = [(torch.randn((hidden, K), generator=g) * K ** -0.5).to(torch.bfloat16) for _ in range(world)]
    resid = torch.randn((T, hidden), generator=g).to(torch.bfloat16).to(dev)
    nw = (1 + 0.1 * torch.randn((hidden,), generator=g)).to(torch.bfloat16).to(dev)
    x, w = xs[rank].to(dev), ws[rank].to(dev)
    parts = [ops.gemm_tn(xs[r].to(dev), ws[r].to(dev)) for r in range(world)]      # every rank recomputes all partials
    acc = torch.zeros((T, hidden), dtype=torch.float32, device=dev)
    for p in parts:
        acc = acc + p.float()                                                      # rank order, fp32
    want_sum = acc.to(torch.bfloat16)
    want_h, want_n = ops.add_rmsnorm(want_sum, resid, nw, 1e-5)

    def fused(h_out, n_out):
        L.atoma_gemm_tn_push(x.data_ptr(), w.data_ptr(), T, hidden, K, K, K, regions, rank, world, comm.slot_bytes, 1, stream())
        L.atoma_allreduce_recv_add_rmsnorm(regions, rank, world, comm.slot_bytes, resid.data_ptr(), h_out.data_ptr(),
                                           nw.data_ptr(), n_out.data_ptr(), T, hidden, hidden, hidden, hidden, 1e-5, 1, stream())
        if L.atoma_last_status() != 0:
            raise RuntimeError(L.atoma_last_error().decode())

    h, n = torch.empty_like(resid), torch.empty_like(resid)
    for it in range(5):                       # eager, odd count
        h.zero_(); n.zero_()
        fused(h, n)
        torch.cuda.synchronize()
        if not (torch.equal(h, want_h) and torch.equal(n, want_n)):
            bad.append(("eager", it, (h.float() - want_h.float()).abs().max().item()))
    ops.set_launch_mode(pdl=True, const_weights=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fused(h, n)
    torch.cuda.synchronize()
    dist.barrier()
    for per_graph in (3, 2):                  # an ODD number of all-reduces per graph, then an even one
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(per_graph):
                fused(h, n)
        for rep in range(3):
            h.zero_(); n.zero_()
            gr.replay()
            torch.cuda.synchronize()
            if not (torch.equal(h, want_h) and torch.equal(n, want_n)):
                bad.append(("graph", per_graph, rep))
        del gr
        dist.barrier()
    ops.set_launch_mode()

    # ---- (2) the whole tensor-parallel step vs tp = 1 ----
    cfg = LlamaConfig(hidden_size=hidden, intermediate_size=2048, vocab_size=4096, num_hidden_layers=3,
                      num_attention_heads=16, num_key_value_heads=8, max_position_embeddings=128)
    res = {}
    for tag, nl, head in (("one_layer", 1, False), ("full", 3, True)):
        shard, kv, bt, lens, pos, slots, tokens = _setup(T, cfg, rank, world, dev, nl)
        st = DecodeStep(shard.prefix(nl, with_head=head), kv, max_batch=T, p2p=comm, want_logits=head, want_hidden=not head)
        st.bind(tokens, pos, slots, bt, lens)
        st.launch()
        torch.cuda.synchronize()
        got = (st.logits if head else st.hidden_out)[:T].float()
        if head:
            # vocab-parallel head: every rank multiplies by its slice of lm_head and the ranks exchange their argmax
            # candidates over peer memory — same sampled tokens and logits as the replicated head, on every rank,
            # eagerly twice (epoch parity) and replayed from a graph
            assert not st.vocab_parallel_head
            vp = DecodeStep(shard.prefix(nl, with_head=True), kv, max_batch=T, p2p=comm)
            assert vp.vocab_parallel_head
            vp.bind(tokens, pos, slots, bt, lens)
            for rep in range(2):
                vp.next_tokens.zero_(); vp.next_logits.zero_()
                vp.launch()
                torch.cuda.synchronize()
                if not (torch.equal(vp.next_tokens[:T], st.next_tokens[:T]) and torch.equal(vp.next_logits[:T], st.next_logits[:T])):
                    bad.append((tag, "vocab-parallel sampling differs (eager)", rep))
            dist.barrier()
            vp.capture()
            for rep in range(3):
                vp.next_tokens.zero_(); vp.next_logits.zero_()
                vp.run()
                torch.cuda.synchronize()
                if not (torch.equal(vp.next_tokens[:T], st.next_tokens[:T]) and torch.equal(vp.next_logits[:T], st.next_logits[:T])):
                    bad.append((tag, "vocab-parallel sampling differs (graph)", rep))
            dist.barrier()
            vp.release()
        outs = [torch.empty_like(got.cpu()) for _ in range(world)]
        dist.all_gather(outs, got.cpu())
        for r in range(1, world):
            if not torch.equal(outs[r], outs[0]):
                bad.append((tag, "ranks disagree", r))
        if rank == 0:
            s1, kv1, *_ = _setup(T, cfg, 0, 1, dev, nl)
            one = DecodeStep(s1.prefix(nl, with_head=head), kv1, max_batch=T, want_logits=head, want_hidden=not head)
            one.bind(tokens, pos, slots, bt, lens)
            one.launch()
            torch.cuda.synchronize()
            want = (one.logits if head else one.hidden_out)[:T].float()
            diff = (got - want).abs()
            rel = (diff.norm() / want.norm()).item()
            ulp = torch.exp2(torch.floor(torch.log2(want.pow(2).mean(dim=-1, keepdim=True).sqrt().clamp_min(2.0 ** -126))) - 7)   # bf16 ulp at the row's rms
            res[tag] = (rel, (diff / ulp).max().item())
            if not head:   # diagnostics for the assertion message: where the largest ulp distance sits
                flat = int((diff / ulp).argmax())
                r_, c_ = divmod(flat, want.shape[1])
                rms = want.pow(2).mean(dim=-1).sqrt()
                res[tag + "_where"] = dict(row=r_, col=c_, want=want[r_, c_].item(), got=got[r_, c_].item(),
                                           row_rms=rms[r_].item(), rms_min=rms.min().item(), rms_max=rms.max().item(),
                                           len=int(lens[r_]), ulp=ulp[r_].item())
            if head:
                if rel > 3e-2:
                    bad.append((tag, "rel err", rel))
                # the ranks' caches together hold what tp = 1 wrote
                for l in range(nl):
                    hs = shard.shard.kv_range
                    a, b = kv[l].float(), kv1[l][:, :, :, hs].float()
                    if ((a - b).norm() / b.norm()).item() > 2e-2:
                        bad.append((tag, "cache", l))
            else:
                if (diff / ulp).max().item() > 8.0 or rel > 6e-3:
                    bad.append((tag, "one layer", rel, (diff / ulp).max().item()))
        dist.barrier()
    comm.check()
    torch.save((bad, res), os.path.join(out_dir, f"r{rank}.pt"))
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_tensor_parallel_step_matches_rank_order_sum_and_tp1(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus N)")
    import torch.multiprocessing as mp
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        bad, res = torch.load(tmp_path / f"r{r}.pt")
        assert bad == [], f"rank {r} (world {world}): {bad}; parity {res}"
