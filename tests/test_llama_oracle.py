"""oracle/llama_oracle.py (the numpy restatement of models/src/llama.rs for one decode step) and the
host-side weight sharding of atoma_infer_b200/llama.py — CPU only.

The candle-side ops are unpinned in the reference (no test there asserts their values), so what is
pinned here is the restatement's internal consistency: hand-computed known answers for each op, the sampling
rule the reference inherits from candle_transformers (`max_by(total_cmp)`: LAST maximum), and that a
tensor-parallel split of the model (llama_nccl.rs:153-202, multi_gpu.rs:8-57) computes the same step as the
unsplit model when no intermediate rounding is applied."""
import numpy as np
import pytest
import torch

from atoma_infer_b200.llama import LlamaConfig, LlamaShard, Llama3RopeScaling, rope_tables
from oracle import attention_oracle as ao
from oracle import llama_oracle as lo


def test_rms_norm_silu_and_add_known_answers():
    x = np.array([[3.0, 4.0, 0.0, 0.0]])
    w = np.array([1.0, 2.0, 1.0, 1.0])
    # mean(x^2) = 25/4 -> scale = 1/2.5
    np.testing.assert_allclose(lo.rms_norm(x, w, 0.0, exact=True), [[1.2, 3.2, 0.0, 0.0]], rtol=1e-12)
    # bf16 rounding of the result: 1.2 -> 1.203125, 3.2 -> 3.203125
    np.testing.assert_array_equal(lo.rms_norm(x, w, 0.0, is_bf16=True), [[1.203125, 3.203125, 0.0, 0.0]])
    assert lo.silu(np.array([0.0]))[0] == 0.0
    np.testing.assert_allclose(lo.silu(np.array([1.0, -1.0])), [0.7310585786300049, -0.2689414213699951], rtol=1e-12)
    # silu rounded to bf16 BEFORE the multiply (two candle kernels), then the product rounded again
    g, u = np.array([1.0]), np.array([3.0])
    s = ao.round_to_io(np.array([0.7310585786300049], np.float32), True)[0]         # 0.73046875
    assert lo.silu_mul(g, u, True)[0] == ao.round_to_io(np.array([s * 3.0], np.float32), True)[0]
    # the model-dtype add: 256 + 1 is not representable in bf16 (8 significant bits) -> ties to even 256
    assert lo.add(np.array([256.0]), np.array([1.0]), True)[0] == 256.0
    assert lo.add(np.array([256.0]), np.array([3.0]), True)[0] == 260.0


def test_greedy_takes_the_last_maximum_under_total_order():
    x = np.zeros((4, 10), np.float32)
    x[0, [2, 7]] = 5.0                      # tie -> 7
    x[1, :] = -np.inf
    x[1, 3] = -1e30
    x[2, 4] = np.nan                        # +NaN sorts above +inf (f32::total_cmp)
    x[2, 9] = np.inf
    ids, vals = lo.greedy(x)                # row 3: all equal -> last index
    assert ids.tolist() == [7, 3, 4, 9]
    assert vals[0] == 5.0 and vals[1] == np.float32(-1e30) and np.isnan(vals[2]) and vals[3] == 0.0


def test_linear_rounds_once_from_an_exact_accumulation():
    x = np.array([[1.0, 2.0 ** -9]])
    w = np.array([[1.0, 1.0]])
    assert lo.linear(x, w, exact=True)[0, 0] == 1.0 + 2.0 ** -9
    assert lo.linear(x, w, is_bf16=True)[0, 0] == 1.0          # 1 + 2^-9 rounds to 1 in bf16 (half an ulp, ties to even)
    assert lo.linear(x, w, is_bf16=False)[0, 0] == 1.0 + 2.0 ** -9


def _tiny(world):
    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, vocab_size=96, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64)
    return cfg, [LlamaShard.random_init(cfg, r, world, dtype=torch.float32, device="cpu", seed=5) for r in range(world)]


def _layers(shard, cfg, h, hk):
    out = []
    for ly in shard.layers:
        f = lambda t: t.numpy().astype(np.float64)
        wq, wk, wv, wg, wu = lo.split_fused(f(ly["wqkv"]), f(ly["w_gate_up"]), h, hk, cfg.head_dim)
        out.append(lo.LayerWeights(wq, wk, wv, f(ly["wo"]), wg, wu, f(ly["w_down"]), f(ly["rms_attn"]), f(ly["rms_mlp"])))
    return out


def test_shards_tile_the_full_model_and_the_tp_step_equals_the_unsplit_step():
    cfg, (full,) = _tiny(1)
    _, shards = _tiny(2)
    D, H, HK = cfg.head_dim, 4, 2
    for l in range(2):
        a, (s0, s1) = full.layers[l], (shards[0].layers[l], shards[1].layers[l])
        nq, nkv = H * D, HK * D
        # column-parallel q/k/v: rank r holds the rows of its heads; row-parallel o: the columns of its heads
        assert torch.equal(torch.cat([s0["wqkv"][:nq // 2], s1["wqkv"][:nq // 2]]), a["wqkv"][:nq])
        assert torch.equal(torch.cat([s0["wqkv"][nq // 2:nq // 2 + nkv // 2], s1["wqkv"][nq // 2:nq // 2 + nkv // 2]]),
                           a["wqkv"][nq:nq + nkv])
        assert torch.equal(torch.cat([s0["wo"], s1["wo"]], dim=1), a["wo"])
        i = cfg.intermediate_size
        assert torch.equal(torch.cat([s0["w_gate_up"][:i // 2], s1["w_gate_up"][:i // 2]]), a["w_gate_up"][:i])
        assert torch.equal(torch.cat([s0["w_gate_up"][i // 2:], s1["w_gate_up"][i // 2:]]), a["w_gate_up"][i:])
        assert torch.equal(torch.cat([s0["w_down"], s1["w_down"]], dim=1), a["w_down"])
        assert torch.equal(s0["rms_attn"], a["rms_attn"]) and torch.equal(s1["rms_mlp"], a["rms_mlp"])
    assert torch.equal(shards[1].embed, full.embed) and torch.equal(shards[0].lm_head, full.lm_head)
    # one decode step, fp64 without intermediate roundings: the split model computes the same numbers
    T, page, MB = 3, 16, 2
    rng = np.random.default_rng(1)
    NB = T * MB
    kv = [rng.standard_normal((2, NB, page, HK, D)) for _ in range(2)]
    bt = rng.permutation(NB).astype(np.int32).reshape(T, MB)
    lens = np.array([20, 1, 32])
    pos = lens - 1
    slots = bt[np.arange(T), pos // page].astype(np.int64) * page + pos % page
    hidden = rng.standard_normal((T, cfg.hidden_size))
    cos, sin = (t.numpy().astype(np.float64) for t in rope_tables(cfg, torch.float32, "cpu"))
    args = (full.norm_f.numpy(), full.lm_head.numpy(), cos, sin, pos, slots, bt, lens, H, HK, D, cfg.rms_norm_eps, True)
    kv_a, kv_b = [c.copy() for c in kv], [c.copy() for c in kv]
    xa, la = lo.decode_step(hidden, _layers(full, cfg, H, HK), kv_a, *args, exact=True)
    xb, lb = lo.decode_step(hidden, _layers(full, cfg, H, HK), kv_b, *args, exact=True,
                            tp_shards=[_layers(s, cfg, H // 2, HK // 2) for s in shards])
    np.testing.assert_allclose(lb, la, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(xb, xa, rtol=1e-9, atol=1e-9)
    for a, b in zip(kv_a, kv_b):
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-12)
    # a prefix model shares the first layers and the replicated tensors with the full one
    pre = LlamaShard.random_init(cfg, dtype=torch.float32, device="cpu", seed=5, num_layers=1)
    assert torch.equal(pre.layers[0]["w_down"], full.layers[0]["w_down"]) and torch.equal(pre.embed, full.embed)


def test_rope_tables_follow_cache_new():
    """models/src/llama.rs:145-199: f32 inverse frequencies, f32 outer product, cos/sin cast to the model dtype;
    Llama-3 scaling leaves high frequencies alone, divides low ones by `factor` and blends in between."""
    cfg = LlamaConfig(256, 512, 10, 1, 4, 2, rope_theta=10000.0, max_position_embeddings=32)
    cos, sin = rope_tables(cfg, torch.float32, "cpu")
    d = cfg.head_dim
    inv = (1.0 / (np.float32(10000.0) ** (np.arange(0, d, 2, dtype=np.float32) / np.float32(d)))).astype(np.float32)
    ang = np.arange(32, dtype=np.float32)[:, None] * inv[None, :]
    np.testing.assert_allclose(cos.numpy(), np.cos(ang), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(sin.numpy(), np.sin(ang), rtol=2e-6, atol=2e-6)
    assert cos.shape == (32, d // 2)
    # the oracle's own table builder agrees (same recipe in numpy)
    c2, s2 = ao.rope_tables(d, 32, theta=10000.0, is_bf16=True)
    cb, sb = rope_tables(cfg, torch.bfloat16, "cpu")
    np.testing.assert_array_equal(cb.float().numpy(), c2)
    scaled = LlamaConfig(256, 512, 10, 1, 4, 2, rope_theta=500000.0, max_position_embeddings=32,
                         rope_scaling=Llama3RopeScaling(factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                                        original_max_position_embeddings=8192))
    cs, _ = rope_tables(scaled, torch.float32, "cpu")
    plain = LlamaConfig(256, 512, 10, 1, 4, 2, rope_theta=500000.0, max_position_embeddings=32)
    cp, _ = rope_tables(plain, torch.float32, "cpu")
    assert torch.equal(cs[:, 0], cp[:, 0])                      # highest frequency: wavelength << 8192/4, untouched
    inv_p = 1.0 / (500000.0 ** (np.arange(0, d, 2) / d))
    low = 2 * np.pi / inv_p > 8192.0                            # wavelength beyond the original context: / factor
    assert low.any()
    j = int(np.argmax(low))
    np.testing.assert_allclose(cs[:, j].numpy(), np.cos(np.arange(32) * inv_p[j] / 8.0), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# Pin against an independent implementation: Hugging Face transformers' Llama in float64
# (tests/golden/make_llama_golden.py -> tests/golden/llama_decode_hf.npz)
# ------------------------------------------------------------------------------------------------
def _hf_fixture():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "llama_decode_hf.npz")


def test_decode_step_oracle_reproduces_hf_llama_logits():
    """One decode step of the oracle (exact = fp64, no model-dtype roundings) on the KV cache Hugging Face's Llama
    built for the first n-1 tokens of each sequence reproduces HF's logits for token n: non-interleaved RoPE with
    Llama-3 frequency scaling, RMSNorm, GQA attention over the paged cache, SwiGLU MLP, residual wiring, final norm
    and lm_head — to 1e-6 (HF evaluates the RMSNorm statistics, the softmax and the rotary tables in float32 even in a
    float64 model; a bf16 rounding anywhere would show as 4e-3).  Sequences of 37, 1, 100 and 64 tokens (a one-token
    sequence: empty cache)."""
    fx = _hf_fixture()
    hid, inter, vocab, L, H, HK, D = [int(v) for v in fx["cfg"]]
    lens = [int(v) for v in fx["lens"]]
    T, page = len(lens), 16
    w = lambda n: np.asarray(fx["w/" + n], np.float64)
    layers = [lo.LayerWeights(w(f"model.layers.{l}.self_attn.q_proj.weight"), w(f"model.layers.{l}.self_attn.k_proj.weight"),
                              w(f"model.layers.{l}.self_attn.v_proj.weight"), w(f"model.layers.{l}.self_attn.o_proj.weight"),
                              w(f"model.layers.{l}.mlp.gate_proj.weight"), w(f"model.layers.{l}.mlp.up_proj.weight"),
                              w(f"model.layers.{l}.mlp.down_proj.weight"), w(f"model.layers.{l}.input_layernorm.weight"),
                              w(f"model.layers.{l}.post_attention_layernorm.weight")) for l in range(L)]
    # paged caches holding HF's K (post-RoPE) and V of every token but the last, blocks in a shuffled order
    mb = (max(lens) + page - 1) // page
    nb = T * mb + 3
    rng = np.random.default_rng(1)
    bt = rng.permutation(nb)[: T * mb].astype(np.int32).reshape(T, mb)
    caches = [np.zeros((2, nb, page, HK, D), np.float64) for _ in range(L)]
    for b, n in enumerate(lens):
        for l in range(L):
            k, v = fx[f"k_{b}_{l}"], fx[f"v_{b}_{l}"]            # [Hkv, n, D]
            for t in range(n - 1):
                caches[l][0][bt[b, t // page], t % page] = k[:, t]
                caches[l][1][bt[b, t // page], t % page] = v[:, t]
    pos = np.asarray(lens) - 1
    slots = np.asarray([bt[b, p // page] * page + p % page for b, p in enumerate(pos)], np.int64)
    hidden = np.stack([w("model.embed_tokens.weight")[int(fx[f"tokens_{b}"][-1])] for b in range(T)])
    _, logits = lo.decode_step(hidden, layers, caches, w("model.norm.weight"), w("lm_head.weight"), fx["rope_cos"], fx["rope_sin"],
                               pos, slots, bt, np.asarray(lens), H, HK, D, 1e-5, True, exact=True)
    for b in range(T):
        want = fx[f"logits_{b}"]
        err = np.abs(logits[b] - want).max() / np.abs(want).max()
        assert err < 1e-6, f"sequence {b} (len {lens[b]}): {err:.3e}"
    # the step also wrote the decoded token's K and V where the slot mapping says — HF's K/V of the last token
    for b, n in enumerate(lens):
        for l in range(L):
            blk, off = bt[b, (n - 1) // page], (n - 1) % page
            assert np.abs(caches[l][0][blk, off] - fx[f"k_{b}_{l}"][:, n - 1]).max() < 1e-5
            assert np.abs(caches[l][1][blk, off] - fx[f"v_{b}_{l}"][:, n - 1]).max() < 1e-5


def test_rope_tables_with_llama3_scaling_match_hf():
    """The product's rotary tables (atoma_infer_b200.llama.rope_tables ≙ Cache::new, llama.rs:145-199, Llama-3 frequency
    scaling) against Hugging Face's LlamaRotaryEmbedding for the same configuration (f32 tables vs f64: 1e-6)."""
    import torch
    from atoma_infer_b200.llama import Llama3RopeScaling, LlamaConfig, rope_tables
    fx = _hf_fixture()
    cfg = LlamaConfig(hidden_size=64, intermediate_size=160, vocab_size=97, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, rope_theta=500000.0, max_position_embeddings=256,
                      rope_scaling=Llama3RopeScaling(8.0, 1.0, 4.0, 64))
    cos, sin = rope_tables(cfg, torch.float32, "cpu", max_pos=128)
    assert np.abs(cos.numpy() - fx["rope_cos"]).max() < 2e-6 and np.abs(sin.numpy() - fx["rope_sin"]).max() < 2e-6


def test_vocab_parallel_greedy_combination_rule():
    """What atoma_sample_greedy_tp computes, restated: every rank takes the greedy token of its vocabulary slice (LAST
    index among equal maxima under f32 total order), the candidates are merged in rank order with 'greater key, or equal
    key and greater index, wins' — equal to the greedy token of the full row, ties across slices and NaNs included."""
    rng = np.random.default_rng(11)
    vocab, T = 4096, 16
    logits = rng.standard_normal((T, vocab)).astype(np.float32)
    logits[0, [5, 700, 4000]] = 9.0            # a three-way tie across different slices: the last index wins
    logits[1, :] = -1.0                        # all equal
    logits[2, 100] = np.inf
    logits[3, 77] = np.nan                     # total order: +NaN above +inf
    logits[3, 3000] = np.inf
    logits[4, 2047:2049] = 5.0                 # a tie straddling a slice boundary
    want_tok, want_val = lo.greedy(logits)
    for world in (2, 4, 8):
        rows = vocab // world
        best_key = np.full(T, np.iinfo(np.int64).min)
        best_idx = np.zeros(T, np.int64)
        for r in range(world):
            tok, _ = lo.greedy(logits[:, r * rows:(r + 1) * rows])
            key = logits[np.arange(T), r * rows + tok].view(np.int32).astype(np.int64)
            key = np.where(key < 0, key ^ 0x7FFFFFFF, key)
            idx = r * rows + tok
            take = (r == 0) | (key > best_key) | ((key == best_key) & (idx > best_idx))
            best_key = np.where(take, key, best_key)
            best_idx = np.where(take, idx, best_idx)
        assert np.array_equal(best_idx, want_tok), world
        got_val = logits[np.arange(T), best_idx]
        assert np.array_equal(got_val.view(np.int32), want_val.view(np.int32))
