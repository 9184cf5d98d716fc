"""Golden vectors for the decode-step oracle (oracle/llama_oracle.py) from an INDEPENDENT implementation of the same
published algorithms: Hugging Face transformers' Llama (the architecture the reference's candle model is a port of —
models/src/llama.rs follows candle-transformers' llama, itself a port of HF's modeling_llama).  The reference's own
ops for RoPE / RMSNorm / SwiGLU / Linear are candle calls that cannot be built or run here (registry crates, no Rust
toolchain), and the reference asserts no values for them; this fixture pins the oracle's restatement of those ops —
non-interleaved RoPE with Llama-3 frequency scaling (llama.rs:145-251), RMSNorm (llama.rs:398-410), SwiGLU MLP
(llama.rs:362-366), GQA attention over a KV cache, residual wiring (llama.rs:398-410), final norm + lm_head
(llama.rs:474-477) — in float64, so that agreement is limited only by rounding of the arithmetic order.

Run here (transformers is importable in the build image; not needed on the GPU box):
    python tests/golden/make_llama_golden.py        -> tests/golden/llama_decode_hf.npz (committed)
"""
from pathlib import Path

import numpy as np
import torch
from transformers import LlamaConfig, LlamaForCausalLM

OUT = Path(__file__).resolve().parent / "llama_decode_hf.npz"


def main():
    torch.manual_seed(20261017)
    cfg = LlamaConfig(vocab_size=97, hidden_size=64, intermediate_size=160, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=16, rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=256,
                      rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                                    "original_max_position_embeddings": 64},
                      tie_word_embeddings=False, attention_bias=False, mlp_bias=False)
    model = LlamaForCausalLM(cfg).double().eval()
    with torch.no_grad():
        for p in model.parameters():            # away from the initialiser's tiny scale: make every op matter
            p.copy_(torch.randn_like(p) * (0.6 if p.dim() > 1 else 1.0) / (p.shape[-1] ** 0.5 if p.dim() > 1 else 1.0))
        for name, p in model.named_parameters():
            if "norm" in name:
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
    lens = [37, 1, 100, 64]                      # tokens per sequence INCLUDING the one being decoded
    out = {"lens": np.asarray(lens)}
    rng = np.random.default_rng(5)
    tokens = [rng.integers(0, cfg.vocab_size, size=n) for n in lens]
    L, HK, D = cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.head_dim
    for b, tok in enumerate(tokens):
        ids = torch.tensor(tok[None, :], dtype=torch.long)
        with torch.no_grad():
            full = model(input_ids=ids, use_cache=True)
        out[f"tokens_{b}"] = tok
        out[f"logits_{b}"] = full.logits[0, -1].numpy()                 # what the decode step of the last token produces
        pkv = full.past_key_values
        for l in range(L):
            layer = pkv.layers[l] if hasattr(pkv, "layers") else None
            k = layer.keys if layer is not None else pkv[l][0]
            v = layer.values if layer is not None else pkv[l][1]
            out[f"k_{b}_{l}"] = k[0].numpy()                             # [Hkv, n, D], post-RoPE, all n tokens
            out[f"v_{b}_{l}"] = v[0].numpy()
    sd = model.state_dict()
    for name, t in sd.items():
        out["w/" + name] = t.numpy()
    # the rotary tables HF uses (cos / sin of position * scaled inv_freq; attention_scaling is 1 for llama3)
    pos = torch.arange(0, 128)[None, :]
    cos, sin = model.model.rotary_emb(torch.zeros(1, 128, cfg.hidden_size, dtype=torch.float64), pos)
    out["rope_cos"] = cos[0, :, : D // 2].numpy()
    out["rope_sin"] = sin[0, :, : D // 2].numpy()
    out["cfg"] = np.asarray([cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, L, cfg.num_attention_heads, HK, D])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
