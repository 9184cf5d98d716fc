"""CPU checks of the rotary-embedding oracle (oracle.attention_oracle.rope / rope_tables): the
reference pins no values for this op (candle-nn is not vendored), so the restatement is checked
against the definition it cites — a rotation of each (x[i], x[i + d/2]) pair by pos * theta^(-2i/d)
(models/src/llama.rs:145-199, 218-251)."""
import numpy as np

from oracle import attention_oracle as orc


def test_rope_is_a_rotation_of_half_split_pairs():
    rng = np.random.default_rng(0)
    T, H, D = 7, 3, 16
    x = rng.standard_normal((T, H, D))
    pos = np.array([0, 1, 5, 100, 4095, 17, 2])
    inv = 1.0 / 500000.0 ** (np.arange(0, D, 2) / D)
    ang = pos[:, None] * inv[None, :]
    got = orc.rope(x, np.arange(T), np.cos(ang), np.sin(ang))      # tables already gathered per token
    z = (x[..., : D // 2] + 1j * x[..., D // 2:]) * np.exp(1j * ang)[:, None, :]
    np.testing.assert_allclose(got[..., : D // 2], z.real, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got[..., D // 2:], z.imag, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(got, axis=-1), np.linalg.norm(x, axis=-1), rtol=1e-12)
    np.testing.assert_array_equal(got[0], x[0])                    # position 0 is the identity


def test_rope_tables_match_the_reference_recipe():
    cos, sin = orc.rope_tables(128, 64, theta=500000.0, is_bf16=True)
    assert cos.shape == (64, 64) and sin.shape == (64, 64)
    assert np.all(cos[0] == 1.0) and np.all(sin[0] == 0.0)
    inv = (1.0 / np.float32(500000.0) ** (np.arange(0, 128, 2, dtype=np.float32) / np.float32(128))).astype(np.float32)
    want = np.cos(np.float32(3) * inv).astype(np.float32)
    np.testing.assert_array_equal(cos[3], orc.round_to_io(want, True))
    # dot products of rotated q and k depend only on the position difference (what RoPE is for)
    rng = np.random.default_rng(1)
    q, k = rng.standard_normal((1, 1, 128)), rng.standard_normal((1, 1, 128))
    c64, s64 = orc.rope_tables(128, 64, is_bf16=False)
    ang = np.arange(64, dtype=np.float64)[:, None] * inv.astype(np.float64)[None, :]
    c, s_ = np.cos(ang), np.sin(ang)
    d1 = (orc.rope(q, [10], c, s_) * orc.rope(k, [4], c, s_)).sum()
    d2 = (orc.rope(q, [36], c, s_) * orc.rope(k, [30], c, s_)).sum()
    np.testing.assert_allclose(d1, d2, rtol=1e-9)
