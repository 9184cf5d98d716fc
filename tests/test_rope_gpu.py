"""GPU parity of the fused rotary-embedding + cache-write kernel (atoma_rope_reshape_and_cache,
SURVEY §8 row f1) against the numpy oracle: rotated values within one rounding of the fp64 result
(fp32 arithmetic, single rounding to the io type), cache placement bit-exact."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from helpers import randn_t, t2f32, ulp16
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("T,H,HK,D,dtype,fused_layout", [
    (32, 32, 8, 128, torch.bfloat16, True),     # decode step, q/k/v as slices of one fused qkv buffer
    (1000, 32, 8, 128, torch.bfloat16, False),  # prefill-sized
    (37, 32, 8, 64, torch.float16, True),
    (5, 8, 1, 128, torch.bfloat16, False),
])
def test_rope_and_cache_vs_oracle(T, H, HK, D, dtype, fused_layout):
    is_bf16 = dtype == torch.bfloat16
    BS, NB, max_pos = 16, (T + 15) // 16 + 3, 4096
    if fused_layout:
        qkv = randn_t((T, (H + 2 * HK) * D), 1, dtype, DEV)
        q = qkv[:, : H * D].unflatten(1, (H, D))
        k = qkv[:, H * D:(H + HK) * D].unflatten(1, (HK, D))
        v = qkv[:, (H + HK) * D:].unflatten(1, (HK, D))
    else:
        q, k, v = randn_t((T, H, D), 1, dtype, DEV), randn_t((T, HK, D), 2, dtype, DEV), randn_t((T, HK, D), 3, dtype, DEV)
    q0, k0, v0 = t2f32(q).copy(), t2f32(k).copy(), v.clone()
    cos_np, sin_np = orc.rope_tables(D, max_pos, is_bf16=is_bf16)
    cos, sin = torch.from_numpy(cos_np).to(dtype).to(DEV), torch.from_numpy(sin_np).to(dtype).to(DEV)
    rng = np.random.default_rng(T)
    pos_np = rng.integers(0, max_pos, size=T)
    slots_np = rng.permutation(NB * BS)[:T].astype(np.int64)
    slots_np[rng.random(T) < 0.15] = -1
    pos, slots = torch.from_numpy(pos_np).to(DEV), torch.from_numpy(slots_np).to(DEV)
    cache = randn_t((2, NB, BS, HK, D), 5, dtype, DEV)
    cache0 = cache.clone()
    ops.rope_reshape_and_cache(q, k, v, cache[0], cache[1], pos, cos, sin, slots)
    torch.cuda.synchronize()
    for name, got, x0 in (("q", q, q0), ("k", k, k0)):
        want = orc.rope(x0, pos_np, cos_np, sin_np)
        g = t2f32(got).astype(np.float64)
        # fp32 arithmetic + one rounding: within 0.5 ulp of the exact value + fp32 evaluation error
        bound = 0.5 * ulp16(want, is_bf16) + 2e-6 * (np.abs(x0).max(-1, keepdims=True) + 1)
        assert (np.abs(g - want) <= bound).all(), f"{name}: max excess {(np.abs(g - want) - bound).max():.3e}"
    assert torch.equal(v, v0), "v must not be modified"
    want_cache = cache0.clone()
    for t, s in enumerate(slots_np):
        if s >= 0:
            want_cache[0, s // BS, s % BS] = k[t]     # the ROTATED k, bit for bit
            want_cache[1, s // BS, s % BS] = v0[t]
    assert torch.equal(cache, want_cache)


def test_rope_only_and_errors():
    T, H, HK, D = 4, 4, 2, 64
    q, k, v = randn_t((T, H, D), 1, torch.bfloat16, DEV), randn_t((T, HK, D), 2, torch.bfloat16, DEV), randn_t((T, HK, D), 3, torch.bfloat16, DEV)
    cos_np, sin_np = orc.rope_tables(D, 16)
    cos, sin = torch.from_numpy(cos_np).to(torch.bfloat16).to(DEV), torch.from_numpy(sin_np).to(torch.bfloat16).to(DEV)
    pos = torch.zeros(T, dtype=torch.int64, device=DEV)
    q0 = q.clone()
    ops.rope_reshape_and_cache(q, k, v, None, None, pos, cos, sin, None)   # position 0: identity, no cache
    assert torch.equal(q, q0)
    with pytest.raises(ops.AtomaError):
        ops.rope_reshape_and_cache(q, k, v, None, None, pos.to(torch.int32), cos, sin, None)
    with pytest.raises(ops.AtomaError):
        ops.rope_reshape_and_cache(q, k, v, None, None, pos, cos[:, :8], sin[:, :8], None)
