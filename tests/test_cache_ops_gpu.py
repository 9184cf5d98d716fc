"""GPU tests for reshape_and_cache_flash / copy_blocks / swap_blocks through the C ABI: byte-exact
against the CPU oracle, mirroring csrc/tests/cache_manager_tests.rs."""
import numpy as np
import pytest
import torch

from atoma_infer_b200 import ops
from atoma_infer_b200.flash_attention import FlashAttention
from helpers import bits, randn_t
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_reshape_and_cache_flash_identity(dtype):
    """cache_manager_tests.rs:552-684"""
    T, H, D, NB, BS = 10, 2, 8, 4, 8
    key, value = randn_t((T, H, D), 1, dtype, DEV), randn_t((T, H, D), 2, dtype, DEV)
    kc = torch.zeros((NB, BS, H, D), dtype=dtype, device=DEV)
    vc = torch.zeros_like(kc)
    ops.reshape_and_cache_flash(key, value, kc, vc, torch.arange(T, dtype=torch.int64, device=DEV))
    assert torch.equal(kc.view(-1, H, D)[:T], key) and torch.equal(vc.view(-1, H, D)[:T], value)
    assert torch.count_nonzero(kc.view(-1, H, D)[T:]) == 0


@pytest.mark.parametrize("T,H,D,BS,NB,dtype", [
    (32, 8, 128, 16, 64, torch.bfloat16),      # config 2 decode step
    (4096, 8, 128, 16, 300, torch.bfloat16),   # prefill-sized scatter
    (37, 8, 64, 16, 8, torch.float16),
    (5, 3, 40, 48, 3, torch.bfloat16),
    (9, 1, 5, 16, 2, torch.float16),           # rows that are not 16-byte multiples -> scalar path
])
def test_reshape_and_cache_flash_vs_oracle(T, H, D, BS, NB, dtype):
    rng = np.random.default_rng(T)
    key, value = randn_t((T, H, D), 3, dtype, DEV), randn_t((T, H, D), 4, dtype, DEV)
    cache = randn_t((2, NB, BS, H, D), 5, dtype, DEV)
    slots = rng.permutation(NB * BS)[:T].astype(np.int64)
    slots[rng.random(T) < 0.2] = -1  # padded tokens (cache_manager.cu:151-155)
    kc_np, vc_np = bits(cache[0]).ravel().copy(), bits(cache[1]).ravel().copy()
    orc.reshape_and_cache_flash(bits(key).ravel(), bits(value).ravel(), kc_np, vc_np, slots, BS * H * D, T, H, D,
                                BS, H * D, H * D)
    ops.reshape_and_cache_flash(key, value, cache[0], cache[1], torch.from_numpy(slots).to(DEV))
    np.testing.assert_array_equal(bits(cache[0]).ravel(), kc_np)
    np.testing.assert_array_equal(bits(cache[1]).ravel(), vc_np)


def test_reshape_and_cache_flash_strided_source():
    """k/v are slices of a fused qkv projection in the model: token stride > num_heads * head_size."""
    T, H, D, BS, NB = 12, 2, 64, 16, 4
    qkv = randn_t((T, 6 * D), 7, torch.bfloat16, DEV)
    key = qkv[:, 2 * D:4 * D].view(T, H, D)
    value = qkv[:, 4 * D:].view(T, H, D)
    cache = torch.zeros((2, NB, BS, H, D), dtype=torch.bfloat16, device=DEV)
    slots = torch.arange(T, dtype=torch.int64, device=DEV) + 3
    ops.reshape_and_cache_flash(key, value, cache[0], cache[1], slots)
    assert torch.equal(cache[0].view(-1, H, D)[3:3 + T], key)
    assert torch.equal(cache[1].view(-1, H, D)[3:3 + T], value)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_copy_blocks_reference_fixture(dtype):
    """cache_manager_tests.rs:241-454: mapping [[0,2],[1,3]], 2 layers, 4 blocks of 64 x 2 x 8."""
    caches = [randn_t((2, 4, 64, 2, 8), 10 + l, dtype, DEV) for l in range(2)]
    before = [c.clone() for c in caches]
    bm = torch.tensor([[0, 2], [1, 3]], dtype=torch.int64, device=DEV)
    FlashAttention.copy_blocks(caches, bm)
    for c, b in zip(caches, before):
        for kvi in range(2):
            assert torch.equal(c[kvi, 2], b[kvi, 0]) and torch.equal(c[kvi, 3], b[kvi, 1])
            assert torch.equal(c[kvi, :2], b[kvi, :2])


@pytest.mark.parametrize("L,NB,BS,H,D,P", [(32, 40, 16, 8, 128, 16), (3, 9, 16, 8, 64, 4), (2, 6, 16, 1, 24, 3),
                                           (2, 5, 16, 3, 5, 2), (1, 4, 256, 8, 128, 2)])
def test_copy_blocks_vs_oracle(L, NB, BS, H, D, P):
    rng = np.random.default_rng(L * 100 + P)
    caches = [randn_t((2, NB, BS, H, D), 20 + l, torch.bfloat16, DEV) for l in range(L)]
    perm = rng.permutation(NB)
    pairs = np.stack([perm[:P], perm[P:2 * P]], axis=1).astype(np.int64)  # disjoint src / dst
    k_np = [bits(c[0]).ravel().copy() for c in caches]
    v_np = [bits(c[1]).ravel().copy() for c in caches]
    orc.copy_blocks(k_np, v_np, pairs, BS * H * D)
    ops.copy_blocks([c[0] for c in caches], [c[1] for c in caches], torch.from_numpy(pairs).to(DEV))
    for l in range(L):
        np.testing.assert_array_equal(bits(caches[l][0]).ravel(), k_np[l])
        np.testing.assert_array_equal(bits(caches[l][1]).ravel(), v_np[l])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("direction", ["d2d", "h2d", "d2h"])
def test_swap_blocks(direction, dtype):
    """cache_manager_tests.rs:64-186: mapped blocks equal, unmapped blocks untouched."""
    NB, BS, H, D = 12, 16, 8, 128
    src = randn_t((NB, BS, H, D), 30, dtype, "cpu")
    dst = randn_t((NB, BS, H, D), 31, dtype, "cpu")
    if direction in ("d2d", "d2h"):
        src = src.to(DEV)
    else:
        src = src.pin_memory()
    if direction in ("d2d", "h2d"):
        dst = dst.to(DEV)
    else:
        dst = dst.pin_memory()
    before = dst.clone()
    mapping = {0: 5, 1: 6, 2: 7, 9: 0, 4: 11}
    ops.swap_blocks(src, dst, mapping)
    torch.cuda.synchronize()
    for s, d in mapping.items():
        assert torch.equal(dst[d].cpu(), src[s].cpu())
    for d in set(range(NB)) - set(mapping.values()):
        assert torch.equal(dst[d].cpu(), before[d].cpu())


def test_swap_blocks_pageable_host_memory():
    """Unpinned host tensors (what candle's CPU storage is, ops.rs:158-164) take the memcpy path."""
    NB, BS, H, D = 10, 16, 2, 64
    gpu = randn_t((NB, BS, H, D), 50, torch.bfloat16, DEV)
    cpu = torch.zeros((NB, BS, H, D), dtype=torch.bfloat16)
    assert not cpu.is_pinned()
    ops.swap_blocks(gpu, cpu, {0: 3, 1: 4, 2: 5, 7: 0})
    for s_, d_ in {0: 3, 1: 4, 2: 5, 7: 0}.items():
        assert torch.equal(cpu[d_], gpu[s_].cpu())
    back = torch.zeros_like(gpu)
    ops.swap_blocks(cpu, back, {3: 9, 4: 8})
    torch.cuda.synchronize()
    assert torch.equal(back[9], gpu[0]) and torch.equal(back[8], gpu[1])


def test_swap_round_trip_full_sequence():
    """Swap one 4096-token sequence (256 blocks, Llama-3.1-8B geometry) out to pinned host memory and
    back into different blocks: encode -> erase -> decode round trip, byte exact."""
    NB, BS, H, D = 600, 16, 8, 128
    fa = FlashAttention(32, 8, 128, 0.088, device=DEV)
    gpu = randn_t((2, NB, BS, H, D), 40, torch.bfloat16, DEV)
    cpu = torch.zeros((2, 300, BS, H, D), dtype=torch.bfloat16).pin_memory()
    rng = np.random.default_rng(4)
    blocks = rng.permutation(NB)[:256].tolist()
    out_map = {b: i for i, b in enumerate(blocks)}
    original = gpu[:, blocks].clone()
    fa.swap_blocks(gpu, cpu, out_map)
    gpu[:, blocks] = 0
    new_blocks = rng.permutation(NB)[:256].tolist()
    fa.swap_blocks(cpu, gpu, {i: b for i, b in enumerate(new_blocks)})
    torch.cuda.synchronize()
    assert torch.equal(gpu[:, new_blocks], original)
