"""Onesweep radix sort (gsb_sort_pairs) against the oracle's stable LSD radix (sort.comp restated)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def run_sort(gs, ctx, keys, vals, key_bits):
    dev = torch.device("cuda:0")
    k = torch.from_numpy(keys.view(np.int64)).to(dev)
    v = torch.from_numpy(vals.view(np.int32)).to(dev)
    kt, vt = torch.empty_like(k), torch.empty_like(v)
    torch.cuda.synchronize()
    ctx.sort_pairs(k.data_ptr(), v.data_ptr(), kt.data_ptr(), vt.data_ptr(), keys.size, key_bits)
    torch.cuda.synchronize()
    return k.cpu().numpy().view(np.uint64), v.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("m", [1, 2, 31, 4095, 4096, 4097, 8192, 100_003, 1_000_000])
@pytest.mark.parametrize("key_bits", [64, 47, 40, 33, 8])
def test_sort_random(gs, oracle, ctx, m, key_bits):
    rng = np.random.default_rng(m * 131 + key_bits)
    keys = rng.integers(0, 2**63, size=m, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=m, dtype=np.uint64)
    if key_bits < 64:
        keys &= np.uint64((1 << key_bits) - 1)
    vals = np.arange(m, dtype=np.uint32)
    ks, vs = run_sort(gs, ctx, keys, vals, key_bits)
    rk, rv = oracle.sort_pairs(keys, vals)
    assert np.array_equal(ks, rk)
    assert np.array_equal(vs, rv)


def test_sort_stability_with_heavy_duplicates(gs, oracle, ctx):
    rng = np.random.default_rng(7)
    m = 300_000
    keys = (rng.integers(0, 5, size=m, dtype=np.uint64) << np.uint64(32)) | rng.integers(0, 3, size=m, dtype=np.uint64)
    vals = np.arange(m, dtype=np.uint32)
    ks, vs = run_sort(gs, ctx, keys, vals, 35)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ks, keys[order])
    assert np.array_equal(vs, vals[order])  # equal keys keep input order


def test_sort_constant_and_presorted(gs, oracle, ctx):
    m = 50_000
    vals = np.arange(m, dtype=np.uint32)
    for keys in (np.full(m, 0x1234_5678_9ABC, np.uint64), np.arange(m, dtype=np.uint64) * 977,
                 (np.arange(m, dtype=np.uint64)[::-1] * 31).copy()):
        ks, vs = run_sort(gs, ctx, keys, vals, 48)
        rk, rv = oracle.sort_pairs(keys, vals)
        assert np.array_equal(ks, rk) and np.array_equal(vs, rv)


def run_sort32(gs, ctx, keys, vals, key_bits):
    dev = torch.device("cuda:0")
    k = torch.from_numpy(keys.view(np.int32)).to(dev)
    v = torch.from_numpy(vals.view(np.int32)).to(dev)
    kt, vt = torch.empty_like(k), torch.empty_like(v)
    torch.cuda.synchronize()
    ctx.sort_pairs32(k.data_ptr(), v.data_ptr(), kt.data_ptr(), vt.data_ptr(), keys.size, key_bits)
    torch.cuda.synchronize()
    return k.cpu().numpy().view(np.uint32), v.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("m", [1, 33, 4095, 4096, 4097, 12_289, 250_001, 3_000_000])
@pytest.mark.parametrize("key_bits", [32, 24, 15, 8, 3])
def test_sort32_random(gs, ctx, m, key_bits):
    """u32 keys: the instantiation the frame runs (depth bits / tile ids) with the atomicOr digit matching."""
    rng = np.random.default_rng(m * 17 + key_bits)
    keys = rng.integers(0, 2**32, size=m, dtype=np.uint64).astype(np.uint32)
    if key_bits < 32:
        keys &= np.uint32((1 << key_bits) - 1)
    vals = np.arange(m, dtype=np.uint32)
    ks, vs = run_sort32(gs, ctx, keys, vals, key_bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ks, keys[order])
    assert np.array_equal(vs, vals[order])  # stable: equal keys keep input order


def test_sort32_skewed_digits(gs, ctx):
    """Whole warps sharing one digit, runs of equal keys (a Gaussian's instances), one hot tile id."""
    m = 700_000
    rng = np.random.default_rng(3)
    runs = np.repeat(rng.integers(0, 17_600, size=m // 7, dtype=np.uint64).astype(np.uint32), 7)[:m]
    hot = np.where(rng.random(m) < 0.6, np.uint32(4242), runs).astype(np.uint32)
    for keys, bits in ((runs, 15), (hot, 15), (np.zeros(m, np.uint32), 15), (np.full(m, 0xFFFFFFFF, np.uint32), 32)):
        vals = rng.permutation(m).astype(np.uint32)
        ks, vs = run_sort32(gs, ctx, keys.copy(), vals, bits)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ks, keys[order]) and np.array_equal(vs, vals[order])
