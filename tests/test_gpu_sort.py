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
