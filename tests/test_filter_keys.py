"""The order-preserving keys of the filter_matches kernels, restated in numpy (csrc/lg_small.cu fm_key / fm_pack_key,
csrc/lg_common.cuh fm_pack): a 64-bit unsigned max over (key << 32 | ~index) must pick what torch.max picks -- the
largest value, the lowest index among equal values, NaN above everything (filter_matches, lightglue.py:294-319)."""
import numpy as np
import torch


def fm_key(v):  # lg_small.cu: signed key, every NaN on top
    u = v.view(np.int32)
    k = u ^ ((u >> 31) & np.int32(0x7FFFFFFF))
    return np.where(np.isnan(v), np.int32(0x7FFFFFFF), k)


def fm_pack_key(key, idx):  # lg_small.cu
    hi = (key.astype(np.int64) & 0xFFFFFFFF) ^ 0x80000000
    return (hi.astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))


def fm_pack(v, idx):  # lg_common.cuh (the epilogue of the bf16 assignment kernel packs with this one)
    u = v.view(np.uint32).astype(np.uint64)
    key = np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
    key = np.where(np.isnan(v), np.uint64(0xFFFFFFFF), key)
    return (key << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.uint64))


def _specials():
    return np.array([-np.inf, -3.4e38, -1.5, -1e-30, -1e-45, 1e-45, 1e-30, 0.5, 1.0, 1.0000001, 3.4e38, np.inf],
                    dtype=np.float32)


def test_signed_key_is_monotonic_and_nan_is_on_top():
    v = _specials()
    k = fm_key(v)
    assert (np.diff(k.astype(np.int64)) > 0).all()
    rng = np.random.default_rng(0)
    r = rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-20, 20, 20000).astype(np.float32)
    order = np.argsort(r, kind="stable")
    assert (np.diff(fm_key(r)[order].astype(np.int64)) >= 0).all()
    for nan in (np.float32(np.nan), np.array([0xFFC00000], dtype=np.uint32).view(np.float32)[0]):  # either sign bit
        assert fm_key(np.array([nan], dtype=np.float32))[0] == 0x7FFFFFFF
    assert fm_key(np.array([np.inf], dtype=np.float32))[0] < 0x7FFFFFFF
    assert fm_key(np.array([-np.inf], dtype=np.float32))[0] > np.int32(-0x80000000)  # the kernels' "nothing yet" key


def test_both_packings_agree_and_pick_torch_max():
    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(1, 70))
        v = rng.choice(np.concatenate([_specials(), rng.standard_normal(8).astype(np.float32)]), size=n).astype(np.float32)
        if trial % 3 == 0:
            v[rng.integers(0, n)] = np.nan
        idx = np.arange(n)
        a, b = fm_pack_key(fm_key(v), idx), fm_pack(v, idx)
        assert (a == b).all()
        best = int(np.argmax(a))  # unsigned 64-bit max
        want = torch.from_numpy(v).max(0)
        assert best == int(want.indices), (v, best, int(want.indices))


def test_known_deviation_signed_zeros():
    """-0.0 sorts below +0.0 in both packings whereas torch.max treats them as equal (first index wins): documented
    in lg_common.cuh; scores are sums of log-probabilities, an exact -0.0 / +0.0 pair of maxima does not occur."""
    v = np.array([-0.0, 0.0], dtype=np.float32)
    assert fm_key(v)[0] < fm_key(v)[1]
