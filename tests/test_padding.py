"""Padding producers with counts: same tensors as the reference's
`pad_to_length` / `pad_and_stack` / `pad_local_features` (models/utils/misc.py:19-62,103-113;
models/cache_loader.py:17-45), plus the counts the B200 matcher masks with."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from glue_factory_colon_b200 import padding  # noqa: E402

REF_MISC = Path("/root/reference/gluefactory/models/utils/misc.py")


def _feat(n, seed, with_extras=True):
    g = torch.Generator().manual_seed(seed)
    f = {
        "keypoints": torch.rand(n, 2, generator=g) * torch.tensor([640.0, 480.0]),
        "descriptors": torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=-1),
        "keypoint_scores": torch.rand(n, generator=g),
    }
    if with_extras:
        f.update(scales=torch.rand(n, generator=g), oris=torch.rand(n, generator=g),
                 sparse_depth=torch.rand(n, generator=g), point3D_ids=torch.arange(n),
                 valid_depth_mask=torch.ones(n, dtype=torch.bool))
    return f


@pytest.mark.parametrize("mode,fill", [("zeros", 0), ("ones", 1), ("minus_one", -1), (False, 0), ("ZEROS", 0)])
def test_constant_modes_keep_the_prefix_dtype_and_count(mode, fill):
    x = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    y, cnt = padding.pad_to_length(x, 10, -2, mode=mode, return_count=True)
    assert cnt == 6 and y.shape == (10, 2) and y.dtype == x.dtype
    assert torch.equal(y[:6], x) and (y[6:] == fill).all()
    ids = torch.arange(5)
    yi = padding.pad_to_length(ids, 8, -1, mode="minus_one")
    assert yi.dtype == torch.int64 and yi.tolist() == [0, 1, 2, 3, 4, -1, -1, -1]
    mask = padding.pad_to_length(torch.ones(3, dtype=torch.bool), 5, -1, mode=False)
    assert mask.dtype == torch.bool and mask.tolist() == [True, True, True, False, False]


def test_full_length_is_returned_as_is_and_errors_follow_the_reference():
    x = torch.rand(4, 2)
    assert padding.pad_to_length(x, 4) is x
    with pytest.raises(AssertionError):
        padding.pad_to_length(x, 3)
    with pytest.raises(ValueError):
        padding.pad_to_length(x, 6, mode="mirror")


def test_random_modes_stay_in_bounds():
    x = torch.rand(50, 2) * torch.tensor([100.0, 10.0])
    y = padding.pad_to_length(x, 80, -2, mode="random_c")
    for c in range(2):  # per-channel range of the data
        assert y[50:, c].min() >= x[:, c].min() and y[50:, c].max() <= x[:, c].max()
    d = torch.randn(50, 16)
    yd = padding.pad_to_length(d, 64, -2, mode="random", bounds=(-0.5, 0.25))
    assert yd[50:].min() >= -0.5 and yd[50:].max() <= 0.25
    # empty sequence: random_c falls back to the bounds (endomapper.py:454-460)
    e = padding.pad_to_length(torch.zeros(0, 2), 7, -2, mode="random_c", bounds=(0, 512))
    assert e.shape == (7, 2) and e.min() >= 0 and e.max() <= 512


def test_pad_and_stack_counts():
    seqs = [torch.rand(n, 2) for n in (5, 0, 9)]
    y, cnt = padding.pad_and_stack(seqs, None, -2, return_counts=True, mode="zeros")
    assert y.shape == (3, 9, 2) and cnt.dtype == torch.int32 and cnt.tolist() == [5, 0, 9]
    assert torch.equal(y[0, :5], seqs[0]) and (y[0, 5:] == 0).all() and (y[1] == 0).all()
    assert padding.pad_and_stack(seqs, 12, -2, mode="minus_one").shape == (3, 12, 2)


def test_pad_local_features_records_the_count_and_feeds_the_matcher_dict():
    f0 = [padding.pad_local_features(_feat(n, 10 + n), 64) for n in (40, 64, 1)]
    f1 = [padding.pad_local_features(_feat(n, 20 + n), 48, deterministic=True) for n in (48, 7, 30)]
    assert [int(f["num_keypoints"]) for f in f0] == [40, 64, 1]
    for f in f0:
        assert f["keypoints"].shape == (64, 2) and f["descriptors"].shape == (64, 256)
        assert f["sparse_depth"].shape == (64,) and f["valid_depth_mask"].dtype == torch.bool
    assert (f1[1]["descriptors"][7:] == 0).all() and (f1[1]["keypoints"][7:] == 0).all()
    assert (f0[0]["point3D_ids"][40:] == -1).all() and not f0[0]["valid_depth_mask"][40:].any()
    data = padding.matcher_inputs(f0, f1, image_sizes0=[[640, 480]] * 3, image_sizes1=[[640, 480]] * 3)
    assert data["keypoints0"].shape == (3, 64, 2) and data["descriptors1"].shape == (3, 48, 256)
    assert data["num_keypoints0"].tolist() == [40, 64, 1] and data["num_keypoints1"].tolist() == [48, 7, 30]
    assert data["scales0"].shape == (3, 64) and data["view1"]["image_size"].shape == (3, 2)
    with pytest.raises(ValueError):
        padding.matcher_inputs(f0, f1[:2])


@pytest.mark.skipif(not REF_MISC.exists(), reason="reference checkout not present")
def test_same_tensors_as_the_reference_producers():
    """misc.py imports only math / typing / torch, so the reference file is loaded directly (read-only)."""
    spec = importlib.util.spec_from_file_location("_ref_misc", REF_MISC)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    x = torch.rand(37, 2) * 300
    d = torch.randn(37, 128)
    s = torch.rand(37)
    cases = [
        (x, 64, -2, dict(mode="random_c")), (x, 64, -2, dict(mode="random_c", bounds=(0, 512))),
        (torch.zeros(0, 2), 16, -2, dict(mode="random_c", bounds=(0, 512))),
        (d, 50, -2, dict(mode="random")), (d, 50, -2, dict(mode="random", bounds=(-1.0, 1.0))),
        (s, 40, -1, dict(mode="zeros")), (s, 40, -1, dict(mode="minus_one")), (s, 40, -1, dict(mode="ones")),
        (s > 0.5, 40, -1, dict(mode=False)), (x, 37, -2, dict(mode="random_c")),
    ]
    for t, length, dim, kw in cases:
        torch.manual_seed(123)
        want = ref.pad_to_length(t, length, dim, **kw)
        torch.manual_seed(123)
        got = padding.pad_to_length(t, length, dim, **kw)
        assert got.dtype == want.dtype and torch.equal(got, want), kw
    seqs = [torch.rand(n, 2) for n in (3, 11, 7)]
    torch.manual_seed(5)
    want = ref.pad_and_stack(seqs, None, -2, mode="random_c")
    torch.manual_seed(5)
    got, cnt = padding.pad_and_stack(seqs, None, -2, return_counts=True, mode="random_c")
    assert torch.equal(got, want) and cnt.tolist() == [3, 11, 7]


@pytest.mark.gpu
def test_random_padding_cannot_influence_valid_matches():
    """End to end (C3): reference-style random padding + the counts == the un-padded pairs, fp32 kernels."""
    from helpers import build_model

    conf = {"filter_threshold": 0.1, "n_layers": 3}
    model = build_model(conf, 3).cuda()
    n0s, n1s = (200, 131, 256), (97, 256, 180)
    raw0 = [_feat(n, 100 + i, with_extras=False) for i, n in enumerate(n0s)]
    raw1 = [_feat(n, 200 + i, with_extras=False) for i, n in enumerate(n1s)]
    torch.manual_seed(9)
    f0 = [padding.pad_local_features(dict(f), 256, bounds=(0, 480)) for f in raw0]
    f1 = [padding.pad_local_features(dict(f), 256, bounds=(0, 480)) for f in raw1]
    sizes = [[640, 480]] * 3
    data = padding.matcher_inputs(f0, f1, sizes, sizes)
    data = {k: ({kk: vv.cuda() for kk, vv in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in data.items()}
    out = model(data)
    for b in range(3):
        one = {
            "keypoints0": raw0[b]["keypoints"][None].cuda(), "keypoints1": raw1[b]["keypoints"][None].cuda(),
            "descriptors0": raw0[b]["descriptors"][None].cuda(), "descriptors1": raw1[b]["descriptors"][None].cuda(),
            "view0": {"image_size": torch.tensor([sizes[b]]).float().cuda()},
            "view1": {"image_size": torch.tensor([sizes[b]]).float().cuda()},
        }
        ref = model(one)
        n0, n1 = n0s[b], n1s[b]
        # identical up to fp32 summation-order ties at the threshold (same bar as tests/test_driver.py)
        assert (out["matches0"][b, :n0] == ref["matches0"][0]).float().mean() >= 0.995
        assert (out["matches1"][b, :n1] == ref["matches1"][0]).float().mean() >= 0.995
        assert (out["matches0"][b, n0:] == -1).all() and (out["matching_scores0"][b, n0:] == 0).all()
        torch.testing.assert_close(out["log_assignment"][b, :n0, :n1], ref["log_assignment"][0, :n0, :n1],
                                   atol=2e-4, rtol=1e-4)


def test_counts_are_what_protects_valid_rows_from_the_padding():
    """Oracle (CPU restatement of the reference, pinned to its goldens): the reference has no mask input
    (lightglue.py:422-553), so reference-style random padding takes part in every attention and in both softmax
    normalisers and changes the valid block of log_assignment; with the recorded counts the padded batch gives exactly
    the un-padded result -- the semantics the B200 matcher implements with `num_keypoints0/1`."""
    from helpers import build_model, oracle_batch

    conf = {"filter_threshold": 0.1, "n_layers": 2}
    model = build_model(conf, 5)
    raw0, raw1 = _feat(90, 1, with_extras=False), _feat(70, 2, with_extras=False)
    torch.manual_seed(3)
    f0 = padding.pad_local_features(dict(raw0), 128, bounds=(0, 480))
    f1 = padding.pad_local_features(dict(raw1), 128, bounds=(0, 480))
    sizes = [[640, 480]]
    data = padding.matcher_inputs([f0], [f1], sizes, sizes)
    plain = padding.matcher_inputs([dict(raw0)], [dict(raw1)], sizes, sizes)
    want = oracle_batch(model, conf, plain)[0]["log_assignment"]                       # [91, 71]
    with_counts = oracle_batch(model, conf, data, num0=data["num_keypoints0"].tolist(),
                               num1=data["num_keypoints1"].tolist())[0]["log_assignment"]
    without = oracle_batch(model, conf, data)[0]["log_assignment"]                      # [129, 129], padding taken as real
    assert torch.equal(with_counts, want)
    assert (without[:90, :70] - want[:90, :70]).abs().max() > 1e-2
