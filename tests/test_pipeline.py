"""Adaptor between the reference's matcher call site / export loop and the batched pair driver
(glue_factory_colon_b200/pipeline.py; reference: models/two_view_pipeline.py:326-335, utils/export_predictions.py:21-90).
CPU tests cover the host logic with a stand-in driver; the GPU test runs the real one against the oracle."""
import numpy as np
import pytest
import torch

from helpers import build_model, make_pairs, oracle_batch, sharp_assignment_overrides
from glue_factory_colon_b200 import pipeline


def _items(counts, seed=81, dim=256, scale=None):
    """Call-site dicts as TwoViewPipeline builds them for its matcher: batch-1 tensors + view dicts + name."""
    out = []
    for i, (n0, n1) in enumerate(counts):
        d = make_pairs(B=1, n0=n0, n1=n1, seed=seed + i, dim=dim)
        d["name"] = [f"pair{i:03d}"]
        d["keypoint_scores0"] = torch.rand(1, n0)
        if scale is not None:
            d["view0"]["scales"] = torch.tensor([[scale, scale]])
            d["view1"]["scales"] = torch.tensor([[scale, scale]])
        out.append(d)
    return out


class _FakeDriver:
    """Stand-in for BatchedPairMatcher on a CPU box: deterministic "matches" derived from the pair's sizes."""

    def __init__(self, matcher, **kw):
        self.kw = kw

    def match(self, pairs):
        buf = list(pairs)  # pulls the whole stream first, like the driver's window
        for p in buf:
            n0, n1 = p["keypoints0"].shape[0], p["keypoints1"].shape[0]
            assert p["keypoints0"].shape == (n0, 2) and p["descriptors1"].shape == (n1, 256)
            assert p["image_size0"].shape == (2,)
            yield {"matches0": torch.arange(n0) % n1, "matches1": torch.arange(n1) % n0,
                   "matching_scores0": torch.full((n0,), 0.5), "matching_scores1": torch.full((n1,), 0.25)}


def test_matcher_inputs_to_pair_shapes():
    d = _items([(7, 5)])[0]
    d["scales0"], d["oris0"] = torch.rand(1, 7, 1), torch.rand(1, 7)
    p = pipeline.matcher_inputs_to_pair(d)
    assert p["keypoints0"].shape == (7, 2) and p["descriptors1"].shape == (5, 256)
    assert p["scales0"].shape == (7,) and p["oris0"].shape == (7,)
    assert torch.equal(p["image_size0"], torch.tensor([640.0, 480.0]))
    with pytest.raises(AssertionError, match="Missing key"):
        pipeline.matcher_inputs_to_pair({"keypoints0": d["keypoints0"]})


def test_export_matches_post_processing(monkeypatch):
    monkeypatch.setattr(pipeline, "BatchedPairMatcher", _FakeDriver)
    items = _items([(9, 6), (4, 8), (5, 5)], scale=2.0)
    w = pipeline.DictWriter()
    n = pipeline.export_matches(items, matcher=None, writer=w, as_half=True,
                                keys=["keypoints0", "matches0", "matching_scores0", "extra"],
                                optional_keys=["keypoint_scores0", "not_there"],
                                callback_fn=lambda pred, data: {"extra": pred["matches0"].float() + 1, "matches0": None})
    assert n == 3 and list(w) == ["pair000", "pair001", "pair002"]  # input order, one group per name
    g = w["pair001"]
    assert set(g) == {"keypoints0", "matches0", "matching_scores0", "extra", "keypoint_scores0"}
    assert g["matches0"].dtype == np.int64 and g["matches0"].shape == (4,)   # pred wins over the callback's key
    assert g["matching_scores0"].dtype == np.float16 and g["extra"].dtype == np.float16
    np.testing.assert_allclose(g["keypoints0"], (items[1]["keypoints0"][0] / 2.0).numpy().astype(np.float16))
    with pytest.raises(ValueError, match="Missing key"):
        pipeline.export_matches(_items([(3, 3)]), None, pipeline.DictWriter(), keys=["nope"])


def test_streamed_matcher_returns_call_site_dicts(monkeypatch):
    monkeypatch.setattr(pipeline, "BatchedPairMatcher", _FakeDriver)
    items = _items([(6, 4), (3, 9)])
    got = list(pipeline.StreamedMatcher(None).match(iter(items)))
    assert [it["name"][0] for it, _ in got] == ["pair000", "pair001"]
    for (it, mp), (n0, n1) in zip(got, [(6, 4), (3, 9)]):
        assert mp["matches0"].shape == (1, n0) and mp["matching_scores1"].shape == (1, n1)
        merged = {**it, **mp}  # two_view_pipeline.py:335
        assert merged["matches0"].dtype == torch.int64


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "fp32_simt", "bf16"])
def test_export_matches_against_oracle(prec):
    """A stream of 11 pairs with different keypoint counts through the export loop -> per pair, what the oracle returns
    for that pair alone (fp32: every index outside numerically tied rows; bf16: >= 98 % of the keypoints)."""
    counts = [(300, 260), (129, 260), (512, 512), (77, 400), (640, 130), (1, 50), (256, 256), (333, 222), (600, 600),
              (50, 1), (420, 421)]
    conf = {"filter_threshold": 0.1, "precision": prec}
    model = build_model(conf, 5, sharp_assignment_overrides())
    items = _items(counts, seed=90)
    exp = [oracle_batch(model, conf, it)[0] for it in items]
    w = pipeline.DictWriter()
    n = pipeline.export_matches(items, model.to("cuda:0"), w, keys=["matches0", "matches1", "matching_scores0"],
                                max_pairs=4, window=6)
    assert n == len(counts) and list(w) == [f"pair{i:03d}" for i in range(len(counts))]
    assert sum(int((r["matches0"] > -1).sum()) for r in exp) > 500  # not vacuous
    for i, r in enumerate(exp):
        g = w[f"pair{i:03d}"]
        assert g["matches0"].shape == (counts[i][0],) and g["matches1"].shape == (counts[i][1],)
        eq0 = (torch.from_numpy(g["matches0"]) == r["matches0"]).float().mean().item()
        eq1 = (torch.from_numpy(g["matches1"]) == r["matches1"]).float().mean().item()
        need = 0.995 if prec != "bf16" else 0.98
        if min(counts[i]) >= 50:  # (small images: at most two keypoints may differ, whatever the rate)
            ok0 = eq0 >= need or round((1 - eq0) * counts[i][0]) <= 2
            ok1 = eq1 >= need or round((1 - eq1) * counts[i][1]) <= 2
            assert ok0 and ok1, (i, counts[i], eq0, eq1)
        same = torch.from_numpy(g["matches0"]) == r["matches0"]
        tol = 1e-3 if prec != "bf16" else 0.12
        assert (torch.from_numpy(g["matching_scores0"])[same] - r["matching_scores0"][same]).abs().max() <= tol
