"""Shared test helpers: build the product module / oracle weights for a golden fixture."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from glue_factory_colon_b200.lightglue import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402
from oracle import lightglue_oracle as oracle  # noqa: E402


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_model(conf, seed, overrides=None):
    """Seeded construction: bit-identical to the reference constructor under the same seed."""
    torch.manual_seed(seed)
    model = LightGlue(conf).eval()
    sd = model.state_dict()
    for k, v in (overrides or {}).items():
        sd[k].copy_(v)
    return model


def load_fixture(path):
    fx = torch.load(path, weights_only=False)
    model = build_model(fx["conf"], fx["seed"], fx["overrides"])
    assert abs(fingerprint(model.state_dict()) - fx["fingerprint"]) < 1e-6 * fx["fingerprint"], (
        "seeded weights differ from the ones the reference used for this fixture"
    )
    data = make_pairs(**fx["data_kwargs"])
    return fx, model, data


def oracle_batch(model, conf, data, dtype=torch.float32, num0=None, num1=None):
    return oracle.forward(model.state_dict(), dict(conf), data, dtype=dtype, num0=num0, num1=num1)


def assert_matches_equal(got_m, got_s, exp_m, exp_s, scores_row_gap=None, what=""):
    """Indices bit-exact; a mismatch is tolerated only where the oracle's own decision
    is numerically ambiguous (top-2 gap below 1e-4), which `scores_row_gap` reports."""
    bad = got_m != exp_m
    if scores_row_gap is not None:
        bad = bad & (scores_row_gap > 1e-4)
    assert not bad.any(), f"{what}: {int(bad.sum())} index mismatches"
    same = got_m == exp_m
    torch.testing.assert_close(got_s[same], exp_s[same], atol=2e-4, rtol=1e-3)


# ---- BASELINE config 1 fixture (oracle/make_golden_c1.py) -------------------------------------------------------


def sharp_assignment_overrides(layer=8, scale=2.0, bias=4.0):
    """Weights under which a random-init LightGlue makes confident matches: the exit layer's MatchAssignment gets
    final_proj = scale * I and a positive matchability bias, so the similarity of two tokens is (scale/4)^2 times the
    dot product of their layer outputs (which still carry the input descriptors through the residual stream).  With
    the synthetic pairs of synthetic.make_pairs (60 % true correspondences) several hundred matches then pass
    filter_threshold 0.1 -- without it none does and every `matches0 == oracle` check would be vacuous."""
    return {
        f"log_assignment.{layer}.final_proj.weight": scale * torch.eye(256),
        f"log_assignment.{layer}.final_proj.bias": torch.zeros(256),
        f"log_assignment.{layer}.matchability.bias": torch.tensor([bias]),
    }


def load_c1_fixture(path):
    """-> (fx, model, data): the boat1/boat2 pair exactly as oracle/make_golden_c1.py fed it to the reference."""
    fx = torch.load(path, weights_only=False)
    model = build_model(fx["conf"], fx["seed"], sharp_assignment_overrides())
    assert abs(fingerprint(model.state_dict()) - fx["fingerprint"]) < 1e-6 * fx["fingerprint"]

    def lift(sift_u8):  # same arithmetic as make_golden_c1.lift_descriptors
        g = torch.Generator().manual_seed(22)
        P = torch.randn(128, 256, generator=g) / 128 ** 0.5
        d = torch.nn.functional.normalize(sift_u8.float(), dim=-1)
        return torch.nn.functional.normalize(d @ P, dim=-1)

    data = {
        "keypoints0": fx["keypoints0"][None], "keypoints1": fx["keypoints1"][None],
        "descriptors0": lift(fx["sift0"])[None], "descriptors1": lift(fx["sift1"])[None],
        "view0": {"image_size": torch.tensor([fx["image_size0"]])},
        "view1": {"image_size": torch.tensor([fx["image_size1"]])},
    }
    return fx, model, data


def oracle_training_step(sd, conf, data, dtype=torch.float32):
    """One training step through the CPU oracle with torch autograd (test infrastructure): forward in training mode,
    LightGlue.loss, `total.mean().backward()`.  Returns (total [B], {state-dict name: gradient}, d descriptors0,
    d descriptors1)."""
    from oracle import loss_oracle

    leaves = {k: v.detach().cpu().to(dtype).clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and k != "confidence_thresholds"}
    d = dict(data)
    d["descriptors0"] = data["descriptors0"].detach().cpu().to(dtype).clone().requires_grad_(True)
    d["descriptors1"] = data["descriptors1"].detach().cpu().to(dtype).clone().requires_grad_(True)
    pred = loss_oracle.forward_collect(leaves, conf, d, keep_graph=True, dtype=dtype)
    losses, _ = loss_oracle.loss(leaves, conf, pred, d, True, keep_graph=True)
    losses["total"].mean().backward()
    return losses["total"].detach(), {k: v.grad for k, v in leaves.items()}, d["descriptors0"].grad, d["descriptors1"].grad
