"""The loss-side oracle (oracle/loss_oracle.py) against goldens produced by the unmodified reference
(oracle/make_golden_loss.py): forward in training mode (all layers collected), LightGlue.loss, matcher_metrics."""
import pytest
import torch

from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs
from oracle import loss_oracle

CASES = ["loss_train", "loss_train_gamma", "loss_eval"]


def _state(fx):
    torch.manual_seed(fx["seed"])
    model = LightGlue(fx["conf"])
    sd = model.state_dict()
    fp = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(fp - fx["fingerprint"]) < 1e-6 * fx["fingerprint"], "constructor no longer reproduces the reference init"
    return sd


@pytest.mark.parametrize("name", CASES)
def test_loss_oracle_against_reference_golden(name, golden_dir):
    fx = torch.load(golden_dir / f"{name}.pt", weights_only=False)
    sd = _state(fx)
    data = make_pairs(with_gt=True, **fx["data_kwargs"])
    pred = loss_oracle.forward_collect(sd, fx["conf"], data)
    if not fx["training"]:  # eval: only the last layer is returned (lightglue.py:495-498)
        pred["ref_descriptors0"] = pred["ref_descriptors0"][:, -1:]
        pred["ref_descriptors1"] = pred["ref_descriptors1"][:, -1:]
    assert tuple(pred["ref_descriptors0"].shape) == fx["ref_desc_shape"]
    for i, am in enumerate(fx["ref_desc_absmean"]):
        assert abs(float(pred["ref_descriptors0"][:, i].abs().mean()) - am) < 1e-4 * am
    torch.testing.assert_close(pred["log_assignment"], fx["pred"]["log_assignment"], atol=2e-4, rtol=0)
    # the loss itself, on the reference's own predictions for log_assignment / matches
    pred_l = {**pred, **fx["pred"]}
    losses, metrics = loss_oracle.loss(sd, fx["conf"], pred_l, data, fx["training"])
    assert set(losses) == set(fx["losses"])
    for k, v in fx["losses"].items():
        torch.testing.assert_close(losses[k].reshape(-1), v.reshape(-1).float(), atol=2e-4, rtol=1e-4, msg=lambda m: f"{k}: {m}")
    assert set(metrics) == set(fx["metrics"])
    for k, v in fx["metrics"].items():
        torch.testing.assert_close(metrics[k], v, atol=1e-6, rtol=1e-6)


def test_matcher_metrics_known_answers(golden_dir):
    fx = torch.load(golden_dir / "metrics_kat.pt", weights_only=False)
    data = make_pairs(with_gt=True, **fx["data_kwargs"])
    met = loss_oracle.matcher_metrics({"matches0": fx["matches0"], "matching_scores0": fx["matching_scores0"]}, data)
    for k, v in fx["metrics"].items():
        assert float(v.abs().sum()) > 0
        torch.testing.assert_close(met[k], v, atol=1e-6, rtol=1e-6)
