"""The differentiable use of the oracle (oracle/loss_oracle.py with keep_graph=True: torch autograd through the CPU
restatement) against gradient goldens of the UNMODIFIED reference (oracle/make_golden_grad.py: one training step,
`losses["total"].mean().backward()`).  This pins the checker that tests/test_gpu_grad.py uses for the hand-written
backward pass."""
import pytest
import torch

from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs
from helpers import oracle_training_step

CASES = ["grad_train", "grad_train_sift"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_autograd_against_reference_gradients(name, golden_dir):
    fx = torch.load(golden_dir / f"{name}.pt", weights_only=False)
    torch.manual_seed(fx["seed"])
    model = LightGlue(fx["conf"])
    sd = model.state_dict()
    fp = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(fp - fx["fingerprint"]) < 1e-6 * fx["fingerprint"]
    data = make_pairs(with_gt=True, **fx["data_kwargs"])
    total, grads, gd0, gd1 = oracle_training_step(sd, fx["conf"], data)
    torch.testing.assert_close(total.float(), fx["total"], atol=2e-4, rtol=1e-4)
    assert fx["no_grad"] == [] and set(fx["grads"]) == {k for k, g in grads.items() if g is not None}
    for k, s in list(fx["grads"].items()) + [("descriptors0", fx["descriptors0"]), ("descriptors1", fx["descriptors1"])]:
        g = {"descriptors0": gd0, "descriptors1": gd1}.get(k, grads.get(k)).double().reshape(-1)
        assert abs(float(g.norm()) - s["norm"]) <= 2e-3 * s["norm"] + 1e-7, k
        torch.testing.assert_close(g[s["idx"]].float(), s["val"], atol=2e-3 * s["norm"] / g.numel() ** 0.5 + 1e-7, rtol=5e-3,
                                   msg=lambda m: f"{k}: {m}")
