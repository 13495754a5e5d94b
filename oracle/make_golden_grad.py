"""Generates tests/golden/grad_train.pt: one training step of the UNMODIFIED reference
(/root/reference/gluefactory/models/matchers/lightglue.py: forward in training mode, LightGlue.loss,
`losses["total"].mean().backward()` as gluefactory/train.py does) on seeded inputs with synthetic ground truth.
Build container only:

    python oracle/make_golden_grad.py

Stored per parameter (252 entries; weights are not stored -- seed + fingerprint as in make_golden.py): the L2 norm of
its gradient, the sum, and 16 entries at seeded positions.  The descriptors are leaves too (an extractor trained
jointly would receive this gradient): same summary for descriptors0/1.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle" / "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

from gluefactory.models import get_model  # noqa: E402

from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

OUT = ROOT / "tests" / "golden"


def summary(name: str, g: torch.Tensor) -> dict:
    g = g.detach().double().reshape(-1)
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    idx = torch.randint(0, g.numel(), (16,), generator=gen)
    return {"norm": float(g.norm()), "sum": float(g.sum()), "idx": idx, "val": g[idx].float().clone()}


def run_case(name, conf, seed, data_kwargs):
    torch.manual_seed(seed)
    model = get_model("matchers.lightglue")(conf)
    model.train(True)
    data = make_pairs(with_gt=True, **data_kwargs)
    data["descriptors0"].requires_grad_(True)
    data["descriptors1"].requires_grad_(True)
    pred = model(data)
    losses, _ = model.loss(pred, data)
    losses["total"].mean().backward()
    fx = {
        "name": name, "conf": conf, "seed": seed, "data_kwargs": data_kwargs,
        "fingerprint": float(sum(v.double().abs().sum() for v in model.state_dict().values())),
        "total": losses["total"].detach().clone(),
        "grads": {k: summary(k, p.grad) for k, p in model.named_parameters() if p.grad is not None},
        "no_grad": [k for k, p in model.named_parameters() if p.grad is None],
        "descriptors0": summary("descriptors0", data["descriptors0"].grad),
        "descriptors1": summary("descriptors1", data["descriptors1"].grad),
    }
    torch.save(fx, OUT / f"{name}.pt")
    big = sorted(fx["grads"].items(), key=lambda kv: -kv[1]["norm"])[:4]
    print(name, "total", [round(float(x), 5) for x in fx["total"]], "params with grad", len(fx["grads"]),
          "without", fx["no_grad"], "largest", [(k, round(v["norm"], 5)) for k, v in big])


def main():
    run_case("grad_train", {"filter_threshold": 0.0, "loss": {"gamma": 1.0, "fn": "nll", "nll_balancing": 0.5}}, 41,
             dict(B=2, n0=160, n1=144, seed=51))
    run_case("grad_train_sift", {"filter_threshold": 0.0, "input_dim": 128, "add_scale_ori": True, "n_layers": 3,
                                 "loss": {"gamma": 0.0, "fn": "nll", "nll_balancing": 0.3}}, 42,
             dict(B=1, n0=128, n1=128, seed=52, dim=128, scale_ori=True, with_size=False))


if __name__ == "__main__":
    main()
