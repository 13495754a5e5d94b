"""Generates tests/golden/loss_train.pt and loss_eval.pt by running the UNMODIFIED reference
(/root/reference/gluefactory/models/matchers/lightglue.py: forward + LightGlue.loss) on seeded inputs with
synthetic ground truth.  Build container only:

    python oracle/make_golden_loss.py

Weights are not stored (seed + fingerprint, as in make_golden.py).  The token-confidence and matchability heads get
random weights under the seed like everything else, so the confidence BCE term is non-trivial.
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


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def run_case(name, conf, seed, data_kwargs, training):
    torch.manual_seed(seed)
    model = get_model("matchers.lightglue")(conf)
    model.train(training)
    data = make_pairs(with_gt=True, **data_kwargs)
    with torch.no_grad():
        pred = model(data)
        losses, metrics = model.loss(pred, data)
    fx = {
        "name": name,
        "conf": conf,
        "seed": seed,
        "training": training,
        "data_kwargs": data_kwargs,
        "fingerprint": fingerprint(model.state_dict()),
        "pred": {k: pred[k].clone() for k in ("matches0", "matching_scores0", "log_assignment")},
        "ref_desc_shape": tuple(pred["ref_descriptors0"].shape),
        "ref_desc_absmean": [float(pred["ref_descriptors0"][:, i].abs().mean()) for i in range(pred["ref_descriptors0"].shape[1])],
        "losses": {k: (v.clone() if isinstance(v, torch.Tensor) else torch.tensor(v)) for k, v in losses.items()},
        "metrics": {k: v.clone() for k, v in metrics.items()},
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, fx["ref_desc_shape"], {k: [round(float(x), 5) for x in v.reshape(-1)[:2]] for k, v in fx["losses"].items()},
          {k: [round(float(x), 4) for x in v.reshape(-1)[:2]] for k, v in fx["metrics"].items()})


def main():
    run_case("loss_train", {"filter_threshold": 0.0, "loss": {"gamma": 1.0, "fn": "nll", "nll_balancing": 0.5}}, 21,
             dict(B=2, n0=160, n1=144, seed=31), True)
    run_case("loss_train_gamma", {"filter_threshold": 0.0, "loss": {"gamma": 0.0, "fn": "nll", "nll_balancing": 0.3}}, 22,
             dict(B=1, n0=96, n1=120, seed=32), True)
    run_case("loss_eval", {"filter_threshold": 0.1}, 23, dict(B=2, n0=150, n1=170, seed=33), False)

    # matcher_metrics known-answer vectors (random init never matches correctly, so the metrics above are all zero):
    # ground truth corrupted at random
    from gluefactory.models.utils.metrics import matcher_metrics

    data = make_pairs(B=3, n0=120, n1=100, seed=34, with_gt=True)
    g = torch.Generator().manual_seed(7)
    m = data["gt_matches0"].clone()
    flip = torch.rand(m.shape, generator=g)
    m[flip < 0.25] = -1
    wrong = (flip > 0.8)
    m[wrong] = torch.randint(0, 100, m.shape, generator=g)[wrong]
    sc = torch.rand(m.shape, generator=g) * (m > -1)
    met = matcher_metrics({"matches0": m, "matching_scores0": sc}, data)
    torch.save({"data_kwargs": dict(B=3, n0=120, n1=100, seed=34), "matches0": m, "matching_scores0": sc,
                "metrics": {k: v.clone() for k, v in met.items()}}, OUT / "metrics_kat.pt")
    print("metrics_kat", {k: [round(float(x), 4) for x in v] for k, v in met.items()})


if __name__ == "__main__":
    main()
