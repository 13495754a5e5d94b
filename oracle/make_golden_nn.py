"""Generates tests/golden/nn_matcher.pt by running the UNMODIFIED reference NearestNeighborMatcher
(/root/reference/gluefactory/models/matchers/nearest_neighbor_matcher.py) on seeded inputs.
Build container only (the reference is not available on the GPU box):  python oracle/make_golden_nn.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle" / "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

from gluefactory.models import get_model  # noqa: E402

from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

CASES = [
    dict(name="mutual", conf={}, data=dict(B=2, n0=140, n1=111, seed=21)),
    dict(name="ratio_dist", conf={"ratio_thresh": 0.9, "distance_thresh": 1.1, "mutual_check": False},
         data=dict(B=1, n0=90, n1=130, seed=22, dim=128)),
    dict(name="ratio_mutual", conf={"ratio_thresh": 0.8}, data=dict(B=2, n0=64, n1=64, seed=23)),
]


def main():
    out = []
    for c in CASES:
        model = get_model("matchers.nearest_neighbor_matcher")(c["conf"]).eval()
        data = make_pairs(**c["data"])
        with torch.no_grad():
            pred = model(data)
        out.append(dict(name=c["name"], conf=c["conf"], data_kwargs=c["data"],
                        out={k: pred[k].clone() for k in ("matches0", "matches1", "matching_scores0",
                                                           "matching_scores1", "similarity", "log_assignment")}))
        print(c["name"], "valid matches:", int((pred["matches0"] > -1).sum()), tuple(pred["similarity"].shape))
    torch.save(out, ROOT / "tests" / "golden" / "nn_matcher.pt")
    # N_pair loss (nearest_neighbor_matcher.py:85-109): forward values of the unmodified reference, eval and train mode
    loss_cases = []
    for name, temp, train, kw in (("eval", 1.0, False, dict(B=2, n0=140, n1=111, seed=24, with_gt=True)),
                                  ("train_T3", 3.0, True, dict(B=1, n0=90, n1=130, seed=25, dim=128, with_gt=True))):
        model = get_model("matchers.nearest_neighbor_matcher")({"loss": "N_pair"})
        model.train(train)
        with torch.no_grad():
            model.temperature.fill_(temp)
            data = make_pairs(**kw)
            pred = model(data)
            losses, metrics = model.loss(pred, data)
        loss_cases.append(dict(name=name, temperature=temp, train=train, data_kwargs=kw,
                               losses={k: v.clone() for k, v in losses.items()},
                               metrics={k: v.clone() for k, v in metrics.items()}))
        print("npair", name, {k: [round(float(x), 5) for x in v.reshape(-1)] for k, v in losses.items()})
    torch.save(loss_cases, ROOT / "tests" / "golden" / "nn_npair_loss.pt")


if __name__ == "__main__":
    main()
