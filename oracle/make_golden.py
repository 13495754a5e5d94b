"""Generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference/gluefactory/models/matchers/lightglue.py) on seeded inputs.

Run in the build container only (the reference is not available on the GPU box):

    python oracle/make_golden.py

Weights are not stored: every fixture records the torch seed under which the
reference constructor was called; `glue_factory_colon_b200.LightGlue` built under
the same seed has a bit-identical state_dict (tests/test_oracle.py checks the
recorded fingerprint), plus explicit `overrides` applied after construction.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle" / "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))

from gluefactory.models import get_model  # noqa: E402
from gluefactory.models.matchers import lightglue as ref_mod  # noqa: E402

from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

OUT = ROOT / "tests" / "golden"


def fingerprint(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_ref(conf, seed, overrides):
    torch.manual_seed(seed)
    model = get_model("matchers.lightglue")(conf).eval()
    sd = model.state_dict()
    for k, v in overrides.items():
        sd[k].copy_(v)
    return model


def run_case(name, conf, seed, data_kwargs, overrides=None):
    overrides = overrides or {}
    model = build_ref(conf, seed, overrides)
    data = make_pairs(**data_kwargs)
    with torch.no_grad():
        out = model(data)
    keep = ["matches0", "matches1", "matching_scores0", "matching_scores1", "log_assignment", "prune0", "prune1"]
    fx = {
        "name": name,
        "conf": conf,
        "seed": seed,
        "overrides": overrides,
        "data_kwargs": data_kwargs,
        "fingerprint": fingerprint(model.state_dict()),
        "out": {k: out[k].clone() for k in keep},
        "ref_desc_absmean": float(out["ref_descriptors0"].abs().mean()),
    }
    torch.save(fx, OUT / f"{name}.pt")
    print(name, {k: tuple(v.shape) for k, v in fx["out"].items()},
          "valid matches:", int((out["matches0"] > -1).sum()))


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    run_case("basic", {"filter_threshold": 0.0}, 1, dict(B=2, n0=200, n1=168, seed=11))
    run_case(
        "nosize_sift",
        {"input_dim": 128, "add_scale_ori": True, "filter_threshold": 0.0},
        2,
        dict(B=1, n0=150, n1=130, seed=12, dim=128, with_size=False, scale_ori=True),
    )
    # point pruning only (early exit crashes in the reference).  The
    # matchability biases are shifted so that a part of the points is pruned.
    ov = {}
    for i, b in enumerate([-4.6, 0.0, -4.4, 0.0, -4.2, 0.0, 0.0, -4.0]):
        ov[f"log_assignment.{i}.matchability.bias"] = torch.tensor([b])
    run_case("prune", {"width_confidence": 0.99, "filter_threshold": 0.0}, 3,
             dict(B=1, n0=190, n1=170, seed=13), ov)

    # filter_matches known-answer vectors (lightglue.py:294-319), including ties and an empty side
    g = torch.Generator().manual_seed(5)
    sc = torch.randn(3, 41, 37, generator=g)
    sc = (sc * 4).round() / 4  # quantise -> plenty of exact ties
    sc[1] -= 3.0
    cases = []
    for th in (0.0, 0.2):
        m0, m1, s0, s1 = ref_mod.filter_matches(sc, th)
        cases.append(dict(scores=sc.clone(), th=th, m0=m0, m1=m1, ms0=s0, ms1=s1))
    e = torch.zeros(2, 1, 9)
    m0, m1, s0, s1 = ref_mod.filter_matches(e, 0.0)
    cases.append(dict(scores=e, th=0.0, m0=m0, m1=m1, ms0=s0, ms1=s1))
    torch.save(cases, OUT / "filter_kat.pt")
    print("filter_kat", len(cases))


if __name__ == "__main__":
    main()
