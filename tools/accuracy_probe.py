"""Accuracy of the bf16 path vs the fp32 CPU oracle, with the fp32 residual stream and with the
residual rounded to bf16 after every block (emulation of a bf16-only residual)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs, to_device
from oracle import lightglue_oracle as oracle

for n in (512, 1024):
    conf = {"filter_threshold": 0.0}
    torch.manual_seed(0)
    model = LightGlue(conf).eval()
    data = make_pairs(B=2, n0=n, n1=n, seed=61)
    ref = oracle.forward(model.state_dict(), conf, data)
    model = model.cuda()
    for mode in ("fp32", "bf16"):
        model.conf.precision = "fp32" if mode == "fp32" else "bf16"
        out = model(to_device(data, "cuda"))
        for b, r in enumerate(ref):
            la, lo = out["log_assignment"][b].cpu(), r["log_assignment"]
            d = (la - lo).abs()
            agree = (la[:-1, :-1].argmax(1) == lo[:-1, :-1].argmax(1)).float().mean()
            mm = (out["matches0"][b].cpu() == r["matches0"]).float().mean()
            print(f"n={n} {mode:11s} pair{b}: mean|d|={d.mean():.4f} max|d|={d.max():.4f} row-argmax agree={agree:.4f} matches0 equal={mm:.4f}")
