"""One warm-up forward + N timed forwards at the bench shape; used under ncu (launch list / --set full)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--kpts", type=int, default=2048)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
torch.manual_seed(0)
model = LightGlue({"precision": a.precision, "filter_threshold": 0.1}).eval().cuda()
data = make_pairs(a.pairs, a.kpts, a.kpts, seed=100, device="cuda")
model(data)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    out = model(data)
e1.record()
torch.cuda.synchronize()
print(f"ms/forward {e0.elapsed_time(e1) / a.iters:.3f}  matches {(out['matches0'] > -1).sum().item()}")
