"""One adaptive forward (batch 1, 2048 keypoints, exit at layer 4) eager, for an ncu launch list."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs
torch.manual_seed(0)
m = LightGlue({"precision": "bf16", "filter_threshold": 0.1, "depth_confidence": 0.95, "width_confidence": 0.99}).eval()
sd = m.state_dict()
for i in range(8):
    sd[f"token_confidence.{i}.token.0.bias"].fill_(3.0 if i >= 4 else -3.0)
    sd[f"log_assignment.{i}.matchability.bias"].fill_(-4.5 if i % 2 == 0 else 0.0)
m = m.cuda()
data = make_pairs(1, 2048, 2048, seed=400, device="cuda")
for _ in range(2):
    out = m(data)
torch.cuda.synchronize()
print(out["log_assignment"].shape, float(out["prune0"].float().mean()))
