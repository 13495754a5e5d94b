"""precision="fp32" (split-fp16 tensor-core mode) forward at the bench shape: 64 pairs x 2048 keypoints, device-resident."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

torch.manual_seed(0)
model = LightGlue({"precision": "fp32", "filter_threshold": 0.1}).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
data = make_pairs(B, 2048, 2048, seed=200, device="cuda")
for _ in range(2):
    out = model(data)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    out = model(data)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"precision": "fp32", "pairs": B, "kpts": 2048, "ms_per_step": round(ms, 3), "pairs_per_s": round(B / ms * 1e3, 1),
                  "checksum": float(out["log_assignment"].float().mean())}))
