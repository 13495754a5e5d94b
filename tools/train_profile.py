"""One training step (forward in training mode + loss + backward) of the drop-in between cudaProfilerStart/Stop, after two
warm-up steps: for `ncu --profile-from-start off --metrics gpu__time_duration.sum`.  Usage: train_profile.py [B] [kpts]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
data = make_pairs(B, N, N, seed=3, device=dev, with_gt=True)
torch.manual_seed(0)
model = LightGlue({"filter_threshold": 0.1}).to(dev).train()


def step():
    model.zero_grad(set_to_none=True)
    pred = model(data)
    losses, _ = model.loss(pred, data)
    losses["total"].mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
