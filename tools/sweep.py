"""BASELINE configs[2..4]: throughput sweep over keypoint counts, variable-count padded batch, adaptive mode.
Prints one JSON line per case (device-resident, CUDA-event timed)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402


def flops_per_pair(n, m, L=9):
    t = n + m
    return float(L * (2_490_368 * t + 1024 * (n * n + m * m) + 1536 * n * m) + 131_584 * t + 512 * n * m)


def timeit(model, data, iters):
    for _ in range(3):
        out = model(data)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        out = model(data)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


torch.manual_seed(0)
model = LightGlue({"precision": "bf16", "filter_threshold": 0.1}).eval().cuda()
# config 5: sweep (total tokens per batch held at 2 * 131072 so every point fills the GPU)
for n in (512, 1024, 2048, 4096, 8192):
    B = max(1, 131072 // n)
    data = make_pairs(B, n, n, seed=200, device="cuda")
    ms, _ = timeit(model, data, 5 if n <= 4096 else 3)
    pps = B / (ms * 1e-3)
    print(json.dumps({"case": "sweep", "kpts": n, "pairs": B, "ms_per_step": round(ms, 3), "pairs_per_s": round(pps, 1),
                      "tflops": round(pps * flops_per_pair(n, n) / 1e12, 1)}), flush=True)
    del data
    torch.cuda.empty_cache()
# config 3: 32 pairs x up to 4096 kpts, per-pair counts, padded to 4096
g = torch.Generator().manual_seed(3)
B = 32
n0 = torch.randint(1024, 4097, (B,), generator=g)
n1 = torch.randint(1024, 4097, (B,), generator=g)
data = make_pairs(B, 4096, 4096, seed=300, image_size=(512.0, 512.0), device="cuda")
data["num_keypoints0"], data["num_keypoints1"] = n0, n1
ms, out = timeit(model, data, 5)
fl = sum(flops_per_pair(int(a), int(b)) for a, b in zip(n0, n1))
assert (out["matches0"][0, int(n0[0]):] == -1).all()
print(json.dumps({"case": "variable_counts", "pairs": B, "pad": 4096, "mean_kpts": float((n0.float().mean() + n1.float().mean()) / 2),
                  "ms_per_step": round(ms, 3), "pairs_per_s": round(B / (ms * 1e-3), 1), "tflops_valid_tokens": round(fl / (ms * 1e-3) / 1e12, 1)}), flush=True)
del data
torch.cuda.empty_cache()
# config 4: adaptive depth/width at 2048 kpts (random-init heads never fire; biases force a realistic mix)
for B in (1, 16):
    torch.manual_seed(0)
    amodel = LightGlue({"precision": "bf16", "filter_threshold": 0.1, "depth_confidence": 0.95, "width_confidence": 0.99}).eval()
    sd = amodel.state_dict()
    for i in range(8):
        sd[f"token_confidence.{i}.token.0.bias"].fill_(3.0 if i >= 4 else -3.0)
        sd[f"log_assignment.{i}.matchability.bias"].fill_(-4.5 if i % 2 == 0 else 0.0)
    amodel = amodel.cuda()
    data = make_pairs(B, 2048, 2048, seed=400, device="cuda")
    ms, out = timeit(amodel, data, 5)
    print(json.dumps({"case": "adaptive", "pairs": B, "kpts": 2048, "ms_per_step": round(ms, 3), "pairs_per_s": round(B / (ms * 1e-3), 1),
                      "log_assignment_shape": list(out["log_assignment"].shape),
                      "mean_prune0": float(out["prune0"].float().mean())}), flush=True)
