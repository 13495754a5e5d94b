"""Training step (forward in training mode + LightGlue.loss + backward) through the drop-in's hand-written backward
pass, and through the UNMODIFIED reference module (baseline/_ref, PyTorch autograd over cuBLAS / SDPA kernels) on the
same GPU and the same batch.  Usage: python tools/train_bench.py [B] [kpts] [reps]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glue_factory_colon_b200 import LightGlue  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
conf = {"filter_threshold": 0.1}
data = make_pairs(B, N, N, seed=3, device=dev, with_gt=True)


def step(model):
    model.zero_grad(set_to_none=True)
    pred = model(data)
    losses, _ = model.loss(pred, data)
    losses["total"].mean().backward()
    return float(losses["total"].mean().detach())


def timed(model, label, extra=None):
    for _ in range(2):
        val = step(model)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        step(model)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    rec = {"impl": label, "pairs": B, "kpts": N, "ms_per_step": round(ms, 2), "pairs_per_s": round(B / ms * 1e3, 1),
           "loss": round(val, 5), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    rec.update(extra or {})
    print(json.dumps(rec), flush=True)
    return rec


for ckpt in (False, True):
    torch.manual_seed(0)
    ours = LightGlue({**conf, "checkpointed": ckpt}).to(dev).train()
    timed(ours, f"glue_factory_colon_b200 (fp32-accurate tcgen05 kernels forward and backward, three-product tensor-core GEMMs), checkpointed={ckpt}")
    del ours
    torch.cuda.empty_cache()
for ckpt in (False, True):
    ref = bench.load_reference_module({**conf, "checkpointed": ckpt})
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing"}))
        break
    ref = ref.to(dev).train()
    try:
        timed(ref, f"unmodified reference module, fp32 autograd, checkpointed={ckpt}")
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"impl": f"reference checkpointed={ckpt}", "error": repr(exc)[:200]}))
    del ref
    torch.cuda.empty_cache()
