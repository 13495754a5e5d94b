"""Config 4 (adaptive depth / width, 2048 keypoints): forward time at batch 1 (the reference's limit) and 16.
usage: [LGB200_LIB=variant.so] python tools/adaptive_bench.py"""
import json, os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue, _abi
from glue_factory_colon_b200.synthetic import make_pairs
if os.environ.get("LGB200_LIB"):
    _abi.load(Path(os.environ["LGB200_LIB"]).resolve())


def timeit(model, data, iters):
    for _ in range(3):
        out = model(data)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        out = model(data)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


for label, conf, bias in (
    ("depth+width, exit at layer 4", {"depth_confidence": 0.95, "width_confidence": 0.99}, None),
    ("width only, every layer prunes", {"width_confidence": 0.99}, -4.0),
):
    for B, graph in ((1, False), (1, True), (16, False), (16, True)):
        torch.manual_seed(0)
        m = LightGlue({"precision": "bf16", "filter_threshold": 0.1, "cuda_graph": graph, **conf}).eval()
        sd = m.state_dict()
        for i in range(8):
            sd[f"token_confidence.{i}.token.0.bias"].fill_(3.0 if i >= 4 else -3.0)
            sd[f"log_assignment.{i}.matchability.bias"].fill_(bias if bias is not None else (-4.5 if i % 2 == 0 else 0.0))
        m = m.cuda()
        data = make_pairs(B, 2048, 2048, seed=400, device="cuda")
        ms, out = timeit(m, data, 20 if B == 1 else 5)
        print(json.dumps({"lib": os.environ.get("LGB200_LIB", "default"), "case": label, "cuda_graph": graph, "pairs": B, "ms_per_forward": round(ms, 3),
                          "pairs_per_s": round(B / ms * 1e3, 1), "log_assignment": list(out["log_assignment"].shape),
                          "mean_prune0": round(float(out["prune0"].float().mean()), 2)}), flush=True)
