"""Config 4 (adaptive depth / width, 2048 keypoints): forward time at batch 1 (the reference's limit) and 16.
usage: [LGB200_LIB=variant.so] python tools/adaptive_bench.py"""
import json, os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))  # (bench.py lives at the repo root)
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


# the unmodified reference module on the same GPU (batch 1 is all its adaptive mode accepts; its early-exit branch
# raises -- SURVEY.md F4 -- so only the width-pruning case can be timed; fp32 and flash=True under autocast(bf16))
import bench  # noqa: E402

for amp, extra in ((False, {}), (True, {"flash": True})):
    ref = bench.load_reference_module({"filter_threshold": 0.1, "width_confidence": 0.99, "depth_confidence": -1, **extra})
    if ref is None:
        break
    sd = ref.state_dict()
    for i in range(8):
        sd[f"token_confidence.{i}.token.0.bias"].fill_(3.0 if i >= 4 else -3.0)
        sd[f"log_assignment.{i}.matchability.bias"].fill_(-4.0)
    ref = ref.cuda()
    data = make_pairs(1, 2048, 2048, seed=400, device="cuda")
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            ms, out = timeit(ref, data, 10)
        print(json.dumps({"lib": "unmodified reference module on this GPU" + (", autocast(bf16) + flash" if amp else ", fp32"),
                          "case": "width only, every layer prunes", "pairs": 1, "ms_per_forward": round(ms, 3),
                          "pairs_per_s": round(1e3 / ms, 1), "log_assignment": list(out["log_assignment"].shape),
                          "mean_prune0": round(float(out["prune0"].float().mean()), 2)}), flush=True)
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"lib": "reference", "amp": amp, "error": repr(exc)[:200]}), flush=True)
