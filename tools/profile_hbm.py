"""The HBM-bound kernels of the path, one launch each, for an `ncu --set full` capture (north_star: achieved HBM GB/s
for the softmax normalisers, filtering and compaction; tensor-pipe share of the similarity GEMM).

  assign   : bf16 forward, 1 layer, 64 pairs x 2048 keypoints -> pack_rows, posenc, QKV / FFN2 / final_proj linears,
             rowdot, tc_assign_kernel<0> (pass 1: normalisers), tc_assign_kernel<1> (pass 2: scores + arg-maxima),
             assign_border, fm_mutual
  filter   : lgb200_filter_matches on a given [64,2049,2049] fp32 score matrix (the reference-matrix entry:
             fm_argmax reads the matrix once, fm_mutual finishes)
  adaptive : 16 pairs x 2048 keypoints, 3 layers, heads biased so that points are pruned -> rowdot (sigmoid),
             exit_check, prune_compact
usage: python tools/profile_hbm.py [assign|filter|adaptive|all] [pairs]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue, _abi  # noqa: E402
from glue_factory_colon_b200._abi import check, ptr  # noqa: E402
from glue_factory_colon_b200.synthetic import make_pairs  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = 2048
dev = "cuda:0"


def timed(fn, label):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1):.3f} ms")
    return out


if what in ("assign", "all"):
    torch.manual_seed(0)
    model = LightGlue({"precision": "bf16", "filter_threshold": 0.1, "n_layers": 1}).eval().to(dev)
    data = make_pairs(B, N, N, seed=100, device=dev)
    out = timed(lambda: model(data), f"assign: 1-layer bf16 forward, {B} pairs")
    print("  matches", int((out["matches0"] > -1).sum()))

if what in ("filter", "all"):
    lib = _abi.load()
    g = torch.Generator(device=dev).manual_seed(1)
    scores = torch.randn(B, N + 1, N + 1, device=dev, generator=g)
    m0 = torch.empty(B, N, device=dev, dtype=torch.int64)
    m1 = torch.empty(B, N, device=dev, dtype=torch.int64)
    ms0 = torch.empty(B, N, device=dev)
    ms1 = torch.empty(B, N, device=dev)
    ws = torch.empty(B * (2 * N + 2), device=dev, dtype=torch.int64)
    st = torch.cuda.current_stream().cuda_stream

    def run_filter():
        check(lib.lgb200_filter_matches(ptr(scores), B, N + 1, N + 1, None, 0.1, None, None, 0, N, N, ptr(m0), ptr(m1),
                                        ptr(ms0), ptr(ms1), ptr(ws), 0, st), "filter_matches")

    timed(run_filter, f"filter: given score matrix, {B} x {N + 1}^2 fp32 = {scores.numel() * 4 / 1e9:.3f} GB")
    # bit-exact against torch on the same matrix (lightglue.py:294-319)
    mx0 = scores[:, :-1, :-1].max(2).indices
    mx1 = scores[:, :-1, :-1].max(1).indices
    mutual0 = torch.arange(N, device=dev)[None] == mx1.gather(1, mx0)
    ok = ((m0 >= 0) <= mutual0).all() and (m0[m0 >= 0] == mx0[m0 >= 0]).all()
    print("  filter indices consistent with torch:", bool(ok))
    assert ok

if what in ("adaptive", "all"):
    Ba = min(B, 16)
    for label, extra in (("width only (prune_compact)", {"width_confidence": 0.99}),
                         ("depth only (exit_check)", {"depth_confidence": 0.95})):
        torch.manual_seed(0)
        conf = {"precision": "bf16", "filter_threshold": 0.1, "n_layers": 3, **extra}
        model = LightGlue(conf).eval()
        sd = model.state_dict()
        for i in range(2):
            sd[f"token_confidence.{i}.token.0.bias"].fill_(-3.0)
            sd[f"log_assignment.{i}.matchability.bias"].fill_(-4.5)  # sigmoid ~ 0.011: about half the points go
        model = model.to(dev)
        data = make_pairs(Ba, N, N, seed=400, device=dev)
        out = timed(lambda: model(data), f"adaptive, {label}: 3 layers, {Ba} pairs")
        print("  log_assignment", tuple(out["log_assignment"].shape), "prune0 mean", float(out["prune0"].float().mean()))
