"""filter_matches on a given fp32 score matrix (the reference-matrix entry point) at the bench shape: time and
achieved HBM GB/s of the streaming pass (reads B*(N+1)*(M+1)*4 bytes once), result checked against torch.
usage: [LGB200_LIB=variant.so] python tools/filter_bench.py [pairs] [kpts]"""
import os, sys, json, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import ptr
lib = _abi.load(Path(os.environ["LGB200_LIB"]).resolve()) if os.environ.get("LGB200_LIB") else _abi.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
scores = torch.randn(B, N + 1, N + 1, device=dev, generator=g)
scores[:, 7, :] = scores[:, 7, 3:4]          # a row of exact ties
scores[0, :, 11] = scores[0, 5:6, 11]        # a column of exact ties
scores[1, 100, 200] = float("nan")
m0 = torch.empty(B, N, device=dev, dtype=torch.int64); m1 = torch.empty_like(m0)
s0 = torch.empty(B, N, device=dev); s1 = torch.empty_like(s0)
ws = torch.empty(B * (2 * N + 2), device=dev, dtype=torch.int64)
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.lgb200_filter_matches(ptr(scores), B, N + 1, N + 1, None, 0.1, None, None, 0, N, N, ptr(m0), ptr(m1), ptr(s0),
                                   ptr(s1), ptr(ws), 0, st)
    assert rc == 0, rc
for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
nbytes = scores.numel() * 4
# reference: filter_matches, lightglue.py:294-319, on the same matrix (torch on the device, two pairs on the CPU as well)
inner = scores[:, :-1, :-1]
mx0, mx1 = inner.max(2), inner.max(1)
i0, i1 = mx0.indices, mx1.indices
ar = torch.arange(N, device=dev)[None]
mut0, mut1 = ar == i1.gather(1, i0), ar == i0.gather(1, i1)
ms0 = torch.where(mut0, mx0.values.exp(), torch.zeros_like(s0))
ms1 = torch.where(mut1, ms0.gather(1, i1), torch.zeros_like(s1))
v0 = mut0 & (ms0 > 0.1)
v1 = mut1 & v0.gather(1, i1)
e_m0, e_m1 = torch.where(v0, i0, -1), torch.where(v1, i1, -1)
ok = bool(torch.equal(m0, e_m0) and torch.equal(m1, e_m1) and torch.equal(s0.nan_to_num(-5), ms0.nan_to_num(-5))
          and torch.equal(s1.nan_to_num(-5), ms1.nan_to_num(-5)))
cpu = scores[:2, :-1, :-1].cpu()
ok_cpu = bool(torch.equal(cpu.max(2).indices, i0[:2].cpu()) and torch.equal(cpu.max(1).indices, i1[:2].cpu()))
print(json.dumps({"lib": os.environ.get("LGB200_LIB", "default"), "pairs": B, "kpts": N, "us": round(ms * 1e3, 1),
                  "GBps_algorithmic": round(nbytes / ms / 1e6), "bit_exact_vs_torch": ok, "torch_cuda_eq_cpu_argmax": ok_cpu}))
assert ok
