"""Attention backward kernels alone (csrc/lg_bwd.cu through lgb200_attention_bwd): S sequences x 4 heads x Lp keypoints,
fp32.  Prints ms per call and the useful fp32-equivalent TFLOP/s (8 tile products of 2.Lp^2.64 per head: 1 statistics +
3 dQ + 4 dK/dV).  Usage: python tools/attn_bwd_bench.py [S] [Lp] [reps]   (LGB200_ATTN_BWD_SIMT=1: CUDA-core kernels)"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi  # noqa: E402
from glue_factory_colon_b200._abi import ptr  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Lp = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
lib = _abi.load()
g = torch.Generator().manual_seed(0)
q, k = ((torch.randn(S, 4, Lp, 64, generator=g) * 0.6).to(dev) for _ in range(2))
v = torch.randn(S, 4, Lp, 64, generator=g).to(dev)
dctx = torch.randn(S, Lp, 256, generator=g).to(dev)
ctx = torch.zeros(S, Lp, 256, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
assert lib.lgb200_attention(_abi.F32, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st) == 0
dq, dk, dv = (torch.empty(S, 4, Lp, 64, device=dev) for _ in range(3))
n_ws = ctypes.c_longlong(0)
assert lib.lgb200_attention_bwd_workspace(S, Lp, ctypes.byref(n_ws)) == 0
ws = torch.empty(n_ws.value, device=dev)


def call():
    rc = lib.lgb200_attention_bwd(ptr(q), ptr(k), ptr(v), ptr(ctx), ptr(dctx), S, Lp, None, 0, ptr(dq), ptr(dk), ptr(dv),
                                  ptr(ws), st)
    assert rc == 0


for _ in range(2):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    call()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / REPS
flop = 8 * 2.0 * Lp * Lp * 64 * 4 * S
print(f"attention backward S={S} Lp={Lp}: {ms:.3f} ms per call, {flop / ms / 1e9:.1f} TFLOP/s useful (fp32-equivalent), "
      f"checksum {float(dq.abs().mean()):.6f} {float(dk.abs().mean()):.6f} {float(dv.abs().mean()):.6f}")
