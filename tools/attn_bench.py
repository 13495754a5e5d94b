"""Times the attention kernel alone at the bench shape (S=128, Lp=2048)."""
import os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, ptr
lib = _abi.load(Path(os.environ['LGB200_LIB']).resolve()) if os.environ.get('LGB200_LIB') else _abi.load()
S, Lp = 128, 2048
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(S * 4 * Lp, 64, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
k = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
ctx = torch.empty(S * Lp, 256, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attention {ms:.3f} ms  {4.0*S*4*Lp*Lp*64/ms/1e9:.0f} TFLOP/s")
