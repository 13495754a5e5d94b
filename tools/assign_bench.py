"""Times the two MatchAssignment passes and filter_matches alone at the bench shape (B=64, N=M=2048)."""
import os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, ptr
lib = _abi.load(Path(os.environ["LGB200_LIB"]).resolve()) if os.environ.get("LGB200_LIB") else _abi.load()
B, Lp = 64, 2048
S = 2 * B
R = C = int(os.environ.get('ASSIGN_RC', Lp + 1))  # ASSIGN_RC=2048: rows of 8192 bytes (aligned stores), n = 2047
g = torch.Generator(device="cuda").manual_seed(0)
md = (torch.randn(S * Lp, 256, device="cuda", generator=g) * 0.25).to(torch.bfloat16)
z = torch.randn(S * Lp, device="cuda", generator=g)
lse = torch.empty(S * Lp, device="cuda")
scores = torch.empty(B, R, C, device="cuda")
ws = torch.empty(B * (R + C), device="cuda", dtype=torch.int64)
print(f"R = C = {R}")
m0 = torch.empty(B, Lp, device="cuda", dtype=torch.int64); m1 = torch.empty_like(m0)
s0 = torch.empty(B, Lp, device="cuda"); s1 = torch.empty_like(s0)
st = torch.cuda.current_stream().cuda_stream


def timeit(name, fn, nbytes):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms * 1e3:.0f} us  {nbytes / ms / 1e6:.0f} GB/s (algorithmic bytes)")


def p1():
    assert lib.lgb200_assign_lse(BF16, ptr(md), S, Lp, None, ptr(lse), st) == 0
def p2():
    assert lib.lgb200_assign_scores(BF16, ptr(md), ptr(z), ptr(lse), B, Lp, None, R, C, ptr(scores), ptr(ws), st) == 0
def fm():
    assert lib.lgb200_filter_matches(None, B, R, C, None, 0.1, None, None, 0, Lp, Lp, ptr(m0), ptr(m1), ptr(s0), ptr(s1), ptr(ws), 1, st) == 0
timeit("assign pass 1 (lse)", p1, S * Lp * 256 * 2 + S * Lp * 4)
timeit("assign pass 2 (scores + border + memset)", p2, B * R * C * 4 + S * Lp * 256 * 2)
def p2n():
    assert lib.lgb200_assign_scores(BF16, ptr(md), ptr(z), ptr(lse), B, Lp, None, R, C, ptr(scores), None, st) == 0
timeit("assign pass 2 without the fused arg-maxima", p2n, B * R * C * 4 + S * Lp * 256 * 2)
timeit("filter_matches (packed maxima in workspace)", fm, B * (R + C) * 8 + B * 2 * Lp * 12)
