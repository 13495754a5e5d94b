import sys, ctypes, torch, numpy as np, os
from pathlib import Path
os.environ["LGB200_ATTN_DBG"] = str(16 + int(sys.argv[1]) if len(sys.argv) > 1 else 16)
os.environ.setdefault("LGB200_ATTN_CL", "1")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, ptr
lib = _abi.load(Path(os.environ['LGB200_LIB']).resolve()) if os.environ.get('LGB200_LIB') else _abi.load()
S, Lp = 128, 2048
q = (torch.randn(S*4*Lp, 64, device="cuda")*0.5).to(torch.bfloat16); k = torch.randn_like(q); v = torch.randn_like(q)
ctx = torch.empty(S*Lp, 256, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.lgb200_debug_attn_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.lgb200_debug_attn_times(buf, 512))
t = np.array(buf[:]).reshape(2, 16, 16)
t0 = t[0, 2, 0]
print("softmax warp (cycles rel. to step 2 start): cols = start, s_full woke, ld done, s_free arrived, max xchg, exp done, pv_done woke, st done, p_ready arrived")
for j in range(2, 12): print(j, (t[0, j, :9] - t0).tolist())
print("mma thread: iter start, k_full, s_free woke, qk issued, p_ready woke, pv issued")
for j in range(2, 12): print(j, (t[1, j, :6] - t0).tolist())

x = np.array(buf[506:512])
print("CTA(0,0,0) thread 64: entry, setup done, softmax loop done, output stored, after final sync (cycles rel. to entry):", (x[:5] - x[0]).tolist())
print("first softmax stamp rel. to entry:", int(t[0, 0, 0] - x[0]), " last p_ready arrive:", int(t[0, 15, 8] - x[0]))
