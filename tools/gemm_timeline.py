"""Wait-cycle accounting of one v2 GEMM launch (library must be built with -DLG_GEMM_DEBUG)."""
import ctypes, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, EPI_LN_GELU, EPI_ROWMAJOR, EPI_HEADS, ptr
import os
lib = _abi.load(Path(os.environ['LGB200_LIB']).resolve()) if os.environ.get('LGB200_LIB') else _abi.load()
lib.lgb200_debug_gemm_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
S, Lp = 128, 2048
T = S * Lp
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(T, 256, device="cuda").to(torch.bfloat16); y = torch.randn_like(x)
def run(name, epi, N, K, **kw):
    W = (torch.randn(N, K, device="cuda") / 16).to(torch.bfloat16); b = torch.randn(N, device="cuda")
    args = dict(A0=x, A1=None, K0=K, resid16=None, out16=None, rot16=None, n_rot=0, outp=(None, None, None), gamma=None, beta=None)
    args.update(kw)
    def call():
        rc = lib.lgb200_linear(BF16, epi, ptr(args["A0"]), ptr(args["A1"]), args["K0"], ptr(W), ptr(b), T, N, K, None, Lp, 1.0, 1.0, 1.0,
                               None, ptr(args["resid16"]), None, ptr(args["out16"]), None, ptr(args["rot16"]), args["n_rot"],
                               ptr(args["outp"][0]), ptr(args["outp"][1]), ptr(args["outp"][2]), ptr(args["gamma"]), ptr(args["beta"]), st)
        assert rc == 0, rc
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)(); lib.lgb200_debug_gemm_times(buf, 16)
    t = list(buf)
    ms = e0.elapsed_time(e1) / 5
    tiles = max(t[5], 1)
    print(f"{name}: {ms*1e3:.1f} us  {2.0*T*N*K/ms/1e9:.0f} TFLOP/s | per tile (cycles): mma-warp total {t[2]/tiles:.0f} = wait-acc {t[0]/tiles:.0f} + wait-A {t[1]/tiles:.0f} + issue; "
          f"epilogue total {t[4]/tiles:.0f}, wait-tfull {t[3]/tiles:.0f}, wait-stats {t[6]/tiles:.0f}, wait-peer {t[7]/tiles:.0f}  (tiles/CTA {tiles})")
hid = torch.empty(T, 512, device="cuda", dtype=torch.bfloat16)
g = torch.ones(512, device="cuda"); be = torch.zeros(512, device="cuda")
run("FFN1 LN  K512 N512", EPI_LN_GELU, 512, 512, A1=y, K0=256, out16=hid, gamma=g, beta=be)
o = torch.empty(T, 256, device="cuda", dtype=torch.bfloat16)
run("FFN2 ROW K512 N256", EPI_ROWMAJOR, 256, 512, A0=hid, resid16=o, out16=o)
run("out  ROW K256 N256", EPI_ROWMAJOR, 256, 256, out16=o)
q2 = [torch.empty(T * 256, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
run("cross HEADS K256 N512", EPI_HEADS, 512, 256, n_rot=0, outp=(q2[0], q2[1], None))
q = [torch.empty(T * 256, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
rot16 = torch.randn(T, 64, device="cuda").to(torch.float16)
run("QKV HEADS K256 N768", EPI_HEADS, 768, 256, rot16=rot16, n_rot=2, outp=q)
