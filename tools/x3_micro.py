"""Numerics of the x3 kernels in isolation: linear (K = 256 / 512) and attention against fp64."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import F32X3, F32, EPI_ROWMAJOR, ptr
lib = _abi.load()
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev).manual_seed(1)

def split(t, scale):
    t = (t.float() * scale)
    hi = t.half(); lo = (t - hi.float()).half()
    return torch.stack([hi, lo]).contiguous()

def stats(tag, got, ref):
    d = (got.double() - ref)
    scale = ref.abs().mean()
    print(f"{tag}: rms err {d.pow(2).mean().sqrt()/scale:.2e} (rel. to mean |ref| {scale:.3f})  max {d.abs().max()/scale:.2e}  "
          f"bias toward zero {(d*ref.sign()).mean()/scale:+.2e}")

for K in (256, 512):
    T, N, Lp = 2048, 256, 1024
    A = torch.randn(T, K, device=dev, generator=g)
    W = (torch.rand(N, K, device=dev, generator=g) * 2 - 1) / K ** 0.5
    b = torch.zeros(N, device=dev)
    ref = A.double() @ W.double().T
    out = torch.empty(T, N, device=dev)
    As, Ws = split(A, 64.0), split(W, 256.0)
    rc = lib.lgb200_linear(F32X3, EPI_ROWMAJOR, ptr(As), None, K, ptr(Ws), ptr(b), T, N, K, None, Lp, 1.0, 1.0, 1.0,
                           None, None, ptr(out), None, None, None, 0, None, None, None, None, None, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    stats(f"x3 linear K={K}", out, ref)
    out2 = torch.empty(T, N, device=dev)
    rc = lib.lgb200_linear(F32, EPI_ROWMAJOR, ptr(A), None, K, ptr(W), ptr(b), T, N, K, None, Lp, 1.0, 1.0, 1.0,
                           None, None, ptr(out2), None, None, None, 0, None, None, None, None, None, st)
    torch.cuda.synchronize()
    stats(f"simt linear K={K}", out2, ref)
    stats(f"torch fp32 matmul K={K}", A @ W.T, ref)
    # representation error only: fp64 product of the planes
    Ar = (As[0].double() + As[1].double()) / 64; Wr = (Ws[0].double() + Ws[1].double()) / 256
    stats(f"planes in fp64 K={K}", (Ar @ Wr.T), ref)
    # positive operands: the accumulator only grows -> a truncating adder shows up as a one-sided bias
    Ap, Wp = A.abs(), W.abs()
    refp = Ap.double() @ Wp.double().T
    Aps, Wps = split(Ap, 64.0), split(Wp, 256.0)
    lib.lgb200_linear(F32X3, EPI_ROWMAJOR, ptr(Aps), None, K, ptr(Wps), ptr(b), T, N, K, None, Lp, 1.0, 1.0, 1.0,
                      None, None, ptr(out), None, None, None, 0, None, None, None, None, None, st)
    torch.cuda.synchronize()
    stats(f"x3 linear K={K}, positive operands", out, refp)

# attention
S, Lp = 2, 2048
q = torch.randn(S, 4, Lp, 64, device=dev, generator=g) * 0.6
k = torch.randn(S, 4, Lp, 64, device=dev, generator=g)
v = torch.randn(S, 4, Lp, 64, device=dev, generator=g)
sc = (q.double() @ k.double().transpose(-1, -2)) * 0.6931471805599453  # log2 domain -> natural
ref = torch.softmax(sc, -1) @ v.double()
ref = ref.permute(0, 2, 1, 3).reshape(S * Lp, 256)
qs, ks, vs = (split(t.reshape(-1, 64), 64.0) for t in (q, k, v))
ctx = torch.empty(2, S * Lp, 256, device=dev, dtype=torch.float16)
rc = lib.lgb200_attention(F32X3, ptr(qs), ptr(ks), ptr(vs), S, Lp, None, 0, ptr(ctx), st)
assert rc == 0, rc
torch.cuda.synchronize()
stats("x3 attention 2048 keys", (ctx[0].float() + ctx[1].float()) / 64, ref)
ctx32 = torch.empty(S * Lp, 256, device=dev)
rc = lib.lgb200_attention(F32, ptr(q.contiguous()), ptr(k.contiguous()), ptr(v.contiguous()), S, Lp, None, 0, ptr(ctx32), st)
torch.cuda.synchronize()
stats("simt attention 2048 keys", ctx32, ref)
