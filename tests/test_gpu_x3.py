"""Kernel-level parity of the fp32-accurate tensor-core mode (LGB200_F32X3: split-fp16 planes, three tcgen05 MMAs per
product; csrc/lg_x3.cu, csrc/lg_x3_attn.cu) against float64 torch on the same inputs.

Tolerances: the tensor core rounds its fp32 accumulator toward zero after every MMA (tools/x3_micro.py), so these
kernels sit at 1-3e-6 relative rms -- about three times the error of an fp32 FMA chain, 1000 times below bf16.  Asserted:
|d| < 2e-5 + 2e-5 |ref| element-wise for the linear layers (values of order 1), 3e-5 for attention outputs."""
import math

import pytest
import torch

from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import EPI_HEADS, EPI_LN_GELU, EPI_ROWMAJOR, F32X3, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EA, EW = 64.0, 256.0  # plane scalings LG_X3_EA / LG_X3_EW (csrc/lg_internal.cuh)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def split(t, scale):
    t = t.float() * scale
    hi = t.half()
    lo = (t - hi.float()).half()
    return torch.stack([hi, lo]).contiguous()


def join(planes, scale):
    return (planes[0].double() + planes[1].double()) / scale


def test_split_rows_kernel_is_the_host_split():
    lib = _abi.load()
    x = torch.randn(1000, 256, device=DEV) * torch.logspace(-4, 2, 256, device=DEV)
    xs = torch.empty(2, 1000, 256, device=DEV, dtype=torch.float16)
    assert lib.lgb200_split_rows(ptr(x), x.numel(), ptr(xs), _stream()) == 0
    assert torch.equal(xs, split(x, EA))
    # hi + lo reproduces x to 22 bits where the low plane is a normal number
    big = x.abs() > 1e-2
    rel = ((join(xs, EA) - x.double()).abs() / x.double().abs())[big]
    assert rel.max() < 2 ** -21


def test_x3_linear_epilogues_against_fp64():
    lib = _abi.load()
    torch.manual_seed(0)
    S, Lp = 4, 256
    T = S * Lp
    lens = torch.tensor([256, 130, 128, 1], dtype=torch.int32, device=DEV)
    tile_valid = (torch.arange(Lp)[None] // 128 * 128 < lens.cpu()[:, None]).reshape(-1)
    x = torch.randn(T, 256, device=DEV)
    y = torch.randn(T, 256, device=DEV) * 0.3
    xs, ys = split(x, EA), split(y, EA)
    st = _stream()

    def lin(epi, A0, W, b, N, K, A1=None, K0=None, scale=(1., 1., 1.), resid=None, out32=None, outs=None, rot=None,
            n_rot=0, outp=(None, None, None), gamma=None, beta=None):
        rc = lib.lgb200_linear(F32X3, epi, ptr(A0), ptr(A1), K if K0 is None else K0, ptr(W), ptr(b), T, N, K,
                               ptr(lens), Lp, scale[0], scale[1], scale[2], ptr(resid), None, ptr(out32), ptr(outs),
                               ptr(rot), None, n_rot, ptr(outp[0]), ptr(outp[1]), ptr(outp[2]), ptr(gamma), ptr(beta), st)
        assert rc == 0, lib.lgb200_error_string(rc)

    def close(got, ref, what, tol=2e-5):
        d = (got.double() - ref).abs()
        assert (d <= tol + tol * ref.abs()).all(), f"{what}: max |d| {d.max():.2e}"

    # ROWMAJOR, K = 512 from two sources, scale, fp32 residual in place, fp32 + split outputs (FFN layer 2)
    W = torch.randn(256, 512, device=DEV) / 22
    b = torch.randn(256, device=DEV)
    resid = torch.randn(T, 256, device=DEV)
    io = resid.clone()
    outs = torch.zeros(2, T, 256, device=DEV, dtype=torch.float16)
    lin(EPI_ROWMAJOR, xs, split(W, EW), b, 256, 512, A1=ys, K0=256, scale=(0.25, 1, 1), resid=io, out32=io, outs=outs)
    ref = (torch.cat([x, y], 1).double() @ W.double().t() + b.double()) * 0.25 + resid.double()
    close(io[tile_valid], ref[tile_valid], "ROWMAJOR fp32 out")
    close(join(outs, EA)[tile_valid], ref[tile_valid], "ROWMAJOR split out")
    assert torch.equal(io[~tile_valid], resid[~tile_valid]), "tiles past lens must be skipped"
    # K = 128 (input_proj shape)
    xin = torch.randn(T, 128, device=DEV)
    W3 = torch.randn(256, 128, device=DEV) / 11
    o3 = torch.zeros(T, 256, device=DEV)
    lin(EPI_ROWMAJOR, split(xin, EA), split(W3, EW), b, 256, 128, out32=o3)
    close(o3[tile_valid], (xin.double() @ W3.double().t() + b.double())[tile_valid], "ROWMAJOR K=128")

    # HEADS with rotary on parts 0, 1 (self-attention Wqkv) and without (cross projections)
    W = torch.randn(768, 256, device=DEV) / 16
    b = torch.randn(768, device=DEV)
    ang = torch.randn(T, 32, device=DEV)
    rot = torch.stack([ang.cos(), ang.sin()], -1).reshape(T, 64).contiguous()
    yref = (x.double() @ W.double().t() + b.double()).view(S, Lp, 3, 4, 64).permute(2, 0, 3, 1, 4)  # [part,S,h,Lp,64]
    c = ang.double().cos().repeat_interleave(2, -1).view(S, 1, Lp, 64)
    s_ = ang.double().sin().repeat_interleave(2, -1).view(S, 1, Lp, 64)

    def rotf(t):
        t2 = t.unflatten(-1, (-1, 2))
        r = torch.stack((-t2[..., 1], t2[..., 0]), -1).flatten(-2)
        return t * c + r * s_

    tv = tile_valid.view(S, 1, Lp, 1).to(DEV)
    parts = [torch.zeros(2, S, 4, Lp, 64, device=DEV, dtype=torch.float16) for _ in range(3)]
    lin(EPI_HEADS, xs, split(W, EW), b, 768, 256, scale=(0.5, 1.0, 2.0), n_rot=2, rot=rot, outp=parts)
    for o, r in zip(parts, [rotf(yref[0]) * 0.5, rotf(yref[1]), yref[2] * 2.0]):
        close(torch.where(tv, join(o, EA), 0), torch.where(tv, r, 0), "HEADS")
    parts = [torch.zeros(2, S, 4, Lp, 64, device=DEV, dtype=torch.float16) for _ in range(2)]
    lin(EPI_HEADS, xs, split(W[:512], EW), b[:512].contiguous(), 512, 256, scale=(0.5, 2.0, 1.0), n_rot=0,
        outp=(parts[0], parts[1], None))
    for o, r in zip(parts, [yref[0] * 0.5, yref[1] * 2.0]):
        close(torch.where(tv, join(o, EA), 0), torch.where(tv, r, 0), "HEADS, no rotary")

    # LayerNorm + GELU(erf) (FFN layer 1 on cat[x, msg])
    W = torch.randn(512, 512, device=DEV) / 22
    b = torch.randn(512, device=DEV)
    gamma = torch.rand(512, device=DEV) + 0.5
    beta = torch.randn(512, device=DEV) * 0.1
    hs = torch.zeros(2, T, 512, device=DEV, dtype=torch.float16)
    lin(EPI_LN_GELU, xs, split(W, EW), b, 512, 512, A1=ys, K0=256, gamma=gamma, beta=beta, outs=hs)
    pre = torch.cat([x, y], 1).double() @ W.double().t() + b.double()
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(pre, (512,), gamma.double(), beta.double(), 1e-5))
    close(join(hs, EA)[tile_valid], ref[tile_valid], "LN + GELU")
    assert (hs[:, ~tile_valid] == 0).all()


@pytest.mark.parametrize("kv_xor", [0, 1])
def test_x3_attention_against_fp64(kv_xor):
    lib = _abi.load()
    torch.manual_seed(1)
    S, Lp = 4, 384
    lens = torch.tensor([384, 200, 129, 77], dtype=torch.int32, device=DEV)
    q = torch.randn(S, 4, Lp, 64, device=DEV) * 1.5
    k = torch.randn(S, 4, Lp, 64, device=DEV)
    v = torch.randn(S, 4, Lp, 64, device=DEV)
    k[0, :, 300] *= 6.0  # a key whose logits tower over the earlier tiles: the reference maximum must move
    qs, ks, vs = (split(t.reshape(-1, 64), EA) for t in (q, k, v))
    ctx = torch.zeros(2, S * Lp, 256, device=DEV, dtype=torch.float16)
    rc = lib.lgb200_attention(F32X3, ptr(qs), ptr(ks), ptr(vs), S, Lp, ptr(lens), kv_xor, ptr(ctx), _stream())
    assert rc == 0, lib.lgb200_error_string(rc)
    got = join(ctx, EA).view(S, Lp, 256)
    for s in range(S):
        nq, skv = int(lens[s]), s ^ kv_xor
        nk = int(lens[skv])
        sc = q[s, :, :nq].double() @ k[skv, :, :nk].double().transpose(-1, -2) * math.log(2.0)  # exp2 domain
        ref = (torch.softmax(sc, -1) @ v[skv, :, :nk].double()).permute(1, 0, 2).reshape(nq, 256)
        d = (got[s, :nq] - ref).abs().max()
        assert d < 3e-5, f"sequence {s}: max |d ctx| = {d:.2e}"
    # no keys at all: zeros
    lens0 = torch.tensor([100, 0, 0, 50], dtype=torch.int32, device=DEV)
    ctx.fill_(1.0)
    assert lib.lgb200_attention(F32X3, ptr(qs), ptr(ks), ptr(vs), S, Lp, ptr(lens0), 1, ptr(ctx), _stream()) == 0
    assert (ctx.view(2, S, Lp, 256)[:, 0, :100] == 0).all() and (ctx.view(2, S, Lp, 256)[:, 3, :50] == 0).all()


def test_x3_similarity_and_assignment_against_fp64():
    """lgb200_x3_similarity + lgb200_x3_assign_lse + lgb200_x3_assign_scores == sigmoid_log_double_softmax
    (lightglue.py:257-269) of the fp64 similarity, with per-pair counts."""
    lib = _abi.load()
    torch.manual_seed(2)
    B, Lp = 3, 512
    n0, n1 = [512, 300, 257], [400, 512, 1]
    lens = torch.tensor([v for p in zip(n0, n1) for v in p], dtype=torch.int32, device=DEV)
    md = torch.randn(2 * B * Lp, 256, device=DEV) * 0.6
    z = torch.randn(2 * B * Lp, device=DEV)
    mds = split(md, EA)
    sim = torch.full((B, Lp, Lp), float("nan"), device=DEV)
    st = _stream()
    assert lib.lgb200_x3_similarity(ptr(mds), B, Lp, ptr(lens), ptr(sim), st) == 0
    lse = torch.zeros(2 * B * Lp, device=DEV)
    assert lib.lgb200_x3_assign_lse(ptr(sim), B, Lp, ptr(lens), Lp, Lp, ptr(lse), st) == 0
    R, C = Lp + 1, Lp + 1
    scores = torch.full((B, R, C), float("nan"), device=DEV)
    assert lib.lgb200_x3_assign_scores(ptr(sim), ptr(z), ptr(lse), B, Lp, ptr(lens), R, C, ptr(scores), st) == 0
    mdv, zv = md.view(B, 2, Lp, 256).double(), z.view(B, 2, Lp).double()
    ls = torch.nn.functional.logsigmoid
    for b in range(B):
        a, c = n0[b], n1[b]
        s64 = mdv[b, 0, :a] @ mdv[b, 1, :c].t()
        assert (sim[b, :a, :c].double() - s64).abs().max() < 2e-5
        ref = (torch.log_softmax(s64, 1) + torch.log_softmax(s64, 0) + ls(zv[b, 0, :a])[:, None] + ls(zv[b, 1, :c])[None])
        got = scores[b].double()
        assert (got[:a, :c] - ref).abs().max() < 1e-4
        assert (got[:a, C - 1] - ls(-zv[b, 0, :a])).abs().max() < 1e-6
        assert (got[R - 1, :c] - ls(-zv[b, 1, :c])).abs().max() < 1e-6
        assert (got[a:R - 1] == 0).all() and (got[:a, c:C - 1] == 0).all() and got[R - 1, C - 1] == 0
