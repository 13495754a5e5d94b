"""Training path of the B200 LightGlue drop-in: forward that keeps what the backward needs, and the backward pass.

The reference trains through `LightGlue.forward` / `LightGlue.loss` with autograd (lightglue.py:484-498 collects every
layer's descriptors in training mode, :588-637 turns them into the deep-supervision NLL; train.py calls
`loss.backward()`).  Here the same contract is met by two `torch.autograd.Function`s whose backward is written by hand:

  * `TransformerFn`  descriptors (+ every transformer / posenc / input_proj parameter) -> ref_descriptors0/1
    [B, n_layers, N, 256].  Forward = the library's fp32-accurate tensor-core kernels (linear layers and attention in the
    split-fp16 / three-MMA mode of lg_x3.cu / lg_x3_attn.cu; LGB200_TRAIN_SIMT_LINEAR=1 selects the CUDA-core fp32 linear
    kernels), keeping the two block inputs of every layer plus -- unless conf.checkpointed
    (lightglue.py:485-494) -- q / k / v / context / message; with conf.checkpointed those are recomputed in the
    backward pass with the same kernels.  The pre-LayerNorm activations are always recomputed (one GEMM).  Backward =
    the hand-written kernels of csrc/lg_bwd.cu:
    flash-attention backward (lgb200_attention_bwd), rotary / head-split backward incl. the gradient of the rotary
    angles (lgb200_heads_bwd), GELU . LayerNorm backward (lgb200_ln_gelu_bwd).
  * `AssignFn`       MatchAssignment of one layer (lightglue.py:272-288) reduced to what NLLLoss reads
    (losses.py:6-26: the sum of log_assignment over the ground-truth matches and the two dustbin vectors).  Forward =
    the library's assignment + reduction kernels; backward = lgb200_assign_dsim (d similarity of the double softmax).

Plain GEMMs of the backward pass (dX = dY.W, dW = dY^T.X) go to cuBLAS -- they are library GEMMs with nothing to fuse --
as three fp16 tensor-core GEMMs on split planes with fp32 accumulation (_Kern.mm3; LGB200_TRAIN_SGEMM=1: one fp32
sgemm on the CUDA cores).  Gradients span a range (1e-8 ... 1e-2) that a fixed plane scaling does not cover, so
lgb200_split_dynamic scales every gradient tensor by its own power of two first.  Everything is fp32-accurate whatever conf.precision says (the bf16 kernels are inference kernels); CPU tensors
raise as everywhere else in this package.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Tuple

import torch
from torch import nn

from . import _abi
from ._abi import EPI_HEADS, EPI_LN_GELU, EPI_ROWMAJOR, F32, F32X3, check, ptr

LOG2E = 1.4426950408889634
Q_SCALE = LOG2E / math.sqrt(64.0)   # folded into q: the attention kernels work in the log2 domain
C_SCALE = math.sqrt(Q_SCALE)        # cross block: to_qk feeds both sides (lightglue.py:208)
AI, WI = 1.0 / 64.0, 1.0 / 256.0       # inverse plane scalings of activations / weights (LG_X3_EA, LG_X3_EW)
N_PARTIALS = 296                    # CTAs of lgb200_ln_gelu_bwd (2 per SM)
_SIMT_ATTN = os.environ.get("LGB200_TRAIN_SIMT_ATTN", "0") == "1"
_SIMT_LINEAR = os.environ.get("LGB200_TRAIN_SIMT_LINEAR", "0") == "1"
# plain GEMMs of the backward pass: 1 = cuBLAS fp32 (sgemm on the CUDA cores), default = three fp16 tensor-core GEMMs on
# split planes with fp32 accumulation (the same hi.hi + hi.lo + lo.hi scheme as the forward's lg_x3.cu kernels)
def _has_mm_out_dtype() -> bool:
    """torch.mm / addmm / bmm with out_dtype (fp16 operands, fp32 accumulation AND fp32 output) exist from PyTorch 2.8 on."""
    try:
        a = torch.empty(1, 1, device="meta", dtype=torch.float16)
        return torch.mm(a, a, out_dtype=torch.float32).dtype == torch.float32
    except (TypeError, RuntimeError, NotImplementedError):
        return False


_SGEMM = os.environ.get("LGB200_TRAIN_SGEMM", "0") == "1" or _SIMT_LINEAR or not _has_mm_out_dtype()
_X3_WEIGHTS = ("qkv_w", "so_w", "sf0_w", "sf3_w", "cqv_w", "co_w", "cf0_w", "cf3_w")
# packed Wqkv row r holds reference row _PERM[r] (lightglue.py:158: head*192 + d*3 + part -> part*256 + head*64 + d)
_PERM = torch.arange(768).view(4, 64, 3).permute(2, 0, 1).reshape(-1)


class _Kern:
    """Launch helpers for one (B, m, n) problem: S = 2B sequences padded to Lp rows, lens on the device."""

    def __init__(self, dev, B: int, m: int, n: int):
        self.lib = _abi.load()
        check(self.lib.lgb200_device_ok(), "device check")
        self.dev, self.B, self.m, self.n = dev, B, m, n
        self.S = 2 * B
        self.Lp = max(128, ((max(m, n) + 127) // 128) * 128)
        self.T = self.S * self.Lp
        self.lens = None
        if m != self.Lp or n != self.Lp:
            self.lens = torch.tensor([m, n] * B, device=dev, dtype=torch.int32)

    @property
    def st(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def zeros(self, *shape):
        return torch.zeros(*shape, device=self.dev, dtype=torch.float32)

    def empty(self, *shape):
        return torch.empty(*shape, device=self.dev, dtype=torch.float32)

    def linear(self, epi, A0, W, b, N, K, A1=None, K0=None, scale=(1.0, 1.0, 1.0), resid=None, out=None, rot=None,
               n_rot=0, outp=(None, None, None), gamma=None, beta=None):
        check(self.lib.lgb200_linear(F32, epi, ptr(A0), ptr(A1), K if K0 is None else K0, ptr(W), ptr(b), self.T, N, K,
                                     ptr(self.lens), self.Lp, scale[0], scale[1], scale[2], ptr(resid), None, ptr(out),
                                     None, ptr(rot), None, n_rot, ptr(outp[0]), ptr(outp[1]), ptr(outp[2]), ptr(gamma),
                                     ptr(beta), self.st), "lgb200_linear")

    def linear3(self, epi, A0, W, b, N, K, A1=None, K0=None, scale=(1.0, 1.0, 1.0), resid=None, out=None, out32=None,
                rot=None, n_rot=0, outp=(None, None, None), gamma=None, beta=None):
        """The same layer in the fp32-accurate tensor-core mode (lg_x3.cu): A0 / A1 / W / out / outp are split-fp16
        planes, resid / out32 / rot fp32."""
        check(self.lib.lgb200_linear(F32X3, epi, ptr(A0), ptr(A1), K if K0 is None else K0, ptr(W), ptr(b), self.T, N, K,
                                     ptr(self.lens), self.Lp, scale[0], scale[1], scale[2], ptr(resid), None, ptr(out32),
                                     ptr(out), ptr(rot), None, n_rot, ptr(outp[0]), ptr(outp[1]), ptr(outp[2]), ptr(gamma),
                                     ptr(beta), self.st), "lgb200_linear (x3)")

    def planes(self, *shape):
        return torch.zeros(2, *shape, device=self.dev, dtype=torch.float16)

    def unsplit(self, p, *shape, out=None):
        """split planes (x * 64 = hi + lo) -> fp32"""
        n = p.numel() // 2
        out = torch.empty(n, device=self.dev, dtype=torch.float32) if out is None else out
        check(self.lib.lgb200_merge_rows(ptr(p), n, AI, ptr(out), self.st), "lgb200_merge_rows")
        return out.view(*shape) if shape else out

    def split(self, t):
        """fp32 -> split-fp16 planes [2][n] (x * 64 = hi + lo), the operand format of the LGB200_F32X3 kernels."""
        out = torch.empty(2, t.numel(), device=self.dev, dtype=torch.float16)
        check(self.lib.lgb200_split_rows(ptr(t), t.numel(), ptr(out), self.st), "lgb200_split_rows")
        return out

    def asplit(self, t):
        """activation [M, N] fp32 -> planes [2, M, N] (scale 64)"""
        return self.split(t).view(2, *t.shape)

    def gsplit(self, t, colsum: bool = False):
        """gradient [M, N] fp32 -> (planes [2, M, N] of g t, 1 / g on the device): lgb200_split_dynamic picks the power of
        two g per tensor (gradients have no fixed range).  colsum: also t.sum(0) (the bias gradient), from per-CTA
        partials of the same pass over t."""
        t = t.contiguous()
        out = torch.empty(2, *t.shape, device=self.dev, dtype=torch.float16)
        inv = torch.empty(2, device=self.dev, dtype=torch.float32)
        part = torch.empty(N_PARTIALS, t.shape[-1], device=self.dev, dtype=torch.float32) if colsum else None
        check(self.lib.lgb200_split_dynamic(ptr(t), t.numel(), ptr(out), ptr(inv), t.shape[-1] if colsum else 0, ptr(part),
                                            N_PARTIALS if colsum else 0, self.st), "lgb200_split_dynamic")
        return (out, inv[0], part.sum(0)) if colsum else (out, inv[0])

    @staticmethod
    def mm3(A, B, scale):
        """fp32-accurate A.B from plane pairs A [2, M, K], B [2, K, N] (views may be transposed): lo.hi + hi.lo + hi.hi as
        three fp16 tensor-core GEMMs with fp32 accumulation and fp32 output (cuBLAS: plain library GEMMs), times `scale`
        (the product of the inverse plane scalings; float or 0-dim device tensor)."""
        f32 = torch.float32
        r = torch.mm(A[1], B[0], out_dtype=f32)
        # beta = 1: accumulated in the GEMM epilogue; out = r: in place (without it ATen copies r into a new output first)
        torch.addmm(r, A[0], B[1], out_dtype=f32, out=r)
        torch.addmm(r, A[0], B[0], out_dtype=f32, out=r)
        return r.mul_(scale)

    def attention(self, q, k, v, kv_xor, ctx):
        """Forward attention on the tensor cores in the fp32-accurate mode (lg_x3_attn.cu: split-fp16 operands, three
        tcgen05 MMAs per product, fp32 softmax) -- 12x faster than the CUDA-core fp32 kernel at 2048 keypoints, same
        1e-6-level agreement with float64.  LGB200_TRAIN_SIMT_ATTN=1 selects the CUDA-core kernel."""
        if _SIMT_ATTN:
            check(self.lib.lgb200_attention(F32, ptr(q), ptr(k), ptr(v), self.S, self.Lp, ptr(self.lens), kv_xor, ptr(ctx),
                                            self.st), "lgb200_attention")
            return
        if q.dtype == torch.float16:  # already split planes (x3 projections)
            qp, kp, vp = q, k, v
        else:
            qp = self.split(q)
            kp = qp if k is q else self.split(k)
            vp = self.split(v)
        cp = torch.zeros(2, self.T, 256, device=self.dev, dtype=torch.float16)
        check(self.lib.lgb200_attention(F32X3, ptr(qp), ptr(kp), ptr(vp), self.S, self.Lp, ptr(self.lens), kv_xor, ptr(cp),
                                        self.st), "lgb200_attention")
        self.unsplit(cp, out=ctx)
        return cp

    def attention_bwd(self, q, k, v, ctx, dctx, kv_xor):
        dq, dk, dv = self.empty(self.T * 256), self.empty(self.T * 256), self.empty(self.T * 256)
        n_ws = C.c_longlong(0)
        check(self.lib.lgb200_attention_bwd_workspace(self.S, self.Lp, C.byref(n_ws)), "lgb200_attention_bwd_workspace")
        ws = self.empty(n_ws.value)
        check(self.lib.lgb200_attention_bwd(ptr(q), ptr(k), ptr(v), ptr(ctx), ptr(dctx), self.S, self.Lp, ptr(self.lens),
                                            kv_xor, ptr(dq), ptr(dk), ptr(dv), ptr(ws), self.st), "lgb200_attention_bwd")
        return dq, dk, dv

    def heads_bwd(self, dq, dk, dv, q, k, rot, n_parts, scales, dtheta):
        out = self.empty(self.T, 256 * n_parts)
        check(self.lib.lgb200_heads_bwd(ptr(dq), ptr(dk), ptr(dv), ptr(q), ptr(k), ptr(rot), self.S, self.Lp,
                                        ptr(self.lens), n_parts, scales[0], scales[1], scales[2], ptr(out), ptr(dtheta),
                                        self.st), "lgb200_heads_bwd")
        return out

    def ln_gelu_bwd(self, h, gamma, beta, da):
        dh, act, part = self.empty(self.T, 512), self.empty(self.T, 512), self.empty(N_PARTIALS, 1024)
        check(self.lib.lgb200_ln_gelu_bwd(ptr(h), ptr(gamma), ptr(beta), ptr(da), self.T, self.Lp, ptr(self.lens), ptr(dh),
                                          ptr(act), ptr(part), N_PARTIALS, self.st), "lgb200_ln_gelu_bwd")
        return dh, act, part.sum(0)

    def pack(self, src, img, dst):
        """src [B, cnt, dim] -> rows of sequences 2b + img of dst [T, dim] (rows past cnt zero-filled)."""
        src = src.to(torch.float32).contiguous()
        check(self.lib.lgb200_pack_rows(ptr(src), self.B, src.shape[1], src.shape[2], img, self.Lp, ptr(dst), None,
                                        self.st), "lgb200_pack_rows")

    def pack2(self, a0, a1):
        dst = self.zeros(self.T, a0.shape[-1])
        self.pack(a0, 0, dst)
        self.pack(a1, 1, dst)
        return dst

    def unpack(self, x):
        xv = x.view(self.B, 2, self.Lp, x.shape[-1])
        return xv[:, 0, : self.m], xv[:, 1, : self.n]


def normalize_keypoints(kpts: torch.Tensor, size) -> torch.Tensor:
    """lightglue.py:28-40 (only needed for the gradient of posenc.Wr; the forward's posenc kernel has it fused)."""
    if size is None:
        size = 1 + kpts.max(-2).values - kpts.min(-2).values
    size = size.to(kpts)
    shift = size / 2
    scale = size.max(-1).values / 2
    return (kpts - shift[..., None, :]) / scale[..., None, None]


def transformer_params(model) -> List[Tuple[object, nn.Parameter]]:
    """(key, parameter) of everything TransformerFn differentiates, in a fixed order."""
    out: List[Tuple[object, nn.Parameter]] = []
    if isinstance(model.input_proj, nn.Linear):
        out += [("in_w", model.input_proj.weight), ("in_b", model.input_proj.bias)]
    out.append(("wr", model.posenc.Wr.weight))
    for i, lyr in enumerate(model.transformers):
        sa, ca = lyr.self_attn, lyr.cross_attn
        out += [
            ((i, "qkv_w"), sa.Wqkv.weight), ((i, "qkv_b"), sa.Wqkv.bias),
            ((i, "so_w"), sa.out_proj.weight), ((i, "so_b"), sa.out_proj.bias),
            ((i, "sf0_w"), sa.ffn[0].weight), ((i, "sf0_b"), sa.ffn[0].bias),
            ((i, "sln_g"), sa.ffn[1].weight), ((i, "sln_b"), sa.ffn[1].bias),
            ((i, "sf3_w"), sa.ffn[3].weight), ((i, "sf3_b"), sa.ffn[3].bias),
            ((i, "cqk_w"), ca.to_qk.weight), ((i, "cqk_b"), ca.to_qk.bias),
            ((i, "cv_w"), ca.to_v.weight), ((i, "cv_b"), ca.to_v.bias),
            ((i, "co_w"), ca.to_out.weight), ((i, "co_b"), ca.to_out.bias),
            ((i, "cf0_w"), ca.ffn[0].weight), ((i, "cf0_b"), ca.ffn[0].bias),
            ((i, "cln_g"), ca.ffn[1].weight), ((i, "cln_b"), ca.ffn[1].bias),
            ((i, "cf3_w"), ca.ffn[3].weight), ((i, "cf3_b"), ca.ffn[3].bias),
        ]
    return out


# ---- one transformer layer, forward pieces (lightglue.py:151-164, 193-222) ---------------------------------------


def x3_weights(w: Dict) -> Dict:
    """Split-fp16 planes [2][N][K] of a layer's weight matrices (256 w = hi + lo, the operand format of lg_x3.cu)."""
    out = {}
    for name in _X3_WEIGHTS:
        t = w[name].detach().float() * 256.0
        hi = t.to(torch.float16)
        out[name] = torch.stack([hi, (t - hi.float()).to(torch.float16)]).contiguous()
    return out


def _self_attend(k: _Kern, w: Dict, x, rot, fp32_heads: bool = True):
    """-> (q, k, v, ctx, msg); with the tensor-core linears q / k / v are None unless `fp32_heads` (the backward kernels
    read them in fp32)."""
    T = k.T
    if _SIMT_LINEAR:
        q, kk, v = k.zeros(T * 256), k.zeros(T * 256), k.zeros(T * 256)
        k.linear(EPI_HEADS, x, w["qkv_w"], w["qkv_b"], 768, 256, scale=(Q_SCALE, 1.0, 1.0), rot=rot, n_rot=2,
                 outp=(q, kk, v))
        ctx = k.zeros(T, 256)
        k.attention(q, kk, v, 0, ctx)
        msg = k.zeros(T, 256)
        k.linear(EPI_ROWMAJOR, ctx, w["so_w"], w["so_b"], 256, 256, out=msg)
        return q, kk, v, ctx, msg
    wx = w["x3"]
    qp, kp, vp = k.planes(T * 256), k.planes(T * 256), k.planes(T * 256)
    k.linear3(EPI_HEADS, k.split(x), wx["qkv_w"], w["qkv_b"], 768, 256, scale=(Q_SCALE, 1.0, 1.0), rot=rot, n_rot=2,
              outp=(qp, kp, vp))
    ctx = k.zeros(T, 256)
    cp = k.attention(qp, kp, vp, 0, ctx)
    msg = k.zeros(T, 256)
    k.linear3(EPI_ROWMAJOR, cp, wx["so_w"], w["so_b"], 256, 256, out32=msg)
    if not fp32_heads:
        return None, None, None, ctx, msg
    return k.unsplit(qp, T * 256), k.unsplit(kp, T * 256), k.unsplit(vp, T * 256), ctx, msg


def _cross_attend(k: _Kern, w: Dict, x, fp32_heads: bool = True):
    T = k.T
    if _SIMT_LINEAR:
        qk, v = k.zeros(T * 256), k.zeros(T * 256)
        k.linear(EPI_HEADS, x, w["cqv_w"], w["cqv_b"], 512, 256, scale=(C_SCALE, 1.0, 1.0), n_rot=0, outp=(qk, v, None))
        ctx = k.zeros(T, 256)
        k.attention(qk, qk, v, 1, ctx)
        msg = k.zeros(T, 256)
        k.linear(EPI_ROWMAJOR, ctx, w["co_w"], w["co_b"], 256, 256, out=msg)
        return qk, v, ctx, msg
    wx = w["x3"]
    qp, vp = k.planes(T * 256), k.planes(T * 256)
    k.linear3(EPI_HEADS, k.split(x), wx["cqv_w"], w["cqv_b"], 512, 256, scale=(C_SCALE, 1.0, 1.0), n_rot=0,
              outp=(qp, vp, None))
    ctx = k.zeros(T, 256)
    cp = k.attention(qp, qp, vp, 1, ctx)
    msg = k.zeros(T, 256)
    k.linear3(EPI_ROWMAJOR, cp, wx["co_w"], w["co_b"], 256, 256, out32=msg)
    if not fp32_heads:
        return None, None, ctx, msg
    return k.unsplit(qp, T * 256), k.unsplit(vp, T * 256), ctx, msg


def _ffn(k: _Kern, w: Dict, pre: str, x, msg):
    """x + ffn(cat[x, msg]) (lightglue.py:164, 220-221)."""
    out = k.zeros(k.T, 256)
    if _SIMT_LINEAR:
        hid = k.zeros(k.T, 512)
        k.linear(EPI_LN_GELU, x, w[pre + "f0_w"], w[pre + "f0_b"], 512, 512, A1=msg, K0=256, gamma=w[pre + "ln_g"],
                 beta=w[pre + "ln_b"], out=hid)
        k.linear(EPI_ROWMAJOR, hid, w[pre + "f3_w"], w[pre + "f3_b"], 256, 512, resid=x, out=out)
        return out
    wx = w["x3"]
    hp = k.planes(k.T, 512)
    k.linear3(EPI_LN_GELU, k.split(x), wx[pre + "f0_w"], w[pre + "f0_b"], 512, 512, A1=k.split(msg), K0=256,
              gamma=w[pre + "ln_g"], beta=w[pre + "ln_b"], out=hp)
    k.linear3(EPI_ROWMAJOR, hp, wx[pre + "f3_w"], w[pre + "f3_b"], 256, 512, resid=x, out32=out)
    return out


def _ffn_bwd(k: _Kern, w: Dict, pre: str, x, msg, dy):
    """Backward of y = ffn(cat[x, msg]) given dy [T,256]: (d x, d msg, parameter gradients)."""
    h = k.zeros(k.T, 512)  # pre-LayerNorm activations, recomputed
    if _SIMT_LINEAR:
        k.linear(EPI_ROWMAJOR, x, w[pre + "f0_w"], w[pre + "f0_b"], 512, 512, A1=msg, K0=256, out=h)
    else:
        k.linear3(EPI_ROWMAJOR, k.split(x), w["x3"][pre + "f0_w"], w[pre + "f0_b"], 512, 512, A1=k.split(msg), K0=256,
                  out32=h)
    if _SGEMM:
        da = dy @ w[pre + "f3_w"]
        dh, act, gsum = k.ln_gelu_bwd(h, w[pre + "ln_g"], w[pre + "ln_b"], da)
        g3, g0 = dy.t() @ act, torch.cat([dh.t() @ x, dh.t() @ msg], 1)
        dcat = dh @ w[pre + "f0_w"]
        b3, b0 = dy.sum(0), dh.sum(0)
    else:
        wx = w["x3"]
        dyp, dyi, b3 = k.gsplit(dy, colsum=True)
        da = k.mm3(dyp, wx[pre + "f3_w"], dyi * WI)
        dh, act, gsum = k.ln_gelu_bwd(h, w[pre + "ln_g"], w[pre + "ln_b"], da)
        g3 = k.mm3(dyp.transpose(1, 2), k.asplit(act), dyi * AI)
        dhp, dhi, b0 = k.gsplit(dh, colsum=True)
        dht = dhp.transpose(1, 2)
        g0 = torch.cat([k.mm3(dht, k.asplit(x), dhi * AI), k.mm3(dht, k.asplit(msg), dhi * AI)], 1)
        dcat = k.mm3(dhp, wx[pre + "f0_w"], dhi * WI)
    g = {
        pre + "f3_w": g3, pre + "f3_b": b3,
        pre + "ln_g": gsum[:512], pre + "ln_b": gsum[512:],
        pre + "f0_w": g0, pre + "f0_b": b0,
    }
    return dcat[:, :256], dcat[:, 256:].contiguous(), g


def _proj_bwd(k: _Kern, w: Dict, name: str, dy, x):
    """Backward of y = x W^T + b for the weight `name` ([N, K], planes in w["x3"]): (dy^T x, dy^T 1, dy W)."""
    if _SGEMM:
        return dy.t() @ x, dy.sum(0), dy @ w[name]
    dyp, dyi, db = k.gsplit(dy, colsum=True)
    return k.mm3(dyp.transpose(1, 2), k.asplit(x), dyi * AI), db, k.mm3(dyp, w["x3"][name], dyi * WI)


class TransformerFn(torch.autograd.Function):
    """ref_descriptors0/1 = all layers' outputs of the transformer stack (lightglue.py:456-498)."""

    @staticmethod
    def forward(ctx, model, geom: Dict, desc0, desc1, *params):
        dev = desc0.device
        B, m, _ = desc0.shape
        n = desc1.shape[1]
        k = _Kern(dev, B, m, n)
        W = model._pack(F32, dev)
        L = model.conf.n_layers
        if not _SIMT_LINEAR:  # split-fp16 planes of the weight matrices, cached with the pack (rebuilt when a weight changes)
            for w in W["layers"]:
                if "x3" not in w:
                    w["x3"] = x3_weights(w)
        xin = None
        if isinstance(model.input_proj, nn.Linear):
            din = model.conf.input_dim
            xin = k.pack2(desc0, desc1)
            x = k.zeros(k.T, 256)
            k.linear(EPI_ROWMAJOR, xin, W["in_w"], W["in_b"], 256, din, out=x)
        else:
            x = k.pack2(desc0, desc1)
        rot = k.zeros(k.T, 64)
        kdim = geom["k0"].shape[-1]
        for img, kk, cnt, sz in ((0, geom["k0"], m, geom["size0"]), (1, geom["k1"], n, geom["size1"])):
            check(k.lib.lgb200_posenc(ptr(kk), B, cnt, kdim, ptr(sz), ptr(W["wr"]), ptr(k.lens), img, k.Lp, ptr(rot),
                                      None, k.st), "lgb200_posenc")
        # conf.checkpointed (lightglue.py:485-494): keep only the block inputs and recompute q / k / v / context /
        # message in the backward pass (26 x T KB per layer less); otherwise they are kept, as autograd would
        keep = not bool(model.conf.checkpointed)
        xs_in, xs_mid, outs, kept = [], [], [], []
        for i in range(L):
            w = W["layers"][i]
            xs_in.append(x)
            sa = _self_attend(k, w, x, rot, fp32_heads=keep)
            x = _ffn(k, w, "s", x, sa[-1])
            xs_mid.append(x)
            ca = _cross_attend(k, w, x, fp32_heads=keep)
            x = _ffn(k, w, "c", x, ca[-1])
            outs.append(x)
            kept.append((sa, ca) if keep else None)
        ctx.k, ctx.W, ctx.geom, ctx.model = k, W, geom, model
        ctx.xs_in, ctx.xs_mid, ctx.rot, ctx.xin, ctx.kept = xs_in, xs_mid, rot, xin, kept
        ctx.keys = [key for key, _ in transformer_params(model)]
        r0 = torch.stack([k.unpack(o)[0] for o in outs], 1)
        r1 = torch.stack([k.unpack(o)[1] for o in outs], 1)
        return r0, r1

    @staticmethod
    def backward(ctx, g0, g1):
        k, W, geom = ctx.k, ctx.W, ctx.geom
        L = len(ctx.xs_in)
        rot = ctx.rot
        dx = k.zeros(k.T, 256)
        dtheta = k.zeros(k.T, 32)
        grads: Dict[object, torch.Tensor] = {}
        for i in reversed(range(L)):
            w = W["layers"][i]
            dx = dx + k.pack2(g0[:, i], g1[:, i])
            x_in, x_mid = ctx.xs_in[i], ctx.xs_mid[i]
            # ---- cross block (lightglue.py:193-222) ----
            qk, v, c, msg = ctx.kept[i][1] if ctx.kept[i] is not None else _cross_attend(k, w, x_mid)
            dxa, dmsg, g = _ffn_bwd(k, w, "c", x_mid, msg, dx)
            for name, val in g.items():
                grads[(i, name)] = val
            grads[(i, "co_w")], grads[(i, "co_b")], dc = _proj_bwd(k, w, "co_w", dmsg, c)
            dq, dk, dv = k.attention_bwd(qk, qk, v, c, dc, 1)
            dqv = k.heads_bwd(dq, dk, dv, None, None, None, 2, (C_SCALE, 1.0, 1.0), None)
            gw, gb, dxp = _proj_bwd(k, w, "cqv_w", dqv, x_mid)
            grads[(i, "cqk_w")], grads[(i, "cv_w")] = gw[:256], gw[256:]
            grads[(i, "cqk_b")], grads[(i, "cv_b")] = gb[:256], gb[256:]
            dxm = dx + dxa + dxp
            # ---- self block (lightglue.py:151-164) ----
            q, kk, v, c, msg = ctx.kept[i][0] if ctx.kept[i] is not None else _self_attend(k, w, x_in, rot)
            ctx.kept[i] = None
            dxa, dmsg, g = _ffn_bwd(k, w, "s", x_in, msg, dxm)
            for name, val in g.items():
                grads[(i, name)] = val
            grads[(i, "so_w")], grads[(i, "so_b")], dc = _proj_bwd(k, w, "so_w", dmsg, c)
            dq, dk, dv = k.attention_bwd(q, kk, v, c, dc, 0)
            dqkv = k.heads_bwd(dq, dk, dv, q, kk, rot, 3, (Q_SCALE, 1.0, 1.0), dtheta)
            perm = _PERM.to(dqkv.device)
            gw = torch.empty(768, 256, device=dqkv.device, dtype=torch.float32)
            gb = torch.empty(768, device=dqkv.device, dtype=torch.float32)
            gw[perm], gb[perm], dxp = _proj_bwd(k, w, "qkv_w", dqkv, x_in)
            grads[(i, "qkv_w")], grads[(i, "qkv_b")] = gw, gb
            dx = dxm + dxa + dxp
        # positional encoding: theta = Wr . normalised keypoints (lightglue.py:61-66)
        kdim = geom["k0"].shape[-1]
        kn = k.zeros(k.T, kdim)
        knv = kn.view(k.B, 2, k.Lp, kdim)
        for img, kk, sz, cnt in ((0, geom["k0"], geom["size0"], k.m), (1, geom["k1"], geom["size1"], k.n)):
            knv[:, img, :cnt, :2] = normalize_keypoints(kk[..., :2], sz)
            knv[:, img, :cnt, 2:] = kk[..., 2:]  # scale / orientation enter un-normalised (lightglue.py:436-454)
        grads["wr"] = dtheta.t() @ kn
        if ctx.xin is not None:  # input_proj (lightglue.py:464-465)
            grads["in_w"], grads["in_b"] = dx.t() @ ctx.xin, dx.sum(0)
            dx = dx @ W["in_w"]
        gd0, gd1 = k.unpack(dx)
        need = ctx.needs_input_grad
        return (None, None, gd0.contiguous() if need[2] else None, gd1.contiguous() if need[3] else None,
                *[grads[key] if need[4 + j] else None for j, key in enumerate(ctx.keys)])


class AssignFn(torch.autograd.Function):
    """MatchAssignment `layer` on (d0, d1) reduced to the inputs of NLLLoss: pos_sum [B] = sum of log_assignment over
    the ground-truth matches, dust0 [B,m] / dust1 [B,n] = the dustbin column / row (differentiable); pos_cnt, row_exp,
    row_arg, col_arg as in lgb200_loss_reduce (not differentiable)."""

    @staticmethod
    def forward(ctx, model, layer: int, gt, d0, d1, fp_w, fp_b, m_w, m_b):
        lib = _abi.load()
        keep: Dict = {}
        pos_sum, pos_cnt, row_exp, row_arg, col_arg, dust0, dust1, _ = model._log_assignment_of(
            lib, F32 if _SIMT_LINEAR else F32X3, d0, d1, layer, gt, keep=keep)
        ctx.save_for_backward(d0, d1, fp_w, fp_b, m_w, gt, keep["z"], keep["lse"])
        ctx.Lp = keep["Lp"]
        # x3 mode: the split planes of final_proj(desc) / 4 (33 MB at 8 x 2048) -- the backward recomputes the similarity
        # from them on tcgen05 instead of redoing final_proj and the N x M product in fp32 on the CUDA cores
        ctx.md_planes, ctx.lens = (None, None) if _SGEMM else (keep["md_planes"], keep["lens"])
        ctx.mark_non_differentiable(pos_cnt, row_exp, row_arg, col_arg)
        return pos_sum, dust0.contiguous(), dust1.contiguous(), pos_cnt, row_exp, row_arg, col_arg

    @staticmethod
    def backward(ctx, g_pos, g_d0, g_d1, *_):
        d0, d1, fp_w, fp_b, m_w, gt, z, lse = ctx.saved_tensors
        lib = _abi.load()
        B, m, _ = d0.shape
        n = d1.shape[1]
        Lp = ctx.Lp
        dev = d0.device
        x0, x1 = d0.to(torch.float32), d1.to(torch.float32)
        st = torch.cuda.current_stream(dev).cuda_stream
        mdp = ctx.md_planes
        if mdp is None:
            md0 = torch.addmm(fp_b, x0.reshape(-1, 256), fp_w.t()).view(B, m, 256) * 0.25  # lightglue.py:281-283
            md1 = torch.addmm(fp_b, x1.reshape(-1, 256), fp_w.t()).view(B, n, 256) * 0.25
            sim = torch.bmm(md0, md1.transpose(1, 2)).contiguous()
        else:
            sim = torch.empty(B, Lp, Lp, device=dev, dtype=torch.float32)
            check(lib.lgb200_x3_similarity(ptr(mdp), B, Lp, ptr(ctx.lens), ptr(sim), st), "x3_similarity")
            sim = sim[:, :m, :n].contiguous()
        g = (g_pos if g_pos is not None else torch.zeros(B, device=dev)).to(torch.float32).contiguous()
        r = (g[:, None] * gt.sum(2)).contiguous()
        c = (g[:, None] * gt.sum(1)).contiguous()
        check(lib.lgb200_assign_dsim(ptr(sim), B, m, n, ptr(lse), Lp, ptr(gt), ptr(g), ptr(r), ptr(c), st), "assign_dsim")
        if mdp is None:
            dmd0 = torch.bmm(sim, md1)
            dmd1 = torch.bmm(sim.transpose(1, 2), md0)
        else:  # three fp16 tensor-core products per GEMM, as _Kern.mm3
            sp = torch.empty(2, B, m, n, device=dev, dtype=torch.float16)
            inv = torch.empty(2, device=dev, dtype=torch.float32)
            check(lib.lgb200_split_dynamic(ptr(sim), sim.numel(), ptr(sp), ptr(inv), 0, None, 0, st), "lgb200_split_dynamic")
            mv = mdp.view(2, B, 2, Lp, 256)
            mp0, mp1 = mv[:, :, 0, :m], mv[:, :, 1, :n]
            f32 = torch.float32

            def bmm3(A, Bm):
                o = torch.bmm(A[1], Bm[0], out_dtype=f32)
                torch.baddbmm(o, A[0], Bm[1], out_dtype=f32, out=o)
                torch.baddbmm(o, A[0], Bm[0], out_dtype=f32, out=o)
                return o.mul_(inv[0] * AI)

            dmd0 = bmm3(sp, mp1)
            dmd1 = bmm3(sp.transpose(2, 3), mp0)
        zv = z.view(B, 2, Lp)
        z0, z1 = zv[:, 0, :m], zv[:, 1, :n]
        gd0 = g_d0 if g_d0 is not None else torch.zeros_like(z0)
        gd1 = g_d1 if g_d1 is not None else torch.zeros_like(z1)
        # la = ... + logsigmoid(z0_i) + logsigmoid(z1_j) inside, logsigmoid(-z) in the dustbins (lightglue.py:262-267)
        dz0 = r * torch.sigmoid(-z0) - gd0 * torch.sigmoid(z0)
        dz1 = c * torch.sigmoid(-z1) - gd1 * torch.sigmoid(z1)
        mw = m_w.reshape(1, 1, 256)
        gx0 = 0.25 * (dmd0 @ fp_w) + dz0[..., None] * mw
        gx1 = 0.25 * (dmd1 @ fp_w) + dz1[..., None] * mw
        g_fp_w = 0.25 * (dmd0.reshape(-1, 256).t() @ x0.reshape(-1, 256) + dmd1.reshape(-1, 256).t() @ x1.reshape(-1, 256))
        g_fp_b = 0.25 * (dmd0.sum((0, 1)) + dmd1.sum((0, 1)))
        g_m_w = ((dz0[..., None] * x0).sum((0, 1)) + (dz1[..., None] * x1).sum((0, 1))).reshape(m_w.shape)
        g_m_b = (dz0.sum() + dz1.sum()).reshape(1)
        return None, None, None, gx0, gx1, g_fp_w, g_fp_b, g_m_w, g_m_b


def wants_grad(model, data=None) -> bool:
    """Training step: module in training mode, autograd on, something to differentiate."""
    if not (model.training and torch.is_grad_enabled()):
        return False
    if any(p.requires_grad for p in model.parameters()):
        return True
    return data is not None and any(
        isinstance(data.get(kk), torch.Tensor) and data[kk].requires_grad for kk in ("descriptors0", "descriptors1"))
