"""GPU tests of the training path: the hand-written backward kernels (csrc/lg_bwd.cu) against
torch autograd of the same op in float64, and a whole training step -- forward in training mode, LightGlue.loss,
`losses["total"].mean().backward()` -- against (a) torch autograd through the CPU oracle, every entry of every
gradient, and (b) gradient goldens of the unmodified reference (oracle/make_golden_grad.py)."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

from glue_factory_colon_b200 import LightGlue, _abi
from glue_factory_colon_b200._abi import ptr
from glue_factory_colon_b200.synthetic import make_pairs, to_device
from helpers import oracle_training_step

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
LOG2E = 1.4426950408889634


def _st():
    return torch.cuda.current_stream(DEV).cuda_stream


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("kv_xor", [0, 1], ids=["self", "cross"])
@pytest.mark.parametrize("lens,gscale", [(None, 1.0), ([200, 131, 64, 256], 1.0), ([70, 0, 129, 5], 1.0),
                                         ([200, 131, 64, 256], 3e-9), (None, 7e3)],
                         ids=["full", "ragged", "empty_side", "ragged_tiny_gradient", "full_large_gradient"])
def test_attention_bwd_against_autograd(kv_xor, lens, gscale):
    """gscale: gradients have no fixed range -- the tcgen05 kernel splits g dO into fp16 planes with a per-call power of
    two g (lg_x3_attn_bwd.cu); 3e-9 would vanish in fp16 without it, 7e3 would overflow the D planes."""
    lib = _abi.load()
    S, Lp = 4, 256
    g = torch.Generator().manual_seed(11 + kv_xor)
    q = (torch.randn(S, 4, Lp, 64, generator=g) * 0.6).to(DEV)
    k = (torch.randn(S, 4, Lp, 64, generator=g) * 0.6).to(DEV)
    v = torch.randn(S, 4, Lp, 64, generator=g).to(DEV)
    dctx = (torch.randn(S, Lp, 256, generator=g) * torch.rand(S, Lp, 1, generator=g) ** 4 * gscale).to(DEV)
    ln = torch.tensor(lens if lens is not None else [Lp] * S)
    lens_d = None if lens is None else ln.to(DEV, torch.int32)
    # reference: float64 autograd; q is in the log2 domain (the kernels exponentiate with exp2)
    qd, kd, vd = (t.double().cpu().requires_grad_(True) for t in (q, k, v))
    ctx_ref = torch.zeros(S, Lp, 256, dtype=torch.float64)
    for s in range(S):
        so = s ^ kv_xor
        nq, nk = int(ln[s]), int(ln[so])
        if nq == 0 or nk == 0:
            continue
        a = torch.softmax(qd[s, :, :nq] @ kd[so, :, :nk].transpose(-1, -2) * math.log(2.0), -1)
        ctx_ref[s, :nq] = (a @ vd[so, :, :nk]).permute(1, 0, 2).reshape(nq, 256)
    (ctx_ref * dctx.double().cpu()).sum().backward()
    ctx = torch.zeros(S, Lp, 256, device=DEV)
    assert lib.lgb200_attention(_abi.F32, ptr(q), ptr(k), ptr(v), S, Lp, ptr(lens_d), kv_xor, ptr(ctx), _st()) == 0
    dq, dk, dv = (torch.full((S, 4, Lp, 64), 7.0, device=DEV) for _ in range(3))
    n_ws = ctypes.c_longlong(0)
    assert lib.lgb200_attention_bwd_workspace(S, Lp, ctypes.byref(n_ws)) == 0
    ws = torch.empty(n_ws.value, device=DEV)
    rc = lib.lgb200_attention_bwd(ptr(q), ptr(k), ptr(v), ptr(ctx), ptr(dctx), S, Lp, ptr(lens_d), kv_xor, ptr(dq), ptr(dk),
                                  ptr(dv), ptr(ws), _st())
    assert rc == 0, lib.lgb200_error_string(rc)
    for name, got, ref in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        assert torch.isfinite(got).all(), name
        assert _rel(got, ref) < 2e-5, f"{name}: {_rel(got, ref)}"
        for s in range(S):  # rows past the valid count carry no gradient
            assert float(got[s, :, int(ln[s]):].abs().max() if int(ln[s]) < Lp else 0.0) == 0.0
    assert lib.lgb200_attention_bwd(ptr(q), ptr(k), ptr(v), ptr(ctx), ptr(dctx), S, Lp, None, kv_xor, None, ptr(dk), ptr(dv),
                                    ptr(ws), _st()) == -2


def test_attention_bwd_long_sequence_accumulation():
    """2048 keys = 384 MMAs per output element: a single tensor-core accumulator (rounded toward zero after every MMA) is
    ~1e-5 low on same-sign sums; the kernel folds its TMEM accumulators into round-to-nearest running sums every 4 tiles
    (lg_x3_attn_bwd.cu).  Reference: float64 autograd on the GPU; inputs with non-zero means so that the sums have a sign."""
    lib = _abi.load()
    S, Lp = 2, 2048
    g = torch.Generator().manual_seed(1)
    q = (torch.randn(S, 4, Lp, 64, generator=g) * 0.6).to(DEV)
    k = (torch.randn(S, 4, Lp, 64, generator=g) * 0.6).to(DEV)
    v = (torch.randn(S, 4, Lp, 64, generator=g) + 0.5).to(DEV)
    dctx = (torch.randn(S, Lp, 256, generator=g) * 1e-3 + 5e-4).to(DEV)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    a = torch.softmax(qd @ kd.transpose(-1, -2) * math.log(2.0), -1)
    ((a @ vd).permute(0, 2, 1, 3).reshape(S, Lp, 256) * dctx.double()).sum().backward()
    ctx = torch.zeros(S, Lp, 256, device=DEV)
    assert lib.lgb200_attention(_abi.F32, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), _st()) == 0
    dq, dk, dv = (torch.empty(S, 4, Lp, 64, device=DEV) for _ in range(3))
    n_ws = ctypes.c_longlong(0)
    assert lib.lgb200_attention_bwd_workspace(S, Lp, ctypes.byref(n_ws)) == 0
    ws = torch.empty(n_ws.value, device=DEV)
    assert lib.lgb200_attention_bwd(ptr(q), ptr(k), ptr(v), ptr(ctx), ptr(dctx), S, Lp, None, 0, ptr(dq), ptr(dk), ptr(dv),
                                    ptr(ws), _st()) == 0
    for name, got, ref in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        d = got.double() - ref
        assert float(d.norm() / ref.norm()) < 4e-6, f"{name}: {float(d.norm() / ref.norm()):.2e}"
        assert abs(float(d.mean() / ref.abs().mean())) < 2e-6, f"{name}: one-sided error {float(d.mean() / ref.abs().mean()):+.2e}"


def test_attention_bwd_shared_qk_tensor_and_split_gemm():
    """Cross block: to_qk feeds both sides, K IS Q (one plane pair inside the kernel wrapper).  Also the backward's
    three-product tensor-core GEMM (_Kern.mm3 on lgb200_split_dynamic planes) against float64."""
    lib = _abi.load()
    S, Lp = 2, 384
    g = torch.Generator().manual_seed(5)
    qk = (torch.randn(S, 4, Lp, 64, generator=g) * 0.5).to(DEV)
    v = torch.randn(S, 4, Lp, 64, generator=g).to(DEV)
    dctx = (torch.randn(S, Lp, 256, generator=g) * 1e-4).to(DEV)
    qd, vd = qk.double().cpu().requires_grad_(True), v.double().cpu().requires_grad_(True)
    ctx_ref = torch.stack([(torch.softmax(qd[s] @ qd[s ^ 1].transpose(-1, -2) * math.log(2.0), -1) @ vd[s ^ 1])
                           .permute(1, 0, 2).reshape(Lp, 256) for s in range(S)])
    (ctx_ref * dctx.double().cpu()).sum().backward()
    ctx = torch.zeros(S, Lp, 256, device=DEV)
    assert lib.lgb200_attention(_abi.F32, ptr(qk), ptr(qk), ptr(v), S, Lp, None, 1, ptr(ctx), _st()) == 0
    dq, dk, dv = (torch.empty(S, 4, Lp, 64, device=DEV) for _ in range(3))
    n_ws = ctypes.c_longlong(0)
    assert lib.lgb200_attention_bwd_workspace(S, Lp, ctypes.byref(n_ws)) == 0
    ws = torch.empty(n_ws.value, device=DEV)
    assert lib.lgb200_attention_bwd(ptr(qk), ptr(qk), ptr(v), ptr(ctx), ptr(dctx), S, Lp, None, 1, ptr(dq), ptr(dk),
                                    ptr(dv), ptr(ws), _st()) == 0
    assert _rel(dq + dk, qd.grad) < 2e-5 and _rel(dv, vd.grad) < 2e-5

    from glue_factory_colon_b200.train import _Kern

    kern = _Kern(DEV, 1, 128, 128)
    a = (torch.randn(1024, 512, generator=g) * torch.rand(1024, 1, generator=g) ** 6 * 2e-7).to(DEV)   # a "gradient"
    b = torch.randn(1024, 256, generator=g).to(DEV)                                                     # an activation
    ap, ai = kern.gsplit(a)
    got = kern.mm3(ap.transpose(1, 2), kern.asplit(b), ai / 64.0)
    assert _rel(got, a.double().t() @ b.double()) < 2e-6
    z = torch.zeros(8, 64, device=DEV)
    zp, zi = kern.gsplit(z)
    assert float(zi) == 1.0 and float(zp.abs().max()) == 0.0


def test_ln_gelu_bwd_against_autograd():
    lib = _abi.load()
    S, Lp = 2, 256
    T = S * Lp
    g = torch.Generator().manual_seed(5)
    h = (torch.randn(T, 512, generator=g) * 1.5 + 0.3).to(DEV)
    gamma = (1 + 0.2 * torch.randn(512, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(512, generator=g)).to(DEV)
    da = torch.randn(T, 512, generator=g).to(DEV)
    lens = torch.tensor([200, 256], dtype=torch.int32, device=DEV)
    valid = torch.zeros(S, Lp, dtype=torch.bool)
    valid[0, :200] = True
    valid[1] = True
    valid = valid.reshape(T)
    hd, gd, bd = h.double().cpu().requires_grad_(True), gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    y = F.gelu(F.layer_norm(hd, (512,), gd, bd, 1e-5))
    (y * da.double().cpu() * valid[:, None]).sum().backward()
    dh, act = torch.full((T, 512), 7.0, device=DEV), torch.full((T, 512), 7.0, device=DEV)
    nP = 296
    part = torch.empty(nP, 1024, device=DEV)
    rc = lib.lgb200_ln_gelu_bwd(ptr(h), ptr(gamma), ptr(beta), ptr(da), T, Lp, ptr(lens), ptr(dh), ptr(act), ptr(part), nP, _st())
    assert rc == 0, lib.lgb200_error_string(rc)
    assert _rel(dh, hd.grad) < 1e-5
    assert float(dh[~valid.to(DEV)].abs().max()) == 0.0 and float(act[~valid.to(DEV)].abs().max()) == 0.0
    assert _rel(act[valid.to(DEV)], y[valid]) < 1e-6
    gs = part.sum(0)
    assert _rel(gs[:512], gd.grad) < 1e-5 and _rel(gs[512:], bd.grad) < 1e-5


def test_heads_bwd_against_autograd():
    lib = _abi.load()
    S, Lp = 2, 128
    T = S * Lp
    g = torch.Generator().manual_seed(9)
    theta = (torch.randn(T, 32, generator=g) * 2).double().requires_grad_(True)
    raw = torch.randn(T, 768, generator=g).double().requires_grad_(True)  # packed columns part*256 + head*64 + d
    s0 = 0.37

    def heads(t):  # [T,256] -> [S,4,Lp,64]
        return t.view(S, Lp, 4, 64).permute(0, 2, 1, 3)

    def rotary(t):  # pairs (2f, 2f+1) rotated by theta_f (lightglue.py:43-50)
        c, s = torch.cos(theta).repeat_interleave(2, -1).repeat(1, 4), torch.sin(theta).repeat_interleave(2, -1).repeat(1, 4)
        tv = t.view(T, 128, 2)
        rot = torch.stack((-tv[..., 1], tv[..., 0]), -1).reshape(T, 256)
        return t * c + rot * s

    q, k, v = heads(rotary(raw[:, :256]) * s0), heads(rotary(raw[:, 256:512])), heads(raw[:, 512:])
    gq, gk, gv = (torch.randn(S, 4, Lp, 64, generator=g).double() for _ in range(3))
    lens = torch.tensor([100, 128], dtype=torch.int32)
    mask = torch.zeros(S, 1, Lp, 1, dtype=torch.float64)
    mask[0, :, :100] = 1
    mask[1] = 1
    ((q * gq + k * gk + v * gv) * mask).sum().backward()
    rot = torch.stack((torch.cos(theta), torch.sin(theta)), -1).reshape(T, 64).float().detach().to(DEV).contiguous()
    out = torch.full((T, 768), 7.0, device=DEV)
    dth = torch.ones(T, 32, device=DEV)  # accumulates: starts from 1
    gq_d, gk_d, gv_d, q_d, k_d = (t.detach().float().contiguous().to(DEV) for t in (gq, gk, gv, q, k))
    lens_d = lens.to(DEV)
    rc = lib.lgb200_heads_bwd(ptr(gq_d), ptr(gk_d), ptr(gv_d), ptr(q_d), ptr(k_d), ptr(rot), S, Lp, ptr(lens_d), 3,
                              s0, 1.0, 1.0, ptr(out), ptr(dth), _st())
    assert rc == 0, lib.lgb200_error_string(rc)
    assert _rel(out, raw.grad) < 1e-5
    vt = mask.expand(S, 1, Lp, 1).reshape(T, 1).bool().expand(T, 32)
    assert _rel((dth.cpu() - 1)[vt], theta.grad[vt]) < 1e-5
    assert float((dth.cpu() - 1)[~vt].abs().max()) == 0.0
    # cross block: to_qk feeds the query and the key side
    out2 = torch.full((T, 512), 7.0, device=DEV)
    rc = lib.lgb200_heads_bwd(ptr(gq_d), ptr(gk_d), ptr(gv_d), None, None, None, S, Lp, ptr(lens_d), 2, s0, 1.0, 1.0,
                              ptr(out2), None, _st())
    assert rc == 0, lib.lgb200_error_string(rc)
    unheads = lambda t: (t * mask).permute(0, 2, 1, 3).reshape(T, 256)  # noqa: E731
    assert _rel(out2[:, :256], s0 * unheads(gq + gk)) < 1e-6 and _rel(out2[:, 256:], unheads(gv)) < 1e-6


def test_assign_dsim_against_autograd():
    lib = _abi.load()
    B, m, n, Lp = 2, 150, 131, 256
    g = torch.Generator().manual_seed(3)
    sim = (torch.randn(B, m, n, generator=g) * 2).double().requires_grad_(True)
    gt = torch.rand(B, m, n, generator=g) < 0.01
    gpos = torch.randn(B, generator=g).double()
    la = F.log_softmax(sim, 2) + F.log_softmax(sim, 1)  # the z terms do not depend on sim (lightglue.py:261-265)
    ((la * gt).sum((1, 2)) * gpos).sum().backward()
    lse = torch.zeros(2 * B, Lp, dtype=torch.float64)
    lse[0::2, :m] = torch.logsumexp(sim.detach(), 2)
    lse[1::2, :n] = torch.logsumexp(sim.detach(), 1)
    sim_d = sim.detach().float().to(DEV).contiguous()
    r = (gpos[:, None] * gt.sum(2)).float().to(DEV).contiguous()
    c = (gpos[:, None] * gt.sum(1)).float().to(DEV).contiguous()
    lse_d, gt_d, gpos_d = lse.float().to(DEV), gt.to(DEV), gpos.float().to(DEV)
    rc = lib.lgb200_assign_dsim(ptr(sim_d), B, m, n, ptr(lse_d), Lp, ptr(gt_d), ptr(gpos_d), ptr(r), ptr(c), _st())
    assert rc == 0, lib.lgb200_error_string(rc)
    assert _rel(sim_d, sim.grad) < 1e-5


def _training_step(model, data):
    model.zero_grad(set_to_none=True)
    pred = model(data)
    losses, metrics = model.loss(pred, data)
    losses["total"].mean().backward()
    return pred, losses, metrics


@pytest.mark.parametrize("checkpointed", [False, True], ids=["kept", "recompute"])
@pytest.mark.parametrize("name", ["grad_train", "grad_train_sift"])
def test_training_step_gradients_against_oracle_autograd_and_reference_golden(name, checkpointed, golden_dir):
    fx = torch.load(golden_dir / f"{name}.pt", weights_only=False)
    torch.manual_seed(fx["seed"])
    model = LightGlue({**fx["conf"], "checkpointed": checkpointed})
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    fp = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(fp - fx["fingerprint"]) < 1e-6 * fx["fingerprint"]
    data_c = make_pairs(with_gt=True, **fx["data_kwargs"])
    total_o, grads_o, gd0_o, gd1_o = oracle_training_step(sd, fx["conf"], data_c, dtype=torch.float64)

    model = model.to(DEV).train()
    data = to_device(data_c, DEV)
    data["descriptors0"].requires_grad_(True)
    data["descriptors1"].requires_grad_(True)
    pred, losses, metrics = _training_step(model, data)
    assert metrics == {} and pred["ref_descriptors0"].requires_grad and losses["total"].requires_grad
    assert not losses["last"].requires_grad  # lightglue.py:601
    torch.testing.assert_close(losses["total"].detach().cpu(), fx["total"], atol=5e-4, rtol=1e-4)
    torch.testing.assert_close(losses["total"].detach().cpu().double(), total_o, atol=5e-4, rtol=1e-4)
    # (a) every entry of every gradient against autograd through the float64 oracle
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        assert p.grad is not None, f"{k} received no gradient"
        assert torch.isfinite(p.grad).all(), k
        e = _rel(p.grad, grads_o[k])
        worst = max(worst, (k, e), key=lambda t: t[1])
    assert worst[1] < 2e-3, f"largest relative gradient error {worst[1]:.2e} at {worst[0]}"
    assert _rel(data["descriptors0"].grad, gd0_o) < 2e-3 and _rel(data["descriptors1"].grad, gd1_o) < 2e-3
    # (b) the unmodified reference's gradients (norms and sampled entries)
    named = dict(model.named_parameters())
    assert set(fx["grads"]) == set(named)
    for k, s in list(fx["grads"].items()) + [("descriptors0", fx["descriptors0"]), ("descriptors1", fx["descriptors1"])]:
        gsrc = data[k].grad if k.startswith("descriptors") else named[k].grad
        gflat = gsrc.detach().double().cpu().reshape(-1)
        assert abs(float(gflat.norm()) - s["norm"]) <= 3e-3 * s["norm"] + 1e-7, k
        torch.testing.assert_close(gflat[s["idx"]].float(), s["val"], atol=3e-3 * s["norm"] / gflat.numel() ** 0.5 + 1e-7,
                                   rtol=1e-2, msg=lambda m: f"{k}: {m}")


def test_training_reduces_the_loss():
    """A few Adam steps on one batch through the drop-in (forward kernels + hand-written backward): the loss drops."""
    torch.manual_seed(0)
    model = LightGlue({"n_layers": 3, "filter_threshold": 0.0}).to(DEV).train()
    data = to_device(make_pairs(B=2, n0=128, n1=100, seed=4, with_gt=True), DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    hist = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        losses, _ = model.loss(model(data), data)
        loss = losses["total"].mean()
        loss.backward()
        opt.step()
        hist.append(float(loss.detach()))
    assert all(math.isfinite(x) for x in hist) and hist[-1] < hist[0] - 0.05, hist


def test_no_grad_training_mode_keeps_the_forward_values_path():
    torch.manual_seed(1)
    model = LightGlue({"n_layers": 2, "precision": "bf16"}).to(DEV).train()
    data = to_device(make_pairs(B=1, n0=128, n1=128, seed=2, with_gt=True), DEV)
    with torch.no_grad():
        pred = model(data)
        losses, _ = model.loss(pred, data)
    assert pred["ref_descriptors0"].dtype == torch.bfloat16 and not losses["total"].requires_grad
