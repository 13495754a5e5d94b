"""GPU tests of the loss side: training-mode forward (all layers collected) and
LightGlue.loss forward values against goldens produced by the unmodified reference, plus the reduction kernel
against torch on matrices with exact ties."""
import pytest
import torch

from glue_factory_colon_b200 import LightGlue, _abi
from glue_factory_colon_b200._abi import ptr
from glue_factory_colon_b200.synthetic import make_pairs, to_device

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _stream():
    return torch.cuda.current_stream(DEV).cuda_stream


def _model(fx, precision):
    torch.manual_seed(fx["seed"])
    model = LightGlue({**fx["conf"], "precision": precision})
    fp = float(sum(v.double().abs().sum() for v in model.state_dict().values()))
    assert abs(fp - fx["fingerprint"]) < 1e-6 * fx["fingerprint"]
    return model.to(DEV).train(fx["training"])


def test_loss_reduce_against_torch():
    lib = _abi.load()
    g = torch.Generator().manual_seed(3)
    B, R, C = 3, 201, 150
    la = ((torch.randn(B, R, C, generator=g) * 3).round() / 2 - 4.0).to(DEV)  # quantised: plenty of exact ties
    la[1, 7, 5] = float("nan")
    la[2, :, 11] = -float("inf")
    gt = (torch.rand(B, R - 1, C - 1, generator=g) < 0.02).to(DEV)
    rows = torch.zeros(3, B, R - 1, device=DEV)
    ra = torch.zeros(B, R - 1, device=DEV, dtype=torch.int32)
    ca = torch.zeros(B, C - 1, device=DEV, dtype=torch.int32)
    rc = lib.lgb200_loss_reduce(ptr(la), B, R, C, ptr(gt), ptr(rows[0]), ptr(rows[1]), ptr(rows[2]), ptr(ra), ptr(ca), _stream())
    assert rc == 0, lib.lgb200_error_string(rc)
    inner = la[:, :-1, :-1]
    ok = ~torch.isnan(inner).any(-1)
    torch.testing.assert_close(rows[0][ok], torch.where(gt, inner, torch.zeros_like(inner)).sum(-1)[ok], atol=1e-4, rtol=1e-5)
    assert torch.equal(rows[1], gt.float().sum(-1))
    torch.testing.assert_close(rows[2][ok], la[:, :-1].exp().sum(-1)[ok], atol=1e-5, rtol=1e-5)
    assert torch.equal(ra.long(), la[:, :-1, :].max(-1).indices)
    assert torch.equal(ca.long(), la[:, :, :-1].max(-2).indices)
    # outputs are optional
    assert lib.lgb200_loss_reduce(ptr(la), B, R, C, None, None, None, None, ptr(ra), None, _stream()) == 0
    assert lib.lgb200_loss_reduce(None, B, R, C, None, None, None, None, None, None, _stream()) == -2


@pytest.mark.parametrize("grad", [False, True], ids=["values", "autograd"])
@pytest.mark.parametrize("name", ["loss_train", "loss_train_gamma", "loss_eval"])
def test_forward_and_loss_fp32_against_reference_golden(name, grad, golden_dir):
    """grad=False: the forward-values path (torch.no_grad(), e.g. validation inside the reference's training loop);
    grad=True: the training path of glue_factory_colon_b200/train.py (same values, with an autograd graph)."""
    fx = torch.load(golden_dir / f"{name}.pt", weights_only=False)
    model = _model(fx, "fp32")
    data = to_device(make_pairs(with_gt=True, **fx["data_kwargs"]), DEV)
    with torch.set_grad_enabled(grad):
        pred = model(data)
        losses, metrics = model.loss(pred, data)
    assert pred["ref_descriptors0"].requires_grad == (grad and fx["training"])
    pred = {k: v.detach() for k, v in pred.items()}
    losses = {k: v.detach() for k, v in losses.items()}
    assert tuple(pred["ref_descriptors0"].shape) == fx["ref_desc_shape"]
    for i, am in enumerate(fx["ref_desc_absmean"]):
        assert abs(float(pred["ref_descriptors0"][:, i].abs().mean()) - am) < 1e-4 * am
    torch.testing.assert_close(pred["log_assignment"].cpu(), fx["pred"]["log_assignment"], atol=1e-3, rtol=0)
    assert set(losses) == set(fx["losses"])
    for k, v in fx["losses"].items():
        torch.testing.assert_close(losses[k].reshape(-1).cpu(), v.reshape(-1).float(), atol=5e-4, rtol=1e-4,
                                   msg=lambda m: f"{k}: {m}")
    assert set(metrics) == set(fx["metrics"])
    if torch.equal(pred["matches0"].cpu(), fx["pred"]["matches0"]):
        for k, v in fx["metrics"].items():
            torch.testing.assert_close(metrics[k].cpu(), v, atol=1e-5, rtol=1e-5)


def test_matcher_metrics_known_answers(golden_dir):
    fx = torch.load(golden_dir / "metrics_kat.pt", weights_only=False)
    data = make_pairs(with_gt=True, **fx["data_kwargs"])
    met = LightGlue._matcher_metrics({"matches0": fx["matches0"].to(DEV), "matching_scores0": fx["matching_scores0"].to(DEV)},
                                     {"gt_matches0": data["gt_matches0"].to(DEV)})
    for k, v in fx["metrics"].items():
        torch.testing.assert_close(met[k].cpu(), v, atol=1e-6, rtol=1e-6)


def test_forward_and_loss_bf16_within_the_bf16_envelope(golden_dir):
    """bf16 (tcgen05) mode: the same loss through the throughput kernels; tolerance = the bf16 envelope of DESIGN.md
    section 2 (mean |d log_assignment| < 0.05) carried to the loss: 0.05 absolute on the NLL terms."""
    fx = torch.load(golden_dir / "loss_train.pt", weights_only=False)
    model = _model(fx, "bf16")
    data = to_device(make_pairs(with_gt=True, **fx["data_kwargs"]), DEV)
    with torch.no_grad():  # (with autograd on, a training-mode forward takes the fp32 training path)
        pred = model(data)
        losses, metrics = model.loss(pred, data)
    assert tuple(pred["ref_descriptors0"].shape) == fx["ref_desc_shape"] and pred["ref_descriptors0"].dtype == torch.bfloat16
    assert metrics == {} and set(losses) == set(fx["losses"])
    for k in ("total", "last", "nll_pos", "nll_neg", "confidence", "row_norm"):
        torch.testing.assert_close(losses[k].cpu(), fx["losses"][k].float(), atol=0.05, rtol=0.02, msg=lambda m: f"{k}: {m}")
    assert torch.equal(losses["num_matchable"].cpu(), fx["losses"]["num_matchable"])


@pytest.mark.parametrize("m,n", [(256, 256), (200, 179), (384, 130)], ids=["full", "ragged", "wide"])
def test_fused_assign_loss_equals_scores_plus_reduce(m, n):
    """lgb200_assign_loss (tcgen05 pass 2 with the loss reductions in its epilogue, nothing N x M written) against
    lgb200_assign_scores + lgb200_loss_reduce on the same operands: arg-maxima identical, sums to fp32 rounding."""
    lib = _abi.load()
    g = torch.Generator().manual_seed(m + n)
    B = 3
    Lp = ((max(m, n) + 127) // 128) * 128
    S = 2 * B
    md = torch.zeros(S, Lp, 256, dtype=torch.bfloat16, device=DEV)
    md[0::2, :m] = (torch.randn(B, m, 256, generator=g) * 0.35).to(DEV).to(torch.bfloat16)
    md[1::2, :n] = (torch.randn(B, n, 256, generator=g) * 0.35).to(DEV).to(torch.bfloat16)
    z = (torch.randn(S, Lp, generator=g) * 2).to(DEV)
    z[0, 3] = 30.0   # dustbin-dominated and matchable-dominated rows / columns
    z[1, 5] = -30.0
    lens = None if m == Lp and n == Lp else torch.tensor([m, n] * B, dtype=torch.int32, device=DEV)
    lse = torch.zeros(S, Lp, device=DEV)
    assert lib.lgb200_assign_lse(_abi.BF16, ptr(md), S, Lp, ptr(lens), ptr(lse), _stream()) == 0
    R, C = m + 1, n + 1
    gt = (torch.rand(B, m, n, generator=g) < 0.01).to(DEV)
    sc = torch.empty(B, R, C, device=DEV)
    assert lib.lgb200_assign_scores(_abi.BF16, ptr(md), ptr(z), ptr(lse), B, Lp, ptr(lens), R, C, ptr(sc), None, _stream()) == 0
    ref_rows = torch.zeros(3, B, m, device=DEV)
    ref_ra = torch.zeros(B, m, device=DEV, dtype=torch.int32)
    ref_ca = torch.zeros(B, n, device=DEV, dtype=torch.int32)
    assert lib.lgb200_loss_reduce(ptr(sc), B, R, C, ptr(gt), ptr(ref_rows[0]), ptr(ref_rows[1]), ptr(ref_rows[2]), ptr(ref_ra),
                                  ptr(ref_ca), _stream()) == 0
    rows = torch.full((3, B, m), 7.0, device=DEV)
    ra = torch.zeros(B, m, device=DEV, dtype=torch.int32)
    ca = torch.zeros(B, n, device=DEV, dtype=torch.int32)
    ws = torch.empty(B * (R + C), device=DEV, dtype=torch.int64)
    rc = lib.lgb200_assign_loss(_abi.BF16, ptr(md), ptr(z), ptr(lse), B, Lp, ptr(lens), R, C, ptr(gt), ptr(rows[0]), ptr(rows[1]),
                                ptr(rows[2]), ptr(ra), ptr(ca), ptr(ws), _stream())
    assert rc == 0, lib.lgb200_error_string(rc)
    assert torch.equal(rows[1], ref_rows[1])
    torch.testing.assert_close(rows[0], ref_rows[0], atol=1e-4, rtol=1e-5)
    torch.testing.assert_close(rows[2], ref_rows[2], atol=1e-5, rtol=1e-4)
    assert torch.equal(ra, ref_ra) and torch.equal(ca, ref_ca)
    assert (ra == n).any() and (ra < n).any()
    # fp32 is refused: the parity mode materialises the matrix
    assert lib.lgb200_assign_loss(_abi.F32, ptr(md), ptr(z), ptr(lse), B, Lp, ptr(lens), R, C, ptr(gt), ptr(rows[0]), ptr(rows[1]),
                                  ptr(rows[2]), ptr(ra), ptr(ca), ptr(ws), _stream()) == -3
