"""CPU restatement of the loss side of the LightGlue matcher (TEST INFRASTRUCTURE ONLY: imported by tests/ and
oracle/make_golden_loss.py, never by the product).

Follows, in plain torch on CPU:
  * weight_loss / NLLLoss        gluefactory/models/utils/losses.py:6-26, :44-73
  * TokenConfidence.loss         gluefactory/models/matchers/lightglue.py:82-95
  * LightGlue.loss               gluefactory/models/matchers/lightglue.py:588-637
  * matcher_metrics              gluefactory/models/utils/metrics.py:5-57
Pinned by tests/golden/loss_train.pt and loss_eval.pt, produced by the unmodified reference
(oracle/make_golden_loss.py).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import lightglue_oracle as lg


def nll_terms(la: torch.Tensor, data: dict, balancing: float = 0.5) -> Dict[str, torch.Tensor]:
    """losses.py:6-26 and :44-60 on one log-assignment matrix [B, M+1, N+1]."""
    pos = data["gt_assignment"].float()
    neg0 = (data["gt_matches0"] == -1).float()
    neg1 = (data["gt_matches1"] == -1).float()
    m, n = pos.shape[1], pos.shape[2]
    num_pos = pos.sum((1, 2)).clamp(min=1.0)
    num_neg0 = neg0.sum(1).clamp(min=1.0)
    num_neg1 = neg1.sum(1).clamp(min=1.0)
    nll_pos = -(la[:, :m, :n] * pos).sum((1, 2)) / num_pos
    nll_neg = (-(la[:, :m, n] * neg0).sum(1) - (la[:, m, :n] * neg1).sum(1)) / (num_neg0 + num_neg1)
    nll = balancing * nll_pos + (1 - balancing) * nll_neg
    return {
        "assignment_nll": nll,
        "nll_pos": nll_pos,
        "nll_neg": nll_neg,
        "num_matchable": num_pos,
        "num_unmatchable": (num_neg0 + num_neg1) / 2.0,
    }


def confidence_loss(sd, i: int, desc0, desc1, la_now, la_final) -> torch.Tensor:
    """lightglue.py:82-95: BCE of the token logits against "this layer's arg-max already equals the final one"."""
    w, b = sd[f"token_confidence.{i}.token.0.weight"], sd[f"token_confidence.{i}.token.0.bias"]
    logit0 = F.linear(desc0.detach(), w, b).squeeze(-1)  # lightglue.py:83-84: the heads do not train the descriptors
    logit1 = F.linear(desc1.detach(), w, b).squeeze(-1)
    la_now, la_final = la_now.detach(), la_final.detach()
    c0 = la_final[:, :-1, :].max(-1).indices == la_now[:, :-1, :].max(-1).indices
    c1 = la_final[:, :, :-1].max(-2).indices == la_now[:, :, :-1].max(-2).indices
    bce = F.binary_cross_entropy_with_logits
    return (bce(logit0, c0.float(), reduction="none").mean(-1) + bce(logit1, c1.float(), reduction="none").mean(-1)) / 2.0


def matcher_metrics(pred: dict, data: dict) -> Dict[str, torch.Tensor]:
    """metrics.py:5-57 for matches0 / matching_scores0."""
    m, gt, sc = pred["matches0"], data["gt_matches0"], pred["matching_scores0"]
    hit = (m == gt).float()
    r_mask = (gt > -1).float()
    a_mask = (gt >= -1).float()
    p_mask = ((m > -1) & (gt >= -1)).float()
    rec = (hit * r_mask).sum(1) / (1e-8 + r_mask.sum(1))
    acc = (hit * a_mask).sum(1) / (1e-8 + a_mask.sum(1))
    prec = (hit * p_mask).sum(1) / (1e-8 + p_mask.sum(1))
    order = torch.argsort(-sc)
    sp, sr, st = p_mask.gather(-1, order), r_mask.gather(-1, order), hit.gather(-1, order)
    p_pts = torch.cumsum(st * sp, -1) / (1e-8 + torch.cumsum(sp, -1))
    r_pts = torch.cumsum(st * sr, -1) / (1e-8 + sr.sum(-1)[:, None])
    # (the reference multiplies every recall step by the LAST precision point: p_pts[:, None, -1])
    ap = torch.sum((r_pts[..., 1:] - r_pts[..., :-1]) * p_pts[:, None, -1], dim=-1)
    return {"match_recall": rec, "match_precision": prec, "accuracy": acc, "average_precision": ap}


def loss(sd: Dict[str, torch.Tensor], conf: dict, pred: dict, data: dict, training: bool, keep_graph: bool = False):
    """lightglue.py:588-637.  pred: ref_descriptors0/1 [B, N, n, 256], log_assignment, matches0, matching_scores0.
    keep_graph: sd / pred are used as given (CPU tensors, any float dtype), so that torch autograd differentiates the
    restatement -- the checker of the hand-written backward pass (tests/test_gpu_grad.py)."""
    c = {**lg.DEFAULT_CONF, **{k: v for k, v in conf.items() if k in lg.DEFAULT_CONF}}
    lc = {"gamma": 1.0, "fn": "nll", "nll_balancing": 0.5, **conf.get("loss", {})}
    if keep_graph:
        r0, r1 = pred["ref_descriptors0"], pred["ref_descriptors1"]
    else:
        sd = {k: v.detach().cpu().float() for k, v in sd.items()}
        r0, r1 = pred["ref_descriptors0"].float().cpu(), pred["ref_descriptors1"].float().cpu()
    N = r0.shape[1]
    n_layers = c["n_layers"]

    def la_of(i):  # loss_params: log_assignment[i] applied to the i-th collected descriptors (i = -1: last module)
        mod = n_layers - 1 if i == -1 else i
        return torch.stack([lg.log_assignment(sd, mod, r0[b, i], r1[b, i]) for b in range(r0.shape[0])])

    la_last = la_of(-1)
    terms = nll_terms(la_last, data, lc["nll_balancing"])
    losses = {"total": terms["assignment_nll"].clone(), "last": terms["assignment_nll"].clone().detach(), **terms}
    if training:
        losses["confidence"] = torch.zeros_like(losses["total"])
    la_pred = pred["log_assignment"].detach().to(r0.dtype).cpu()
    losses["row_norm"] = la_pred.exp()[:, :-1].sum(2).mean(1)
    sum_w = 1.0
    for i in range(N - 1):
        la_i = la_of(i)
        nll_i = nll_terms(la_i, data, lc["nll_balancing"])["assignment_nll"]
        w = lc["gamma"] ** (N - i - 1) if lc["gamma"] > 0.0 else i + 1
        sum_w += w
        losses["total"] = losses["total"] + nll_i * w
        losses["confidence"] = losses["confidence"] + confidence_loss(sd, i, r0[:, i], r1[:, i], la_i, la_pred) / (N - 1)
    losses["total"] = losses["total"] / sum_w
    if training:
        losses["total"] = losses["total"] + losses["confidence"]
    metrics = {} if training else matcher_metrics({k: v.cpu() for k, v in pred.items() if k in ("matches0", "matching_scores0")}, data)
    return losses, metrics


def forward_collect(sd, conf: dict, data: dict, keep_graph: bool = False, dtype=torch.float32) -> dict:
    """Training-mode forward (lightglue.py:483-498, :541-553): all layers' descriptors are collected, no early exit /
    pruning.  Returns the batched prediction dict (full-size pairs only).  keep_graph: sd is used as given."""
    outs: List[dict] = []
    traces = []
    sd_c = sd if keep_graph else {k: v.detach().cpu() for k, v in sd.items()}
    B = data["keypoints0"].shape[0]
    conf_t = {**conf, "depth_confidence": -1, "width_confidence": -1}
    size0 = data.get("view0", {}).get("image_size")
    size1 = data.get("view1", {}).get("image_size")
    k0, k1 = lg._kpts_with_scale_ori(data, 0, conf), lg._kpts_with_scale_ori(data, 1, conf)  # lightglue.py:436-454
    for b in range(B):
        tr: dict = {}
        outs.append(
            lg.forward_pair(
                sd_c, conf_t, k0[b], k1[b], data["descriptors0"][b],
                data["descriptors1"][b], None if size0 is None else size0[b].float(),
                None if size1 is None else size1[b].float(), dtype=dtype, trace=tr,
            )
        )
        traces.append(tr)
    pred = {k: torch.stack([o[k] for o in outs]) for k in ("matches0", "matches1", "matching_scores0", "matching_scores1", "log_assignment")}
    pred["ref_descriptors0"] = torch.stack([torch.stack(t["desc0"]) for t in traces])
    pred["ref_descriptors1"] = torch.stack([torch.stack(t["desc1"]) for t in traces])
    return pred
