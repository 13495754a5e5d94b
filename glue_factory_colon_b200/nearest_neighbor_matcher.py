"""B200 drop-in for `gluefactory.models.matchers.nearest_neighbor_matcher.NearestNeighborMatcher`
(SURVEY.md 8(f) rank 3): same conf keys, same input / output dict, fp32.

Compute runs in the C-ABI library: the similarity and the dual log-softmax reuse the fp32 similarity kernels of
MatchAssignment (`lgb200_assign_lse`, `lgb200_nn_scores`), `find_nn` + `mutual_check` are `lgb200_nn_match`.
There is no CPU path.  Select it in glue-factory with `model.matcher.name=glue_factory_colon_b200.nearest_neighbor_matcher`.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from . import _abi
from ._abi import F32, check, ptr


class NearestNeighborMatcher(nn.Module):
    # nearest_neighbor_matcher.py:47-53
    default_conf = {"ratio_thresh": None, "distance_thresh": None, "mutual_check": True, "loss": None}
    required_data_keys = ["descriptors0", "descriptors1"]

    def __init__(self, conf=None):
        super().__init__()
        merged = dict(self.default_conf)
        merged.update(dict(conf or {}))  # unknown keys (name, trainable, ...) are accepted like BaseModel does
        self.conf = SimpleNamespace(**merged)
        if self.conf.loss == "N_pair":  # :55-58
            self.register_parameter("temperature", nn.Parameter(torch.tensor(1.0)))

    @torch.no_grad()
    def forward(self, data: dict) -> dict:
        for key in self.required_data_keys:
            assert key in data, f"Missing key {key} in data"
        d0, d1 = data["descriptors0"], data["descriptors1"]
        if not d0.is_cuda:
            raise _abi.LightGlueB200Error("glue_factory_colon_b200.NearestNeighborMatcher runs on CUDA tensors only")
        lib = _abi.load()
        check(lib.lgb200_device_ok(), "device check")
        B, N, D = d0.shape
        M = d1.shape[1]
        assert d1.shape[0] == B and d1.shape[2] == D
        if D > 256:
            raise ValueError("descriptor dimension above 256 is not supported by the similarity kernel")
        dev = d0.device
        f32 = dict(device=dev, dtype=torch.float32)
        i64 = dict(device=dev, dtype=torch.int64)
        if N == 0 or M == 0:  # find_nn with no candidates (:17-18): nothing to compute
            return {
                "matches0": torch.full((B, N), -1, **i64), "matches1": torch.full((B, M), -1, **i64),
                "matching_scores0": torch.zeros(B, N, **f32), "matching_scores1": torch.zeros(B, M, **f32),
                "similarity": torch.zeros(B, N, M, **f32), "log_assignment": torch.zeros(B, N + 1, M + 1, **f32),
            }
        Lp = (max(N, M) + 127) // 128 * 128
        S = 2 * B
        # staging: [S, Lp, 256] fp32, sequence s = 2b + image, rows >= count and columns >= D zero
        md = torch.zeros(B, 2, Lp, 256, **f32)
        md[:, 0, :N, :D] = d0
        md[:, 1, :M, :D] = d1
        counts = torch.empty(B, 2, dtype=torch.int32)
        counts[:, 0], counts[:, 1] = N, M
        for idx in (0, 1):  # B200 extension, as in the LightGlue drop-in: per-pair valid counts of a padded batch
            cnt = data.get(f"num_keypoints{idx}")
            if cnt is not None:
                counts[:, idx] = torch.as_tensor(cnt).to("cpu", torch.int32).reshape(-1)
        lens = counts.reshape(-1).to(dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        lse = torch.empty(S * Lp, **f32)
        check(lib.lgb200_assign_lse(F32, ptr(md), S, Lp, ptr(lens), ptr(lse), st), "assign_lse")
        sim = torch.empty(B, N, M, **f32)
        la = torch.empty(B, N + 1, M + 1, **f32)
        check(lib.lgb200_nn_scores(ptr(md), ptr(lse), B, Lp, ptr(lens), N + 1, M + 1, ptr(sim), ptr(la), st), "nn_scores")
        ws = torch.empty(B * (N + M), **i64)
        m0, m1 = torch.empty(B, N, **i64), torch.empty(B, M, **i64)
        ms0, ms1 = torch.empty(B, N, **f32), torch.empty(B, M, **f32)
        check(
            lib.lgb200_nn_match(ptr(sim), B, N, M, ptr(lens), float(self.conf.ratio_thresh or 0.0),
                                float(self.conf.distance_thresh or 0.0), int(bool(self.conf.mutual_check)),
                                ptr(ws), ptr(m0), ptr(m1), ptr(ms0), ptr(ms1), st),
            "nn_match",
        )
        return {"matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1,
                "similarity": sim, "log_assignment": la}

    def loss(self, pred, data):
        """nearest_neighbor_matcher.py:85-109, FORWARD VALUES (no autograd graph: the temperature is not trained here).
        `N_pair`: three streaming passes over the similarity matrix in `lgb200_npair_loss`; metrics as the
        reference's matcher_metrics in eval mode.  Any other conf.loss raises NotImplementedError, which
        TwoViewPipeline.loss treats as "skip" (two_view_pipeline.py:417-429)."""
        if self.conf.loss != "N_pair":
            raise NotImplementedError
        with torch.no_grad():
            sim = pred["similarity"]
            if not sim.is_cuda:
                raise _abi.LightGlueB200Error("glue_factory_colon_b200.NearestNeighborMatcher.loss runs on CUDA tensors only")
            lib = _abi.load()
            sim = sim.to(torch.float32).contiguous()
            B, N, M = sim.shape
            dev = sim.device
            gt = data["gt_assignment"].to(dev).to(torch.bool).contiguous()
            assert gt.shape == (B, N, M), "gt_assignment must be [B, N, M]"
            f32 = dict(device=dev, dtype=torch.float32)
            rows = torch.empty(2, B, N, **f32)
            ws = torch.empty(B * (N + M), **f32)
            check(lib.lgb200_npair_loss(ptr(sim), ptr(gt), B, N, M, float(self.temperature), ptr(ws), ptr(rows[0]),
                                        ptr(rows[1]), torch.cuda.current_stream(dev).cuda_stream), "npair_loss")
            num = rows[1].sum(1).clamp(min=1.0)
            nll = -(rows[0].sum(1) / num) / 2
            losses = {"n_pair_nll": nll, "total": nll, "num_matchable": num,
                      "n_pair_temperature": self.temperature.detach()[None]}
            if self.training:
                return losses, {}
            from .lightglue import LightGlue

            return losses, LightGlue._matcher_metrics(pred, {"gt_matches0": data["gt_matches0"].to(dev)})


__main_model__ = NearestNeighborMatcher
