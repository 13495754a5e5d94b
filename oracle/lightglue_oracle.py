"""CPU oracle for the LightGlue matcher forward pass.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (functional, state-dict driven, one
pair at a time) of the algorithm in the reference
`gluefactory/models/matchers/lightglue.py` (cited below as `lightglue.py:LINE`,
relative to the reference checkout).  It exists to *check* the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
    `--impl reference` legs may import it;
  * nothing under `glue_factory_colon_b200/` may import it -- the product path
    has no CPU fallback and fails loudly without its CUDA library.

Parity pinning.  The reference holds NO golden vectors or tests for this path
(SURVEY.md F9).  The oracle is therefore pinned against outputs of the
reference itself: `oracle/make_golden.py` imports the unmodified reference
from /root/reference (with the test-only `oracle/_shim/omegaconf`), runs it on
seeded inputs and commits inputs+outputs under `tests/golden/`;
`tests/test_oracle.py` checks this file against those fixtures on every run and,
when /root/reference is present, against the live reference as well.

Differences from the reference that are deliberate and documented:
  * variable keypoint counts: the oracle runs each pair on its valid prefix
    (`num0[b]`, `num1[b]`), which is the semantics SURVEY.md 8(c) defines for a
    padded batch (lightglue.py:494 never passes masks; masked_forward
    lightglue.py:248-254 equals the un-padded run on valid rows);
  * early exit: the reference crashes on `torch.stack([])` (SURVEY.md F4,
    lightglue.py:495-498,546-547).  The oracle returns the exit-layer
    descriptors as `ref_descriptors*` with one collected entry instead;
  * adaptive mode accepts B > 1 as independent pairs (reference asserts b == 1,
    lightglue.py:502,507).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

DEFAULT_CONF = {  # lightglue.py:323-343 (keys read on the forward path)
    "input_dim": 256,
    "add_scale_ori": False,
    "descriptor_dim": 256,
    "n_layers": 9,
    "num_heads": 4,
    "depth_confidence": -1,
    "width_confidence": -1,
    "filter_threshold": 0.0,
}


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"].to(x), sd[name + ".bias"].to(x))


def normalize_keypoints(kpts: torch.Tensor, size: Optional[torch.Tensor]) -> torch.Tensor:
    """lightglue.py:28-40.  kpts [n,2]; size [2] (W,H) or None."""
    if size is None:
        size = 1 + kpts.max(-2).values - kpts.min(-2).values
    size = size.to(kpts)
    return (kpts - size / 2) / (size.max(-1).values / 2)


def positional_encoding(wr: torch.Tensor, kpts: torch.Tensor):
    """lightglue.py:61-66.  Returns (cos, sin), each [n, head_dim], with every
    frequency repeated twice (repeat_interleave(2, -1))."""
    proj = kpts @ wr.to(kpts).t()
    return (
        torch.cos(proj).repeat_interleave(2, dim=-1),
        torch.sin(proj).repeat_interleave(2, dim=-1),
    )


def rotary(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """lightglue.py:43-50.  t [h, n, dh]; pairs are interleaved (2i, 2i+1)."""
    te, to = t[..., 0::2], t[..., 1::2]
    rot = torch.stack((-to, te), dim=-1).flatten(-2)
    return t * cos + rot * sin


def _ffn(sd, prefix, x, msg):
    """lightglue.py:144-149 / 179-184: Linear(512,512) LayerNorm GELU(erf) Linear(512,256)."""
    h = _lin(sd, prefix + ".ffn.0", torch.cat([x, msg], -1))
    h = F.layer_norm(
        h, (h.shape[-1],), sd[prefix + ".ffn.1.weight"].to(h), sd[prefix + ".ffn.1.bias"].to(h), 1e-5
    )
    return _lin(sd, prefix + ".ffn.3", F.gelu(h))


def self_block(sd, prefix, x, cos, sin, heads):
    """lightglue.py:151-164.  x [n, d]."""
    n, d = x.shape
    dh = d // heads
    qkv = _lin(sd, prefix + ".Wqkv", x).reshape(n, heads, dh, 3).permute(3, 1, 0, 2)
    q, k, v = rotary(qkv[0], cos, sin), rotary(qkv[1], cos, sin), qkv[2]
    attn = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ctx = (attn @ v).permute(1, 0, 2).reshape(n, d)
    return x + _ffn(sd, prefix, x, _lin(sd, prefix + ".out_proj", ctx))


def cross_block(sd, prefix, x0, x1, heads):
    """lightglue.py:193-222 (einsum branch).  x0 [n0,d], x1 [n1,d]."""
    d = x0.shape[-1]
    dh = d // heads

    def split(t):
        return t.reshape(t.shape[0], heads, dh).permute(1, 0, 2)

    qk0, qk1 = split(_lin(sd, prefix + ".to_qk", x0)), split(_lin(sd, prefix + ".to_qk", x1))
    v0, v1 = split(_lin(sd, prefix + ".to_v", x0)), split(_lin(sd, prefix + ".to_v", x1))
    sim = (qk0 @ qk1.transpose(-1, -2)) / math.sqrt(dh)
    m0 = torch.softmax(sim, dim=-1) @ v1
    m1 = torch.softmax(sim.transpose(-1, -2), dim=-1) @ v0
    m0 = _lin(sd, prefix + ".to_out", m0.permute(1, 0, 2).reshape(-1, d))
    m1 = _lin(sd, prefix + ".to_out", m1.permute(1, 0, 2).reshape(-1, d))
    return x0 + _ffn(sd, prefix, x0, m0), x1 + _ffn(sd, prefix, x1, m1)


def transformer_layer(sd, i, x0, x1, enc0, enc1, heads):
    """lightglue.py:231-245."""
    p = f"transformers.{i}"
    x0 = self_block(sd, p + ".self_attn", x0, enc0[0], enc0[1], heads)
    x1 = self_block(sd, p + ".self_attn", x1, enc1[0], enc1[1], heads)
    return cross_block(sd, p + ".cross_attn", x0, x1, heads)


def log_assignment(sd, i, x0, x1) -> torch.Tensor:
    """lightglue.py:257-288.  Returns scores [n0+1, n1+1]."""
    p = f"log_assignment.{i}"
    d = x0.shape[-1]
    md0 = _lin(sd, p + ".final_proj", x0) / d**0.25
    md1 = _lin(sd, p + ".final_proj", x1) / d**0.25
    sim = md0 @ md1.t()
    z0 = _lin(sd, p + ".matchability", x0)[:, 0]
    z1 = _lin(sd, p + ".matchability", x1)[:, 0]
    n0, n1 = sim.shape
    out = sim.new_zeros(n0 + 1, n1 + 1)
    if n0 > 0 and n1 > 0:
        out[:n0, :n1] = (
            F.log_softmax(sim, 1)
            + F.log_softmax(sim, 0)
            + F.logsigmoid(z0)[:, None]
            + F.logsigmoid(z1)[None, :]
        )
    out[:n0, n1] = F.logsigmoid(-z0)
    out[n0, :n1] = F.logsigmoid(-z1)
    return out


def filter_matches(scores: torch.Tensor, th: float):
    """lightglue.py:294-319 for one pair.  scores [n0+1, n1+1]."""
    n0, n1 = scores.shape[0] - 1, scores.shape[1] - 1
    if n0 == 0 or n1 == 0:
        return (
            torch.full((n0,), -1, dtype=torch.long),
            torch.full((n1,), -1, dtype=torch.long),
            scores.new_zeros(n0),
            scores.new_zeros(n1),
        )
    inner = scores[:-1, :-1]
    max0, max1 = inner.max(1), inner.max(0)
    m0, m1 = max0.indices, max1.indices
    mutual0 = torch.arange(n0) == m1[m0]
    mutual1 = torch.arange(n1) == m0[m1]
    ms0 = torch.where(mutual0, max0.values.exp(), max0.values.new_zeros(()))
    ms1 = torch.where(mutual1, ms0[m1], ms0.new_zeros(()))
    valid0 = mutual0 & (ms0 > th)
    valid1 = mutual1 & valid0[m1]
    return (
        torch.where(valid0, m0, torch.full_like(m0, -1)),
        torch.where(valid1, m1, torch.full_like(m1, -1)),
        ms0,
        ms1,
    )


def confidence_threshold(layer: int, n_layers: int) -> float:
    """lightglue.py:555-558; stored as an fp32 buffer (lightglue.py:403-408)."""
    t = 0.8 + 0.1 * math.exp(-4.0 * layer / n_layers)
    return float(torch.tensor(min(max(t, 0.0), 1.0), dtype=torch.float32))


def _token_conf(sd, i, x):
    """lightglue.py:75-80."""
    return torch.sigmoid(_lin(sd, f"token_confidence.{i}.token.0", x))[:, 0]


def _matchability(sd, i, x):
    """lightglue.py:290-291."""
    return torch.sigmoid(_lin(sd, f"log_assignment.{i}.matchability", x))[:, 0]


def forward_pair(
    sd: Dict[str, torch.Tensor],
    conf: dict,
    kpts0: torch.Tensor,
    kpts1: torch.Tensor,
    desc0: torch.Tensor,
    desc1: torch.Tensor,
    size0: Optional[torch.Tensor],
    size1: Optional[torch.Tensor],
    dtype=torch.float32,
    trace: Optional[dict] = None,
) -> dict:
    """One un-padded pair through lightglue.py:422-553.  kpts [n, 2 or 4]
    (scale/orientation already concatenated, lightglue.py:436-454)."""
    c = {**DEFAULT_CONF, **{k: v for k, v in conf.items() if k in DEFAULT_CONF}}
    heads, n_layers = c["num_heads"], c["n_layers"]
    n0, n1 = kpts0.shape[0], kpts1.shape[0]
    k0 = torch.cat([normalize_keypoints(kpts0[:, :2].float(), size0), kpts0[:, 2:].float()], -1)
    k1 = torch.cat([normalize_keypoints(kpts1[:, :2].float(), size1), kpts1[:, 2:].float()], -1)
    k0, k1 = k0.to(dtype), k1.to(dtype)
    x0, x1 = desc0.to(dtype), desc1.to(dtype)
    if c["input_dim"] != c["descriptor_dim"]:  # lightglue.py:352-355,464-465
        x0, x1 = _lin(sd, "input_proj", x0), _lin(sd, "input_proj", x1)
    enc0 = positional_encoding(sd["posenc.Wr.weight"], k0)
    enc1 = positional_encoding(sd["posenc.Wr.weight"], k1)

    early = c["depth_confidence"] > 0
    prune = c["width_confidence"] > 0
    ind0, ind1 = torch.arange(n0), torch.arange(n1)
    prune0 = torch.ones(n0, dtype=torch.long)
    prune1 = torch.ones(n1, dtype=torch.long)
    exit_layer = n_layers - 1
    for i in range(n_layers):
        x0, x1 = transformer_layer(sd, i, x0, x1, enc0, enc1, heads)
        if trace is not None:
            trace.setdefault("desc0", []).append(x0.clone())
            trace.setdefault("desc1", []).append(x1.clone())
        if i == n_layers - 1:
            break
        tok0 = tok1 = None
        if early:  # lightglue.py:501-505, 569-580
            tok0, tok1 = _token_conf(sd, i, x0), _token_conf(sd, i, x1)
            thr = confidence_threshold(i, n_layers)
            n_low = (tok0 < thr).float().sum() + (tok1 < thr).float().sum()
            if 1.0 - n_low / (n0 + n1) > c["depth_confidence"]:
                exit_layer = i
                break
        if prune:  # lightglue.py:506-521, 560-567
            thr = confidence_threshold(i, n_layers)
            keep0 = _matchability(sd, i, x0) > (1 - c["width_confidence"])
            keep1 = _matchability(sd, i, x1) > (1 - c["width_confidence"])
            if tok0 is not None:
                keep0 |= tok0 <= thr
                keep1 |= tok1 <= thr
            ind0, x0, enc0 = ind0[keep0], x0[keep0], (enc0[0][keep0], enc0[1][keep0])
            ind1, x1, enc1 = ind1[keep1], x1[keep1], (enc1[0][keep1], enc1[1][keep1])
            prune0[ind0] += 1
            prune1[ind1] += 1

    scores = log_assignment(sd, exit_layer, x0, x1)
    m0, m1, ms0, ms1 = filter_matches(scores, c["filter_threshold"])
    if prune:  # lightglue.py:527-536
        m0_ = torch.full((n0,), -1, dtype=torch.long)
        m1_ = torch.full((n1,), -1, dtype=torch.long)
        m0_[ind0] = torch.where(m0 == -1, m0, ind1[m0.clamp(min=0)]) if ind1.numel() else m0
        m1_[ind1] = torch.where(m1 == -1, m1, ind0[m1.clamp(min=0)]) if ind0.numel() else m1
        ms0_, ms1_ = torch.zeros(n0), torch.zeros(n1)
        ms0_[ind0], ms1_[ind1] = ms0.float(), ms1.float()
        m0, m1, ms0, ms1 = m0_, m1_, ms0_, ms1_
        p0, p1 = prune0, prune1
    else:  # lightglue.py:538-539
        p0 = torch.full((n0,), float(n_layers))
        p1 = torch.full((n1,), float(n_layers))
    return {
        "matches0": m0,
        "matches1": m1,
        "matching_scores0": ms0,
        "matching_scores1": ms1,
        "ref_descriptors0": x0[None],
        "ref_descriptors1": x1[None],
        "log_assignment": scores,
        "prune0": p0,
        "prune1": p1,
        "exit_layer": exit_layer,
        "ind0": ind0,
        "ind1": ind1,
    }


def _kpts_with_scale_ori(data, idx, conf):
    k = data[f"keypoints{idx}"]
    if conf.get("add_scale_ori", False):
        sc, ori = data[f"scales{idx}"], data[f"oris{idx}"]
        sc = sc if sc.dim() == 3 else sc[..., None]
        ori = ori if ori.dim() == 3 else ori[..., None]
        k = torch.cat([k, sc, ori], -1)
    return k


def forward(
    sd: Dict[str, torch.Tensor],
    conf: dict,
    data: dict,
    dtype=torch.float32,
    num0: Optional[List[int]] = None,
    num1: Optional[List[int]] = None,
) -> List[dict]:
    """Batch front end: returns one un-padded result dict per pair."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    k0, k1 = _kpts_with_scale_ori(data, 0, conf), _kpts_with_scale_ori(data, 1, conf)
    B = k0.shape[0]
    size0 = data.get("view0", {}).get("image_size")
    size1 = data.get("view1", {}).get("image_size")
    out = []
    for b in range(B):
        a0 = k0.shape[1] if num0 is None else int(num0[b])
        a1 = k1.shape[1] if num1 is None else int(num1[b])
        out.append(
            forward_pair(
                sd,
                conf,
                k0[b, :a0].cpu(),
                k1[b, :a1].cpu(),
                data["descriptors0"][b, :a0].cpu(),
                data["descriptors1"][b, :a1].cpu(),
                None if size0 is None else size0[b].cpu().float(),
                None if size1 is None else size1[b].cpu().float(),
                dtype=dtype,
            )
        )
    return out


def flops_per_pair(n: int, m: int, n_layers: int = 9, input_dim: int = 256) -> float:
    """Algorithmic FLOPs of one pair."""
    t = n + m
    f = n_layers * (2_490_368 * t + 1024 * (n * n + m * m) + 1536 * n * m) + 131_584 * t + 512 * n * m
    if input_dim != 256:
        f += 2 * input_dim * 256 * t
    return float(f)
