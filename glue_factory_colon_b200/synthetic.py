"""Seeded synthetic LightGlue inputs: keypoints uniform in the
image, unit descriptors; image 1 holds a permuted, warped, noisy copy of 60 % of
image 0's points so that the matcher has real structure to find."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


def make_pairs(
    B: int,
    n0: int,
    n1: int,
    seed: int = 0,
    image_size=(640.0, 480.0),
    dim: int = 256,
    with_size: bool = True,
    scale_ori: bool = False,
    overlap: float = 0.6,
    device: Optional[torch.device] = None,
    with_gt: bool = False,
) -> dict:
    W, H = image_size
    wh = torch.tensor([W, H], dtype=torch.float32)
    Hm = torch.tensor([[0.98, 0.03, 6.0], [-0.02, 1.01, -4.0], [1e-5, -2e-5, 1.0]])
    k0s, k1s, d0s, d1s = [], [], [], []
    gm0 = torch.full((B, n0), -1, dtype=torch.long)
    gm1 = torch.full((B, n1), -1, dtype=torch.long)
    for p in range(B):
        g = torch.Generator().manual_seed(1000 * seed + p)
        k0 = torch.rand(n0, 2, generator=g) * wh
        d0 = F.normalize(torch.randn(n0, dim, generator=g), dim=-1)
        k1 = torch.rand(n1, 2, generator=g) * wh
        d1 = F.normalize(torch.randn(n1, dim, generator=g), dim=-1)
        nm = int(min(n0, n1) * overlap)
        if nm > 0:
            src = torch.randperm(n0, generator=g)[:nm]
            dst = torch.randperm(n1, generator=g)[:nm]
            ph = torch.cat([k0[src], torch.ones(nm, 1)], -1) @ Hm.t()
            k1[dst] = (ph[:, :2] / ph[:, 2:]).clamp_(min=0.0) + torch.randn(nm, 2, generator=g)
            k1[dst] = torch.minimum(k1[dst].clamp_(min=0.0), wh - 1)
            d1[dst] = F.normalize(d0[src] + 0.05 * torch.randn(nm, dim, generator=g), dim=-1)
            gm0[p, src], gm1[p, dst] = dst, src
        k0s.append(k0), k1s.append(k1), d0s.append(d0), d1s.append(d1)
    data = {
        "keypoints0": torch.stack(k0s),
        "keypoints1": torch.stack(k1s),
        "descriptors0": torch.stack(d0s),
        "descriptors1": torch.stack(d1s),
        "view0": {},
        "view1": {},
    }
    if with_size:
        data["view0"]["image_size"] = wh[None].repeat(B, 1)
        data["view1"]["image_size"] = wh[None].repeat(B, 1)
    if scale_ori:
        g = torch.Generator().manual_seed(1000 * seed + 999)
        for i, n in ((0, n0), (1, n1)):
            data[f"scales{i}"] = torch.rand(B, n, generator=g) * 4 + 1
            data[f"oris{i}"] = (torch.rand(B, n, generator=g) - 0.5) * 6.28
    if with_gt:
        # ground truth in the reference's format (gluefactory/geometry/gt_generation.py): matches (-1 = unmatchable,
        # -2 = ignored) and the dense boolean assignment read by NLLLoss (models/utils/losses.py:62-73).  Every 11th
        # unmatched point is marked "ignored" so that the -1 / -2 distinction is exercised.  (No extra random draws:
        # the inputs above are identical with and without ground truth.)
        for gm in (gm0, gm1):
            un = (gm == -1).nonzero()
            gm[un[::11, 0], un[::11, 1]] = -2
        ga = torch.zeros(B, n0, n1, dtype=torch.bool)
        bi, ii = (gm0 >= 0).nonzero(as_tuple=True)
        ga[bi, ii, gm0[bi, ii]] = True
        data["gt_matches0"], data["gt_matches1"], data["gt_assignment"] = gm0, gm1, ga
    if device is not None:
        data = to_device(data, device)
    return data


def to_device(data: dict, device, non_blocking: bool = False) -> dict:
    out = {}
    for k, v in data.items():
        if isinstance(v, dict):
            out[k] = to_device(v, device, non_blocking)
        elif isinstance(v, torch.Tensor):
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out
