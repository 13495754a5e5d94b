"""Padding producers that keep the per-sequence counts.

The reference pads variable-length local features to a fixed length so that they can be batched
(`pad_to_length` / `pad_and_stack`, gluefactory/models/utils/misc.py:19-62 and :103-113; `pad_local_features`,
models/cache_loader.py:17-45; the Endomapper loader, datasets/endomapper.py:452-488) and then forgets how many
entries were real: the padding is *random* keypoints and descriptors so that the matcher, which has no mask
input (lightglue.py:422-553), is merely unlikely to match them.  The B200 matcher takes the counts
(`num_keypoints0/1`) and masks padded keys in every attention and in both softmax normalisers, so padded rows can
never match and never change a valid row.  These producers therefore return the same padded tensors as the
reference's (same modes, same bounds, same values on the valid prefix; the same torch RNG draws for the random
modes) **plus the counts**, and `matcher_inputs` assembles the dict `LightGlue.forward` takes.

Host-side tensor bookkeeping only; no arithmetic of the hot path lives here.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

# constant fills by mode name (misc.py:35-44: "zeros", "ones", "minus_one", False / "false")
_CONSTANT_FILL = {"zeros": 0, "false": 0, "ones": 1, "minus_one": -1}

# which mode each local-feature key is padded with, and along which dim (cache_loader.py:17-45;
# the Endomapper loader adds the last four, endomapper.py:474-488)
_LOCAL_FEATURE_RULES = {
    "keypoints": (-2, "random_c"),
    "descriptors": (-2, "random"),
    "keypoint_scores": (-1, "zeros"),
    "scales": (-1, "zeros"),
    "oris": (-1, "zeros"),
    "depth_keypoints": (-1, "zeros"),
    "valid_depth_keypoints": (-1, "zeros"),
    "sparse_depth": (-1, "minus_one"),
    "point3D_ids": (-1, "minus_one"),
    "valid_depth_mask": (-1, False),
    "valid_3D_mask": (-1, False),
}


def _mode_key(mode) -> str:
    if mode is False:
        return "false"
    if not isinstance(mode, str):
        raise ValueError(mode)
    return mode.lower()


def _tail(x: torch.Tensor, shape: List[int], d: int, key: str, bounds) -> torch.Tensor:
    """The padding block of `shape` for mode `key` (values as the reference draws them, misc.py:35-62)."""
    if key in _CONSTANT_FILL:
        return torch.full(shape, _CONSTANT_FILL[key], device=x.device, dtype=x.dtype)
    lo, hi = bounds
    if key == "random":  # one uniform range for the whole tensor: the data's own min / max unless bounded
        lo = x.min() if lo is None else lo
        hi = x.max() if hi is None else hi
        return torch.empty(shape, device=x.device).uniform_(lo, hi)
    if key == "random_c":  # one range per channel of the last dim; `bounds` only used for empty sequences
        cols = []
        for c in range(shape[-1]):
            c_lo, c_hi = (x[..., c].min(), x[..., c].max()) if d > 0 else (lo, hi)
            cols.append(torch.empty(shape[:-1] + [1], device=x.device).uniform_(c_lo, c_hi))
        return torch.cat(cols, dim=-1)
    raise ValueError(key)


def pad_to_length(
    x: torch.Tensor,
    length: int,
    pad_dim: int = -2,
    mode: Union[str, bool] = "zeros",
    bounds: Tuple[Optional[float], Optional[float]] = (None, None),
    return_count: bool = False,
):
    """`x` extended along `pad_dim` to `length` entries (misc.py:19-62: same modes, bounds and error behaviour --
    AssertionError when `x` is longer than `length`, ValueError on an unknown mode; `x` itself is returned when
    it already has the length).  With `return_count` the number of real entries comes back as well."""
    d = x.shape[pad_dim]
    assert d <= length
    if d == length:
        return (x, d) if return_count else x
    key = _mode_key(mode)
    shape = list(x.shape)
    shape[pad_dim] = length - d
    out = torch.cat([x, _tail(x, shape, d, key, bounds)], dim=pad_dim)
    return (out, d) if return_count else out


def pad_and_stack(
    sequences: Sequence[torch.Tensor],
    length: Optional[int] = None,
    pad_dim: int = -2,
    return_counts: bool = False,
    **kwargs,
):
    """Stack of the sequences, each padded to `length` (default: the longest; misc.py:103-113).  With
    `return_counts` also the int32 vector of real lengths -- what `num_keypoints0/1` wants."""
    if length is None:
        length = max(int(s.shape[pad_dim]) for s in sequences)
    stacked = torch.stack([pad_to_length(s, length, pad_dim, **kwargs) for s in sequences], 0)
    if not return_counts:
        return stacked
    return stacked, torch.tensor([int(s.shape[pad_dim]) for s in sequences], dtype=torch.int32)


def pad_local_features(pred: Dict[str, torch.Tensor], seq_l: int, bounds=(None, None), deterministic: bool = False) -> dict:
    """Pads one image's local features to `seq_l` entries in place and records `num_keypoints` (the count before
    padding).  Keys, dims and modes follow cache_loader.py:17-45; `bounds` is forwarded to the keypoint padding as
    the Endomapper loader does (endomapper.py:454-460).  `deterministic=True` pads keypoints and descriptors with
    zeros instead of random draws: with the counts recorded the values of the padding cannot influence the
    matcher, and a loader that wants bit-reproducible batches can drop the RNG dependence."""
    count = int(pred["keypoints"].shape[-2])
    for name, (dim, mode) in _LOCAL_FEATURE_RULES.items():
        if name not in pred:
            continue
        if deterministic and mode in ("random", "random_c"):
            mode = "zeros"
        kw = {"bounds": bounds} if name == "keypoints" else {}
        pred[name] = pad_to_length(pred[name], seq_l, dim, mode=mode, **kw)
    pred["num_keypoints"] = torch.tensor(count, dtype=torch.int32)
    return pred


def matcher_inputs(features0: Sequence[dict], features1: Sequence[dict], image_sizes0=None, image_sizes1=None) -> dict:
    """Batches padded per-image feature dicts (outputs of `pad_local_features`, all of one length per side) into
    the input dict of `LightGlue.forward`: `keypoints0/1`, `descriptors0/1`, optional `scales*/oris*`,
    `num_keypoints0/1` and `view0/1.image_size` ([B,2] as (W,H)) when sizes are given."""
    if len(features0) != len(features1) or not features0:
        raise ValueError("need the same, non-zero number of feature dicts for both images")
    data: dict = {"view0": {}, "view1": {}}
    for side, feats, sizes in ((0, features0, image_sizes0), (1, features1, image_sizes1)):
        for name in ("keypoints", "descriptors", "scales", "oris"):
            have = [name in f for f in feats]
            if not any(have):
                continue
            if not all(have):
                raise ValueError(f"{name} present for some images of side {side} only")
            data[f"{name}{side}"] = torch.stack([f[name] for f in feats], 0)
        data[f"num_keypoints{side}"] = torch.stack(
            [torch.as_tensor(f.get("num_keypoints", f["keypoints"].shape[-2]), dtype=torch.int32) for f in feats], 0
        )
        if sizes is not None:
            data[f"view{side}"]["image_size"] = torch.as_tensor(sizes, dtype=torch.float32).reshape(len(feats), 2)
    return data
