"""Adaptor between the reference's matcher call site and the batched pair driver.

The reference reaches the matcher in two places:

  * `TwoViewPipeline._forward` (models/two_view_pipeline.py:326-335): `pred = {**pred, **self.matcher({**data, **pred})}`
    on one dataloader batch, between two device synchronisations (`_profile_call`, :78-102);
  * the eval loops, which call the whole pipeline at batch 1 from `export_predictions`
    (utils/export_predictions.py:21-90): model(data) -> callback_fn -> key filter -> keypoint renormalisation ->
    `v[0].cpu().numpy()` -> one HDF5 group per `data["name"][0]`.

A one-pair call leaves a B200 idle (a 2048-keypoint pair is ~1 ms of launches), so this module lets the same flow
feed `driver.BatchedPairMatcher` instead:

  * `matcher_inputs_to_pair(d)`  -- the dict `{**data, **pred}` a pipeline hands to its matcher (batch-1 tensors,
    `view0/1.image_size`) -> the per-pair dict the driver takes;
  * `StreamedMatcher(matcher)`   -- `.match(items)` takes the stream of those call-site dicts and yields, in input
    order, `(item, matcher_pred)` with `matcher_pred` = the batch-1 dict the reference's matcher would have returned
    for that item (`matches0/1` int64 [1,N], `matching_scores0/1` fp32 [1,N]), ready for `pred = {**pred, **matcher_pred}`;
  * `export_matches(items, matcher, writer, ...)` -- export_predictions' post-processing (same arguments, same
    errors) on top of the streamed matcher; `writer(name, arrays)` receives what the reference writes into the
    HDF5 group (`H5Writer` does exactly that when h5py is importable; `DictWriter` collects in memory).

Host logic only: all compute is `LightGlue.forward` (C-ABI kernels) through the driver.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .driver import BatchedPairMatcher

MATCHER_KEYS = ("matches0", "matches1", "matching_scores0", "matching_scores1")


def matcher_inputs_to_pair(d: dict) -> dict:
    """`{**data, **pred}` as built at two_view_pipeline.py:327 (batch dimension 1) -> per-pair dict for the driver
    (CPU tensors: the driver stages them in pinned memory).  Required keys as lightglue.py:345,423-424."""
    for key in ("keypoints0", "keypoints1", "descriptors0", "descriptors1"):
        assert key in d, f"Missing key {key} in data"
    if d["keypoints0"].dim() == 3:
        assert d["keypoints0"].shape[0] == 1, "the streamed matcher takes one pair per call-site dict"
    pair = {}
    for k in ("keypoints0", "keypoints1", "descriptors0", "descriptors1", "scales0", "scales1", "oris0", "oris1"):
        if isinstance(d.get(k), torch.Tensor):
            t = d[k].detach()
            if k.startswith(("scales", "oris")):  # [1,N] or [1,N,1] (lightglue.py:436-454)
                t = t.reshape(-1)
            elif t.dim() == 3:
                t = t[0]
            pair[k] = t.to("cpu", torch.float32)
    for i in (0, 1):
        view = d.get(f"view{i}")
        size = view.get("image_size") if isinstance(view, dict) else None
        if size is not None:
            pair[f"image_size{i}"] = torch.as_tensor(size, dtype=torch.float32).reshape(-1)[:2].cpu()
    return pair


class StreamedMatcher:
    """Runs the matcher stage of a stream of pipeline items through `BatchedPairMatcher`."""

    def __init__(self, matcher, out_device: Optional[torch.device] = None, **driver_kwargs):
        self.driver = BatchedPairMatcher(matcher, **driver_kwargs)
        self.out_device = out_device

    def match(self, items: Iterable[dict]) -> Iterator[Tuple[dict, Dict[str, torch.Tensor]]]:
        """items: call-site dicts (`{**data, **pred}`).  The driver pulls ahead of the results it has yielded (it
        plans batches over a window of pairs), so the items are remembered until their result is out."""
        kept: List[dict] = []

        def pairs():
            for it in items:
                kept.append(it)
                yield matcher_inputs_to_pair(it)

        for i, res in enumerate(self.driver.match(pairs())):
            item, kept[i] = kept[i], None
            pred = {k: res[k][None] for k in MATCHER_KEYS}  # batch-1 tensors, like the matcher's own output
            if "log_assignment" in res:
                pred["log_assignment"] = res["log_assignment"][None]
            if self.out_device is not None:
                pred = {k: v.to(self.out_device) for k, v in pred.items()}
            yield item, pred


# ------------------------------------------------------------------------------------------------ export


class DictWriter(dict):
    """`writer(name, arrays)`: keeps everything in memory ({name: {key: ndarray}})."""

    def __call__(self, name, arrays):
        self[name] = arrays


class H5Writer:
    """The reference's storage (utils/export_predictions.py:33,83-88): one HDF5 group per name, one dataset per key;
    a group that cannot be created (duplicate name) is skipped, as the reference does."""

    def __init__(self, output_file):
        import h5py  # absent in the build image; present wherever the reference's eval flow runs
        from pathlib import Path

        Path(output_file).parent.mkdir(exist_ok=True, parents=True)
        self.file = h5py.File(str(output_file), "w")

    def __call__(self, name, arrays):
        try:
            grp = self.file.create_group(name)
            for k, v in arrays.items():
                grp.create_dataset(k, data=v)
        except (RuntimeError, ValueError):
            pass

    def close(self):
        self.file.close()


def export_matches(items: Iterable[dict], matcher, writer: Callable[[str, Dict[str, np.ndarray]], None],
                   as_half: bool = False, keys="*", callback_fn=None, optional_keys: Sequence[str] = (),
                   **driver_kwargs) -> int:
    """utils/export_predictions.py:21-90 with the matcher stage streamed in padded batches.

    `items` yields what the reference's loop would hand to the matcher: the dataloader's `data` merged with the
    extractor's `pred` (keypoints / descriptors of both views, batch 1), plus `name`.  Per item, in input order:
    pred = {**item_pred, **matcher_pred}; callback_fn(pred, data) merged underneath; key filter (ValueError on a
    missing key); keypoints multiplied by 1 / view.scales; `[0].cpu().numpy()`; optional fp16 cast; writer(name, ...).
    Returns the number of items written."""
    assert keys == "*" or isinstance(keys, (tuple, list))
    n = 0
    sm = StreamedMatcher(matcher, **driver_kwargs)
    with torch.no_grad():
        for data, mpred in sm.match(items):
            name = data.get("name", [None])[0] if not isinstance(data.get("name"), str) else data["name"]
            # what the pipeline's `pred` holds at this point: the extractor outputs it was given, then the matcher's
            pred = {k: v for k, v in data.items()
                    if isinstance(v, torch.Tensor) and k.startswith(("keypoints", "descriptors", "keypoint_scores"))}
            pred.update(mpred)
            if callback_fn is not None:
                pred = {**callback_fn(pred, data), **pred}
            if keys != "*":
                if len(set(keys) - set(pred.keys())) > 0:
                    raise ValueError(f"Missing key {set(keys) - set(pred.keys())}")
                pred = {k: v for k, v in pred.items() if k in list(keys) + list(optional_keys)}
            assert len(pred) > 0
            for k in list(pred.keys()):  # renormalization (export_predictions.py:55-73)
                if k.startswith("keypoints"):
                    idx = k.replace("keypoints", "")
                    src = data if len(idx) == 0 else data.get(f"view{idx}", {})
                    if isinstance(src, dict) and "scales" in src:
                        scales = 1.0 / torch.as_tensor(src["scales"]).to(pred[k])
                        pred[k] = pred[k] * scales[None]
            arrays = {k: v[0].cpu().numpy() for k, v in pred.items()}
            if as_half:
                for k in arrays:
                    if arrays[k].dtype == np.float32:
                        arrays[k] = arrays[k].astype(np.float16)
            writer(name, arrays)
            n += 1
    return n
