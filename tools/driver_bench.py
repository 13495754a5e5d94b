"""Pair driver on a stream of pairs of mixed size (256 pairs, 512-4096 keypoints): order-preserving vs size-sorted batches."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.driver import BatchedPairMatcher
from glue_factory_colon_b200.synthetic import make_pairs
torch.manual_seed(0)
model = LightGlue({"precision": "bf16", "filter_threshold": 0.1}).eval().cuda()
g = torch.Generator().manual_seed(5)
sizes = [(int(a), int(b)) for a, b in torch.randint(512, 4097, (256, 2), generator=g)]
pairs = []
for i, (a, b) in enumerate(sizes):
    d = make_pairs(1, a, b, seed=1000 + i)
    pairs.append({"keypoints0": d["keypoints0"][0], "keypoints1": d["keypoints1"][0], "descriptors0": d["descriptors0"][0],
                  "descriptors1": d["descriptors1"][0], "image_size0": d["view0"]["image_size"][0], "image_size1": d["view1"]["image_size"][0]})
for by_size in (False, True, False, True):
    drv = BatchedPairMatcher(model, max_pairs=32, max_tokens=32 * 2 * 4096, sort_by_size=by_size)
    list(drv.match(pairs[:32]))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = list(drv.match(pairs))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"sort_by_size={by_size}: {len(out)} pairs in {dt * 1e3:.1f} ms = {len(out) / dt:.0f} pairs/s (host collation + upload + forward + read-back)")
