"""Single-pair latency (the reference's own eval flow runs the matcher at batch 1): eager launches vs CUDA-graph replay."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs
torch.manual_seed(0)
for kpts in (512, 1024, 2048):
    data = make_pairs(1, kpts, kpts, seed=7, device="cuda")
    for graph in (False, True):
        m = LightGlue({"precision": "bf16", "filter_threshold": 0.1, "cuda_graph": graph}).eval().cuda()
        for _ in range(3): m(data)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 50
        for _ in range(n): out = m(data)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        print(f"kpts {kpts} cuda_graph {graph}: {dt * 1e3:.3f} ms per pair (wall, synchronised at the end)  {1 / dt:.0f} pairs/s")
