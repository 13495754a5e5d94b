"""Config 3 (32 pairs, 1024-4096 keypoints padded to 4096, key-padding masks): one warm-up + timed forwards; used under ncu."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs
from oracle.lightglue_oracle import flops_per_pair
torch.manual_seed(0)
B = 32
g = torch.Generator().manual_seed(3000)
n0 = torch.randint(1024, 4097, (B,), generator=g)
n1 = torch.randint(1024, 4097, (B,), generator=g)
model = LightGlue({"precision": "bf16", "filter_threshold": 0.1}).eval().cuda()
data = make_pairs(B, 4096, 4096, seed=300, image_size=(512.0, 512.0), device="cuda")
data["num_keypoints0"], data["num_keypoints1"] = n0, n1
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
model(data); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): out = model(data)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = sum(flops_per_pair(int(a), int(b)) for a, b in zip(n0, n1))
print(f"C3: {ms:.3f} ms/step  {B / ms * 1e3:.1f} pairs/s  {fl / ms / 1e9:.1f} TFLOP/s over valid tokens  (attention share of flops {sum(9*(1024*(int(a)**2+int(b)**2)+1536*int(a)*int(b)) for a,b in zip(n0,n1))/fl:.2f})")
