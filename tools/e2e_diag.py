"""Why is the end-to-end step slow on some boxes of the pool?  Times, with CUDA events: the 270 MB pinned upload alone,
the forward alone, both concurrently (copy stream + compute stream), and the bench's e2e loop without the clock sampler."""
import json, sys, time, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import LightGlue
from glue_factory_colon_b200.synthetic import make_pairs, to_device

dev = "cuda:0"
torch.manual_seed(0)
model = LightGlue({"precision": "bf16", "filter_threshold": 0.1}).eval().to(dev)
host = make_pairs(64, 2048, 2048, seed=100)
pinned = {k: ({kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v.pin_memory()) for k, v in host.items()}
data = to_device(pinned, dev)
copy_stream, compute = torch.cuda.Stream(device=dev), torch.cuda.current_stream(dev)
for _ in range(3):
    model(data)
torch.cuda.synchronize()


def ev():
    return torch.cuda.Event(enable_timing=True)


res = {}
# 1. upload alone
a, b = ev(), ev()
with torch.cuda.stream(copy_stream):
    a.record(copy_stream)
    for _ in range(5):
        d = to_device(pinned, dev, non_blocking=True)
    b.record(copy_stream)
torch.cuda.synchronize()
res["upload_alone_ms"] = a.elapsed_time(b) / 5
# 2. forward alone
a, b = ev(), ev()
a.record()
for _ in range(5):
    model(data)
b.record()
torch.cuda.synchronize()
res["forward_alone_ms"] = a.elapsed_time(b) / 5
# 3. concurrent: 5 uploads on the copy stream while 5 forwards run
a, b, c, e = ev(), ev(), ev(), ev()
t0 = time.perf_counter()
with torch.cuda.stream(copy_stream):
    a.record(copy_stream)
    for _ in range(5):
        d = to_device(pinned, dev, non_blocking=True)
    b.record(copy_stream)
t1 = time.perf_counter()
c.record()
for _ in range(5):
    model(data)
e.record()
t2 = time.perf_counter()
torch.cuda.synchronize()
res["concurrent_upload_ms"] = a.elapsed_time(b) / 5
res["concurrent_forward_ms"] = c.elapsed_time(e) / 5
res["host_enqueue_upload_ms"] = (t1 - t0) * 1e3 / 5
res["host_enqueue_forward_ms"] = (t2 - t1) * 1e3 / 5
# 4. the bench's pipeline, no sampler
host_out = {k: torch.empty(s, dtype=t).pin_memory() for k, s, t in (("matches0", (64, 2048), torch.int64), ("matching_scores0", (64, 2048), torch.float32))}


def upload():
    with torch.cuda.stream(copy_stream):
        d = to_device(pinned, dev, non_blocking=True)
        x = torch.cuda.Event()
        x.record(copy_stream)
    return d, x


def run(steps):
    nxt = upload()
    for i in range(steps):
        d, x = nxt
        compute.wait_event(x)
        if i + 1 < steps:
            nxt = upload()
        o = model(d)
        for k in host_out:
            host_out[k].copy_(o[k], non_blocking=True)
        for t in (d["keypoints0"], d["keypoints1"], d["descriptors0"], d["descriptors1"]):
            t.record_stream(compute)


run(2)
torch.cuda.synchronize()
a, b = ev(), ev()
t0 = time.perf_counter()
a.record()
run(10)
b.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
res["pipeline_ms_per_step"] = a.elapsed_time(b) / 10
res["pipeline_host_ms_per_step"] = (t1 - t0) * 1e3 / 10
res["mem_reserved_GB"] = torch.cuda.memory_reserved() / 1e9
print(json.dumps({k: round(v, 3) for k, v in res.items()}))
