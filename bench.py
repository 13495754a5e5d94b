#!/usr/bin/env python
"""LightGlue matcher throughput bench (BASELINE.json metric: pairs/sec @ 2048 kpts).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one LightGlue forward over one batch of synthetic pairs.  At N = 1 the
workload is BASELINE.json configs[1]: 64 pairs x 2048 keypoints x 256-d, 9 layers,
random-init weights, bf16 tcgen05 kernels.  At N > 1 (torchrun, one rank per GPU) every
rank runs the same per-GPU workload on its own pairs (pair-sharded, no collective on the
hot path, "weak" scaling); NCCL only gathers the matches after the timed region.

Printed JSON line (rank 0):
  value        pairs/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          pairs/s through the public API with pinned HOST inputs: H2D of the step's
               keypoints/descriptors and D2H of matches/scores inside the timed region
  roofline     dominant kernel (flash attention) timed alone with CUDA events: algorithmic
               FLOPs / duration vs the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (oracle/lightglue_oracle.py, a port of the reference) timed on
               this box's host cores on a bounded sample of the same workload
--impl reference times that CPU port as the whole arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LightGlue pairs/sec @2048 kpts"
UNIT = "pairs/s"
KPTS = 2048
PAIRS_PER_GPU = 64


ATT_DRAM_BYTES_PER_LAUNCH = 402_726_144 + 116_338_176  # measured, see roofline.traffic below


def flops_per_pair(n, m, n_layers=9):
    t = n + m
    return float(n_layers * (2_490_368 * t + 1024 * (n * n + m * m) + 1536 * n * m) + 131_584 * t + 512 * n * m)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), src="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        # under load = samples in the upper half of the observed range
        load = [c for c in sm if c >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(precision="bf16", seed=0):
    from glue_factory_colon_b200 import LightGlue

    torch.manual_seed(seed)
    return LightGlue({"precision": precision, "filter_threshold": 0.1}).eval()


# ----------------------------------------------------------------------------- CPU arm


def cpu_port_time(budget_s, kpts, seed):
    """Times the CPU oracle (port of the reference) pair by pair; returns pairs/s and the sample."""
    from glue_factory_colon_b200.synthetic import make_pairs
    from oracle import lightglue_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model("fp32")
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    data = make_pairs(1, kpts, kpts, seed=seed)
    conf = {"filter_threshold": 0.1}
    with torch.no_grad():
        oracle.forward(sd, conf, data)  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            oracle.forward(sd, conf, data)
            n += 1
            el = time.perf_counter() - t0
            if el >= budget_s and n >= 2 or n >= 64:
                break
    return n / el, cores, f"{n} pairs x {kpts} kpts, fp32, {cores} torch threads, {el:.1f} s"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = args.steps, args.warmup
    from glue_factory_colon_b200.synthetic import make_pairs
    from oracle import lightglue_oracle as oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model("fp32")
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    conf = {"filter_threshold": 0.1}
    data = make_pairs(1, KPTS, KPTS, seed=7)
    steps = min(steps, 8)  # each step = 1 pair (~seconds on CPU); keep the arm within minutes
    with torch.no_grad():
        for _ in range(min(warm, 1)):
            oracle.forward(sd, conf, data)
        t0 = time.perf_counter()
        for _ in range(steps):
            oracle.forward(sd, conf, data)
        el = time.perf_counter() - t0
    v = steps / el
    sample = f"{steps} steps x 1 pair x {KPTS} kpts (bounded sample of the 64-pair batch), fp32, {cores} torch threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(warm, 1), "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.pairs} pairs/GPU x {KPTS} kpts x 256-d descriptors, 9 layers, random-init, "
                               f"{args.precision} (BASELINE configs[1])",
                   "pairs_per_gpu": args.pairs, "kpts": KPTS,
                   "note": "this arm: CPU port of the reference (oracle/) in fp32 on the host cores, one pair per step "
                           "(bounded sample of the same workload)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- GPU arm


def time_attention_alone(lib, S, Lp, reps=10):
    """Dominant kernel timed alone: self-attention at the bench shape, CUDA events on the launch stream."""
    from glue_factory_colon_b200._abi import BF16, ptr

    g = torch.Generator(device="cuda").manual_seed(0)
    q = (torch.randn(S * 4 * Lp, 64, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    k = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
    ctx = torch.empty(S * Lp, 256, device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 4.0 * S * 4 * Lp * Lp * 64  # QK^T and PV, 2 flops per MAC
    return ms, flops


def run_gpu_arm(args):
    import torch.distributed as dist

    from glue_factory_colon_b200 import _abi
    from glue_factory_colon_b200.synthetic import make_pairs, to_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.load()
    B = args.pairs
    model = build_model(args.precision).to(dev)
    host = make_pairs(B, KPTS, KPTS, seed=100 + rank)
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else {kk: vv.pin_memory() for kk, vv in v.items()})
              for k, v in host.items()}
    data = to_device(host, dev)
    h2d = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))
    h2d += sum(vv.numel() * vv.element_size() for v in host.values() if isinstance(v, dict) for vv in v.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        out = model(data)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.lgb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = model(data)
    e1.record()
    barrier()
    launches = lib.lgb200_launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * B / (ms * 1e-3)

    # ---- dominant kernel inside the running step: the same loop once more with CUDA events (launch stream) around
    # every self-attention launch (kept out of the timed region above: 36 event records per step) ----
    att_in_step_ms = None
    if rank == 0:
        model._attn_events = []
        for _ in range(args.steps):
            out = model(data)
        torch.cuda.synchronize()
        spans = [a.elapsed_time(b) for a, b in model._attn_events]
        model._attn_events = None
        att_in_step_ms = sum(spans) / len(spans)
    barrier()

    # ---- end to end through the public API with pinned host buffers ----
    d2h_keys = ["matches0", "matches1", "matching_scores0", "matching_scores1"]
    host_out = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in d2h_keys}
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    # Public-API pipeline a user would write: a copy stream uploads step i+1's pinned host batch while the
    # compute stream runs model(step i); each step's inputs are uploaded inside the timed region and each
    # step's matches/scores are read back to pinned host memory.  The device-side input buffers are two
    # persistent sets filled alternately (allocated once, outside the timed region): the loop itself never goes
    # through the caching allocator for them -- on some boxes of the pool (virtualised hosts) the earlier version,
    # which let `.to(device)` allocate 270 MB of fresh device tensors on the copy stream every step, ran at 55-60 ms
    # per step instead of 21 with identical kernels and an idle-link H2D rate of 55 GB/s.
    copy_stream = torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream(dev)
    dev_sets = [to_device(pinned, dev), to_device(pinned, dev)]
    free_ev = [None, None]  # compute-stream event after the last forward that read set j
    torch.cuda.synchronize()

    def copy_into(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                copy_into(dst[k], v)
            elif isinstance(v, torch.Tensor):
                dst[k].copy_(v, non_blocking=True)

    def upload(j):
        with torch.cuda.stream(copy_stream):
            if free_ev[j] is not None:
                copy_stream.wait_event(free_ev[j])
            copy_into(dev_sets[j], pinned)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    # The read-back runs on a third stream behind an event of the forward, so that the next forward does not queue
    # behind a device->host copy either; the timed region ends only after the last step's results are on the host.
    d2h_stream = torch.cuda.Stream(device=dev)

    def run_e2e(steps):
        ev = upload(0)
        for i in range(steps):
            j = i & 1
            compute.wait_event(ev)
            if i + 1 < steps:
                ev = upload(j ^ 1)
            o = model(dev_sets[j])
            done = torch.cuda.Event()
            done.record(compute)
            free_ev[j] = done
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                for k in d2h_keys:
                    host_out[k].copy_(o[k], non_blocking=True)
                    o[k].record_stream(d2h_stream)  # allocated on the compute stream, read on this one
        compute.wait_stream(d2h_stream)

    run_e2e(2)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * B / (ms_e2e * 1e-3)

    # ---- result gather (NCCL, outside the timed region) ----
    if world > 1:
        from glue_factory_colon_b200.shard import gather_to_rank0

        gathered = gather_to_rank0(out["matches0"], [B] * world)
        if rank == 0:
            assert gathered.shape[0] == world * B

    if rank == 0:
        peaks = load_peaks()
        F = flops_per_pair(KPTS, KPTS)
        att_alone_ms, att_flops = time_attention_alone(lib, 2 * B, KPTS)
        att_ms = att_in_step_ms
        att_tf = att_flops / (att_ms * 1e-3) / 1e12
        step_tf = (value / world) * F / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_port_time(args.cpu_budget, KPTS, seed=7)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {
                "workload": f"{B} pairs/GPU x {KPTS} kpts x 256-d descriptors, 9 layers, random-init, "
                            f"{args.precision} (BASELINE configs[1])",
                "pairs_per_gpu": B, "kpts": KPTS, "parallelism": f"pair-sharded x{world}, no hot-path collective",
                "l2": "working set per step (~3 GB activations + 1.07 GB log_assignment) exceeds the 126 MB L2",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "kernel": "tc_attention_kernel (self-attention, S=%d x 4 heads x %d^2)" % (2 * B, KPTS),
                "achieved": att_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": att_tf / peaks["bf16_sustained"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, ncu --set full
                # (profiles/r1_ncu_attention_summary.txt); algorithmic bytes are 4 x 134.2 MB
                "traffic": ATT_DRAM_BYTES_PER_LAUNCH if (2 * B, KPTS) == (128, 2048) else None,
                "traffic_unit": "bytes",
                "peak_source": peaks["src"] + " sustained (kernel timed inside the running step: average over the "
                                              "self-attention launches of %d steps, CUDA events on the launch stream)" % args.steps,
                "kernel_ms": att_ms, "flops_per_launch": att_flops,
                # the same kernel launched alone right after the loops, against the burst peak
                "alone": {"kernel_ms": att_alone_ms, "achieved": att_flops / (att_alone_ms * 1e-3) / 1e12,
                          "peak": peaks["bf16"], "frac": att_flops / (att_alone_ms * 1e-3) / 1e12 / peaks["bf16"],
                          "peak_source": peaks["src"] + " burst"},
                "whole_step": {"achieved": step_tf, "peak": peaks["bf16_sustained"],
                               "frac": step_tf / peaks["bf16_sustained"], "flops_per_pair": F,
                               "peak_source": peaks["src"] + " sustained"},
            },
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
