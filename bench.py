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
  cpu_baseline the reference's own CPU path timed on this box's host cores on a bounded sample of the
               same workload: the UNMODIFIED reference module from git-ignored baseline/_ref (kind
               "reference") when that install travelled with the snapshot, else the CPU oracle
               (oracle/lightglue_oracle.py, kind "port")
  gpu_library  (N = 1) the unmodified reference module on the SAME GPU through PyTorch's library kernels
               (cuBLAS / SDPA): fp32, and flash=True under autocast(bf16) -- the "library path on the same box"
--impl reference times the CPU arm as the whole run (rank 0 only).
--workload c3 | c5 selects BASELINE configs[2] (ragged, <= 4096 kpts, cost-balanced pair shards) / configs[4]
(one point of the 512-8192 sweep, --kpts); the default c2 is the configuration the metric is quoted on.
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


ATT_DRAM_BYTES_PER_LAUNCH = 402_691_072 + 120_926_720  # dram__bytes_read.sum + dram__bytes_write.sum of the r2 capture
ATT_DRAM_SOURCE = "profiles/r2_ncu_full_attention_raw.csv (ncu --set full, one launch at S=128, Lp=2048; DRAM counters need the profiler, so this field is the committed capture of the same kernel, not a live measurement)"


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


# ----------------------------------------------------------------------------- shared config


def workload_config(args, world):
    """The `config` object of the JSON line -- identical in both arms (the driver compares them)."""
    if args.workload == "c3":
        wl = (f"{args.pairs} pairs/GPU, per-pair counts uniform in 1024..4096 padded to 4096, 256-d descriptors, "
              f"9 layers, random-init, {args.precision} (BASELINE configs[2])")
    elif args.workload == "c5":
        wl = (f"{args.pairs} pairs/GPU x {args.kpts} kpts x 256-d descriptors, 9 layers, random-init, "
              f"{args.precision} (BASELINE configs[4], one sweep point)")
    else:
        wl = (f"{args.pairs} pairs/GPU x {args.kpts} kpts x 256-d descriptors, 9 layers, random-init, "
              f"{args.precision} (BASELINE configs[1])")
    return {
        "workload": wl, "pairs_per_gpu": args.pairs, "kpts": args.kpts,
        "parallelism": f"pair-sharded x{world}, no hot-path collective",
        "l2": "working set per step (~3 GB activations + 1.07 GB log_assignment) exceeds the 126 MB L2",
    }


# ----------------------------------------------------------------------------- reference arms (CPU, library GPU)


def load_reference_module(conf):
    """The UNMODIFIED reference LightGlue from baseline/_ref (pip --target install of /root/reference made in the build
    container; git-ignored, travels with the snapshot) under torch.manual_seed(0), or None when it is not there.  Its
    one missing dependency, omegaconf, is served by the test-side shim in oracle/_shim."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "gluefactory").exists():
        return None
    for p in (str(ROOT / "oracle" / "_shim"), str(ref)):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        from gluefactory.models import get_model  # type: ignore

        torch.manual_seed(0)
        return get_model("matchers.lightglue")(dict(conf)).eval()
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] baseline/_ref present but not importable ({exc!r}); falling back to the port", file=sys.stderr)
        return None


def cpu_reference_runner(kpts, seed):
    """-> (run_one_pair, kind, cores): one fp32 forward of one pair on the host cores."""
    from glue_factory_colon_b200.synthetic import make_pairs

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    conf = {"filter_threshold": 0.1}
    data = make_pairs(1, kpts, kpts, seed=seed)
    ref = load_reference_module(conf)
    if ref is not None:
        def run():
            with torch.no_grad():
                ref(data)
        return run, "reference", cores
    from oracle import lightglue_oracle as oracle

    model = build_model("fp32")
    sd = {k: v.detach() for k, v in model.state_dict().items()}

    def run():
        with torch.no_grad():
            oracle.forward(sd, conf, data)
    return run, "port", cores


def cpu_reference_time(budget_s, kpts, seed):
    """Bounded sample: pairs of the bench shape, one per forward (the batch size the reference's eval flow uses)."""
    run, kind, cores = cpu_reference_runner(kpts, seed)
    run()  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        run()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s and n >= 2 or n >= 64:
            break
    what = "unmodified reference module (baseline/_ref)" if kind == "reference" else "oracle port"
    return n / el, cores, kind, f"{n} pairs x {kpts} kpts, {what}, fp32, {cores} torch threads, {el:.1f} s"


def gpu_library_time(dev, B, kpts, seed, reps=10, warm=3):
    """The unmodified reference module on the same GPU (PyTorch library kernels: cuBLAS, SDPA), protocol of the
    reference's utils/benchmark.py:7-33 (warm-ups, CUDA events around the forward only, inputs resident)."""
    from glue_factory_colon_b200.synthetic import make_pairs

    out = {}
    data = make_pairs(B, kpts, kpts, seed=seed, device=dev)
    for name, conf, amp in (("fp32", {"filter_threshold": 0.1}, False),
                            ("bf16_autocast_flash", {"filter_threshold": 0.1, "flash": True}, True)):
        ref = load_reference_module(conf)
        if ref is None:
            return None
        ref = ref.to(dev)
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                for _ in range(warm):
                    ref(data)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    ref(data)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "reps": reps, "warmup": warm}
        except Exception as exc:  # noqa: BLE001  (e.g. an SDPA backend missing for this shape)
            out[name] = {"error": repr(exc)[:200]}
        del ref
        torch.cuda.empty_cache()
    out["what"] = (f"unmodified reference module (baseline/_ref) on this GPU, {B} pairs x {kpts} kpts per forward, "
                   f"inputs resident, CUDA events")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    run, kind, cores = cpu_reference_runner(args.kpts, seed=7)
    steps, warm = min(args.steps, 64), min(args.warmup, 8)  # each step = 1 pair (~0.5 s on 16 cores): well under minutes
    for _ in range(warm):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    el = time.perf_counter() - t0
    v = steps / el
    what = "unmodified reference module from baseline/_ref" if kind == "reference" else "CPU port of the reference (oracle/)"
    sample = (f"{steps} steps x 1 pair x {args.kpts} kpts (bounded sample of the {args.pairs}-pair batch), {what}, fp32, "
              f"{cores} torch threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "arm": f"{what} in fp32 on the host cores, one pair per step (bounded sample of the same workload)",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
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
    KPTS = args.kpts
    model = build_model(args.precision).to(dev)
    total_pairs = world * B
    flops_step_all = total_pairs * flops_per_pair(KPTS, KPTS)  # algorithmic FLOPs of one step over all ranks
    if args.workload == "c3":
        # BASELINE configs[2]: world*B pairs with per-pair counts in 1024..4096 (one global, seeded list), split into
        # contiguous cost-balanced shards (shard.shard_bounds); every rank pads its own shard to 4096
        from glue_factory_colon_b200.shard import pair_cost, shard_bounds

        g = torch.Generator().manual_seed(3)
        n0_all = torch.randint(1024, 4097, (total_pairs,), generator=g)
        n1_all = torch.randint(1024, 4097, (total_pairs,), generator=g)
        costs = [pair_cost(int(a), int(b)) for a, b in zip(n0_all, n1_all)]
        lo, hi = shard_bounds(costs, world)[rank]
        flops_step_all = sum(flops_per_pair(int(a), int(b)) for a, b in zip(n0_all, n1_all))
        B = hi - lo
        KPTS = 4096
        host = make_pairs(B, KPTS, KPTS, seed=300 + rank, image_size=(512.0, 512.0))
        host["num_keypoints0"], host["num_keypoints1"] = n0_all[lo:hi].clone(), n1_all[lo:hi].clone()
    else:
        host = make_pairs(B, KPTS, KPTS, seed=100 + rank)
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else {kk: vv.pin_memory() for kk, vv in v.items()})
              for k, v in host.items()}
    counts = {k: host.pop(k) for k in ("num_keypoints0", "num_keypoints1") if k in host}  # read on the host
    data = {**to_device(host, dev), **counts}
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
    value = total_pairs / (ms * 1e-3)

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
    dev_sets = [{**to_device(pinned, dev), **counts}, {**to_device(pinned, dev), **counts}]
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
    e2e_value = total_pairs / (ms_e2e * 1e-3)

    # ---- result gather (NCCL, outside the timed region) ----
    if world > 1:
        from glue_factory_colon_b200.shard import gather_to_rank0

        sizes = [None] * world
        dist.all_gather_object(sizes, B)
        gathered = gather_to_rank0(out["matches0"], sizes)
        if rank == 0:
            assert gathered.shape[0] == total_pairs

    if rank == 0:
        peaks = load_peaks()
        c2_shape = args.workload != "c3"
        if c2_shape:
            att_alone_ms, att_flops = time_attention_alone(lib, 2 * B, KPTS)
        else:  # ragged launches differ in work: the kernel is timed alone at the C2 shape
            att_alone_ms, att_flops = time_attention_alone(lib, 128, 2048)
        att_ms = att_in_step_ms if c2_shape else att_alone_ms
        att_tf = att_flops / (att_ms * 1e-3) / 1e12
        step_tf = flops_step_all / world / (ms * 1e-3) / 1e12
        cpu = lib_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, kind, sample = cpu_reference_time(args.cpu_budget, args.kpts, seed=7)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        if world == 1 and not args.no_gpu_library and args.workload == "c2":
            del out, data
            torch.cuda.empty_cache()
            lib_gpu = gpu_library_time(dev, min(B, 16), KPTS, seed=100)
        fp32_tc = None
        if world == 1 and args.workload == "c2" and args.precision == "bf16" and not args.no_fp32_mode:
            # the drop-in's DEFAULT numerics (conf.mp False, no autocast -> precision "fp32"): the fp32-accurate
            # tensor-core mode (split-fp16 x3, csrc/lg_x3*.cu) on the same batch, device-resident, CUDA events
            m32 = build_model("fp32").to(dev)
            d32 = to_device(pinned, dev)
            for _ in range(2):
                m32(d32)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record()
            for _ in range(5):
                m32(d32)
            f1.record()
            torch.cuda.synchronize()
            ms32 = f0.elapsed_time(f1) / 5
            fp32_tc = {"value": B / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "steps": 5, "warmup": 2,
                       "what": "precision=fp32 (fp32-accurate mode on tcgen05: split-fp16 operands, 3 MMAs per product), "
                               "same batch, inputs resident"}
            del m32, d32
            torch.cuda.empty_cache()
        in_step = "kernel timed inside the running step: average over the self-attention launches of %d steps, CUDA events on the launch stream" % args.steps
        line = {
            "metric": METRIC if args.workload == "c2" and args.kpts == 2048 else f"LightGlue pairs/sec ({args.workload}, {args.kpts} kpts)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "kernel": "tc_attention_kernel (self-attention, S=%d x 4 heads x %d^2)" % (
                    (2 * B, KPTS) if c2_shape else (128, 2048)),
                "achieved": att_tf, "peak": peaks["bf16_sustained"] if c2_shape else peaks["bf16"], "unit": "TFLOP/s",
                "frac": att_tf / (peaks["bf16_sustained"] if c2_shape else peaks["bf16"]),
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape from the committed ncu
                # --set full capture (not re-measured by this run: DRAM counters need the profiler); algorithmic
                # bytes are 4 x 134.2 MB
                "traffic": ATT_DRAM_BYTES_PER_LAUNCH if (2 * B, KPTS) == (128, 2048) and args.precision == "bf16" else None,
                "traffic_unit": "bytes", "traffic_source": ATT_DRAM_SOURCE,
                "peak_source": peaks["src"] + (" sustained (" + in_step + ")" if c2_shape else " burst (kernel timed alone)"),
                "kernel_ms": att_ms, "flops_per_launch": att_flops,
                # the same kernel launched alone right after the loops, against the burst peak
                "alone": {"kernel_ms": att_alone_ms, "achieved": att_flops / (att_alone_ms * 1e-3) / 1e12,
                          "peak": peaks["bf16"], "frac": att_flops / (att_alone_ms * 1e-3) / 1e12 / peaks["bf16"],
                          "peak_source": peaks["src"] + " burst"},
                "whole_step": {"achieved": step_tf, "peak": peaks["bf16_sustained"],
                               "frac": step_tf / peaks["bf16_sustained"], "flops_per_step_per_gpu": flops_step_all / world,
                               "peak_source": peaks["src"] + " sustained"},
            },
            "cpu_baseline": cpu,
            "gpu_library": lib_gpu,
            "fp32_mode": fp32_tc,
            "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def _claim_stdout():
    """Rank 0 prints ONE JSON line on stdout and nothing else: keep a private duplicate of file descriptor 1 for that
    line and point fd 1 at stderr, so that whatever native libraries write to "stdout" (NCCL prints its version banner
    there under torchrun when NCCL_DEBUG is set) lands on stderr instead of next to the result."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-library", action="store_true", help="skip the reference-module-on-this-GPU comparator")
    ap.add_argument("--no-fp32-mode", action="store_true", help="skip the extra precision=fp32 (tensor-core) measurement")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"],
                    help="c2: BASELINE configs[1] (default); c3: ragged <= 4096 kpts, cost-balanced shards; c5: sweep point")
    ap.add_argument("--kpts", type=int, default=0, help="keypoints per image (c2/c5; default 2048)")
    args = ap.parse_args()
    if args.workload == "c3":
        args.kpts = 4096
        if args.pairs == PAIRS_PER_GPU:
            args.pairs = 32
    elif args.kpts <= 0:
        args.kpts = KPTS
    if args.workload == "c5" and args.pairs == PAIRS_PER_GPU:
        args.pairs = max(1, 131072 // args.kpts)  # total tokens per batch held at 2 x 131072
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
