"""Attention kernel alone under the power cap: 600 warm-up launches (the board needs ~0.4 s to reach its 1000 W limit,
after which the SM clock settles near 1.67 GHz), then 1500 timed launches.  This, not a 10-launch burst at 1.96 GHz, is
the regime the kernel runs in inside the bench step.  usage: LGB200_LIB=... python tools/attn_sustained.py [kv_xor]"""
import os, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, ptr
lib = _abi.load()
S, Lp = 128, 2048
kv_xor = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(S * 4 * Lp, 64, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
k = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
ctx = torch.empty(S * Lp, 256, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, kv_xor, ptr(ctx), st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
import subprocess, threading, time, statistics
run(3)
time.sleep(0.5)
burst = run(10)
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip().split(", ")) for l in proc.stdout], daemon=True).start()
run(600)
n0 = len(rows)
sus = run(1500)
load = rows[n0:]
proc.terminate()
clk = statistics.median(float(r[0]) for r in load) if load else 0
pw = statistics.median(float(r[1]) for r in load) if load else 0
print(f"burst {burst:.3f} ms  sustained {sus:.3f} ms  ({4.0*S*4*Lp*Lp*64/sus/1e9:.0f} TFLOP/s sustained)  clock {clk:.0f} MHz  power {pw:.0f} W  Mcycles/launch {sus*clk/1e3:.3f}")
