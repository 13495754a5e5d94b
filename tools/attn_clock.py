"""Attention kernel alone in a long loop with nvidia-smi clock / power samples and an in-kernel-free estimate of the SM
clock (cycles are not needed: the question is whether the power cap pulls the clock down under this kernel)."""
import os, subprocess, sys, threading, time, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from glue_factory_colon_b200 import _abi
from glue_factory_colon_b200._abi import BF16, ptr
lib = _abi.load()
S, Lp = 128, 2048
g = torch.Generator(device="cuda").manual_seed(0)
q = (torch.randn(S * 4 * Lp, 64, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
k = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(S * 4 * Lp, 64, device="cuda", generator=g).to(torch.bfloat16)
ctx = torch.empty(S * Lp, 256, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
time.sleep(0.3)
n0 = len(rows)
for n in (20, 100, 400, 1500):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        lib.lgb200_attention(BF16, ptr(q), ptr(k), ptr(v), S, Lp, None, 0, ptr(ctx), st)
    e1.record(); torch.cuda.synchronize()
    print(f"{n} launches: {e0.elapsed_time(e1)/n:.4f} ms each")
time.sleep(0.1)
proc.terminate()
print("idle samples:", rows[:n0][-3:])
load = rows[n0:]
print("samples under load:", len(load))
for i in range(0, len(load), max(1, len(load)//25)):
    print(load[i])
