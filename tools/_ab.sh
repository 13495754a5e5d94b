cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/gputest_all.log 2>&1; tail -8 gpurun_out/gputest_all.log | cut -c1-300
python - <<'PY' 2>&1 | tail -5
import torch
a = torch.randn(64, 128, device="cuda").half(); b = torch.randn(128, 32, device="cuda").half()
r = torch.mm(a, b, out_dtype=torch.float32)
try:
    r2 = torch.addmm(r, a, b, out_dtype=torch.float32)
    print("addmm out_dtype ok", float((r2 - 2 * r).abs().max()))
except Exception as e:
    print("addmm ERR", str(e)[:200])
try:
    r3 = torch.baddbmm(torch.zeros(2, 64, 32, device="cuda"), a[None].expand(2, -1, -1), b[None].expand(2, -1, -1), out_dtype=torch.float32)
    print("baddbmm out_dtype ok")
except Exception as e:
    print("baddbmm ERR", str(e)[:200])
PY
