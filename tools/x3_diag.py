"""fp32 tensor-core mode (x3) vs CUDA-core fp32 vs the CPU oracle: where does the error come from?"""
import sys, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_model, load_c1_fixture, make_pairs, oracle_batch, sharp_assignment_overrides
from glue_factory_colon_b200.synthetic import to_device

def report(tag, model, conf, data):
    res = oracle_batch(model.cpu(), conf, data)
    res64 = oracle_batch(model.cpu(), conf, data, dtype=torch.float64)
    r, r64 = res[0], res64[0]
    la_o, x_o = r64["log_assignment"].float(), r64["ref_descriptors0"][0].float()
    print(f"[{tag}] oracle fp32 vs fp64: la max {(r['log_assignment']-la_o).abs().max():.2e}  x max {(r['ref_descriptors0'][0]-x_o).abs().max():.2e}  |la|max {la_o.abs().max():.1f} |x|max {x_o.abs().max():.2f}")
    for prec in ("fp32", "fp32_simt"):
        model.conf.precision = prec
        out = model.to("cuda")(to_device(data, "cuda"))
        la = out["log_assignment"][0].float().cpu()
        x = out["ref_descriptors0"][0, 0].float().cpu()
        n0, n1 = la_o.shape[0] - 1, la_o.shape[1] - 1
        d = (la[:n0, :n1] - la_o[:n0, :n1]).abs()
        dx = (x[:n0] - x_o).abs()
        rel = d / la_o[:n0, :n1].abs().clamp(min=1.0)
        print(f"[{tag}] {prec:10s} vs fp64 oracle: la max {d.max():.2e} mean {d.mean():.2e} rel max {rel.max():.2e} | x max {dx.max():.2e} mean {dx.mean():.2e} | matches== {(out['matches0'][0].cpu()[:n0]==r64['matches0']).float().mean():.4f}")
        model.cpu()

fx, model, data = load_c1_fixture(ROOT / "tests" / "golden" / "c1_boat.pt")
report("C1 boat sharp", model, fx["conf"], data)
for sharp in (False, True):
    conf = {"filter_threshold": 0.1 if sharp else 0.0}
    model = build_model(conf, 0, sharp_assignment_overrides() if sharp else None)
    data = make_pairs(B=1, n0=2048, n1=2048, seed=51)
    report(f"C2 2048 sharp={sharp}", model, conf, data)
