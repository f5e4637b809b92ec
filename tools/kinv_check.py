"""Gradient deviation at full C2 size between arithmetic modes of the K^-1 (LAUUM) step:
pure DMMA vs INT8 path with 7 slices everywhere vs 7 slices + 6 in LAUUM (the default)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from projected_lmc_b200 import ProjectedLMCmll
from projected_lmc_b200.engine import LatentEngine

cfg = bench.WORKLOADS["c2"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n"]
torch.set_default_dtype(torch.float64)
X, Y = bench.make_data(n, cfg["d"], cfg["p"], cfg["q"], seed=0)
out = {}
for name, (s, sk) in {"dmma": (0, 0), "int8 7/7": (7, 7), "int8 7/6": (7, 6), "int8 7/5": (7, 5)}.items():
    LatentEngine.fp64_slices, LatentEngine.fp64_slices_kinv = s, sk
    m = bench.build_model(X.clone(), Y.clone(), cfg["q"], cfg["kernel"]).cuda()
    m.train()
    loss = -ProjectedLMCmll(m.likelihood, m)(m(m.train_inputs[0]), m.train_y)
    loss.backward()
    out[name] = (loss.item(), torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]).cpu())
    m._engine.release()
    del m, loss
    import gc
    gc.collect()
    torch.cuda.empty_cache()
ref_l, ref_g = out["dmma"]
for name, (l, g) in out.items():
    print(f"{name:9s} loss {l!r}  |dloss|/|loss| {abs(l - ref_l) / abs(ref_l):.2e}  "
          f"grad rel err (2-norm) {((g - ref_g).norm() / ref_g.norm()).item():.2e}  "
          f"max elementwise rel {((g - ref_g).abs() / ref_g.abs().clamp_min(1e-300)).max().item():.2e}")
