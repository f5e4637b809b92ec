"""How many moduli does the explicit inverse need at scale?  Gradients of an ILL-conditioned problem (RBF, long
lengthscales, noise at its e^-9 floor) at n = 20000 with the inverse on 16 / 13 / 12 / 11 moduli, against the pure
FP64 (DMMA) path and against 16 moduli.   python tools/kinv_scale_check.py [n] [ell]"""
import gc, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from projected_lmc_b200 import ProjectedLMCmll
from projected_lmc_b200.engine import LatentEngine
from tests.helpers import cpu_copy, rel_err
from tests.test_conditioning_gpu import ill_conditioned_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ell = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
m, X, Y, _ = ill_conditioned_model(n, 3, 5, 2, ell, seed=n)
Xg, Yg = X.cuda(), Y.cuda()
out = {}
for name, mode, prec, kinv in [("fp64", "fp64", 0, 0), ("rns 16/16", "rns", 16, 16), ("rns 16/13", "rns", 16, 13),
                               ("rns 16/12", "rns", 16, 12), ("rns 16/11", "rns", 16, 11), ("rns 15/12", "rns", 15, 12)]:
    LatentEngine.gemm_mode = mode
    if mode == "rns":
        LatentEngine.rns_moduli, LatentEngine.rns_moduli_kinv = prec, kinv
    mg = cpu_copy(m)
    mg._engine = LatentEngine()
    mg = mg.cuda()
    loss = -ProjectedLMCmll(mg.likelihood, mg)(mg(Xg), Yg)
    loss.backward()
    out[name] = (loss.item(), {k: p.grad.detach().clone() for k, p in mg.named_parameters() if p.grad is not None})
    mg._engine.release()
    del mg, loss
    gc.collect()
    torch.cuda.empty_cache()
for name, (l, g) in out.items():
    row = {"n": n, "ell": ell, "mode": name, "loss": l}
    for ref in ("fp64", "rns 16/16"):
        rl, rg = out[ref]
        row[f"loss_vs_{ref}"] = abs(l - rl) / abs(rl)
        row[f"max_grad_rel_vs_{ref}"] = max(rel_err(g[k], rg[k]) for k in g)
    print(json.dumps(row), flush=True)
