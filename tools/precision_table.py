"""Errors of the CUDA path against the CPU oracle as a function of the INT8 precision settings, on well- and
ill-conditioned problems (default routing thresholds, so the INT8 kernels run).  Prints one JSON line per case.
    python tools/precision_table.py"""
import json, math, os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedLMCmll
from projected_lmc_b200.engine import LatentEngine
from tests.helpers import cpu_copy, make_model, oracle_params, rel_err, synth
from tests.test_conditioning_gpu import ill_conditioned_model

CASES = [("well n=3000 matern52 default ell", None, 3000, 0.0), ("ill n=2048 rbf ell=3 noise=e^-9", "ill", 2048, 3.0),
         ("ill n=3000 rbf ell=5 noise=e^-9", "ill", 3000, 5.0)]
SETTINGS = [("fp64", 0, 0), ("digits", 7, 6), ("rns", 16, 16), ("rns", 16, 14), ("rns", 16, 13), ("rns", 16, 12), ("rns", 15, 13)]
for name, kind, n, ell in CASES:
    if kind == "ill":
        m, X, Y, Xs = ill_conditioned_model(n, 3, 5, 2, ell, seed=n)
    else:
        X, Y, Xs, _ = synth(n, 6, 5, 2, seed=7, ns=64)
        m = make_model(X, Y, 2, variant="PLMC", kernel="matern52")
    mc = cpu_copy(m)
    ref = -O.mll(oracle_params(mc), X, Y)
    ref.backward()
    refg = {k: v.grad.clone() for k, v in mc.named_parameters() if v.grad is not None}
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean_ref, _, var_ref = O.predict(oracle_params(mc), X, Y, Xs)
    for mode, prec, kinv in SETTINGS:
        LatentEngine.gemm_mode = mode
        LatentEngine.rns_min_k, LatentEngine.rns_min_mnk = 0, 0       # residues for every routed product
        if mode == "rns":
            LatentEngine.rns_moduli, LatentEngine.rns_moduli_kinv = prec, kinv
        mg = cpu_copy(m)
        mg._engine = LatentEngine()
        mg = mg.cuda()
        for p in mg.parameters():
            p.grad = None
        loss = -ProjectedLMCmll(mg.likelihood, mg)(mg(X.cuda()), Y.cuda())
        loss.backward()
        gerr = max(rel_err(p.grad, refg[k]) for k, p in mg.named_parameters() if k in refg)
        mg.eval()
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pred = mg.full_likelihood()(mg(Xs.cuda()))
            cond = float(mg.kernel_cond().max())
        print(json.dumps({"case": name, "cond": cond, "mode": mode, "precision": prec, "kinv": kinv,
                          "mll_rel_err": abs(loss.item() - ref.item()) / abs(ref.item()), "max_grad_rel_err": gerr,
                          "pred_mean_rel_err": rel_err(pred.mean, mean_ref),
                          "pred_var_rel_err": rel_err(pred.variance, var_ref)}), flush=True)
        del mg
