"""Generate the golden fixtures of tests/golden/ from the reference checkout.

Run in the BUILD container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

1. fragments.pt   -- outputs of the gpytorch-free fragments of the reference executed
                     as they are: init_lmc_coefficients (projected_lmc.py:183-201), the four
                     parametrisation classes (:207-258) and LMCMixingMatrix (:819-890).
2. reference_runs.pt -- the reference's own ProjectedGPModel / ProjectedLMCmll source
                     (projected_lmc.py:893-1241), imported unmodified over oracle/gpytorch_shim
                     (gpytorch itself is not installable here): loss, gradients w.r.t. every raw
                     parameter, predictive mean / variance, for every model variant.
The tests compare the oracle, the host-side modules and (on the GPU) the CUDA path with them.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/projectedlmc/projected_lmc.py"

torch.set_default_dtype(torch.float64)


def run_fragments():
    src = open(REF).read().splitlines()
    pieces = src[182:201] + [""] + src[206:258] + [""] + src[818:890]
    from typing import List, Union  # noqa: F401

    from sklearn.utils.extmath import randomized_svd  # noqa: F401
    from torch import Tensor  # noqa: F401

    ns = dict(torch=torch, np=np, Tensor=Tensor, List=List, Union=Union, randomized_svd=randomized_svd)
    exec("\n".join(pieces), ns)
    out = {}
    g = torch.Generator().manual_seed(123)
    # init_lmc_coefficients
    Y = torch.randn(40, 6, generator=g)
    out["init.Y"] = Y
    U, S = ns["init_lmc_coefficients"](Y, 3, QR_form=True)
    out["init.q3.U"], out["init.q3.S"] = U, S
    U, S = ns["init_lmc_coefficients"](Y, 6, QR_form=True)
    out["init.q6.U"], out["init.q6.S"] = U, S
    out["init.q3.plain"] = ns["init_lmc_coefficients"](Y, 3, QR_form=False)
    Ysmall = torch.randn(2, 6, generator=g)
    out["init.small.Y"] = Ysmall
    U, S = ns["init_lmc_coefficients"](Ysmall, 4, QR_form=True)
    out["init.small.U"], out["init.small.S"] = U, S
    # parametrisations
    A = torch.randn(5, 5, generator=g)
    v = torch.randn(7, generator=g)
    out["param.A"], out["param.v"] = A, v
    out["param.scalar.fwd"] = ns["ScalarParam"](bounds=(-9.0, 9.0)).forward(v.clone())
    out["param.scalar.fwd_clamped"] = ns["ScalarParam"](bounds=(-0.01, 0.01)).forward(v.clone() + 5)
    out["param.posdiag.fwd"] = ns["PositiveDiagonalParam"]().forward(A.clone())
    out["param.posdiag.inv"] = ns["PositiveDiagonalParam"]().right_inverse(torch.diag_embed(torch.rand(5, generator=g) + 0.5))
    out["param.upper.fwd"] = ns["UpperTriangularParam"]().forward(A.clone())
    Apos = A.clone().triu()
    Apos[range(5), range(5)] = torch.rand(5, generator=g) + 0.5
    out["param.upper.inv_in"] = Apos.clone()
    out["param.upper.inv"] = ns["UpperTriangularParam"]().right_inverse(Apos.clone())
    out["param.lower.fwd"] = ns["LowerTriangularParam"](bounds=(-9.0, 9.0)).forward(A.clone() * 6)
    out["param.lower.inv"] = ns["LowerTriangularParam"]().right_inverse(Apos.T.clone())
    # LMCMixingMatrix
    Qp, _ = torch.linalg.qr(torch.randn(6, 6, generator=g))
    R = torch.diag_embed(torch.rand(3, generator=g) + 0.5)
    out["mix.Qp"], out["mix.R"] = Qp, R
    for tag, Q_in in (("plus", Qp), ("q", Qp[:, :3].contiguous())):
        for bulk in (True, False):
            mm = ns["LMCMixingMatrix"](Q_in.clone(), R.clone(), bulk=bulk)
            Q, Rr, Qo = mm.QR()
            key = f"mix.{tag}.{'bulk' if bulk else 'split'}"
            out[key + ".mode"] = mm.mode
            out[key + ".Q"], out[key + ".R"] = Q.detach(), Rr.detach()
            out[key + ".Qo"] = None if Qo is None else Qo.detach()
            out[key + ".fwd"] = mm().detach()
            if bulk:
                out[key + ".H"] = mm.H.detach()
    torch.save(out, os.path.join(HERE, "fragments.pt"))
    print("fragments.pt:", len(out), "entries")


VARIANTS = {
    "PLMC": dict(BDN=False, diagonal_B=False, scalar_B=False),
    "PLMC_fast": dict(BDN=True, diagonal_B=True, scalar_B=True),
    "BDN_full": dict(BDN=True, diagonal_B=False, scalar_B=False),
    "BDN_diag": dict(BDN=True, diagonal_B=True, scalar_B=False),
    "M_diag": dict(BDN=False, diagonal_B=True, scalar_B=False),
    "M_scalar": dict(BDN=False, diagonal_B=True, scalar_B=True),
    "oilmm": dict(BDN=True, diagonal_B=True, scalar_B=True, diagonal_R=True, bulk=False),
    "nonbulk_tri": dict(BDN=False, diagonal_B=False, scalar_B=False, bulk=False),
}


def run_reference():
    from oracle import gpytorch_shim
    from tests.helpers import synth

    ref = gpytorch_shim.load_reference(REF)
    gp = sys.modules["gpytorch"]
    out = {"meta": {"n": 70, "d": 3, "p": 6, "q": 2, "ns": 9, "seed": 5}}
    X, Y, Xs, _ = synth(70, 3, 6, 2, seed=5, ns=9)
    out["X"], out["Y"], out["Xs"] = X, Y, Xs
    for vname, vkw in VARIANTS.items():
        for kname, ktype, kk in (("rbf", gp.kernels.RBFKernel, {}), ("matern52", gp.kernels.MaternKernel, {})):
            for outputscales in (False, True):
                if outputscales and not (vname == "PLMC" and kname == "matern52"):
                    continue
                torch.manual_seed(11)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    m = ref.ProjectedGPModel(X, Y, 6, 2, proj_likelihood=None, mean_type=gp.means.ZeroMean,
                                             kernel_type=ktype, ker_kwargs=kk, init_lmc_coeffs=True,
                                             outputscales=outputscales, **vkw)
                gen = torch.Generator().manual_seed(12)
                with torch.no_grad():
                    for name, prm in m.named_parameters():
                        if "Q_plus" in name:
                            continue
                        prm.add_(0.1 * torch.randn(prm.shape, generator=gen))
                state0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
                m.train()
                mll = ref.ProjectedLMCmll(m.likelihood, m)
                with gp.settings.cholesky_max_tries(8):
                    loss = -mll(m(X), Y)
                loss.backward()
                grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}
                terms = [float(t) for t in mll.proj_term_list]
                m.eval()
                with torch.no_grad(), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    fl = m.full_likelihood()
                    lat = m(Xs)
                    obs = fl(lat)
                    entry = dict(
                        state=state0, loss=float(loss), grads=grads, proj_terms=terms,
                        mean=lat.mean.clone(), var_f=lat.variance.clone(), var_y=obs.variance.clone(),
                        T=m.projection_matrix().clone(), TY=m.project_data(Y).clone(),
                        Sigma_factor=fl.task_noise_covar_factor.detach().clone(),
                        lscales=m.lscales().clone(), B_tilde=m.B_tilde().clone(),
                        kwargs=dict(vkw), kernel=kname, outputscales=outputscales,
                    )
                out[f"{vname}/{kname}/{'os' if outputscales else 'noos'}"] = entry
                print(vname, kname, outputscales, "loss", float(loss))
    torch.save(out, os.path.join(HERE, "reference_runs.pt"))
    print("reference_runs.pt:", len(out) - 4, "cases")


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("reference checkout not found: fixtures can only be regenerated in the build container")
    run_fragments()
    run_reference()
