"""Which piece of the training step cannot be captured into a CUDA graph on this torch build?  Captures the pieces one
by one (forward + backward each) and prints the first error of each.   python tools/capture_probe.py"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
from projected_lmc_b200 import ProjectedLMCmll  # noqa: E402
from projected_lmc_b200.mll import projection_terms  # noqa: E402
from tests.helpers import make_model, synth  # noqa: E402

warnings.simplefilter("ignore")
variant = sys.argv[1] if len(sys.argv) > 1 else "PLMC"
X, Y, _, _ = synth(300, 3, 6, 2, seed=1)
m = make_model(X, Y, 2, variant=variant, kernel="matern52").cuda()
mll = ProjectedLMCmll(m.likelihood, m)
Xd, Yd = m.train_inputs[0], m.train_y
params = [p for p in m.parameters() if p.requires_grad]


def eager_step():
    for p in params:
        p.grad = None
    loss = -mll(m(Xd), Yd)
    loss.backward()


for _ in range(3):
    eager_step()
mll.proj_term_list = None
torch.cuda.synchronize()


def attempt(name, fn):
    for p in params:
        p.grad = None
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        print(f"[ok]   {name}")
    except Exception as ex:  # noqa: BLE001
        print(f"[FAIL] {name}: {str(ex).splitlines()[0][:160]}")
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass


H = m.lmc_coefficients.H if hasattr(m.lmc_coefficients, "H") else None
S = mll._second_moment(Yd)
m._engine.capture_mode = True
attempt("qr forward", lambda: torch.linalg.qr(H if H is not None else torch.randn(6, 6, device="cuda")))
if H is not None:
    attempt("qr fwd+bwd", lambda: sum(t.sum() for t in torch.linalg.qr(H)).backward())
attempt("lmc_coefficients.QR fwd+bwd", lambda: sum(t.sum() for t in m.lmc_coefficients.QR()).backward())
attempt("projection_matrix fwd+bwd", lambda: m.projection_matrix().sum().backward())
attempt("projected_noise fwd+bwd", lambda: m.projected_noise().sum().backward())
attempt("project_data fwd+bwd", lambda: m.project_data(Yd).sum().backward())
attempt("projection_terms fwd+bwd", lambda: sum(projection_terms(m, S, 300)).backward())
attempt("B_tilde fwd+bwd", lambda: m.B_tilde().sum().backward() if hasattr(m, "B_tilde") else None)
attempt("mll forward only", lambda: mll(m(Xd), Yd))
attempt("mll fwd+bwd", lambda: (-mll(m(Xd), Yd)).backward())
m._engine.capture_mode = False
