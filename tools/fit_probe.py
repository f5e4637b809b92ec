import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.set_default_dtype(torch.float64)
from projected_lmc_b200 import ProjectedLMCmll, fit
from tests.helpers import make_model, synth
warnings.simplefilter("ignore")
for (n, variant, kernel, kw) in [(120, "PLMC", "matern52", dict(n_iter=12, lr=1e-2, lr_min=1e-3, check_every=5, patience=500)),
                                 (64, "PLMC_fast", "rbf", dict(n_iter=60, lr=1e-6, lr_min=None, loss_thresh=1e-2, patience=7, check_every=4))]:
    X, Y, _, _ = synth(n, 2, 5 if n == 120 else 4, 2, seed=8)
    m = make_model(X, Y, 2, variant=variant, kernel=kernel).cuda()
    out = fit(m, ProjectedLMCmll(m.likelihood, m), X.cuda(), Y.cuda(), **kw)
    print(n, variant, out["cuda_graph"], out["cuda_graph_note"], out["n_iter"], out["stopped_at"])
    try:
        print("randn ok", torch.randn(3, device="cuda").sum().item())
    except Exception as ex:
        print("randn FAILED:", str(ex)[:100])

# a capture that MUST fail (host read inside the step): the fallback has to leave torch's RNG usable
X, Y, _, _ = synth(200, 2, 4, 2, seed=3)
m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
mll = ProjectedLMCmll(m.likelihood, m)
orig = mll.forward
def leaky(*a, **k):
    out = orig(*a, **k)
    float(out.detach().cpu())          # host read: not capturable
    return out
mll.forward = leaky
out = fit(m, mll, m.train_inputs[0], m.train_y, n_iter=8, lr=1e-3, lr_min=None, cuda_graph=True)
print("forced failure:", out["cuda_graph"], (out["cuda_graph_note"] or "")[:60])
print("randn after failed capture:", torch.randn(3, device="cuda").sum().item())
