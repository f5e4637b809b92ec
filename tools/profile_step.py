"""One reduced-size training iteration + prediction for ncu (launch list / --set full captures of the kernels):
    python tools/profile_step.py [n] [q]"""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from projected_lmc_b200 import ProjectedLMCmll
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.set_default_dtype(torch.float64)
cfg = bench.WORKLOADS["c2"]
X, Y = bench.make_data(n, cfg["d"], cfg["p"], q, seed=0)
model = bench.build_model(X, Y, q, cfg["kernel"]).cuda()
mll = ProjectedLMCmll(model.likelihood, model)
Xd, Yd = model.train_inputs[0], model.train_y
for it in range(2):
    for prm in model.parameters():
        prm.grad = None
    loss = -mll(model(Xd), Yd)
    loss.backward()
torch.cuda.synchronize()
model.eval()
with torch.no_grad(), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    Xs = torch.rand(4096, cfg["d"], dtype=torch.float64, device="cuda") * 2 - 1
    pred = model.full_likelihood()(model(Xs))
torch.cuda.synchronize()
print("loss", float(loss), "pred", float(pred.mean.sum()))
