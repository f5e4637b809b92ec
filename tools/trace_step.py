"""Per-GEMM-shape timing of one training iteration (CUDA events around every GEMM of the factorisation):
    python tools/trace_step.py [--workload c2] [--points N]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from projected_lmc_b200 import ProjectedLMCmll, ops

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--points", type=int, default=0)
ap.add_argument("--predict", type=int, default=0, help="trace the prediction of this many test points instead")
a = ap.parse_args()
cfg = bench.WORKLOADS[a.workload]
n = a.points or cfg["n"]
torch.set_default_dtype(torch.float64)
p_, q_ = (cfg["named_p"], cfg["named_q"]) if a.predict else (cfg["p"], cfg["q"])
X, Y = bench.make_data(n, cfg["d"], p_, q_, seed=0)
model = bench.build_model(X, Y, q_, cfg["kernel"]).cuda()
if a.predict:
    import warnings
    warnings.simplefilter("ignore")
    model.eval()
    if "PLMC_PREDICT_INVERSE" not in os.environ:
        model._engine.predict_inverse = True
    Xs = (torch.rand(a.predict, cfg["d"]) * 2 - 1).cuda()
    with torch.no_grad():
        model(Xs[:128]); model(Xs); torch.cuda.synchronize()
        ops.trace_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); model._engine.profile = []; model(Xs); e1.record(); torch.cuda.synchronize()
    print("predict ms", e0.elapsed_time(e1), file=sys.stderr)
    ops.trace_report()
    sys.exit(0)
model.train()
mll = ProjectedLMCmll(model.likelihood, model)
Xd, Yd = model.train_inputs[0], model.train_y


def step():
    for prm in model.parameters():
        prm.grad = None
    loss = -mll(model(Xd), Yd)
    loss.backward()
    return loss


step(); torch.cuda.synchronize()
ops.trace_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1), file=sys.stderr)
ops.trace_report()
