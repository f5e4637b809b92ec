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
a = ap.parse_args()
cfg = bench.WORKLOADS[a.workload]
n = a.points or cfg["n"]
torch.set_default_dtype(torch.float64)
X, Y = bench.make_data(n, cfg["d"], cfg["p"], cfg["q"], seed=0)
model = bench.build_model(X, Y, cfg["q"], cfg["kernel"]).cuda()
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
