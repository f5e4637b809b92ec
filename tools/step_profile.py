"""Where a launch-bound training step (BASELINE config 1: n=1000, p=50, q=10) spends its time.

torch.profiler over a few steps of the bench's own step(): host time per op (self CPU), device time per kernel,
and a wall-clock split of the step into its sections (forward host algebra / engine / backward / optimiser), each
bracketed by a device synchronise so the figures add up.   python tools/step_profile.py [--workload c1]
"""
import argparse
import math
import os
import sys
import time
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rows", type=int, default=45)
    args = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    from projected_lmc_b200 import ProjectedLMCmll

    cfg = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    Xh, Yh = bench.make_data(cfg["n"], cfg["d"], cfg["p"], cfg["q"], seed=0)
    model = bench.build_model(Xh.clone(), Yh.clone(), cfg["q"], cfg["kernel"]).to(dev)
    model.train()
    mll = ProjectedLMCmll(model.likelihood, model)
    Xd, Yd = model.train_inputs[0], model.train_y
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-2)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=math.exp(math.log(1e-3 / 1e-2) / 10000))

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    sect = {"zero_grad": 0.0, "model(X)": 0.0, "mll": 0.0, "backward": 0.0, "opt+sched": 0.0}

    def step(timed=False):
        t0 = sync() if timed else 0
        opt.zero_grad(set_to_none=True)
        t1 = sync() if timed else 0
        out = model(Xd)
        t2 = sync() if timed else 0
        loss = -mll(out, Yd)
        t3 = sync() if timed else 0
        loss.backward()
        t4 = sync() if timed else 0
        opt.step()
        sched.step()
        t5 = sync() if timed else 0
        if timed:
            for k, v in zip(sect, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                sect[k] += v
        return loss

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(5):
            step()
        t0 = sync()
        for _ in range(args.steps):
            step()
        t1 = sync()
        print("free-running step: %.3f ms" % ((t1 - t0) / args.steps * 1e3))
        for _ in range(args.steps):
            step(timed=True)
        print("sections (each synchronised), ms per step:",
              {k: round(v / args.steps * 1e3, 3) for k, v in sect.items()})
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize()
    ka = prof.key_averages()
    print(ka.table(sort_by="self_cpu_time_total", row_limit=args.rows, max_name_column_width=60))
    print(ka.table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
