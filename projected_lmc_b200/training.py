"""Training-loop helper: the loop of the reference drivers (experiments.py:259-284,
realdata_experiments.py:174-210) around the CUDA path -- AdamW, exponential learning-rate decay,
plateau-based early stop -- without the per-iteration host synchronisation.

The reference calls ``loss.item()`` every iteration to evaluate its stopping rule
(experiments.py:275); here losses stay on the device and the same rule is evaluated over a
buffered history every ``check_every`` iterations.  Inside an iteration the only host wait is on
the event recorded right after the Cholesky factorisation (its per-latent status decides on the
psd_safe_cholesky jitter retry); it is taken AFTER the solves, the inverse and the gradient
sweep of the same iteration have been queued behind it, so the GPU queue does not drain on it
(engine.LatentEngine.log_prob_and_grads).
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import gp


def fit(model, mll, X, Y, n_iter: int, lr: float = 1e-2, lr_min: Optional[float] = 1e-3, loss_thresh: float = 1e-4,
        patience: int = 500, check_every: int = 25, cholesky_max_tries: int = 8, optimizer=None, scheduler=None,
        print_loss: bool = False, freq_print: int = 100, sync_grads=None):
    """Train ``model`` by maximising ``mll``; returns a dict with the loss history and timings.

    Stopping rule of the reference: ``|1 - loss_i / loss_{i-1}| < loss_thresh`` for more than
    ``patience`` consecutive iterations.  ``sync_grads(loss, params)`` is called after backward in
    latent-parallel runs (``distributed.allreduce_loss_and_grads``)."""
    if optimizer is None:
        optimizer = torch.optim.AdamW(model.parameters(), lr=lr)
    if scheduler is None and lr_min is not None:
        scheduler = torch.optim.lr_scheduler.ExponentialLR(optimizer, gamma=float(np.exp(np.log(lr_min / lr) / n_iter)))
    params = [p for p in model.parameters() if p.requires_grad]
    model.train()
    losses = torch.empty(n_iter, dtype=torch.float64, device=X.device)
    plateau_id, last, checked, stop_at = 0, None, 0, None
    start = time.time()
    it = 0
    with gp.settings.cholesky_max_tries(cholesky_max_tries):
        for it in range(n_iter):
            optimizer.zero_grad(set_to_none=True)
            loss = -mll(model(X), Y)
            loss.backward()
            if sync_grads is not None:
                loss = sync_grads(loss, params)
            optimizer.step()
            if scheduler is not None:
                scheduler.step()
            losses[it] = loss.detach()
            if (it + 1) % check_every == 0 or it == n_iter - 1:
                hist = losses[checked:it + 1].tolist()        # one host sync per check_every iterations
                for k, new in enumerate(hist, start=checked):
                    if print_loss and k % freq_print == 0:
                        print(new)
                    if k > 0 and abs(1 - new / last) < loss_thresh:
                        plateau_id += 1
                        if plateau_id > patience and stop_at is None:
                            stop_at = k
                    else:
                        plateau_id = 0
                    last = new
                checked = it + 1
                if stop_at is not None:
                    break
    if X.is_cuda:
        torch.cuda.synchronize()
    n_done = it + 1
    return {"losses": losses[:n_done].cpu(), "n_iter": n_done, "stopped_at": stop_at,
            "train_time": time.time() - start}
