"""Latent-parallel path on CPU: world_size-2 gloo run of the host-side sharding logic.

The CUDA engine cannot run here, so the two kernels entry points the host code calls are
replaced IN THE TEST by oracle-backed stand-ins (test infrastructure only); what is under
test is the product's sharding, loss bookkeeping and the single flat all-reduce."""
import os
import socket
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedLMCmll, distributed as pdist, ops

from .helpers import KNAMES, cpu_copy, make_model, oracle_params, rel_err, synth


class OracleEngine:
    """Stand-in for LatentEngine with the same contract, backed by the CPU oracle."""

    def log_prob_and_grads(self, X, TY, comps, noise, need_grad, max_tries=None):
        with torch.enable_grad():  # autograd.Function.forward runs with grad mode off
            ty, nz = TY.clone().requires_grad_(), noise.clone().requires_grad_()
            leaves, n, K = [], X.shape[0], 0.0
            for kid, dims, ell, os_ in comps:
                e = ell.clone().requires_grad_()
                leaves.append(e)
                Xg = X[:, list(dims)]
                Kg = O.base_kernel(KNAMES[kid], Xg, Xg, e[:, None, :], zero_diag=False)
                if os_ is not None:
                    o = os_.clone().requires_grad_()
                    leaves.append(o)
                    Kg = Kg * o[:, None, None]
                K = K + Kg
            K = K + torch.diag_embed(nz[:, None].expand(-1, n))
            lp = O.mvn_log_prob(K, ty)
            g = torch.autograd.grad(lp.sum(), [ty, nz] + leaves)
        return lp.detach(), (g[0], g[1], list(g[2:]))


def _patch_kernels():
    ops.project_fwd = lambda Y, T: (Y @ T).T.contiguous()
    ops.project_bwd = lambda Y, G: Y.T @ G.T


def _worker(rank, world, port, variant, q, out_path):
    torch.set_default_dtype(torch.float64)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _patch_kernels()
        X, Y, _, _ = synth(48, 2, 7, q, seed=9)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = make_model(X, Y, q, variant=variant, kernel="matern52")
        ref_model = cpu_copy(m)
        m._engine = OracleEngine()
        pdist.shard_latents(m, rank, world)
        assert m._latent_range == pdist.latent_block(q, rank, world)
        mll = ProjectedLMCmll(m.likelihood, m)
        loss = -mll(m(X), Y)
        loss.backward()
        params = [p for p in m.parameters() if p.requires_grad]
        total = pdist.allreduce_loss_and_grads(loss, params)
        # single-process oracle on the full model
        ref = -O.mll(oracle_params(ref_model), X, Y)
        ref.backward()
        assert abs(total.item() - ref.item()) <= 1e-11 * abs(ref.item()), (total.item(), ref.item())
        refg = dict(ref_model.named_parameters())
        for name, prm in m.named_parameters():
            if refg[name].grad is not None:
                assert rel_err(prm.grad, refg[name].grad) <= 1e-9, name
        if rank == 0:
            with open(out_path, "w") as f:
                f.write("ok")
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("variant,q", [("PLMC", 3), ("PLMC_fast", 2)])
def test_world_size_2_matches_single_process(tmp_path, variant, q):
    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(2, _free_port(), variant, q, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_latent_block_partition():
    for q in range(1, 40):
        for world in range(1, min(q, 8) + 1):
            blocks = [pdist.latent_block(q, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == q
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    assert pdist.latent_block(32, 3, 8) == (12, 16)       # C4: 32 latents, 4 per GPU


def test_cannot_shard_more_ranks_than_latents():
    X, Y, _, _ = synth(20, 2, 4, 2)
    m = make_model(X, Y, 2)
    with pytest.raises(ValueError):
        pdist.shard_latents(m, rank=0, world=3)
