"""GPU: the FP64-via-INT8 tensor path (csrc/ozaki.cu, tcgen05.mma.kind::i8 + TMEM + bulk async copies).

* the stand-alone GEMM against torch float64 for every operand layout, SYRK mode, K-chunking;
* exactness of the slicing on inputs that fit one plane;
* the whole model path with the emulation forced onto every GEMM (min_dim = 128) against the CPU oracle at
  the same tolerances as the pure-DMMA path, and emulated-vs-DMMA agreement at a size beyond the oracle."""
import pytest
import torch

from oracle import plmc_oracle as O
from projected_lmc_b200 import ProjectedLMCmll, ops
from projected_lmc_b200.engine import LatentEngine

from .helpers import cpu_copy, make_model, oracle_params, rel_err, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(DEV)


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (0.5, -1.0)])
def test_ozaki_gemm_layouts(layout, alpha, beta):
    M, N, K = 384, 256, 640
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = rnd(K, M, seed=1) if a_mc else rnd(M, K, seed=1)
    B = rnd(K, N, seed=2) if b_nc else rnd(N, K, seed=2)
    C0 = rnd(M, N, seed=3)
    C = C0.clone()
    ref = alpha * (A.T if a_mc else A) @ (B if b_nc else B.T) + beta * C0
    ops.ozaki_gemm(layout, A, B, C, M, N, K, alpha=alpha, beta=beta, slices=7)
    assert (C - ref).abs().max().item() < 5e-14 * ref.abs().max().item()


def test_single_plane_inputs_are_exact():
    M, N, K = 256, 128, 512
    g = torch.Generator().manual_seed(0)
    # |x| * 2^-(e+1) < 1/2 and 7 bits in the leading plane: integers / 64 up to 63 fit one plane exactly
    A = (torch.randint(-63, 64, (M, K), generator=g).double() / 64.0).to(DEV)
    B = (torch.randint(-63, 64, (N, K), generator=g).double() / 64.0).to(DEV)
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    ops.ozaki_gemm(0, A, B, C, M, N, K, slices=1)
    assert torch.equal(C, A @ B.T)


def test_scratch_of_exactly_the_advertised_size_at_any_alignment():
    M, N, K = 256, 128, 256
    A, B = rnd(M, K, seed=12), rnd(N, K, seed=13)
    need = ops.lib().plmc_ozaki_ws_bytes(M, N, K, 7, 0)
    buf = torch.empty(need + 2048, dtype=torch.uint8, device=DEV)
    for off in (0, 8, 520, 1016):
        C = torch.empty(M, N, dtype=torch.float64, device=DEV)
        ops.ozaki_gemm(0, A, B, C, M, N, K, slices=7, ws=buf[off:off + need])
        assert rel_err(C, A @ B.T) < 1e-13


def test_lower_needs_a_square_product():
    A, B = rnd(256, 64, seed=14), rnd(128, 64, seed=15)
    C = torch.zeros(256, 128, dtype=torch.float64, device=DEV)
    with pytest.raises(Exception):
        ops.ozaki_gemm(0, A, B, C, 256, 128, 64, lower=True)


def test_gemm_trace_reports_every_shape(capfd):
    X, Y, _, _ = synth(1500, 3, 4, 2, seed=5)
    m = make_model(X, Y, 2, variant="PLMC", kernel="rbf").cuda()
    ops.trace_enable(True)
    try:
        loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
        loss.backward()
    finally:
        ops.trace_report()
        ops.trace_enable(False)
    err = capfd.readouterr().err
    assert "TFLOP/s" in err and ("int8" in err or "dmma" in err)


def test_rows_with_very_different_scales_and_zero_rows():
    M, N, K = 256, 256, 256
    A, B = rnd(M, K, seed=4), rnd(N, K, seed=5)
    A *= torch.logspace(-150, 150, M, dtype=torch.float64, device=DEV)[:, None]   # per-row exponents matter
    A[7] = 0.0
    B[:, 3] = 0.0
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    ops.ozaki_gemm(0, A, B, C, M, N, K, slices=7)
    ref = A @ B.T
    scale = A.abs().max(1).values[:, None] * B.abs().max(1).values[None, :] * K ** 0.5 + 1e-300
    assert ((C - ref).abs() / scale).max().item() < 1e-14
    assert C[7].abs().max().item() == 0.0


def test_syrk_lower_same_operand_and_k_chunks():
    n, K = 384, 20480                      # K > 16384: two INT32 accumulation chunks
    P = rnd(n, K, seed=6)
    C0 = rnd(n, n, seed=7)
    C = C0.clone()
    ops.ozaki_gemm(0, P, P, C, n, n, K, alpha=-1.0, beta=1.0, lower=True, same_operand=True)
    ref = C0 - P @ P.T
    for ti in range(n // 128):
        for tj in range(n // 64):
            blk = (slice(128 * ti, 128 * ti + 128), slice(64 * tj, 64 * tj + 64))
            if 64 * tj > 128 * ti + 127:
                assert torch.equal(C[blk], C0[blk])                       # above the diagonal: untouched
            else:
                assert (C[blk] - ref[blk]).abs().max().item() < 5e-14 * ref.abs().max().item()


@pytest.mark.parametrize("slices,tol", [(4, 2e-7), (5, 1e-9), (6, 5e-12), (7, 3e-14)])   # 8s - 1 bits
def test_accuracy_scales_with_slices(slices, tol):
    n = 512
    A, B = rnd(n, 1024, seed=8), rnd(n, 1024, seed=9)
    C = torch.empty(n, n, dtype=torch.float64, device=DEV)
    ops.ozaki_gemm(0, A, B, C, n, n, 1024, slices=slices)
    ref = A @ B.T
    assert rel_err(C, ref) < tol


@pytest.fixture(params=["digits", "rns"])
def force_emulation_everywhere(request):
    E = LatentEngine
    old = (E.fp64_slices, E.fp64_min_dim, E.gemm_mode, E.fp64_min_mnk, E.fp64_min_order, E.rns_min_k, E.rns_min_mnk)
    E.fp64_slices, E.fp64_min_dim, E.gemm_mode, E.fp64_min_mnk, E.fp64_min_order = 7, 128, request.param, 0, 256
    E.rns_min_k, E.rns_min_mnk = 0, 0            # "rns" really means residues for every product
    yield
    (E.fp64_slices, E.fp64_min_dim, E.gemm_mode, E.fp64_min_mnk, E.fp64_min_order, E.rns_min_k, E.rns_min_mnk) = old


@pytest.mark.parametrize("variant,kernel", [("PLMC", "matern52"), ("PLMC_fast", "rbf")])
def test_model_path_with_every_gemm_emulated_matches_oracle(force_emulation_everywhere, variant, kernel):
    X, Y, Xs, _ = synth(700, 4, 6, 3, seed=3, ns=40)
    m = make_model(X, Y, 3, variant=variant, kernel=kernel)
    mc = cpu_copy(m)
    m = m.cuda()
    loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
    loss.backward()
    ref = -O.mll(oracle_params(mc), X, Y)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-9 * abs(ref.item())
    refg = dict(mc.named_parameters())
    for name, prm in m.named_parameters():
        if refg[name].grad is not None:
            assert rel_err(prm.grad, refg[name].grad) <= 1e-7, name
    m.eval()
    with torch.no_grad():
        pred = m(Xs.cuda())
        mean_ref, var_ref, _ = O.predict(oracle_params(mc), X, Y, Xs)
    assert rel_err(pred.mean, mean_ref) <= 1e-7 and rel_err(pred.variance, var_ref) <= 1e-7


def test_emulated_and_dmma_paths_agree_at_n_6000():
    """Beyond the oracle's reach: same model, all three arithmetic paths (default thresholds)."""
    X, Y, _, _ = synth(6000, 8, 6, 2, seed=11)
    out = {}
    old = LatentEngine.gemm_mode
    try:
        for mode in ("fp64", "digits", "rns"):
            LatentEngine.gemm_mode = mode
            m = make_model(X, Y, 2, variant="PLMC", kernel="matern52").cuda()
            loss = -ProjectedLMCmll(m.likelihood, m)(m(X.cuda()), Y.cuda())
            loss.backward()
            out[mode] = (loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()})
            del m
    finally:
        LatentEngine.gemm_mode = old
    for mode in ("digits", "rns"):
        assert abs(out["fp64"][0] - out[mode][0]) <= 1e-10 * abs(out["fp64"][0]), mode
        for k in out["fp64"][1]:
            assert rel_err(out[mode][1][k], out["fp64"][1][k]) <= 1e-7, (mode, k)
