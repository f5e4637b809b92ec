"""GPU: the residue-number-system FP64 GEMM (csrc/ozaki2.cu: tcgen05.mma.kind::i8 with CTA pairs, TMEM, bulk async
copies, Chinese-remainder reconstruction) against torch float64, for both kernels (cta_group::2 and ::1)."""
import pytest
import torch

from projected_lmc_b200 import ops

from .helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
import os

# both kernels by default; PLMC_TEST_RNS_FLAGS=0 / 1 restricts the run to the CTA-pair / single-CTA kernel
FLAGS = [int(os.environ["PLMC_TEST_RNS_FLAGS"])] if "PLMC_TEST_RNS_FLAGS" in os.environ else [0, ops.GEMM_FLAG_SINGLE_CTA]


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(DEV)


@pytest.mark.parametrize("flags", FLAGS)
def test_integer_inputs_are_recovered_to_the_last_bits(flags):
    """The integer product is exact; what is left is the FP64 rounding of the reconstruction (c'/P, times P)."""
    M, N, K = 256, 256, 256
    g = torch.Generator().manual_seed(0)
    A = torch.randint(-1000, 1001, (M, K), generator=g).double().to(DEV)
    B = torch.randint(-1000, 1001, (N, K), generator=g).double().to(DEV)
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    ops.rns_gemm(0, A, B, C, M, N, K, moduli=16, flags=flags)
    ref = A @ B.T                                                   # exact: |entries| < 2^53
    assert ((C - ref).abs() <= 4 * 2.0 ** -53 * ref.abs()).all()    # a few ulp of EACH entry, not of the largest
    assert torch.equal(C.round(), ref)


@pytest.mark.parametrize("flags", FLAGS)
@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (0.5, -1.0)])
def test_rns_gemm_layouts(layout, alpha, beta, flags):
    M, N, K = 384, 640, 640             # odd multiples of 128: padded 256-tiles on both sides
    a_mc, b_nc = bool(layout & 2), bool(layout & 1)
    A = rnd(K, M, seed=1) if a_mc else rnd(M, K, seed=1)
    B = rnd(K, N, seed=2) if b_nc else rnd(N, K, seed=2)
    C0 = rnd(M, N, seed=3)
    C = C0.clone()
    ref = alpha * (A.T if a_mc else A) @ (B if b_nc else B.T) + beta * C0
    ops.rns_gemm(layout, A, B, C, M, N, K, alpha=alpha, beta=beta, moduli=16, flags=flags)
    assert (C - ref).abs().max().item() < 5e-14 * ref.abs().max().item()


@pytest.mark.parametrize("flags", FLAGS)
def test_submatrix_views_and_leading_dimensions(flags):
    big = rnd(1024, 1536, seed=21)
    A = big[128:128 + 256, 256:256 + 512]      # row-major view, ld = 1536
    Bt = big[512:512 + 512, 128:128 + 384]     # used as B(k, n) n-contiguous (layout bit 0)
    Cbig = rnd(512, 1024, seed=22)
    C = Cbig[128:128 + 256, 256:256 + 384]
    ref = C - A @ Bt
    ops.rns_gemm(1, A, Bt, C, 256, 384, 512, alpha=-1.0, beta=1.0, moduli=16, flags=flags)
    assert (C - ref).abs().max().item() < 5e-14 * ref.abs().max().item()
    assert torch.equal(Cbig[:128], rnd(512, 1024, seed=22)[:128])           # outside the view: untouched


@pytest.mark.parametrize("flags", FLAGS)
def test_rows_with_very_different_scales_and_zero_rows(flags):
    M, N, K = 256, 256, 256
    A, B = rnd(M, K, seed=4), rnd(N, K, seed=5)
    A *= torch.logspace(-150, 150, M, dtype=torch.float64, device=DEV)[:, None]   # per-row exponents matter
    A[7] = 0.0
    B[:, 3] = 0.0
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    ops.rns_gemm(0, A, B, C, M, N, K, moduli=16, flags=flags)
    ref = A @ B.T
    scale = A.abs().max(1).values[:, None] * B.abs().max(1).values[None, :] * K ** 0.5 + 1e-300
    assert ((C - ref).abs() / scale).max().item() < 1e-14
    assert C[7].abs().max().item() == 0.0


@pytest.mark.parametrize("flags", FLAGS)
def test_syrk_lower_same_operand(flags):
    n, K = 640, 1152
    P = rnd(n, K, seed=6)
    C0 = rnd(n, n, seed=7)
    C = C0.clone()
    ops.rns_gemm(0, P, P, C, n, n, K, alpha=-1.0, beta=1.0, lower=True, same_operand=True, moduli=16, flags=flags)
    ref = C0 - P @ P.T
    for ti in range(n // 128):
        for tj in range(n // 128):
            blk = (slice(128 * ti, 128 * ti + 128), slice(128 * tj, 128 * tj + 128))
            if tj > ti:
                assert torch.equal(C[blk], C0[blk])                       # above the diagonal: untouched
            else:
                assert (C[blk] - ref[blk]).abs().max().item() < 5e-14 * ref.abs().max().item()


def test_long_inner_dimension_and_many_tiles():
    M, N, K = 1280, 2304, 20480
    A, B = rnd(M, K, seed=30), rnd(N, K, seed=31)
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    ops.rns_gemm(0, A, B, C, M, N, K, moduli=16)
    ref = A @ B.T
    assert rel_err(C, ref) < 5e-14
    assert ops.rns_bits(16, K) == 55 and ops.rns_bits(16, 16384) == 55 and ops.rns_bits(16, 32768) == 54


@pytest.mark.parametrize("moduli,tol", [(10, 2e-8), (12, 2e-10), (14, 4e-12), (15, 3e-13), (16, 3e-14), (18, 4e-15)])
def test_accuracy_scales_with_moduli(moduli, tol):
    n = 512
    A, B = rnd(n, 1024, seed=8), rnd(n, 1024, seed=9)
    C = torch.empty(n, n, dtype=torch.float64, device=DEV)
    ops.rns_gemm(0, A, B, C, n, n, 1024, moduli=moduli)
    assert rel_err(C, A @ B.T) < tol


def test_scratch_too_small_for_one_pass_splits_the_product():
    M, N, K = 1024, 768, 512
    A, B = rnd(M, K, seed=12), rnd(N, K, seed=13)
    C0 = rnd(M, N, seed=14)
    full = ops.lib().plmc_rns_ws_bytes(M, N, K, 16, 0, 0)
    for frac in (0.6, 0.3, 0.12):
        C = C0.clone()
        ws = torch.empty(int(full * frac), dtype=torch.uint8, device=DEV)
        ops.rns_gemm(0, A, B, C, M, N, K, alpha=2.0, beta=1.0, moduli=16, ws=ws[8:])      # odd alignment too
        assert rel_err(C, C0 + 2.0 * A @ B.T) < 5e-14
    P = rnd(1024, 512, seed=15)
    full = ops.lib().plmc_rns_ws_bytes(1024, 1024, 512, 16, 1, 1)
    C = torch.zeros(1024, 1024, dtype=torch.float64, device=DEV)
    ops.rns_gemm(0, P, P, C, 1024, 1024, 512, lower=True, same_operand=True, moduli=16,
                 ws=torch.empty(int(full * 0.4), dtype=torch.uint8, device=DEV))
    assert rel_err(torch.tril(C), torch.tril(P @ P.T)) < 5e-14


def test_bad_arguments_are_refused():
    A, B = rnd(256, 64, seed=14), rnd(128, 64, seed=15)
    C = torch.zeros(256, 128, dtype=torch.float64, device=DEV)
    with pytest.raises(Exception):
        ops.rns_gemm(0, A, B, C, 256, 128, 64)                    # K not a multiple of 128
    A = rnd(256, 128, seed=16)
    B = rnd(128, 128, seed=17)
    with pytest.raises(Exception):
        ops.rns_gemm(0, A, B, C, 256, 128, 128, lower=True)       # lower needs a square product
    with pytest.raises(Exception):
        ops.rns_gemm(0, A, B, C, 256, 128, 128, moduli=19)


def test_two_streams_with_different_precisions_run_concurrently():
    """Re-entrancy: per-call configuration and scratch, no process-wide state."""
    n, K = 1024, 2048
    A, B = rnd(n, K, seed=40), rnd(n, K, seed=41)
    ref = A @ B.T
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    C1 = torch.empty(n, n, dtype=torch.float64, device=DEV)
    C2 = torch.empty(n, n, dtype=torch.float64, device=DEV)
    ws1 = torch.empty(ops.lib().plmc_rns_ws_bytes(n, n, K, 16, 0, 0), dtype=torch.uint8, device=DEV)
    ws2 = torch.empty(ops.lib().plmc_rns_ws_bytes(n, n, K, 12, 0, 0), dtype=torch.uint8, device=DEV)
    torch.cuda.synchronize()
    for _ in range(4):
        with torch.cuda.stream(s1):
            ops.rns_gemm(0, A, B, C1, n, n, K, moduli=16, ws=ws1)
        with torch.cuda.stream(s2):
            ops.rns_gemm(0, A, B, C2, n, n, K, moduli=12, ws=ws2)
    torch.cuda.synchronize()
    assert rel_err(C1, ref) < 5e-14
    assert 1e-13 < rel_err(C2, ref) < 2e-10          # really the 12-moduli result, not the other stream's setting
