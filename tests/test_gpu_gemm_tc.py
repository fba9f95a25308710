"""-m gpu: the tcgen05/TMEM/TMA GEMM against the fp32 CUDA-core GEMM (library self-test through the C-ABI)."""
import pytest

from ecog2txt_b200 import Engine, EngineConfig
import parity_common as pc

pytestmark = pytest.mark.gpu

# (M, N, K): recurrent step, input projection, output projection, ragged edges, K tail, tiny N
SHAPES = [(256, 1600, 400), (8704, 1600, 800), (2816, 1806, 800), (300, 200, 100), (129, 72, 36), (128, 32, 8),
          (1000, 100, 3072), (64, 1600, 400)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tcgen05_gemm_matches_simt(gpu_lib, M, N, K):
    eng = pc.engine_for(pc.TINY, gpu_lib, 2, 8, 4)
    d = eng.selftest_gemm(M, N, K)
    # operands in [-0.5, 0.5): |a.b| terms <= 0.25; kind::tf32 keeps 10 mantissa bits of each operand
    # (relative error <= 2^-10 each), errors add like a random walk over K
    tol = 4.0 * (K ** 0.5) * 0.25 * 2 * 2 ** -10 + 1e-5
    print(f"tcgen05 gemm {M}x{N}x{K}: max|diff| = {d:.3e} (tol {tol:.3e})")
    assert d <= tol, (d, tol)
    eng.close()


# weight-gradient shapes C[M,N] = A[K,M]^T B[K,N]: dWx (800x1600 over 8704 rows), dWh, dWp, conv, ragged
TN_SHAPES = [(800, 1600, 8704), (400, 1600, 8448), (1806, 800, 2816), (100, 1600, 8704), (3072, 100, 8704),
             (150, 3200, 2816), (130, 72, 100), (32, 32, 64)]


@pytest.mark.parametrize("M,N,K", TN_SHAPES)
def test_tcgen05_gemm_tn_matches_simt(gpu_lib, M, N, K):
    eng = pc.engine_for(pc.TINY, gpu_lib, 2, 8, 4)
    d = eng.selftest_gemm(-M, N, K)   # negative M selects the TN variant
    tol = 4.0 * (K ** 0.5) * 0.25 * 2 * 2 ** -10 + 1e-5
    print(f"tcgen05 gemm TN {M}x{N}x{K}: max|diff| = {d:.3e} (tol {tol:.3e})")
    assert d <= tol, (d, tol)
    eng.close()
