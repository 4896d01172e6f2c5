"""GPU parity: K5 Dense-layer GEMM + fused epilogues, column sums, activation
backward, BCE loss — through the C-ABI vs numpy fp64."""
import numpy as np
import pytest

from util import REL_BF16, REL_F32, assert_close

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(dtype) if dtype is not None else t


def _ref_epi(acc, epi, bias, aux, cold):
    from recommendsystem_b200 import cabi
    if epi == cabi.EPI_BIAS:
        return acc + bias
    if epi == cabi.EPI_BIAS_RELU:
        return np.maximum(acc + bias, 0)
    if epi == cabi.EPI_BIAS_SIGMOID:
        return 1 / (1 + np.exp(-(acc + bias)))
    if epi == cabi.EPI_MUL_RELU_MASK:
        return acc * (aux > 0)
    if epi == cabi.EPI_MUL_DSIGMOID:
        return acc * aux * (1 - aux)
    if epi == cabi.EPI_ACCUM:
        return acc + cold
    return acc


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (64, 64, 16), (100, 37, 53), (1024, 256, 624), (333, 1, 752),
                                   (257, 400, 130)])
@pytest.mark.parametrize("tA,tB", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("epi", range(7))
def test_gemm_f32(cuda_dev, M, N, K, tA, tB, epi):
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N + K + epi)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K) if tB else (K, N)) / np.sqrt(K)).astype(np.float32)   # Glorot-like scale
    bias = rng.standard_normal(N).astype(np.float32)
    aux = rng.random((M, N)).astype(np.float32) - 0.3
    cold = rng.standard_normal((M, N)).astype(np.float32)
    acc = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    ref = _ref_epi(acc, epi, bias.astype(np.float64), aux.astype(np.float64), cold.astype(np.float64))
    C = _t(cold, cuda_dev)
    ops.gemm(_t(A, cuda_dev), _t(B, cuda_dev), C, bias=_t(bias, cuda_dev), aux=_t(aux, cuda_dev), epilogue=epi,
             transA=tA, transB=tB)
    assert_close(C.cpu().numpy(), ref, REL_F32, "gemm f32")


def test_gemm_strided_views(cuda_dev):
    """Operands and outputs that are column slices of wider buffers (concat-free towers)."""
    from recommendsystem_b200 import cabi, ops
    rng = np.random.default_rng(0)
    big = _t(rng.standard_normal((50, 100)).astype(np.float32), cuda_dev)
    Wm = _t(rng.standard_normal((30, 20)).astype(np.float32), cuda_dev)
    out = torch.zeros(50, 64, device=cuda_dev)
    ops.gemm(big[:, 10:40], Wm, out[:, 8:28])
    ref = big[:, 10:40].double().cpu().numpy() @ Wm.double().cpu().numpy()
    assert_close(out[:, 8:28].cpu().numpy(), ref, REL_F32, "strided gemm")
    assert float(out[:, :8].abs().sum()) == 0 and float(out[:, 28:].abs().sum()) == 0


@pytest.mark.parametrize("M,N", [(1, 1), (1000, 37), (8192, 256), (300, 1)])
def test_colsum_act_bwd(cuda_dev, M, N):
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N)
    x = rng.standard_normal((M, N)).astype(np.float32)
    assert_close(ops.colsum(_t(x, cuda_dev)).cpu().numpy(), x.astype(np.float64).sum(0), REL_F32, "colsum")
    ref = rng.standard_normal((M, N)).astype(np.float32)
    y0 = ops.act_bwd(_t(x, cuda_dev), _t(ref, cuda_dev), 0).cpu().numpy()
    assert np.array_equal(y0, x * (ref > 0))
    sg = 1 / (1 + np.exp(-ref))
    y1 = ops.act_bwd(_t(x, cuda_dev), _t(sg, cuda_dev), 1).cpu().numpy()
    assert_close(y1, x.astype(np.float64) * sg * (1 - sg), REL_F32, "dsigmoid")


@pytest.mark.parametrize("B,k", [(1, 1), (1024, 1), (4096, 7)])
def test_bce(cuda_dev, B, k):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B + k)
    p = rng.random((B, k)).astype(np.float32)
    p.reshape(-1)[:3] = [0.0, 1.0, 5e-7][: min(3, p.size)]
    y = (rng.random((B, k)) < 0.25).astype(np.float32)
    loss, dp = onp.bce_loss(p.astype(np.float64), y.astype(np.float64))
    dz_ref = dp * p.astype(np.float64) * (1 - p.astype(np.float64))
    l, dz = ops.bce_sigmoid_fwd_bwd(_t(p, cuda_dev), _t(y, cuda_dev))
    assert abs(float(l) - loss) <= REL_F32 * abs(loss)
    assert_close(dz.cpu().numpy(), dz_ref, REL_F32, "bce dz")
