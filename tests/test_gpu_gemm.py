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


# ----------------------------------------------------------------- tcgen05 bf16 path
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (8192, 256, 624), (8192, 128, 256), (1000, 624, 256),
                                   (8192, 624, 256), (300, 72, 136), (129, 33, 8), (64, 16, 1024)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("cdt", ["bf16", "f32"])
def test_gemm_bf16_tc(cuda_dev, M, N, K, epi, cdt):
    """A[M,K] bf16, B stored [N,K] bf16 (K-major both), fp32 accumulation in TMEM."""
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N + K + epi)
    cd = torch.bfloat16 if cdt == "bf16" else torch.float32
    if cdt == "bf16" and N % 8 != 0:
        pytest.skip("bf16 C rows must be 16-byte aligned for this shape's vector path; covered by f32")
    A = _t(rng.standard_normal((M, K)).astype(np.float32), cuda_dev, torch.bfloat16)
    Bm = _t((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32), cuda_dev, torch.bfloat16)
    bias = rng.standard_normal(N).astype(np.float32)
    aux = _t((rng.random((M, N)) - 0.3).astype(np.float32), cuda_dev, cd)
    C = _t(rng.standard_normal((M, N)).astype(np.float32), cuda_dev, cd)
    acc = A.double().cpu().numpy() @ Bm.double().cpu().numpy().T
    ref = _ref_epi(acc, epi, bias.astype(np.float64), aux.double().cpu().numpy(), C.double().cpu().numpy())
    ops.gemm(A, Bm, C, bias=_t(bias, cuda_dev), aux=aux, epilogue=epi, transB=True)
    # fp32 outputs check the tensor-core accumulation tightly; bf16 outputs add one rounding
    assert_close(C.double().cpu().numpy(), ref, REL_F32 if cdt == "f32" else REL_BF16, f"gemm tc {cdt}")


@pytest.mark.parametrize("M,N,K", [(624, 256, 8192), (256, 128, 8192), (752, 8, 4096), (100, 50, 10000)])
def test_gemm_bf16_tc_splitk_wgrad(cuda_dev, M, N, K):
    """Weight-gradient shape: dW[M,N] = X^T[M,K=batch] dY[K,N] with both operands transposed to
    K-major first (rs_transpose2d); split-K partials summed in order => deterministic."""
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N)
    Kp = (K + 7) // 8 * 8
    X = _t(rng.standard_normal((K, M)).astype(np.float32), cuda_dev, torch.bfloat16)      # [batch, in]
    dY = _t(rng.standard_normal((K, N)).astype(np.float32) / 64, cuda_dev, torch.bfloat16)  # [batch, out]
    XT = torch.zeros(M, Kp, dtype=torch.bfloat16, device=cuda_dev)
    dYT = torch.zeros(N, Kp, dtype=torch.bfloat16, device=cuda_dev)
    ops.transpose2d(X, XT[:, :K])
    ops.transpose2d(dY, dYT[:, :K])
    assert torch.equal(XT[:, :K], X.t())
    dW = ops.gemm(XT[:, :K], dYT[:, :K], transB=True, out_dtype=torch.float32)
    dW2 = ops.gemm(XT[:, :K], dYT[:, :K], transB=True, out_dtype=torch.float32)
    ref = X.double().cpu().numpy().T @ dY.double().cpu().numpy()
    assert_close(dW.cpu().numpy(), ref, REL_F32, "wgrad split-K")
    assert torch.equal(dW, dW2)
    acc0 = torch.ones(M, N, device=cuda_dev)
    ops.gemm(XT[:, :K], dYT[:, :K], acc0, epilogue=6, transB=True)
    assert_close(acc0.cpu().numpy(), ref + 1.0, REL_F32, "wgrad split-K accumulate")


@pytest.mark.parametrize("M,N,K", [(624, 256, 8192), (256, 128, 8192), (752, 128, 4096), (128, 64, 64), (200, 72, 10000),
                                   (624, 40, 1000)])
def test_gemm_bf16_tc_mnmajor_wgrad(cuda_dev, M, N, K):
    """Weight gradient with NO transposed copies: dW[M,N] = X^T dY with X stored [K=batch, M] and dY stored
    [K, N] (transA=1, transB=0): both operands are MN-major TMA tiles / UMMA descriptors."""
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N + 1)
    X = _t(rng.standard_normal((K, M)).astype(np.float32), cuda_dev, torch.bfloat16)       # [batch, in]
    dY = _t(rng.standard_normal((K, N)).astype(np.float32) / 64, cuda_dev, torch.bfloat16)  # [batch, out]
    dW = ops.gemm(X, dY, transA=True, out_dtype=torch.float32)
    dW2 = ops.gemm(X, dY, transA=True, out_dtype=torch.float32)
    ref = X.double().cpu().numpy().T @ dY.double().cpu().numpy()
    assert_close(dW.cpu().numpy(), ref, REL_F32, "wgrad MN-major")
    assert torch.equal(dW, dW2)
    # strided views (columns of a wider buffer), as the trainer passes them
    big = torch.zeros(K, M + 16, dtype=torch.bfloat16, device=cuda_dev)
    big[:, 8:8 + M] = X
    dW3 = ops.gemm(big[:, 8:8 + M], dY, transA=True, out_dtype=torch.float32)
    assert torch.equal(dW, dW3)


def test_gemm_bf16_rejects_non_kmajor(cuda_dev):
    from recommendsystem_b200 import cabi, ops
    A = torch.zeros(64, 64, dtype=torch.bfloat16, device=cuda_dev)
    with pytest.raises(cabi.RsError):
        ops.gemm(A, A)            # A K-major with B MN-major: mixed majors are not built


@pytest.mark.parametrize("M,N,dt", [(1, 1, "f32"), (100, 37, "f32"), (8192, 624, "bf16"), (33, 65, "bf16")])
def test_transpose2d(cuda_dev, M, N, dt):
    from recommendsystem_b200 import ops
    d = torch.float32 if dt == "f32" else torch.bfloat16
    x = torch.randn(M, N, device=cuda_dev).to(d)
    assert torch.equal(ops.transpose2d(x), x.t().contiguous())


@pytest.mark.parametrize("B,zw,dt", [(1, 4, "f32"), (1000, 752, "f32"), (8192, 752, "f32"), (513, 100, "f32"),
                                     (4096, 2048, "f32"), (1000, 752, "bf16")])
def test_logit_head(cuda_dev, B, zw, dt):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B + zw)
    d = torch.float32 if dt == "f32" else torch.bfloat16
    rel = REL_F32 if dt == "f32" else REL_BF16
    Zt = _t(rng.standard_normal((B, zw)).astype(np.float32), cuda_dev, d)
    Z = Zt.double().cpu().numpy()
    w = (rng.standard_normal((zw, 1)) / np.sqrt(zw)).astype(np.float32)
    b = np.array([0.1], np.float32)
    y = (rng.random((B, 1)) < 0.25).astype(np.float32)
    p_raw = onp.dense(Z, w.astype(np.float64), b.astype(np.float64), "sigmoid")
    loss, dp = onp.bce_loss(p_raw, y.astype(np.float64))
    dz = dp * p_raw * (1 - p_raw)
    dZ = torch.empty(B, zw, dtype=d, device=cuda_dev)
    dw = torch.empty(zw, device=cuda_dev)
    db = torch.empty(1, device=cuda_dev)
    p, l = ops.logit_head(Zt, _t(w, cuda_dev), _t(b, cuda_dev), _t(y, cuda_dev), dZ, dw, db)
    assert_close(p.double().cpu().numpy(), p_raw, rel, "p")
    assert abs(float(l) - loss) <= rel * abs(loss)
    assert_close(dZ.double().cpu().numpy(), dz @ w.astype(np.float64).T, rel, "dZ")
    assert_close(dw.cpu().numpy(), (Z.T @ dz)[:, 0], 3 * rel, "dw")
    assert_close(db.cpu().numpy(), dz.sum(0), 3 * rel, "db")
    # deferred reduction (the trainer sums the partials on its side stream): same dw / db / loss bit for bit
    ws = ops.logit_head_workspace(B, zw, cuda_dev)
    dZ2, dw2, db2 = torch.empty_like(dZ), torch.full_like(dw, float("nan")), torch.full_like(db, float("nan"))
    l2 = torch.full((1,), float("nan"), device=cuda_dev)
    p2, _ = ops.logit_head(Zt, _t(w, cuda_dev), _t(b, cuda_dev), _t(y, cuda_dev), dZ2, dw2, db2, loss=l2, ws=ws,
                           defer_reduce=True)
    assert torch.isnan(dw2).all() and torch.isnan(l2).all()      # untouched until the reduce
    ops.logit_head_reduce(ws, dw2, db2, l2, B, zw)
    assert torch.equal(dZ2, dZ) and torch.equal(p2, p) and torch.equal(dw2, dw) and torch.equal(db2, db)
    assert torch.equal(l2, l)


@pytest.mark.parametrize("M,N,K", [(624, 256, 8192), (752, 1, 8192), (130, 70, 3000)])
def test_gemm_f32_splitk_wgrad(cuda_dev, M, N, K):
    """fp32 weight-gradient shape (A stored [K,M]): split-K partials, ordered reduce."""
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(M + N)
    X = _t(rng.standard_normal((K, M)).astype(np.float32), cuda_dev)
    dY = _t(rng.standard_normal((K, N)).astype(np.float32) / 64, cuda_dev)
    dW = ops.gemm(X, dY, transA=True)
    dW2 = ops.gemm(X, dY, transA=True)
    ref = X.double().cpu().numpy().T @ dY.double().cpu().numpy()
    assert_close(dW.cpu().numpy(), ref, REL_F32, "f32 wgrad split-K")
    assert torch.equal(dW, dW2)


@pytest.mark.parametrize("M,N,K", [(512, 200, 100), (2048, 256, 1712), (128, 32, 4096), (300, 72, 260), (4096, 8, 64),
                                   (16384, 8, 16), (16, 8, 16384), (8200, 12, 16), (16384, 256, 96), (9500, 512, 72)])
@pytest.mark.parametrize("tA,tB", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("epi", [0, 2, 6])
def test_gemm_f32_tf32x3(cuda_dev, M, N, K, tA, tB, epi):
    """fp32 operands on the tensor cores (3xTF32, hi/lo split in the kernel): every storage order of A and B,
    K / M / N tails, fused epilogues, split-K; fp32-grade accuracy and the FFMA kernel as a cross-check."""
    from recommendsystem_b200 import cabi, ops
    rng = np.random.default_rng(M + N + K + epi)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K) if tB else (K, N)) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    cold = rng.standard_normal((M, N)).astype(np.float32)
    acc = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    ref = _ref_epi(acc, epi, bias.astype(np.float64), None, cold.astype(np.float64))
    lib = cabi.load()
    n0 = lib.rs_launch_count()
    C = _t(cold, cuda_dev)
    ops.gemm(_t(A, cuda_dev), _t(B, cuda_dev), C, bias=_t(bias, cuda_dev), epilogue=epi, transA=tA, transB=tB)
    launches_tc = lib.rs_launch_count() - n0
    prev = lib.rs_set_fp32_gemm_mode(1)
    try:
        C2 = _t(cold, cuda_dev)
        ops.gemm(_t(A, cuda_dev), _t(B, cuda_dev), C2, bias=_t(bias, cuda_dev), epilogue=epi, transA=tA, transB=tB)
    finally:
        lib.rs_set_fp32_gemm_mode(prev)
    assert prev == 0 and launches_tc >= 1
    assert_close(C.cpu().numpy(), ref, REL_F32, "gemm 3xTF32")
    assert_close(C2.cpu().numpy(), ref, REL_F32, "gemm FFMA")
    # element-wise too: a dropped lo term would show up as ~1e-3 errors on individual outputs
    scale = np.abs(ref).max()
    assert np.abs(C.cpu().numpy() - ref).max() <= 2e-5 * scale


def test_gemm_f32_tf32x3_is_deterministic_and_handles_extremes(cuda_dev):
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(4)
    A = rng.standard_normal((1024, 512)).astype(np.float32)
    A[0, :8] = [0.0, -0.0, 1e-38, -1e-38, 3e38, 0.0, 1.0, 2.0 ** -126]        # zeros, near-denormal, near-max
    B = np.zeros((512, 64), np.float32)
    B[:8, 0] = 1.0
    B[8:, 1:] = rng.standard_normal((504, 63)).astype(np.float32) / 16
    At, Bt = _t(A, cuda_dev), _t(B, cuda_dev)
    C1, C2 = ops.gemm(At, Bt), ops.gemm(At, Bt)
    assert torch.equal(C1, C2)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    got = C1.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got[1:], ref[1:], rtol=0, atol=2e-5 * np.abs(ref).max())
    assert got[0, 0] == pytest.approx(ref[0, 0], rel=1e-6)
