"""GPU parity: K4 fused InteractingLayer forward/backward through the C-ABI vs
oracle/oracle_np.py (fp64 restatement of InteractingLayer.py:37-61)."""
import numpy as np
import pytest

from util import REL_BF16, REL_F32, assert_close, interacting_params

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

# (B, F, D, U, H, L)
SHAPES = [
    (5, 39, 16, 16, 2, 3),      # BASELINE cfg1/cfg2 layer
    (64, 39, 16, 16, 2, 1),
    (33, 7, 8, 8, 2, 1),        # rank/multi_head: unit 8, 2 heads (multidnn.py:54)
    (3, 175, 8, 8, 2, 1),       # rank/ctr: F=175 (model_init.py:54-59)
    (17, 39, 16, 8, 2, 1),      # D != U, single iteration
    (9, 1, 16, 16, 4, 2),       # F=1: softmax over a single field
    (130, 128, 16, 16, 1, 2),
    (257, 40, 8, 8, 1, 2),
]


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(dtype) if dtype is not None else t


@pytest.mark.parametrize("B,F,D,U,H,L", SHAPES)
@pytest.mark.parametrize("ln_eps", [1e-3, 1e-9])
@pytest.mark.parametrize("use_res", [True, False])
def test_interacting_fp32(cuda_dev, B, F, D, U, H, L, ln_eps, use_res):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B + F + D + H + L)
    W, b, gamma, beta = interacting_params(rng, D, U)
    x = rng.standard_normal((B, F, D)).astype(np.float32)
    dy = rng.standard_normal((B, F, U)).astype(np.float32)
    f64 = lambda a: a.astype(np.float64)
    ref = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), ln_eps, H, L, use_res)
    rdx, rdW, rdb, rdg, rdbt = onp.interacting_bwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), ln_eps, H, L,
                                                   f64(dy), use_res)
    xt = _t(x, cuda_dev)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, ln_eps, H, L, use_res)
    assert_close(y.cpu().numpy(), ref, REL_F32, "interacting fwd")
    dx, dW, db, dg, dbt = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, ln_eps, H, L, _t(dy, cuda_dev), use_res)
    assert_close(dx.cpu().numpy(), rdx, REL_F32, "dx")
    assert_close(dW.cpu().numpy(), rdW, REL_F32, "dW")
    assert_close(db.cpu().numpy(), rdb, REL_F32, "db")
    assert_close(dg.cpu().numpy(), rdg, REL_F32, "dgamma")
    assert_close(dbt.cpu().numpy(), rdbt, REL_F32, "dbeta")
    # deterministic: a second run gives identical bits
    dx2, dW2, *_ = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, ln_eps, H, L, _t(dy, cuda_dev), use_res)
    assert torch.equal(dx, dx2) and torch.equal(dW, dW2)


@pytest.mark.parametrize("B,F,D,U,H,L", SHAPES[:3])
def test_interacting_bf16(cuda_dev, B, F, D, U, H, L):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B + F)
    W, b, gamma, beta = interacting_params(rng, D, U)
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    dyt = _t(rng.standard_normal((B, F, U)).astype(np.float32), cuda_dev, torch.bfloat16)
    f64 = lambda a: a.astype(np.float64)
    x, dy = xt.float().cpu().numpy(), dyt.float().cpu().numpy()
    ref = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L)
    rdx, rdW, rdb, rdg, rdbt = onp.interacting_bwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, f64(dy))
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L)
    assert_close(y.float().cpu().numpy(), ref, REL_BF16, "bf16 fwd")
    dx, dW, db, dg, dbt = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt)
    assert_close(dx.float().cpu().numpy(), rdx, 2 * REL_BF16, "bf16 dx")
    assert_close(dW.cpu().numpy(), rdW, 2 * REL_BF16, "bf16 dW")
    assert_close(dg.cpu().numpy(), rdg, 2 * REL_BF16, "bf16 dgamma")


def test_interacting_unsupported_shape_raises(cuda_dev):
    from recommendsystem_b200 import cabi, ops
    x = torch.zeros(2, 3, 24, device=cuda_dev)
    with pytest.raises(cabi.RsError):
        ops.interacting_fwd(x, torch.zeros(24, 96, device=cuda_dev), torch.zeros(96, device=cuda_dev),
                            torch.ones(24, device=cuda_dev), torch.zeros(24, device=cuda_dev), 1e-3, 2, 1)


TC_SHAPES = [(1, 39, 1), (3, 39, 1), (5, 39, 3), (1000, 39, 3), (64, 33, 2), (7, 40, 1),
             (37, 26, 3), (100, 16, 2), (50, 13, 2), (21, 48, 2), (300, 5, 1), (9, 1, 2), (33, 24, 3)]


@pytest.mark.parametrize("B,F,L", TC_SHAPES)
@pytest.mark.parametrize("use_res", [True, False])
def test_interacting_tc_fwd(cuda_dev, B, F, L, use_res):
    """tcgen05 forward (compute_bf16=1): bf16 projection operands, tf32 QK^T, bf16 P.V, fp32 accumulation
    and fp32 softmax / LayerNorm — against the fp64 oracle on the same bf16 inputs (1e-2 bf16 bar; the
    bf16 rounding of W and P inside the kernel is part of what the bar covers)."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    D = U = 16
    H = 2
    rng = np.random.default_rng(B + F + L)
    W, b, gamma, beta = interacting_params(rng, D, U)
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    f64 = lambda a: a.astype(np.float64)
    ref = onp.interacting_fwd(f64(xt.float().cpu().numpy()), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, use_res)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    from recommendsystem_b200 import cabi
    assert ops.interacting_path(F, D, U, H, torch.bfloat16, True) == cabi.PATH_TCGEN05      # the kernels under test
    assert ops.interacting_path(F, D, U, H, torch.bfloat16, False) == cabi.PATH_FFMA
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, use_res, compute_bf16=True)
    y0, _ = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, use_res, compute_bf16=False)
    from util import rel_err
    print("tc vs oracle", rel_err(y.float().cpu().numpy(), ref), "ffma vs oracle", rel_err(y0.float().cpu().numpy(), ref))
    assert_close(y.float().cpu().numpy(), ref, REL_BF16, "tc fwd")
    if L > 1:
        ref1 = onp.interacting_fwd(f64(xt.float().cpu().numpy()), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, 1, use_res)
        # tensor-core mode saves the pre-LayerNorm activations; LayerNorm(saved[0]) is iteration 1's input
        y1 = onp.layer_norm(f64(saved[0].cpu().numpy()).reshape(B, F, U), f64(gamma), f64(beta), 1e-3)
        assert_close(y1, ref1, REL_BF16, "LayerNorm(saved[0])")


@pytest.mark.parametrize("B,F,L", TC_SHAPES)
@pytest.mark.parametrize("use_res", [True, False])
def test_interacting_tc_bwd(cuda_dev, B, F, L, use_res):
    """tcgen05 backward (compute_bf16=1) against the fp64 oracle gradient evaluated at the activations the
    tcgen05 forward stored (`saved` = pre-LayerNorm activations), the point the kernel differentiates at.  db is checked per
    q|k|v|r segment so that a wrong MMA is localised."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    from util import rel_err
    D = U = 16
    H = 2
    rng = np.random.default_rng(100 + B + F + L)
    W, b, gamma, beta = interacting_params(rng, D, U)
    gamma = (gamma + 0.1 * rng.standard_normal(U)).astype(np.float32)
    beta = (beta + 0.1 * rng.standard_normal(U)).astype(np.float32)
    b = (b + 0.1 * rng.standard_normal(4 * U)).astype(np.float32)
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    dyt = _t(rng.standard_normal((B, F, U)).astype(np.float32), cuda_dev, torch.bfloat16)
    f64 = lambda a: a.astype(np.float64)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    from recommendsystem_b200 import cabi
    assert ops.interacting_path(F, D, U, H, torch.bfloat16, True) == cabi.PATH_TCGEN05
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, use_res, compute_bf16=True)
    acts = [f64(saved[i].cpu().numpy()) for i in range(L)]
    rdx, rdW, rdb, rdg, rdbt = onp.interacting_bwd(f64(xt.float().cpu().numpy()), f64(W), f64(b), f64(gamma), f64(beta),
                                                   1e-3, H, L, f64(dyt.float().cpu().numpy()), use_res, stored_act=acts)
    dx, dW, db, dg, dbt = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, use_res, compute_bf16=True)
    names = ["dx", "dW", "db", "dgamma", "dbeta"]
    got = [t.float().cpu().numpy() for t in (dx, dW, db, dg, dbt)]
    refs = [rdx, rdW, rdb, rdg, rdbt]
    for n, a, r in zip(names, got, refs):
        print(n, "tc", rel_err(a, r))
    for i, seg in enumerate("qkvr"):
        print("db", seg, rel_err(got[2][i * U:(i + 1) * U], rdb[i * U:(i + 1) * U]))
    for n, a, r in zip(names, got, refs):
        assert_close(a, r, 2 * REL_BF16, "tc " + n)
    dx2, dW2, *_ = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, use_res, compute_bf16=True)
    assert torch.equal(dx, dx2) and torch.equal(dW, dW2)


@pytest.mark.parametrize("B,F,D,U,H,L", [(5, 39, 16, 16, 2, 3), (33, 7, 8, 8, 2, 1), (3, 175, 8, 8, 2, 1), (9, 40, 16, 16, 4, 2)])
@pytest.mark.parametrize("rate", [0.2, 0.5])
def test_interacting_attention_dropout(cuda_dev, B, F, D, U, H, L, rate):
    """Training-mode attention-weight dropout (InteractingLayer.py:53-54; multidnn.py:54, model_init.py:54-59)
    fused into forward and backward: the counter-based mask is regenerated by the oracle from the same seed, so
    outputs and every gradient are compared at the fp32 bar; rate 0 reproduces the no-dropout kernels bit for bit."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    rng = np.random.default_rng(B + F + int(rate * 10))
    W, b, gamma, beta = interacting_params(rng, D, U)
    x = rng.standard_normal((B, F, D)).astype(np.float32)
    dy = rng.standard_normal((B, F, U)).astype(np.float32)
    seed = 0x1234567890ABCDEF + B
    f64 = lambda a: a.astype(np.float64)
    drop = (rate, seed)
    ref = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, True, dropout=drop)
    rdx, rdW, rdb, rdg, rdbt = onp.interacting_bwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, f64(dy), True,
                                                   dropout=drop)
    xt, dyt = _t(x, cuda_dev), _t(dy, cuda_dev)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, dropout_rate=rate, dropout_seed=seed)
    assert_close(y.cpu().numpy(), ref, REL_F32, "dropout fwd")
    ref0 = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, True)
    assert np.abs(ref - ref0).max() > 1e-3            # the mask really changes the output
    dx, dW, db, dg, dbt = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, True, dropout_rate=rate,
                                              dropout_seed=seed)
    for n, a, r in zip(["dx", "dW", "db", "dgamma", "dbeta"], [dx, dW, db, dg, dbt], [rdx, rdW, rdb, rdg, rdbt]):
        assert_close(a.cpu().numpy(), r, REL_F32, "dropout " + n)
    y0, s0 = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, dropout_rate=0.0, dropout_seed=seed)
    y1, s1 = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True)
    assert torch.equal(y0, y1)
    # a different seed gives a different mask
    y2, _ = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, dropout_rate=rate, dropout_seed=seed + 1)
    assert not torch.equal(y, y2)


def test_interacting_layer_module_dropout(cuda_dev):
    """api.InteractingLayer(use_dropout=True): active in train mode (fresh mask per call, reproducible from
    last_dropout_seed), inactive in eval mode; bf16 input with dropout falls back to the FFMA kernels."""
    from oracle import oracle_np as onp
    from recommendsystem_b200.api import InteractingLayer
    torch.manual_seed(0)
    layer = InteractingLayer(layer_num=1, unit_num=8, head_num=2, use_dropout=True, dropout_rate=0.2, use_res=True)
    x = torch.randn(6, 11, 8, device=cuda_dev, requires_grad=True)
    layer.train()
    y_a = layer(x)
    seed_a = layer.last_dropout_seed
    y_b = layer(x)
    assert not torch.equal(y_a, y_b) and layer.last_dropout_seed != seed_a
    W, b = layer.packed()
    f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
    ref = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(layer.layer_norm_gamma), f64(layer.layer_norm_beta), 1e-3, 2, 1,
                              True, dropout=(0.2, seed_a))
    assert_close(y_a.detach().cpu().numpy(), ref, REL_F32, "module dropout fwd")
    y_a.sum().backward()
    assert torch.isfinite(x.grad).all()
    layer.eval()
    ref0 = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(layer.layer_norm_gamma), f64(layer.layer_norm_beta), 1e-3, 2, 1, True)
    assert_close(layer(x).detach().cpu().numpy(), ref0, REL_F32, "module eval fwd")


def test_interacting_path_is_exported(cuda_dev):
    """rs_interacting_path reports which kernels a call runs, so a test cannot pass on the wrong kernel silently:
    tcgen05 only for (D = U = 16, H = 2, F <= 48, bf16, no dropout); FFMA for every other built shape."""
    from recommendsystem_b200 import cabi, ops
    bf, f32 = torch.bfloat16, torch.float32
    assert ops.interacting_path(39, 16, 16, 2, bf, True) == cabi.PATH_TCGEN05
    assert ops.interacting_path(48, 16, 16, 2, bf, True) == cabi.PATH_TCGEN05
    assert ops.interacting_path(49, 16, 16, 2, bf, True) == cabi.PATH_FFMA       # F > 48
    assert ops.interacting_path(39, 16, 16, 2, bf, True, 0.2) == cabi.PATH_FFMA  # attention dropout
    assert ops.interacting_path(39, 16, 16, 2, f32, True) == cabi.PATH_FFMA      # fp32 activations
    assert ops.interacting_path(39, 16, 16, 2, bf, False) == cabi.PATH_FFMA      # parity mode
    assert ops.interacting_path(175, 8, 8, 2, bf, True) == cabi.PATH_FFMA        # rank/ctr shape
    assert ops.interacting_path(39, 24, 24, 2, bf, True) == cabi.PATH_NONE       # not built at all


@pytest.mark.parametrize("B,F", [(257, 39), (2000, 20), (3, 48)])
def test_interacting_tc_bwd_deferred_reduce(cuda_dev, B, F):
    """rs_interacting_bwd_scatter(dparams = NULL) + rs_interacting_bwd_reduce (the trainer runs the reduction on its
    side stream) gives the parameter gradients of the undeferred call bit for bit, and the same dx."""
    from recommendsystem_b200 import cabi, ops
    L, D, U, H = 2, 16, 16, 2
    rng = np.random.default_rng(B + F)
    W, b, gamma, beta = [_t(a, cuda_dev) for a in interacting_params(rng, D, U)]
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    dy = _t(rng.standard_normal((B, F, U)).astype(np.float32), cuda_dev, torch.bfloat16)
    assert ops.interacting_path(F, D, U, H, torch.bfloat16, True) == cabi.PATH_TCGEN05
    _, saved = ops.interacting_fwd(xt, W, b, gamma, beta, 1e-3, H, L, True, compute_bf16=True)
    dx0, dW0, db0, dg0, dbt0 = ops.interacting_bwd(xt, saved, W, b, gamma, beta, 1e-3, H, L, dy, compute_bf16=True)
    ref = torch.cat([dW0.reshape(-1), db0, dg0, dbt0]).clone()
    ws = torch.empty(cabi.load().rs_interacting_workspace_bytes(B, F, D, U), dtype=torch.uint8, device=cuda_dev)
    dx = torch.empty_like(xt)
    st = ops._stream()
    cabi.call("rs_interacting_bwd_scatter", xt.data_ptr(), D, 0, saved.data_ptr(), cabi.RS_BF16, W.data_ptr(), b.data_ptr(),
              gamma.data_ptr(), beta.data_ptr(), 1e-3, dy.data_ptr(), U, 0, dx.data_ptr(), D, 0, None, None, 1, 0, None, 0,
              None, B, F, D, U, H, L, 1, ws.data_ptr(), ws.numel(), st)
    out = torch.full_like(ref, float("nan"))
    cabi.call("rs_interacting_bwd_reduce", ws.data_ptr(), ws.numel(), out.data_ptr(), B, F, D, U, H, st)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx0)
    assert torch.equal(out, ref)
    with pytest.raises(cabi.RsError):      # too small a workspace is refused, not read out of bounds
        cabi.call("rs_interacting_bwd_reduce", ws.data_ptr(), 16, out.data_ptr(), B, F, D, U, H, st)


def test_interacting_tc_bench_size(cuda_dev):
    """The benchmarked launch itself: B = 8192, F = 39, L = 3 (2731 tiles over the persistent CTAs, every
    accumulator alias and the cross-tile prefetch exercised thousands of times) forward and backward against the
    fp64 oracle; same tolerances as the small cases."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import cabi, ops
    from util import rel_err
    B, F, L, D, U, H = 8192, 39, 3, 16, 16, 2
    rng = np.random.default_rng(8192)
    W, b, gamma, beta = interacting_params(rng, D, U)
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    dyt = _t(rng.standard_normal((B, F, U)).astype(np.float32), cuda_dev, torch.bfloat16)
    f64 = lambda a: a.astype(np.float64)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    assert ops.interacting_path(F, D, U, H, torch.bfloat16, True) == cabi.PATH_TCGEN05
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, compute_bf16=True)
    x64 = f64(xt.float().cpu().numpy())
    ref = onp.interacting_fwd(x64, f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, True)
    # 5.1 M outputs: the 1e-2 bar is applied to all but the worst 1e-5 of them (51 elements) and 2e-2 to the very
    # worst — at this count the extreme tail of the ReLU / LayerNorm conditioning shows up: the same
    # statistic of the fp32-arithmetic FFMA kernels on the same bf16 inputs is printed beside it
    def tail(t):
        e = np.abs(f64(t.float().cpu().numpy()) - ref).ravel() / np.max(np.abs(ref))
        return float(np.quantile(e, 1 - 1e-5)), float(e.max()), float(np.sqrt(np.mean(e * e)))
    y0, _ = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, compute_bf16=False)
    q_tc, m_tc, r_tc = tail(y)
    q_ff, m_ff, r_ff = tail(y0)
    print(f"tc   : q(1-1e-5) {q_tc:.3e} max {m_tc:.3e} rms {r_tc:.3e}")
    print(f"ffma : q(1-1e-5) {q_ff:.3e} max {m_ff:.3e} rms {r_ff:.3e}")
    assert q_tc <= REL_BF16 and m_tc <= 2 * REL_BF16, (q_tc, m_tc)
    acts = [f64(saved[i].cpu().numpy()) for i in range(L)]
    refs = onp.interacting_bwd(x64, f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, f64(dyt.float().cpu().numpy()),
                               True, stored_act=acts)
    got = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, True, compute_bf16=True)
    for n, a, r in zip(["dx", "dW", "db", "dgamma", "dbeta"], got, refs):
        print(n, rel_err(a.float().cpu().numpy(), r))
        assert_close(a.float().cpu().numpy(), r, 2 * REL_BF16, "tc B=8192 " + n)
    got2 = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, True, compute_bf16=True)
    assert all(torch.equal(a, c) for a, c in zip(got, got2))          # deterministic at 296 CTAs


@pytest.mark.parametrize("B,F,L", [(512, 39, 3), (2048, 39, 1), (300, 26, 2)])
def test_interacting_tc_bwd_end_to_end(cuda_dev, B, F, L):
    """bf16 gradients against the fp64 oracle END TO END on the same bf16-rounded inputs — the oracle is NOT handed
    the kernel's stored activations here.  The layer's gradient is ill-conditioned in its input (rounding x alone
    moves dX by 18-42 % in max norm at L = 3: ReLU masks flip under LayerNorm), so the element-wise
    1e-2 bar cannot apply; what must hold for training is direction and size: cosine similarity >= 0.99 for every
    gradient and the norm within 5 %.  The norm-wise max error is printed for the record."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import ops
    from util import rel_err
    D = U = 16
    H = 2
    rng = np.random.default_rng(900 + B + F + L)
    W, b, gamma, beta = interacting_params(rng, D, U)
    xt = _t(rng.standard_normal((B, F, D)).astype(np.float32), cuda_dev, torch.bfloat16)
    dyt = _t(rng.standard_normal((B, F, U)).astype(np.float32), cuda_dev, torch.bfloat16)
    f64 = lambda a: a.astype(np.float64)
    Wt, bt, gt, bet = (_t(a, cuda_dev) for a in (W, b, gamma, beta))
    y, saved = ops.interacting_fwd(xt, Wt, bt, gt, bet, 1e-3, H, L, True, compute_bf16=True)
    got = ops.interacting_bwd(xt, saved, Wt, bt, gt, bet, 1e-3, H, L, dyt, True, compute_bf16=True)
    refs = onp.interacting_bwd(f64(xt.float().cpu().numpy()), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L,
                               f64(dyt.float().cpu().numpy()), True)
    for n, a, r in zip(["dx", "dW", "db", "dgamma", "dbeta"], got, refs):
        a = f64(a.float().cpu().numpy()).ravel()
        r = f64(r).ravel()
        cos = float(a @ r / (np.linalg.norm(a) * np.linalg.norm(r)))
        ratio = float(np.linalg.norm(a) / np.linalg.norm(r))
        print(f"{n}: cosine {cos:.5f}  |kernel|/|oracle| {ratio:.4f}  norm-wise max error {rel_err(a, r):.3e}")
        assert cos >= 0.99, (n, cos)
        assert abs(ratio - 1.0) <= 0.05, (n, ratio)
