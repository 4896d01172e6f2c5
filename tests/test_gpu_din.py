"""GPU parity: K6 fused DIN attention unit (both reference variants), forward and
backward, through the C-ABI vs oracle/oracle_np.py (din.py:18-47, staytime/layer.py:16-41)."""
import numpy as np
import pytest

from util import REL_BF16, REL_F32, assert_close, glorot_uniform

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(dtype) if dtype is not None else t


def _params(rng, nin, Hd=16):
    W1 = glorot_uniform(rng, nin, Hd)
    b1 = (0.1 * rng.standard_normal(Hd)).astype(np.float32)
    W2 = glorot_uniform(rng, Hd, 1)
    b2 = (0.1 * rng.standard_normal(1) + 0.2).astype(np.float32)
    return W1, b1, W2, b2


f64 = lambda a: a.astype(np.float64)

SHAPES = [(1, 1, 16), (5, 50, 16), (64, 100, 16), (33, 7, 8), (130, 128, 16), (9, 200, 16), (257, 65, 8)]


@pytest.mark.parametrize("B,T,H", SHAPES)
def test_din_a(cuda_dev, B, T, H):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import cabi, ops
    rng = np.random.default_rng(B + T + H)
    q = rng.standard_normal((B, H)).astype(np.float32)
    keys = rng.standard_normal((B, T, H)).astype(np.float32)
    values = rng.standard_normal((B, T, H)).astype(np.float32)
    seq_len = rng.integers(0, T + 1, size=B).astype(np.int32)
    seq_len[0] = T                                   # max(seq_length) == T (din.py:24)
    W1, b1, W2, b2 = _params(rng, 3 * H)
    dout = rng.standard_normal((B, H)).astype(np.float32)
    ref = onp.din_a_fwd(f64(q), f64(keys), f64(values), seq_len, f64(W1), f64(b1), f64(W2), f64(b2))
    refs = onp.din_a_bwd(f64(q), f64(keys), f64(values), seq_len, f64(W1), f64(b1), f64(W2), f64(b2), f64(dout))
    args = [_t(a, cuda_dev) for a in (q, keys, values)]
    sl = _t(seq_len, cuda_dev)
    Ws = [_t(a, cuda_dev) for a in (W1, b1, W2, b2)]
    out = ops.din_fwd(cabi.DIN_A, *args, sl, None, *Ws)
    assert_close(out.cpu().numpy(), ref, REL_F32, "din A fwd")
    got = ops.din_bwd(cabi.DIN_A, *args, sl, None, *Ws, _t(dout, cuda_dev))
    for g, r, name in zip(got, refs, ["dq", "dkeys", "dvalues", "dW1", "db1", "dW2", "db2"]):
        assert_close(g.cpu().numpy().reshape(r.shape), r, REL_F32, "din A " + name)
    got2 = ops.din_bwd(cabi.DIN_A, *args, sl, None, *Ws, _t(dout, cuda_dev))
    assert all(torch.equal(a, b) for a, b in zip(got, got2))          # deterministic


@pytest.mark.parametrize("B,T,H", SHAPES)
@pytest.mark.parametrize("with_mask", [True, False])
def test_din_b(cuda_dev, B, T, H, with_mask):
    from oracle import oracle_np as onp
    from recommendsystem_b200 import cabi, ops
    rng = np.random.default_rng(B + T + H + 1)
    q = rng.standard_normal((B, H)).astype(np.float32)
    facts = rng.standard_normal((B, T, H)).astype(np.float32)
    mask = None
    if with_mask:
        lens = rng.integers(0, T + 1, size=B)
        lens[0] = 0                                  # all-masked row -> uniform softmax = mean of facts
        mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.uint8)
    W1, b1, W2, b2 = _params(rng, 4 * H)
    dout = rng.standard_normal((B, H)).astype(np.float32)
    ref = onp.din_b_fwd(f64(q), f64(facts), mask, f64(W1), f64(b1), f64(W2), f64(b2))
    refs = onp.din_b_bwd(f64(q), f64(facts), mask, f64(W1), f64(b1), f64(W2), f64(b2), f64(dout))
    qt, ft = _t(q, cuda_dev), _t(facts, cuda_dev)
    mt = _t(mask, cuda_dev) if mask is not None else None
    Ws = [_t(a, cuda_dev) for a in (W1, b1, W2, b2)]
    out = ops.din_fwd(cabi.DIN_B, qt, ft, None, None, mt, *Ws)
    assert_close(out.cpu().numpy(), ref, REL_F32, "din B fwd")
    if with_mask:
        assert_close(out[0].cpu().numpy(), facts[0].astype(np.float64).mean(0), REL_F32, "all-masked row")
    dq, dfacts, _none, dW1, db1, dW2, db2 = ops.din_bwd(cabi.DIN_B, qt, ft, None, None, mt, *Ws, _t(dout, cuda_dev))
    for g, r, name in zip((dq, dfacts, dW1, db1, dW2, db2), refs, ["dq", "dfacts", "dW1", "db1", "dW2", "db2"]):
        # db2 is analytically 0 (softmax is shift invariant): absolute bound instead
        assert_close(g.cpu().numpy().reshape(r.shape), r, 2 * REL_F32, "din B " + name,
                     atol=1e-4 if name == "db2" else 0.0)


def test_din_b_strided_bf16(cuda_dev):
    """facts = first 16 columns of a [B,T,32] sequence embedding (staytime/VideoDnn.py:68), bf16."""
    from oracle import oracle_np as onp
    from recommendsystem_b200 import cabi, ops
    rng = np.random.default_rng(9)
    B, T, H = 40, 50, 16
    seq = _t(rng.standard_normal((B, T, 32)).astype(np.float32), cuda_dev, torch.bfloat16)
    qt = _t(rng.standard_normal((B, H)).astype(np.float32), cuda_dev, torch.bfloat16)
    facts = seq[:, :, :16]
    lens = rng.integers(1, T + 1, size=B)
    mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.uint8)
    W1, b1, W2, b2 = _params(rng, 4 * H)
    ref = onp.din_b_fwd(f64(qt.float().cpu().numpy()), f64(facts.float().cpu().numpy()), mask, f64(W1), f64(b1),
                        f64(W2), f64(b2))
    out = ops.din_fwd(cabi.DIN_B, qt, facts, None, None, _t(mask, cuda_dev),
                      *[_t(a, cuda_dev) for a in (W1, b1, W2, b2)], kv_ld=32)
    assert_close(out.float().cpu().numpy(), ref, REL_BF16, "din B strided bf16")
