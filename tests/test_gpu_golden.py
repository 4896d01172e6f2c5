"""GPU parity against the committed golden fixtures (tests/golden/*.npz, written by
oracle/gen_golden.py): the CUDA path replays the stored inputs and must reproduce the stored
oracle outputs — bit-exact for integer/index/copy work, 1e-5 relative for fp32 arithmetic."""
import os

import numpy as np
import pytest

from util import REL_F32, assert_close

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _g(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_golden_interacting(cuda_dev):
    from recommendsystem_b200 import ops
    g = _g("interacting_cfg1")
    H, L, eps = int(g["H"]), int(g["L"]), float(g["ln_eps"])
    x, W, b, gm, bt, dy = (_t(g[k], cuda_dev) for k in ("x", "W", "b", "gamma", "beta", "dy"))
    y, saved = ops.interacting_fwd(x, W, b, gm, bt, eps, H, L)
    assert_close(y.cpu().numpy(), g["y"], REL_F32, "y")
    dx, dW, db, dg, dbt = ops.interacting_bwd(x, saved, W, b, gm, bt, eps, H, L, dy)
    for got, k in ((dx, "dx"), (dW, "dW"), (db, "db"), (dg, "dgamma"), (dbt, "dbeta")):
        assert_close(got.cpu().numpy(), g[k], REL_F32, k)


def test_golden_din(cuda_dev):
    from recommendsystem_b200 import cabi, ops
    g = _g("din_a")
    Ws = [_t(g[k], cuda_dev) for k in ("W1", "b1", "W2", "b2")]
    q, keys, values, sl, dout = (_t(g[k], cuda_dev) for k in ("q", "keys", "values", "seq_len", "dout"))
    out = ops.din_fwd(cabi.DIN_A, q, keys, values, sl, None, *Ws)
    assert_close(out.cpu().numpy(), g["out"], REL_F32, "din_a out")
    got = ops.din_bwd(cabi.DIN_A, q, keys, values, sl, None, *Ws, dout)
    for t, k in zip(got, ("dq", "dkeys", "dvalues", "dW1", "db1", "dW2", "db2")):
        assert_close(t.cpu().numpy().reshape(g[k].shape), g[k], REL_F32, "din_a " + k)
    g = _g("din_b")
    Ws = [_t(g[k], cuda_dev) for k in ("W1", "b1", "W2", "b2")]
    q, facts, mask, dout = (_t(g[k], cuda_dev) for k in ("q", "facts", "mask", "dout"))
    out = ops.din_fwd(cabi.DIN_B, q, facts, None, None, mask, *Ws)
    assert_close(out.cpu().numpy(), g["out"], REL_F32, "din_b out")
    dq, dfacts, _, dW1, db1, dW2, db2 = ops.din_bwd(cabi.DIN_B, q, facts, None, None, mask, *Ws, dout)
    for t, k in ((dq, "dq"), (dfacts, "dfacts"), (dW1, "dW1"), (db1, "db1"), (dW2, "dW2"), (db2, "db2")):
        assert_close(t.cpu().numpy().reshape(g[k].shape), g[k], 2 * REL_F32, "din_b " + k,
                     atol=1e-5 if k == "db2" else 0.0)


def test_golden_embedding(cuda_dev):
    from recommendsystem_b200 import ops
    g = _g("embedding")
    table, ids, rows, base = (_t(g[k], cuda_dev) for k in ("table", "ids", "rows", "row_base"))
    emb, keys, arows = ops.embed_gather(table, ids, base, rows, want_keys=True, want_rows=True)
    assert np.array_equal(emb.cpu().numpy(), g["emb"])                       # bit-exact
    assert np.array_equal(arows.cpu().numpy().astype(np.int64), g["arena_rows"])
    sr, inv, cnt, off = ops.route_ids(ids, ids.shape[1], rows, _t(g["local_base"], cuda_dev), int(g["world"]))
    assert np.array_equal(sr.cpu().numpy(), g["send_rows"]) and np.array_equal(inv.cpu().numpy(), g["inverse"])
    assert np.array_equal(cnt.cpu().numpy(), g["send_counts"]) and np.array_equal(off.cpu().numpy(), g["send_offsets"])
    w, m, v = table.clone(), torch.zeros_like(table), torch.zeros_like(table)
    scal = torch.zeros(4, device=cuda_dev)
    ops.adam_advance(scal, 0.9, 0.999)
    ks = ops.sort_keys(keys, ops.row_bits(table.shape[0]))
    ops.segsum_adam(w, m, v, _t(g["grad"], cuda_dev), ks, 1e-2, 0.9, 0.999, 1e-8, scal)
    assert_close(w.cpu().numpy(), g["adam_w"], REL_F32, "adam w")
    assert_close(m.cpu().numpy(), g["adam_m"], REL_F32, "adam m")
    assert_close(v.cpu().numpy(), g["adam_v"], REL_F32, "adam v")
