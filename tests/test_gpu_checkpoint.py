"""GPU: checkpoint save / load of the AutoInt trainer (tables + sparse Adam state + dense weights + dense
Adam state) and Keras-weight import.  Resuming must be exact: the kernels are deterministic, so
train(2) -> save -> load into a fresh trainer -> train(2) equals train(4) bit for bit."""
import numpy as np
import pytest

from util import REL_F32

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _cfg(dtype, B=256, hidden=(64, 32)):
    from recommendsystem_b200.autoint import AutoIntConfig
    rng = np.random.default_rng(5)
    return AutoIntConfig(num_fields=39, rows_per_field=[int(r) for r in rng.integers(50, 4000, size=39)],
                         embed_dim=16, unit_num=16, head_num=2, layer_num=3, mlp_hidden=hidden, batch=B,
                         dtype=dtype, lr_dense=1e-3, lr_sparse=1e-2)


def _batches(cfg, n, seed=11):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ids = rng.integers(0, 2 ** 40, size=(cfg.batch, cfg.num_fields)).astype(np.int64)
        y = (rng.random((cfg.batch, 1)) < 0.25).astype(np.float32)
        out.append((torch.from_numpy(ids), torch.from_numpy(y)))
    return out


def _state(tr):
    torch.cuda.synchronize()
    return (tr.arena.cpu().numpy().copy(), tr.flat.cpu().numpy().copy(), tr.flat_m.cpu().numpy().copy(),
            tr.flat_v.cpu().numpy().copy(), tr.adam_scalars.cpu().numpy().copy())


@pytest.mark.parametrize("dtype,graph", [("f32", False), ("bf16", True)])
def test_resume_is_exact(cuda_dev, tmp_path, dtype, graph):
    from recommendsystem_b200.autoint import AutoIntTrainer
    from recommendsystem_b200 import checkpoint as ck
    cfg = _cfg(dtype, hidden=(64, 32) if dtype == "f32" else (128, 64))
    batches = _batches(cfg, 4)
    dev = torch.device(cuda_dev)

    def run(tr, bs):
        losses = []
        for ids, y in bs:
            losses.append(float(tr.step(ids.to(dev), y.to(dev))))
        return losses

    a = AutoIntTrainer(cfg, cuda_dev)
    init = a.dense_state()
    tab0 = a.table.clone()
    la = run(a, batches)

    b = AutoIntTrainer(cfg, cuda_dev, tables=tab0, dense_init=init)
    lb = run(b, batches[:2])
    ck.save_checkpoint(b, tmp_path / "ck", step=2)
    meta = ck.read_meta(tmp_path / "ck")
    assert meta["step"] == 2 and meta["world"] == 1 and meta["config"]["num_fields"] == 39

    c = AutoIntTrainer(cfg, cuda_dev)            # different random init, everything overwritten by the load
    if graph:
        c.capture()                              # load must work INTO the captured buffers
    ck.load_checkpoint(c, tmp_path / "ck")
    for x, y in zip(_state(b), _state(c)):
        np.testing.assert_array_equal(x, y)
    lc = run(c, batches[2:])
    assert la == lb + lc                         # losses identical, bit for bit
    for x, y in zip(_state(a), _state(c)):
        np.testing.assert_array_equal(x, y)


def test_load_rejects_other_geometry(cuda_dev, tmp_path):
    from recommendsystem_b200.autoint import AutoIntConfig, AutoIntTrainer
    from recommendsystem_b200 import checkpoint as ck
    cfg = _cfg("f32", B=32)
    a = AutoIntTrainer(cfg, cuda_dev)
    ck.save_checkpoint(a, tmp_path / "ck")
    import dataclasses
    other = AutoIntTrainer(dataclasses.replace(cfg, mlp_hidden=(48, 32)), cuda_dev)
    with pytest.raises(ValueError):
        ck.load_checkpoint(other, tmp_path / "ck")
    ck.load_checkpoint(other, tmp_path / "ck", dense=False)        # tables alone still load
    np.testing.assert_array_equal(other.arena.cpu().numpy(), a.arena.cpu().numpy())
    rows = list(cfg.rows()); rows[0] += 1
    third = AutoIntTrainer(dataclasses.replace(cfg, rows_per_field=rows), cuda_dev)
    with pytest.raises(ValueError):
        ck.load_checkpoint(third, tmp_path / "ck")


def test_keras_import_export_roundtrip_and_oracle(cuda_dev):
    """Weights under the reference's Keras variable names go in, the forward agrees with the oracle run on
    the same arrays, and export gives the same arrays back."""
    from oracle import oracle_np as onp
    from recommendsystem_b200.autoint import AutoIntTrainer
    from recommendsystem_b200 import checkpoint as ck
    cfg = _cfg("f32", B=64)
    tr = AutoIntTrainer(cfg, cuda_dev)
    rng = np.random.default_rng(2)
    d = U = 16
    widths = [39 * d, 64, 32]
    w = {}
    for nm in ("query", "key", "value", "res"):
        w[f"{nm}_dense/kernel:0"] = rng.standard_normal((d, U)).astype(np.float32) * 0.3
        w[f"{nm}_dense/bias:0"] = rng.standard_normal(U).astype(np.float32) * 0.1
    w["layer_normalization/gamma:0"] = 1 + 0.1 * rng.standard_normal(U).astype(np.float32)
    w["layer_normalization/beta:0"] = 0.1 * rng.standard_normal(U).astype(np.float32)
    for i in range(2):
        w[f"mlp_{i}/kernel:0"] = rng.standard_normal((widths[i], widths[i + 1])).astype(np.float32) * 0.05
        w[f"mlp_{i}/bias:0"] = rng.standard_normal(widths[i + 1]).astype(np.float32) * 0.1
    w["logits/kernel:0"] = rng.standard_normal((32 + 39 * U, 1)).astype(np.float32) * 0.05
    w["logits/bias:0"] = np.array([0.1], np.float32)
    rep = ck.import_keras_autoint(tr, w)
    assert rep["Wqkvr"] == "query_dense/kernel|key_dense/kernel|value_dense/kernel|res_dense/kernel"
    back = ck.export_keras_autoint(tr)
    assert sorted(back) == sorted(w)
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])

    ids = rng.integers(0, 2 ** 40, size=(cfg.batch, 39)).astype(np.int64)
    p = tr.predict(torch.from_numpy(ids).to(cuda_dev)).cpu().numpy().reshape(-1)
    f64 = lambda a: np.asarray(a, np.float64)
    X, _ = onp.embed_gather(tr.table.cpu().numpy(), ids, tr.rows_host, tr.base_host)
    P = dict(Wqkvr=f64(np.concatenate([w[f"{n}_dense/kernel:0"] for n in ("query", "key", "value", "res")], 1)),
             bqkvr=f64(np.concatenate([w[f"{n}_dense/bias:0"] for n in ("query", "key", "value", "res")])),
             gamma=f64(w["layer_normalization/gamma:0"]), beta=f64(w["layer_normalization/beta:0"]),
             mlp_W=[f64(w[f"mlp_{i}/kernel:0"]) for i in range(2)], mlp_b=[f64(w[f"mlp_{i}/bias:0"]) for i in range(2)],
             out_W=f64(w["logits/kernel:0"]), out_b=f64(w["logits/bias:0"]))
    res = onp.autoint_fwd_bwd(f64(X), P, np.zeros((cfg.batch, 1)), cfg.head_num, cfg.layer_num, cfg.ln_eps)
    np.testing.assert_allclose(p, np.asarray(res["p"]).reshape(-1), rtol=REL_F32 * 10, atol=1e-6)

    with pytest.raises(KeyError):
        ck.import_keras_autoint(tr, {**w, "stray/kernel:0": np.zeros((1, 1), np.float32)})
    bad = dict(w); bad["logits/kernel:0"] = np.zeros((3, 1), np.float32)
    with pytest.raises(ValueError):
        ck.import_keras_autoint(tr, bad)
