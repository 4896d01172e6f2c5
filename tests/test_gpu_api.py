"""GPU parity of the drop-in layer API (recommendsystem_b200.api): same constructor surface and error
behaviour as the reference classes, outputs and autograd gradients against the oracle."""
import numpy as np
import pytest

from util import REL_F32, assert_close

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
f64 = lambda t: t.detach().double().cpu().numpy()


def test_interacting_layer_module(cuda_dev):
    from oracle import oracle_np as onp
    from recommendsystem_b200.api import InteractingLayer
    torch.manual_seed(0)
    layer = InteractingLayer(layer_num=3, unit_num=16, head_num=2, use_dropout=False, dropout_rate=0.3, use_res=True)
    x = torch.randn(9, 39, 16, device=cuda_dev, requires_grad=True)
    y = layer(x)
    assert y.shape == (9, 39, 16)
    names = dict(layer.named_parameters())
    assert {"query_dense_kernel", "key_dense_kernel", "value_dense_kernel", "res_dense_kernel",
            "layer_norm_gamma", "layer_norm_beta"} <= set(names)
    assert names["query_dense_kernel"].shape == (16, 16)            # Keras [in, out]
    W, b = layer.packed()
    ref = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(layer.layer_norm_gamma), f64(layer.layer_norm_beta),
                              1e-3, 2, 3, True)
    assert_close(f64(y), ref, REL_F32, "InteractingLayer fwd")
    dy = torch.randn_like(y)
    y.backward(dy)
    rdx, rdW, rdb, rdg, rdbt = onp.interacting_bwd(f64(x), f64(W), f64(b), f64(layer.layer_norm_gamma),
                                                   f64(layer.layer_norm_beta), 1e-3, 2, 3, f64(dy), True)
    assert_close(f64(x.grad), rdx, REL_F32, "dx")
    assert_close(f64(layer.key_dense_kernel.grad), rdW[:, 16:32], REL_F32, "dWk")
    assert_close(f64(layer.res_dense_bias.grad), rdb[48:], REL_F32, "dbr")
    assert_close(f64(layer.layer_norm_gamma.grad), rdg, REL_F32, "dgamma")


def test_interacting_layer_errors(cuda_dev):
    from recommendsystem_b200.api import InteractingLayer
    with pytest.raises(ValueError, match="The rank of input of InteractingLayer must be 3, but now is 2"):
        InteractingLayer(1, 16, 2)(torch.zeros(4, 16, device=cuda_dev))
    with pytest.raises(ValueError):
        InteractingLayer(1, 15, 2)
    with pytest.raises(ValueError):
        InteractingLayer(2, 8, 2)(torch.zeros(2, 3, 16, device=cuda_dev))     # layer_num>1 needs D == unit_num
    alias = InteractingLayer.from_deepctr(att_embedding_size=8, head_num=2, use_res=True)
    assert alias.unit_num == 16 and alias.layer_num == 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        InteractingLayer(1, 16, 2)(torch.zeros(2, 3, 16))


def test_din_modules(cuda_dev):
    from oracle import oracle_np as onp
    from recommendsystem_b200.api.din import DIN as DinA
    from recommendsystem_b200.api.staytime_layer import DIN as DinB
    torch.manual_seed(1)
    B, T, H = 33, 50, 16
    q = torch.randn(B, H, device=cuda_dev, requires_grad=True)
    keys = torch.randn(B, T, H, device=cuda_dev, requires_grad=True)
    vals = torch.randn(B, T, H, device=cuda_dev, requires_grad=True)
    sl = torch.randint(0, T + 1, (B,), device=cuda_dev)
    sl[0] = T
    a = DinA()
    out = a(q, keys, vals, sl)
    ref = onp.din_a_fwd(f64(q), f64(keys), f64(vals), sl.cpu().numpy(), f64(a.din_nn_0_kernel), f64(a.din_nn_0_bias),
                        f64(a.din_nn_1_kernel), f64(a.din_nn_1_bias))
    assert_close(f64(out), ref, REL_F32, "DIN-A")
    dout = torch.randn_like(out)
    out.backward(dout)
    refs = onp.din_a_bwd(f64(q), f64(keys), f64(vals), sl.cpu().numpy(), f64(a.din_nn_0_kernel), f64(a.din_nn_0_bias),
                         f64(a.din_nn_1_kernel), f64(a.din_nn_1_bias), f64(dout))
    assert_close(f64(q.grad), refs[0], REL_F32, "DIN-A dq")
    assert_close(f64(keys.grad), refs[1], REL_F32, "DIN-A dkeys")
    assert_close(f64(vals.grad), refs[2], REL_F32, "DIN-A dvalues")
    assert_close(f64(a.din_nn_0_kernel.grad), refs[3], REL_F32, "DIN-A dW1")
    with pytest.raises(ValueError):
        a(q, keys, vals, None)
    # variant B on a column slice of a wider sequence embedding (staytime/VideoDnn.py:68)
    seq = torch.randn(B, T, 32, device=cuda_dev, requires_grad=True)
    q2 = torch.randn(B, H, device=cuda_dev, requires_grad=True)
    mask = torch.arange(T, device=cuda_dev)[None, :] < sl[:, None]
    b = DinB()
    out = b(q2, seq[:, :, 0:16], mask)
    ref = onp.din_b_fwd(f64(q2), f64(seq)[:, :, :16], mask.cpu().numpy(), f64(b.layer_1_kernel), f64(b.layer_1_bias),
                        f64(b.layer_2_kernel), f64(b.layer_2_bias))
    assert_close(f64(out), ref, REL_F32, "DIN-B")
    dout = torch.randn_like(out)
    out.backward(dout)
    refs = onp.din_b_bwd(f64(q2), f64(seq)[:, :, :16], mask.cpu().numpy(), f64(b.layer_1_kernel), f64(b.layer_1_bias),
                         f64(b.layer_2_kernel), f64(b.layer_2_bias), f64(dout))
    assert_close(f64(q2.grad), refs[0], 2 * REL_F32, "DIN-B dq")
    assert_close(f64(seq.grad)[:, :, :16], refs[1], 2 * REL_F32, "DIN-B dfacts")
    assert float(seq.grad[:, :, 16:].abs().max()) == 0.0


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_dnn_mmoe_ple(cuda_dev, dt):
    from oracle import oracle_np as onp
    from recommendsystem_b200.api.rough_rank_layer import DNN, MMOE, PLE
    torch.manual_seed(2)
    d = torch.float32 if dt == "f32" else torch.bfloat16
    rel = REL_F32 if dt == "f32" else 2e-2
    x = torch.randn(300, 48, device=cuda_dev).to(d).requires_grad_(True)
    dnn = DNN((32, 16), activation="relu", output_activation="linear", seed=7)
    y = dnn(x)
    Ws = [f64(k) if dt == "f32" else f64(k.to(torch.bfloat16)) for k in dnn.kernels]
    ref, hs = onp.mlp_fwd(f64(x), Ws, [f64(b) for b in dnn.bias], ["relu", None], keep=True)
    assert_close(f64(y), ref, rel, "DNN fwd")
    dy = torch.randn_like(y)
    y.backward(dy)
    rdx, rdW, rdb = onp.mlp_bwd(hs, Ws, ["relu", None], f64(dy))
    assert_close(f64(x.grad), rdx, rel, "DNN dx")
    assert_close(f64(dnn.kernels[0].grad), rdW[0], rel, "DNN dW0")
    assert_close(f64(dnn.bias[1].grad), rdb[1], rel, "DNN db1")
    assert dnn.get_config()["hidden_units"] == [32, 16]
    x2 = torch.randn(64, 24, device=cuda_dev).to(d)
    outs = MMOE(num_tasks=3, num_experts=4, expert_dnn_units=(16,))(x2)
    assert len(outs) == 3 and outs[0].shape == (64, 16)
    ple = PLE(num_tasks=2, num_shared_experts=4, num_specific_experts=4, expert_dnn_units=(32,))
    outs = ple(x2)
    assert len(outs) == 2 and outs[1].shape == (64, 32)
    # gate-weighted sum against the oracle for task 0
    g = ple.gate_nets[0]
    gate = onp.softmax(f64(x2) @ f64(g.kernels[0].to(d)) + f64(g.bias[0]))
    ex = [np.maximum(f64(x2) @ f64(n.kernels[0].to(d)) + f64(n.bias[0]), 0)
          for n in list(ple.shared_expert_nets) + list(ple.specific_expert_nets[0])]
    ref0 = sum(gate[:, [e]] * ex[e] for e in range(8))
    assert_close(f64(outs[0]), ref0, rel, "PLE task0")


def test_crossnet_deepcross_fm_similarity(cuda_dev):
    from recommendsystem_b200.api.rough_rank_layer import CrossNet, KDLoss, Similarity
    from recommendsystem_b200.api.staytime_layer import DeepCrossLayer, FMLayer
    torch.manual_seed(3)
    x = torch.randn(20, 12, device=cuda_dev)
    cn = CrossNet(layer_num=2)
    y = cn(x)
    xl = f64(x)
    for i in range(2):                                     # x_{l+1} = x0 (x_l . w) + b + x_l
        xl = f64(x) * (xl @ f64(cn.kernels[i])) + f64(cn.bias[i]).T + xl
    assert_close(f64(y), xl, REL_F32, "CrossNet")
    dc = DeepCrossLayer(3)
    y = dc(x)
    c = f64(x)
    for i in range(3):
        c = f64(x) * (c @ f64(dc.W[i])) + f64(dc.b[i]) + c
    assert_close(f64(y), c, REL_F32, "DeepCrossLayer")
    e = torch.randn(5, 7, 16, device=cuda_dev)
    fm = FMLayer()(e)
    ref = 0.5 * ((f64(e).sum(1) ** 2) - (f64(e) ** 2).sum(1)).sum(-1, keepdims=True)   # [B,1]
    assert_close(f64(fm), ref, REL_F32, "FMLayer")
    with pytest.raises(ValueError):
        FMLayer()(x)
    s = Similarity(use_sigmoid=True)([x, x])
    assert s.shape == (20, 1)
    assert KDLoss()(x[:, :1], x[:, 1:2]).shape == (20,)


def test_embedding_features_and_sparse_optimizers(cuda_dev):
    from oracle import oracle_np as onp
    from recommendsystem_b200.api.embedding import AdaGrad, Adam, EmbeddingFeatures, category_column, embedding_column
    cols = [embedding_column(category_column("1591", 500), 16, combiner="mean"),
            embedding_column(category_column("1593", 300), 16, combiner="mean"),
            embedding_column(category_column("2125", 700), 16, combiner=None, seq_max_len=10)]
    emb = EmbeddingFeatures(cols, Adam(1e-2, 0.9, 0.999, 1e-8), "t", device=cuda_dev)
    g = torch.Generator().manual_seed(0)
    B = 40
    seq = torch.randint(0, 10 ** 6, (B, 10), generator=g)
    seq[torch.arange(10)[None, :] >= torch.randint(0, 11, (B, 1), generator=g)] = -1
    inputs = {"1591": torch.randint(0, 10 ** 6, (B,), generator=g), "1593": torch.randint(0, 10 ** 6, (B, 3), generator=g),
              "2125": seq}
    inputs["1593"][0, 1] = -1
    table0 = emb.table.cpu().numpy().copy()
    out = emb({k: v.to(cuda_dev) for k, v in inputs.items()})
    r0 = inputs["1591"].numpy() % 500
    assert np.array_equal(out["1591"].cpu().numpy(), table0[r0])                         # bit-exact gather
    e3, m3 = out["2125"]
    assert m3.shape == (B, 10) and np.array_equal(m3.cpu().numpy(), seq.numpy() >= 0)
    rseq = 800 + seq.numpy() % 700
    assert np.array_equal(e3.cpu().numpy(), np.where((seq.numpy() >= 0)[..., None], table0[rseq], 0))
    bag = inputs["1593"].numpy()
    ref_mean = np.stack([np.mean([table0[500 + i % 300] for i in row if i >= 0], axis=0) for row in bag])
    assert_close(out["1593"].cpu().numpy(), ref_mean, REL_F32, "bag mean")
    # sparse Adam push for the single-valued column
    grads = {"1591": torch.randn(B, 16, generator=g).to(cuda_dev), "1593": torch.zeros(B, 16, device=cuda_dev),
             "2125": torch.zeros(B, 10, 16, device=cuda_dev)}
    emb.backward(grads)
    _, _, corr = onp.adam_scalars(1, 0.9, 0.999)
    w, _, _ = onp.sparse_adam(table0.astype(np.float64), np.zeros_like(table0, np.float64), np.zeros_like(table0, np.float64),
                              r0, grads["1591"].cpu().numpy(), 1e-2, 0.9, 0.999, 1e-8, float(corr))
    assert_close(emb.table.cpu().numpy()[:500], w[:500], REL_F32, "sparse Adam via EmbeddingFeatures")
    # AdaGrad variant constructs and updates
    emb2 = EmbeddingFeatures(cols[:1], AdaGrad(0.005, 0.1, 0.1), "t2", device=cuda_dev)
    o = emb2({"1591": inputs["1591"].to(cuda_dev)})
    before = emb2.table.clone()
    emb2.backward({"1591": torch.ones_like(o["1591"])})
    assert not torch.equal(before, emb2.table)


def test_multihead_autoint_builder(cuda_dev):
    from recommendsystem_b200.api.builders import AUTOINT, AutoInt, cross_entropy
    slots = [str(1000 + i) for i in range(12)]
    # the reference signature (rank/multi_head/multidnn.py:214) and return type (ModelResult, :246-250)
    ret = AUTOINT(slots, [], True, dnn_hidden_units=(32, 16), bucket_size=1000, device=cuda_dev)
    model = ret.model
    assert ret.sub_model is model.sub_model and callable(ret.model_predict)
    g = torch.Generator().manual_seed(0)
    B = 64
    inputs = {s: torch.randint(0, 10 ** 6, (B,), generator=g).to(cuda_dev) for s in slots}
    labels = (torch.rand(B, 7, generator=g) < 0.3).float().to(cuda_dev)
    l0, pred = model.train_step(inputs, labels)
    assert pred.shape == (B, 7) and torch.isfinite(l0)
    assert ret.model_predict(inputs).shape == (B, 7)
    names = set(dict(ret.sub_model.named_parameters()))
    for k in ("dnn_0.kernel", "expert_7_fc1.kernel", "gate_6_fc2.bias", "unlike_pred.kernel",
              "interacting_layer.query_dense_kernel"):
        assert k in names, k                                   # Keras layer names of the reference graph
    model.opt.lr = 1e-2
    model.emb.opt.learning_rate = 1e-2
    for _ in range(60):
        l, _ = model.train_step(inputs, labels)
    assert float(l) < float(l0)
    assert cross_entropy(labels, pred).shape == (B, 1)
    cfg = {"model_param": {"interact": {"layer_num": 3, "unit_num": 16, "head_num": 2, "use_dropout": False,
                                         "dropout_rate": 0.0, "use_res": True},
                           "mlp": {"hidden_units": [32, 16], "activation": "relu"},
                           "logits": {"hidden_units": [1], "activation": "sigmoid"}},
           "feature": {"num_fields": 39, "rows_per_field": 100, "embed_dim": 16}, "batch": 32}
    out = AutoInt(cfg, device=cuda_dev).run()
    p = out["predict"](torch.randint(0, 100, (32, 39), device=cuda_dev))
    assert p.shape == (32, 1) and float(p.min()) >= 1e-6 and float(p.max()) <= 1.0


@pytest.mark.parametrize("B,dim,L", [(20, 12, 2), (700, 1712, 3), (1025, 400, 3), (3, 8, 1)])
def test_cross_network_kernels_fwd_bwd(cuda_dev, B, dim, L):
    """rs_cross_fwd / rs_cross_bwd (one fused row kernel for all cross layers) against the reference recurrence
    x_{l+1} = x0 (x_l . w_l) + b_l + x_l written out in float64 (rough_rank/layer.py:256-264, staytime/layer.py:66-72),
    forward 1e-5 and every gradient (inputs, kernels, biases) 1e-5 via torch float64 autograd; deterministic."""
    from recommendsystem_b200.api.functional import CrossFn
    g = torch.Generator(device=cuda_dev).manual_seed(B + dim)
    x = torch.randn(B, dim, device=cuda_dev, generator=g, requires_grad=True)
    W = (torch.randn(L, dim, device=cuda_dev, generator=g) * (2.0 / dim) ** 0.5).requires_grad_(True)
    b = (0.1 * torch.randn(L, dim, device=cuda_dev, generator=g)).requires_grad_(True)
    dout = torch.randn(B, dim, device=cuda_dev, generator=g)
    y = CrossFn.apply(x, W, b)
    y.backward(dout)
    x64, W64, b64 = (t.detach().double().requires_grad_(True) for t in (x, W, b))
    xl = x64
    for l in range(L):
        xl = x64 * (xl @ W64[l]).unsqueeze(1) + b64[l] + xl
    xl.backward(dout.double())
    assert_close(f64(y), f64(xl), REL_F32, "cross fwd")
    assert_close(f64(x.grad), f64(x64.grad), REL_F32, "cross dx")
    assert_close(f64(W.grad), f64(W64.grad), REL_F32, "cross dW")
    assert_close(f64(b.grad), f64(b64.grad), REL_F32, "cross db")
    x2 = x.detach().clone().requires_grad_(True)
    W2 = W.detach().clone().requires_grad_(True)
    CrossFn.apply(x2, W2, b.detach()).backward(dout)
    assert torch.equal(x2.grad, x.grad) and torch.equal(W2.grad, W.grad)


@pytest.mark.parametrize("opt_name", ["adam", "adagrad"])
def test_embedding_features_csr_bags(cuda_dev, opt_name):
    """True variable-length bags (VarLenFeature ids as CSR values/offsets, staytime/parse.py:22-23) through
    EmbeddingFeatures with combiner='mean' (VideoDnn.py:224-226): forward = mean over the valid ids of each bag (empty
    bag -> zeros), backward = 1/count per occurrence, duplicate rows summed, sparse optimizer applied — against a
    dense float64 restatement; and identical to the padded-[B, bag] input path."""
    from recommendsystem_b200.api.embedding import Adam, AdaGrad, EmbeddingFeatures, category_column, embedding_column
    mk = (lambda: Adam(1e-2, 0.9, 0.999, 1e-8)) if opt_name == "adam" else (lambda: AdaGrad(1e-2, 0.1, 0.1))
    rng = np.random.default_rng(4)
    B, R, d, maxbag = 37, 50, 16, 6
    lens = rng.integers(0, maxbag + 1, size=B)
    lens[0], lens[1] = 0, maxbag
    vals = rng.integers(0, 10 ** 9, size=int(lens.sum())).astype(np.int64)
    vals[3] = -1                                             # a padding id inside a bag is ignored
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    padded = -np.ones((B, maxbag), np.int64)
    for b in range(B):
        padded[b, :lens[b]] = vals[offs[b]:offs[b + 1]]
    col = lambda: [embedding_column(category_column("s", R), d, combiner="mean")]
    a = EmbeddingFeatures(col(), mk(), device=cuda_dev, seed=5)
    c = EmbeddingFeatures(col(), mk(), device=cuda_dev, seed=5)
    t0 = a.table.clone().double().cpu().numpy()
    csr = (torch.from_numpy(vals).to(cuda_dev), torch.from_numpy(offs).to(cuda_dev))
    oa = a({"s": csr})["s"]
    oc = c({"s": torch.from_numpy(padded).to(cuda_dev)})["s"]
    ref = np.zeros((B, d))
    for b in range(B):
        v = [x for x in vals[offs[b]:offs[b + 1]] if x >= 0]
        if v:
            ref[b] = t0[np.asarray(v) % R].mean(0)
    assert_close(f64(oa), ref, REL_F32, "CSR bag mean")
    assert_close(f64(oc), ref, REL_F32, "padded bag mean")
    g = torch.from_numpy(rng.standard_normal((B, d)).astype(np.float32)).to(cuda_dev)
    a.backward({"s": g})
    c.backward({"s": g})
    assert_close(f64(a.table), f64(c.table), 1e-6, "CSR update == padded update")
    # rows that no bag touched did not move; touched rows did
    touched = np.unique(vals[vals >= 0] % R)
    moved = np.abs(f64(a.table) - t0).max(1) > 0
    assert set(np.nonzero(moved)[0]) == set(touched.tolist())
