"""oracle_np.py — CPU restatement (numpy, fp64 by default) of the reference's CTR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under recommendsystem_b200/ imports this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may.  It is the checker, never the product.

PINNING: the reference (yueshifeng/recommendSystem) ships no tests, no golden
vectors and depends on TensorFlow + tensornet, neither of which is installable
here (no network), so parity against TensorFlow's own arithmetic stays UNPINNED.
What IS pinned: the layer functions below (InteractingLayer, both DIN units,
Dense stacks) reproduce, to fp64 round-off, the outputs of the reference's own
layer files EXECUTED unmodified with `tensorflow` replaced by a numpy stand-in of
the individual TF ops (oracle/tf_numpy_shim.py, tools/gen_reference_layer_golden.py,
tests/golden/ref_layers.npz, tests/test_oracle_reference_pin.py) — i.e. the
reference's composition of those ops, not a re-reading of it.  The functions are
line-by-line restatements of the reference's Python sources (cited per function,
paths relative to the reference root); the other tests/golden/ vectors are
produced by THIS module (oracle/gen_golden.py), cross-checked against an
independent torch-CPU restatement (oracle/oracle_torch.py).  Third-party arithmetic restated from
published semantics: tf.keras.layers.Dense (y = act(x @ kernel[in,out] + bias)),
tf.nn.softmax (last axis), tf.keras.layers.LayerNormalization (biased variance
over the last axis, gamma/beta), tf.sequence_mask, Keras/TF Adam.

All functions are pure; arrays are row-major numpy.  `dt` selects the working
precision (np.float64 for the oracle proper, np.float32 to mimic TF's fp32).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------- helpers


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def dense(x, kernel, bias, act=None):
    """tf.keras.layers.Dense: act(x @ kernel[in,out] + bias) on the last axis."""
    y = x @ kernel + bias
    if act == "relu":
        return relu(y)
    if act == "sigmoid":
        return sigmoid(y)
    if act == "softmax":
        return softmax(y)
    assert act in (None, "linear"), act
    return y


def layer_norm(x, gamma, beta, eps):
    """LayerNormalization over the last axis (module missing from the reference,
    InteractingLayer.py:4; tf.keras semantics: biased variance, gamma, beta)."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    xhat = (x - mean) / np.sqrt(var + eps)
    return xhat * gamma + beta


# ------------------------------------------------------------ K1/K2 embedding


def embed_rows(ids, rows, row_base):
    """Arena row of every lookup.  ids int64 [..., F]; field f = last-axis index.
    row = row_base[f] + (id mod rows[f]); id < 0 (padding) -> -1.
    Restates tn.feature_column.category_column(key, bucket_size) + the table
    lookup of tn.layers.EmbeddingFeatures (call sites staytime/VideoDnn.py:217-244,
    rough_rank/model.py:89-115, rank/ctr/base_model.py:203-217); bucket_size is
    treated as the table's row count and ids are reduced mod it."""
    ids = np.asarray(ids, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    row_base = np.asarray(row_base, dtype=np.int64)
    r = row_base + np.mod(ids, rows)  # broadcasts over the last axis (F)
    return np.where(ids >= 0, r, -1).astype(np.int64)


def embed_gather(table, ids, rows, row_base):
    """[B,F] ids -> [B,F,d] rows of the fp32 arena (bag = 1: combiner='mean'
    degenerates to a copy).  Bit-exact operation."""
    r = embed_rows(ids, rows, row_base)
    out = table[np.maximum(r, 0)]
    out = np.where((r >= 0)[..., None], out, 0).astype(table.dtype)
    return out, r


def embed_gather_rows(table, rowidx):
    """Gather by arena row; rowidx < 0 -> zeros and mask 0 (sequence slots,
    staytime/VideoDnn.py:228-231 `combiner=None, seq_max_len=N` -> (emb, mask))."""
    rowidx = np.asarray(rowidx, dtype=np.int64)
    out = table[np.maximum(rowidx, 0)]
    mask = rowidx >= 0
    out = np.where(mask[..., None], out, 0).astype(table.dtype)
    return out, mask.astype(np.uint8)


def embed_bag_mean(table, ids, offsets, rows, row_base, F):
    """combiner='mean' over CSR bags (staytime/VideoDnn.py:224-226 with
    VarLenFeature ids, staytime/parse.py:22-23).  Bag k belongs to field k % F.
    fp32 accumulation in id order, then divide by the count; empty bag -> 0."""
    n_bags = len(offsets) - 1
    d = table.shape[1]
    out = np.zeros((n_bags, d), dtype=table.dtype)
    for k in range(n_bags):
        f = k % F
        acc = np.zeros(d, dtype=table.dtype)
        cnt = 0
        for p in range(int(offsets[k]), int(offsets[k + 1])):
            if ids[p] < 0:
                continue
            acc = acc + table[row_base[f] + ids[p] % rows[f]]
            cnt += 1
        if cnt:
            acc = acc / table.dtype.type(cnt)
        out[k] = acc
    return out


# ------------------------------------------------- K3 sparse grad + optimizers


def segment_sum_sorted(rowidx, grad):
    """Deterministic sorted-segment sum: stable sort of lookups by arena row,
    each run summed left-to-right in lookup order (fp32 if grad is fp32).
    Returns (unique rows ascending, summed grads).  Padding rows (<0) dropped."""
    rowidx = np.asarray(rowidx, dtype=np.int64).reshape(-1)
    grad = grad.reshape(len(rowidx), -1)
    order = np.argsort(rowidx, kind="stable")
    uniq, sums = [], []
    prev = None
    for p in order:
        r = rowidx[p]
        if r < 0:
            continue
        if r != prev:
            uniq.append(r)
            sums.append(grad[p].copy())
            prev = r
        else:
            sums[-1] = sums[-1] + grad[p]
    if not uniq:
        return np.zeros(0, np.int64), np.zeros((0, grad.shape[1]), grad.dtype)
    return np.asarray(uniq, np.int64), np.stack(sums)


def adam_scalars(step, beta1, beta2, dt=np.float32):
    """State after `step` calls of rs_adam_advance: powers carried
    multiplicatively in fp32 (like TF's beta_power variables), corr =
    sqrt(1-b2^t)/(1-b1^t) (Keras/TF Adam)."""
    p1 = dt(1.0)
    p2 = dt(1.0)
    for _ in range(step):
        p1 = dt(p1 * dt(beta1))
        p2 = dt(p2 * dt(beta2))
    corr = dt(np.sqrt(dt(1.0) - p2) / (dt(1.0) - p1))
    return p1, p2, corr


def _scaled_segment_sum(rowidx, grad, grad_scale, sum_dtype, out_dtype):
    """Per-row gradient: each occurrence scaled, then summed left-to-right in
    `sum_dtype` (fp32 = what an fp32 implementation computes; the order is the
    lookup order, so an fp32 kernel using the same order matches bit for bit)."""
    g = (grad.astype(sum_dtype) * sum_dtype(grad_scale)).astype(sum_dtype)
    uniq, sums = segment_sum_sorted(rowidx, g)
    return uniq, sums.astype(out_dtype)


def sparse_adam(w, m, v, rowidx, grad, lr, beta1, beta2, eps, corr, grad_scale=1.0,
                sum_dtype=np.float32):
    """tn.core.Adam(learning_rate, beta1, beta2, epsilon) applied to touched rows
    (call sites rank/multi_head/multidnn.py:235, rank/ctr/base_model.py:163,
    rough_rank/model.py:106).  TensorNet's source is not vendored: this is
    standard bias-corrected Adam in the Keras form, one update per touched row
    with the summed gradient.  In-place on copies; returns (w,m,v)."""
    w, m, v = w.copy(), m.copy(), v.copy()
    dt = w.dtype.type
    uniq, g = _scaled_segment_sum(rowidx, grad, grad_scale, sum_dtype, w.dtype)
    # hyper-parameters are fp32 values (TF holds them as float32 tensors): 1 - float32(0.999)
    # differs from 1 - 0.999 by 1.3e-5 relative
    b1, b2, lr_, eps_ = (dt(np.float32(t)) for t in (beta1, beta2, lr, eps))
    for r, gr in zip(uniq, g):
        m[r] = b1 * m[r] + (dt(1) - b1) * gr
        v[r] = b2 * v[r] + (dt(1) - b2) * gr * gr
        w[r] = w[r] - lr_ * dt(corr) * m[r] / (np.sqrt(v[r]) + eps_)
    return w, m, v


def sparse_adagrad(w, g2sum, rowidx, grad, lr, eps, per_element=False, grad_scale=1.0,
                   sum_dtype=np.float32):
    """tn.core.AdaGrad(learning_rate, initial_g2sum, initial_scale) (call sites
    staytime/VideoDnn.py:233,259).  Row mode (TensorNet-style): one scalar per
    row, g2sum += mean_d(g*g); w -= lr*g/(sqrt(g2sum)+eps).  per_element: the
    classic accumulator.  Arithmetic unpinned (TensorNet not vendored)."""
    w, g2sum = w.copy(), g2sum.copy()
    dt = w.dtype.type
    uniq, g = _scaled_segment_sum(rowidx, grad, grad_scale, sum_dtype, w.dtype)
    d = w.shape[1]
    for r, gr in zip(uniq, g):
        if per_element:
            g2sum[r] = g2sum[r] + gr * gr
            w[r] = w[r] - dt(lr) * gr / (np.sqrt(g2sum[r]) + dt(eps))
        else:
            g2sum[r] = g2sum[r] + np.sum(gr * gr) / dt(d)
            w[r] = w[r] - dt(lr) * gr / (np.sqrt(g2sum[r]) + dt(eps))
    return w, g2sum


def dense_adam(w, m, v, g, lr, beta1, beta2, eps, corr):
    """tn.optimizer.Optimizer(tn.core.Adam(...)) on a flat dense buffer
    (rank/multi_head/model.py:53, staytime/model.py:72)."""
    dt = w.dtype.type
    b1, b2, lr_, eps_ = (dt(np.float32(t)) for t in (beta1, beta2, lr, eps))
    m2 = b1 * m + (dt(1) - b1) * g
    v2 = b2 * v + (dt(1) - b2) * g * g
    w2 = w - lr_ * dt(corr) * m2 / (np.sqrt(v2) + eps_)
    return w2, m2, v2


# ---------------------------------------------------------------- K7 routing


def route_ids(ids, F, rows, local_base, world, pad_spread=False):
    """Row-sharded tables: r = id mod rows[f]; owner = r mod world;
    local row = local_base[f] + r // world.  Stable bucket-by-owner.
    Returns send_rows[n], inverse[n], send_counts[world], send_offsets[world+1].
    Padding ids (<0) go to owner 0 with row -1 (pad_spread: to owner ((uint32(i) * 0x9E3779B1) >> 16) mod world).  (Replaces TensorNet's
    sign->shard routing; only trace in the reference: staytime/parse.py:78-79.)"""
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    n = len(ids)
    f = np.arange(n) % F
    r = np.mod(ids, np.asarray(rows, np.int64)[f])
    spread = ((((np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)) >> np.uint64(16)) % np.uint64(world)).astype(np.int64)
    owner = np.where(ids >= 0, r % world, spread if pad_spread else 0)
    lrow = np.where(ids >= 0, np.asarray(local_base, np.int64)[f] + r // world, -1)
    order = np.argsort(owner, kind="stable")
    send_rows = lrow[order].astype(np.int32)
    inverse = np.empty(n, np.int32)
    inverse[order] = np.arange(n, dtype=np.int32)
    counts = np.bincount(owner, minlength=world).astype(np.int32)
    offsets = np.zeros(world + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    return send_rows, inverse, counts, offsets


def route_ids_padded(ids, F, rows, local_base, world, capacity, pad_spread=False):
    """Fixed-capacity layout of route_ids: bucket o owns slots [o*capacity, (o+1)*capacity),
    unused slots hold row -1; returns send_rows[world*capacity], inverse[n], counts, overflow."""
    sr, inv, cnt, off = route_ids(ids, F, rows, local_base, world, pad_spread)
    send = np.full(world * capacity, -1, np.int32)
    inverse = np.full(len(inv), -1, np.int32)
    owner_of_slot = np.searchsorted(off, np.arange(len(sr)), side="right") - 1
    k = np.arange(len(sr)) - off[owner_of_slot]
    ok = k < capacity
    new_slot = owner_of_slot * capacity + k
    send[new_slot[ok]] = sr[ok]
    remap = np.where(ok, new_slot, -1).astype(np.int32)
    inverse[:] = remap[inv]
    return send, inverse, cnt, int((cnt > capacity).any())


# ------------------------------------------------------- K4 InteractingLayer


def dropout_scale(seed, it, B, H, F, rate):
    """Inverted-dropout scale [H,B,F,F] of the attention weights of iteration `it` (InteractingLayer.py:53-54).
    TensorFlow's random stream cannot be reproduced; the kernels use a counter-based mask instead
    (include/rs_b200.h, rs_interacting_fwd_dropout): element (it, b, h, i, j) is kept iff the top 24 bits of
    splitmix64(linear index + seed) are >= rate * 2^24, kept weights scale by 1 / (1 - rate)."""
    h, b, i, j = np.meshgrid(np.arange(H, dtype=np.uint64), np.arange(B, dtype=np.uint64),
                             np.arange(F, dtype=np.uint64), np.arange(F, dtype=np.uint64), indexing="ij")
    with np.errstate(over="ignore"):
        idx = ((((np.uint64(it) * np.uint64(B) + b) * np.uint64(H) + h) * np.uint64(F) + i) * np.uint64(F)) + j
        z = idx + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    keep = u >= np.float32(rate)
    return np.where(keep, np.float32(1.0) / (np.float32(1.0) - np.float32(rate)), np.float32(0.0)).astype(np.float64)


def interacting_fwd(x, Wqkvr, bqkvr, gamma, beta, ln_eps, H, L, use_res=True, keep_cache=False,
                    iter_inputs=None, stored_act=None, dropout=None):
    """InteractingLayer.call (InteractingLayer.py:37-61; duplicate at
    rank/multi_head/interacting_layer.py).  Wqkvr = [Wq|Wk|Wv|Wr] ([D,4U], Keras
    [in,out] kernels side by side).  The four Dense(relu) layers and the
    LayerNorm are created once (:24-31) and re-applied in every iteration (:41)."""
    if x.ndim != 3:
        raise ValueError("The rank of input of InteractingLayer must be 3, but now is %d" % x.ndim)
    U = Wqkvr.shape[1] // 4
    dh = U // H
    B, F, _ = x.shape
    out = x
    cache = []
    for it in range(L):
        if iter_inputs is not None and it > 0:
            # evaluate iteration `it` at a GIVEN input (the activations an implementation stored for
            # its backward) instead of this function's own chain — used to check a backward pass
            # against the exact gradient at the same stored activations
            out = np.asarray(iter_inputs[it - 1], dtype=x.dtype).reshape(x.shape[0], x.shape[1], -1)
        if stored_act is not None and it > 0:
            # same idea when the implementation stored the pre-LayerNorm activations `act` of every
            # iteration (the tcgen05 path): iteration `it` starts from LayerNorm(stored act of it-1)
            out = layer_norm(np.asarray(stored_act[it - 1], dtype=x.dtype).reshape(B, F, U), gamma, beta, ln_eps)
        z = out @ Wqkvr + bqkvr                       # :42-46 (pre-activation)
        a = relu(z)
        q, k, v, r = a[..., :U], a[..., U:2 * U], a[..., 2 * U:3 * U], a[..., 3 * U:]
        # :47-49 split heads on the last axis, stack on the batch axis
        qh = q.reshape(B, F, H, dh).transpose(2, 0, 1, 3)   # [H,B,F,dh]
        kh = k.reshape(B, F, H, dh).transpose(2, 0, 1, 3)
        vh = v.reshape(B, F, H, dh).transpose(2, 0, 1, 3)
        s = qh @ kh.transpose(0, 1, 3, 2) / (dh ** 0.5)     # :50-51
        p = softmax(s)                                      # :52
        pm = p
        if dropout is not None and dropout[0] > 0:          # :53-54 (training): dropout = (rate, seed)
            pm = p * dropout_scale(dropout[1], it, B, H, F, dropout[0])
        o = (pm @ vh).transpose(1, 2, 0, 3).reshape(B, F, U)  # :55-56
        t = o + r if use_res else o                         # :57-58
        act = relu(t)                                       # :59
        if stored_act is not None:
            # ... and its LayerNorm/ReLU are differentiated at the stored `act` itself
            act = np.asarray(stored_act[it], dtype=x.dtype).reshape(B, F, U)
            t = act
        y = layer_norm(act, gamma, beta, ln_eps)            # :60
        if keep_cache:
            cache.append(dict(x=out, z=z, p=p, pm=pm, it=it, qh=qh, kh=kh, vh=vh, t=t, act=act))
        out = y
    return (out, cache) if keep_cache else out


def interacting_bwd(x, Wqkvr, bqkvr, gamma, beta, ln_eps, H, L, dy, use_res=True, iter_inputs=None,
                    stored_act=None, dropout=None):
    """Manual backward of interacting_fwd.  Returns dx, dW[D,4U], db[4U], dgamma, dbeta.
    iter_inputs: optional stored inputs of iterations 1..L-1; stored_act: optional stored
    pre-LayerNorm activations of iterations 0..L-1 (see interacting_fwd)."""
    U = Wqkvr.shape[1] // 4
    dh = U // H
    B, F, _ = x.shape
    _, cache = interacting_fwd(x, Wqkvr, bqkvr, gamma, beta, ln_eps, H, L, use_res, keep_cache=True,
                               iter_inputs=iter_inputs, stored_act=stored_act, dropout=dropout)
    dW = np.zeros_like(Wqkvr)
    db = np.zeros_like(bqkvr)
    dgamma = np.zeros_like(gamma)
    dbeta = np.zeros_like(beta)
    g = dy
    for c in reversed(cache):
        act = c["act"]
        mean = act.mean(-1, keepdims=True)
        var = ((act - mean) ** 2).mean(-1, keepdims=True)
        rstd = 1.0 / np.sqrt(var + ln_eps)
        xhat = (act - mean) * rstd
        dgamma = dgamma + (g * xhat).sum((0, 1))
        dbeta = dbeta + g.sum((0, 1))
        gg = g * gamma
        dact = (gg - gg.mean(-1, keepdims=True) - xhat * (gg * xhat).mean(-1, keepdims=True)) * rstd
        dt = dact * (c["t"] > 0)
        dr = dt if use_res else np.zeros_like(dt)
        do = dt.reshape(B, F, H, dh).transpose(2, 0, 1, 3)          # [H,B,F,dh]
        p, qh, kh, vh = c["p"], c["qh"], c["kh"], c["vh"]
        dv = c["pm"].transpose(0, 1, 3, 2) @ do
        dp = do @ vh.transpose(0, 1, 3, 2)
        if dropout is not None and dropout[0] > 0:                  # chain rule through p -> p * scale
            dp = dp * dropout_scale(dropout[1], c["it"], B, H, F, dropout[0])
        ds = p * (dp - (dp * p).sum(-1, keepdims=True)) / (dh ** 0.5)
        dq = ds @ kh
        dk = ds.transpose(0, 1, 3, 2) @ qh
        unh = lambda a: a.transpose(1, 2, 0, 3).reshape(B, F, U)
        da = np.concatenate([unh(dq), unh(dk), unh(dv), dr], axis=-1)
        dz = da * (c["z"] > 0)
        xin = c["x"]
        dW = dW + xin.reshape(-1, xin.shape[-1]).T @ dz.reshape(-1, 4 * U)
        db = db + dz.sum((0, 1))
        g = dz @ Wqkvr.T
    return g, dW, db, dgamma, dbeta


# ----------------------------------------------------------------- K6 DIN


def din_a_fwd(q, keys, values, seq_len, W1, b1, W2, b2):
    """DIN variant A (din.py:18-47): relu(relu([q,k,q*k]W1+b1)W2+b2), scores of
    positions t >= seq_len zeroed (tf.sequence_mask + tf.where, :24,39-42),
    out = scores @ values (:44-45).  No softmax."""
    B, T, Hd = keys.shape
    qt = np.broadcast_to(q[:, None, :], keys.shape)                 # :26,28
    z = np.concatenate([qt, keys, qt * keys], -1)                   # :31
    h = relu(z @ W1 + b1)                                           # din_nn_0 (:14-16,33)
    s = relu(h @ W2 + b2)[..., 0]                                   # din_nn_1, squeeze (:34,37)
    mask = np.arange(T)[None, :] < np.asarray(seq_len)[:, None]     # :24
    s = np.where(mask, s, 0)                                        # :42
    return np.einsum("bt,bth->bh", s, values)                       # :44-45


def din_b_fwd(q, facts, mask, W1, b1, W2, b2):
    """DIN variant B (staytime/layer.py:16-41): sigmoid hidden, linear score,
    masked positions set to -2**32+1 (:32-34), softmax over T (:35), p @ facts (:36)."""
    B, T, Hd = facts.shape
    qt = np.broadcast_to(q[:, None, :], facts.shape)                # :20-21
    z = np.concatenate([qt, facts, qt - facts, qt * facts], -1)     # :22-23
    h = sigmoid(z @ W1 + b1)                                        # :24
    s = (h @ W2 + b2)[..., 0]                                       # :25-26
    if mask is not None:
        pad = q.dtype.type(-2 ** 32 + 1)
        s = np.where(np.asarray(mask).astype(bool), s, pad)         # :30-34
    p = softmax(s)                                                  # :35
    return np.einsum("bt,bth->bh", p, facts)                        # :36-39


def din_a_bwd(q, keys, values, seq_len, W1, b1, W2, b2, dout):
    B, T, Hd = keys.shape
    qt = np.broadcast_to(q[:, None, :], keys.shape)
    z = np.concatenate([qt, keys, qt * keys], -1)
    h1 = z @ W1 + b1
    h = relu(h1)
    s1 = (h @ W2 + b2)[..., 0]
    mask = np.arange(T)[None, :] < np.asarray(seq_len)[:, None]
    s = np.where(mask, relu(s1), 0)
    dvalues = s[..., None] * dout[:, None, :]
    ds = np.einsum("bh,bth->bt", dout, values)
    ds1 = ds * mask * (s1 > 0)
    dW2 = (h * ds1[..., None]).sum((0, 1))[:, None]
    db2 = np.array([ds1.sum()])
    dh1 = (ds1[..., None] * W2[:, 0]) * (h1 > 0)
    dW1 = z.reshape(-1, 3 * Hd).T @ dh1.reshape(-1, W1.shape[1])
    db1 = dh1.sum((0, 1))
    dz = dh1 @ W1.T
    dqt, dk, dqk = dz[..., :Hd], dz[..., Hd:2 * Hd], dz[..., 2 * Hd:]
    dq = (dqt + dqk * keys).sum(1)
    dkeys = dk + dqk * qt
    return dq, dkeys, dvalues, dW1, db1, dW2, db2


def din_b_bwd(q, facts, mask, W1, b1, W2, b2, dout):
    B, T, Hd = facts.shape
    qt = np.broadcast_to(q[:, None, :], facts.shape)
    z = np.concatenate([qt, facts, qt - facts, qt * facts], -1)
    h = sigmoid(z @ W1 + b1)
    s = (h @ W2 + b2)[..., 0]
    mk = np.ones((B, T), bool) if mask is None else np.asarray(mask).astype(bool)
    s = np.where(mk, s, q.dtype.type(-2 ** 32 + 1))
    p = softmax(s)
    dfacts = p[..., None] * dout[:, None, :]
    dp = np.einsum("bh,bth->bt", dout, facts)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    ds = ds * mk                                   # tf.where routes no gradient to masked scores
    dW2 = (h * ds[..., None]).sum((0, 1))[:, None]
    db2 = np.array([ds.sum()])
    dh1 = (ds[..., None] * W2[:, 0]) * h * (1 - h)
    dW1 = z.reshape(-1, 4 * Hd).T @ dh1.reshape(-1, W1.shape[1])
    db1 = dh1.sum((0, 1))
    dz = dh1 @ W1.T
    dqt, df, dd, dm = dz[..., :Hd], dz[..., Hd:2 * Hd], dz[..., 2 * Hd:3 * Hd], dz[..., 3 * Hd:]
    dq = (dqt + dd + dm * facts).sum(1)
    dfacts = dfacts + df - dd + dm * qt
    return dq, dfacts, dW1, db1, dW2, db2


# ------------------------------------------------------------- K5 MLP / loss


def mlp_fwd(x, weights, biases, acts, keep=False):
    """MultiLayerDense (autoint:40-41,49-50; module missing -> `for u in units:
    x = Dense(u, activation)(x)`) and DNN.call (rough_rank/layer.py:100-109)."""
    hs = [x]
    for W, b, a in zip(weights, biases, acts):
        x = dense(x, W, b, a)
        hs.append(x)
    return (x, hs) if keep else x


def mlp_bwd(hs, weights, acts, dy):
    """Backward of mlp_fwd given the saved layer outputs hs[0..n]."""
    dWs, dbs = [], []
    g = dy
    for i in reversed(range(len(weights))):
        y = hs[i + 1]
        if acts[i] == "relu":
            g = g * (y > 0)
        elif acts[i] == "sigmoid":
            g = g * y * (1 - y)
        else:
            assert acts[i] in (None, "linear")
        dWs.append(hs[i].T @ g)
        dbs.append(g.sum(0))
        g = g @ weights[i].T
    return g, dWs[::-1], dbs[::-1]


def bce_loss(p_raw, y, a=1.0):
    """clip (autoint:52) + cross_entropy (rank/ctr/base_model.py:7-12):
    mean_b sum_k(-y log(p+1e-6) - (a-y) log(1-p+1e-6)).  Returns (loss, dL/dp_raw)."""
    dt = p_raw.dtype.type
    p = np.clip(p_raw, dt(1e-6), dt(1.0))
    B = p.shape[0]
    loss = np.mean(np.sum(-y * np.log(p + dt(1e-6)) - (a - y) * np.log(dt(1) - p + dt(1e-6)), axis=1))
    dp = (-y / (p + dt(1e-6)) + (a - y) / (dt(1) - p + dt(1e-6))) / B
    dp = dp * ((p_raw >= dt(1e-6)) & (p_raw <= dt(1.0)))
    return loss, dp


# ------------------------------------------------------------ AutoInt model


def autoint_fwd_bwd(X, P, y, H, L, ln_eps, use_res=True):
    """AutoInt.model_layer (autoint:18-56) + BaseModel.output_layer loss
    (rank/ctr/base_model.py:7-12,169-201) on the gathered field block X [B,F,d].
    P: dict with Wqkvr,bqkvr,gamma,beta, mlp_W[],mlp_b[] (relu), out_W,out_b (sigmoid).
    Returns dict(logits, loss, dX, grads{...})."""
    B, F, d = X.shape
    A = interacting_fwd(X, P["Wqkvr"], P["bqkvr"], P["gamma"], P["beta"], ln_eps, H, L, use_res)
    A2 = A.reshape(B, -1)                                           # Flatten (:36)
    deep, hs = mlp_fwd(X.reshape(B, -1), P["mlp_W"], P["mlp_b"], ["relu"] * len(P["mlp_W"]), keep=True)  # :39-41
    Z = np.concatenate([deep, A2], axis=1)                          # :45
    p_raw = dense(Z, P["out_W"], P["out_b"], "sigmoid")             # :49-50
    loss, dp = bce_loss(p_raw, y)                                   # :52 + loss
    dzl = dp * p_raw * (1 - p_raw)
    g_outW = Z.T @ dzl
    g_outb = dzl.sum(0)
    dZ = dzl @ P["out_W"].T
    n_deep = deep.shape[1]
    ddeep, dA2 = dZ[:, :n_deep], dZ[:, n_deep:]
    dXm, g_mlpW, g_mlpb = mlp_bwd(hs, P["mlp_W"], ["relu"] * len(P["mlp_W"]), ddeep)
    dXi, dW, db, dgamma, dbeta = interacting_bwd(
        X, P["Wqkvr"], P["bqkvr"], P["gamma"], P["beta"], ln_eps, H, L, dA2.reshape(A.shape), use_res)
    dX = dXm.reshape(X.shape) + dXi
    return dict(p=np.clip(p_raw, 1e-6, 1.0), p_raw=p_raw, loss=loss, dX=dX, A=A,
                grads=dict(Wqkvr=dW, bqkvr=db, gamma=dgamma, beta=dbeta, mlp_W=g_mlpW, mlp_b=g_mlpb,
                           out_W=g_outW, out_b=g_outb))


# ---------------------------------------------------------------- numerics model of the 3xTF32 GEMM
def tf32_truncate(x):
    """fp32 -> the tf32 value the tensor core sees when it ignores the low 13 mantissa bits."""
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def gemm_tf32x3_model(A, B):
    """What csrc/gemm_tc.cu::gemm_tf32x3_kernel computes, with exact (fp64) accumulation: x = hi + lo,
    hi = tf32_truncate(x), lo = x - hi (exact in fp32; the tensor core keeps its top 11 bits), and
    C = hi.hi + hi.lo + lo.hi.  The dropped lo.lo term and the truncation of lo bound the relative error of
    every product by ~2^-20; the GPU adds the accumulator's own rounding on top (tests/test_gpu_gemm.py)."""
    A = np.asarray(A, np.float32)
    B = np.asarray(B, np.float32)
    ah, bh = tf32_truncate(A), tf32_truncate(B)
    al, bl = tf32_truncate(A - ah), tf32_truncate(B - bh)
    f = lambda a: a.astype(np.float64)
    return f(ah) @ f(bh) + f(ah) @ f(bl) + f(al) @ f(bh)
