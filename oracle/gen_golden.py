"""Generates the small golden input/output fixtures under tests/golden/ from the oracle.

    python -m oracle.gen_golden        # rewrites tests/golden/*.npz

The reference has no golden vectors of its own and cannot run here (TensorFlow and
tensornet are not installable), so these fixtures pin the ORACLE (seeded inputs ->
outputs of oracle_np in fp64, inputs stored as fp32) — "parity unpinned" with respect
to TensorFlow itself.  The GPU parity tests replay them through the CUDA path.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import numpy as np

from . import oracle_np as onp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _glorot(rng, fi, fo):
    lim = np.sqrt(6.0 / (fi + fo))
    return rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)


def generate():
    out = {}
    f64 = lambda a: np.asarray(a, np.float64)
    # --- InteractingLayer, BASELINE cfg1 layer shape (F=39, d=U=16, H=2, L=3) on 6 samples
    rng = np.random.default_rng(20261018)
    B, F, D, U, H, L = 6, 39, 16, 16, 2, 3
    W = np.concatenate([_glorot(rng, D, U) for _ in range(4)], axis=1)
    b = (0.1 * rng.standard_normal(4 * U)).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(U)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(U)).astype(np.float32)
    x = rng.standard_normal((B, F, D)).astype(np.float32)
    dy = rng.standard_normal((B, F, U)).astype(np.float32)
    y = onp.interacting_fwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L)
    dx, dW, db, dg, dbt = onp.interacting_bwd(f64(x), f64(W), f64(b), f64(gamma), f64(beta), 1e-3, H, L, f64(dy))
    out["interacting_cfg1"] = dict(x=x, W=W, b=b, gamma=gamma, beta=beta, dy=dy, ln_eps=np.float64(1e-3),
                                   H=np.int64(H), L=np.int64(L), y=y, dx=dx, dW=dW, db=db, dgamma=dg, dbeta=dbt)
    # --- DIN A (din.py) and DIN B (staytime/layer.py), T=20, H=16
    rng = np.random.default_rng(20261019)
    B, T, Hd = 5, 20, 16
    q = rng.standard_normal((B, Hd)).astype(np.float32)
    keys = rng.standard_normal((B, T, Hd)).astype(np.float32)
    values = rng.standard_normal((B, T, Hd)).astype(np.float32)
    seq_len = np.array([20, 0, 7, 13, 1], np.int32)
    W1a, b1 = _glorot(rng, 3 * Hd, 16), (0.1 * rng.standard_normal(16)).astype(np.float32)
    W2, b2 = _glorot(rng, 16, 1), np.array([0.2], np.float32)
    dout = rng.standard_normal((B, Hd)).astype(np.float32)
    oa = onp.din_a_fwd(f64(q), f64(keys), f64(values), seq_len, f64(W1a), f64(b1), f64(W2), f64(b2))
    ga = onp.din_a_bwd(f64(q), f64(keys), f64(values), seq_len, f64(W1a), f64(b1), f64(W2), f64(b2), f64(dout))
    out["din_a"] = dict(q=q, keys=keys, values=values, seq_len=seq_len, W1=W1a, b1=b1, W2=W2, b2=b2, dout=dout,
                        out=oa, dq=ga[0], dkeys=ga[1], dvalues=ga[2], dW1=ga[3], db1=ga[4], dW2=ga[5], db2=ga[6])
    W1b = _glorot(rng, 4 * Hd, 16)
    mask = (np.arange(T)[None, :] < seq_len[:, None]).astype(np.uint8)
    ob = onp.din_b_fwd(f64(q), f64(keys), mask, f64(W1b), f64(b1), f64(W2), f64(b2))
    gb = onp.din_b_bwd(f64(q), f64(keys), mask, f64(W1b), f64(b1), f64(W2), f64(b2), f64(dout))
    out["din_b"] = dict(q=q, facts=keys, mask=mask, W1=W1b, b1=b1, W2=W2, b2=b2, dout=dout,
                        out=ob, dq=gb[0], dfacts=gb[1], dW1=gb[2], db1=gb[3], dW2=gb[4], db2=gb[5])
    # --- embedding gather + routing + sparse Adam, 4 fields
    rng = np.random.default_rng(20261020)
    Fe, d, Be = 4, 16, 12
    rows = np.array([50, 7, 1000, 3], np.int64)
    base = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table = (0.1 * rng.standard_normal((int(rows.sum()), d))).astype(np.float32)
    ids = rng.integers(0, 2 ** 40, size=(Be, Fe)).astype(np.int64)
    ids[0, 1] = -1
    emb, r = onp.embed_gather(table, ids, rows, base)
    grad = rng.standard_normal((Be * Fe, d)).astype(np.float32)
    _, _, corr = onp.adam_scalars(1, 0.9, 0.999)
    w2, m2, v2 = onp.sparse_adam(f64(table), np.zeros_like(f64(table)), np.zeros_like(f64(table)), r.reshape(-1),
                                 grad, 1e-2, 0.9, 0.999, 1e-8, float(corr))
    world = 4
    per = (rows + world - 1) // world
    lbase = np.concatenate([[0], np.cumsum(per)[:-1]]).astype(np.int64)
    sr, inv, cnt, off = onp.route_ids(ids, Fe, rows, lbase, world)
    out["embedding"] = dict(table=table, ids=ids, rows=rows, row_base=base, emb=emb, arena_rows=r, grad=grad,
                            adam_w=w2, adam_m=m2, adam_v=v2, world=np.int64(world), local_base=lbase,
                            send_rows=sr, inverse=inv, send_counts=cnt, send_offsets=off)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, arrs in generate().items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
