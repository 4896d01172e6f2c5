"""World-size-2 gloo tests (CPU) of the sharded-embedding exchange protocol: routing into fixed-capacity
buckets, all-to-all of ids / rows / gradients through recommendsystem_b200.sharded.Exchange, and the
owner-side segment sums.  The per-rank compute (routing, lookup, segment sum) is done by the oracle here —
the CUDA kernels that do it in the product are parity-tested on the GPU (tests/test_gpu_embed.py,
tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_np as onp
        from recommendsystem_b200.sharded import Exchange, bucket_capacity, shard_layout
        ex = Exchange()
        F, d, b = 5, 8, 64
        rng = np.random.default_rng(1234)                       # same stream on all ranks: global state
        rows = np.array([37, 100, 3, 1000, 64], np.int64)
        gbase = np.concatenate([[0], np.cumsum(rows)[:-1]])
        table = rng.standard_normal((int(rows.sum()), d)).astype(np.float32)
        ids_all = rng.integers(0, 10 ** 9, size=(world, b, F)).astype(np.int64)
        ids_all[:, 0, 0] = -1                                   # a padding id on every rank
        grads_all = rng.standard_normal((world, b * F, d)).astype(np.float32)
        lrows, lbase = shard_layout(rows, world)
        # my shard: global rows rank, rank+W, ... of every field
        shard = np.zeros((int(lrows.sum()), d), np.float32)
        for f in range(F):
            src = table[gbase[f] + rank: gbase[f] + rows[f]: world]
            shard[lbase[f]: lbase[f] + len(src)] = src
        ids = ids_all[rank]
        cap = bucket_capacity(b * F, world)
        send, inv, cnt, ovf = onp.route_ids_padded(ids, F, rows, lbase, world, cap)
        assert ovf == 0 and (inv[ids.reshape(-1) >= 0] >= 0).all()
        recv = ex.all_to_all(torch.empty(world * cap, dtype=torch.int32), torch.from_numpy(send)).numpy()
        out_rows, _ = onp.embed_gather_rows(shard, recv)
        got = ex.all_to_all(torch.empty(world * cap, d), torch.from_numpy(out_rows)).numpy()
        X = np.where((inv >= 0)[:, None], got[np.maximum(inv, 0)], 0)
        ref, _ = onp.embed_gather(table, ids, rows, gbase)
        assert np.array_equal(X.reshape(b, F, d), ref), "sharded lookup != single-table lookup"
        # gradients: scatter into the slots, exchange, owner-side sorted segment sum
        gsend = np.zeros((world * cap, d), np.float32)
        gsend[inv[inv >= 0]] = grads_all[rank][inv >= 0]
        grecv = ex.all_to_all(torch.empty(world * cap, d), torch.from_numpy(gsend)).numpy()
        uniq, sums = onp.segment_sum_sorted(recv, grecv)
        # reference: global segment sum over all ranks' lookups in (rank, lookup) order, my rows only
        _, grow = onp.embed_gather(table, ids_all.reshape(world * b, F), rows, gbase)
        guniq, gsums = onp.segment_sum_sorted(grow.reshape(-1), grads_all.reshape(-1, d))
        mine = {}
        for r, s in zip(guniq, gsums):
            f = int(np.searchsorted(gbase, r, side="right") - 1)
            rl = r - gbase[f]
            if rl % world == rank:
                mine[int(lbase[f] + rl // world)] = s
        assert sorted(mine) == [int(u) for u in uniq]
        for u, s in zip(uniq, sums):
            assert np.array_equal(s, mine[int(u)]), "owner-side segment sum differs (order must be rank-major)"
        # dense gradient averaging
        t = torch.full((7,), float(rank + 1))
        ex.all_reduce_mean(t)
        assert torch.allclose(t, torch.full((7,), (world + 1) / 2.0))
        q.put((rank, "ok"))
    except Exception as e:                                       # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_exchange_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"


def test_bucket_capacity_and_layout():
    from recommendsystem_b200.sharded import bucket_capacity, shard_layout
    assert bucket_capacity(319488, 8) % 128 == 0 and bucket_capacity(319488, 8) >= 319488 // 8 + 1000
    assert bucket_capacity(100, 1) == 100
    lr, lb = shard_layout([10, 7, 1], 4)
    assert lr.tolist() == [3, 2, 1] and lb.tolist() == [0, 3, 5]


def test_route_padded_oracle_overflow_flag():
    from oracle import oracle_np as onp
    ids = np.zeros((50, 1), np.int64)          # every lookup goes to owner 0
    send, inv, cnt, ovf = onp.route_ids_padded(ids, 1, [100], [0], 2, 16)
    assert ovf == 1 and (inv >= 0).sum() == 16 and cnt.tolist() == [50, 0]


def _metric_state_np(y, p, T=200):
    """The accumulator rs_binary_metrics_update builds, restated with the oracle (int64 words + 2 double sums)."""
    from oracle import oracle_metrics as om
    thr = om.keras_thresholds(T)
    k = (np.asarray(p, np.float32)[:, None] > thr[None, :]).sum(1)
    pos = np.asarray(y) > 0.5
    st = np.zeros(2 * T + 6, np.int64)
    st[:T + 1] = np.bincount(k[pos], minlength=T + 1)
    st[T + 1:2 * T + 2] = np.bincount(k[~pos], minlength=T + 1)
    st[2 * T + 2] = len(y)
    st[2 * T + 3] = int(((np.asarray(p, np.float32) > 0.5) == pos).sum())
    st[2 * T + 4:] = np.array([np.sum(y, dtype=np.float64), np.sum(np.asarray(p, np.float64))]).view(np.int64)
    return st


def _metrics_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_metrics as om
        from recommendsystem_b200.api.metrics import merge_metric_states
        rng = np.random.default_rng(77)
        y = (rng.random(4000) < 0.3).astype(np.float32)
        p = rng.random(4000).astype(np.float32)
        lo, hi = rank * 2000, (rank + 1) * 2000
        st = torch.from_numpy(_metric_state_np(y[lo:hi], p[lo:hi]))
        merge_metric_states(st, 200)
        full = _metric_state_np(y, p)
        T = 200
        assert np.array_equal(st.numpy()[:2 * T + 4], full[:2 * T + 4]), "integer words"
        s, f = st.numpy()[2 * T + 4:].view(np.float64), full[2 * T + 4:].view(np.float64)
        assert np.allclose(s, f, rtol=1e-14), "double sums"
        # the merged histogram gives the AUC of the union
        ph, nh = st.numpy()[:T + 1], st.numpy()[T + 1:2 * T + 2]
        tp = np.cumsum(ph[::-1])[::-1][1:].astype(np.float64)
        fp = np.cumsum(nh[::-1])[::-1][1:].astype(np.float64)
        auc = om.keras_auc_from_counts(tp, fp, nh.sum() - fp, ph.sum() - tp)
        assert abs(auc - om.keras_auc(y, p)) < 1e-12
        q.put((rank, "ok"))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_metric_states_merge_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metrics_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(30)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}:\n{msg}"
