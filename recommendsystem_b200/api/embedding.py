"""Sparse embedding + sparse optimizer surface the reference reaches through `tensornet` (tn):
    tn.feature_column.category_column / embedding_column, tn.layers.EmbeddingFeatures,
    tn.core.Adam / tn.core.AdaGrad
(call sites staytime/VideoDnn.py:217-244, rough_rank/model.py:89-115, rank/ctr/base_model.py:203-217).
TensorNet itself is not vendored by the reference: table storage, the `bucket_size` semantics (ids
are reduced mod bucket_size here) and the exact update rules are this repo's documented choices
(DESIGN.md).  All tables of one EmbeddingFeatures layer live in ONE fp32 arena so a
batch is served by a single gather launch and a single sorted-segment update launch."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import cabi, ops


@dataclass
class CategoryColumn:
    key: str
    bucket_size: int


@dataclass
class EmbeddingColumn:
    categorical_column: CategoryColumn
    dimension: int
    combiner: Optional[str] = "mean"
    seq_max_len: Optional[int] = None
    name: Optional[str] = None

    @property
    def key(self):
        return self.name or self.categorical_column.key


def category_column(key, bucket_size):
    return CategoryColumn(str(key), int(bucket_size))


def embedding_column(categorical_column, dimension, combiner="mean", seq_max_len=None, name=None):
    return EmbeddingColumn(categorical_column, int(dimension), combiner, seq_max_len, name)


class Adam:
    """tn.core.Adam(learning_rate, beta1, beta2, epsilon) — sparse rows touched by the batch."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon


class AdaGrad:
    """tn.core.AdaGrad(learning_rate, initial_g2sum, initial_scale[, feature_drop_show]).
    One g2sum scalar per row (TensorNet style); rows are initialised N(0,1) * initial_scale."""

    def __init__(self, learning_rate=0.01, initial_g2sum=0, initial_scale=1, epsilon=1e-8, feature_drop_show=-1,
                 per_element=False):
        self.learning_rate, self.initial_g2sum, self.initial_scale = learning_rate, initial_g2sum, initial_scale
        self.epsilon, self.per_element = epsilon, per_element


class EmbeddingFeatures:
    """EmbeddingFeatures(embedding_columns, sparse_opt, name)(inputs: dict[key -> ids]) ->
    dict[column key -> [B, d]]  (combiner='mean', dense [B] or [B, bag] ids; id < 0 = padding) or
    ([B, T, d], mask [B, T]) for `combiner=None, seq_max_len=T` sequence columns.

    `backward(grads)` applies the fused sorted-segment sparse optimizer to the rows the last
    forward touched (the reference's server-side push); `grads` maps column key -> d(loss)/d(output)."""

    def __init__(self, embedding_columns: List[EmbeddingColumn], sparse_opt, name="embedding", device="cuda:0",
                 out_dtype=torch.float32, seed=0):
        self.cols = list(embedding_columns)
        self.opt = sparse_opt
        self.dev = torch.device(device)
        dims = {c.dimension for c in self.cols}
        if len(dims) != 1:
            raise ValueError("one EmbeddingFeatures layer holds columns of a single dimension (one arena)")
        self.d = dims.pop()
        self.out_dtype = out_dtype
        rows = np.asarray([c.categorical_column.bucket_size for c in self.cols], np.int64)
        self.base = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
        self.rows = rows
        if not isinstance(sparse_opt, (Adam, AdaGrad)):
            raise TypeError("sparse_opt must be api.embedding.Adam or AdaGrad")
        total = self._alloc_tables(seed)
        self.row_bits = ops.row_bits(total)
        self._last = None
        self._geom_cache = {}

    def _init_scale(self):
        return float(getattr(self.opt, "initial_scale", 0.1) if isinstance(self.opt, AdaGrad) else 0.1)

    def _alloc_state(self, n_rows):
        """Storage for `n_rows` table rows + optimizer state: Adam keeps one [w | m | v] record per row (DESIGN.md 2:
        the sparse update touches ONE contiguous 12d-byte block per row instead of three far-apart ones)."""
        if isinstance(self.opt, Adam):
            self.arena = torch.zeros(n_rows, 3, self.d, device=self.dev)
            self.table, self.m, self.v = self.arena[:, 0, :], self.arena[:, 1, :], self.arena[:, 2, :]
            self.table_ld = 3 * self.d
            self.scalars = torch.zeros(4, device=self.dev)
        else:
            self.table = torch.zeros(n_rows, self.d, device=self.dev)
            self.table_ld = self.d
            shape = (n_rows, self.d) if self.opt.per_element else (n_rows,)
            self.g2sum = torch.full(shape, float(self.opt.initial_g2sum), device=self.dev)

    def _alloc_tables(self, seed) -> int:
        """Allocate and initialise (N(0, scale), one generator stream over the rows in arena order); returns the number
        of rows the sort keys must address."""
        total = int(self.rows.sum())
        self._alloc_state(total)
        gen = torch.Generator(device=self.dev).manual_seed(seed)
        chunk = 1 << 22
        for r0 in range(0, total, chunk):
            r1 = min(total, r0 + chunk)
            self.table[r0:r1] = torch.empty(r1 - r0, self.d, device=self.dev).normal_(0.0, self._init_scale(), generator=gen)
        return total

    def __call__(self, inputs: Dict[str, torch.Tensor]):
        out, plan = {}, []
        # single-valued mean columns (ids [B] / [B,1], the Criteo-shaped case) share ONE gather launch over
        # the stacked ids [B, F'] (per-column row_base / bucket size) and one sorted-segment push
        single = [ci for ci, c in enumerate(self.cols) if c.combiner is not None and
                  not isinstance(inputs[c.categorical_column.key], (tuple, list)) and
                  (inputs[c.categorical_column.key].dim() == 1 or inputs[c.categorical_column.key].shape[1] == 1)]
        if len(single) > 1:
            ids = torch.stack([inputs[self.cols[ci].categorical_column.key].reshape(-1) for ci in single], dim=1)
            ids = ids.to(self.dev, torch.int64).contiguous()
            base, rows = self._geom(tuple(single))
            emb, keys, _ = ops.embed_gather(self.table, ids, base, rows, torch.float32, want_keys=True)
            for j, ci in enumerate(single):
                out[self.cols[ci].key] = emb[:, j, :].to(self.out_dtype)
            plan.append(([self.cols[ci].key for ci in single], keys, 1.0, None))
            # the gather output as it lies: (column keys, [B, F', d]) for consumers that work on the stacked tensor
            self.last_stacked = ([self.cols[ci].key for ci in single], emb.to(self.out_dtype))
        else:
            single = []
            self.last_stacked = None
        for ci, c in enumerate(self.cols):
            if ci in single:
                continue
            raw = inputs[c.categorical_column.key]
            base, rows = self._geom((ci,))
            if isinstance(raw, (tuple, list)):
                # variable-length bag in CSR form (values int64 [nnz], offsets int64 [B+1]): the VarLenFeature /
                # SparseTensor input of the reference (staytime/parse.py:22-23), combiner='mean' (VideoDnn.py:224-226)
                if c.combiner is None:
                    raise ValueError("CSR (values, offsets) inputs are for combiner='mean' columns")
                vals = raw[0].to(self.dev, torch.int64).contiguous()
                offs = raw[1].to(self.dev, torch.int64).contiguous()
                nb = offs.numel() - 1
                emb = torch.empty(nb, self.d, dtype=torch.float32, device=self.dev)
                keys = torch.empty(vals.numel(), dtype=torch.int64, device=self.dev)
                inv = torch.empty(nb, dtype=torch.float32, device=self.dev)
                cabi.call("rs_embed_bag_fwd_ld", self.table.data_ptr(), self.table_ld, vals.data_ptr(), offs.data_ptr(),
                          base.data_ptr(), rows.data_ptr(), nb, 1, self.d, emb.data_ptr(), cabi.RS_F32, keys.data_ptr(),
                          inv.data_ptr(), ops._stream())
                out[c.key] = emb.to(self.out_dtype)
                plan.append((c.key, keys, None, ("csr", offs, inv)))
                continue
            ids = raw.to(self.dev, torch.int64)
            if c.combiner is None:                               # sequence column -> ([B,T,d], mask)
                T = c.seq_max_len or ids.shape[1]
                seq = ids[:, :T].contiguous()
                emb, keys, arows = ops.embed_gather(self.table, seq.reshape(-1, 1), base, rows, self.out_dtype,
                                                    want_keys=True, want_rows=True)
                out[c.key] = (emb.view(seq.shape[0], T, self.d), (arows.view(seq.shape[0], T) >= 0))
                plan.append((c.key, keys, 1.0, None))
            else:
                if ids.dim() == 1:
                    ids = ids[:, None]
                B, bag = ids.shape
                emb, keys, arows = ops.embed_gather(self.table, ids.reshape(-1, 1).contiguous(), base, rows,
                                                    torch.float32, want_keys=True, want_rows=True)
                if bag == 1:
                    out[c.key] = emb.view(B, self.d).to(self.out_dtype)
                    plan.append((c.key, keys, 1.0, None))
                else:                                            # combiner='mean' over the valid ids of the bag
                    valid = (arows.view(B, bag) >= 0)
                    cnt = valid.sum(1).clamp(min=1).to(torch.float32)
                    out[c.key] = (emb.view(B, bag, self.d).sum(1) / cnt[:, None]).to(self.out_dtype)
                    plan.append((c.key, keys, None, (bag, cnt)))
        self._last = plan
        return out

    # ---- state one train step changes (GraphedTrainStep undoes its warm-up steps with these) -----------
    def touched_rows(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Sorted unique arena rows a step on `inputs` reads and updates."""
        rows = []
        for ci, c in enumerate(self.cols):
            ids = inputs[c.categorical_column.key]
            if isinstance(ids, (tuple, list)):                   # CSR bag: (values, offsets)
                ids = ids[0]
            ids = ids.to(self.dev, torch.int64).reshape(-1)
            if c.combiner is None and c.seq_max_len:
                ids = inputs[c.categorical_column.key].to(self.dev, torch.int64)[:, :c.seq_max_len].reshape(-1)
            ids = ids[ids >= 0]
            rows.append(int(self.base[ci]) + torch.remainder(ids, int(self.rows[ci])))
        return torch.unique(torch.cat(rows)) if rows else torch.empty(0, dtype=torch.int64, device=self.dev)

    def snapshot(self, inputs):
        rows = self.touched_rows(inputs)
        if isinstance(self.opt, Adam):
            return {"rows": rows, "arena": self.arena[rows].clone(), "scalars": self.scalars.clone()}
        return {"rows": rows, "table": self.table[rows].clone(), "g2sum": self.g2sum[rows].clone()}

    def restore(self, snap):
        rows = snap["rows"]
        if isinstance(self.opt, Adam):
            self.arena[rows] = snap["arena"]
            self.scalars.copy_(snap["scalars"])
        else:
            self.table[rows] = snap["table"]
            self.g2sum[rows] = snap["g2sum"]

    def _geom(self, cols):
        """(row_base, rows) device tensors of a column group, created once (no host->device copy per call:
        the step stays CUDA-graph capturable)."""
        g = self._geom_cache.get(cols)
        if g is None:
            idx = list(cols)
            g = (torch.as_tensor(np.asarray(self.base)[idx], device=self.dev),
                 torch.as_tensor(np.asarray(self.rows)[idx], device=self.dev))
            self._geom_cache[cols] = g
        return g

    def backward(self, grads: Dict[str, torch.Tensor]):
        """Push d(loss)/d(output) of every column: segment-sum per touched row + optimizer update."""
        if self._last is None:
            raise RuntimeError("EmbeddingFeatures.backward called before a forward")
        if isinstance(self.opt, Adam):
            ops.adam_advance(self.scalars, self.opt.beta1, self.opt.beta2)
        for key, keys, scale, bag in self._last:
            if isinstance(key, list):                             # stacked single-valued columns: [B, F', d]
                g = torch.stack([grads[k] for k in key], dim=1)
            else:
                g = grads[key]
            if isinstance(g, tuple):
                g = g[0]
            g = g.reshape(-1, self.d)
            if bag is not None and bag[0] == "csr":               # CSR bag: one gradient row per occurrence
                _, offs, inv = bag
                gc = g.contiguous()
                g = torch.empty(keys.numel(), self.d, dtype=torch.float32, device=self.dev)
                cabi.call("rs_embed_bag_grad", gc.data_ptr(), ops._dt(gc), offs.data_ptr(), inv.data_ptr(),
                          offs.numel() - 1, self.d, g.data_ptr(), ops._stream())
            elif bag is not None:                                 # mean combiner: 1/count per occurrence
                b, cnt = bag
                g = (g.float() / cnt[:, None]).repeat_interleave(b, dim=0)
            g = g.contiguous()
            ks = ops.sort_keys(keys, self.row_bits)
            if isinstance(self.opt, Adam):
                ops.segsum_adam(self.table, self.m, self.v, g, ks, self.opt.learning_rate, self.opt.beta1,
                                self.opt.beta2, self.opt.epsilon, self.scalars)
            else:
                ops.segsum_adagrad(self.table, self.g2sum, g, ks, self.opt.learning_rate, self.opt.epsilon,
                                   self.opt.per_element)
        self._last = None
