import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import oracle_np as onp
from recommendsystem_b200.autoint import AutoIntConfig, AutoIntTrainer
from util import rel_err
f64 = lambda a: np.asarray(a, np.float64)
dev = torch.device("cuda:0")
rng = np.random.default_rng(7)
B, F, d, H, L, hidden = 256, 39, 16, 2, 3, (256, 128)
for dt in ["f32", "bf16"]:
    cfg = AutoIntConfig(num_fields=F, rows_per_field=200, embed_dim=d, unit_num=d, head_num=H, layer_num=L,
                        mlp_hidden=hidden, batch=B, dtype=dt, lr_dense=1e-3, lr_sparse=1e-2)
    tr = AutoIntTrainer(cfg, dev)
    P0 = tr.dense_state()
    table0 = tr.table.cpu().numpy().copy()
    ids = rng.integers(0, 2 ** 40, size=(B, F)).astype(np.int64)
    y = (rng.random((B, 1)) < 0.25).astype(np.float32)
    loss = tr.step(torch.from_numpy(ids).to(dev), torch.from_numpy(y).to(dev))
    torch.cuda.synchronize()
    X, rows = onp.embed_gather(table0, ids, tr.rows_host, tr.base_host)
    n = len(hidden)
    P = dict(Wqkvr=f64(P0["Wqkvr"]), bqkvr=f64(P0["bqkvr"]), gamma=f64(P0["gamma"]), beta=f64(P0["beta"]),
             mlp_W=[f64(P0[f"mlp_W{i}"]) for i in range(n)], mlp_b=[f64(P0[f"mlp_b{i}"]) for i in range(n)],
             out_W=f64(P0["out_W"]), out_b=f64(P0["out_b"]))
    res = onp.autoint_fwd_bwd(f64(X), P, f64(y), H, L, cfg.ln_eps)
    # recompute pieces
    A = res["A"]
    print(dt, "loss", float(loss), res["loss"])
    print(" A", rel_err(tr.Z[:, tr.n_deep:].float().cpu().numpy().reshape(B, F, d), A))
    print(" p", rel_err(tr.p_raw.float().cpu().numpy(), res["p_raw"]))
    print(" dX", rel_err(tr.dX.float().cpu().numpy(), res["dX"]))
    # split dX into parts via oracle
    dXi, *_ = onp.interacting_bwd(f64(X), P["Wqkvr"], P["bqkvr"], P["gamma"], P["beta"], cfg.ln_eps, H, L,
                                  f64(tr.A.float().cpu().numpy()))   # tr.A holds dA after the step
    print(" |dXi|max", np.abs(dXi).max(), "|dX|max", np.abs(res["dX"]).max(), "|dX - dXi|max", np.abs(res["dX"]-dXi).max())
    G = {k: v.cpu().numpy() for k, v in tr.G.items()}
    for k in ["Wqkvr", "bqkvr", "gamma", "beta", "out_W", "out_b"]:
        print(" d" + k, rel_err(G[k], res["grads"][k]))
    for i in range(n):
        print(f" dmlp_W{i}", rel_err(G[f"mlp_W{i}"], res["grads"]["mlp_W"][i]), f"dmlp_b{i}", rel_err(G[f"mlp_b{i}"], res["grads"]["mlp_b"][i]))
    print(" dH0 max", tr.dH[0].float().abs().max().item(), "dZ max", tr.dZ.float().abs().max().item())
