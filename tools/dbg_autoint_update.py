"""Developer tool (1 GPU): one AUTOINT train step on the 512-sample batch of the W = 8 parity worker; the sparse
Adam update of the unsharded EmbeddingFeatures against the numpy oracle on the captured embedding gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle_np as onp
from recommendsystem_b200.api.builders import AUTOINT, AUTOINT_LABELS
from recommendsystem_b200.api.optim import DenseAdam
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(17)
world, b = 8, 64
slots = [str(3000 + i) for i in range(13)]
ids_all = {s: torch.randint(0, 10 ** 9, (world * b,), generator=g) for s in slots}
ids_all[slots[2]][5] = -1
y_all = (torch.rand(world * b, 7, generator=g) < 0.3).float()
m = AUTOINT(slots, [], False, dnn_hidden_units=(32, 16), bucket_size=997, device=str(dev), seed=3).model
torch.manual_seed(11)
inp = {s: v.to(dev) for s, v in ids_all.items()}
lab = {k: y_all[:, i:i + 1].to(dev) for i, k in enumerate(AUTOINT_LABELS)}
m.predict(inp)
m.emb.opt.learning_rate = 1e-2
m.opt = DenseAdam(m.sub_model.parameters(), lr=1e-3)
cap = {}
orig = m.emb.backward
m.emb.backward = lambda gr: (cap.update({k: v.clone() for k, v in gr.items()}), orig(gr))[1]
for step in range(3):
    t0 = m.emb.arena.clone().cpu().numpy()
    m.train_step(inp, lab)
    torch.cuda.synchronize()
    G = torch.stack([cap[s] for s in slots], 1).reshape(-1, 8).cpu().numpy()
    ids = torch.stack([ids_all[s] for s in slots], 1)
    rows = torch.where(ids >= 0, ids % 997 + torch.arange(13)[None, :] * 997, torch.tensor(-1)).reshape(-1).numpy()
    sc = m.emb.scalars.cpu().numpy()
    t = step + 1
    corr = np.sqrt(1 - np.float32(0.999) ** t) / (1 - np.float32(0.9) ** t)
    w, mm, vv = onp.sparse_adam(t0[:, 0], t0[:, 1], t0[:, 2], rows, G, 1e-2, 0.9, 0.999, 1e-8, corr)
    got = m.emb.arena.cpu().numpy()
    err = np.abs(got[:, 0] - w).max(1)
    bad = np.nonzero(err > 1e-6)[0]
    print(f"step {t}: max |table - oracle| = {err.max():.3e}; rows off: {bad[:10].tolist()}; scalars {sc}")
    for r in (7487, 12127, 12261, 12651):
        i = np.nonzero(rows == r)[0]
        print("   row", r, "occurrences", i.tolist(), "grad", G[i[0]][:4], "w0", t0[r, 0, :2], "got", got[r, 0, :2], "oracle", w[r, :2])
