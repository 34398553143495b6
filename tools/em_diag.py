#!/usr/bin/env python
"""Where does update_GMM deviate from the fp64 oracle at the headline shape?  (GPU diagnostic, not a test.)
Prints per-stage errors: the fused gradient of one class, one zero-gradient replay, then per-class errors of a
whole update_GMM call keyed by the class's role (active / inactive) and its order among the active classes."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import headline_case as HC                      # noqa: E402
from oracle import mgproto_oracle as O          # noqa: E402
import mgproto_b200 as M                        # noqa: E402
from mgproto_b200 import ops                    # noqa: E402

dev = torch.device("cuda:0")
C, K, D, T, cap = 200, 10, 128, 20, 800
t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)   # noqa: E731
nw = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())              # noqa: E731

mu, sg, wt = HC.mixture(C, K, D)
rows = HC.bank_rows(C, K, D, cap, mu)
am, av, flags, short, step0 = HC.em_state(C, K, D)

# (1) gradient of one class from the statistics kernels vs oracle
c = 17
bank = t(rows)
order = torch.full((C,), -1, dtype=torch.int32, device=dev)
order[c] = 0
sched = torch.tensor([1, 0], dtype=torch.int32, device=dev)
for n_split in (1, 2):
    stats = torch.empty(C, n_split, ops.em_stat_stride(K, D), device=dev)
    ops.em_stats(bank, order, t(mu), t(sg), t(wt), 0.1, stats, n_split)
    grad = torch.zeros(C, K, D, device=dev)
    w2 = t(wt).clone()
    ops.em_update(stats, n_split, cap, order, sched, t(mu), t(sg), w2, None, None, 0, 3, 1, 0.0, 0.9, 0.999, 1e-8, 0.99,
                  grad_out=grad)
    x64 = rows[c].astype(np.float64)
    pi_old = wt[c, c * K:(c + 1) * K].astype(np.float64)
    _, lr = O.e_step(x64, mu[c].astype(np.float64), sg[c].astype(np.float64), pi_old)
    pin, g, _ = O.m_step_diversified_grad(x64, lr, mu[c].astype(np.float64), sg[c].astype(np.float64), pi_old, 0.1)
    r = O.smooth_resp(lr, 0.1)
    s0 = stats[c].sum(0)[:K].cpu().numpy()
    s1 = stats[c].sum(0)[K:K + K * D].cpu().numpy().reshape(K, D)
    print("n_split %d: grad normwise %.2e | S0 %.2e | S1 %.2e | grad max %.3e" % (
        n_split, nw(grad[c].cpu().numpy(), g), nw(s0, r.sum(0)), nw(s1, r.T @ x64), np.abs(g).max()))

# (2) whole update_GMM, per-class error vs fp64 oracle
net = M.MGProto(features=nn.Sequential(nn.Conv2d(3, 8, 1)), img_size=224, prototype_shape=(C * K, D, 1, 1),
                proto_layer_rf_info=None, num_classes=C, add_on_layers_type="regular", sz_embedding=8,
                mem_capacity=cap, mine_K=T).to(dev)
net.prototype_means.data.copy_(t(mu)); net.prototype_covs.data.copy_(t(sg)); net.last_layer.weight.data.copy_(t(wt))
net.prototype_optimizer = torch.optim.Adam([{"params": net.prototype_means, "lr": 3e-3}])
net.queue.bank.copy_(t(rows)); net.queue.mem_len.fill_(cap)
net.prototype_optimizer.state[net.prototype_means] = {"step": torch.tensor(float(step0)), "exp_avg": t(am).clone(),
                                                      "exp_avg_sq": t(av).clone()}
f = flags[0].copy()
f[short] = False
net.queue.updated.copy_(t(f, torch.uint8))
net.update_GMM()
net.sync_optimizer_state()
ob = O.MemoryBankOracle(C, D, cap, dtype=np.float64)
ob.data[:] = rows; ob.mem_len[:] = cap
ad = O.AdamOracle((C, K, D), lr=3e-3)
ad.m, ad.v, ad.t = am.astype(np.float64), av.astype(np.float64), step0
mu_ref, wt_ref, _ = O.update_gmm(ob, f, mu.astype(np.float64), sg.astype(np.float64), wt.astype(np.float64), ad)
got = net.prototype_means.detach().cpu().numpy().astype(np.float64)
err = np.abs(got - mu_ref).reshape(C, -1).max(1) / np.abs(mu_ref).max()
mv = np.abs(mu_ref - mu).reshape(C, -1).max(1)
act = np.nonzero(f)[0]
ordr = {int(cc): i for i, cc in enumerate(act)}
print("overall normwise %.2e; inactive classes max %.2e; active max %.2e" % (err.max(), err[~f].max(), err[f].max()))
worst = np.argsort(-err)[:12]
for cc in worst:
    print("class %3d active=%d order=%s err %.2e movement %.2e" % (cc, f[cc], ordr.get(int(cc), "-"), err[cc], mv[cc]))
st = net.prototype_optimizer.state[net.prototype_means]
print("adam m %.2e v %.2e" % (nw(st["exp_avg"].cpu().numpy(), ad.m), nw(st["exp_avg_sq"].cpu().numpy(), ad.v)))
em = np.abs(st["exp_avg"].cpu().numpy() - ad.m).reshape(C, -1).max(1) / np.abs(ad.m).max()
print("adam m err: inactive max %.2e active max %.2e" % (em[~f].max(), em[f].max()))
# replay only: no active class at all is impossible (n_active = 0 -> nothing happens); isolate via first-order classes
lo = [cc for cc in act if ordr[int(cc)] < 3]
hi = [cc for cc in act if ordr[int(cc)] > len(act) - 4]
print("first-ordered active classes err", [float("%.2e" % err[cc]) for cc in lo], "last-ordered", [float("%.2e" % err[cc]) for cc in hi])
