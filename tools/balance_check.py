#!/usr/bin/env python
"""Per-CTA busy cycles vs the cost model used for the static schedule: python tools/balance_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.engine import Engine
from drl_urban_planning_b200.packing import pack_states
dev = torch.device("cuda", 0)
count = 256
states, actions = synth.make_states(111, "hlg", count)
blob = pack_states(states).to(dev)
eng = Engine(dev, blob.n_cap, blob.e_cap)
t = lambda x: torch.as_tensor(x, device=dev)
adv, ret, exps = synth.make_ppo_targets(1, count)
fixed = np.full((count, 1), -4.0, np.float32)
params = t(PL.default_init(1))
info = blob.info.astype(np.int64)
cost = Engine.graph_cost(info)
ids_np = eng.balance_ids(np.arange(count), cost)
ids = t(ids_np.astype(np.int32))
stamps = torch.zeros(384, dtype=torch.int64, device=dev)
args = (blob, params, t(actions), t(adv), t(ret), t(fixed), t(exps), 1.0 / count, 1.0 / count)
for _ in range(3): eng.ppo_grad(*args, ids=ids)
eng.set_stamp_buffer(stamps); eng.ppo_grad(*args, ids=ids); torch.cuda.synchronize()
st = stamps.cpu().numpy(); g = eng.grid
busy = st[64:64 + g] - st[224:224 + g]
pred = np.zeros(g); ng = np.zeros(g, int)
for i, gid in enumerate(ids_np): pred[i % g] += cost[gid]; ng[i % g] += 1
A = np.stack([info[ids_np[:g], 1], info[ids_np[:g], 0], info[ids_np[:g], 2], np.ones(g)], 1).astype(float)
single = ng == 1
coef, *_ = np.linalg.lstsq(A[single], busy[single], rcond=None)
print("fit on single-graph CTAs: cycles = %.1f e + %.1f n + %.1f k + %.0f" % tuple(coef))
# all CTAs: features summed over the CTA's graphs (e, n, k of land-use graphs, k of road graphs, graph count)
F = np.zeros((g, 5))
for i, gid in enumerate(ids_np):
    n_, e_, k_, st_ = info[gid]
    F[i % g] += [e_, n_, k_ if st_ == 0 else 0, k_ if st_ == 1 else 0, 1]
coef2, *_ = np.linalg.lstsq(F, busy.astype(float), rcond=None)
res = busy - F @ coef2
print("fit on all CTAs: cycles = %.1f e + %.1f n + %.1f k_landuse + %.1f k_road + %.0f per graph;  residual rms %.0f (%.1f%% of mean)"
      % (*coef2, np.sqrt((res ** 2).mean()), 100 * np.sqrt((res ** 2).mean()) / busy.mean()))
# what LPT would achieve with the refitted model (true cost taken as the fitted per-graph cost)
true_cost = coef2[0] * info[:, 1] + coef2[1] * info[:, 0] + coef2[2] * info[:, 2] * (info[:, 3] == 0) + coef2[3] * info[:, 2] * (info[:, 3] == 1) + coef2[4]
ids2 = eng.balance_ids(np.arange(count), true_cost)
load = np.zeros(g)
for i, gid in enumerate(ids2): load[i % g] += true_cost[gid]
print("refitted model + LPT: predicted max %.0f mean %.0f (balance %.3f); current: measured max %d mean %.0f (balance %.3f)"
      % (load.max(), load.mean(), load.mean() / load.max(), busy.max(), busy.mean(), busy.mean() / busy.max()))
print("busy max %d mean %.0f; predicted max %.0f mean %.0f; corr %.3f" % (busy.max(), busy.mean(), pred.max(), pred.mean(), np.corrcoef(busy, pred)[0, 1]))
top = np.argsort(-busy)[:5]
for c in top:
    gs = [ids_np[c + r * g] for r in range(ng[c])]
    print("CTA", c, "busy", busy[c], "pred", pred[c], [tuple(info[x][:3]) for x in gs])
