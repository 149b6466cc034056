#!/usr/bin/env python
"""Cycle budget of the fused kernel's phases for one HLG graph (CTA 0, first graph): python tools/phase_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.engine import Engine
from drl_urban_planning_b200.packing import pack_states

NAMES = {1: "setup (lists, VN weights)", 2: "h0 + hc + numeric L0", 3: "epq L0 (+num L1, Weff)", 4: "pull fwd L0", 5: "epq L1",
         6: "pull fwd L1 + means", 7: "attention fwd", 8: "  warp 0: value tail done", 9: "value head || policy head (barrier)", 20: "  warp 0: softmax done", 11: "  warp 0: value/num bwd done", 10: "softmax + seeds + g_z",
         12: "head bwd A || value/num bwd", 13: "head bwd B/C", 14: "head grads + attention bwd", 15: "pull bwd L1", 16: "gW/g_h L1",
         17: "epq L0 (recompute)", 18: "pull bwd L0", 19: "gW/g_h L0", 21: "node encoder bwd"}
dev = torch.device("cuda", 0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
states, actions = synth.make_states(111, "hlg", count)
blob = pack_states(states).to(dev)
eng = Engine(dev, blob.n_cap, blob.e_cap)
t = lambda x: torch.as_tensor(x, device=dev)
adv, ret, exps = synth.make_ppo_targets(1, count)
fixed = np.full((count, 1), -4.0, np.float32)
params = t(PL.default_init(1))
stamps = torch.zeros(384, dtype=torch.int64, device=dev)
info = blob.info.astype(np.int64)
ids = t(eng.balance_ids(np.arange(count), Engine.graph_cost(info)).astype(np.int32))
args = (blob, params, t(actions), t(adv), t(ret), t(fixed), t(exps), 1.0 / count, 1.0 / count)
for _ in range(3):
    eng.ppo_grad(*args, ids=ids)
eng.set_stamp_buffer(stamps)
eng.ppo_grad(*args, ids=ids)
torch.cuda.synchronize()
st = stamps.cpu().numpy()
print("graph 0: n, e, k, stage =", blob.info[0])
prev, tot = st[0], st[21] - st[0]
for i in [1, 2, 3, 4, 5, 6, 7, 8, 9, 20, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21]:
    if st[i] == 0: continue
    print(f"{i:3d} {NAMES.get(i, ''):28s} {st[i] - prev:8d} cycles  {100.0 * (st[i] - prev) / tot:5.1f}%")
    prev = st[i]
print("total", tot, "cycles")
if st[46:62].any():
    print("warp probe (cycles after stamp 1):", [int(x - st[1]) for x in st[46:62]])
print("stamps rel. to stamp 1:", {i: int(st[i] - st[1]) for i in range(2, 22) if st[i]})
if st[212:224].any():
    print("warp probe B, warps 0-11 (cycles after stamp 1):", [int(x - st[1]) for x in st[212:224]])
busy = st[64:64 + eng.grid]; pro = st[224:224 + eng.grid]
print(f"per-CTA busy cycles: max {busy.max()}  mean {busy.mean():.0f}  min {busy.min()}  (balance {busy.mean() / busy.max():.2f});"
      f" launch prologue mean {pro.mean():.0f} max {pro.max()}")
