#!/usr/bin/env python
"""Cycle budget of the fused tail (CTA 0): python tools/tail_times.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.engine import Engine
from drl_urban_planning_b200.packing import pack_states
dev = torch.device("cuda", 0); count = 256
states, actions = synth.make_states(111, "hlg", count)
blob = pack_states(states).to(dev)
eng = Engine(dev, blob.n_cap, blob.e_cap)
t = lambda x: torch.as_tensor(x, device=dev)
adv, ret, exps = synth.make_ppo_targets(1, count)
fixed = np.full((count, 1), -4.0, np.float32)
params = t(PL.default_init(1)).clone()
info = blob.info.astype(np.int64)
ids = t(eng.balance_ids(np.arange(count), Engine.graph_cost(info)).astype(np.int32))
args = (blob, params, t(actions), t(adv), t(ret), t(fixed), t(exps), 1.0 / count, 1.0 / count)
for _ in range(4): eng.ppo_step(*args, ids=ids)
stamps = torch.zeros(384, dtype=torch.int64, device=dev)
eng.set_stamp_buffer(stamps); eng.ppo_step(*args, ids=ids); torch.cuda.synchronize()
st = stamps.cpu().numpy()
print("CTA0: wait at the grid barrier %d | slice sums + push + flags %d | flag wait + reduce + Adam %d cycles"
      % (st[41] - st[40], st[42] - st[41], st[43] - st[42]))
busy = st[64:64 + eng.grid]
print("busy (graphs only) max %d mean %.0f" % (busy.max(), busy.mean()))
dc, dg = st[31] - st[30], st[33] - st[32]
print("CTA 0 graph loop: %d cycles in %d ns of globaltimer -> SM clock %.1f MHz" % (dc, dg, 1e3 * dc / max(dg, 1)))
# whole-step time by CUDA events for comparison
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.set_stamp_buffer(None)
for _ in range(5): eng.ppo_step(*args, ids=ids)
torch.cuda.synchronize(); ev0.record()
for _ in range(50): eng.ppo_step(*args, ids=ids)
ev1.record(); torch.cuda.synchronize()
print("events: %.2f us per step (same minibatch, L2-resident)" % (ev0.elapsed_time(ev1) * 1e3 / 50))
