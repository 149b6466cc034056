#!/usr/bin/env python
"""Stage times of the end-to-end path (host pack / H2D / step+D2H) for one 256-graph minibatch: python tools/e2e_stages.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.engine import Engine
from drl_urban_planning_b200.packing import pack_states, _pointer_table

dev = torch.device("cuda", 0)
B, POOL = 256, 8
states, actions = synth.make_states(111, "hlg", B * POOL)
host = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
t0 = time.perf_counter()
for r in range(POOL):
    _pointer_table(states[r * B:(r + 1) * B])
print(f"pointer table: {(time.perf_counter() - t0) / POOL * 1e3:.3f} ms")
for thr in (1, 4, 8, 16, 24, 32, 64):
    pack_states(states[:B], threads=thr, out_host=host)
    t0 = time.perf_counter()
    for r in range(3 * POOL):
        lo = (r % POOL) * B
        b = pack_states(states[lo:lo + B], threads=thr, out_host=host)
    print(f"pack, {thr:3d} threads: {(time.perf_counter() - t0) / (3 * POOL) * 1e3:.3f} ms  ({b.nbytes / 1e6:.1f} MB)")
dbuf = torch.empty(16 << 20, dtype=torch.uint8, device=dev)
b.to(dev, out=dbuf); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    b.to(dev, out=dbuf)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"H2D: {dt * 1e3:.3f} ms  ({b.nbytes / dt / 1e9:.1f} GB/s)")
eng = Engine(dev, b.n_cap, b.e_cap)
t = lambda x: torch.as_tensor(x, device=dev)
adv, ret, exps = synth.make_ppo_targets(1, B)
args = (b, t(PL.default_init(1)), t(actions[:B]), t(adv), t(ret), t(np.full((B, 1), -4.0, np.float32)), t(exps), 1.0 / B, 1.0 / B)
grad = eng.new_grad_buffer()
for _ in range(3):
    eng.ppo_step(*args, out=grad)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    eng.ppo_step(*args, out=grad)
    eng.read_losses(grad)
print(f"step + D2H losses (sync every step): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
