#!/usr/bin/env python
"""Timeline of the pipelined end-to-end path (same structure as bench.py's e2e leg): python tools/e2e_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.engine import Engine
from drl_urban_planning_b200.packing import pack_states

dev = torch.device("cuda", 0)
B, POOL = 256, 16
states, actions = synth.make_states(111, "hlg", B * POOL)
cap = 16 << 20
host_bufs = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
dev_bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(2)]
copy_stream = torch.cuda.Stream(device=dev)
ev_h2d = [torch.cuda.Event() for _ in range(2)]
ev_done = [torch.cuda.Event() for _ in range(2)]
for ev in ev_h2d + ev_done:
    ev.record()
b0 = pack_states(states[:B], out_host=host_bufs[0]).to(dev, out=dev_bufs[0])
eng = Engine(dev, b0.n_cap, b0.e_cap)
t = lambda x: torch.as_tensor(x, device=dev)
adv, ret, exps = synth.make_ppo_targets(1, B)
rest = (t(PL.default_init(1)), t(actions[:B]), t(adv), t(ret), t(np.full((B, 1), -4.0, np.float32)), t(exps), 1.0 / B, 1.0 / B)
grad = eng.new_grad_buffer()
log = []
T0 = time.perf_counter()
now = lambda: (time.perf_counter() - T0) * 1e3

def pack_job(i):
    j = i & 1
    a = now()
    ev_h2d[j].synchronize()
    c = now()
    lo = (i % POOL) * B
    b = pack_states(states[lo:lo + B], b0.n_cap, b0.e_cap, out_host=host_bufs[j])
    log.append((i, "pack", a, c, now()))
    return b

def upload(i):
    j = i & 1
    a = now()
    b = pending.pop(i).result()
    c = now()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev_done[j])
        b.to(dev, out=dev_bufs[j])
        ev_h2d[j].record(copy_stream)
    log.append((i, "upload call", a, c, now()))
    return b

pool_ex = ThreadPoolExecutor(1)
pending = {0: pool_ex.submit(pack_job, 0)}
uploaded = {0: upload(0)}
pending[1] = pool_ex.submit(pack_job, 1)
for i in range(12):
    j = i & 1
    a = now()
    pending[i + 2] = pool_ex.submit(pack_job, i + 2)
    uploaded[i + 1] = upload(i + 1)
    b = uploaded.pop(i)
    torch.cuda.current_stream().wait_event(ev_h2d[j])
    c = now()
    eng.ppo_step(b, *rest, out=grad)
    ev_done[j].record()
    d = now()
    eng.read_losses(grad)
    log.append((i, "step", a, c, d, now()))
for f in pending.values():
    f.result()
torch.cuda.synchronize()
for rec in sorted(log, key=lambda r: r[2]):
    print(rec[0], rec[1], " ".join(f"{x:8.3f}" for x in rec[2:]))
