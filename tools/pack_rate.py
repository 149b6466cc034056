#!/usr/bin/env python
"""Host packer throughput, back-to-back calls on a DRAM-resident pool of states: python tools/pack_rate.py [threads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import synth
from drl_urban_planning_b200.packing import pack_states
B, POOL = 256, 16
thr = int(sys.argv[1]) if len(sys.argv) > 1 else 0
states, _ = synth.make_states(111, "hlg", B * POOL)
pin = torch.cuda.is_available()
bufs = [torch.empty(16 << 20, dtype=torch.uint8, pin_memory=pin) for _ in range(2)]
for rep in range(3):
    t0 = time.perf_counter()
    for r in range(2 * POOL):
        lo = (r % POOL) * B
        b = pack_states(states[lo:lo + B], 1000, 3000, threads=thr, out_host=bufs[r & 1])
    dt = (time.perf_counter() - t0) / (2 * POOL)
    print(f"env NUMA={os.environ.get('UPB_PACK_NUMA', '-')} CHUNK={os.environ.get('UPB_PACK_CHUNK', '-')} threads={thr}: "
          f"{dt * 1e3:.3f} ms per {B} states, {B / dt / 1e3:.0f} k states/s")
