import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from drl_urban_planning_b200 import synth
from drl_urban_planning_b200.packing import pack_states, _pointer_table
dev = torch.device("cuda", 0)
B = 256
for POOL in (4, 16):
    states, actions = synth.make_states(111, "hlg", B * POOL)
    bufs = [torch.empty(16 << 20, dtype=torch.uint8).pin_memory() for _ in range(2)]
    dbuf = torch.empty(16 << 20, dtype=torch.uint8, device=dev)
    for mode in ("same buffer", "alternating + H2D between", "alternating + H2D concurrent"):
        tp = 0.0
        for r in range(2 * POOL + 2):
            lo = (r % POOL) * B
            hb = bufs[0] if mode == "same buffer" else bufs[r & 1]
            t0 = time.perf_counter()
            b = pack_states(states[lo:lo + B], 1000, 3000, out_host=hb)
            if r >= 2: tp += time.perf_counter() - t0
            if mode != "same buffer":
                b.to(dev, out=dbuf)
                if mode == "alternating + H2D between": torch.cuda.synchronize()
        torch.cuda.synchronize()
        print(f"POOL {POOL:2d} {mode:30s}: pack {tp / (2 * POOL) * 1e3:.3f} ms")
    t0 = time.perf_counter()
    for r in range(POOL): _pointer_table(states[r * B:(r + 1) * B])
    print("pointer table", (time.perf_counter() - t0) / POOL * 1e3)
