import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drl_urban_planning_b200 import params as PL, synth
from drl_urban_planning_b200.packing import pack_states
from drl_urban_planning_b200.engine import Engine
from oracle import sgnn_numpy as ON
dev = torch.device('cuda', 0)
spec = synth.COMMUNITIES["tiny"]
rng = np.random.default_rng(0)
s0, a0 = synth.make_state(rng, spec, n=2, stage=1, e=1)
s1, a1 = synth.make_state(rng, spec, n=spec.max_num_nodes, stage=0)
s2, _ = synth.make_state(rng, spec, n=10, stage=0); s2[6][:] = False
s3, _ = synth.make_state(rng, spec, n=12, stage=0); s3[6][:int(s3[5].sum())] = True
s4, a4 = synth.make_state(rng, spec, n=15, stage=1)
states = [s0, s1, s2, s3, s4]
actions = np.zeros((5, 2), np.float32)
actions[0, 1], actions[1, 0], actions[2, 0], actions[3, 0], actions[4, 1] = a0, a1, 0, 3, a4
adv, ret, exps = synth.make_ppo_targets(3, 5)
flat = PL.default_init(4)
fixed = np.full((5, 1), -2.0, np.float32)
t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
for i in range(5):
    st = [states[i]]
    ref = ON.ppo_minibatch(flat, st, actions[i:i+1], adv[i:i+1], ret[i:i+1], fixed[i:i+1], exps[i:i+1])
    blob = pack_states(st).to(dev)
    eng = Engine(dev, blob.n_cap, blob.e_cap)
    grad = eng.ppo_grad(blob, t(flat), t(actions[i:i+1]), t(adv[i:i+1]), t(ret[i:i+1]), t(fixed[i:i+1]), t(exps[i:i+1]), 1.0, 1.0).cpu().numpy()
    print('graph', i, 'info', blob.info[0], 'logp', ref['log_prob'], 'ent', ref['entropy'], 'val', ref['value'])
    for s in PL.SLOTS.values():
        a = grad[s.offset:s.offset+s.size]; b = ref['grad'][s.offset:s.offset+s.size]
        d = np.abs(a-b).max(); m = np.abs(b).max()
        if d > 1e-5*max(m,1e-9) and d > 1e-9:
            print('   ', s.name, 'maxdiff', d, 'maxref', m, 'maxgot', np.abs(a).max())
