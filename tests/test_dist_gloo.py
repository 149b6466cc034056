"""CPU, world_size 2 (gloo): the data-parallel decomposition of one global minibatch -- each rank takes
perm[...][rank::world], uses the GLOBAL 1/B and 1/|ind|, and a sum all-reduce of the flat [gradient | statistics]
buffer reproduces the single-process batch gradient and losses.  The per-rank gradient here comes from the numpy
oracle (the CUDA engine needs a GPU); the sharding / reduction host logic is what is under test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drl_urban_planning_b200 import _lib, params as PL, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _shard_buffer(flat, states, actions, adv, ret, fixed, exps, ids, B, n_ind):
    """What upb_ppo_grad returns for the graphs `ids` of a global minibatch of B graphs (oracle stand-in)."""
    from oracle import sgnn_numpy as ON
    buf = np.zeros(_lib.UPB_GRAD_STRIDE, np.float64)
    if len(ids) == 0:
        return buf
    sub = [states[i] for i in ids]
    r = ON.ppo_minibatch(flat, sub, actions[ids], adv[ids], ret[ids], fixed[ids], exps[ids])
    # ppo_minibatch normalises by the shard's own sizes; rescale the three loss terms to the global ones
    b, ni = len(ids), max(int((exps[ids] != 0).sum()), 1)
    P = ON._p64(flat)
    g = np.zeros(PL.NUM_PARAMS)
    for j, i in enumerate(ids):
        gph = ON.unpad(states[i]); sid = int(np.argmax(gph.stage[:2]))
        fw = ON.forward(P, gph, action=int(actions[i, sid]), keep=True)
        g_v = 2 * 0.5 * (fw["value"] - float(ret[i, 0])) / B
        g_lp = g_en = 0.0
        if exps[i] != 0:
            ratio = np.exp(fw["log_prob"] - float(fixed[i, 0])); A = float(adv[i, 0])
            s1, s2 = ratio * A, np.clip(ratio, 0.8, 1.2) * A
            if 0.8 <= ratio <= 1.2 or s1 < s2:
                g_lp = -A * ratio / n_ind
            g_en = -0.01 / n_ind
        G = ON.backward(P, gph, fw, g_v, g_lp, g_en)
        for s in PL.SLOTS.values():
            g[s.offset:s.offset + s.size] += G[s.name].reshape(-1)
    buf[:PL.NUM_PARAMS] = g
    st = buf[_lib.UPB_STAT_OFFSET:]
    st[0] = r["value_loss"] * b; st[1] = r["surr_loss"] * ni; st[2] = r["entropy_loss"] * ni
    st[3] = b; st[4] = int((exps[ids] != 0).sum())
    return buf


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 10
    states, actions = synth.make_states(3, "tiny", B)
    adv, ret, exps = synth.make_ppo_targets(3, B)
    exps[2] = 0
    fixed = np.full((B, 1), -2.5, np.float32)
    flat = PL.default_init(3)
    perm = np.random.default_rng(0).permutation(B)
    n_ind = int((exps != 0).sum())
    ids = perm[rank::world]                                   # PPOUpdater.update_policy sharding rule
    buf = torch.tensor(_shard_buffer(flat, states, actions, adv, ret, fixed, exps, ids, B, n_ind))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)                # PPOUpdater.allreduce
    if rank == 0:
        out.put(buf.numpy())
    dist.destroy_process_group()


def test_two_rank_shards_sum_to_batch_gradient():
    from oracle import sgnn_numpy as ON
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    got = q.get(timeout=300)
    for p in procs: p.join(timeout=60)
    B = 10
    states, actions = synth.make_states(3, "tiny", B)
    adv, ret, exps = synth.make_ppo_targets(3, B)
    exps[2] = 0
    fixed = np.full((B, 1), -2.5, np.float32)
    ref = ON.ppo_minibatch(PL.default_init(3), states, actions, adv, ret, fixed, exps)
    g = got[:PL.NUM_PARAMS]
    assert np.abs(g - ref["grad"]).max() <= 1e-9 * max(np.abs(ref["grad"]).max(), 1)
    st = got[_lib.UPB_STAT_OFFSET:]
    assert st[3] == B and st[4] == int((exps != 0).sum())
    assert np.isclose(st[0] / st[3], ref["value_loss"]) and np.isclose(st[1] / st[4], ref["surr_loss"])
    assert np.isclose(st[2] / st[4], ref["entropy_loss"])
