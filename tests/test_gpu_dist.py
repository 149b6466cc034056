"""GPU, 2 ranks (NCCL): data-parallel update_policy reproduces the single-GPU parameter trajectory (strong scaling:
the same global minibatches, each rank takes perm[...][rank::2], one all-reduce of the gradient buffer per step)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make_case():
    from drl_urban_planning_b200 import params as PL, synth
    T = 96
    states, actions = synth.make_states(77, "small", T)
    rng = np.random.default_rng(77)
    rewards = rng.standard_normal(T).astype(np.float32)
    masks = np.ones(T, np.float32); masks[7::8] = 0.0
    return PL.default_init(77), states, actions, rewards, masks


def _run(updater, case, seed):
    flat, states, actions, rewards, masks = case
    np.random.seed(seed)
    updater.update_params(states, actions, rewards, masks)
    return updater.flat_params()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from drl_urban_planning_b200 import synth
    from drl_urban_planning_b200.ppo import PPOUpdater
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    case = _make_case()
    spec = synth.COMMUNITIES["small"]
    up = PPOUpdater(case[0], spec.max_num_nodes, spec.max_num_edges, torch.device("cuda", rank), gamma=0.99, tau=0.95,
                    opt_num_epochs=2, mini_batch_size=32)
    assert up.world == world
    out = _run(up, case, seed=5)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_two_gpu_update_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from drl_urban_planning_b200 import synth
    from drl_urban_planning_b200.ppo import PPOUpdater
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=600)
    for p in procs: p.join(timeout=120)
    case = _make_case()
    spec = synth.COMMUNITIES["small"]
    single = PPOUpdater(case[0], spec.max_num_nodes, spec.max_num_edges, torch.device("cuda", 0), gamma=0.99, tau=0.95,
                        opt_num_epochs=2, mini_batch_size=32, process_group=None)
    want = _run(single, case, seed=5)
    assert np.abs(got - want).max() <= 2e-6 * max(np.abs(want).max(), 1.0)
