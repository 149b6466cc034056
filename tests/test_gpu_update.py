"""GPU: a whole `update_params` iteration (values -> GAE -> fixed log-probs -> epochs x minibatches, reference
urban_planning_agent.py:248-361) on the B200 path against the padded eager-PyTorch oracle port driven by the same
np.random permutations; and the drop-in nn.Modules dispatching to the CUDA library."""
import math

import numpy as np
import pytest
import torch

from drl_urban_planning_b200 import _lib, params as PL, synth
from oracle import torch_port as TP

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-9))


def port_update_params(flat, states, actions, rewards, masks, exps, gamma, tau, epochs, B, seed):
    """The reference's update_params/update_policy control flow on the oracle port (CPU)."""
    agent = TP.PortAgent(flat)
    T = len(states)
    b_all = TP.stack_states(states)
    act = torch.tensor(actions)
    with torch.no_grad():
        values = TP.value(agent.params(), b_all)
    adv, ret = TP.estimate_advantages(torch.tensor(rewards), torch.tensor(masks), values, gamma, tau)
    with torch.no_grad():
        fixed, _ = TP.log_prob_entropy(agent.params(), b_all, act)
    exps_t = torch.tensor(exps)
    np.random.seed(seed)
    losses = []
    for _ in range(epochs):
        perm = np.arange(T)
        np.random.shuffle(perm)
        for i in range(int(math.floor(T / B))):
            idx = perm[i * B:(i + 1) * B]
            b = TP.stack_states([states[j] for j in idx])
            ind = exps_t[idx].nonzero(as_tuple=False).squeeze(1)
            losses.append(agent.step(b, act[idx], adv[idx], ret[idx], fixed[idx], ind))
    return agent.flat(), np.array(losses), adv.numpy(), ret.numpy(), fixed.numpy()


def test_update_params_iteration_matches_oracle_port():
    from drl_urban_planning_b200.ppo import PPOUpdater
    dev = torch.device("cuda", 0)
    T, B, epochs = 72, 16, 2
    states, actions = synth.make_states(41, "small", T)
    rng = np.random.default_rng(41)
    rewards = rng.standard_normal(T).astype(np.float32)
    masks = np.ones(T, np.float32); masks[11::12] = 0.0
    exps = np.ones(T, np.float32); exps[5] = 0.0
    flat = PL.default_init(41)
    want, want_losses, adv, ret, fixed = port_update_params(flat, states, actions, rewards, masks, exps, 0.99, 0.95,
                                                           epochs, B, seed=7)
    spec = synth.COMMUNITIES["small"]
    up = PPOUpdater(flat, spec.max_num_nodes, spec.max_num_edges, dev, gamma=0.99, tau=0.95, opt_num_epochs=epochs,
                    mini_batch_size=B, clip_mode=_lib.CLIP_REFERENCE)
    logged = []
    np.random.seed(7)
    up.update_params(states, actions, rewards, masks, exps, log_fn=lambda tag, v, s: logged.append((tag, v, s)))
    assert rel(up.advantages.cpu().numpy(), adv.ravel()) < 1e-5
    assert rel(up.fixed_log_probs.cpu().numpy(), fixed.ravel()) < 1e-5
    got_losses = np.array([v for tag, v, s in logged if tag == "loss/loss"])
    assert got_losses.shape[0] == epochs * (T // B)
    assert np.allclose(got_losses, want_losses[:, 0], rtol=2e-4, atol=2e-5)
    assert rel(up.flat_params(), want) < 2e-5
    steps = [s for tag, v, s in logged if tag == "loss/value_loss"]
    assert steps == list(range(epochs * (T // B)))          # same TensorBoard step indices as the reference


def test_dropin_modules_dispatch_to_cuda():
    from drl_urban_planning_b200.model import ActorCritic, create_sgnn_model
    from test_model_dropin import Agent, Cfg, tensorfy
    dev = torch.device("cuda", 0)
    spec = synth.COMMUNITIES["small"]
    states, actions = synth.make_states(5, "small", 12)
    torch.manual_seed(3)
    p, v = create_sgnn_model(Cfg(spec.max_num_nodes, spec.max_num_edges), Agent())
    ac = ActorCritic(p, v)
    ts = tensorfy(states)
    with torch.no_grad():
        val_c = v(ts); lp_c, ent_c = p.get_log_prob_entropy(ts, torch.tensor(actions)); gr_c = p.select_action(ts, True)
    ac.to(dev)                                               # to_device(device, actor_critic_net) in the reference
    val_g = v(states)                                        # numpy states: packed, no tensorfy needed
    lp_g, ent_g = p.get_log_prob_entropy([[t.to(dev) for t in s] for s in ts], torch.tensor(actions).to(dev))
    gr_g = p.select_action(states, mean_action=True)
    assert val_g.shape == (12, 1) and val_g.is_cuda
    assert rel(val_g.cpu().numpy(), val_c.numpy()) < 1e-4
    assert rel(lp_g.cpu().numpy(), lp_c.numpy()) < 1e-4 and rel(ent_g.cpu().numpy(), ent_c.numpy()) < 1e-4
    assert np.array_equal(gr_g.cpu().numpy(), gr_c.numpy())
    ac.to("cpu")                                             # to_cpu(policy_net) before forking rollout workers
    assert np.array_equal(p.select_action(ts, True).numpy(), gr_c.numpy())
