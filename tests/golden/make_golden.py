"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each fixture it builds the reference model with `create_sgnn_model` under torch seed 111
(README.md:47 `--global_seed 111`), draws a seeded synthetic minibatch, and records

  * the flat parameter vector (drl_urban_planning_b200/params.py order),
  * value / log-prob / entropy / greedy action per state from the reference modules,
  * the four losses and the full gradient of one minibatch, computed with the reference's own
    `AgentPG.value_loss` (khrylib/rl/agents/agent_pg.py:19-23) and
    `UrbanPlanningAgent.ppo_entropy_loss` (urban_planning/agents/urban_planning_agent.py:363-371),
  * the parameter trajectory over 3 optimiser steps using the reference's `AgentPPO.clip_policy_grad`
    (generator-exhaustion quirk included) and `torch.optim.Adam(lr=4e-4, eps=1e-5)`,
  * GAE outputs of `khrylib.rl.core.estimate_advantages` on a seeded trajectory batch.

The reference pins torch<=1.13; these vectors are the behaviour under this image's torch (printed below).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402

from drl_urban_planning_b200 import params as PL, synth  # noqa: E402
from fixtures_io import compact_states  # noqa: E402

torch.set_num_threads(4)

FIXTURES = [
    # name, community, seed, count
    ("tiny_mixed", "tiny", 3, 12),
    ("small_mixed", "small", 5, 8),
    ("hlg", "hlg", 111, 4),
    ("concept", "hlg_concept", 7, 2),
]


class _Duck:
    pass


def tensorfy(states):
    return [[torch.tensor(x) for x in s] for s in states]


def run_fixture(name, community, seed, count):
    from urban_planning.agents.urban_planning_agent import UrbanPlanningAgent
    from khrylib.rl.agents import AgentPG, AgentPPO

    spec = synth.COMMUNITIES[community]
    policy_net, value_net, ac = ref_shim.build_reference_model(spec.max_num_nodes, spec.max_num_edges, 111)
    states, actions = synth.make_states(seed, community, count)
    adv, ret, exps = synth.make_ppo_targets(seed, count)
    if count >= 8:
        exps[1] = 0.0     # exercise ind != all (mean_action rollouts store exp=0, agent.py:52)
    sd0 = {k: v.detach().clone() for k, v in ac.state_dict().items()}
    flat0 = PL.from_state_dict(sd0)

    ts = tensorfy(states)
    act_t = torch.tensor(actions)
    # old log-probs from a perturbed copy of the weights so ratios straddle the clip range
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in ac.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g) * p.abs().mean())
        fixed_lp, _ = policy_net.get_log_prob_entropy(ts, act_t)
        ac.load_state_dict(sd0)
        values = value_net(ts)
        lp, ent = policy_net.get_log_prob_entropy(ts, act_t)
        greedy = policy_net.select_action(ts, mean_action=True)

    duck = _Duck()
    duck.policy_net, duck.value_net = policy_net, value_net
    duck.trans_policy = duck.trans_value = lambda s: s
    duck.clip_epsilon = 0.2
    duck.policy_grad_clip = [(policy_net.parameters(), 1), (value_net.parameters(), 1)]  # as :46
    opt = torch.optim.Adam(ac.parameters(), lr=4e-4, eps=1e-5, weight_decay=0.0)
    adv_t, ret_t, exps_t = torch.tensor(adv), torch.tensor(ret), torch.tensor(exps)
    ind = exps_t.nonzero(as_tuple=False).squeeze(1)

    losses, grads, traj = [], [], []
    for step in range(3):
        value_loss = AgentPG.value_loss(duck, ts, ret_t)
        surr, ent_loss = UrbanPlanningAgent.ppo_entropy_loss(duck, ts, act_t, adv_t, fixed_lp, ind)
        loss = surr + 0.5 * value_loss + 0.01 * ent_loss
        opt.zero_grad()
        loss.backward()
        named = {}
        for s in PL.SLOTS.values():
            p = dict(ac.named_parameters())[PL.state_dict_keys(s)[0]]
            named[s.name] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        grads.append(PL.flatten(named))
        AgentPPO.clip_policy_grad(duck)
        opt.step()
        losses.append([loss.item(), value_loss.item(), surr.item(), ent_loss.item()])
        traj.append(PL.from_state_dict(ac.state_dict()))

    out = dict(compact_states(states))
    out.update(
        params=flat0, actions=actions, advantages=adv, returns=ret, exps=exps,
        fixed_log_probs=fixed_lp.numpy().astype(np.float32),
        values=values.numpy().astype(np.float32), log_probs=lp.numpy().astype(np.float32),
        entropies=ent.numpy().astype(np.float32), greedy=greedy.numpy().astype(np.float32),
        losses=np.array(losses, np.float64), grads=np.stack(grads), params_after=np.stack(traj),
        meta=np.array([f"torch {torch.__version__}", f"community {community}", f"seed {seed}"]),
    )
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: B={count} losses0={losses[0]} |g|={np.linalg.norm(grads[0]):.4g} -> {path}"
          f" ({os.path.getsize(path) / 1024:.0f} KB)")


def run_gae():
    from khrylib.rl.core import estimate_advantages
    rng = np.random.default_rng(11)
    T = 700
    rewards = rng.standard_normal(T).astype(np.float32)
    masks = np.ones(T, np.float32)
    ends = np.sort(rng.choice(np.arange(5, T - 1), size=20, replace=False))
    masks[ends] = 0.0
    masks[-1] = 0.0
    values = rng.standard_normal((T, 1)).astype(np.float32)
    out = {"rewards": rewards, "masks": masks, "values": values}
    for tag, (gamma, tau) in {"g1t0": (1.0, 0.0), "g99t95": (0.99, 0.95)}.items():
        a, r = estimate_advantages(torch.tensor(rewards), torch.tensor(masks), torch.tensor(values), gamma, tau)
        out[f"adv_{tag}"] = a.numpy()
        out[f"ret_{tag}"] = r.numpy()
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)
    print("gae: T=700 written")


if __name__ == "__main__":
    print("torch", torch.__version__, "reference at", ref_shim.REFERENCE_ROOT)
    for f in FIXTURES:
        run_fixture(*f)
    run_gae()
