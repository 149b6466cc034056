"""Drop-in for the update half of `UrbanPlanningAgent` (reference urban_planning/agents/urban_planning_agent.py).

Usage in the reference tree:

    from drl_urban_planning_b200.agent import use_b200_update
    agent = UrbanPlanningAgent(cfg, dtype, device, num_threads, ...)      # unchanged reference constructor
    use_b200_update(agent)                                                 # update_params now runs on the B200 path

`use_b200_update` replaces `agent.update_params` (reference :248-271, which calls estimate_advantages and
update_policy :281-361).  Sampling (`sample_worker`), evaluation, logging and checkpointing stay the reference's;
after every update the new weights are written back into `agent.actor_critic_net` so `save_checkpoint` /
`sample` see them.  `estimate_advantages` is also exported with the reference's signature.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import _lib, params as PL
from .ppo import PPOUpdater


def estimate_advantages(rewards, masks, values, gamma, tau, engine=None):
    """khrylib/rl/core/common.py:5-26 on the GPU: rewards (T,), masks (T,), values (T,1) -> (T,1), (T,1) on the
    input's device.  Bit-identical to the reference's sequential fp32 scan (tests/test_gpu_parity.py)."""
    from .engine import Engine
    device = rewards.device
    if engine is None:
        dev = device if device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        engine = _default_engine(dev)
    adv, ret = engine.gae(rewards, masks, values, gamma, tau)
    return adv.reshape(-1, 1).to(device), ret.reshape(-1, 1).to(device)


_ENGINES = {}


def _default_engine(dev):
    from .engine import Engine
    key = (dev.type, dev.index)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(dev, 64, 64)
    return _ENGINES[key]


class B200Update:
    """Owns the PPOUpdater of one agent and mirrors the weights between it and the agent's torch modules."""

    def __init__(self, agent, clip_mode: int = _lib.CLIP_REFERENCE, process_group="auto", device=None):
        cfg = agent.cfg
        self.agent = agent
        dev = torch.device(device) if device is not None else agent.device
        if dev.type != "cuda":
            raise _lib.UpbError("use_b200_update needs agent.device to be a CUDA device (train.py --use_nvidia_gpu)")
        se = cfg.state_encoder_specs
        self.updater = PPOUpdater(
            PL.from_state_dict(agent.actor_critic_net.state_dict()), se["max_num_nodes"], se["max_num_edges"], dev,
            lr=cfg.lr, eps=cfg.eps, clip_epsilon=cfg.clip_epsilon, value_pred_coef=cfg.value_pred_coef,
            entropy_coef=cfg.entropy_coef, gamma=cfg.gamma, tau=cfg.tau, opt_num_epochs=cfg.num_optim_epoch,
            mini_batch_size=cfg.mini_batch_size, clip_mode=clip_mode, process_group=process_group)
        if getattr(cfg, "weightdecay", 0.0) != 0.0:
            raise NotImplementedError("weight decay != 0 is not used by any shipped cfg")
        if cfg.agent_specs.get("batch_stage", False):
            raise NotImplementedError("agent_specs.batch_stage is false in every shipped cfg")

    def push_weights(self):
        """agent modules -> updater (e.g. after load_checkpoint / freeze_*)."""
        flat = PL.from_state_dict(self.agent.actor_critic_net.state_dict())
        self.updater.params.copy_(torch.as_tensor(flat, device=self.updater.params.device))

    def pull_weights(self):
        sd = PL.to_state_dict(self.updater.flat_params())
        ref = self.agent.actor_critic_net.state_dict()
        self.agent.actor_critic_net.load_state_dict({k: torch.as_tensor(v).to(ref[k].device) for k, v in sd.items()})

    # ---- optimiser state.  The reference's checkpoints hold no Adam state (save_checkpoint :172-193): a
    # resumed run restarts the moments, and so does a fresh B200Update.  These two calls let a caller keep them.
    def optimizer_state(self) -> dict:
        m, v, steps = self.updater.engine.get_opt_state()
        return {"exp_avg": m, "exp_avg_sq": v, "steps": steps}

    def load_optimizer_state(self, state: dict) -> None:
        self.updater.engine.set_opt_state(state["exp_avg"], state["exp_avg_sq"], state["steps"])

    def update_params(self, batch, iteration):
        """Signature and effects of UrbanPlanningAgent.update_params (:248-271)."""
        t0 = time.time()
        agent = self.agent
        self.push_weights()
        tb = getattr(agent, "tb_logger", None)
        log_fn = (lambda tag, val, step: tb.add_scalar(tag, val, step)) if tb is not None else None
        self.updater.loss_iter = getattr(agent, "loss_iter", 0)
        self.updater.update_params(batch.states, batch.actions, batch.rewards, batch.masks, batch.exps,
                                   log_fn=log_fn, iteration=iteration)
        agent.loss_iter = self.updater.loss_iter
        self.pull_weights()
        return time.time() - t0


def use_b200_update(agent, **kw) -> B200Update:
    """Route `agent.update_params` through the B200 path; returns the controller object."""
    ctl = B200Update(agent, **kw)
    agent.update_params = ctl.update_params
    agent._b200_update = ctl
    return ctl
