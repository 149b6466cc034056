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
        # everything the kernels are specialised for is checked BEFORE a CUDA context / updater is built
        # (heads = 2 or another layer count has the same flat shapes but different maths)
        kind = getattr(cfg, "agent", "rl-sgnn")
        if kind == "rl-sgnn":
            from .model import _check_specs
            _check_specs(cfg)
            self.layout, model = PL.SGNN, "sgnn"
        elif kind == "rl-mlp":                     # the ablation agent of train.py:18 (models/model.py:22-33)
            from .mlp import _check_mlp_specs
            _check_mlp_specs(cfg)
            self.layout, model = PL.MLP, "mlp"
        else:
            raise NotImplementedError(f"agent '{kind}' has no learned update (rule / GA baselines)")
        if getattr(cfg, "weightdecay", 0.0) != 0.0:
            raise NotImplementedError("weight decay != 0 is not used by any shipped cfg")
        se = cfg.state_encoder_specs
        self.updater = PPOUpdater(
            self.layout.from_state_dict(agent.actor_critic_net.state_dict()), se["max_num_nodes"], se["max_num_edges"], dev,
            lr=cfg.lr, eps=cfg.eps, clip_epsilon=cfg.clip_epsilon, value_pred_coef=cfg.value_pred_coef,
            entropy_coef=cfg.entropy_coef, gamma=cfg.gamma, tau=cfg.tau, opt_num_epochs=cfg.num_optim_epoch,
            mini_batch_size=cfg.mini_batch_size, clip_mode=clip_mode, process_group=process_group,
            batch_stage=bool(cfg.agent_specs.get("batch_stage", False)), model=model)

    def push_weights(self):
        """agent modules -> updater (e.g. after load_checkpoint / freeze_*)."""
        flat = self.layout.from_state_dict(self.agent.actor_critic_net.state_dict())
        self.updater.params.copy_(torch.as_tensor(flat, device=self.updater.params.device))

    def pull_weights(self):
        sd = self.layout.to_state_dict(self.updater.flat_params())
        ref = self.agent.actor_critic_net.state_dict()
        self.agent.actor_critic_net.load_state_dict({k: torch.as_tensor(v).to(ref[k].device) for k, v in sd.items()})

    # ---- optimiser state.  The reference's checkpoints hold no Adam state (save_checkpoint :172-193): a
    # resumed run restarts the moments, and so does a fresh B200Update.  These calls keep them, inside the reference's
    # own checkpoint files under a key the reference ignores.
    CHECKPOINT_KEY = "b200_optimizer"

    def optimizer_state(self) -> dict:
        m, v, steps = self.updater.engine.get_opt_state()
        return {"exp_avg": m, "exp_avg_sq": v, "steps": steps}

    def load_optimizer_state(self, state: dict, clip_like_new_process: bool = True) -> None:
        """Restore the Adam moments / step counts.  `clip_like_new_process` (default) keeps the reference's behaviour
        that the FIRST optimiser step of every process clips gradients (the parameters() generators of
        urban_planning_agent.py:46 are fresh in a new process): the restored global step count is kept
        for Adam's bias correction but the clip-once latch is re-armed.  False = continue as if never interrupted."""
        self.updater.engine.set_opt_state(state["exp_avg"], state["exp_avg_sq"], state["steps"],
                                          rearm_first_step_clip=clip_like_new_process)

    def checkpoint_paths(self, iteration: int):
        """The files `UrbanPlanningAgent.save_checkpoint(iteration)` writes (urban_planning_agent.py:185-193)."""
        cfg, agent = self.agent.cfg, self.agent
        paths = []
        if cfg.save_model_interval > 0 and (iteration + 1) % cfg.save_model_interval == 0:
            paths.append("{}/iteration_{:04d}.p".format(cfg.model_dir, iteration + 1))
        if getattr(agent, "save_best_flag", False):
            paths.append("{}/best.p".format(cfg.model_dir))
            paths.append("{}/best_reward{:.2f}_iteration_{:04d}.p".format(cfg.model_dir, agent.best_rewards, iteration + 1))
        return paths

    def save_checkpoint(self, iteration: int):
        """`agent.save_checkpoint` (:172-193) followed by adding the Adam state to every file it wrote."""
        import os
        import pickle
        paths = self.checkpoint_paths(iteration)
        self._ref_save_checkpoint(iteration)
        state = None
        for p in paths:
            if os.path.exists(p):
                if state is None:
                    state = self.optimizer_state()
                with open(p, "rb") as f:
                    cp = pickle.load(f)
                cp[self.CHECKPOINT_KEY] = state
                with open(p, "wb") as f:
                    pickle.dump(cp, f)

    def load_checkpoint(self, checkpoint, restore_best_rewards: bool = True):
        """`agent.load_checkpoint` (:153-170) followed by restoring the Adam state if the file carries it."""
        import pickle
        start = self._ref_load_checkpoint(checkpoint, restore_best_rewards)
        cfg = self.agent.cfg
        cp_path = ("%s/iteration_%04d.p" % (cfg.model_dir, checkpoint)) if isinstance(checkpoint, int) else \
            ("%s/%s.p" % (cfg.model_dir, checkpoint))
        with open(cp_path, "rb") as f:
            cp = pickle.load(f)
        self.push_weights()
        if self.CHECKPOINT_KEY in cp:
            self.load_optimizer_state(cp[self.CHECKPOINT_KEY])
        return start

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
    # checkpoints: the reference's files, plus the Adam moments under a key it ignores
    if hasattr(agent, "save_checkpoint"):
        ctl._ref_save_checkpoint = agent.save_checkpoint
        agent.save_checkpoint = ctl.save_checkpoint
    if hasattr(agent, "load_checkpoint"):
        ctl._ref_load_checkpoint = agent.load_checkpoint
        agent.load_checkpoint = ctl.load_checkpoint
    agent._b200_update = ctl
    return ctl
