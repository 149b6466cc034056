"""Drop-in for the reference's rl-mlp ablation model (`create_mlp_model`, urban_planning/models/model.py:22-33;
`MLPStateEncoder`, urban_planning/models/state_encoder.py:217-308; selectable with `train.py --agent rl-mlp`).

Same recipe as drl_urban_planning_b200/model.py for the SGNN: ordinary torch modules built from the same layers in the
same order as the reference (bit-identical seeded initialisation, the same 26 checkpoint keys), a plain-PyTorch CPU path
for the forked rollout workers, and on CUDA every call goes through libupb200.so (`k_mlp`, csrc/mlp_kernel.cuh).

What the encoder computes per graph (no message passing, no attention):
    h_i   = W_e x_i + b_e                                   node embeddings (state_encoder.py:283)
    sel_j = v_j if type(v_j) == FEASIBLE else u_j           the edge's "selected" endpoint, by raw node type (:269-273)
    he_j  = W_e x_{sel_j} + b_e = h_{sel_j}                 edge embeddings (:284; the node encoder is linear)
    s_v   = [h_num | mean_i h_i | mean_j he_j | stage]      51 value features (:291-292)
    land-use head on [he_j | h_c | he_j*h_c | he_j-h_c], road head on h_i   (:294-300)
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import params as PL
from .model import ActorCritic, MASK_FILL, UrbanPlanningPolicy, UrbanPlanningValue, _seq  # noqa: F401

FEASIBLE = 1            # urban_planning/envs/city_config.py:24
NUM_TYPE_SLOTS = 14     # city_config.NUM_TYPES + 1


def _check_mlp_specs(cfg):
    se, ps, vs = cfg.state_encoder_specs, cfg.policy_specs, cfg.value_specs
    ok = (list(se["state_encoder_hidden_size"]) == [64, 16] and se["gcn_node_dim"] == 16
          and list(ps["policy_land_use_head_hidden_size"]) == [32, 1]
          and list(ps["policy_road_head_hidden_size"]) == [32, 1]
          and list(vs["value_head_hidden_size"]) == [32, 32, 1])
    if not ok:
        raise NotImplementedError("the sm_100a kernels are specialised for the shapes of every shipped cfg "
                                  "(state_encoder [64,16], gcn_node_dim 16, heads [32,1], value [32,32,1])")


class MLPStateEncoder(nn.Module):
    """Parameter container + CPU forward of the rl-mlp encoder (reference state_encoder.py:217-308)."""
    model_kind = "mlp"

    def __init__(self, cfg, agent):
        super().__init__()
        self.cfg, self.agent = cfg, agent
        if agent.node_dim != PL.NODE_DIM or agent.numerical_feature_size != PL.NUMERICAL_DIM:
            raise NotImplementedError("node_dim must be 23 and numerical_feature_size 52")
        d = cfg["gcn_node_dim"]
        self.numerical_feature_encoder = _seq([
            ("flatten_0", nn.Flatten()), ("linear_0", nn.Linear(PL.NUMERICAL_DIM, 64)), ("tanh_0", nn.Tanh()),
            ("linear_1", nn.Linear(64, 16)), ("tanh_1", nn.Tanh())])
        self.node_encoder = nn.Linear(agent.node_dim, d)
        self.max_num_nodes, self.max_num_edges = cfg["max_num_nodes"], cfg["max_num_edges"]
        self.output_policy_land_use_size = d * 4
        self.output_policy_road_size = d
        self.output_value_size = d * 2 + cfg["state_encoder_hidden_size"][-1] + 3

    # CPU rollout path: one unpadded graph.  Returns (he (e,16), h (n,16), hc (16,), sv (51,)) like the SGNN encoder.
    def encode_one(self, state):
        numerical, nf, ei, cur, nmask, emask, _, _, stage = state
        n, e = int(nmask.sum()), int(emask.sum())
        x, edges = nf[:n], ei[:e]
        h_num = self.numerical_feature_encoder(numerical.reshape(1, -1))[0]
        h = self.node_encoder(x)
        hc = self.node_encoder(cur)
        u, v = edges[:, 0], edges[:, 1]
        feas = torch.argmax(x[:, :NUM_TYPE_SLOTS], dim=1) == FEASIBLE
        sel = torch.where(feas[v], v, u)
        he = h[sel]
        m_e = he.mean(0) if e > 0 else h.new_full((h.shape[1],), float("nan"))
        sv = torch.cat([h_num, h.mean(0), m_e, stage.to(h.dtype)])
        return he, h, hc, sv


def create_mlp_model(cfg, agent):
    """reference models/model.py:22-33 (same construction order -> same seeded initialisation)."""
    _check_mlp_specs(cfg)
    shared_net = MLPStateEncoder(cfg.state_encoder_specs, agent)
    policy_net = UrbanPlanningPolicy(cfg.policy_specs, agent, shared_net)
    value_net = UrbanPlanningValue(cfg.value_specs, agent, shared_net)
    policy_net._peer_params = lambda: dict(value_net.named_parameters())
    value_net._peer_params = lambda: dict(policy_net.named_parameters())
    return policy_net, value_net
