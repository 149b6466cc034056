"""TEST INFRASTRUCTURE ONLY.  Import the unmodified reference modules in this build container.

`/root/reference` is a Python project whose model/PPO code imports the geometry stack (geopandas,
shapely, libpysal, momepy, matplotlib, osmnx) transitively (`urban_planning/envs/__init__.py:1`,
`khrylib/utils/__init__.py:9`).  Those packages are absent here and irrelevant to the update path, so
they are replaced by inert stub modules before the import (recipe of SURVEY.md section 8(c)).  Nothing in
the reference is modified or copied; this file only exists so `tests/golden/make_golden.py` can run the
real reference to produce golden vectors.  `/root/reference` does not exist on the GPU box: nothing
under `tests/` (other than the golden generator), `bench.py` or the product imports this module.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("UPB_REFERENCE_ROOT", "/root/reference")

_STUBS = (
    "geopandas", "shapely", "shapely.geometry", "shapely.ops", "shapely.affinity", "shapely.validation",
    "libpysal", "momepy", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "osmnx", "pygad",
)


class _Stub(types.ModuleType):
    __path__: list = []
    __all__: list = []

    def __getattr__(self, name):  # any attribute -> a mock (never executed on the update path)
        if name.startswith("__"):
            raise AttributeError(name)
        return mock.MagicMock(name=f"{self.__name__}.{name}")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "urban_planning"))


def install() -> None:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


class DuckCfg:
    """The three spec dicts `create_sgnn_model` reads (cfg/exp_cfg/real/hlg.yaml:21-33)."""

    def __init__(self, max_num_nodes: int, max_num_edges: int):
        self.state_encoder_specs = dict(
            state_encoder_hidden_size=[64, 16], gcn_node_dim=16, num_gcn_layers=2, num_edge_fc_layers=1,
            max_num_nodes=max_num_nodes, max_num_edges=max_num_edges, num_attention_heads=1)
        self.policy_specs = dict(policy_land_use_head_hidden_size=[32, 1], policy_road_head_hidden_size=[32, 1])
        self.value_specs = dict(value_head_hidden_size=[32, 32, 1])


class DuckAgent:
    """The attributes of `UrbanPlanningAgent` the model factory reads (urban_planning_agent.py:117-126)."""

    def __init__(self):
        import torch
        self.node_dim = 23
        self.numerical_feature_size = 52
        self.dtype = torch.float32


def build_reference_model(max_num_nodes: int, max_num_edges: int, seed: int):
    """(policy_net, value_net, actor_critic) built by the reference's own factory under `seed`."""
    install()
    import torch
    from urban_planning.models.model import create_sgnn_model, ActorCritic
    torch.manual_seed(seed)
    policy_net, value_net = create_sgnn_model(DuckCfg(max_num_nodes, max_num_edges), DuckAgent())
    return policy_net, value_net, ActorCritic(policy_net, value_net)


def build_reference_mlp_model(max_num_nodes: int, max_num_edges: int, seed: int):
    """(policy_net, value_net, actor_critic) of the rl-mlp ablation, built by the reference's `create_mlp_model`
    (urban_planning/models/model.py:22-33) under `seed`."""
    install()
    import torch
    from urban_planning.models.model import create_mlp_model, ActorCritic
    torch.manual_seed(seed)
    policy_net, value_net = create_mlp_model(DuckCfg(max_num_nodes, max_num_edges), DuckAgent())
    return policy_net, value_net, ActorCritic(policy_net, value_net)
