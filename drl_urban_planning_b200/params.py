"""Flat fp32 parameter layout shared by the CUDA library, the oracle and the drop-in modules.

One contiguous buffer of 13,729 floats holds the 32 unique tensors of the reference's
`ActorCritic.state_dict()` in `ActorCritic.parameters()` order.  The C side
(`include/upb200.h`, `upb_param_layout`) carries the same table; `tests/test_abi.py` checks they agree.

Key names are the reference's (`urban_planning/models/state_encoder.py:13-33`, `policy.py:19-43`,
`value.py:15-34`); the shared encoder appears under both `actor_net.shared_net.*` and
`value_net.shared_net.*` in a checkpoint (`model.py:36-47`).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Tuple

import numpy as np

# dims fixed by every shipped config (cfg/exp_cfg/**.yaml: state_encoder_specs / policy_specs / value_specs)
NODE_DIM = 23
NUMERICAL_DIM = 52
NUM_HIDDEN = (64, 16)
GCN_DIM = 16
NUM_GCN_LAYERS = 2
HEAD_HIDDEN = 32
VALUE_HIDDEN = (32, 32)
STAGE_DIM = 3
VALUE_IN = 3 * GCN_DIM + NUM_HIDDEN[-1] + STAGE_DIM  # 67 (state_encoder.py:33)

# (short name, reference key relative to the owning module, owner, shape)
#   owner 'enc'  -> {actor_net,value_net}.shared_net.<key>
#   owner 'pol'  -> actor_net.<key>
#   owner 'val'  -> value_net.<key>
_TABLE: List[Tuple[str, str, str, Tuple[int, ...]]] = [
    ("num_w0", "numerical_feature_encoder.linear_0.weight", "enc", (64, 52)),
    ("num_b0", "numerical_feature_encoder.linear_0.bias", "enc", (64,)),
    ("num_w1", "numerical_feature_encoder.linear_1.weight", "enc", (16, 64)),
    ("num_b1", "numerical_feature_encoder.linear_1.bias", "enc", (16,)),
    ("enc_w", "node_encoder.weight", "enc", (16, 23)),
    ("enc_b", "node_encoder.bias", "enc", (16,)),
    ("gcn0_w", "edge_fc_layers.0.linear_0.weight", "enc", (16, 32)),
    ("gcn0_b", "edge_fc_layers.0.linear_0.bias", "enc", (16,)),
    ("gcn1_w", "edge_fc_layers.1.linear_0.weight", "enc", (16, 32)),
    ("gcn1_b", "edge_fc_layers.1.linear_0.bias", "enc", (16,)),
    ("mha_in_w", "attention_layer.in_proj_weight", "enc", (48, 16)),
    ("mha_in_b", "attention_layer.in_proj_bias", "enc", (48,)),
    ("mha_out_w", "attention_layer.out_proj.weight", "enc", (16, 16)),
    ("mha_out_b", "attention_layer.out_proj.bias", "enc", (16,)),
    ("att_q_w", "attention_query_layer.weight", "enc", (16, 16)),
    ("att_q_b", "attention_query_layer.bias", "enc", (16,)),
    ("att_k_w", "attention_key_layer.weight", "enc", (16, 16)),
    ("att_k_b", "attention_key_layer.bias", "enc", (16,)),
    ("att_v_w", "attention_value_layer.weight", "enc", (16, 16)),
    ("att_v_b", "attention_value_layer.bias", "enc", (16,)),
    ("lu_w0", "policy_land_use_head.land_use_linear_0.weight", "pol", (32, 64)),
    ("lu_b0", "policy_land_use_head.land_use_linear_0.bias", "pol", (32,)),
    ("lu_w1", "policy_land_use_head.land_use_linear_1.weight", "pol", (1, 32)),
    ("road_w0", "policy_road_head.road_linear_0.weight", "pol", (32, 16)),
    ("road_b0", "policy_road_head.road_linear_0.bias", "pol", (32,)),
    ("road_w1", "policy_road_head.road_linear_1.weight", "pol", (1, 32)),
    ("val_w0", "value_head.linear_0.weight", "val", (32, 67)),
    ("val_b0", "value_head.linear_0.bias", "val", (32,)),
    ("val_w1", "value_head.linear_1.weight", "val", (32, 32)),
    ("val_b1", "value_head.linear_1.bias", "val", (32,)),
    ("val_w2", "value_head.linear_2.weight", "val", (1, 32)),
    ("val_b2", "value_head.linear_2.bias", "val", (1,)),
]


class Slot:
    __slots__ = ("name", "key", "owner", "shape", "offset", "size")

    def __init__(self, name, key, owner, shape, offset):
        self.name, self.key, self.owner, self.shape, self.offset = name, key, owner, tuple(shape), offset
        self.size = int(np.prod(shape))

    def __repr__(self):
        return f"Slot({self.name}, off={self.offset}, shape={self.shape})"


def _build() -> "OrderedDict[str, Slot]":
    out, off = OrderedDict(), 0
    for name, key, owner, shape in _TABLE:
        s = Slot(name, key, owner, shape, off)
        out[name] = s
        off += s.size
    return out


SLOTS: "OrderedDict[str, Slot]" = _build()
NUM_PARAMS: int = sum(s.size for s in SLOTS.values())
assert NUM_PARAMS == 13729

# Parameter groups for the reference's `clip_policy_grad` (agent_ppo.py:43-46 with the ctor argument of
# urban_planning_agent.py:46): group 0 = policy_net.parameters() = encoder + policy heads,
# group 1 = value_net.parameters() = encoder + value head.
ENCODER_END = SLOTS["lu_w0"].offset          # [0, ENCODER_END)   shared encoder
POLICY_END = SLOTS["val_w0"].offset          # [ENCODER_END, POLICY_END) policy heads
# [POLICY_END, NUM_PARAMS) value head


def state_dict_keys(slot: Slot) -> List[str]:
    """Checkpoint key(s) of one slot in `ActorCritic.state_dict()` order."""
    if slot.owner == "enc":
        return [f"actor_net.shared_net.{slot.key}", f"value_net.shared_net.{slot.key}"]
    if slot.owner == "pol":
        return [f"actor_net.{slot.key}"]
    return [f"value_net.{slot.key}"]


def flatten(named: Dict[str, np.ndarray]) -> np.ndarray:
    """{short name -> array} -> flat float32 vector."""
    flat = np.zeros(NUM_PARAMS, dtype=np.float32)
    for s in SLOTS.values():
        flat[s.offset:s.offset + s.size] = np.asarray(named[s.name], dtype=np.float32).reshape(-1)
    return flat


def unflatten(flat: np.ndarray) -> Dict[str, np.ndarray]:
    flat = np.asarray(flat)
    return {s.name: flat[s.offset:s.offset + s.size].reshape(s.shape) for s in SLOTS.values()}


def from_state_dict(sd) -> np.ndarray:
    """Reference checkpoint dict (`actor_critic_dict`) -> flat float32 vector."""
    named = {}
    for s in SLOTS.values():
        v = sd[state_dict_keys(s)[0]]
        named[s.name] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return flatten(named)


def to_state_dict(flat: np.ndarray):
    """flat vector -> OrderedDict with the reference's 52 keys (numpy arrays)."""
    named = unflatten(np.asarray(flat, dtype=np.float32))
    actor, value = OrderedDict(), OrderedDict()
    for s in SLOTS.values():
        for k in state_dict_keys(s):
            (actor if k.startswith("actor_net.") else value)[k] = named[s.name].copy()
    out = OrderedDict()
    out.update(actor)
    out.update(value)
    return out


def default_init(seed: int) -> np.ndarray:
    """torch-default initialisation of the 32 tensors (nn.Linear Kaiming-uniform(a=sqrt 5) -> U(+-1/sqrt(fan_in))
    for weight and bias; nn.MultiheadAttention: Xavier-uniform in_proj, zero in/out-proj biases), from a numpy
    stream so weights can be made without torch.  Not bit-identical to `torch.manual_seed(seed)` init."""
    rng = np.random.default_rng(seed)
    named = {}
    for s in SLOTS.values():
        if s.name == "mha_in_w":
            bound = np.sqrt(6.0 / (48 + 16))
            named[s.name] = rng.uniform(-bound, bound, s.shape)
        elif s.name in ("mha_in_b", "mha_out_b"):
            named[s.name] = np.zeros(s.shape)
        else:
            fan_in = s.shape[1] if len(s.shape) == 2 else {
                "num_b0": 52, "num_b1": 64, "enc_b": 23, "gcn0_b": 32, "gcn1_b": 32, "att_q_b": 16,
                "att_k_b": 16, "att_v_b": 16, "lu_b0": 64, "road_b0": 16, "val_b0": 67, "val_b1": 32,
                "val_b2": 32}[s.name]
            bound = 1.0 / np.sqrt(fan_in)
            named[s.name] = rng.uniform(-bound, bound, s.shape)
    return flatten(named)


# ---------------------------------------------------------------------------------------------------------------------
# rl-mlp ablation (`create_mlp_model`, reference urban_planning/models/model.py:22-33, state_encoder.py:217-308): no
# message passing and no attention; the value features are [h_num | mean h_nodes | mean h_edges | stage] = 51 wide.
VALUE_IN_MLP = 2 * GCN_DIM + NUM_HIDDEN[-1] + STAGE_DIM        # 51 (state_encoder.py:236)
_MLP_TABLE: List[Tuple[str, str, str, Tuple[int, ...]]] = [
    ("num_w0", "numerical_feature_encoder.linear_0.weight", "enc", (64, 52)),
    ("num_b0", "numerical_feature_encoder.linear_0.bias", "enc", (64,)),
    ("num_w1", "numerical_feature_encoder.linear_1.weight", "enc", (16, 64)),
    ("num_b1", "numerical_feature_encoder.linear_1.bias", "enc", (16,)),
    ("enc_w", "node_encoder.weight", "enc", (16, 23)),
    ("enc_b", "node_encoder.bias", "enc", (16,)),
    ("lu_w0", "policy_land_use_head.land_use_linear_0.weight", "pol", (32, 64)),
    ("lu_b0", "policy_land_use_head.land_use_linear_0.bias", "pol", (32,)),
    ("lu_w1", "policy_land_use_head.land_use_linear_1.weight", "pol", (1, 32)),
    ("road_w0", "policy_road_head.road_linear_0.weight", "pol", (32, 16)),
    ("road_b0", "policy_road_head.road_linear_0.bias", "pol", (32,)),
    ("road_w1", "policy_road_head.road_linear_1.weight", "pol", (1, 32)),
    ("val_w0", "value_head.linear_0.weight", "val", (32, VALUE_IN_MLP)),
    ("val_b0", "value_head.linear_0.bias", "val", (32,)),
    ("val_w1", "value_head.linear_1.weight", "val", (32, 32)),
    ("val_b1", "value_head.linear_1.bias", "val", (32,)),
    ("val_w2", "value_head.linear_2.weight", "val", (1, 32)),
    ("val_b2", "value_head.linear_2.bias", "val", (1,)),
]


class Layout:
    """A flat parameter layout (table of tensors in `ActorCritic.parameters()` order) with the conversions above."""

    def __init__(self, table):
        self.slots: "OrderedDict[str, Slot]" = OrderedDict()
        off = 0
        for name, key, owner, shape in table:
            sl = Slot(name, key, owner, shape, off)
            self.slots[name] = sl
            off += sl.size
        self.num_params = off
        self.encoder_end = self.slots["lu_w0"].offset
        self.policy_end = self.slots["val_w0"].offset

    def flatten(self, named) -> np.ndarray:
        flat = np.zeros(self.num_params, dtype=np.float32)
        for sl in self.slots.values():
            flat[sl.offset:sl.offset + sl.size] = np.asarray(named[sl.name], dtype=np.float32).reshape(-1)
        return flat

    def unflatten(self, flat) -> Dict[str, np.ndarray]:
        flat = np.asarray(flat)
        return {sl.name: flat[sl.offset:sl.offset + sl.size].reshape(sl.shape) for sl in self.slots.values()}

    def from_state_dict(self, sd) -> np.ndarray:
        named = {}
        for sl in self.slots.values():
            v = sd[state_dict_keys(sl)[0]]
            named[sl.name] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        return self.flatten(named)

    def to_state_dict(self, flat):
        named = self.unflatten(np.asarray(flat, dtype=np.float32))
        actor, value = OrderedDict(), OrderedDict()
        for sl in self.slots.values():
            for k in state_dict_keys(sl):
                (actor if k.startswith("actor_net.") else value)[k] = named[sl.name].copy()
        out = OrderedDict()
        out.update(actor)
        out.update(value)
        return out

    def default_init(self, seed: int) -> np.ndarray:
        """nn.Linear default init (U(+-1/sqrt(fan_in)) for weight and bias) from a numpy stream."""
        rng = np.random.default_rng(seed)
        named, fan = {}, {}
        for sl in self.slots.values():
            if len(sl.shape) == 2:
                fan[sl.name[:-2] if sl.name[-2] == "w" else sl.name] = sl.shape[1]
        for sl in self.slots.values():
            if len(sl.shape) == 2:
                bound = 1.0 / np.sqrt(sl.shape[1])
            else:
                w = sl.name.replace("_b", "_w")
                bound = 1.0 / np.sqrt(self.slots[w].shape[1])
            named[sl.name] = rng.uniform(-bound, bound, sl.shape)
        return self.flatten(named)


SGNN = Layout(_TABLE)
MLP = Layout(_MLP_TABLE)
assert SGNN.num_params == NUM_PARAMS and MLP.num_params == 10257
