"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product path.

Padded, eager, fp32 PyTorch restatement ("port") of the reference's update path: the SGNN state encoder,
the two masked-categorical action heads, the value head, the PPO-clip / value / entropy losses, the
clip-then-Adam step and GAE.  It follows the reference's dataflow on the padded (B,N,.) / (B,E,.) layout
op for op so that (a) it reproduces the reference's numbers to fp32 round-off (pinned by
`tests/golden/*.npz`, generated from the unmodified reference by `tests/golden/make_golden.py`), and
(b) timing it on host cores is a fair stand-in for "the reference's own CPU PyTorch path" on a machine
where `/root/reference` is absent (bench.py `--impl reference`, `cpu_baseline.kind == "port"`).

Reference sites restated here:
  encoder    urban_planning/models/state_encoder.py:84-214
  heads      urban_planning/models/policy.py:45-104, urban_planning/models/value.py:36-39
  losses     khrylib/rl/agents/agent_pg.py:19-23, urban_planning/agents/urban_planning_agent.py:363-371
  step       urban_planning/agents/urban_planning_agent.py:322-337, khrylib/rl/agents/agent_ppo.py:43-46
  GAE        khrylib/rl/core/common.py:5-26
Parameters are addressed by the short names of `drl_urban_planning_b200/params.py`.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from drl_urban_planning_b200 import params as PL

EPS_DEG = 1e-6                 # SGNNStateEncoder.EPSILON (state_encoder.py:11)
MASK_FILL = -2.0 ** 32 + 1     # policy.py:50,59


# ----------------------------------------------------------------------------- containers
def params_from_flat(flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Views of one flat fp32 tensor, keyed by short name (autograd flows back into `flat`)."""
    return {s.name: flat[s.offset:s.offset + s.size].view(s.shape) for s in PL.SLOTS.values()}


def stack_states(states: Sequence[Sequence[np.ndarray]], device="cpu") -> Dict[str, torch.Tensor]:
    """tensorfy + batch_data (urban_planning_agent.py:16-20, state_encoder.py:163-177): one tensor per
    array per state, then nine stacks."""
    names = ("numerical", "node_features", "edge_index", "current_node", "node_mask", "edge_mask",
             "land_use_mask", "road_mask", "stage")
    per_state = [[torch.tensor(x).to(device) for x in st] for st in states]
    cols = list(zip(*per_state))
    return {k: torch.stack(c) for k, c in zip(names, cols)}


# ----------------------------------------------------------------------------- encoder
def _edge_messages(h: torch.Tensor, edge_index: torch.Tensor, edge_mask: torch.Tensor,
                   w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """state_encoder.py:110-130 -- symmetric edge MLP on gathered endpoint embeddings, padded rows zeroed."""
    d = h.size(-1)
    iu = edge_index[:, :, 0].unsqueeze(-1).expand(-1, -1, d)
    iv = edge_index[:, :, 1].unsqueeze(-1).expand(-1, -1, d)
    hu = torch.gather(h, 1, iu)
    hv = torch.gather(h, 1, iv)
    fwd = torch.tanh(F.linear(torch.cat([hu, hv], -1), w, b))
    rev = torch.tanh(F.linear(torch.cat([hv, hu], -1), w, b))
    he = (fwd + rev) / 2
    keep = torch.broadcast_to(edge_mask.unsqueeze(-1), he.shape)
    return torch.where(keep, he, torch.zeros_like(he))


def _scatter_side(he: torch.Tensor, idx: torch.Tensor, edge_mask: torch.Tensor, n_max: int):
    """state_encoder.py:84-108 -- sums and (broadcast) incident counts for one endpoint column."""
    bsz, _, d = he.shape
    acc = torch.zeros(bsz, n_max, d).to(he.device)
    cnt = torch.zeros_like(acc)
    ones = torch.broadcast_to(edge_mask.unsqueeze(-1), he.shape).float()
    ix = idx.unsqueeze(-1).expand(-1, -1, d)
    acc = acc.scatter_add_(1, ix, he)
    cnt = cnt.scatter_add_(1, ix, ones)
    return acc, cnt


def _aggregate(he, edge_index, edge_mask, n_max):
    """state_encoder.py:132-148."""
    a0, c0 = _scatter_side(he, edge_index[:, :, 0], edge_mask, n_max)
    a1, c1 = _scatter_side(he, edge_index[:, :, 1], edge_mask, n_max)
    return (a0 + a1) / (c0 + c1 + EPS_DEG)


def _masked_mean(x, mask):
    """state_encoder.py:179-182."""
    return (x * mask.unsqueeze(-1).float()).sum(dim=1) / mask.float().sum(dim=1, keepdim=True)


def _attend(P, h_cur, h, node_mask):
    """state_encoder.py:150-161 with nn.MultiheadAttention(16, 1) (:26); query is not key, so the generic
    in-projection + softmax(q k^T / sqrt(16)) route of F.multi_head_attention_forward is taken."""
    q = F.linear(h_cur, P["att_q_w"], P["att_q_b"]).transpose(0, 1)
    k = F.linear(h, P["att_k_w"], P["att_k_b"]).transpose(0, 1)
    v = F.linear(h, P["att_v_w"], P["att_v_b"]).transpose(0, 1)
    out, _ = F.multi_head_attention_forward(
        q, k, v, embed_dim_to_check=PL.GCN_DIM, num_heads=1,
        in_proj_weight=P["mha_in_w"], in_proj_bias=P["mha_in_b"], bias_k=None, bias_v=None,
        add_zero_attn=False, dropout_p=0.0, out_proj_weight=P["mha_out_w"], out_proj_bias=P["mha_out_b"],
        training=True, key_padding_mask=~node_mask, need_weights=True)
    return out.transpose(0, 1).squeeze(1)


def encoder(P: Dict[str, torch.Tensor], b: Dict[str, torch.Tensor]):
    """state_encoder.py:184-214.  Returns (land-use features (B,E,64), road features (B,N,16),
    value features (B,67))."""
    n_max, e_max = b["node_features"].shape[1], b["edge_index"].shape[1]
    h_num = torch.tanh(F.linear(b["numerical"].flatten(1), P["num_w0"], P["num_b0"]))
    h_num = torch.tanh(F.linear(h_num, P["num_w1"], P["num_b1"]))
    h = F.linear(b["node_features"], P["enc_w"], P["enc_b"])
    h_cur = F.linear(b["current_node"].unsqueeze(1), P["enc_w"], P["enc_b"])
    he = None
    for layer in range(PL.NUM_GCN_LAYERS):
        he = _edge_messages(h, b["edge_index"], b["edge_mask"], P[f"gcn{layer}_w"], P[f"gcn{layer}_b"])
        h = h + _aggregate(he, b["edge_index"], b["edge_mask"], n_max)
    he_mean = _masked_mean(he, b["edge_mask"])
    h_mean = _masked_mean(h, b["node_mask"])
    att = _attend(P, h_cur, h, b["node_mask"])
    sv = torch.cat([h_num, h_mean, he_mean, att, b["stage"]], dim=1)
    cur = h_cur.repeat(1, e_max, 1)
    lu = torch.cat([he, cur, he * cur, he - cur], dim=-1)
    return lu, h, sv


# ----------------------------------------------------------------------------- heads
def value(P, b) -> torch.Tensor:
    """value.py:36-39 -> (B,1)."""
    _, _, sv = encoder(P, b)
    x = torch.tanh(F.linear(sv, P["val_w0"], P["val_b0"]))
    x = torch.tanh(F.linear(x, P["val_w1"], P["val_b1"]))
    return F.linear(x, P["val_w2"], P["val_b2"])


def _distributions(P, b):
    """policy.py:45-65: per-stage masked Categorical over all E edges / all N nodes."""
    lu, road, _ = encoder(P, b)
    stage = b["stage"]
    sel0, sel1 = stage[:, 0].bool(), stage[:, 1].bool()
    d0 = d1 = None
    if stage[:, 0].sum() > 0:
        z = F.linear(torch.tanh(F.linear(lu[sel0], P["lu_w0"], P["lu_b0"])), P["lu_w1"]).flatten(1)
        fill = torch.ones_like(b["land_use_mask"][sel0], dtype=torch.float32) * MASK_FILL
        d0 = torch.distributions.Categorical(logits=torch.where(b["land_use_mask"][sel0], z, fill))
    if stage[:, 1].sum() > 0:
        z = F.linear(torch.tanh(F.linear(road[sel1], P["road_w0"], P["road_b0"])), P["road_w1"]).flatten(1)
        fill = torch.ones_like(b["road_mask"][sel1], dtype=torch.float32) * MASK_FILL
        d1 = torch.distributions.Categorical(logits=torch.where(b["road_mask"][sel1], z, fill))
    return d0, d1, sel0, sel1


def log_prob_entropy(P, b, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """policy.py:87-104 -> ((B,1), (B,1))."""
    d0, d1, sel0, sel1 = _distributions(P, b)
    bsz = b["stage"].shape[0]
    lp = torch.zeros(bsz, dtype=torch.float32, device=b["stage"].device)
    ent = torch.zeros_like(lp)
    if d0 is not None:
        lp[sel0] = d0.log_prob(actions[sel0, 0])
        ent[sel0] = d0.entropy()
    if d1 is not None:
        lp[sel1] = d1.log_prob(actions[sel1, 1])
        ent[sel1] = d1.entropy()
    return lp.unsqueeze(1), ent.unsqueeze(1)


def greedy_action(P, b) -> torch.Tensor:
    """policy.py:67-85 with mean_action=True -> (B,2) float32 (argmax of probs, first max on ties)."""
    d0, d1, sel0, sel1 = _distributions(P, b)
    act = torch.zeros(b["stage"].shape[0], 2, dtype=torch.float32)
    if d0 is not None:
        act[sel0, 0] = d0.probs.argmax(dim=1).float()
    if d1 is not None:
        act[sel1, 1] = d1.probs.argmax(dim=1).float()
    return act


# ----------------------------------------------------------------------------- losses and step
def ppo_losses(P, b, actions, advantages, returns, fixed_log_probs, ind, clip_epsilon: float = 0.2):
    """agent_pg.py:19-23 + urban_planning_agent.py:363-371.  Two encoder passes, like the reference.
    Returns (surr_loss, value_loss, entropy_loss)."""
    v = value(P, b)
    value_loss = (v - returns).pow(2).mean()
    lp, ent = log_prob_entropy(P, b, actions)
    ratio = torch.exp(lp[ind] - fixed_log_probs[ind])
    adv = advantages[ind]
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon) * adv
    surr = -torch.min(s1, s2).mean()
    entropy_loss = -ent[ind].mean()
    return surr, value_loss, entropy_loss


class PortAgent:
    """Minimal stand-in for the update half of UrbanPlanningAgent (urban_planning_agent.py:145-151,
    322-337): 32 leaf tensors, torch.optim.Adam(lr, eps, weight_decay=0) over them, and the reference's
    clipping behaviour.

    reference_clip=True reproduces SURVEY.md A.6-2: `policy_grad_clip` holds two `parameters()` generators
    (urban_planning_agent.py:46) that the first `clip_policy_grad()` call exhausts (agent_ppo.py:43-46), so
    clipping (policy group, then value group, max-norm 1; the shared encoder is scaled twice) happens on
    the first optimiser step of the agent's lifetime only.  reference_clip=False clips on every step.
    A head whose stage is absent from the minibatch keeps `grad is None` and is skipped by Adam (A.6-7).
    """

    def __init__(self, flat_init: np.ndarray, lr=4e-4, eps=1e-5, clip_epsilon=0.2, value_pred_coef=0.5,
                 entropy_coef=0.01, reference_clip=True, device="cpu"):
        flat_init = np.asarray(flat_init, dtype=np.float32)
        self.P = {s.name: torch.tensor(flat_init[s.offset:s.offset + s.size].reshape(s.shape).copy(), device=device,
                                       requires_grad=True) for s in PL.SLOTS.values()}
        self.opt = torch.optim.Adam(list(self.P.values()), lr=lr, eps=eps, weight_decay=0.0)
        self.clip_epsilon, self.value_pred_coef, self.entropy_coef = clip_epsilon, value_pred_coef, entropy_coef
        self.reference_clip = reference_clip
        self.steps_done = 0
        pol = [s.name for s in PL.SLOTS.values() if s.owner in ("enc", "pol")]
        val = [s.name for s in PL.SLOTS.values() if s.owner in ("enc", "val")]
        self._clip_groups = [pol, val]

    def params(self) -> Dict[str, torch.Tensor]:
        return self.P

    def flat(self) -> np.ndarray:
        return PL.flatten({k: v.detach().numpy() for k, v in self.P.items()})

    def flat_grad(self) -> np.ndarray:
        return PL.flatten({k: (v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32))
                           for k, v in self.P.items()})

    def clip(self):
        if self.reference_clip and self.steps_done > 0:
            return
        for names in self._clip_groups:
            torch.nn.utils.clip_grad_norm_([self.P[n] for n in names], 1.0)

    def backward(self, b, actions, advantages, returns, fixed_log_probs, ind):
        surr, vl, el = ppo_losses(self.P, b, actions, advantages, returns, fixed_log_probs, ind,
                                  self.clip_epsilon)
        loss = surr + self.value_pred_coef * vl + self.entropy_coef * el
        self.opt.zero_grad()
        loss.backward()
        return loss.item(), vl.item(), surr.item(), el.item()

    def step(self, b, actions, advantages, returns, fixed_log_probs, ind):
        """One minibatch step (urban_planning_agent.py:330-341); returns (loss, value, surr, entropy)."""
        losses = self.backward(b, actions, advantages, returns, fixed_log_probs, ind)
        self.clip()
        self.opt.step()
        self.steps_done += 1
        return losses


# ----------------------------------------------------------------------------- GAE
def estimate_advantages(rewards: torch.Tensor, masks: torch.Tensor, values: torch.Tensor,
                        gamma: float, tau: float):
    """khrylib/rl/core/common.py:5-26 -- sequential reverse scan in fp32, no normalisation.
    rewards (T,), masks (T,), values (T,1) -> advantages (T,1), returns (T,1)."""
    T = rewards.size(0)
    deltas = torch.zeros(T, 1)
    adv = torch.zeros(T, 1)
    prev_v = 0
    prev_a = 0
    for i in reversed(range(T)):
        deltas[i] = rewards[i] + gamma * prev_v * masks[i] - values[i]
        adv[i] = deltas[i] + gamma * tau * prev_a * masks[i]
        prev_v = values[i, 0]
        prev_a = adv[i, 0]
    return adv, values + adv
