"""CPU ORACLE (test infrastructure — NOT part of the product path).

Pure-PyTorch, op-for-op restatement of the reference's heterogeneous-graph message-passing hot path:

* ``dglnn.GATv2Conv`` as called at reference ``algos/madrqn/agents/gnn_agents.py:92-97,103-104`` and
  ``algos/drqn/agents/gnn_agents.py:17-18,27``;
* ``GraphObservationEncoder`` (``gnn_agents.py:80-107``), ``TarMAC`` (``:232-271``), ``GnnAgent`` (``:12-56``),
  DRQN ``GnnAgent`` (``algos/drqn/agents/gnn_agents.py:9-30``), ``DuelingLayer`` (``agents/dueling.py:4-16``);
* ``nn.GRUCell`` restated explicitly (and pinned against ``torch.nn.GRUCell`` in tests).

PARITY UNPINNED BY THE REFERENCE: the arithmetic lives in ``dgl==0.9.0`` (``requirements.txt:17``), which is
neither vendored under /root/reference nor installable here (no network), and the reference ships no tests,
fixtures or golden vectors.  The DGL op semantics below are restated from DGL 0.9.0's
published ``GATv2Conv.forward`` / ``edge_softmax`` / ``update_all`` behaviour.  Our own
pins: an independent float64 loop restatement (``oracle/loops.py``), ``torch.autograd.gradcheck`` in fp64,
hand-computed micro cases, ``torch.nn.GRUCell`` / ``torch.nn.functional`` as live references for the parts that
ARE importable, and the committed golden vectors under ``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this module.  Every function works on raw edge lists ``(src, dst)`` in the reference's edge order, so it does not
depend on the product's CSR machinery.
"""
from __future__ import annotations

import math

import torch as th
import torch.nn as nn
import torch.nn.functional as F


# ============================================================================================== functional
def edge_softmax(dst: th.Tensor, e: th.Tensor, n_dst: int) -> th.Tensor:
    """``dgl.nn.functional.edge_softmax(g, e)`` with ``norm_by='dst'``: DGL runs copy_e→max, sub, exp,
    copy_e→sum, div.  ``e`` is ``(E, ...)``; softmax runs over the in-edges of each dst."""
    idx = dst.view((-1,) + (1,) * (e.dim() - 1)).expand_as(e)
    mx = th.full((n_dst,) + tuple(e.shape[1:]), float("-inf"), dtype=e.dtype, device=e.device)
    mx = mx.scatter_reduce(0, idx, e.detach(), reduce="amax", include_self=True)
    ex = th.exp(e - mx.index_select(0, dst))
    den = th.zeros_like(mx).index_add_(0, dst, ex)
    return ex / den.index_select(0, dst)


def gatv2_conv(src, dst, n_dst, feat_src, feat_dst, fc_src_w, fc_src_b, fc_dst_w, fc_dst_b, attn,
               res_w=None, res_b=None, negative_slope=0.2, activation=F.relu, identity_res=False):
    """DGL 0.9.0 ``GATv2Conv.forward`` with a (src, dst) feature pair, ``feat_drop = attn_drop = 0``:

        el = fc_src(h_src)            er = fc_dst(h_dst)                 (views (·, heads, D))
        e  = leaky_relu(el[u] + er[v])                                  apply_edges(u_add_v)
        e  = (e * attn).sum(-1, keepdim)                                (E, heads, 1)
        a  = edge_softmax(g, e)
        ft = update_all(u_mul_e('el', 'a'), sum)                        messages use el, not e
        rst = ft + res_fc(h_dst);  rst = activation(rst)                 -> (n_dst, heads, D)

    Destinations without in-edges get ``ft = 0`` (``allow_zero_in_degree=True`` at every call site)."""
    heads, D = attn.shape[-2], attn.shape[-1]
    el = F.linear(feat_src, fc_src_w, fc_src_b).view(-1, heads, D)
    er = F.linear(feat_dst, fc_dst_w, fc_dst_b).view(-1, heads, D)
    e = F.leaky_relu(el.index_select(0, src) + er.index_select(0, dst), negative_slope)
    e = (e * attn.view(1, heads, D)).sum(-1, keepdim=True)
    a = edge_softmax(dst, e, n_dst)
    m = el.index_select(0, src) * a
    rst = th.zeros(n_dst, heads, D, dtype=m.dtype, device=m.device).index_add_(0, dst, m)
    if res_w is not None:
        rst = rst + F.linear(feat_dst, res_w, res_b).view(n_dst, heads, D)
    elif identity_res:          # res_fc = Identity(): h_dst viewed (N, -1, out_feats) = (N, 1, D), broadcast over heads
        rst = rst + feat_dst.view(n_dst, -1, D)
    if activation is not None:
        rst = activation(rst)
    return rst


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """``torch.nn.GRUCell`` (gate order r, z, n)."""
    H = h.shape[1]
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    r = th.sigmoid(gi[:, :H] + gh[:, :H])
    z = th.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = th.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def tarmac_comm(src, dst, x, h, val_w, val_b, sign_w, sign_b, que_w, que_b, w_ih, w_hh, b_ih, b_hh,
                key_size, n_rounds=1):
    """Reference ``TarMAC.forward`` (``gnn_agents.py:248-271``): messages see the *detached* hidden state,
    scores are divided by ``key_size`` itself (``:262``), softmax over in-edges, GRUCell on ``[x ‖ c]``."""
    n = x.shape[0]
    for _ in range(n_rounds):
        inputs = th.cat((x, h.detach()), 1)
        v = F.linear(inputs, val_w, val_b)
        s = F.linear(inputs, sign_w, sign_b)
        q = F.linear(inputs, que_w, que_b)
        e = (s.index_select(0, src) * q.index_select(0, dst)).sum(-1, keepdim=True) / key_size
        a = edge_softmax(dst, e, n)
        c = th.zeros(n, v.shape[1], dtype=v.dtype, device=v.device).index_add_(0, dst, v.index_select(0, src) * a)
        h = gru_cell(th.cat((x, c), 1), h, w_ih, w_hh, b_ih, b_hh)
    return h


# ============================================================================================== modules
class GATv2Conv(nn.Module):
    """Parameter container with DGL 0.9.0's names, shapes and init order."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False, bias=True, share_weights=False):
        super().__init__()
        assert feat_drop == 0. and attn_drop == 0., "dropout is 0 at every reference call site"
        fs, fd = in_feats if isinstance(in_feats, tuple) else (in_feats, in_feats)
        self._num_heads, self._out_feats, self._slope = num_heads, out_feats, negative_slope
        self._allow_zero_in_degree = allow_zero_in_degree
        self.fc_src = nn.Linear(fs, out_feats * num_heads, bias=bias)
        if share_weights and not isinstance(in_feats, tuple):
            self.fc_dst = self.fc_src
        else:
            self.fc_dst = nn.Linear(fd, out_feats * num_heads, bias=bias)
        self.attn = nn.Parameter(th.empty(1, num_heads, out_feats))
        if residual:
            self.res_fc = nn.Linear(fd, num_heads * out_feats, bias=bias) if fd != out_feats else nn.Identity()
        else:
            self.register_buffer("res_fc", None)
        self.activation = activation
        self.share_weights, self.bias = share_weights, bias
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.fc_src.weight, gain=gain)
        if self.bias:
            nn.init.constant_(self.fc_src.bias, 0)
        if not (self.fc_dst is self.fc_src):
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.fc_dst.bias, 0)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.res_fc.bias, 0)

    def forward(self, graph, feat, get_attention=False):
        src, dst = graph.edges()
        n_dst = graph.num_dst_nodes()
        h_src, h_dst = feat if isinstance(feat, tuple) else (feat, feat)
        if not self._allow_zero_in_degree and n_dst and int(th.bincount(dst, minlength=n_dst).min()) == 0:
            raise RuntimeError("There are 0-in-degree nodes in the graph")
        if isinstance(self.res_fc, nn.Linear):
            rw, rb = self.res_fc.weight, self.res_fc.bias
        else:
            rw = rb = None
        return gatv2_conv(src, dst, n_dst, h_src, h_dst, self.fc_src.weight, self.fc_src.bias,
                          self.fc_dst.weight, self.fc_dst.bias, self.attn, rw, rb, self._slope, self.activation,
                          identity_res=isinstance(self.res_fc, nn.Identity))


class DuelingLayer(nn.Module):
    """Reference ``algos/madrqn/agents/dueling.py:4-16``."""

    def __init__(self, in_feats, n_actions):
        super().__init__()
        self.adv_head = nn.Linear(in_feats, n_actions)
        self.v_head = nn.Linear(in_feats, 1)

    def forward(self, x):
        advs = self.adv_head(x)
        return self.v_head(x) + (advs - advs.mean(-1, keepdim=True))


class DenseObservationEncoder(nn.Module):
    """Reference ``gnn_agents.py:62-77``."""

    def __init__(self, obs_shape, args):
        super().__init__()
        layers = [nn.Linear(obs_shape, args.hidden_size), nn.ReLU()]
        for _ in range(args.n_layers - 1):
            layers += [nn.Linear(args.hidden_size, args.hidden_size), nn.ReLU()]
        self.enc = nn.Sequential(*layers)

    def forward(self, g, x):
        return self.enc(x["agent"])


class GraphObservationEncoder(nn.Module):
    """Reference ``gnn_agents.py:80-107``: one GATv2 per relation, head-major flatten, concat, Linear+ReLU."""

    def __init__(self, obs_shape, args):
        super().__init__()
        n_heads, out_feats = args.n_heads, args.hidden_size
        assert out_feats % n_heads == 0, "out_feats cannot be divided by n_heads in GraphObservationLayer."
        d = out_feats // n_heads
        self.f_conv = nn.ModuleDict({
            "seen": GATv2Conv((obs_shape["gt"], obs_shape["agent"]), d, n_heads, residual=True,
                              allow_zero_in_degree=True, activation=nn.ReLU()),
            "near": GATv2Conv((obs_shape["ubs"], obs_shape["agent"]), d, n_heads, residual=True,
                              allow_zero_in_degree=True, activation=nn.ReLU()),
        })
        self.f_aggr = nn.Sequential(nn.Linear(len(self.f_conv) * out_feats, out_feats), nn.ReLU())

    def forward(self, g, x):
        n = g.num_nodes("agent")
        x_gt = self.f_conv["seen"](g["seen"], (x["gt"], x["agent"])).view(n, -1)
        x_ubs = self.f_conv["near"](g["near"], (x["ubs"], x["agent"])).view(n, -1)
        return self.f_aggr(th.cat((x_gt, x_ubs), 1))


class TarMAC(nn.Module):
    """Reference ``gnn_agents.py:232-271``."""

    def __init__(self, args):
        super().__init__()
        H, M, K = args.hidden_size, args.msg_size, args.key_size
        self._key_size, self._n_rounds = K, args.n_rounds
        self.f_val = nn.Linear(2 * H, M)
        self.f_sign = nn.Linear(2 * H, K)
        self.f_que = nn.Linear(2 * H, K)
        self.f_udt = nn.GRUCell(H + M, H)

    def forward(self, g, x, h):
        src, dst = g.edges()
        u = self.f_udt
        return tarmac_comm(src, dst, x, h, self.f_val.weight, self.f_val.bias, self.f_sign.weight,
                           self.f_sign.bias, self.f_que.weight, self.f_que.bias, u.weight_ih, u.weight_hh,
                           u.bias_ih, u.bias_hh, self._key_size, self._n_rounds)


def _segment_mean(dst, m, n):
    out = th.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device).index_add_(0, dst, m)
    deg = th.bincount(dst, minlength=n).clamp(min=1).to(m.dtype)
    return out / deg.unsqueeze(1)


class BaseComm(nn.Module):
    """Reference ``gnn_agents.py:113-148``: Linear message on [x_u ‖ h_u.detach()], mean aggregate, GRU."""

    def __init__(self, args):
        super().__init__()
        H, M = args.hidden_size, args.msg_size
        self._hidden_size = H
        self.f_msg = nn.Linear(2 * H, M)
        self.f_udt = nn.GRUCell(H + M, H)

    def forward(self, g, x, h):
        src, dst = g.edges()
        if g.number_of_edges() == 0:
            c = th.zeros(x.shape[0], self._hidden_size)
        else:
            m = self.f_msg(th.cat((x, h.detach()), 1).index_select(0, src))
            c = _segment_mean(dst, m, x.shape[0])
        return self.f_udt(th.cat((x, c), 1), h)


class CommNet(nn.Module):
    """Reference ``gnn_agents.py:196-229``."""

    def __init__(self, args):
        super().__init__()
        H = args.hidden_size
        self._hidden_size, self._n_rounds = H, args.n_rounds
        self.c_mod = nn.Linear(H, H)
        self.f_mod = nn.GRUCell(H, H)

    def forward(self, g, x, h):
        src, dst = g.edges()
        for _ in range(self._n_rounds):
            if g.number_of_edges() == 0:
                c = th.zeros(x.shape[0], self._hidden_size)
            else:
                c = _segment_mean(dst, h.detach().index_select(0, src), x.shape[0])
            h = self.f_mod(x + self.c_mod(c), h)
        return h


def gumbel_softmax_hard(logits, gumbels, tau):
    """``F.gumbel_softmax(logits, tau, hard=True)`` of PyTorch with the noise made explicit
    (``gumbels = -log(Exponential(1))``): straight-through one-hot, ``y_hard - y_soft.detach() + y_soft``."""
    y_soft = ((logits + gumbels) / tau).softmax(-1)
    index = y_soft.max(-1, keepdim=True)[1]
    y_hard = th.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


def _segment_max(dst, m, n):
    """UDF reduce ``nodes.mailbox['m'].max(1)[0]`` (``gnn_agents.py:176-179``): mailbox rows in edge-id order, the
    gradient follows ``torch.max(dim)`` (ONE winner per (node, feature)); nodes without messages get zeros."""
    deg = th.bincount(dst, minlength=n)
    dmax = int(deg.max()) if dst.numel() else 0
    if dmax == 0:
        return th.zeros((n,) + tuple(m.shape[1:]), dtype=m.dtype)
    order = th.sort(dst, stable=True)[1]
    start = th.cumsum(deg, 0) - deg
    slot = th.arange(dst.numel()) - start[dst[order]]
    box = th.full((n, dmax) + tuple(m.shape[1:]), float("-inf"), dtype=m.dtype)
    box = box.index_put((dst[order], slot), m[order])
    out = box.max(1)[0]
    return th.where((deg > 0).view((-1,) + (1,) * (m.dim() - 1)), out, th.zeros_like(out))


class DiscreteComm(nn.Module):
    """Reference ``gnn_agents.py:151-193``: per-EDGE hard Gumbel-softmax bits of ``f_enc([x_u ‖ h_u.detach()])``
    (tau = 0.5), element-wise max over in-edges, ``f_dec``, GRU.  ``exponential_feed``: optional iterator of the
    ``Exponential(1)`` draws (E, msg, 2) in edge-id order, one per call (else drawn from the torch RNG)."""

    def __init__(self, args):
        super().__init__()
        H, M = args.hidden_size, args.msg_size
        self._hidden_size, self._msg_size = H, M
        self.f_enc = nn.Linear(2 * H, 2 * M)
        self.f_dec = nn.Linear(2 * M, 2 * M)
        self.f_udt = nn.GRUCell(H + 2 * M, H)
        self.exponential_feed = None

    def forward(self, g, x, h):
        src, dst = g.edges()
        if g.number_of_edges() == 0:
            c = th.zeros(x.shape[0], 2 * self._msg_size, dtype=x.dtype)
        else:
            logits = self.f_enc(th.cat((x, h.detach()), 1).index_select(0, src)).view(-1, self._msg_size, 2)
            if self.exponential_feed is not None:
                e = next(self.exponential_feed).to(logits.dtype)
            else:
                e = th.empty_like(logits).exponential_()
            m = gumbel_softmax_hard(logits, -e.log(), 0.5).flatten(1)
            c = _segment_max(dst, m, x.shape[0])
        return self.f_udt(th.cat((x, self.f_dec(c)), 1), h)


class EdgeConv(nn.Module):
    """Reference ``gnn_agents.py:274-299``: message ``f_msg([x_u ‖ h_u ‖ x_v ‖ h_v])`` (h detached), mean, GRU."""

    def __init__(self, args):
        super().__init__()
        H, M = args.hidden_size, args.msg_size
        self._hidden_size, self._n_rounds = H, args.n_rounds
        self.f_msg = nn.Linear(4 * H, M)
        self.f_udt = nn.GRUCell(H + M, H)

    def forward(self, g, x, h):
        src, dst = g.edges()
        for _ in range(self._n_rounds):
            if g.number_of_edges() == 0:
                c = th.zeros(x.shape[0], self._hidden_size, dtype=x.dtype)
            else:
                xh = th.cat((x, h.detach()), 1)
                m = self.f_msg(th.cat((xh.index_select(0, src), xh.index_select(0, dst)), 1))
                c = _segment_mean(dst, m, x.shape[0])
            h = self.f_udt(th.cat((x, c), 1), h)
        return h


class GnnAgent(nn.Module):
    """Reference MADRQN ``GnnAgent`` (``gnn_agents.py:12-56``), every comm protocol of ``:27-41``."""

    def __init__(self, obs_shape, n_actions, args):
        super().__init__()
        self._hidden_size, self._comm_protocol = args.hidden_size, args.c
        if isinstance(obs_shape, int):
            self.enc = DenseObservationEncoder(obs_shape, args)
        elif isinstance(obs_shape, dict):
            self.enc = GraphObservationEncoder(obs_shape, args)
        if self._comm_protocol is None:
            self.rnn = nn.GRUCell(self._hidden_size, self._hidden_size)
        elif self._comm_protocol == "tarmac":
            self.f_comm = TarMAC(args)
        elif self._comm_protocol == "base":
            self.f_comm = BaseComm(args)
        elif self._comm_protocol == "commnet":
            self.f_comm = CommNet(args)
        elif self._comm_protocol == "disc":
            self.f_comm = DiscreteComm(args)
        elif self._comm_protocol == "econv":
            self.f_comm = EdgeConv(args)
        else:
            raise KeyError("Unsupported communication scheme.")
        self.f_out = DuelingLayer(self._hidden_size, n_actions) if args.dueling else nn.Linear(self._hidden_size, n_actions)

    def init_hidden(self):
        return th.zeros(1, self._hidden_size)

    def forward(self, g, h):
        x = self.enc(g, g.ndata["feat"]).view(g.num_nodes("agent"), -1)
        h = self.f_comm(g["talk"], x, h) if self._comm_protocol is not None else self.rnn(x, h)
        return self.f_out(h), h


class DrqnGnnAgent(nn.Module):
    """Reference DRQN ``GnnAgent`` (``algos/drqn/agents/gnn_agents.py:9-30``)."""

    def __init__(self, obs_shape, n_actions, args):
        super().__init__()
        self._hidden_size, self._n_heads = args.hidden_size, args.n_heads
        self.enc = GATv2Conv((obs_shape["gt"], obs_shape["agent"]), self._hidden_size // self._n_heads,
                             self._n_heads, residual=True, allow_zero_in_degree=True, activation=nn.ReLU())
        self.rnn = nn.GRUCell(self._hidden_size, self._hidden_size)
        self.f_out = nn.Linear(self._hidden_size, n_actions)

    def init_hidden(self):
        return th.zeros(1, self._hidden_size)

    def forward(self, g, h):
        rel = g[g.canonical_etypes[0]]
        x = self.enc(rel, (g.nodes["gt"].data["feat"], g.nodes["agent"].data["feat"])).flatten(start_dim=1)
        h = self.rnn(x, h)
        return self.f_out(h), h


# ============================================================================================== learner math
def bptt_loss(policy, target, obs_seq, h0, h0_targ, acts, rews, dones, gamma, double_q, n_agents, mixer=None,
              target_mixer=None, states=None):
    """Loss of reference ``MultiAgentQLearner.update`` (``algos/madrqn/learner.py:118-154``): ``obs_seq`` has T+1
    batched graphs; ``acts (T,N,1)``, ``rews (T,B,n_agents|1)``, ``dones (T,B,1)``; with QMIX (``:144-148``) ``states
    (T+1,B,S)`` and the two mixers turn the per-agent values into ``Q_tot``."""
    T = len(obs_seq) - 1
    h, h_targ = h0, h0_targ
    agent_out, target_out = [], []
    for t in range(T):
        logits, h = policy(obs_seq[t], h)
        agent_out.append(logits)
        with th.no_grad():
            nl, h_targ = target(obs_seq[t + 1], h_targ)
            target_out.append(nl)
    logits, h = policy(obs_seq[T], h)
    agent_out.append(logits)
    agent_out, target_out = th.stack(agent_out), th.stack(target_out)
    qvals = agent_out[:-1].gather(2, acts)
    if not double_q:
        next_vals = target_out.max(2, keepdim=True)[0]
    else:
        next_acts = th.argmax(agent_out[1:].clone().detach(), 2, keepdim=True)
        next_vals = target_out.gather(2, next_acts)
    B = rews.shape[1]
    qvals = qvals.view(T, B, n_agents)
    next_vals = next_vals.view(T, B, n_agents)
    if mixer is not None:
        qvals = mixer(qvals, states[:-1])
        next_vals = target_mixer(next_vals, states[1:])
    rews, dones = rews.expand_as(next_vals), dones.expand_as(next_vals)
    target_q = rews + gamma * (1 - dones) * next_vals
    return F.mse_loss(qvals, target_q), qvals
