"""Agent modules of the hot path with the reference's constructor / forward signatures and ``state_dict`` keys.

Mirrors reference ``algos/madrqn/agents/gnn_agents.py`` (``GnnAgent`` :12-56, ``DenseObservationEncoder`` :62-77,
``GraphObservationEncoder`` :80-107, ``BaseComm`` :113-148, ``DiscreteComm`` :151-193, ``CommNet`` :196-229,
``TarMAC`` :232-271, ``EdgeConv`` :274-299), ``algos/drqn/agents/gnn_agents.py`` (``GnnAgent`` :9-30) and DGL 0.9.0's
``dglnn.GATv2Conv``, but the graph arithmetic runs in the sm_100a kernels behind
``include/ubs_gnn.h`` (``ops.py``); DGL is not used.  ``learner._build_agent`` can instantiate these through
``REGISTRY['gnn']`` unchanged.
"""
from __future__ import annotations

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .dueling import DuelingLayer


class GRUCell(nn.GRUCell):
    """``nn.GRUCell`` parameters (``weight_ih, weight_hh, bias_ih, bias_hh``); gate math in ``ubs_gru_gates_*``."""

    def forward(self, x, h):
        return ops.gru_cell(x, h, self.weight_ih, self.weight_hh, self.bias_ih, self.bias_hh)


class GATv2Conv(nn.Module):
    """Drop-in for ``dglnn.GATv2Conv`` (DGL 0.9.0 signature, parameter names, init order).

    ``forward(graph, (feat_src, feat_dst)) -> (N_dst, heads, out_feats)`` where ``graph`` is a relation view
    (``g['seen']``).  The fused kernel covers the reference's shapes (``F_src <= 4``, ``F_dst <= 2``,
    ``heads*out_feats in {32, 64, 128}``); other shapes raise — there is no silent eager fallback."""

    def __init__(self, in_feats, out_feats, num_heads, feat_drop=0., attn_drop=0., negative_slope=0.2,
                 residual=False, activation=None, allow_zero_in_degree=False, bias=True, share_weights=False):
        super().__init__()
        if feat_drop != 0. or attn_drop != 0.:
            raise NotImplementedError("feat_drop / attn_drop are 0 at every reference call site")
        self._num_heads = num_heads
        self._in_src_feats, self._in_dst_feats = in_feats if isinstance(in_feats, tuple) else (in_feats, in_feats)
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self._negative_slope = negative_slope
        self.fc_src = nn.Linear(self._in_src_feats, out_feats * num_heads, bias=bias)
        if share_weights and not isinstance(in_feats, tuple):
            self.fc_dst = self.fc_src
        else:
            self.fc_dst = nn.Linear(self._in_dst_feats, out_feats * num_heads, bias=bias)
        self.attn = nn.Parameter(th.empty(1, num_heads, out_feats))
        if residual:
            if self._in_dst_feats != out_feats:
                self.res_fc = nn.Linear(self._in_dst_feats, num_heads * out_feats, bias=bias)
            else:
                self.res_fc = nn.Identity()          # DGL: h_dst viewed (N, 1, out_feats), broadcast over the heads
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
        if self.fc_dst is not self.fc_src:
            nn.init.xavier_normal_(self.fc_dst.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.fc_dst.bias, 0)
        nn.init.xavier_normal_(self.attn, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)
            if self.bias:
                nn.init.constant_(self.res_fc.bias, 0)

    def forward(self, graph, feat, get_attention=False):
        if get_attention:
            raise NotImplementedError("get_attention=True is not used on this path")
        h_src, h_dst = feat if isinstance(feat, tuple) else (feat, feat)
        csr = graph.csr()
        if not self._allow_zero_in_degree and csr.n_dst and int((csr.indptr[1:] == csr.indptr[:-1]).any()):
            raise RuntimeError("There are 0-in-degree nodes in the graph; set allow_zero_in_degree=True")
        relu = isinstance(self.activation, nn.ReLU) or self.activation in (F.relu, th.relu)
        if self.activation is not None and not relu:
            flags, post = 0, self.activation
        else:
            flags, post = (ops.GAT_RELU if relu else 0), None
        has_res = isinstance(self.res_fc, nn.Linear)
        flags |= ops.GAT_RESIDUAL if has_res else 0
        if not ops.gatv2_fused_supported(self._in_src_feats, self._in_dst_feats, self._num_heads, self._out_feats,
                                         self._negative_slope):
            # wide inputs (synthetic sweep): dense projections on the tensor cores, then ONE gather/softmax/aggregate pass
            if not ops.gat_aggregate_supported(self._num_heads, self._out_feats, self._negative_slope):
                raise NotImplementedError(
                    f"GATv2Conv shape (F_src={self._in_src_feats}, F_dst={self._in_dst_feats}, "
                    f"heads={self._num_heads}, D={self._out_feats}) is outside the kernels' range")
            # dense per-relation feature projections: tcgen05 3xTF32 GEMMs (ubs_tf32x3_gemm / _tn) with autograd
            el = ops.linear(h_src, self.fc_src.weight, self.fc_src.bias)
            er = ops.linear(h_dst, self.fc_dst.weight, self.fc_dst.bias)
            if has_res:
                rs = ops.linear(h_dst, self.res_fc.weight, self.res_fc.bias)
            elif isinstance(self.res_fc, nn.Identity):
                rs = h_dst.repeat(1, self._num_heads)
            else:
                rs = None
            out = ops.GATAggregate.apply(el, er, rs, csr.indptr, csr.src_idx, self.attn, self._num_heads,
                                         self._out_feats, self._negative_slope, flags & ops.GAT_RELU)
            out = out.view(-1, self._num_heads, self._out_feats)
            return post(out) if post is not None else out
        out = ops.GATv2Fused.apply(h_src, h_dst, csr.indptr, csr.src_idx, self.fc_src.weight, self.fc_src.bias,
                                   self.fc_dst.weight, self.fc_dst.bias, self.attn,
                                   self.res_fc.weight if has_res else None, self.res_fc.bias if has_res else None,
                                   self._num_heads, self._out_feats, self._negative_slope, flags)
        out = out.view(-1, self._num_heads, self._out_feats)
        return post(out) if post is not None else out


# ============================================================================================== encoders
class DenseObservationEncoder(nn.Module):
    """MLP observation encoder (reference ``gnn_agents.py:62-77``)."""

    def __init__(self, obs_shape, args):
        super().__init__()
        self._n_layers, self._hidden_size = args.n_layers, args.hidden_size
        layers = [nn.Linear(obs_shape, self._hidden_size), nn.ReLU()]
        for _ in range(self._n_layers - 1):
            layers += [nn.Linear(self._hidden_size, self._hidden_size), nn.ReLU()]
        self.enc = nn.Sequential(*layers)

    def forward(self, g, x):
        return self.enc(x["agent"])


class GraphObservationEncoder(nn.Module):
    """Heterogeneous graph observation encoder (reference ``gnn_agents.py:80-107``): one GATv2 per relation
    (``seen``: gt→agent, ``near``: ubs→agent), head-major flatten, concat, ``Linear(2H, H) + ReLU``."""

    def __init__(self, obs_shape, args):
        super().__init__()
        n_heads, out_feats = args.n_heads, args.hidden_size
        assert out_feats % n_heads == 0, "out_feats cannot be divided by n_heads in GraphObservationLayer."
        feats_per_head = out_feats // n_heads
        self.f_conv = nn.ModuleDict({
            "seen": GATv2Conv((obs_shape["gt"], obs_shape["agent"]), feats_per_head, n_heads, residual=True,
                              allow_zero_in_degree=True, activation=nn.ReLU()),
            "near": GATv2Conv((obs_shape["ubs"], obs_shape["agent"]), feats_per_head, n_heads, residual=True,
                              allow_zero_in_degree=True, activation=nn.ReLU()),
        })
        self.f_aggr = nn.Sequential(nn.Linear(len(self.f_conv) * out_feats, out_feats), nn.ReLU())

    def forward_relations(self, g, x):
        """``[x_gt ‖ x_ubs]`` (N, 2H): the per-relation GATv2 outputs before the aggregator (the fused agent step
        applies ``f_aggr`` itself)."""
        n = g.num_nodes("agent")
        x_gt = self.f_conv["seen"](g["seen"], (x["gt"], x["agent"])).view(n, -1)
        x_ubs = self.f_conv["near"](g["near"], (x["ubs"], x["agent"])).view(n, -1)
        return th.cat((x_gt, x_ubs), 1)

    def forward(self, g, x):
        return self.f_aggr(self.forward_relations(g, x))


# ============================================================================================== comm protocols
def _mean_of_sources(g, per_node_msg):
    """Mean over in-edges of a message that depends on the source node only (BaseComm / CommNet).  Batched per-env
    comm graphs on the GPU go through the block-mean kernel (no edge list); anything else through torch ops."""
    blk = g.block_mask() if per_node_msg.is_cuda else None
    if blk is not None:
        block, mask = blk
        return ops.BlockMean.apply(per_node_msg, mask, block)
    src, _ = g.edges()
    return _mean_by_dst(g, per_node_msg.index_select(0, src), per_node_msg.shape[0])


def _mean_by_dst(g, msg, n):
    _, dst = g.edges()
    out = th.zeros((n,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device).index_add_(0, dst, msg)
    deg = th.bincount(dst, minlength=n).clamp_(min=1).to(msg.dtype)
    return out / deg.unsqueeze(1)


class TarMAC(nn.Module):
    """TarMAC targeted communication (reference ``gnn_agents.py:232-271``).

    Per round: ``[v|s|q] = W_vsq [x ‖ h.detach()]`` (one library GEMM), block attention kernel
    (``e = <s_u, q_v> / key_size``, softmax over in-edges, ``c_v = Σ a v_u``), then ``GRUCell([x ‖ c], h)``."""

    def __init__(self, args):
        super().__init__()
        self._hidden_size, self._msg_size = args.hidden_size, args.msg_size
        self._key_size, self._n_rounds = args.key_size, args.n_rounds
        self.f_val = nn.Linear(2 * self._hidden_size, self._msg_size)
        self.f_sign = nn.Linear(2 * self._hidden_size, self._key_size)
        self.f_que = nn.Linear(2 * self._hidden_size, self._key_size)
        self.f_udt = GRUCell(self._hidden_size + self._msg_size, self._hidden_size)

    def forward(self, g, x, h):
        blk = g.block_mask()
        if blk is None:
            raise NotImplementedError("TarMAC needs a batch of equally sized comm graphs with <= 32 agents each")
        block, mask = blk
        K, M = self._key_size, self._msg_size
        w = th.cat((self.f_val.weight, self.f_sign.weight, self.f_que.weight), 0)
        b = th.cat((self.f_val.bias, self.f_sign.bias, self.f_que.bias), 0)
        for _ in range(self._n_rounds):
            inputs = th.cat((x, h.detach()), 1)
            vsq = th.addmm(b, inputs, w.t())
            c = ops.BlockAttention.apply(vsq, mask, block, K, M, 1.0 / K)
            h = self.f_udt(th.cat((x, c), 1), h)
        return h


class BaseComm(nn.Module):
    """Reference ``gnn_agents.py:113-148``: message ``f_msg([x_u ‖ h_u.detach()])``, mean over in-edges, GRU."""

    def __init__(self, args):
        super().__init__()
        self._hidden_size, self._msg_size = args.hidden_size, args.msg_size
        self.f_msg = nn.Linear(2 * self._hidden_size, self._msg_size)
        self.f_udt = GRUCell(self._hidden_size + self._msg_size, self._hidden_size)

    def forward(self, g, x, h):
        if g.number_of_edges() == 0:
            c = th.zeros(x.shape[0], self._hidden_size, device=x.device)
        else:
            c = _mean_of_sources(g, self.f_msg(th.cat((x, h.detach()), 1)))
        return self.f_udt(th.cat((x, c), 1), h)


def _first_max_by_dst(g, msg, n):
    """``nodes.mailbox['m'].max(1)[0]`` of the reference's UDF reduce (``gnn_agents.py:176-179``) for per-EDGE messages
    ``msg (E, F)`` in edge-id order: element-wise max over the in-edges of every destination, zeros where there are
    none, and — like ``torch.max(dim)`` on the degree-bucketed mailbox — the gradient goes to ONE winner per
    (destination, feature): the first maximal entry in mailbox (= edge-id) order."""
    src, dst = g.edges()
    E, F_ = msg.shape
    if E == 0:
        return th.zeros(n, F_, dtype=msg.dtype, device=msg.device)
    order = th.sort(dst, stable=True)[1]                       # mailbox order: a destination's in-edges by edge id
    d_sorted = dst[order]
    deg = th.bincount(dst, minlength=n)
    start = th.cumsum(deg, 0) - deg
    slot = th.arange(E, device=dst.device) - start[d_sorted]
    dmax = int(deg.max())
    box = th.full((n, dmax, F_), float("-inf"), dtype=msg.dtype, device=msg.device)
    box = box.index_put((d_sorted, slot), msg[order])
    top = box.detach().max(1, keepdim=True)[0]
    is_top = box.detach() == top
    first = is_top & (th.cumsum(is_top, 1) == 1)               # exactly one winner per (destination, feature)
    out = th.where(first, box, th.zeros_like(box)).sum(1)
    return th.where((deg > 0).unsqueeze(1), out, th.zeros_like(out))


class DiscreteComm(nn.Module):
    """Reference ``gnn_agents.py:151-193``: 2-digit one-hot bits via hard Gumbel-softmax (tau=0.5) on every edge,
    element-wise OR (max) over in-edges, ``f_dec``, GRU.  Noise comes from the torch RNG, per edge in edge-id order, as in
    the reference's edge UDF (``exponential_feed``: optional iterator of pre-drawn ``Exponential(1)`` tensors
    ``(E, msg, 2)``, one per call — how the parity tests replay the noise the reference consumed).  The encoder is
    applied per SOURCE NODE (``f_enc`` sees only the source's ``[x ‖ h]``), the noise per edge."""

    def __init__(self, args):
        super().__init__()
        self._hidden_size, self._msg_size = args.hidden_size, args.msg_size
        self.f_enc = nn.Linear(2 * self._hidden_size, 2 * self._msg_size)
        self.f_dec = nn.Linear(2 * self._msg_size, 2 * self._msg_size)
        self.f_udt = GRUCell(self._hidden_size + 2 * self._msg_size, self._hidden_size)
        self.exponential_feed = None

    def forward(self, g, x, h):
        n, M = x.shape[0], self._msg_size
        E = g.number_of_edges()
        if E == 0:
            c = th.zeros(n, 2 * M, device=x.device)
        else:
            logits_n = self.f_enc(th.cat((x, h.detach()), 1))                  # per source NODE (N, 2M)
            if self.exponential_feed is not None:
                e = next(self.exponential_feed).to(logits_n.dtype)
            else:
                e = th.empty(E, M, 2, dtype=logits_n.dtype, device=x.device).exponential_()
            blk = g.block_mask() if x.is_cuda else None
            if blk is not None:
                # batched per-env comm graphs: fused kernel (noise per edge, max over the destination's bit mask)
                block, mask = blk
                c = ops.BlockBitMax.apply(logits_n, e.reshape(E, M, 2), mask, block, 0.5)
            else:
                src, _ = g.edges()
                logits = logits_n.index_select(0, src).view(-1, M, 2)
                # F.gumbel_softmax(logits, tau=0.5, hard=True) with the noise made explicit (straight-through one-hot)
                y_soft = ((logits - e.log()) / 0.5).softmax(-1)
                y_hard = th.zeros_like(y_soft).scatter_(-1, y_soft.argmax(-1, keepdim=True), 1.0)
                m = (y_hard - y_soft.detach() + y_soft).flatten(1)
                c = _first_max_by_dst(g, m, n)
        return self.f_udt(th.cat((x, self.f_dec(c)), 1), h)


class CommNet(nn.Module):
    """Reference ``gnn_agents.py:196-229``: mean of neighbours' detached hidden states, skip connection."""

    def __init__(self, args):
        super().__init__()
        self._hidden_size, self._n_rounds = args.hidden_size, args.n_rounds
        self.c_mod = nn.Linear(self._hidden_size, self._hidden_size)
        self.f_mod = GRUCell(self._hidden_size, self._hidden_size)

    def forward(self, g, x, h):
        for _ in range(self._n_rounds):
            if g.number_of_edges() == 0:
                c = th.zeros(x.shape[0], self._hidden_size, device=x.device)
            else:
                c = _mean_of_sources(g, h.detach())
            h = self.f_mod(x + self.c_mod(c), h)
        return h


class EdgeConv(nn.Module):
    """Reference ``gnn_agents.py:274-299``: message ``f_msg([x_u ‖ h_u ‖ x_v ‖ h_v])`` (h detached), mean, GRU.

    The message is linear in its source and destination halves, so the mean over in-edges is
    ``W_src · mean_u([x_u ‖ h_u]) + W_dst · [x_v ‖ h_v] + b`` (zero for destinations without in-edges): ONE block-mean
    launch on the node features + one library GEMM instead of a per-edge ``4H``-wide projection."""

    def __init__(self, args):
        super().__init__()
        self._hidden_size, self._msg_size, self._n_rounds = args.hidden_size, args.msg_size, args.n_rounds
        self.f_msg = nn.Linear(4 * self._hidden_size, self._msg_size)
        self.f_udt = GRUCell(self._hidden_size + self._msg_size, self._hidden_size)

    def forward(self, g, x, h):
        for _ in range(self._n_rounds):
            if g.number_of_edges() == 0:
                c = th.zeros(x.shape[0], self._hidden_size, device=x.device)
            else:
                xh = th.cat((x, h.detach()), 1)
                has_in = (g.in_degrees() > 0).unsqueeze(1).to(xh.dtype)
                c = self.f_msg(th.cat((_mean_of_sources(g, xh), xh), 1)) * has_in
            h = self.f_udt(th.cat((x, c), 1), h)
        return h


# ============================================================================================== agents
class GnnAgent(nn.Module):
    """Recurrent agent with graph observation encoder and comm protocol (reference ``gnn_agents.py:12-56``)."""

    def __init__(self, obs_shape, n_actions, args):
        super().__init__()
        self._hidden_size = args.hidden_size
        self._comm_protocol = args.c
        if isinstance(obs_shape, int):
            self.enc = DenseObservationEncoder(obs_shape, args)
        elif isinstance(obs_shape, dict):
            self.enc = GraphObservationEncoder(obs_shape, args)
        if self._comm_protocol is None:
            self.rnn = GRUCell(self._hidden_size, self._hidden_size)
        elif self._comm_protocol == "base":
            self.f_comm = BaseComm(args)
        elif self._comm_protocol == "disc":
            self.f_comm = DiscreteComm(args)
        elif self._comm_protocol == "commnet":
            self.f_comm = CommNet(args)
        elif self._comm_protocol == "tarmac":
            self.f_comm = TarMAC(args)
        elif self._comm_protocol == "econv":
            self.f_comm = EdgeConv(args)
        else:
            raise KeyError("Unsupported communication scheme.")
        self.f_out = DuelingLayer(self._hidden_size, n_actions) if args.dueling else nn.Linear(self._hidden_size, n_actions)
        self._n_actions = n_actions
        self._msg_size, self._key_size = getattr(args, "msg_size", 0), getattr(args, "key_size", 0)
        self._n_rounds = getattr(args, "n_rounds", 1)
        self._pack_cache = {}
        self._relpack_cache = None
        self._param_gen = 0
        self._all_params, self._refresh_fp = None, None
        self.use_rel_act = True       # act step with the two observation relations folded into the act kernel (one launch)
        self.use_seq2 = True          # resident-weight sequence kernels when they fit (False: weight-streaming kernels)
        self.use_seq2_act = False     # act step through the resident-weight kernel (T = 1) instead of the streaming one

    def init_hidden(self):
        return th.zeros(1, self._hidden_size)

    # ---- fused path (one kernel for aggregator + comm + GRU + Q head; whole sequences in one launch) -------------
    def _fused_params(self):
        p = {k: None for k in ops.PARAM_ORDER}
        if isinstance(self.enc, GraphObservationEncoder):
            p["W_aggr"], p["b_aggr"] = self.enc.f_aggr[0].weight, self.enc.f_aggr[0].bias
        if self._comm_protocol == "tarmac":
            c = self.f_comm
            p.update(W_val=c.f_val.weight, b_val=c.f_val.bias, W_sign=c.f_sign.weight, b_sign=c.f_sign.bias,
                     W_que=c.f_que.weight, b_que=c.f_que.bias)
            cell = c.f_udt
        else:
            cell = self.rnn
        p.update(W_ih=cell.weight_ih, b_ih=cell.bias_ih, W_hh=cell.weight_hh, b_hh=cell.bias_hh,
                 W_out=self.f_out.weight, b_out=self.f_out.bias)
        return p

    def fused_dims(self, agents_per_env):
        """``ops.AgentDims`` of the fused step for this agent, or ``None`` when the configuration is outside it
        (other comm protocols, dueling head, multi-round TarMAC, > 16 agents per env)."""
        if self._comm_protocol not in (None, "tarmac") or not isinstance(self.f_out, nn.Linear):
            return None
        if self._comm_protocol == "tarmac" and self._n_rounds != 1:
            return None
        graph_enc = isinstance(self.enc, GraphObservationEncoder)
        flags = (ops.STEP_AGGR if graph_enc else 0) | (ops.STEP_TARMAC if self._comm_protocol == "tarmac" else 0)
        H = self._hidden_size
        if self._comm_protocol == "tarmac" and agents_per_env is None:
            return None
        d = ops.AgentDims(H, self._msg_size, self._key_size, self._n_actions, agents_per_env or 1,
                          2 * H if graph_enc else H, flags)
        return d if d.supported() else None

    def _packed(self, dims, params):
        key = (self._param_gen,) + tuple((t.data_ptr(), t._version) for t in params.values() if t is not None)
        hit = self._pack_cache.get(dims.ints())
        if hit is None or hit[0] != key:
            buf = ops.agent_pack(dims, params, None if hit is None else hit[1])
            self._pack_cache[dims.ints()] = (key, buf)
            return buf
        return hit[1]

    def _relpacked(self):
        """Constant tables of the two observation relations for the fused act step, rebuilt (in place) when a parameter
        of the relation encoders changed."""
        convs = (self.enc.f_conv["seen"], self.enc.f_conv["near"])
        ps = [(c.fc_src.weight, c.fc_src.bias, c.fc_dst.weight, c.fc_dst.bias, c.attn, c.res_fc.weight, c.res_fc.bias)
              for c in convs]
        key = (self._param_gen,) + tuple((t.data_ptr(), t._version) for p in ps for t in p)
        hit = self._relpack_cache
        if hit is None or hit[0] != key:
            c0 = convs[0]
            buf = ops.gatv2_rel_pack(ps, [c._in_src_feats for c in convs], c0._in_dst_feats, c0._num_heads, c0._out_feats,
                                     c0._negative_slope, ops.GAT_RESIDUAL | ops.GAT_RELU, None if hit is None else hit[1])
            self._relpack_cache = hit = (key, buf)
        return hit[1]

    def rel_act_supported(self, arena):
        """Whether ``arena_step`` runs as ONE kernel (relations + agent step) for this agent on this arena layout."""
        dims = self.arena_dims(arena)
        if dims is None or not self.use_rel_act or not isinstance(self.enc, GraphObservationEncoder):
            return False
        L, c0 = arena.layout, self.enc.f_conv["seen"]
        return ops.agent_act_rel_supported(dims, c0._num_heads, L.F_gt, L.G, L.F_ubs, max(L.U - 1, 0), L.F_ag)

    def mark_params_changed(self):
        """For writers that do not bump ``Tensor._version`` (the fused multi-tensor AdamW kernel): the packed copies are
        rebuilt at the next ``refresh_packed`` / act step."""
        self._param_gen += 1

    def refresh_packed(self, arena=None):
        """Re-packs (in place) every act-step weight buffer built so far if a parameter changed since — a version check
        when nothing did.  Captured CUDA graphs read these buffers by address, so the learner calls this before every
        graph replay and after everything that writes parameters (optimizer step, polyak update, ``load_checkpoint``).
        With ``arena``: also builds the buffers that arena's act step needs (must exist before a capture starts).
        Hit path: ONE fingerprint over all parameters (address + version, ≈ 5 µs of Python) compared with the one this
        method last left the buffers consistent with — it runs before every act step of a host-driven loop."""
        if self._all_params is None:
            self._all_params = list(self.parameters())
        fp = (self._param_gen, len(self._pack_cache), self._relpack_cache is None,
              None if arena is None else arena.layout.key(),
              tuple((p.data_ptr(), p._version) for p in self._all_params))
        if fp == self._refresh_fp:
            return
        self._refresh_fp = None
        params = None
        for ints in list(self._pack_cache):
            params = params or self._fused_params()
            self._packed(ops.AgentDims(*ints), params)
        if self._relpack_cache is not None:
            self._relpacked()
        if arena is not None:
            dims = self.arena_dims(arena)
            if dims is not None:
                self._packed(dims, params or self._fused_params())
                if self.rel_act_supported(arena):
                    self._relpacked()
        # fingerprint AFTER the work: the caches may have grown (first call for this arena)
        self._refresh_fp = (self._param_gen, len(self._pack_cache), self._relpack_cache is None,
                            None if arena is None else arena.layout.key(), fp[4])

    def _encode_pre(self, g):
        """Input of the fused step: ``[x_gt ‖ x_ubs]`` for the graph encoder (aggregator fused), else the encoder output."""
        if isinstance(self.enc, GraphObservationEncoder):
            return self.enc.forward_relations(g, g.ndata["feat"])
        return self.enc(g, g.ndata["feat"]).view(g.num_nodes("agent"), -1)

    def _talk_mask(self, g):
        if self._comm_protocol != "tarmac":
            return None, None
        blk = g["talk"].block_mask()
        return (None, None) if blk is None else blk

    def can_fuse(self, g):
        return self.fused_dims(self._talk_mask(g)[0]) is not None

    def forward_sequence(self, graphs, h0):
        """``T`` timesteps in one go: ``graphs`` is a list of T batched observation graphs with the same agent
        rows (or ONE graph that already is their ``batch`` plus ``T`` inferred from ``h0``).  Returns
        ``(q (T,N,A), h_last (N,H))`` — identical math to calling ``forward`` T times (the encoder does not depend on
        ``h``, reference ``gnn_agents.py:53``), but the encoder runs once over all T·N agent rows and the recurrent
        part is one persistent kernel forward and one backward."""
        from ..graph import batch as graph_batch
        gb = graph_batch(list(graphs)) if isinstance(graphs, (list, tuple)) else graphs
        N = h0.shape[0]
        TN = gb.num_nodes("agent")
        T = TN // N
        block, mask = self._talk_mask(gb)
        dims = self.fused_dims(block)
        if dims is None:
            raise NotImplementedError("forward_sequence needs the fused configuration (c in {None,'tarmac'}, Linear head)")
        xin = self._encode_pre(gb).view(T, N, dims.Fin)
        return self._run_sequence(dims, xin, h0.contiguous(), None if mask is None else mask.view(T, N))

    def _run_sequence(self, dims, xin, h0, mask):
        """Recurrent part of a whole window.  Resident-weight kernels (``ubs_agent_seq2_*``) when the recurrent
        weights fit shared memory (H = 64 configs), else the weight-streaming kernels (``ubs_agent_seq_*``)."""
        params = self._fused_params()
        T = xin.shape[0]
        grad = th.is_grad_enabled() and (xin.requires_grad or any(p is not None and p.requires_grad for p in params.values()))
        if self.use_seq2 and ops.seq2_supported(dims):
            if grad:
                q, h_last, _ = ops.AgentSequence2.apply(xin, h0, mask, dims, *[params[k] for k in ops.PARAM_ORDER])
                return q, h_last
            q, h_all = ops.agent_seq2_infer(dims, params, xin, h0, mask)
            return q, h_all[T - 1]
        packed = self._packed(dims, params)
        if grad:
            q, h_last, _ = ops.AgentSequence.apply(xin, h0, mask, dims, packed, *[params[k] for k in ops.PARAM_ORDER])
            return q, h_last
        q, h_all = ops.agent_seq_infer(dims, packed, xin, h0, mask)
        return q, h_all[T - 1]

    # ---- sequence-arena path: no graph objects at all (arena.py) ---------------------------------------------------
    def _arena_xin(self, arena, t0, T):
        """``[x_gt ‖ x_ubs] (T, N, 2H)`` for arena slots ``t0 .. t0+T-1``: ONE strided-segment launch per relation."""
        L, W = arena.layout, arena.layout.words
        convs = (self.enc.f_conv["seen"], self.enc.f_conv["near"])
        specs = [ops.RelSpec(arena.ptr("x_gt", t0), W, L.F_gt, arena.ptr("ip_seen", t0), W, T * L.cap_gt),
                 ops.RelSpec(arena.ptr("x_ubs", t0), W, L.F_ubs, arena.ptr("ip_near", t0), W, T * L.cap_ubs)]
        params = []
        for c in convs:
            params += [c.fc_src.weight, c.fc_src.bias, c.fc_dst.weight, c.fc_dst.bias, c.attn, c.res_fc.weight,
                       c.res_fc.bias]
        if not th.is_grad_enabled():
            # inside Function.forward the grad mode is always off, so the op decides from requires_grad what to save for
            # a backward: under no_grad (act step, target network) hand it detached parameters — no stats, no scores
            params = [p.detach() for p in params]
        c0 = convs[0]
        out = ops.SegmentEncode.apply(arena.buf, specs, arena.ptr("x_agent", t0), W, L.F_ag, T, L.N, c0._num_heads,
                                      c0._out_feats, c0._negative_slope, ops.GAT_RESIDUAL | ops.GAT_RELU, *params)
        return out.view(T, L.N, 2 * self._hidden_size)

    def arena_dims(self, arena):
        if not isinstance(self.enc, GraphObservationEncoder) and not arena.layout.flat_dim:
            return None
        return self.fused_dims(arena.layout.U)

    def _arena_flat_x(self, arena, t0, T):
        """MLP encoder (``DenseObservationEncoder``, reference ``gnn_agents.py:62-77``) over the flattened observations
        of arena slots ``t0 .. t0+T-1`` -> ``(T, N, H)``: plain library GEMMs (with autograd when enabled) reading
        the packets in place."""
        return self.enc.enc(arena.flat_obs(t0, T))

    def arena_sequence(self, arena, t0, T, h0):
        """``forward_sequence`` over arena slots ``t0 .. t0+T-1`` (with autograd when enabled).  ``arena`` may be a LIST of
        arenas (sampled replay windows, ``arena.ArenaReplay``): every window is encoded in place and the recurrent kernels
        run over the agent rows of all of them (``h0`` = their initial hidden states concatenated in list order)."""
        arenas = list(arena) if isinstance(arena, (list, tuple)) else [arena]
        dims = self.arena_dims(arenas[0])
        if dims is None:
            raise NotImplementedError("arena path needs the fused step configuration (and, for the MLP encoder, an arena "
                                      "with flattened observations)")
        graph_enc = isinstance(self.enc, GraphObservationEncoder)
        xs = [self._arena_xin(a, t0, T) if graph_enc else self._arena_flat_x(a, t0, T) for a in arenas]
        xin = xs[0] if len(xs) == 1 else th.cat(xs, 1)
        mask = None
        if dims.tarmac:
            ms = [a.sec("mask")[t0:t0 + T] for a in arenas]
            mask = (ms[0] if len(ms) == 1 else th.cat(ms, 1)).contiguous()
        return self._run_sequence(dims, xin, h0.contiguous(), mask)

    @th.no_grad()
    def arena_step(self, arena, t, q_out=None, explore=None, pdl=False):
        """Inference on slot t: reads ``arena.h[t]``, writes ``arena.h[t+1]`` and the greedy actions ``arena.acts[t]``;
        returns the Q values.  Three launches (two relations + the fused step), no allocation-dependent host logic,
        so the call can be captured in a CUDA graph."""
        dims = self.arena_dims(arena)
        mask = arena.sec("mask", t) if dims.tarmac else None
        if self.rel_act_supported(arena) and not (self.use_seq2_act and self.use_seq2):
            L, c0 = arena.layout, self.enc.f_conv["seen"]
            q = q_out.view(L.N, dims.A) if q_out is not None else th.empty(L.N, dims.A, dtype=th.float32, device=arena.device)
            ops.agent_act_rel(dims, self._packed(dims, self._fused_params()), self._relpacked(), arena.ptr("x_gt", t),
                              arena.ptr("ip_seen", t), L.F_gt, L.G, arena.ptr("x_ubs", t), arena.ptr("ip_near", t), L.F_ubs,
                              max(L.U - 1, 0), arena.ptr("x_agent", t), L.F_ag, c0._num_heads,
                              ops.GAT_RESIDUAL | ops.GAT_RELU | (ops.ACT_PDL if pdl else 0), arena.h[t], mask,
                              arena.h[t + 1], q, arena.acts[t], explore)
            return q
        xin = self._arena_xin(arena, t, 1) if isinstance(self.enc, GraphObservationEncoder) else self._arena_flat_x(arena, t, 1)
        if self.use_seq2_act and self.use_seq2 and ops.seq2_supported(dims):
            # two small library GEMMs (aggregator; fused [pv | pg]) + the resident-weight kernel with T = 1 + Q head
            q, h_all = ops.agent_seq2_infer(dims, self._fused_params(), xin, arena.h[t], mask)
            arena.h[t + 1].copy_(h_all[0])
            th.argmax(q[0], 1, out=arena.acts[t])
            if explore is not None:
                th.where(explore[0] <= explore[2], explore[1], arena.acts[t], out=arena.acts[t])
            if q_out is not None:
                q_out.copy_(q)
            return q[0]
        packed = self._packed(dims, self._fused_params())
        q, _, _ = ops.agent_seq_infer(dims, packed, xin, arena.h[t], mask, h_out=arena.h[t + 1].unsqueeze(0),
                                      q=q_out, acts=arena.acts[t].unsqueeze(0), explore=explore)
        return q[0]

    def forward(self, g, h):
        if not th.is_grad_enabled() and h.is_cuda:
            block, mask = self._talk_mask(g)
            dims = self.fused_dims(block)
            if dims is not None:
                packed = self._packed(dims, self._fused_params())
                xin = self._encode_pre(g).unsqueeze(0)
                q, h_all = ops.agent_seq_infer(dims, packed, xin, h.contiguous(), mask)
                return q[0], h_all[0]
        x = self.enc(g, g.ndata["feat"]).view(g.num_nodes("agent"), -1)
        h = self.f_comm(g["talk"], x, h) if self._comm_protocol is not None else self.rnn(x, h)
        return self.f_out(h), h


class DrqnGnnAgent(nn.Module):
    """Single-UBS agent (reference ``algos/drqn/agents/gnn_agents.py:9-30``): one GATv2 (gt→agent), GRU, Linear."""

    def __init__(self, obs_shape, n_actions, args):
        super().__init__()
        self._hidden_size, self._n_heads = args.hidden_size, args.n_heads
        feats_per_head = self._hidden_size // self._n_heads
        self.enc = GATv2Conv((obs_shape["gt"], obs_shape["agent"]), feats_per_head, self._n_heads, residual=True,
                             allow_zero_in_degree=True, activation=nn.ReLU())
        self.rnn = GRUCell(self._hidden_size, self._hidden_size)
        self.f_out = nn.Linear(self._hidden_size, n_actions)

    def init_hidden(self):
        return th.zeros(1, self._hidden_size)

    def forward(self, g, h):
        rel = g[g.canonical_etypes[0]]
        x = self.enc(rel, (g.nodes["gt"].data["feat"], g.nodes["agent"].data["feat"])).flatten(start_dim=1)
        h = self.rnn(x, h)
        return self.f_out(h), h
