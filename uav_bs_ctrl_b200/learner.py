"""Learner mirrors: the callers of the hot path (reference ``algos/madrqn/learner.py:14-201``, ``algos/drqn/learner.py``).

Same constructor / ``init_hidden`` / ``act`` / ``cache`` / ``update`` / checkpoint surface as the reference's
``MultiAgentQLearner``, with three additions the north-star needs:

* **vector envs** — one graph carries ``n_envs`` env instances (``N = n_envs * n_agents`` agent rows); ε-greedy draws
  one uniform per env (the reference draws one per call for its single env, ``learner.py:75-78``);
* **device-resident replay** — ``cache`` keeps whatever graph it is handed (the device copy ``act`` staged), so
  ``update`` never re-uploads observations (reference: ``obs[t].to(device)`` at ``learner.py:116``);
* **data parallel** — a flat-bucket NCCL all-reduce of the policy gradients between ``loss.backward()`` and
  ``clip_grad_value_`` (``learner.py:158-159``), see ``dist.py``.

Loss math (``learner.py:118-154``), value clipping, AdamW, polyak target and checkpoint dict keys are the reference's.
"""
from __future__ import annotations

import random
import weakref
from collections import deque
import os
from copy import deepcopy
from typing import List

import torch as th
import torch.nn as nn
from torch.optim import AdamW

from . import dist, _lib
from .agents import REGISTRY as agent_REGISTRY, DRQN_REGISTRY
from .graph import HeteroGraph, batch as graph_batch

SCHEME = ("obs", "h", "state", "act", "rew", "done")


def cat(data_list):
    """Reference ``algos/common.py:40-47``: ``th.cat`` for tensors, graph batching for graphs."""
    if isinstance(data_list[0], th.Tensor):
        return th.cat(data_list)
    if isinstance(data_list[0], HeteroGraph):
        return graph_batch(data_list)
    raise TypeError("Unrecognised observation type.")


class ReplayBuffer:
    """Replay of fixed-length sequences (reference ``algos/madrqn/buffer.py:7-42``).  Entries may live on the GPU."""

    def __init__(self, capacity, max_seq_len, scheme=SCHEME):
        self.memory = deque(maxlen=capacity)
        self.max_seq_len, self.scheme = max_seq_len, scheme
        self.curr_seq = {k: [] for k in scheme}
        self.ptr = 0

    def push(self, transition: dict):
        for k, v in transition.items():
            if k in self.scheme:
                self.curr_seq[k].append(v)
        self.ptr += 1
        if self.ptr == self.max_seq_len:
            for k in ("obs", "h", "state"):
                if k in self.scheme:
                    self.curr_seq[k].append(transition.get("next_" + k))
            self.memory.append(self.curr_seq)
            self.curr_seq = {k: [] for k in self.scheme}
            self.ptr = 0

    def sample(self, batch_size: int):
        return random.sample(self.memory, batch_size)

    def __len__(self):
        return len(self.memory)


def _bump_versions(params):
    """The fused multi-tensor AdamW kernel writes the parameters without touching ``Tensor._version``, which every
    derived-weight cache of the package is keyed on (packed act-step weights, window weight layouts): bump it —
    bookkeeping only; if this torch has no such entry point, an in-place ``+= 0`` does the same with one kernel."""
    try:
        th._C._autograd._unsafe_set_version_counter(params, [p._version + 1 for p in params])
    except (AttributeError, TypeError):
        with th.no_grad():
            th._foreach_add_(params, 0.0)


class MultiAgentQLearner:
    """Multi-agent recurrent Q-learner (reference ``algos/madrqn/learner.py:14-201``), optionally with the QMIX mixer."""

    def __init__(self, env_info, args):
        self.args = args
        self.device = th.device(args.device)
        self.obs_shape, self.state_shape = env_info["obs_shape"], env_info.get("state_shape")
        self.n_actions, self.n_agents = env_info["n_actions"], env_info["n_agents"]
        self.n_envs = int(getattr(args, "n_envs", 1))

        self.policy_net = self._build_agent().to(self.device)
        self.target_net = self._build_agent().to(self.device)
        dist.sync_params(self.policy_net)
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.target_net.eval()
        self.params = list(self.policy_net.parameters())
        self._n_policy = sum(p.numel() for p in self.params)
        self.mixer = None
        if getattr(args, "mixer", False):                      # QMIX (reference learner.py:35-40)
            from .agents.mixers import QMixer
            if not getattr(args, "share_reward", False):
                raise ValueError("QMIX mixes the agents' values into one Q_tot: it needs share_reward=True "
                                 "(reference learner.py:151 expands the rewards to Q_tot's shape)")
            if not self.state_shape:
                raise ValueError("QMIX needs the global state size in env_info['state_shape']")
            if not hasattr(args, "embed_dim"):
                args.embed_dim = 32                           # reference config.py:19
            self.mixer = QMixer(self.state_shape, self.n_agents, args).to(self.device)
            dist.sync_params(self.mixer)
            self.target_mixer = deepcopy(self.mixer).to(self.device)
            self.params += list(self.mixer.parameters())

        self.max_seq_len = args.max_seq_len if args.max_seq_len is not None else env_info["episode_limit"]
        self.gamma, self.polyak, self.batch_size = args.gamma, args.polyak, args.batch_size
        self.buffer = ReplayBuffer(args.replay_size, self.max_seq_len)
        self.loss_fn = nn.MSELoss()
        # the reference's AdamW(lr) (learner.py:44); on the GPU its fused multi-tensor implementation: one kernel per
        # step for all parameters instead of a dozen foreach launches (same update rule, same state_dict layout)
        self.optimizer = AdamW(self.params, lr=args.lr, fused=bool(self.params[0].is_cuda))
        self.grad_bucket = dist.FlatGradBucket(self.params)
        self.anneal_lr = getattr(args, "anneal_lr", False)
        if self.anneal_lr:
            self.lr_scheduler = th.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=lambda e: max(0.4, 1 - e / 100))
        self.double_q = args.double_q
        self.fused = bool(getattr(args, "fused", True))      # sequence-fused update (False: per-step module calls)

    # ------------------------------------------------------------------------------------------ acting
    def init_hidden(self, batch_size=1):
        return self.policy_net.init_hidden().expand(self.n_agents * batch_size, -1)

    def _build_agent(self):
        if (self.args.o == "mlp") and (self.args.c is None):
            return agent_REGISTRY["rnn"](self.obs_shape, self.n_actions, self.args)
        return agent_REGISTRY["gnn"](self.obs_shape, self.n_actions, self.args)

    def stage(self, obs):
        """Host → device copy of an observation (asynchronous when the host graph is pinned)."""
        return obs.to(self.device, non_blocking=True)

    def act(self, obs, h, eps_thres):
        """ε-greedy action selection (reference ``learner.py:69-80``).  Single env: returns ``(list, h)`` exactly
        like the reference.  Vector envs: returns a device tensor of ``n_envs * n_agents`` actions."""
        obs, h = obs.to(self.device, non_blocking=True), h.to(self.device)
        with th.no_grad():
            logits, h = self.policy_net(obs, h)
        if self.n_envs == 1:
            if random.random() > eps_thres:
                acts = th.argmax(logits, 1)
            else:
                acts = th.randint(self.n_actions, size=(self.n_agents,), dtype=th.long)
            return acts.tolist(), h
        greedy = th.argmax(logits, 1)
        explore = (th.rand(self.n_envs, device=self.device) <= eps_thres).repeat_interleave(self.n_agents)
        rnd = th.randint(self.n_actions, size=greedy.shape, device=self.device)
        return th.where(explore, rnd, greedy), h

    def cache(self, obs, h, state, act, rew, next_obs, next_h, next_state, done, bad_mask):
        """Reference ``learner.py:82-92``.  ``done`` / ``bad_mask`` are scalars (single env) or ``(n_envs,)`` tensors."""
        as_t = lambda x, dt: x.to(dt) if isinstance(x, th.Tensor) else th.tensor(x, dtype=dt)
        rew = as_t(rew, th.float32)
        if self.args.share_reward:
            rew = rew.reshape(self.n_envs, -1).mean(1, keepdim=True)
        done_t = as_t(done, th.float32).reshape(self.n_envs, 1)
        bad_t = as_t(bad_mask if bad_mask is not None else 0, th.float32).reshape(-1, 1)
        keep = (1 - done_t).repeat_interleave(self.n_agents, 0).to(next_h.device)
        transition = dict(obs=obs, h=h, state=state, act=as_t(act, th.long).reshape(-1, 1),
                          rew=rew.reshape(self.n_envs, -1), next_obs=next_obs, next_h=keep * next_h,
                          next_state=next_state, done=(1 - bad_t) * done_t)
        self.buffer.push(transition)

    # ------------------------------------------------------------------------------------------ learning
    def _unroll(self, obs: List[HeteroGraph], h, h_targ):
        """Reference ``learner.py:118-129``: T policy steps with grad + T target steps without + one more policy step."""
        agent_out, target_out = [], []
        T = self.max_seq_len
        can_fuse = getattr(self.policy_net, "can_fuse", None)
        if self.fused and can_fuse is not None and can_fuse(obs[0]) and h.is_cuda:
            # same math, but the encoder runs once over all timesteps and the recurrence is one persistent kernel
            agent_out, _ = self.policy_net.forward_sequence(obs, h)
            with th.no_grad():
                target_out, _ = self.target_net.forward_sequence(obs[1:], h_targ)
            return agent_out, target_out
        for t in range(T):
            logits, h = self.policy_net(obs[t], h)
            agent_out.append(logits)
            with th.no_grad():
                next_logits, h_targ = self.target_net(obs[t + 1], h_targ)
                target_out.append(next_logits)
        logits, h = self.policy_net(obs[T], h)
        agent_out.append(logits)
        return th.stack(agent_out), th.stack(target_out)

    def compute_loss(self, obs, h, h_targ, acts, rews, dones, states=None):
        agent_out, target_out = self._unroll(obs, h, h_targ)
        return self._td_loss(agent_out, target_out, acts, rews, dones, states)

    def _td_loss(self, agent_out, target_out, acts, rews, dones, states=None):
        """Reference ``learner.py:134-154``: ``agent_out (T+1,N,A)``, ``target_out (T,N,A)``; with the mixer,
        ``states (T+1, n_seq, S)`` turn the per-agent values into ``Q_tot`` (``:144-148``)."""
        T = target_out.shape[0]
        qvals = agent_out[:-1].gather(2, acts)
        if not self.double_q:
            next_vals = target_out.max(2, keepdim=True)[0]
        else:
            next_acts = th.argmax(agent_out[1:].detach(), 2, keepdim=True)
            next_vals = target_out.gather(2, next_acts)
        n_seq = rews.shape[1]
        qvals = qvals.view(T, n_seq, self.n_agents)
        next_vals = next_vals.view(T, n_seq, self.n_agents)
        if self.mixer is not None:
            if states is None:
                raise ValueError("QMIX update without global states")
            qvals = self.mixer(qvals, states[:-1])
            with th.no_grad():
                next_vals = self.target_mixer(next_vals, states[1:])
        rews, dones = rews.expand_as(next_vals), dones.expand_as(next_vals)
        target_qvals = rews + self.gamma * (1 - dones) * next_vals
        return self.loss_fn(qvals, target_qvals), qvals

    def gather_batch(self, samples):
        """Reference ``learner.py:99-116``: per-timestep ``cat`` over the sampled sequences, then move to device."""
        T = self.max_seq_len
        batch = {k: [] for k in self.buffer.scheme}
        keys = [k for k in batch if samples[0][k] and samples[0][k][0] is not None]
        for t in range(T):
            for k in keys:
                batch[k].append(cat([s[k][t] for s in samples]))
        for k in ("obs", "h", "state"):
            if k in keys:
                batch[k].append(cat([s[k][T] for s in samples]))
        dev = self.device
        acts = th.stack(batch["act"]).to(dev)
        rews = th.stack(batch["rew"]).to(dev)
        dones = th.stack(batch["done"]).to(dev)
        h, h_targ = batch["h"][0].to(dev), batch["h"][1].to(dev)
        obs = [o.to(dev) for o in batch["obs"]]
        self._batch_states = th.stack(batch["state"]).to(dev) if batch.get("state") else None   # (T+1, n_seq·n_envs, S)
        return obs, h, h_targ, acts, rews, dones

    def update(self, samples=None, sync=True):
        """One BPTT update (reference ``learner.py:94-173``).  ``samples`` overrides the random replay draw."""
        if samples is None:
            assert len(self.buffer) >= self.batch_size, "Insufficient samples for update."
            samples = self.buffer.sample(self.batch_size)
        obs, h, h_targ, acts, rews, dones = self.gather_batch(samples)
        loss, qvals = self.compute_loss(obs, h, h_targ, acts, rews, dones, self._batch_states)
        return self._optimise(loss, qvals, sync)

    def _optimise(self, loss, qvals, sync):
        """Reference ``learner.py:157-173``: backward, (DP all-reduce,) value clip, AdamW, polyak target."""
        self._backward_gather(loss)
        return self._apply_gradients(loss, qvals, sync)

    def _backward_gather(self, loss):
        self.grad_bucket.release()                                           # autograd assigns instead of accumulating ...
        loss.backward()
        self.grad_bucket.gather()                                            # ... and one cat fills the flat bucket

    def _apply_gradients(self, loss, qvals, sync):
        dist.avg_grads(self.grad_bucket)                                     # DP: one flat all-reduce
        # clip_grad_value_(self.policy_net.parameters(), 1) (reference learner.py:159): the policy parameters come first in
        # the flat bucket; the QMIX mixer's gradients stay unclipped like in the reference
        self.grad_bucket.flat[:self._n_policy].clamp_(-1.0, 1.0)
        self.optimizer.step()
        _bump_versions(self.params)
        with th.no_grad():
            pp, tp = list(self.policy_net.parameters()), list(self.target_net.parameters())
            if self.mixer is not None:                                       # reference learner.py:168-171
                pp, tp = pp + list(self.mixer.parameters()), tp + list(self.target_mixer.parameters())
            th._foreach_mul_(tp, self.polyak)
            th._foreach_add_(tp, pp, alpha=1 - self.polyak)
        self._refresh_packed()
        if sync:
            return dict(LossQ=loss.item(), QVals=qvals.detach().cpu().numpy())
        return dict(LossQ=loss.detach(), QVals=qvals.detach())

    def _refresh_packed(self):
        """AdamW / polyak / checkpoint loading wrote the parameters in place: bring the packed act-step copies (read by
        address from captured CUDA graphs) up to date."""
        for net in (self.policy_net, self.target_net):
            if hasattr(net, "refresh_packed"):
                net.mark_params_changed()                 # the fused AdamW kernel leaves Tensor._version alone
                net.refresh_packed()

    # ------------------------------------------------------------------------------------------ sequence-arena path
    def new_arena(self, n_gts, n_slots=None):
        """Device-resident replay entry: ``max_seq_len + 1`` observation packets + hidden states + actions."""
        from .arena import PacketLayout, SequenceArena
        graph_obs = isinstance(self.obs_shape, dict)
        fg = self.obs_shape["gt"] if graph_obs else getattr(self.args, "F_gt", 4)
        L = PacketLayout(self.n_envs, self.n_agents, n_gts, F_ag=self.obs_shape["agent"] if graph_obs else 2, F_gt=fg,
                         F_ubs=self.obs_shape["ubs"] if graph_obs else 2,
                         state_dim=int(self.state_shape or 0) if self.mixer is not None else 0,
                         flat_dim=0 if graph_obs else int(self.obs_shape))
        return SequenceArena(L, n_slots or self.max_seq_len + 1, self.args.hidden_size, self.device)

    def begin_sequence(self, arena, h0=None):
        """Start of a T-step window on ``arena``: initial hidden state and the window's exploration noise."""
        if h0 is None:
            arena.h[0].zero_()                       # init_hidden() is zeros (reference gnn_agents.py:48-49)
        else:
            arena.h[0].copy_(h0)
        if not hasattr(arena, "explore_u"):          # fixed addresses: captured act graphs read these buffers
            arena.explore_u = th.empty(arena.S, arena.layout.N, device=self.device)
            arena.explore_a = th.empty(arena.S, arena.layout.N, dtype=th.int64, device=self.device)
        u = th.rand(arena.S, self.n_envs, 1, device=self.device)
        arena.explore_u.copy_(u.expand(-1, -1, self.n_agents).reshape(arena.S, -1))
        arena.explore_a.random_(0, self.n_actions)

    def _act_arena_eager(self, arena, t, pdl=False):
        # ε-greedy is fused into the step kernel: one uniform per env (explore_u holds it repeated for the env's agents).
        # pdl: inside a captured rollout the step may launch programmatically behind the previous kernel (UBS_ACT_PDL)
        return self.policy_net.arena_step(arena, t, explore=(arena.explore_u[t], arena.explore_a[t], self._eps_dev), pdl=pdl)

    def act_arena(self, arena, t, eps_thres):
        """``act`` on arena slot t (observation already staged with ``arena.load``): writes ``arena.h[t+1]`` and the
        ε-greedy actions ``arena.acts[t]`` (returned, on the device).  With ``args.cuda_graphs`` the three kernels and
        the ε-greedy ops of every slot are captured once and replayed."""
        self._set_eps(eps_thres)
        if not getattr(self.args, "cuda_graphs", False):
            self._act_arena_eager(arena, t)
            return arena.acts[t]
        # captured graphs hold raw device addresses of the arena (and env) they were captured on: the cache lives and
        # dies with those objects (weak keys) instead of being keyed by id(), which a later object could reuse
        graphs = self._graph_cache(arena)
        key = ("act", t)
        g = graphs.get(key)
        # packed weights are read by address inside the graph: re-pack (outside any capture) if a parameter changed
        # since — optimizer step, load_checkpoint, load_state_dict, a replay-buffer update()
        self.policy_net.refresh_packed(arena)
        if g is None:
            side = th.cuda.Stream()
            side.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(side):
                self._act_arena_eager(arena, t)          # warm-up outside capture
            th.cuda.current_stream().wait_stream(side)
            g = th.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with th.cuda.graph(g):
                self._act_arena_eager(arena, t)
            g = (g, _lib.launch_count() - n0)             # library kernels inside the graph (for the launch counter)
            graphs[key] = g
        g[0].replay()
        _lib.add_launches(g[1])
        return arena.acts[t]

    def _set_eps(self, eps_thres):
        if not hasattr(self, "_eps_dev"):
            self._eps_dev = th.zeros((), device=self.device)
            self._eps_host = None
        if self._eps_host != eps_thres:
            self._eps_dev.fill_(float(eps_thres))
            self._eps_host = eps_thres

    def _graph_cache(self, arena, env=None):
        if not hasattr(self, "_graphs"):
            self._graphs = weakref.WeakKeyDictionary()
        per_arena = self._graphs.setdefault(arena, {})
        if env is None:
            return per_arena.setdefault("self", {})
        if "env" not in per_arena:
            per_arena["env"] = weakref.WeakKeyDictionary()
        return per_arena["env"].setdefault(env, {})

    def rollout_arena(self, env, arena, eps_thres):
        """One whole episode window on the device env (``envs.MultiUbsCoverageVecEnv``): for t in 0..T-1 the fused act
        step on slot t, then ``env.step`` writing observation / reward / done of slot t+1 — the reference's
        ``act -> env.step -> cache`` loop (``algos/madrqn/run.py:81-90``) with nothing returning to the host.  With
        ``args.cuda_graphs`` all 5·T kernels are captured once per (arena, env) and replayed as ONE graph launch.
        ``env=None``: the T act steps alone, on observations that are already resident in the arena (replayed
        episodes) — the same graph without the env kernels."""
        T = self.max_seq_len
        if env is not None and T > getattr(env, "episode_limit", T):
            # the device env flags done at t == episode_limit and never resets itself or the hidden state (the reference
            # loop does both, run.py:91-94): a longer window would keep stepping a finished episode.  A shorter window is
            # the prefix of an episode; either way the caller resets the env (and begin_sequence) before every window.
            raise ValueError(f"rollout_arena: max_seq_len ({T}) exceeds the env's episode_limit ({env.episode_limit})")
        self._set_eps(eps_thres)

        use_pdl = bool(getattr(self.args, "act_pdl", True))

        def run(captured=False):
            for t in range(T):
                # inside the captured graph the packed weights are constant (refreshed before the replay) and the kernel
                # ahead of step t >= 1 is the previous act step or the env's pack kernel: neither writes them
                self._act_arena_eager(arena, t, pdl=captured and use_pdl and t > 0)
                if env is not None:
                    env.step(arena, t)

        from . import ops
        if not getattr(self.args, "cuda_graphs", False) or ops.TIMER is not None:
            run()
            return
        graphs = self._graph_cache(arena, env)
        key = "rollout"
        g = graphs.get(key)
        self.policy_net.refresh_packed(arena)
        if g is None:
            snap = env.buf.snapshot() if env is not None else None   # the warm-up step must not advance the episode
            side = th.cuda.Stream()
            side.wait_stream(th.cuda.current_stream())
            with th.cuda.stream(side):
                self._act_arena_eager(arena, 0)          # warm-up outside capture (slot 1 is rewritten by the replay)
                if env is not None:
                    env.step(arena, 0)
            th.cuda.current_stream().wait_stream(side)
            if env is not None:
                env.buf.restore(snap)
            g = th.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with th.cuda.graph(g):
                run(captured=True)
            g = (g, _lib.launch_count() - n0)
            graphs[key] = g
        g[0].replay()
        _lib.add_launches(g[1])

    def new_replay(self, n_gts, capacity=None):
        """Device-resident replay ring of ``capacity`` (default ``args.replay_size``) sequence arenas."""
        from .arena import ArenaReplay
        return ArenaReplay(lambda: self.new_arena(n_gts), capacity or self.args.replay_size)

    def update_arena(self, arena, sync=True):
        """One BPTT update on the window held by ``arena`` — or on a LIST of windows sampled from an ``ArenaReplay``
        (``buffer.sample(batch_size)`` + the per-timestep ``cat`` of reference ``learner.py:99-116``, as pointer
        selection).  Same math as ``update``; no graph objects, no re-batching: one strided-segment encoder launch per
        relation and window over all T+1 timesteps, and one persistent recurrent kernel over the agent rows of all
        windows, for the policy (with grad) and for the target network."""
        from . import ops
        arenas = list(arena) if isinstance(arena, (list, tuple)) else [arena]
        # one arena, fixed addresses: the update replays as a CUDA graph (UBS_UPDATE_GRAPH=0 or args.update_graph = False:
        # eager launches)
        if (len(arenas) == 1 and arenas[0].h.is_cuda and getattr(self.args, "cuda_graphs", False)
                and getattr(self.args, "update_graph", True) and ops.TIMER is None
                and os.environ.get("UBS_UPDATE_GRAPH", "1") != "0"):
            out = self._update_arena_graphed(arenas[0], sync)
            if out is not None:
                return out
        loss, qvals = self._arena_loss(arenas)
        return self._optimise(loss, qvals, sync)

    def _update_arena_graphed(self, arena, sync):
        """``update_arena`` with forward + TD loss + backward + gradient gather replayed as ONE CUDA graph (captured once
        per arena: every operand lives at a fixed address — the arena, the parameters, the flat gradient bucket); the
        all-reduce, clip, AdamW, polyak and re-pack follow eagerly (a dozen launches).  The update is then immune to a
        slow host: eagerly it is ≈ 180 launches that a busy CPU cannot enqueue as fast as the device retires them.
        Returns None (and remembers it) if this configuration cannot be captured."""
        from . import ops
        graphs = self._graph_cache(arena)
        ent = graphs.get("update")
        if ent is False:
            return None
        if ent is None:
            cur = th.cuda.current_stream()
            try:
                side = th.cuda.Stream()
                side.wait_stream(cur)
                with th.cuda.stream(side):                 # warm-up outside capture: lazy inits, workspaces; no parameter changes
                    loss, qvals = self._arena_loss([arena])
                    self._backward_gather(loss)
                    del loss, qvals                         # nothing of the warm-up's autograd graph survives into the capture
                cur.wait_stream(side)
                th.cuda.synchronize()
                # Everything derived from the parameters (window weight layouts, packed weights) must be REBUILT INSIDE the
                # capture: a cache hit here would bake the warm-up's buffers — stale after the first optimizer step, and
                # freed below — into the graph.
                self._invalidate_derived_weights()
                g = th.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                if getattr(self, "_update_pool", None) is None:
                    self._update_pool = th.cuda.graph_pool_handle()   # the update graphs of several arenas (a replay ring)
                with th.cuda.graph(g, pool=self._update_pool):        # share ONE activation pool: they never overlap
                    loss, qvals = self._arena_loss([arena])
                    self._backward_gather(loss)
                # keep the static outputs, not their autograd graph (it would pin the AccumulateGrad nodes of this capture)
                ent = (g, _lib.launch_count() - n0, loss.detach(), qvals.detach())
                del loss, qvals
            except Exception as e:                         # e.g. an op of this configuration that syncs with the host
                import warnings
                th.cuda.synchronize()
                warnings.warn(f"update_arena: CUDA-graph capture of the update failed ({type(e).__name__}: {e}); running eagerly")
                graphs["update"] = False
                ent = None
            finally:
                # derived-weight caches filled during the capture hold buffers whose contents only exist after a replay
                self._invalidate_derived_weights()
            if ent is None:
                return None
            graphs["update"] = ent
        ent[0].replay()
        _lib.add_launches(ent[1])
        # the graph's outputs live at fixed addresses that the next replay overwrites: hand out copies
        return self._apply_gradients(ent[2].detach().clone(), ent[3].detach().clone(), sync)

    def _invalidate_derived_weights(self):
        from . import ops
        ops._SEQ2_CACHE.clear()
        for net in (self.policy_net, self.target_net):
            if hasattr(net, "mark_params_changed"):
                net.mark_params_changed()

    def _arena_loss(self, arenas):
        """TD loss of the window(s): policy window with grad, target window without (reference ``learner.py:118-154``)."""
        from . import ops
        T = self.max_seq_len
        cat = (lambda xs, d: xs[0] if len(xs) == 1 else th.cat(xs, d))
        acts = cat([a.acts[:T] for a in arenas], 1).unsqueeze(-1)
        rews, dones = cat([a.rewards(T) for a in arenas], 1), cat([a.dones(T) for a in arenas], 1)
        if self.args.share_reward:
            rews = rews.mean(2, keepdim=True)
        # next_h = (1 - done) * next_h  (reference cache(), learner.py:90)
        keep = cat([(1 - a.sec("done", 1)).repeat_interleave(self.n_agents) for a in arenas], 0).unsqueeze(1)
        h0, h_targ = cat([a.h[0] for a in arenas], 0), cat([a.h[1] for a in arenas], 0) * keep
        # (small windows are bound by the host enqueueing launches: a second stream only adds waits there — measured at
        # 32 envs per GPU: 3.18 ms in order, 3.53 ms overlapped)
        big = h0.shape[0] * (T + 1) >= 32768
        if h0.is_cuda and big and getattr(self.args, "overlap_target", True) and ops.TIMER is None:
            # The target window (no grad) is independent of the policy window: run it on a second stream.  Its kernels
            # interleave with the policy's on the SMs — the GATv2 kernels are FP32-issue bound, the recurrent window
            # kernels tensor-pipe bound, the projections wait on HBM — instead of queueing behind them.
            cur = th.cuda.current_stream()
            if getattr(self, "_side_stream", None) is None:
                self._side_stream = th.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
            with th.cuda.stream(side), th.no_grad():
                target_out, _ = self.target_net.arena_sequence(arenas, 1, T, h_targ)
            agent_out, _ = self.policy_net.arena_sequence(arenas, 0, T + 1, h0)
            cur.wait_stream(side)
            if not th.cuda.is_current_stream_capturing():
                target_out.record_stream(cur)         # allocated on the side stream, consumed (and freed) on this one
        else:
            agent_out, _ = self.policy_net.arena_sequence(arenas, 0, T + 1, h0)
            with th.no_grad():
                target_out, _ = self.target_net.arena_sequence(arenas, 1, T, h_targ)
        states = cat([a.states(T + 1) for a in arenas], 1) if self.mixer is not None else None
        return self._td_loss(agent_out, target_out, acts, rews, dones, states)

    # ------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, path, stamp):
        checkpoint = dict(stamp)
        checkpoint["model_state_dict"] = self.policy_net.state_dict()
        checkpoint["optimizer_state_dict"] = self.optimizer.state_dict()
        if self.mixer is not None:
            checkpoint["mixer_state_dict"] = self.mixer.state_dict()
        if self.anneal_lr:
            checkpoint["lr_scheduler_state_dict"] = self.lr_scheduler.state_dict()
        th.save(checkpoint, path)

    def load_checkpoint(self, path):
        checkpoint = th.load(path, map_location=self.device)
        stamp = dict(epoch=checkpoint["epoch"], t=checkpoint["t"])
        self.policy_net.load_state_dict(checkpoint["model_state_dict"])
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if self.mixer is not None:
            self.mixer.load_state_dict(checkpoint["mixer_state_dict"])
            self.target_mixer.load_state_dict(self.mixer.state_dict())
        if self.anneal_lr:
            self.lr_scheduler.load_state_dict(checkpoint["lr_scheduler_state_dict"])
        self._refresh_packed()
        return stamp


class QLearner(MultiAgentQLearner):
    """Single-agent DRQN learner (reference ``algos/drqn/learner.py:15-150``): one agent, ``max`` target, no double-Q."""

    def __init__(self, env_info, args):
        info = dict(env_info)
        info.setdefault("n_agents", 1)
        info.setdefault("state_shape", None)
        for k, v in dict(o="gnn", c=None, share_reward=False, double_q=False, dueling=False, mixer=False).items():
            if not hasattr(args, k):
                setattr(args, k, v)
        super().__init__(info, args)

    def _build_agent(self):
        kind = "rnn" if isinstance(self.obs_shape, int) else "gnn"
        return DRQN_REGISTRY[kind](self.obs_shape, self.n_actions, self.args)

    def act(self, obs, h, eps_thres):
        acts, h = super().act(obs, h, eps_thres)
        return (acts[0] if isinstance(acts, list) else acts), h

    def cache(self, obs, h, act, rew, next_obs, next_h, done, bad_mask):
        """Reference ``algos/drqn/learner.py:67-73``: the DRQN loop (``algos/drqn/run.py:82``) passes 8 arguments —
        no global state."""
        super().cache(obs, h, None, act, rew, next_obs, next_h, None, done, bad_mask)
