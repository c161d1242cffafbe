"""GPU parity of the CUDA path against fixtures produced by the reference's OWN source (``tests/golden/ref_golden_v1.pt``,
generator ``tests/golden/make_ref_golden.py``: the unmodified ``/root/reference`` agent classes, learners, replay
buffer, env and wrapper with ``dgl`` replaced by a container).  Everything goes through the C ABI (``ops.py``).

Tolerances.  north_star: 1e-5 relative, fp32.  Per output:

* Q values / hidden states: ``allclose(rtol=1e-5, atol=1e-6 * max|ref|)`` against the reference's fp32 result — met by
  every case below — AND at least as close to the reference's fp64 result as its own fp32 run is (x4).
* parameter gradients: sums of 1e3..1e6 cancelling terms whose fp32 value depends on the summation order (the
  reference's own fp32 run is 1e-6..1e-4 away from its fp64 run on the small ones): ``rtol=1e-5`` + an absolute floor of
  ``1e-5 * max|grad of that parameter| + 2e-7 * max|grad of any parameter|`` against fp32, and the fp64 criterion (as
  accurate as the reference's fp32 run, x6).  The measured worst ratios are printed with ``-s``.
"""
import copy
from types import SimpleNamespace as SN

import pytest
import torch as th

from uav_bs_ctrl_b200 import agents as A, ops
from uav_bs_ctrl_b200 import learner as L
from uav_bs_ctrl_b200.arena import ObsPacket
from uav_bs_ctrl_b200.builder import build_obs_graph_batch, build_drqn_graph_batch
from helpers import (assert_close, assert_grads_close, assert_as_accurate, assert_param_sums, functional_inputs,
                     ref_golden, ref_agent_graph, ref_learner_batch, unpack64)

pytestmark = pytest.mark.gpu
DEV = "cuda"
AGENT_TAGS = [k.split("/", 1)[1] for k in ref_golden() if k.startswith("agent/")]
FUSED_TAGS = [t for t in AGENT_TAGS if ref_golden()[f"agent/{t}"]["args"]["c"] in (None, "tarmac")
              and ref_golden()[f"agent/{t}"]["args"]["n_rounds"] == 1 and not ref_golden()[f"agent/{t}"]["args"]["dueling"]]


def _agent(case):
    th.manual_seed(case["seed"])
    net = A.GnnAgent(case["obs_shape"], case["n_actions"], SN(**case["args"]))
    assert_param_sums(net.state_dict(), case["param_sums"], "product init vs reference init")
    return net.to(DEV)


def _feed_noise(net, case):
    if case["args"]["c"] == "disc":
        net.f_comm.exponential_feed = iter([e.to(DEV) for e in case["exponential"]])


def _check(case, q, h, grads, tag):
    r32, r64 = case["r32"], case["r64"]
    q6, h6 = unpack64(r32["q"], r64["q"]), unpack64(r32["h"], r64["h"])
    assert_close(q, r32["q"], rtol=1e-5, atol_scale=1e-6, what=f"{tag}: q vs reference fp32")
    assert_close(h, r32["h"], rtol=1e-5, atol_scale=1e-6, what=f"{tag}: h vs reference fp32")
    assert_as_accurate(q, r32["q"], q6, what=f"{tag}: q vs reference fp64", slack=4.0, floor_scale=2e-6)
    assert_as_accurate(h, r32["h"], h6, what=f"{tag}: h vs reference fp64", slack=4.0, floor_scale=2e-6)
    if grads is None:
        return
    assert_grads_close(grads, r32["grads"], rtol=1e-5, atol_scale=1e-5, floor_scale=2e-7, what=tag)
    worst = 0.0
    gmax = max(float(v.abs().max()) for v in r32["grads"].values())
    for k, g in grads.items():
        g6 = unpack64(r32["grads"][k], r64["grads"][k])
        assert_as_accurate(g, r32["grads"][k], g6, what=f"{tag}: grad {k} vs reference fp64", slack=6.0, floor_scale=5e-6)
        # mathematically-zero gradients (f_sign.bias) hold cancellation noise only: scale by the largest gradient then
        worst = max(worst, float((g.double().cpu() - g6).abs().max()) / max(float(g6.abs().max()), 1e-3 * gmax))
    print(f"[{tag}] worst |grad - ref64| / max|ref64| over parameters: {worst:.2e}; "
          f"q err vs fp64 {float((q.double().cpu() - q6).abs().max()):.2e} (reference fp32 run: {case['err64']:.2e})")


@pytest.mark.parametrize("tag", AGENT_TAGS)
def test_module_steps_match_reference_classes(tag):
    """``GnnAgent.forward`` called step by step with autograd (the path the reference's unmodified learner loop drives):
    Q, h_T and every parameter gradient vs the reference's ``GnnAgent``."""
    case = ref_golden()[f"agent/{tag}"]
    net = _agent(case)
    _feed_noise(net, case)
    N, H, T = case["B"] * case["U"], case["args"]["hidden_size"], case["T"]
    h0, w, hw = (t.to(DEV) for t in functional_inputs(case["seed"], N, H, T, case["n_actions"]))
    h, qs = h0, []
    for t in range(T):
        q, h = net(ref_agent_graph(case, t, DEV), h)
        qs.append(q)
    qs = th.stack(qs)
    ((qs * w).sum() + (qs ** 2).mean() + (h * hw).sum()).backward()
    _check(case, qs, h, {k: p.grad for k, p in net.named_parameters()}, tag)


@pytest.mark.parametrize("tag", AGENT_TAGS)
def test_inference_steps_match_reference_classes(tag):
    """The no-grad ``forward`` (``learner.act`` / target network: the fused act kernel where the configuration fits)."""
    case = ref_golden()[f"agent/{tag}"]
    net = _agent(case)
    _feed_noise(net, case)
    N, H, T = case["B"] * case["U"], case["args"]["hidden_size"], case["T"]
    h = functional_inputs(case["seed"], N, H, T, case["n_actions"])[0].to(DEV)
    ops.TIMER = ops.KernelTimer()
    qs = []
    with th.no_grad():
        for t in range(T):
            q, h = net(ref_agent_graph(case, t, DEV), h)
            qs.append(q)
    used = ops.TIMER.summary()
    ops.TIMER = None
    if tag in FUSED_TAGS:
        assert used.get("agent_seq_fwd", {}).get("count") == T, "the fused act kernel must be the path that ran"
    _check(case, th.stack(qs), h, None, tag)


@pytest.mark.parametrize("use_seq2", [True, False], ids=["resident", "streaming"])
@pytest.mark.parametrize("tag", FUSED_TAGS)
def test_window_kernels_match_reference_classes(tag, use_seq2):
    """``forward_sequence`` (one encoder launch per relation over all timesteps + one persistent recurrent kernel
    forward / backward) vs the reference's T separate ``GnnAgent.forward`` calls."""
    case = ref_golden()[f"agent/{tag}"]
    net = _agent(case)
    net.use_seq2 = use_seq2
    N, H, T = case["B"] * case["U"], case["args"]["hidden_size"], case["T"]
    h0, w, hw = (t.to(DEV) for t in functional_inputs(case["seed"], N, H, T, case["n_actions"]))
    ops.TIMER = ops.KernelTimer()
    q, h = net.forward_sequence([ref_agent_graph(case, t, DEV) for t in range(T)], h0)
    ((q * w).sum() + (q ** 2).mean() + (h * hw).sum()).backward()
    used = ops.TIMER.summary()
    ops.TIMER = None
    assert any(k in used for k in ("agent_seq2_bwd", "agent_seq_bwd")), "window kernels must be the path that ran"
    _check(case, q, h, {k: p.grad for k, p in net.named_parameters()}, f"{tag}/{'seq2' if use_seq2 else 'seq'}")


# ====================================================================================================== learner level
LEARNER_TAGS = [k.split("/", 1)[1] for k in ref_golden() if k.startswith("learner/") and "drqn" not in k]


def _graph_of(step, comm):
    if "flat" in step:
        B, U = step["flat"].shape[:2]
        g = build_obs_graph_batch(step["flat"], th.zeros(B, U, 0, 5), th.zeros(B, U, 0, 3), step["adj"] if comm else None)
    else:
        g = build_obs_graph_batch(step["agent"], step["gt"], step["ubs"], step["adj"] if comm else None)
    return g.to(DEV)


def _learner(case, n_envs=1, **kw):
    cfg = dict(case["config"])
    cfg.update(device=DEV, n_envs=n_envs, **kw)
    th.manual_seed(case["seed"])
    lr = L.MultiAgentQLearner(case["env_info"], SN(**cfg))
    assert_param_sums(lr.policy_net.state_dict(), case["policy0_sums"], "learner policy init vs reference")
    if lr.mixer is not None:
        lr.mixer.load_state_dict(case["mixer0"])
        lr.target_mixer.load_state_dict(case["mixer0"])
    return lr


def _check_update(case, lr, out, p0, tag):
    upd, cfg = case["updates"][0], case["config"]
    assert abs(out["LossQ"] - upd["LossQ"]) <= 1e-5 * abs(upd["LossQ"]), (out["LossQ"], upd["LossQ"])
    assert_close(th.as_tensor(out["QVals"]), upd["QVals"], rtol=1e-5, atol_scale=1e-6, what=f"{tag}: QVals")
    grads = {k: p.grad for k, p in lr.policy_net.named_parameters()}
    assert_grads_close(grads, upd["grads"], rtol=2e-5, atol_scale=1e-5, floor_scale=2e-7, what=f"{tag}: clipped")
    if lr.mixer is not None:
        assert_grads_close({k: p.grad for k, p in lr.mixer.named_parameters()}, upd["mixer_grads"], rtol=2e-5,
                           atol_scale=1e-5, floor_scale=2e-7, what=f"{tag}: mixer (unclipped, learner.py:159)")
    # AdamW step (learner.py:160).  The first Adam step is lr * g / (|g| + eps): where |g| >> eps it is ~lr * sign(g)
    # and must match tightly; entries with |g| ~ eps amplify 1e-9 gradient noise to a fraction of lr.
    lr_ = cfg["lr"]
    for k, p in lr.policy_net.named_parameters():
        ref_p, g = upd["policy"][k].to(DEV), upd["grads"][k].to(DEV)
        delta, ref_delta = p.detach() - p0[k], ref_p - p0[k]
        big = g.abs() > 1e-5
        assert float((delta - ref_delta).abs().max()) <= 2.01 * lr_, f"{tag}: AdamW step of {k}"
        if bool(big.any()):
            assert float((delta - ref_delta)[big].abs().max()) <= 2e-3 * lr_, f"{tag}: AdamW step of {k} (|g| > 1e-5)"
    # polyak target (learner.py:163-166): pinned through the reference's own per-parameter sums
    for k, v in lr.target_net.state_dict().items():
        s, a = upd["target_sums"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-5 * a + 2.01 * lr_ * (1 - cfg["polyak"]) * v.numel(), f"{tag}: target {k}"


@pytest.mark.parametrize("fused", [True, False], ids=["window", "stepwise"])
@pytest.mark.parametrize("tag", LEARNER_TAGS)
def test_update_matches_reference_learner(tag, fused):
    """The reference loop's transitions go through OUR ``cache`` (graph objects on the device); ``update`` on the
    sequences the reference sampled must give the reference's LossQ, QVals, clipped gradients, AdamW step and polyak
    target — with the window kernels (``fused``) and through per-step module calls."""
    case = ref_golden()[f"learner/{tag}"]
    cfg = case["config"]
    lr = _learner(case, fused=fused)
    p0 = {k: p.detach().clone() for k, p in lr.policy_net.named_parameters()}
    obs, comm = case["obs"], cfg["c"]
    graphs = {}
    g = lambda i: graphs.setdefault(i, _graph_of({k: v.unsqueeze(0) for k, v in obs[i].items()}, comm))
    for x in case["transitions"]:
        # act parity on the way (learner.act is the no-grad forward): logits via the policy net, h2 as the reference stored it
        with th.no_grad():
            q, h2 = lr.policy_net(g(x["obs"]), x["h"].to(DEV))
        assert_close(q, x["logits"], rtol=1e-5, atol_scale=1e-6, what=f"{tag}: act logits")
        assert_close(h2, x["next_h"], rtol=1e-5, atol_scale=1e-6, what=f"{tag}: act h")
        lr.cache(g(x["obs"]), x["h"].to(DEV), x["state"].to(DEV), x["act"].tolist(), x["rew"].numpy(), g(x["next_obs"]),
                 x["next_h"].to(DEV), x["next_state"].to(DEV), float(x["done"]), float(x["bad"]))
    assert len(lr.buffer) == len(case["transitions"]) // case["max_seq_len"]
    out = lr.update(samples=[lr.buffer.memory[i] for i in case["sample_idx"][0]])
    _check_update(case, lr, out, p0, tag)
    out2 = lr.update(samples=[lr.buffer.memory[i] for i in case["sample_idx"][1]])
    ref2 = case["updates"][1]["LossQ"]
    assert abs(out2["LossQ"] - ref2) <= 2e-3 * abs(ref2), "second update (after one AdamW / polyak step)"


@pytest.mark.parametrize("tag", [t for t in LEARNER_TAGS if ref_golden()[f"learner/{t}"]["max_seq_len"]
                                 == ref_golden()[f"learner/{t}"]["env_info"]["episode_limit"]])
def test_update_arena_matches_reference_learner(tag):
    """Same mini-batch through the sequence arena (no graph objects): the sampled sequences become the arena's env
    rows, ``update_arena`` must reproduce the reference ``update()``."""
    case = ref_golden()[f"learner/{tag}"]
    cfg, info = case["config"], case["env_info"]
    idx = case["sample_idx"][0]
    lr = _learner(case, n_envs=len(idx), cuda_graphs=False)
    if lr.policy_net.arena_dims(lr.new_arena(case["n_gts"])) is None:
        pytest.skip("configuration outside the fused arena path")
    p0 = {k: p.detach().clone() for k, p in lr.policy_net.named_parameters()}
    b = ref_learner_batch(case, idx)
    T = case["max_seq_len"]
    ar = lr.new_arena(case["n_gts"])
    tr = case["transitions"]
    for t in range(T + 1):
        s = b["steps"][t]
        rew = done = bad = None
        if t > 0:                                           # reward / done / bad of the transition that LED to slot t
            rew = th.stack([tr[k * T + t - 1]["rew"].float() for k in idx])
            done = th.tensor([float(tr[k * T + t - 1]["done"]) for k in idx])
            bad = th.tensor([float(tr[k * T + t - 1]["bad"]) for k in idx])
        pk = ObsPacket(ar.layout)
        if "flat" in s:
            B, U = s["flat"].shape[:2]
            pk.fill_from_dense(th.zeros(B, U, 2), th.zeros(B, U, case["n_gts"], 5), th.zeros(B, U, U - 1, 3), s["adj"], rew=rew,
                               done=done, bad=bad, state=None if b["states"] is None else b["states"][t])
            pk.sec("x_flat").view(B * U, ar.layout.flat_ld)[:, :ar.layout.flat_dim] = s["flat"].reshape(B * U, -1)
        else:
            pk.fill_from_dense(s["agent"], s["gt"], s["ubs"], s["adj"] if cfg["c"] else None, rew=rew, done=done, bad=bad,
                               state=None if b["states"] is None else b["states"][t])
        ar.load(t, pk)
    ar.h[0].copy_(b["h0"])
    ar.h[1].copy_(b["h1"])
    ar.acts[:T].copy_(b["acts"].squeeze(-1))
    out = lr.update_arena(ar)
    _check_update(case, lr, out, p0, f"{tag}/arena")


def test_qlearner_matches_reference_drqn():
    """exp1: the DRQN mirror (``QLearner``) driven with the reference loop's 8-argument ``cache`` calls."""
    case = ref_golden()["learner/exp1-drqn"]
    cfg, info = dict(case["config"]), case["env_info"]
    cfg.update(device=DEV)
    th.manual_seed(case["seed"])
    lr = L.QLearner(info, SN(**cfg))
    assert_param_sums(lr.policy_net.state_dict(), case["policy0_sums"], "QLearner init vs reference")
    p0 = {k: p.detach().clone() for k, p in lr.policy_net.named_parameters()}
    obs = case["obs"]
    g = lambda i: build_drqn_graph_batch(obs[i]["agent"], obs[i]["gt"].unsqueeze(0)).to(DEV)
    for x in case["transitions"]:
        with th.no_grad():
            q, h2 = lr.policy_net(g(x["obs"]), x["h"].to(DEV))
        assert_close(q, x["logits"], rtol=1e-5, atol_scale=1e-6, what="drqn act logits")
        assert_close(h2, x["next_h"], rtol=1e-5, atol_scale=1e-6, what="drqn act h")
        lr.cache(g(x["obs"]), x["h"].to(DEV), x["act"], x["rew"], g(x["next_obs"]), x["next_h"].to(DEV), float(x["done"]),
                 float(x["bad"]))
    out = lr.update(samples=[lr.buffer.memory[i] for i in case["sample_idx"][0]])
    upd = case["updates"][0]
    assert abs(out["LossQ"] - upd["LossQ"]) <= 1e-5 * abs(upd["LossQ"]), (out["LossQ"], upd["LossQ"])
    assert_close(th.as_tensor(out["QVals"]).reshape(upd["QVals"].shape), upd["QVals"], rtol=1e-5, atol_scale=1e-6, what="drqn QVals")
    assert_grads_close({k: p.grad for k, p in lr.policy_net.named_parameters()}, upd["grads"], rtol=2e-5, atol_scale=1e-5,
                       floor_scale=2e-7, what="drqn clipped")
    for k, p in lr.policy_net.named_parameters():
        big = upd["grads"][k].to(DEV).abs() > 1e-5
        d = (p.detach() - p0[k]) - (upd["policy"][k].to(DEV) - p0[k])
        assert float(d.abs().max()) <= 2.01 * cfg["lr"]
        if bool(big.any()):
            assert float(d[big].abs().max()) <= 2e-3 * cfg["lr"], f"drqn AdamW step of {k}"


@pytest.mark.parametrize("tag", ["exp3", "qmix"])
def test_replay_ring_update_matches_reference_learner(tag):
    """Device-resident replay (``ArenaReplay``): the reference loop's sequences are stored as single-env windows of a
    ring of arenas in the order the reference's ``ReplayBuffer`` received them; ``update_arena`` on the windows at the
    indices the reference sampled (``random.sample(memory, batch_size)``, ``buffer.py:37-39``) must reproduce the
    reference ``update()`` — off-policy replay with ``batch_size > 1`` on the fast path."""
    case = ref_golden()[f"learner/{tag}"]
    cfg, T = case["config"], case["max_seq_len"]
    lr = _learner(case, n_envs=1, cuda_graphs=False)
    p0 = {k: p.detach().clone() for k, p in lr.policy_net.named_parameters()}
    n_seq = len(case["transitions"]) // T
    replay = lr.new_replay(case["n_gts"], capacity=n_seq)
    tr = case["transitions"]
    for k in range(n_seq):
        b = ref_learner_batch(case, [k])
        ar = replay.next_arena()
        for t in range(T + 1):
            s = b["steps"][t]
            rew = done = bad = None
            if t > 0:
                rew = tr[k * T + t - 1]["rew"].float().unsqueeze(0)
                done = th.tensor([float(tr[k * T + t - 1]["done"])])
                bad = th.tensor([float(tr[k * T + t - 1]["bad"])])
            pk = ObsPacket(ar.layout).fill_from_dense(s["agent"], s["gt"], s["ubs"], s["adj"], rew=rew, done=done, bad=bad,
                                                     state=None if b["states"] is None else b["states"][t])
            ar.load(t, pk)
        ar.h[0].copy_(b["h0"])
        ar.h[1].copy_(b["h1"])
        ar.acts[:T].copy_(b["acts"].squeeze(-1))
        replay.commit()
    assert len(replay) == n_seq
    windows = replay.windows()
    out = lr.update_arena([windows[i] for i in case["sample_idx"][0]])
    _check_update(case, lr, out, p0, f"{tag}/replay")
    # the ring overwrites its oldest window and keeps the order of the reference's deque
    first = replay.next_arena()
    assert first is windows[0]
    replay.commit()
    assert replay.windows()[-1] is windows[0] and replay.windows()[0] is windows[1]
