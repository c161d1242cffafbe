"""QMIX mixer (reference algos/madrqn/agents/mixers.py) and its use in the TD loss (learner.py:144-152).

The golden vectors (tests/golden/mixer_golden_v1.pt) were produced by the reference's own QMixer class, loaded from
/root/reference by tests/golden/make_mixer_golden.py — this component is pinned by the reference itself."""
import os
from types import SimpleNamespace

import pytest
import torch as th

from uav_bs_ctrl_b200.agents import QMixer

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = th.load(os.path.join(HERE, "golden", "mixer_golden_v1.pt"), weights_only=False)


@pytest.mark.parametrize("tag", sorted(GOLD))
def test_qmixer_matches_the_reference_class(tag):
    d = GOLD[tag]
    S, U, E, L, B = d["dims"]
    m = QMixer(S, U, SimpleNamespace(embed_dim=E))
    assert set(m.state_dict()) == set(d["state_dict"])                  # checkpoint compatible
    m.load_state_dict(d["state_dict"])
    qs = d["qs"].clone().requires_grad_(True)
    y = m(qs, d["states"])
    assert y.shape == (L, B, 1)
    assert th.allclose(y, d["y"], rtol=1e-6, atol=1e-6)
    grads = th.autograd.grad(y, [qs] + list(m.parameters()), d["gy"])
    assert th.allclose(grads[0], d["grad_qs"], rtol=1e-5, atol=1e-6)
    for (k, _), g in zip(m.named_parameters(), grads[1:]):
        if k in d["grad_params"]:
            assert th.allclose(g, d["grad_params"][k], rtol=1e-5, atol=2e-6), k
    assert bool((grads[0] * 0 + th.autograd.grad(m(qs, d["states"]).sum(), qs)[0] >= 0).all())   # monotonic in every q_i


def _learner(mixer, **kw):
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    args = SimpleNamespace(device="cpu", o="mlp", c=None, share_reward=mixer, hidden_size=16, n_layers=1, n_heads=4,
                           msg_size=8, key_size=4, n_rounds=1, lr=1e-3, gamma=0.9, polyak=0.5, batch_size=1, replay_size=2,
                           max_seq_len=3, anneal_lr=False, double_q=True, dueling=False, mixer=mixer, embed_dim=8, n_envs=2)
    for k, v in kw.items():
        setattr(args, k, v)
    return MultiAgentQLearner(dict(obs_shape=5, state_shape=7, n_actions=4, n_agents=3, episode_limit=3), args)


def test_td_loss_with_mixer_follows_the_reference_formula():
    th.manual_seed(0)
    lr = _learner(True)
    T, B, U, A, S = 3, 2, 3, 4, 7
    agent_out = th.randn(T + 1, B * U, A, requires_grad=True)
    target_out = th.randn(T, B * U, A)
    acts = th.randint(0, A, (T, B * U, 1))
    rews, dones, states = th.randn(T, B, 1), th.zeros(T, B, 1), th.randn(T + 1, B, S)
    loss, qvals = lr._td_loss(agent_out, target_out, acts, rews, dones, states)
    # restatement of learner.py:134-154
    q = agent_out[:-1].gather(2, acts).view(T, B, U)
    nxt = target_out.gather(2, agent_out[1:].detach().argmax(2, keepdim=True)).view(T, B, U)
    q_tot, n_tot = lr.mixer(q, states[:-1]), lr.target_mixer(nxt, states[1:])
    ref = th.nn.functional.mse_loss(q_tot, rews + 0.9 * (1 - dones) * n_tot)
    assert th.allclose(loss, ref) and qvals.shape == (T, B, 1)
    loss.backward()
    assert all(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in lr.mixer.parameters())
    assert all(p.grad is None for p in lr.target_mixer.parameters())


def test_mixer_needs_shared_reward_and_state_size():
    with pytest.raises(ValueError, match="share_reward"):
        _learner(True, share_reward=False)


def test_polyak_moves_the_target_mixer_and_checkpoints_carry_it(tmp_path):
    th.manual_seed(1)
    lr = _learner(True)
    with th.no_grad():
        for p in lr.mixer.parameters():
            p.add_(1.0)
    before = [p.clone() for p in lr.target_mixer.parameters()]
    lr._optimise(sum((p ** 2).sum() for p in lr.params) * 0.0 + sum(p.sum() for p in lr.params) * 0.0, th.zeros(1), sync=False)
    for b, t, p in zip(before, lr.target_mixer.parameters(), lr.mixer.parameters()):
        assert th.allclose(t, 0.5 * b + 0.5 * p, atol=1e-5)
    path = str(tmp_path / "ckpt.pt")
    lr.save_checkpoint(path, dict(epoch=1, t=2))
    lr2 = _learner(True)
    lr2.load_checkpoint(path)
    for a, b in zip(lr.mixer.parameters(), lr2.mixer.parameters()):
        assert th.equal(a, b)


@pytest.mark.gpu
def test_full_loop_with_qmix_on_the_device_env():
    """reset -> rollout -> update with the mixer fed by the global states the env kernel wrote into the arena."""
    import numpy as np
    from uav_bs_ctrl_b200 import envs as E
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    B, T = 8, 5
    env = E.MultiUbsCoverageVecEnv("8ubs", B)
    info = env.get_env_info()
    info["episode_limit"] = T
    assert info["state_shape"] == 2 * 8 + 4 * 50
    args = SimpleNamespace(device="cuda", o="gnn", c="tarmac", share_reward=True, hidden_size=64, n_layers=2, n_heads=4,
                           msg_size=64, key_size=16, n_rounds=1, lr=2.5e-4, gamma=0.99, polyak=0.999, batch_size=1,
                           replay_size=2, max_seq_len=T, anneal_lr=False, double_q=True, dueling=False, mixer=True,
                           embed_dim=32, n_envs=B, cuda_graphs=True)
    th.manual_seed(0)
    learner = MultiAgentQLearner(info, args)
    arena = learner.new_arena(env.cfg.n_gts)
    assert arena.layout.state_dim == info["state_shape"]
    for cycle in range(2):
        learner.begin_sequence(arena)
        env.reset(arena, 0, seeds=range(cycle * B, (cycle + 1) * B))
        learner.rollout_arena(env, arena, 0.3)
        w0 = [p.detach().clone() for p in learner.mixer.parameters()]
        out = learner.update_arena(arena, sync=True)
        assert np.isfinite(out["LossQ"]) and out["QVals"].shape == (T, B, 1)
    st = arena.states(T + 1)
    assert float(st[..., :16].min()) >= 0 and float(st[..., :16].max()) <= 1          # normalised UBS positions
    assert any(not th.equal(a, b) for a, b in zip(w0, learner.mixer.parameters()))    # AdamW moved the mixer
