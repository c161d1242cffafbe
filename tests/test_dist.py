"""CPU, world_size 2, gloo: flat-bucket gradient averaging == full-batch gradient; parameter broadcast."""
import os
import socket

import pytest
import torch as th
import torch.distributed as td
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model(seed):
    th.manual_seed(seed)
    return nn.Sequential(nn.Linear(6, 16), nn.Tanh(), nn.GRUCell(16, 8)) if False else nn.Sequential(
        nn.Linear(6, 16), nn.Tanh(), nn.Linear(16, 3))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from uav_bs_ctrl_b200 import dist
    dist.init_from_env("gloo")
    assert dist.world_size() == world and dist.rank() == rank
    model = _make_model(seed=100 + rank)                # different init per rank on purpose
    dist.sync_params(model)                             # ... rank 0's parameters win
    bucket = dist.FlatGradBucket(model.parameters())
    th.manual_seed(0)
    x, y = th.randn(8, 6), th.randn(8, 3)               # global batch; rank r owns rows [4r, 4r+4)
    xs, ys = x[4 * rank:4 * rank + 4], y[4 * rank:4 * rank + 4]
    bucket.zero_()
    nn.functional.mse_loss(model(xs), ys).backward()
    bucket.rebind()
    dist.avg_grads(bucket)
    # module-level variant (per-call flatten) must agree
    model2 = _make_model(seed=100)
    nn.functional.mse_loss(model2(xs), ys).backward()
    dist.avg_grads(model2)
    mx = dist.all_reduce_max_scalar(float(rank + 1), "cpu")
    th.save(dict(flat=bucket.flat.clone(), params=[p.detach().clone() for p in model.parameters()],
                 flat2=th.cat([p.grad.reshape(-1) for p in model2.parameters()]), mx=mx),
            os.path.join(out_dir, f"r{rank}.pt"))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_bucket_allreduce_matches_full_batch(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = th.load(tmp_path / "r0.pt"), th.load(tmp_path / "r1.pt")
    ref = _make_model(seed=100)
    for p, q0, q1 in zip(ref.parameters(), r0["params"], r1["params"]):
        assert th.equal(p, q0) and th.equal(p, q1), "sync_params must broadcast rank 0's parameters"
    th.manual_seed(0)
    x, y = th.randn(8, 6), th.randn(8, 3)
    nn.functional.mse_loss(ref(x), y).backward()        # equal shard sizes + mean loss => averaged grads == global grads
    want = th.cat([p.grad.reshape(-1) for p in ref.parameters()])
    assert th.allclose(r0["flat"], want, rtol=1e-5, atol=1e-7) and th.equal(r0["flat"], r1["flat"])
    assert th.allclose(r0["flat2"], want, rtol=1e-5, atol=1e-7)
    assert r0["mx"] == r1["mx"] == 2.0


def test_bucket_views_survive_zero_and_rebind():
    from uav_bs_ctrl_b200 import dist
    m = _make_model(0)
    b = dist.FlatGradBucket(m.parameters())
    m(th.randn(2, 6)).sum().backward()
    g = th.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert th.equal(g, b.flat) and all(p.grad.data_ptr() >= b.flat.data_ptr() for p in m.parameters())
    for p in m.parameters():
        p.grad = None                                     # e.g. optimizer.zero_grad(set_to_none=True)
    m(th.randn(2, 6)).sum().backward()
    b.rebind()
    assert th.equal(th.cat([p.grad.reshape(-1) for p in m.parameters()]), b.flat)
    b.zero_()
    assert float(b.flat.abs().sum()) == 0 and all(float(p.grad.abs().sum()) == 0 for p in m.parameters())


def _learner_worker(rank, world, port, out_dir):
    """Two ranks, different seeds and different local batches: after init and after one QMIX update the policy and the
    mixer must be identical on both ranks (sync_params at construction, ONE flat all-reduce over policy + mixer
    gradients per update)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from types import SimpleNamespace
    from uav_bs_ctrl_b200 import dist
    from uav_bs_ctrl_b200.learner import MultiAgentQLearner
    dist.init_from_env("gloo")
    th.manual_seed(1000 + rank)
    args = SimpleNamespace(device="cpu", o="mlp", c=None, share_reward=True, hidden_size=16, n_layers=1, n_heads=4,
                           msg_size=8, key_size=4, n_rounds=1, lr=1e-2, gamma=0.9, polyak=0.5, batch_size=1, replay_size=2,
                           max_seq_len=3, anneal_lr=False, double_q=True, dueling=False, mixer=True, embed_dim=8, n_envs=2)
    lr = MultiAgentQLearner(dict(obs_shape=5, state_shape=7, n_actions=4, n_agents=3, episode_limit=3), args)
    init = [p.detach().clone() for p in lr.params]
    T, B, U, A, S = 3, 2, 3, 4, 7
    g = th.Generator().manual_seed(rank)                                     # rank-local data
    # the recurrent core is a CUDA kernel (no CPU fallback): drive the Q head with rank-local features instead
    agent_out = lr.policy_net.f_out(th.randn(T + 1, B * U, 16, generator=g))
    target_out = th.randn(T, B * U, A, generator=g)
    acts = th.randint(0, A, (T, B * U, 1), generator=g)
    loss, qv = lr._td_loss(agent_out, target_out, acts, th.randn(T, B, 1, generator=g), th.zeros(T, B, 1),
                           th.randn(T + 1, B, S, generator=g))
    lr._optimise(loss, qv, sync=False)
    th.save(dict(init=init, after=[p.detach().clone() for p in lr.params],
                 n_mixer=sum(p.numel() for p in lr.mixer.parameters()), n_bucket=lr.grad_bucket.flat.numel()),
            os.path.join(out_dir, f"l{rank}.pt"))
    td.barrier()
    td.destroy_process_group()


@pytest.mark.timeout(180)
def test_qmix_learner_stays_in_lockstep_across_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_learner_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = th.load(tmp_path / "l0.pt"), th.load(tmp_path / "l1.pt")
    assert r0["n_bucket"] > r0["n_mixer"] > 0                              # the mixer's gradients ride in the same bucket
    for a, b in zip(r0["init"], r1["init"]):
        assert th.equal(a, b)
    moved = False
    for a0, a1, i0 in zip(r0["after"], r1["after"], r0["init"]):
        assert th.allclose(a0, a1, rtol=0, atol=1e-7)
        moved = moved or not th.equal(a0, i0)
    assert moved


def test_bucket_release_gather_equals_accumulating_into_the_views():
    from uav_bs_ctrl_b200 import dist
    th.manual_seed(0)
    net = th.nn.Sequential(th.nn.Linear(5, 7), th.nn.ReLU(), th.nn.Linear(7, 3), th.nn.Linear(3, 2))
    for p in net[3].parameters():
        p.requires_grad_(True)
    x = th.randn(11, 5)
    b = dist.FlatGradBucket(net.parameters())
    b.zero_()
    net[:3](x).pow(2).sum().backward()                    # the last layer gets no gradient
    b.rebind()
    ref = b.flat.clone()
    b.flat.fill_(7.0)
    b.release()
    assert all(p.grad is None for p in net.parameters())
    net[:3](x).pow(2).sum().backward()
    b.gather()
    assert th.equal(b.flat, ref)
    o = 0
    for p in net.parameters():                            # .grad are views of the flat buffer again
        assert p.grad.data_ptr() == b.flat.data_ptr() + 4 * o and th.equal(p.grad.flatten(), b.flat[o:o + p.numel()])
        o += p.numel()
    assert float(b.flat[-8:].abs().max()) == 0            # Linear(3, 2): 6 + 2 zeros
