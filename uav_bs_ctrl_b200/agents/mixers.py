"""QMIX monotonic mixing network (reference ``algos/madrqn/agents/mixers.py:6-54``).

``Q_tot = elu(q · |W1(s)| + b1(s)) · |w_f(s)| + V(s)`` with the weights produced by hyper-networks of the global state,
so ``∂Q_tot/∂q_i ≥ 0``.  Same constructor, parameter names (``hyper_w_1``, ``hyper_b_1``, ``hyper_w_final``, ``V.0``,
``V.2``) and ``forward(agent_qs (L, B, U), states (L, B, S)) -> (L, B, 1)`` as the reference, so its checkpoints load.
The three first-level hyper-networks read the same state rows: they run as ONE library GEMM over ``[W1; b1; w_f; V.0]``.
"""
import torch as th
import torch.nn as nn
import torch.nn.functional as F


class QMixer(nn.Module):
    def __init__(self, state_shape, n_agents, args):
        super().__init__()
        self.n_agents, self.state_dim, self.embed_dim = n_agents, int(state_shape), args.embed_dim
        E, S, U = self.embed_dim, self.state_dim, n_agents
        self.hyper_w_1 = nn.Linear(S, E * U)
        self.hyper_w_final = nn.Linear(S, E)
        self.hyper_b_1 = nn.Linear(S, E)
        self.V = nn.Sequential(nn.Linear(S, E), nn.ReLU(), nn.Linear(E, 1))

    def forward(self, agent_qs, states):
        L, B = agent_qs.shape[0], agent_qs.shape[1]
        E, U = self.embed_dim, self.n_agents
        s = states.reshape(-1, self.state_dim)
        heads = (self.hyper_w_1, self.hyper_b_1, self.hyper_w_final, self.V[0])
        hyper = th.addmm(th.cat([m.bias for m in heads]), s, th.cat([m.weight for m in heads]).t())
        w1, b1, wf, v0 = hyper.split((E * U, E, E, E), dim=1)
        q = agent_qs.reshape(-1, 1, U)
        hidden = F.elu(th.bmm(q, w1.abs().view(-1, U, E)) + b1.unsqueeze(1))
        v = self.V[2](th.relu(v0))
        q_tot = th.bmm(hidden, wf.abs().unsqueeze(2)).squeeze(2) + v
        return q_tot.view(L, B, 1)
