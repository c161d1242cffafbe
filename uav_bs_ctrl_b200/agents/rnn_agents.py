"""MLP + GRU baseline agent (reference ``algos/madrqn/agents/rnn_agents.py:6-35``, ``algos/drqn/agents/rnn_agents.py:5-27``).
Not on the graph hot path; kept so that ``REGISTRY['rnn']`` resolves for ``o='mlp', c=None``."""
import torch as th
import torch.nn as nn

from .dueling import DuelingLayer
from .gnn_agents import GRUCell


class RnnAgent(nn.Module):
    def __init__(self, obs_shape, n_actions, args):
        super().__init__()
        self._n_layers, self._hidden_size = args.n_layers, args.hidden_size
        layers = [nn.Linear(obs_shape, self._hidden_size), nn.ReLU()]
        for _ in range(self._n_layers - 1):
            layers += [nn.Linear(self._hidden_size, self._hidden_size), nn.ReLU()]
        self.enc = nn.Sequential(*layers)
        self.rnn = GRUCell(self._hidden_size, self._hidden_size)
        if getattr(args, "dueling", False):
            self.f_out = DuelingLayer(self._hidden_size, n_actions)
        else:
            self.f_out = nn.Linear(self._hidden_size, n_actions)

    def init_hidden(self):
        return th.zeros(1, self._hidden_size)

    def forward(self, obs, h):
        h = self.rnn(self.enc(obs), h)
        return self.f_out(h), h
