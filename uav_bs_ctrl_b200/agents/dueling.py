"""Dueling head (reference ``algos/madrqn/agents/dueling.py:4-16``); runs after the hot path, plain torch."""
import torch.nn as nn


class DuelingLayer(nn.Module):
    def __init__(self, in_feats, n_actions):
        super().__init__()
        self.adv_head = nn.Linear(in_feats, n_actions)
        self.v_head = nn.Linear(in_feats, 1)

    def forward(self, x):
        advs = self.adv_head(x)
        return self.v_head(x) + (advs - advs.mean(-1, keepdim=True))
