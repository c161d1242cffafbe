"""Agent registries with the reference's keys (``algos/madrqn/agents/__init__.py:1-7``, ``algos/drqn/agents/__init__.py``)."""
from .gnn_agents import (GATv2Conv, GRUCell, GraphObservationEncoder, DenseObservationEncoder, TarMAC, BaseComm,
                         DiscreteComm, CommNet, EdgeConv, GnnAgent, DrqnGnnAgent)
from .rnn_agents import RnnAgent
from .dueling import DuelingLayer
from .mixers import QMixer

REGISTRY = {"rnn": RnnAgent, "gnn": GnnAgent}            # MADRQN
DRQN_REGISTRY = {"rnn": RnnAgent, "gnn": DrqnGnnAgent}   # DRQN
