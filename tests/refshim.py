"""Runs the reference's OWN Python sources (``/root/reference``, unmodified, byte for byte) in the build container.

TEST INFRASTRUCTURE.  The reference imports four packages that are not installed (and not installable: no network):
``dgl``, ``gym``, ``matplotlib`` (and ``mpi4py`` in files this path never imports).  ``install()`` registers stand-ins
in ``sys.modules`` so that ``algos/madrqn/agents/gnn_agents.py``, ``algos/madrqn/learner.py``, ``buffer.py``,
``algos/common.py``, ``algos/madrqn/agents/mixers.py``, ``algos/madrqn/utils/env_wrappers.py``, the DRQN twins and
``envs/mubs_cov`` import and run AS THEY ARE:

* ``dgl``                  -> the DGL-free CPU graph container ``uav_bs_ctrl_b200.graph`` (``heterograph`` / ``batch`` /
                              ``merge`` / ``DGLGraph``): structure only;
* ``dgl.function``,
  ``dgl.nn.functional``    -> ``uav_bs_ctrl_b200.function`` (``u_dot_v``, ``u_mul_e``, ``sum``, UDF ``update_all``,
                              ``edge_softmax``): plain differentiable torch ops on CPU following SURVEY.md Appendix A.4;
* ``dgl.nn.pytorch``       -> ``GATv2Conv`` = the restated DGL 0.9.0 module of ``oracle/gnn_oracle.py`` (the one class of
                              this path whose source is NOT under /root/reference);
* ``gym.spaces``,
  ``matplotlib``           -> inert stand-ins (only constructed / imported, never used on this path).

Everything else that executes — ``GnnAgent``, ``GraphObservationEncoder``, ``TarMAC``, ``BaseComm``, ``CommNet``,
``DiscreteComm``, ``EdgeConv``, ``DuelingLayer``, ``QMixer``, ``MultiAgentQLearner.act / cache / update``, ``QLearner``,
``ReplayBuffer``, ``common.cat``, the env, the maps and the observation wrapper — is the reference's own code.  The
fixtures this produces (``tests/golden/make_ref_golden.py``) therefore pin the hot path to the reference's source
rather than to our restatement; what stays restated is listed above and in DESIGN.md §c.

``/root/reference`` does not exist on the GPU box: ``available()`` is False there and everything that needs the live
reference skips; the committed fixtures travel instead.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_installed = False


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "algos", "madrqn", "agents", "gnn_agents.py"))


def install():
    """Registers the stand-ins and puts /root/reference on ``sys.path`` (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"{REF} is not mounted: the live reference only exists in the build container")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    # ---- gym.spaces / matplotlib: inert --------------------------------------------------------------------------
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class _Space:
        def __init__(self, *a, **k):
            self.args, self.kwargs = a, k

    for name, cls in (("discrete", "Discrete"), ("box", "Box"), ("dict", "Dict")):
        m = types.ModuleType(f"gym.spaces.{name}")
        setattr(m, cls, type(cls, (_Space,), {}))
        sys.modules[f"gym.spaces.{name}"] = m
        setattr(spaces, name, m)
    utils = types.ModuleType("gym.spaces.utils")

    def flatten_space(space):
        # gym: a Dict of Boxes flattens to one Box whose length is the total element count
        n = sum(int(np.prod(b.kwargs["shape"])) for b in space.kwargs["spaces"].values())
        return types.SimpleNamespace(shape=(n,))

    utils.flatten_space = flatten_space
    # gym's flatten of a Dict space concatenates the entries in sorted-key order: agent, gt, ubs
    utils.flatten = lambda s, o: np.concatenate([np.asarray(o[k]).ravel() for k in ("agent", "gt", "ubs")])
    sys.modules["gym.spaces.utils"] = utils
    spaces.utils = utils
    gym.spaces = spaces
    sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    mpl = types.ModuleType("matplotlib")
    for sub in ("pyplot", "gridspec"):
        m = types.ModuleType(f"matplotlib.{sub}")
        sys.modules[f"matplotlib.{sub}"] = m
        setattr(mpl, sub, m)
    sys.modules["matplotlib"] = mpl
    # ---- dgl -------------------------------------------------------------------------------------------------------
    import torch.nn as nn
    from uav_bs_ctrl_b200 import graph as G, function as FN
    from oracle import gnn_oracle as O

    class GATv2Conv(O.GATv2Conv):
        """``dglnn.GATv2Conv`` stand-in: DGL accepts the relation view (``g['seen']``, madrqn) as well as a whole
        single-relation heterograph (drqn ``gnn_agents.py:27``)."""

        def forward(self, graph, feat, get_attention=False):
            if isinstance(graph, G.HeteroGraph):
                graph = graph[graph.canonical_etypes[0]]
            return super().forward(graph, feat, get_attention)

    dgl = types.ModuleType("dgl")
    dgl.heterograph, dgl.batch, dgl.merge, dgl.DGLGraph = G.heterograph, G.batch, G.merge, G.HeteroGraph
    dgl_fn = types.ModuleType("dgl.function")
    for k in FN.__all__:
        setattr(dgl_fn, k, getattr(FN, k))
    dgl_nn = types.ModuleType("dgl.nn")
    dgl_nn_pt = types.ModuleType("dgl.nn.pytorch")
    dgl_nn_pt.GATv2Conv = GATv2Conv
    dgl_nn_fn = types.ModuleType("dgl.nn.functional")
    dgl_nn_fn.edge_softmax = FN.edge_softmax
    dgl.function, dgl.nn = dgl_fn, dgl_nn
    dgl_nn.pytorch, dgl_nn.functional = dgl_nn_pt, dgl_nn_fn
    sys.modules.update({"dgl": dgl, "dgl.function": dgl_fn, "dgl.nn": dgl_nn, "dgl.nn.pytorch": dgl_nn_pt,
                        "dgl.nn.functional": dgl_nn_fn})
    assert isinstance(nn.Module, type)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _installed = True


def madrqn():
    """The reference's MADRQN modules (``gnn_agents``, ``learner``, ``buffer``, ``common``, ``mixers``, ``env_wrappers``)."""
    install()
    import importlib
    names = dict(agents="algos.madrqn.agents.gnn_agents", rnn="algos.madrqn.agents.rnn_agents",
                 learner="algos.madrqn.learner", buffer="algos.madrqn.buffer", common="algos.common",
                 mixers="algos.madrqn.agents.mixers", wrappers="algos.madrqn.utils.env_wrappers",
                 config="algos.madrqn.config")
    return types.SimpleNamespace(**{k: importlib.import_module(v) for k, v in names.items()})


def drqn():
    install()
    import importlib
    names = dict(agents="algos.drqn.agents.gnn_agents", learner="algos.drqn.learner", buffer="algos.drqn.buffer",
                 wrappers="algos.drqn.utils.env_wrappers", config="algos.drqn.config")
    return types.SimpleNamespace(**{k: importlib.import_module(v) for k, v in names.items()})


def mubs_env():
    install()
    from envs.mubs_cov.mubs_cov import MultiUbsCoverageEnv
    from envs.mubs_cov.maps import MAPS, DenseHotSpot
    if "8ubs80" not in MAPS:
        MAPS["8ubs80"] = DenseHotSpot(n_ubs=8, n_grps=16)      # BASELINE config from the reference's own map class
    return MultiUbsCoverageEnv, MAPS
