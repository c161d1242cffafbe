"""Host logic: DGL-surface graph container, batch / merge semantics, CSR composition, vectorised builder."""
import torch as th
import pytest

from uav_bs_ctrl_b200 import graph as G, function as fn
from uav_bs_ctrl_b200.builder import build_obs_graph_batch, build_drqn_graph_batch
from uav_bs_ctrl_b200.synth import synth_dense_obs
from helpers import batched_graph_ref, env_graph_ref


def test_heterograph_basic():
    g = G.heterograph({('gt', 'seen', 'agent'): (th.arange(3), th.zeros(3, dtype=th.long)),
                       ('ubs', 'near', 'agent'): ([], []), ('agent', 'talk', 'agent'): ([], [])},
                      num_nodes_dict={'gt': 3, 'ubs': 0, 'agent': 1})
    assert g.ntypes == ['agent', 'gt', 'ubs']
    assert g.num_nodes('gt') == 3 and g.num_nodes('agent') == 1 and g.number_of_edges() == 3
    assert isinstance(g, G.DGLGraph)
    g.ndata['feat'] = {'gt': th.ones(3, 4), 'agent': th.zeros(1, 2), 'ubs': th.zeros(0, 2)}
    assert set(g.ndata['feat'].keys()) == {'gt', 'agent', 'ubs'}
    assert g.nodes['agent'].data['feat'].shape == (1, 2)
    c = g['seen'].csr()
    assert c.is_star and c.indptr.tolist() == [0, 3] and c.eid is None
    assert g['seen'].num_dst_nodes() == 1 and g['seen'].num_src_nodes() == 3
    with pytest.raises(KeyError):
        g['nope']


def test_local_scope_drops_writes():
    g = G.heterograph({('agent', 'talk', 'agent'): ([0, 1], [1, 0])}, num_nodes_dict={'agent': 2})
    rel = g['talk']
    with rel.local_scope():
        rel.ndata['x'] = th.ones(2, 3)
        rel.edata['a'] = th.ones(2, 1)
        assert 'x' in rel.srcdata
    assert 'x' not in rel.ndata and 'a' not in rel.edata


def test_debug_map_layout():
    """Reference Debug map (envs/mubs_cov/maps.py:38-50): 3 UBS, 4 GT, r_sns=300 ⇒ visibility ⇒ degrees."""
    pos_ubs = 100 * th.tensor([[3., 3], [8, 2], [8, 9]])
    pos_gts = 100 * th.tensor([[3., 4], [4, 2], [3, 1], [6, 9]])
    d = (pos_ubs[:, None] - pos_gts[None]).norm(dim=-1)
    vis = d <= 300.
    assert vis.sum(1).tolist() == [3, 0, 1]          # UBS0 sees GT0..2, UBS1 none, UBS2 sees GT3
    U, Gn = 3, 4
    gt = th.zeros(1, U, Gn, 5)
    gt[0, ..., 0] = vis.float()
    gt[0, ..., 1:3] = (pos_gts[None] - pos_ubs[:, None]) / 300.
    ubs = th.zeros(1, U, U - 1, 3)
    ubs[..., 0] = 1
    g = build_obs_graph_batch(pos_ubs.view(1, U, 2) / 1000., gt, ubs, th.ones(1, U, U, dtype=th.bool))
    assert g['seen'].csr().indptr.tolist() == [0, 3, 3, 4]
    assert g['near'].csr().indptr.tolist() == [0, 2, 4, 6]
    assert g.num_nodes('gt') == 4 and g['talk'].num_edges() == 9
    assert th.allclose(g.nodes['gt'].data['feat'][3, :2], (pos_gts[3] - pos_ubs[2]) / 300.)


@pytest.mark.parametrize("profile,comm_p", [("full", 1.0), ("realistic", 0.5), ("random", 0.3)])
def test_vectorised_builder_matches_reference_style(profile, comm_p):
    B, U, Gn = 5, 4, 6
    a, gt, ubs, adj = synth_dense_obs(B, U, Gn, profile, seed=3, comm_p=comm_p, near_p=0.7)
    ref = batched_graph_ref(a, gt, ubs, adj)
    vec = build_obs_graph_batch(a, gt, ubs, adj)
    assert ref.batch_size == vec.batch_size == B
    for nt in ('agent', 'gt', 'ubs'):
        assert ref.num_nodes(nt) == vec.num_nodes(nt)
        assert th.equal(ref.nodes[nt].data['feat'], vec.nodes[nt].data['feat'])
        assert ref.batch_num_nodes(nt).tolist() == vec.batch_num_nodes(nt).tolist()
    for et in ('seen', 'near', 'talk'):
        (ru, rv), (vu, vv) = ref.edges(et), vec.edges(et)
        assert th.equal(ru, vu) and th.equal(rv, vv), et
        rc, vc = ref[et].csr(), vec[et].csr()
        assert th.equal(rc.indptr, vc.indptr) and rc.is_star == vc.is_star
        assert th.equal(rc.src_of_slot(), vc.src_of_slot())
        if rc.eid is not None or vc.eid is not None:
            E = rc.n_edges
            re = rc.eid if rc.eid is not None else th.arange(E)
            ve = vc.eid if vc.eid is not None else th.arange(E)
            assert th.equal(re, ve)
    assert ref['seen'].csr().is_star and ref['near'].csr().is_star
    assert ref.uniform_block('agent') == U


def test_csr_matches_fresh_sort():
    a, gt, ubs, adj = synth_dense_obs(3, 5, 7, "realistic", seed=9, comm_p=0.4)
    g = batched_graph_ref(a, gt, ubs, adj)
    for et in ('seen', 'near', 'talk'):
        rel = g[et]
        u, v = rel.edges()
        fresh = G.RelCSR.from_edges(u, v, rel.num_src_nodes(), rel.num_dst_nodes())
        c = rel.csr()
        assert th.equal(c.indptr, fresh.indptr)
        assert th.equal(c.src_of_slot(), fresh.src_of_slot())
        assert th.equal(c.dst_of_slot(), v if c.eid is None else v[c.eid])


def test_merge_semantics():
    a, gt, ubs, adj = synth_dense_obs(1, 3, 4, "full", seed=1)
    g = env_graph_ref(a[0], gt[0], ubs[0], adj[0])
    assert g.batch_size == 1 and g.num_nodes('agent') == 3 and g['talk'].num_edges() == 9
    assert g.nodes['gt'].data['feat'].shape == (12, 4)
    u, v = g.edges('talk')
    assert u.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2] and v.tolist() == [0, 1, 2] * 3     # src-major
    c = g['talk'].csr()
    assert not c.is_star and c.src_of_slot().tolist() == [0, 1, 2] * 3 and c.eid.tolist() == [0, 3, 6, 1, 4, 7, 2, 5, 8]


def test_to_and_empty():
    a, gt, ubs, adj = synth_dense_obs(2, 2, 3, "random", seed=5)
    gt[..., 0] = 0                                            # nobody sees anything
    g = batched_graph_ref(a, gt, ubs, adj)
    assert g.num_nodes('gt') == 0 and g['seen'].csr().indptr.tolist() == [0] * 5
    g2 = g.to('cpu')
    assert g2 is g
    v = build_obs_graph_batch(a, gt, ubs, None)
    assert v['talk'].num_edges() == 0 and v['talk'].csr().n_edges == 0


def test_message_passing_api_matches_manual():
    th.manual_seed(0)
    g = G.heterograph({('agent', 'talk', 'agent'): ([0, 1, 2, 2], [1, 1, 0, 2])}, num_nodes_dict={'agent': 4})
    rel = g['talk']
    s, q, v = th.randn(4, 3), th.randn(4, 3), th.randn(4, 5)
    with rel.local_scope():
        rel.srcdata.update(dict(s=s, v=v))
        rel.dstdata.update(dict(q=q))
        rel.apply_edges(fn.u_dot_v('s', 'q', 'e'))
        e = rel.edata.pop('e')
        assert th.allclose(e[:, 0], th.stack([s[0] @ q[1], s[1] @ q[1], s[2] @ q[0], s[2] @ q[2]]))
        rel.edata['a'] = fn.edge_softmax(rel, e)
        a = rel.edata['a']
        assert th.allclose(a[0] + a[1], th.ones(1)) and th.allclose(a[2], th.ones(1))
        rel.update_all(fn.u_mul_e('v', 'a', 'm'), fn.sum('m', 'c'))
        c = rel.dstdata['c']
        assert th.allclose(c[1], a[0] * v[0] + a[1] * v[1]) and th.allclose(c[3], th.zeros(5))
        # UDF path (mailbox, degree bucketing) — mean / max reducers of the other comm protocols
        rel.update_all(lambda ed: {'m': ed.src['v']}, lambda nd: {'mx': nd.mailbox['m'].max(1)[0],
                                                                  'mean': nd.mailbox['m'].mean(1)})
        assert th.allclose(rel.dstdata['mx'][1], th.maximum(v[0], v[1]))
        assert th.allclose(rel.dstdata['mean'][1], (v[0] + v[1]) / 2)
        assert th.allclose(rel.dstdata['mean'][3], th.zeros(5))


def test_drqn_builder():
    g = build_drqn_graph_batch(th.rand(3, 2), th.rand(3, 10, 4))
    rel = g[('gt', 'seen-by', 'agent')]
    assert rel.csr().is_star and rel.csr().indptr.tolist() == [0, 10, 20, 30]
    assert g.nodes['gt'].data['feat'].shape == (30, 4)
