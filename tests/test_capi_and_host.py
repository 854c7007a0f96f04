"""CPU-only checks: the C-ABI library loads and exports every symbol include/pgsd_b200.h
declares, argument validation fails loudly without touching the GPU, and the host-side layer
classes mirror the reference's constructor / state_dict / repr / error contracts."""
import ctypes as C
import os
import re

import pytest
import torch

import pytorch_geometric_signed_directed_b200 as pg
from pytorch_geometric_signed_directed_b200 import _lib, nn, synthetic
from pytorch_geometric_signed_directed_b200._lib import PgsdError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pgsd_b200.h")).read()
    declared = set(re.findall(r"PGSD_API\s+[\w\s\*]+?\b(pgsd_\w+)\s*\(", hdr))
    assert declared, "no prototypes parsed from the header"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pgsd_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.pgsd_abi_version() == 1


def test_ctypes_struct_sizes_match_the_compiled_header():
    lib = _lib.load()
    a, b = C.c_size_t(0), C.c_size_t(0)
    assert lib.pgsd_sizeof_args(C.byref(a), C.byref(b)) == 0
    assert a.value == C.sizeof(_lib.SpmmArgs) and b.value == C.sizeof(_lib.DenseArgs)
    assert lib.pgsd_sizeof_magnet_fused_args() == C.sizeof(_lib.MagnetFusedArgs)
    assert lib.pgsd_sizeof_push_args() == C.sizeof(_lib.PushArgs)
    assert lib.pgsd_sizeof_attn_bwd_args() == C.sizeof(_lib.AttnBwdArgs)


def test_exchange_and_attention_backward_entry_points_validate_arguments():
    """The entry points added in round 2 fail loudly on bad arguments, without touching the GPU."""
    lib = _lib.load()
    assert lib.pgsd_shard_push(None, None) == 1 and b"null" in lib.pgsd_last_error()
    p = _lib.PushArgs()
    p.world, p.rank, p.n_tensors, p.n_slices, p.row_bytes = 40, 0, 1, 1, 256
    assert lib.pgsd_shard_push(C.byref(p), None) == 1 and b"world" in lib.pgsd_last_error()
    p.world, p.row_bytes = 2, 100
    assert lib.pgsd_shard_push(C.byref(p), None) == 1 and b"row_bytes" in lib.pgsd_last_error()
    assert lib.pgsd_wait_flags(None, None, 1, 1, 1000, None, None) == 1
    assert lib.pgsd_edge_softmax_backward(None, None) == 1
    b = _lib.AttnBwdArgs()
    b.n_rows, b.n_types = 5, 3
    assert lib.pgsd_edge_softmax_backward(C.byref(b), None) == 1 and b"n_types" in lib.pgsd_last_error()
    assert lib.pgsd_signal_flag(None, 1, None) == 1
    assert lib.pgsd_last_spmm_kernel() is not None


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.pgsd_spmm_csr(None, None) == 1
    assert b"null" in lib.pgsd_last_error()
    a = _lib.SpmmArgs()
    a.n_ops = 3
    assert lib.pgsd_spmm_csr(C.byref(a), None) == 1
    assert b"n_ops" in lib.pgsd_last_error()
    d = _lib.DenseArgs()
    d.n_terms = 0
    assert lib.pgsd_dense_transform(C.byref(d), None) == 1
    nbytes = C.c_size_t(0)
    assert lib.pgsd_plan_workspace_bytes(1000, 5000, C.byref(nbytes)) == 0 and nbytes.value > 5000 * 16
    assert lib.pgsd_plan_workspace_bytes(10, 2 ** 31, C.byref(nbytes)) == 4   # PGSD_ERR_RANGE
    with pytest.raises(PgsdError):
        _lib.check(1, "x")


def test_layers_refuse_cpu_tensors_loudly():
    conv = nn.MagNetConv(4, 4, K=1, q=0.25, trainable_q=False)
    x = torch.randn(10, 4)
    ei = torch.randint(0, 10, (2, 30))
    with pytest.raises(PgsdError, match="no CPU path"):
        conv(x, x, ei)
    with pytest.raises(PgsdError):
        nn.DiGCNConv(4, 4)(x, ei, torch.rand(30))
    with pytest.raises(PgsdError):
        nn.SGCNConv(4, 2, True)(x, ei, ei)
    with pytest.raises(PgsdError):
        nn.Conv_Base()(x, ei)


def test_reference_constructor_contracts():
    c = nn.MagNetConv(3, 2, K=2, q=0.25, trainable_q=False)
    assert repr(c) == 'MagNetConv(3, 2, filter size=3, normalization=sym)'   # test/directed_test.py:73-74
    assert c.weight.shape == (3, 3, 2) and c.bias.shape == (2,) and c.cached is False
    assert c.cached_result is None and c.cached_num_edges is None
    assert set(c.state_dict()) == {"weight", "bias"}
    ct = nn.MagNetConv(3, 2, K=1, q=0.1, trainable_q=True, normalization=None, bias=False)
    assert set(ct.state_dict()) == {"weight", "q"} and ct.bias is None
    with pytest.raises(AssertionError):
        nn.MagNetConv(3, 2, K=0, q=0.25, trainable_q=False)
    with pytest.raises(AssertionError):
        nn.MagNetConv(3, 2, K=1, q=0.25, trainable_q=False, normalization='rw')
    m = nn.MSConv(3, 2, 2, 0.25, False, 'sym', True, True, False)   # bias, cached, absolute_degree order
    assert m.cached is True and m.absolute_degree is False and repr(m).startswith('MSConv(3, 2')
    d = nn.DiGCNConv(5, 4)
    assert d.cached is True and repr(d) == 'DiGCNConv(5, 4)' and set(d.state_dict()) == {"weight", "bias"}
    b = nn.DiGCN_InceptionBlock(5, 4)
    assert set(b.state_dict()) == {"ln.weight", "ln.bias", "conv1.weight", "conv1.bias",
                                   "conv2.weight", "conv2.bias"}
    s = nn.SGCNConv(8, 5, first_aggr=False)
    assert repr(s) == 'SGCNConv(8, 5, first_aggr=False)'   # test/signed_test.py:100,103
    assert s.lin_b.weight.shape == (5, 24) and set(s.state_dict()) == {
        "lin_b.weight", "lin_b.bias", "lin_u.weight", "lin_u.bias"}
    assert nn.SGCNConv(8, 5, first_aggr=True).lin_u.weight.shape == (5, 16)
    dm = nn.DIMPA(hop=2)
    assert dm._w_s.shape == (3, 1) and float(dm._w_t.sum()) == 3.0
    assert set(dm.state_dict()) == {"_w_s", "_w_t"}


def test_state_dict_round_trip_from_reference_shapes():
    ref_like = {"weight": torch.randn(2, 6, 4), "bias": torch.randn(4)}
    c = nn.MagNetConv(6, 4, K=1, q=0.25, trainable_q=False)
    c.load_state_dict(ref_like)
    assert torch.equal(c.weight, ref_like["weight"])


def test_synthetic_generators_follow_the_reference_distributions():
    ei, labels = synthetic.dsbm_edges(3000, k=3, num_edges=30000, eta=0.1, size_ratio=1.5, seed=3)
    assert ei.dtype == torch.int64 and ei.shape[0] == 2 and abs(ei.shape[1] - 30000) < 1500
    assert int((ei[0] == ei[1]).sum()) == 0
    assert torch.unique(ei[0] * 3000 + ei[1]).numel() == ei.shape[1]          # simple digraph
    sizes = torch.bincount(labels)
    assert sizes.sum() == 3000 and abs(sizes.max().item() / sizes.min().item() - 1.5) < 0.05
    # cyclic meta-graph: cluster c -> c+1 edges outnumber c+1 -> c by about (1-eta)/eta = 9
    fwd = ((labels[ei[1]] - labels[ei[0]]) % 3 == 1).sum().item()
    bwd = ((labels[ei[0]] - labels[ei[1]]) % 3 == 1).sum().item()
    assert 6.0 < fwd / bwd < 13.0
    pos, neg, lab = synthetic.ssbm_edges(2000, k=3, num_entries=40000, eta=0.1, seed=4)
    assert abs(pos.shape[1] + neg.shape[1] - 40000) < 2500
    inside = (lab[pos[0]] == lab[pos[1]]).float().mean().item()
    assert inside > 0.7 and (lab[neg[0]] != lab[neg[1]]).float().mean().item() > 0.85
    # both directions stored
    assert torch.equal(torch.sort(pos[0] * 2000 + pos[1]).values, torch.sort(pos[1] * 2000 + pos[0]).values)
    assert synthetic.meta_graph_cyclic(3, 0.1)[0, 1] == pytest.approx(0.9)


def test_argument_validation_of_the_newer_entry_points():
    """Entry points added after the first C-ABI batch reject bad arguments before touching the device."""
    lib = _lib.load()
    one = C.c_int64(0)
    buf = (C.c_float * 64)()
    ibuf = (C.c_int32 * 8)()
    # pgsd_gat_aggregate: feature width must be a multiple of 4, rows 16-byte aligned
    rc = lib.pgsd_gat_aggregate(ibuf, ibuf, buf, buf, 0.2, buf, 3, 3, 2, None, None, 0, 0.0, buf, 3, None)
    assert rc == 1 and b"multiple of 4" in lib.pgsd_last_error()
    rc = lib.pgsd_gat_aggregate(None, ibuf, buf, buf, 0.2, buf, 4, 4, 2, None, None, 0, 0.0, buf, 4, None)
    assert rc == 1 and b"null" in lib.pgsd_last_error()
    assert lib.pgsd_gat_aggregate(None, None, None, None, 0.2, None, 4, 4, 0, None, None, 0, 0.0, None, 4, None) == 0
    # pgsd_signed_triangle_counts: null list table
    rc = lib.pgsd_signed_triangle_counts(None, None, ibuf, ibuf, 5, 10, ibuf, None)
    assert rc == 1 and b"null" in lib.pgsd_last_error()
    assert lib.pgsd_signed_triangle_counts(None, None, None, None, 0, 10, None, None) == 0      # no edges: no-op
    # row-range builder: the range must lie inside [0, num_nodes]
    rc = lib.pgsd_build_magnetic_rows_begin(ibuf, ibuf, None, 1, 10, 7, 3, 0, ibuf, ibuf, buf, buf, buf,
                                            C.byref(one), buf, 1 << 20, None)
    assert rc != 0 and b"row range" in lib.pgsd_last_error()
    rc = lib.pgsd_build_magnetic_rows_finish(ibuf, ibuf, buf, buf, buf, 10, 4, 12, 0.25, 1, 2.0, buf, buf, buf, None)
    assert rc != 0 and b"row range" in lib.pgsd_last_error()
    assert lib.pgsd_build_magnetic_rows_finish(None, None, None, None, None, 10, 4, 4, 0.25, 1, 2.0, None, None, None,
                                               None) == 0                                         # empty shard


def test_new_model_wrappers_keep_the_reference_contracts():
    from pytorch_geometric_signed_directed_b200 import distributed as pgd
    m = nn.DiGCN_Inception_Block_node_classification(7, 16, 4, dropout=0.5)
    assert sorted(k for k, _ in m.named_parameters())[:3] == ["ib1.conv1.bias", "ib1.conv1.weight", "ib1.conv2.bias"]
    assert m.ib3.conv1.weight.shape == (16, 4) and m.ib1.ln.weight.shape == (16, 7)
    es = torch.tensor([[0, 1, 1], [1, 2, -1], [2, 0, 1]])
    with pytest.raises(NotImplementedError, match="init_emb"):
        nn.SGCN(3, es)
    with pytest.raises(NotImplementedError, match="init_emb"):
        nn.SDGNN(3, es)
    with pytest.raises(NotImplementedError, match="init_emb"):
        nn.SiGAT(3, es)
    with pytest.raises(PgsdError, match="no CPU path"):
        nn.SDGNN(3, es, in_dim=4, out_dim=4, init_emb=torch.randn(3, 4))       # motif mining needs the device
    sg = nn.SGCN(3, es, in_dim=4, out_dim=4, init_emb=torch.randn(3, 4))
    assert sg.pos_edge_index.tolist() == [[0, 2], [1, 0]] and sg.neg_edge_index.tolist() == [[1], [2]]
    assert [n_ for n_, _ in sg.named_parameters()][:2] == ["x", "conv1.lin_b.weight"]
    with pytest.raises(NotImplementedError):
        sg.loss()
    assert pgd.halo_fraction([torch.arange(2, dtype=torch.int32)] * 2, [0, 4, 8], 0) == 1.0


def test_plan_cache_modes_on_host(monkeypatch):
    """PlanCache: identity key + (on CUDA) content fingerprint; `off` rebuilds every call; capacity evicts the oldest;
    get_many serves several requests with one read-back.  (CPU tensors have no fingerprint: identity semantics.)"""
    from pytorch_geometric_signed_directed_b200.plan import PlanCache
    built = []

    def mk(tag):
        def build():
            built.append(tag)
            return tag
        return build

    a, b, c = torch.arange(10), torch.arange(6), torch.arange(3)
    cache = PlanCache(capacity=2)
    assert cache.get((a,), "k", mk("A")) == "A" and cache.get((a,), "k", mk("A'")) == "A"
    assert cache.get((a,), "other extra", mk("A2")) == "A2"
    a.add_(1)                                                    # in-place write bumps the version counter
    assert cache.get((a,), "k", mk("A3")) == "A3"
    assert cache.get_many([((b,), 1, mk("B")), ((c,), 1, mk("C"))]) == ["B", "C"]
    assert len(cache._items) == 2                                # capacity: the two oldest entries are gone
    monkeypatch.setenv("PGSD_PLAN_REUSE", "off")
    assert cache.get((b,), 1, mk("B again")) == "B again"
    monkeypatch.setenv("PGSD_PLAN_REUSE", "identity")
    assert cache.get((b,), 1, mk("never")) == "B"
    cache.clear()
    assert not cache._items
    assert built == ["A", "A2", "A3", "B", "C", "B again"]


def test_push_exchange_slice_defaults():
    from pytorch_geometric_signed_directed_b200 import distributed as pgd
    assert pgd.stage_fractions(world=2) == [0.0, 1.0]
    assert pgd.stage_fractions(world=4) == [0.0, 0.5, 1.0]
    assert pgd.stage_fractions(world=8) == [0.0, 0.25, 0.5, 0.75, 1.0]
    assert pgd.stage_fractions("1,1,2") == [0.0, 0.25, 0.5, 1.0]
    for n in (0, 1, 7, 1000):
        for cum in (pgd.stage_fractions(3), pgd.stage_fractions("0.22,0.2,0.18,0.16,0.14,0.1")):
            r = pgd.slice_rows(n, cum)
            assert r[0] == 0 and r[-1] == n and all(y >= x for x, y in zip(r, r[1:]))


def test_sparse_adjacency_inputs_are_read_as_transposed_adjacency():
    """SGCNConv accepts `Adj` (SGCNConv.py:94-95,131-134): a SparseTensor / torch sparse `adj_t[target, source]` is
    turned into the same [2, E] (source, target) edge list a COO call passes (reference test: test/signed_test.py:94-111)."""
    from pytorch_geometric_signed_directed_b200.plan import as_edge_index
    ei = torch.tensor([[0, 2, 2, 5, 1], [3, 1, 4, 0, 1]])
    assert as_edge_index(ei) is ei
    adj_t = torch.sparse_coo_tensor(torch.stack([ei[1], ei[0]]), torch.ones(5), (6, 6))
    assert torch.equal(as_edge_index(adj_t), ei)
    got = as_edge_index(adj_t.coalesce().to_sparse_csr())                 # CSR: rows sorted, same edge set
    key = lambda e: sorted(map(tuple, e.t().tolist()))
    assert key(got) == key(ei)

    class FakeSparseTensor:                                                # torch_sparse.SparseTensor duck type
        def coo(self):
            return ei[1], ei[0], None
    assert torch.equal(as_edge_index(FakeSparseTensor()), ei)
    with pytest.raises(NotImplementedError):
        as_edge_index("not an adjacency")
