"""CPU, world_size = 2, gloo: the host logic of the node-range sharded path (row split, column
blocks by owner, ring exchange schedule, block-wise accumulation incl. the Chebyshev epilogue).
The CUDA aggregation kernel is replaced by a torch stand-in passed as `aggregate_fn` (checker
only); results are compared with the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port
from pytorch_geometric_signed_directed_b200 import distributed as pgd
from pytorch_geometric_signed_directed_b200.plan import CSRPlan


def _plan_from_oracle(ei, n, q=0.25):
    """CSR-by-destination plan built on the CPU from the oracle's cached_result layout."""
    ei_r, ei_i, nr, ni = port.magnet_norm(ei, None, n, q, "sym", 2.0)
    nnz = ei_i.size(1) - n                       # sorted (row, col) block; loops follow
    src, dst = ei_i[0, :nnz], ei_i[1, :nnz]      # source_to_target: gather src, reduce at dst
    order = torch.sort(dst * n + src, stable=True).indices
    rp = torch.zeros(n + 1, dtype=torch.int32)
    rp[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0).int()
    return CSRPlan(n, n, nnz, ei.size(1), rp, src[order].int(), [nr[:nnz][order], ni[:nnz][order]],
                   [None, None], [0.25, 0.0])      # non-zero real diagonal (lambda_max != 2)


def _torch_aggregate(block, xs, op_ids, alpha, beta, zs, out):
    counts = (block.row_ptr[1:] - block.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(block.n_dst), counts)
    ys = []
    for k, op in enumerate(op_ids):
        agg = torch.zeros(block.n_dst, xs[k].size(1))
        if block.nnz:
            agg.index_add_(0, rows, block.val[op].view(-1, 1) * xs[k][block.col.long()])
        if block.diag_const[op] != 0.0:
            off = block.meta.get("diag_row_offset", 0)
            agg += block.diag_const[op] * xs[k][off:off + block.n_dst]
        y = alpha * agg
        if zs is not None and zs[k] is not None:
            y = y + beta * zs[k]
        ys.append(y)
    return ys


def _worker(rank, world, port_no, n, e, f):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ei = torch.randint(0, n, (2, e), generator=g)
        xr = torch.rand(n, f, generator=g) * 2 - 1
        xi = torch.rand(n, f, generator=g) * 2 - 1
        full = _plan_from_oracle(ei, n)
        bounds = pgd.node_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        local = pgd.split_rows(full, lo, hi)
        assert local.n_dst == hi - lo
        # the single-launch "gather" mode must agree with the pipelined "ring" mode
        agg_g = pgd.ShardedAggregator(local, bounds, rank, world, aggregate_fn=_torch_aggregate, mode="gather")
        tg = agg_g([xr[lo:hi], xi[lo:hi]])
        agg = pgd.ShardedAggregator(local, bounds, rank, world, aggregate_fn=_torch_aggregate, mode="ring")
        assert sum(b.nnz for b in agg.blocks) == local.nnz
        for b, blk in enumerate(agg.blocks):
            assert blk.n_src == bounds[b + 1] - bounds[b]
            if blk.nnz:
                assert 0 <= int(blk.col.min()) and int(blk.col.max()) < blk.n_src
        # reference: unsharded aggregation restricted to this rank's rows
        ref = _torch_aggregate(full, [xr, xi], (0, 1), 1.0, 0.0, None, None)
        t1 = agg([xr[lo:hi], xi[lo:hi]])
        for k in range(2):
            assert torch.allclose(t1[k], ref[k][lo:hi], atol=1e-5), f"rank {rank} op {k}"
            assert torch.allclose(tg[k], ref[k][lo:hi], atol=1e-5), f"rank {rank} gather-mode op {k}"
        # one tensor for both operators: the shard travels once (half-width buffers), same result
        ref_s = _torch_aggregate(full, [xr, xr], (0, 1), 1.0, 0.0, None, None)
        xs_loc = xr[lo:hi]
        ts = agg([xs_loc, xs_loc])
        assert agg._recv[0][-1] == f                       # exchange width F, not 2F
        for k in range(2):
            assert torch.allclose(ts[k], ref_s[k][lo:hi], atol=1e-5), f"rank {rank} shared op {k}"
        # Chebyshev step: T2 = 2 L T1 - T0 with a second exchange
        ref2 = _torch_aggregate(full, ref, (0, 1), 2.0, -1.0, [xr, xi], None)
        t2 = agg(t1, alpha=2.0, beta=-1.0, zs=[xr[lo:hi], xi[lo:hi]])
        for k in range(2):
            assert torch.allclose(t2[k], ref2[k][lo:hi], atol=1e-5), f"rank {rank} cheb op {k}"
    finally:
        dist.destroy_process_group()


def _torch_gather(x, index, out=None):
    res = x[index.long()]
    if out is None:
        return res
    out.copy_(res)
    return out


def _halo_worker(rank, world, port_no, n, e, f, band, long_range):
    """Halo path: request lists through all-to-all, per-step pack -> all_to_all_single -> two blocks."""
    from pytorch_geometric_signed_directed_b200 import synthetic
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei = synthetic.locality_edges(n, e, band, long_range, seed=3)
        g = torch.Generator().manual_seed(1)
        xr = torch.rand(n, f, generator=g) * 2 - 1
        xi = torch.rand(n, f, generator=g) * 2 - 1
        full = _plan_from_oracle(ei, n)
        bounds = pgd.node_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        local = pgd.split_rows(full, lo, hi)
        own, halo, need = pgd.split_local_and_halo(local, bounds, rank)
        assert own.nnz + halo.nnz == local.nnz and need[rank].numel() == 0
        # the compact columns of the halo block index the concatenation of the need lists
        cat = torch.cat([need[b].long() + bounds[b] for b in range(world)])
        counts = (local.row_ptr[1:] - local.row_ptr[:-1]).long()
        colg = local.col.long()
        remote = (colg < lo) | (colg >= hi)
        assert torch.equal(cat[halo.col.long()], colg[remote])                  # same entries, same order
        assert torch.equal(cat, torch.unique(colg[remote]))                     # sorted, distinct
        agg = pgd.ShardedAggregator(local, bounds, rank, world, aggregate_fn=_torch_aggregate, mode="auto",
                                    gather_fn=_torch_gather)
        expect_halo = long_range == 0.0
        assert agg.mode == ("halo" if expect_halo else "ring"), (agg.mode, agg.halo_fraction)
        if expect_halo:
            assert agg.halo.n_recv == cat.numel() and agg.halo_fraction < 0.5
            # what the peers were told to serve is exactly what this rank needs (round trip)
            got = torch.empty(agg.halo.n_recv, dtype=torch.int32)
            pgd._all_to_all(got, agg.halo.serve + lo, agg.halo.need_splits, agg.halo.serve_splits)
            assert torch.equal(got.long(), cat)
        ref = _torch_aggregate(full, [xr, xi], (0, 1), 1.0, 0.0, None, None)
        t1 = agg([xr[lo:hi], xi[lo:hi]])
        for k in range(2):
            assert torch.allclose(t1[k], ref[k][lo:hi], atol=1e-5), f"rank {rank} op {k}"
        # one tensor for both operators: halo rows are packed and exchanged once
        ref_s = _torch_aggregate(full, [xr, xr], (0, 1), 1.0, 0.0, None, None)
        xs_loc = xr[lo:hi]
        ts = agg([xs_loc, xs_loc])
        for k in range(2):
            assert torch.allclose(ts[k], ref_s[k][lo:hi], atol=1e-5), f"rank {rank} shared op {k}"
        ref2 = _torch_aggregate(full, ref, (0, 1), 2.0, -1.0, [xr, xi], None)
        t2 = agg(t1, alpha=2.0, beta=-1.0, zs=[xr[lo:hi], xi[lo:hi]])
        t2b = agg(t1, alpha=2.0, beta=-1.0, zs=[xr[lo:hi], xi[lo:hi]])          # buffers are reused
        for k in range(2):
            assert torch.allclose(t2[k], ref2[k][lo:hi], atol=1e-5), f"rank {rank} cheb op {k}"
            assert torch.equal(t2[k], t2b[k])
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_aggregation_matches_unsharded(world):
    mp.spawn(_worker, args=(world, _free_port(), 301, 4000, 8), nprocs=world, join=True)


def test_node_bounds_and_ring_schedule():
    assert pgd.node_bounds(10, 3) == [0, 3, 6, 10]
    assert pgd.node_bounds(8_000_000, 8)[-1] == 8_000_000
    ring = pgd.RingExchange(rank=1, world=4)
    assert [ring.source_of_round(s) for s in (1, 2, 3)] == [0, 3, 2]


@pytest.mark.parametrize("world,long_range", [(2, 0.0), (3, 0.0), (4, 0.0), (2, 0.2)])
def test_halo_exchange_on_a_naturally_sharding_graph(world, long_range):
    """Locality-ordered graph: thin halo -> `auto` picks the all-to-all halo path; with 20 % long-range
    edges the halo is most of the matrix and `auto` must fall back to the ring all-gather."""
    mp.spawn(_halo_worker, args=(world, _free_port(), 600, 9000, 8, 25, long_range), nprocs=world, join=True)


def test_halo_fraction():
    need = [torch.arange(3, dtype=torch.int32), torch.empty(0, dtype=torch.int32), torch.arange(5, dtype=torch.int32)]
    assert pgd.halo_fraction(need, [0, 10, 20, 30], 1) == 8 / 20


def _route_worker(rank, world, port_no, n, e):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        ei = torch.randint(0, n, (2, e), generator=g)
        ew = torch.rand(e, generator=g) - 0.5
        bounds = pgd.node_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        cuts = [(r * e) // world for r in range(world + 1)]            # contiguous slices of the edge list
        sl = slice(cuts[rank], cuts[rank + 1])
        want_ei, want_ew = pgd.incident_edges(ei, ew, lo, hi)
        got_ei, got_ew = pgd.route_edges(ei[:, sl].contiguous(), ew[sl].contiguous(), bounds, rank, world)
        assert torch.equal(got_ei, want_ei) and torch.equal(got_ew, want_ew)     # same edges, ORIGINAL order
        got_ei2, none = pgd.route_edges(ei[:, sl].contiguous(), None, bounds, rank, world)
        assert none is None and torch.equal(got_ei2, want_ei)
        # uneven shards: all-gather of per-shard vectors
        mine = torch.arange(lo, hi, dtype=torch.float32) * 0.5
        assert torch.equal(pgd.allgather_rows(mine, bounds), torch.arange(n, dtype=torch.float32) * 0.5)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_edge_routing_for_the_distributed_build(world):
    mp.spawn(_route_worker, args=(world, _free_port(), 101, 3000), nprocs=world, join=True)


def test_locality_graph_generator_properties():
    from pytorch_geometric_signed_directed_b200 import synthetic
    n, band = 5000, 40
    ei = synthetic.locality_edges(n, 60_000, band, 0.0, seed=2)
    assert ei.dtype == torch.long and ei.shape[0] == 2
    assert int(ei.min()) >= 0 and int(ei.max()) < n
    assert bool((ei[0] != ei[1]).all())                                   # no self-loops
    assert bool(((ei[0] - ei[1]).abs() <= band).all())                    # banded
    key = ei[0] * n + ei[1]
    assert key.unique().numel() == key.numel()                            # no duplicates
    far = synthetic.locality_edges(n, 60_000, band, 0.3, seed=2)
    assert float(((far[0] - far[1]).abs() > band).float().mean()) > 0.2   # long-range share lands anywhere
    # a 1-D node-range split of the banded graph needs only a thin band of each neighbour shard
    bounds = pgd.node_bounds(n, 4)
    lo, hi = bounds[1], bounds[2]
    rows = (ei[1] >= lo) & (ei[1] < hi)                                   # entries aggregated by rank 1 (dst rows)
    src = ei[0][rows]
    remote = src[(src < lo) | (src >= hi)]
    assert remote.unique().numel() <= 2 * band


def test_stage_split_covers_the_shard_and_matches_the_slices():
    """Push exchange, host side: the stage blocks partition a row shard's entries (block 0 = own columns,
    block s = columns in slice s-1 of any peer's shard, GLOBAL numbering), slices use one formula on sender
    and receiver, and block-wise accumulation reproduces the unsharded aggregation."""
    n, e, f, world = 1003, 9000, 8, 3
    g = torch.Generator().manual_seed(5)
    ei = torch.randint(0, n, (2, e), generator=g)
    xr = torch.rand(n, f, generator=g) * 2 - 1
    xi = torch.rand(n, f, generator=g) * 2 - 1
    full = _plan_from_oracle(ei, n)
    bounds = pgd.node_bounds(n, world)
    ref = _torch_aggregate(full, [xr, xi], (0, 1), 1.0, 0.0, None, None)
    for spec in (None, 1, 3, 5):
        cum = pgd.stage_fractions(spec)
        assert cum[0] == 0.0 and cum[-1] == 1.0 and all(b >= a for a, b in zip(cum, cum[1:]))
        for rank in range(world):
            lo, hi = bounds[rank], bounds[rank + 1]
            local = pgd.split_rows(full, lo, hi)
            blocks = pgd.split_columns_by_stage(local, bounds, rank, cum)
            assert len(blocks) == len(cum) and sum(b.nnz for b in blocks) == local.nnz
            assert blocks[0].n_src == hi - lo and all(b.n_src == n for b in blocks[1:])
            cuts = {b: pgd.slice_rows(bounds[b + 1] - bounds[b], cum) for b in range(world)}
            for s, blk in enumerate(blocks[1:]):
                c = blk.col.long()
                own = (c >= lo) & (c < hi)
                assert not bool(own.any())
                for b in range(world):
                    inb = c[(c >= bounds[b]) & (c < bounds[b + 1])] - bounds[b]
                    if inb.numel():
                        assert int(inb.min()) >= cuts[b][s] and int(inb.max()) < cuts[b][s + 1]
            # own block reads the local rows, the stage blocks the global planes; (+=) over the blocks
            y = _torch_aggregate(blocks[0], [xr[lo:hi], xi[lo:hi]], (0, 1), 1.0, 0.0, None, None)
            for blk in blocks[1:]:
                y = _torch_aggregate(blk, [xr, xi], (0, 1), 1.0, 1.0, y, None)
            for k in range(2):
                assert torch.allclose(y[k], ref[k][lo:hi], atol=1e-5)
