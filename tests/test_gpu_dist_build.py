"""Row-range (sharded) operator build == the same rows of the single-GPU build, bit for bit
(`pgsd_build_magnetic_rows_begin/_finish`).  One GPU plays every rank in turn; the
all-gather of the degrees is the concatenation of what the ranks' first phases produced."""
import pytest
import torch

from pytorch_geometric_signed_directed_b200 import distributed as pgd, plan as planmod, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sharded_rows(ei, ew, n, world, q, normalization, lmax, signed_mode, whole_list):
    bounds = pgd.node_bounds(n, world)
    degs = {}

    def run(collect):
        plans = []
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            e_r, w_r = (ei, ew) if whole_list else pgd.incident_edges(ei, ew, lo, hi)

            def ag(d, r=r):
                if collect:
                    degs[r] = d.clone()
                    return torch.zeros(n, device=DEV)
                return torch.cat([degs[b] for b in range(world)])
            plans.append(planmod.build_magnetic_rows(e_r, w_r, n, lo, hi, q, normalization, lmax, signed_mode, ag))
        return plans
    run(True)
    return bounds, run(False)


@pytest.mark.parametrize("world,normalization,signed_mode,weighted,whole_list", [
    (3, "sym", 0, False, False), (2, "sym", 0, True, False), (4, None, 0, True, True),
    (3, "sym", 1, True, False), (3, "sym", 2, True, False), (1, "sym", 0, True, False)])
def test_row_range_build_is_bit_identical(world, normalization, signed_mode, weighted, whole_list):
    g = torch.Generator().manual_seed(world * 10 + signed_mode)
    n, e = 5001, 60_000
    ei = torch.randint(0, n - 17, (2, e), generator=g)              # last nodes isolated (deg 0 -> inf -> 0)
    ei[:, :500] = ei[:, 500:1000]                                   # duplicates (summed in edge order)
    ei[:, 1000:1200] = ei[:, 1200:1400].flip(0)                     # reciprocal pairs
    ei[1, 1400:1500] = ei[0, 1400:1500]                             # self-loops (dropped)
    ew = None
    if weighted:
        ew = torch.rand(e, generator=g) + 0.25
        if signed_mode:
            ew = ew * torch.where(torch.rand(e, generator=g) < 0.3, -1.0, 1.0)
        ew = ew.to(DEV)
    ei = ei.to(DEV)
    lmax = 2.0 if normalization == "sym" else 7.5
    full = planmod.build_magnetic(ei, ew, n, 0.2, normalization, lmax, signed_mode)
    bounds, plans = _sharded_rows(ei, ew, n, world, 0.2, normalization, lmax, signed_mode, whole_list)
    assert sum(p.nnz for p in plans) == full.nnz
    for r, p in enumerate(plans):
        ref = pgd.split_rows(full, bounds[r], bounds[r + 1])
        assert p.n_dst == ref.n_dst and p.n_src == n and p.nnz == ref.nnz
        assert torch.equal(p.row_ptr, ref.row_ptr) and torch.equal(p.col, ref.col)
        assert torch.equal(p.val[0], ref.val[0]) and torch.equal(p.val[1], ref.val[1])
        assert p.diag_const == ref.diag_const
        assert torch.equal(p.meta["diag_real"], full.meta["diag_real"][bounds[r]:bounds[r + 1]])
        if ref.diag[0] is not None:
            assert torch.equal(p.diag[0], ref.diag[0])


def test_row_range_build_edge_cases():
    # empty graph, empty shard, and a proper range without the exchange callback
    ei = torch.zeros((2, 0), dtype=torch.long, device=DEV)
    p = planmod.build_magnetic_rows(ei, None, 10, 3, 7, 0.25, "sym", 2.0, 0, lambda d: torch.zeros(10, device=DEV))
    assert p.nnz == 0 and p.row_ptr.tolist() == [0] * 5
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]], device=DEV)
    p = planmod.build_magnetic_rows(ei, None, 3, 1, 1, 0.25, "sym", 2.0, 0, lambda d: torch.ones(3, device=DEV))
    assert p.n_dst == 0 and p.nnz == 0
    with pytest.raises(ValueError):
        planmod.build_magnetic_rows(ei, None, 3, 0, 2, 0.25, "sym", 2.0)
    whole = planmod.build_magnetic_rows(ei, None, 3, 0, 3, 0.25, "sym", 2.0)
    ref = planmod.build_magnetic(ei, None, 3, 0.25, "sym", 2.0)
    assert torch.equal(whole.col, ref.col) and torch.equal(whole.val[1], ref.val[1])


def test_row_range_build_at_scale_matches_replicated_build():
    """1M nodes / 20M edges, 8 row shards built one after the other."""
    n, world = 1_000_000, 8
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=DEV)
    full = planmod.build_magnetic(ei, None, n, 0.25, "sym", 2.0)
    bounds, plans = _sharded_rows(ei, None, n, world, 0.25, "sym", 2.0, 0, False)
    for r, p in enumerate(plans):
        ref = pgd.split_rows(full, bounds[r], bounds[r + 1])
        assert torch.equal(p.row_ptr, ref.row_ptr) and torch.equal(p.col, ref.col)
        assert torch.equal(p.val[0], ref.val[0]) and torch.equal(p.val[1], ref.val[1])
