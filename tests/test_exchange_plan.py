"""Multi-GPU gradient exchange plan (cb200_engine_exchange_plan), checked on the CPU with
planning-only engines: for every rank of world sizes 2, 3 and 8 the chunks must tile the rank's
residual blocks, their exclusive gradient ranges must tile the rank's exclusive range, and no
residual block outside a chunk may touch the chunk's range - the property that lets the
evaluation kernel copy a chunk's gradient entries to the other ranks as soon as the chunk is
done.  The protocol itself (copy exclusive ranges, sum the shared entries in rank order) is then
replayed with the oracle's per-rank gradients and must reproduce the unsharded gradient."""
import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import binding as B, problems as P


def _owner_columns(cp):
    """delta offset and tangent size of every active parameter block, by program order."""
    tab = cp.pb_table()  # [size, tangent, state_offset, delta_offset] per program block
    return tab


def _rank_gradient(spec, lo, hi, x, fmt=0):
    sub = P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[lo:hi],
        rb_pb=spec.rb_pb[2 * lo:2 * hi], fdata=spec.fdata[2 * lo:2 * hi],
        rb_loss_kind=spec.rb_loss_kind[lo:hi], rb_loss_a=spec.rb_loss_a[lo:hi],
        rb_loss_b=spec.rb_loss_b[lo:hi], num_eliminate_blocks=spec.num_eliminate_blocks)
    sop = O.OracleProblem(sub, jacobian_format=fmt, reduce=False)
    ok, c, r, g, j = sop.evaluate(x)
    assert ok
    return c, g


@pytest.mark.parametrize("world", [2, 3, 8])
def test_chunks_tile_the_rank_and_own_their_gradient_range(world):
    spec = P.bal_problem(11, 1500, 9000, seed=21)
    npts = 1500
    point_of_rb = spec.rb_pb.reshape(-1, 2)[:, 1]
    op = O.OracleProblem(spec, jacobian_format=0)
    x = op.initial_state()
    ok, c_full, r_full, g_full, j_full = op.evaluate(x)
    assert ok
    plans, infos = [], []
    for rank in range(world):
        cp = B.CudaProblem(spec, jacobian_format=0, device=-1, rank=rank, world_size=world)
        plan = cp.exchange_plan()
        assert plan is not None, "a Schur-ordered bundle adjustment problem allows the peer exchange"
        infos.append(cp.shard_info())
        plans.append(plan)
        cp.close()
    shared = plans[0]["shared_count"]
    assert all(p["shared_count"] == shared for p in plans)
    # shared = 9 per camera + the points at rank boundaries (3 each, at most world - 1) + 2
    assert 9 * 11 + 2 <= shared <= 9 * 11 + 2 + 3 * (world - 1)

    total_exclusive = 0
    for rank, (plan, info) in enumerate(zip(plans, infos)):
        lo, hi = info["rb_begin"], info["rb_end"]
        ch = plan["chunks"]
        eb, el = plan["exclusive"]
        total_exclusive += el
        assert ch[0, 0] == 0 and ch[-1, 1] == hi - lo
        assert np.array_equal(ch[1:, 0], ch[:-1, 1])          # blocks: contiguous tiling
        assert ch[0, 2] == eb and ch[-1, 3] == eb + el
        assert np.array_equal(ch[1:, 2], ch[:-1, 3])          # gradient ranges: contiguous tiling
        assert np.all(ch[:, 1] > ch[:, 0]) and np.all(ch[:, 3] >= ch[:, 2])
        # points are parameter blocks 0..npts-1 with delta offset 3 * index (cameras after them)
        for c0, c1, d0, d1 in ch:
            pts = point_of_rb[lo + c0:lo + c1]
            cols = 3 * pts
            inside = (cols >= d0) & (cols < d1)
            exclusive = (cols >= eb) & (cols < eb + el)
            assert np.array_equal(inside, exclusive)           # every exclusive point it touches
            others = np.concatenate([point_of_rb[:lo + c0], point_of_rb[lo + c1:]])
            oc = 3 * others
            assert not np.any((oc >= d0) & (oc < d1))          # ... is touched by it alone
    assert total_exclusive + shared - 2 == g_full.size

    # replay the exchange: exclusive ranges are copied, shared entries summed in rank order
    parts = [_rank_gradient(spec, i["rb_begin"], i["rb_end"], x) for i in infos]
    g = np.zeros_like(g_full)
    is_shared = np.ones(g.size, dtype=bool)
    for (c, gr), plan in zip(parts, plans):
        eb, el = plan["exclusive"]
        g[eb:eb + el] = gr[eb:eb + el]
        is_shared[eb:eb + el] = False
    for c, gr in parts:
        g[is_shared] += gr[is_shared]
    assert np.max(np.abs(g - g_full)) <= 1e-12 * np.max(np.abs(g_full))
    assert abs(sum(c for c, _ in parts) - c_full) <= 1e-12 * abs(c_full)


def test_unordered_problems_fall_back_to_the_all_reduce():
    # two residual-block types interleaved: no chunked plan, the ranks use NCCL's all-reduce
    spec = P.evaluator_cuda_test_problem()
    cp = B.CudaProblem(spec, jacobian_format=0, device=-1, rank=0, world_size=2)
    assert cp.exchange_plan() is None
    cp.close()


def test_long_chunks_are_cut_near_the_target_size():
    spec = P.bal_problem(40, 6000, 40000, seed=4)
    cp = B.CudaProblem(spec, jacobian_format=0, device=-1, rank=1, world_size=2)
    ch = cp.exchange_plan()["chunks"]
    sizes = ch[:, 1] - ch[:, 0]
    # target: ~12 chunks per warp of a full persistent grid, 128..4096 blocks; a chunk ends at
    # the first valid cut that makes its length a multiple of 32 (no partial tile) once half
    # the target is reached, and never grows beyond three times the target when the points have a
    # handful of observations each (a valid cut every few blocks)
    local = 40000 // 2
    target = min(4096, max(128, 2 * (local // (148 * 12 * 12) - 104) // 32 * 32))
    assert target == 128
    assert sizes.max() <= 3 * target and np.all(sizes[:-1] >= target // 2)
    assert np.mean(sizes[:-1] % 32 == 0) >= 0.75
    cp.close()
