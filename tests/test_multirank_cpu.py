"""N > 1 path on CPU: two gloo ranks plan their shards with the product's host code
(planning-only engine, device = -1), check that the shards tile the problem, evaluate
their residual-block range with the oracle and combine cost/gradient with one
all-reduce — the same data flow the GPU ranks use with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fmt, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ceres_b200  # noqa: F401
    from ceres_b200 import binding as B, problems as P
    import oracle_py as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = P.bal_problem(9, 300, 1100, seed=11)
    cp = B.CudaProblem(spec, jacobian_format=fmt, device=-1, rank=rank, world_size=world)
    info = cp.shard_info()
    infos = [None] * world
    dist.all_gather_object(infos, info)

    # the full problem on every rank (checker), this rank's range as a sub-problem
    op = O.OracleProblem(spec, jacobian_format=fmt)
    x = op.initial_state()
    _, c_full, r_full, g_full, j_full = op.evaluate(x)
    lo, hi = info["rb_begin"], info["rb_end"]
    sub = P.ProblemSpec(
        pb_size=spec.pb_size, pb_values=spec.pb_values, rb_type=spec.rb_type[lo:hi],
        rb_pb=spec.rb_pb[2 * lo:2 * hi], fdata=spec.fdata[2 * lo:2 * hi],
        rb_loss_kind=spec.rb_loss_kind[lo:hi], rb_loss_a=spec.rb_loss_a[lo:hi],
        rb_loss_b=spec.rb_loss_b[lo:hi], num_eliminate_blocks=spec.num_eliminate_blocks)
    sop = O.OracleProblem(sub, jacobian_format=fmt, reduce=False)
    _, c, r, g, j = sop.evaluate(x)

    # one all-reduce over [gradient | cost], as the engine does with NCCL
    packed = torch.from_numpy(np.concatenate([g, [c]]))
    dist.all_reduce(packed)
    g_sum, c_sum = packed[:-1].numpy(), float(packed[-1])
    ok = True
    ok &= abs(c_sum - c_full) <= 1e-10 * abs(c_full)
    ok &= np.max(np.abs(g_sum - g_full)) <= 1e-10 * np.max(np.abs(g_full))
    # this rank's residual slice and Jacobian segments hold exactly its values
    ok &= np.array_equal(r, r_full[info["residual_begin"]:info["residual_end"]])
    covered = 0
    if fmt == 1:
        # compressed rows: one contiguous slice, same order as the sub-problem's values
        (gb, ln, lb), = info["segments"]
        ok &= np.array_equal(j[:ln], j_full[gb:gb + ln])
        covered = ln
    else:
        covered = sum(s[1] for s in info["segments"])
        ok &= np.isclose(np.sum(j[:sop.num_jacobian_values] ** 2),
                         sum(np.sum(j_full[gb:gb + ln] ** 2) for gb, ln, _ in info["segments"]),
                         rtol=1e-12)
    if rank == 0:
        q.put((bool(ok), infos, op.num_residual_blocks, op.num_residuals,
               op.num_jacobian_values, covered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fmt", [0, 1])
def test_two_ranks_shard_and_allreduce(fmt):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fmt, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, infos, nrb, nres, nnz, _ = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    # residual-block ranges and residual slices tile [0, n)
    assert infos[0]["rb_begin"] == 0 and infos[-1]["rb_end"] == nrb
    assert infos[0]["residual_begin"] == 0 and infos[-1]["residual_end"] == nres
    for a, b in zip(infos, infos[1:]):
        assert a["rb_end"] == b["rb_begin"] and a["residual_end"] == b["residual_begin"]
    assert abs(infos[0]["rb_end"] - nrb // 2) <= 1
    # Jacobian segments are disjoint and cover every value exactly once
    spans = sorted((gb, gb + ln) for info in infos for gb, ln, _ in info["segments"])
    assert spans[0][0] == 0 and spans[-1][1] == nnz
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0
    # BlockSparseMatrix with Schur ordering: an E slice and an F slice per rank
    assert all(len(i["segments"]) == (2 if fmt == 0 else 1) for i in infos)
    # local offsets pack the segments back to back
    for info in infos:
        off = 0
        for gb, ln, lb in info["segments"]:
            assert lb == off
            off += ln
