"""The BAL text reader of examples/bundle_adjuster.cu (reference: examples/bal_problem.cc:72-132:
"num_cameras num_points num_observations", one "camera point x y" line per observation, then
9 doubles per camera and 3 per point, one per line).  A synthetic problem is written in that
format; the example must read back exactly what was written (CPU), and solving it must start
from the cost the oracle computes for the same problem (GPU)."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_py as O
from ceres_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "build", "examples", "bundle_adjuster")


def _write_bal(spec, path):
    nc, npts, nobs = (spec.meta[k] for k in ("num_cameras", "num_points", "num_observations"))
    rb = spec.rb_pb.reshape(-1, 2)
    obs = spec.fdata.reshape(-1, 2)
    with open(path, "w") as f:
        f.write(f"{nc} {npts} {nobs}\n")
        for (cam, pt), (x, y) in zip(rb, obs):
            f.write(f"{cam - npts} {pt} {float(x)!r} {float(y)!r}\n")  # repr: exact round trip
        for v in spec.pb_values[3 * npts:]:                   # cameras, 9 per camera
            f.write(f"{float(v)!r}\n")
        for v in spec.pb_values[:3 * npts]:                   # points
            f.write(f"{float(v)!r}\n")
    return rb[:, 0] - npts, rb[:, 1]


def test_reader_reads_back_what_was_written(tmp_path):
    spec = P.bal_problem(7, 90, 400, seed=17)
    path = str(tmp_path / "problem-7-90-pre.txt")
    cams, pts = _write_bal(spec, path)
    out = subprocess.run([EXAMPLE, f"--input={path}", "--check_input"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    m = re.search(r"cameras (\d+) points (\d+) observations (\d+) index_sum (-?\d+) value_sum (\S+)",
                  out.stdout)
    assert m, out.stdout
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (7, 90, 400)
    assert int(m.group(4)) == int(cams.sum() + 3 * pts.sum())
    # same summation order as the example: observations, cameras, points
    want = 0.0
    for v in np.concatenate([spec.fdata, spec.pb_values[270:], spec.pb_values[:270]]):
        want += float(v)
    assert float(m.group(5)) == want
    # a truncated file is refused
    with open(path) as f:
        text = f.read()
    with open(path, "w") as f:
        f.write(text[: len(text) // 2])
    out = subprocess.run([EXAMPLE, f"--input={path}", "--check_input"], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode != 0


@pytest.mark.gpu
def test_solving_a_bal_file_starts_from_the_oracle_cost(tmp_path):
    spec = P.bal_problem(9, 300, 1300, seed=18)
    path = str(tmp_path / "problem-9-300-pre.txt")
    _write_bal(spec, path)
    op = O.OracleProblem(spec, jacobian_format=0)
    ok, cost, *_ = op.evaluate(op.initial_state())
    assert ok
    for solver in ("cgnr_cuda", "iterative_schur"):
        out = subprocess.run([EXAMPLE, f"--input={path}", "--robustify", "--num_iterations=6",
                              f"--linear_solver={solver}"], capture_output=True, text=True,
                             timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        m = re.search(r"^\s*0\s+(\S+)\s", out.stdout, re.M)   # iteration 0 of the progress table
        assert m, out.stdout
        assert abs(float(m.group(1)) - cost) <= 1e-6 * cost   # printed with 7 digits
        final = re.search(r"Final\s+(\S+)", out.stdout)
        assert final and float(final.group(1)) < cost
