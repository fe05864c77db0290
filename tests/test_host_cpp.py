"""The C++ host side (lcqpow_b200/host): the LCQPow API in front of the C ABI.  The test binary mirrors the
reference's own suite (/root/reference/test/RunUnitTests.cpp, test/examples/*.cpp); see host/test/host_tests.cpp."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "lcqpow_b200", "host")
BIN = os.path.join(HOST, "bin")


@pytest.fixture(scope="module")
def host_bin():
    from lcqpow_b200 import build
    build.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    return BIN


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)
    return r.returncode, r.stdout + r.stderr


def test_host_unit_tests_cpu(host_bin, tmp_path):
    import torch
    args = [os.path.join(host_bin, "host_tests"), "--cpu", str(tmp_path)]
    if not torch.cuda.is_available():
        args.append("--expect-no-device")   # runSolver / SubsolverCUDA / LCQProblemBatch must fail loudly, not fall back
    rc, out = _run(args)
    assert rc == 0, out
    assert re.search(r"--cpu: \d+ checks, 0 failed", out), out


def test_host_library_exports_the_api(host_bin):
    out = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(ROOT, "lcqpow_b200", "lib", "liblcqpow_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    for sym in ("LCQPow::LCQProblem::loadLCQP(double const*", "LCQPow::LCQProblem::loadLCQP(char const*",
                "LCQPow::LCQProblem::loadLCQP(LCQPow::csc const*", "LCQPow::LCQProblem::runSolver()",
                "LCQPow::LCQProblem::getPrimalSolution(double*) const", "LCQPow::LCQProblem::getDualSolution(double*) const",
                "LCQPow::LCQProblem::switchToSparseMode()", "LCQPow::LCQProblem::switchToDenseMode()",
                "LCQPow::LCQProblem::getOutputStatistics(LCQPow::OutputStatistics&) const",
                "LCQPow::SubsolverCUDA::solve(bool, int&, int&, double const*", "LCQPow::SubsolverCUDA::getSolution(double*, double*)",
                "LCQPow::Subsolver::solve(bool, int&, int&", "LCQPow::Options::setToDefault()",
                "LCQPow::OutputStatistics::updateTrackingVectors(", "LCQPow::LCQProblemBatch::runSolver()",
                "LCQPow::Utilities::MatrixSymmetrizationProduct(double const*"):
        assert sym in out, sym


@pytest.mark.gpu
def test_host_solver_tests_gpu(host_bin, tmp_path):
    import torch
    rc, out = _run([os.path.join(host_bin, "host_tests"), "--gpu", str(tmp_path), str(torch.cuda.device_count())])
    assert rc == 0, out
    assert re.search(r"--gpu: \d+ checks, 0 failed", out), out


@pytest.mark.gpu
def test_examples_gpu(host_bin, tmp_path, example_data):
    # examples/warm_up.cpp: prints the iteration table and one of the two solutions
    rc, out = _run([os.path.join(host_bin, "warm_up")])
    assert rc == 0, out
    assert " outer |  inner |" in out and "S-stationary" in out, out
    m = re.search(r"xOpt = \[ ([-\d.e+]+), ([-\d.e+]+) \]", out)
    x = sorted(abs(float(v)) for v in m.groups())
    assert x[0] < 1e-6 and abs(x[1] - 1) < 1e-6, out
    # examples/OptimizeOnCircle.cpp: the global solution (0.1811, -0.9835)
    rc, out = _run([os.path.join(host_bin, "OptimizeOnCircle")])
    assert rc == 0, out
    m = re.search(r"xOpt = \[ ([-\d.e+]+), ([-\d.e+]+) \];  i = (\d+); k = (\d+); rho = ([\d.e+-]+)", out)
    assert abs(float(m.group(1)) - 0.181111) < 1e-5 and abs(float(m.group(2)) + 0.983483) < 1e-5, out
    # examples/solve_lcqp_from_file.cpp on the reference's example_data (written back to text files from the fixture)
    d = tmp_path / "example_data"
    d.mkdir()
    for name in ("Q", "g", "L", "R", "lbL", "ubL", "lbR", "ubR", "A", "lbA", "ubA", "lb", "ub", "x0"):
        if name in example_data:
            with open(d / f"{name}.txt", "w") as f:
                for v in np.asarray(example_data[name]).ravel():
                    f.write("Inf\n" if v == np.inf else "-Inf\n" if v == -np.inf else f"{float(v):.17g}\n")
    rc, out = _run([os.path.join(host_bin, "solve_lcqp_from_file"), str(d)])
    assert rc == 0, out
    assert "nV = 151, nC = 50, nComp = 100" in out, out
    m = re.search(r"status = (\d); x\[0..2\] = \[ ([-\d.e+]+), ([-\d.e+]+), ([-\d.e+]+) \]; i = (\d+); k = (\d+); rho = ([\d.e+-]+)", out)
    assert int(m.group(1)) == 4 and int(m.group(6)) == 8 and int(m.group(5)) == 34 and abs(float(m.group(7)) - 2.56) < 1e-12, out
    assert abs(float(m.group(2)) + 1.48) < 0.01 and abs(float(m.group(3)) + 1.36) < 0.01, out   # SURVEY.md section 4 golden row
    # the batched door from C++
    rc, out = _run([os.path.join(host_bin, "batch_circle"), "512"])
    assert rc == 0, out
    m = re.search(r"(\d+) of 512 LCQPs solved .* instance 0: x = \[ ([-\d.]+), ([-\d.]+) \]", out)
    assert int(m.group(1)) >= 490 and abs(float(m.group(2)) - 0.181110968) < 1e-6, out
