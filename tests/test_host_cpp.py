"""The C++ host class (pism_b200/host/SIAFD_B200.hh: the reference's SSB_Modifier / SIAFD interface over the C ABI,
on PETSc-free mirrors of PISM's containers) driven by tests/host_cpp/siafd_test.cc, which has the shape of the
reference's src/stressbalance/sia/siafd_test.cc (verification test F through the class)."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
DIR = os.path.join(HERE, "host_cpp")
EXE = os.path.join(DIR, "siafd_test")


def build():
    if not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libpism_exact.so")):
        pytest.skip("oracle/_ref/libpism_exact.so (the reference's exact solutions) has not been built")
    subprocess.run(["make", "-C", DIR, "all"], check=True, stdout=subprocess.DEVNULL)


def test_host_class_compiles_and_fails_loudly_without_a_gpu():
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    r = subprocess.run([EXE, "-Mx", "21", "-My", "21", "-Mz", "11"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "PISM ERROR" in r.stderr and "no CUDA device" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_siafd_test_F_through_the_cpp_class():
    build()
    r = subprocess.run([EXE, "-Mx", "61", "-My", "61", "-Mz", "61"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    m = re.search(r"surf vels :\s+maxUvec\s+avUvec\s+maxW\s+avW\s+([-\d.eE+]+)\s+([-\d.eE+]+)\s+([-\d.eE+]+)\s+([-\d.eE+]+)", out)
    maxU, avU, maxW, avW = (float(m.group(q)) for q in (1, 2, 3, 4))
    # vertical velocity at the surface (StressBalance::compute_vertical_velocity, 8(f) N2) against exactFG's w:
    # the reference's own 31^2 golden error after 1000 a is 0.028 / 0.004 m/a (test_17.sh, maxW / avW)
    assert maxW < 0.05 and avW < 0.01, out
    # Test F surface speeds are O(1-5 m/a); one update on a 61^3 grid is within 0.21 m/a of exact at the
    # worst point (near the margin) and 0.008 m/a on average -- the same numbers the oracle gives
    # (the reference's own golden error after 1000 a at 31^2 is 0.95 m/a, test/regression/test_17.sh)
    assert 0.0 < maxU < 0.3 and avU < 0.02, out
    vals = {k: float(v) for k, v in re.findall(r"^(D_max|sum_D|sum_absQ|sum_absU_mid) (\S+)$", out, re.M)}
    assert "flux_only_ok 1" in out
    assert re.search(r"error_path status 4: .*above top of computational grid", out), out
    # the same state through the Python mirror (tests/cases.py "F") must give the same numbers
    import gpu_util as U
    grid, cfg, inputs, gb = cases.case("F")
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    assert abs(sia.max_diffusivity() - vals["D_max"]) <= 1e-12 * vals["D_max"]
    D = cases.interior(sia.diffusivity(), cfg.w_stag)
    assert abs(D.sum() - vals["sum_D"]) <= 1e-11 * vals["sum_D"]
    u = cases.interior(sia.velocity_u(), cfg.w_uv)
    assert abs(np.abs(u[:, :, grid.Mz // 2]).sum() - vals["sum_absU_mid"]) <= 1e-11 * vals["sum_absU_mid"]


@pytest.mark.gpu
def test_pismv_test_C_through_the_cpp_classes_reproduces_test_15():
    """`pismv -test C` time-stepped in C++ (tests/host_cpp/pismv_test_C.cc: IceModel::step over StressBalance_B200 and
    GeometryEvolution_B200, host arrays in and out of the C ABI every step, like a drop-in under PISM) prints the
    reference's golden rows of test/regression/test_15.sh."""
    build()
    import pismv_oracle as PO
    for M, golden in PO.TEST_15_GOLDEN.items():
        r = subprocess.run([os.path.join(DIR, "pismv_test_C"), "-Mx", str(M), "-My", str(M)], capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0] == "NUMERICAL ERRORS evaluated at final time (relative to exact solution):"
        assert lines[1] == "geometry  :    prcntVOL        maxH         avH   relmaxETA"
        assert lines[2] == "           " + golden, (lines[2], golden)
        assert lines[3] == "NUM ERRORS DONE" and lines[4] == "steps 84"


def test_pismv_test_C_cpp_fails_loudly_without_a_gpu():
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    r = subprocess.run([os.path.join(DIR, "pismv_test_C")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def _run_ranks(size, extra, timeout=600):
    """`size` processes of pismv_test_C, one per rank, as mpiexec would start them; they share GPU 0 unless the box has
    more (CUDA IPC works between processes on one device too).  Returns the CompletedProcess of every rank."""
    import tempfile
    import torch
    import uuid
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    prefix = os.path.join(d, "siafd_b200_test_%s" % uuid.uuid4().hex)
    env = dict(os.environ, PISMV_NDEV=str(max(1, min(size, torch.cuda.device_count()))))
    procs = [subprocess.Popen([os.path.join(DIR, "pismv_test_C"), "-rank", str(r), "-size", str(size), "-prefix", prefix] + extra,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for r in range(size)]
    out = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        out.append(subprocess.CompletedProcess(p.args, p.returncode, o, e))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("size", [2, 4])
def test_pismv_test_C_on_several_ranks_in_cpp_reproduces_test_15(size):
    """The same C++ driver as one process per rank (PISM's decomposition of the 31 x 31 grid): SIAFD_B200::update on
    every rank's host arrays, the library's communicator underneath (no MPI, no torch), ghost updates and reductions of
    the host side through it.  Rank 0 prints the golden row of test/regression/test_15.sh."""
    build()
    import pismv_oracle as PO
    res = _run_ranks(size, ["-Mx", "31", "-My", "31"])
    for r in res:
        assert r.returncode == 0, (r.stdout, r.stderr)
    lines = res[0].stdout.splitlines()
    assert lines[2] == "           " + PO.TEST_15_GOLDEN[31], (lines, PO.TEST_15_GOLDEN[31])
    assert lines[4] == "steps 84"
    for q in range(1, size):
        assert res[q].stdout.strip() == "rank %d done, steps 84" % q


@pytest.mark.gpu
def test_an_error_on_one_rank_raises_on_every_rank_in_cpp():
    """A negative thickness at one point owned by rank 1: BOTH processes must throw the reference's RuntimeError
    (sia/BedSmoother.cc:303-305 under ParallelSection, util/error_handling.cc:189-214)."""
    build()
    res = _run_ranks(2, ["-Mx", "31", "-My", "31", "-poison_rank", "1"], timeout=300)
    for q, r in enumerate(res):
        assert r.returncode == 1, (q, r.stdout, r.stderr)
        assert "PISM ERROR (rank %d)" % q in r.stderr and "negative original thickness" in r.stderr, r.stderr
