"""Device code on the CPU: the sources of openifem_b200/csrc are rewritten mechanically (kernel launches -> calls into a SIMT
emulator with one fiber per CUDA thread and real __syncthreads / __syncwarp / __shfl_xor_sync barriers), compiled with g++
against a stand-in cuda_runtime.h (tests/cpu_emul/) and a selection of the `-m gpu` parity tests is replayed on the result in a
child process. This is test infrastructure: it checks kernel arithmetic, indexing, shared-memory staging, launch geometry and
the host orchestration of the product sources against the oracle on a machine without a GPU; it says nothing about
performance, and the product library (nvcc, sm_100a) is a different file that never loads any of it.

The selection is sized for the CPU suite (about a minute after the one-off build); any other gpu test can be replayed with
    python tests/cpu_emul/run_tests.py <pytest arguments> -m gpu
(time-stepping goldens with thousands of kernel launches take many minutes there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "cpu_emul", "run_tests.py")

SELECTION = [
    ("tests/test_ins_gpu.py", "assembly_2d_cavity or assembly_3d_neumann_fsi"),                # verified on a B200: calibrates the emulator
    ("tests/test_linear_elasticity_gpu.py", "linear_assembly or strain_and_stress"),           # verified on a B200
    ("tests/test_scns_gpu.py", "assembly"),                                                    # verified on a B200
    ("tests/test_zz_insimex_gpu.py", "assembly"),                                              # not yet run on hardware
    ("tests/test_zz_supg_insim_gpu.py", "assembly"),
    ("tests/test_zz_kirchhoff_gpu.py", "kirchhoff"),
]


@pytest.fixture(scope="module")
def emulated_library():
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emul"))
    import build_emulated

    return build_emulated.build()


def _replay(path, expr, env=None):
    return subprocess.run([sys.executable, RUNNER, os.path.join(ROOT, path), "-q", "-x", "-m", "gpu", "-k", expr, "-p", "no:cacheprovider"],
                          capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("path,expr", SELECTION)
def test_gpu_parity_tests_pass_on_the_emulated_device(emulated_library, path, expr):
    r = _replay(path, expr)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


@pytest.mark.parametrize("path,expr", [SELECTION[0], SELECTION[3], SELECTION[4]])
def test_results_do_not_depend_on_the_thread_schedule(emulated_library, path, expr):
    """IFEM_EMUL_SHUFFLE: threads of a block and blocks of a launch run in a pseudo-random order that changes at every
    barrier pass - a kernel that relies on lane or block order (a missing barrier) would no longer reproduce the oracle"""
    r = _replay(path, expr, {"IFEM_EMUL_SHUFFLE": "12345"})
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
