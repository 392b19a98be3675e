"""The C++ facade (include/openifem/openifem.h) compiles against the C ABI with plain g++ (CPU check) and
the reference's fluid_pipe_mpi driver passes its golden assertion through it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "fluid_pipe_mpi")
EXE_CYL = os.path.join(ROOT, "tests", "cpp", "_build", "fluid_cylinder_mpi")


def _build(name="fluid_pipe_mpi"):
    from openifem_b200 import build

    lib = build.build()
    exe = os.path.join(ROOT, "tests", "cpp", "_build", name)
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-o", exe, "-L", os.path.dirname(lib), "-lopenifem_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.check_call(cmd)


def test_cpp_driver_compiles_and_fails_loudly_without_gpu(golden_dir):
    _build()
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE, os.path.join(golden_dir, "ins_pipe_2d.prm")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_reference_golden(golden_dir):
    if os.environ.get("IFEM_CPU_EMULATION"):
        pytest.skip("compiled drivers link the product library: not replayable on the emulated device")
    _build()
    r = subprocess.run([EXE, os.path.join(golden_dir, "ins_pipe_2d.prm")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout


def test_cpp_cylinder_driver_compiles_and_fails_loudly_without_gpu(golden_dir):
    """tests/fluid_cylinder_mpi through the facade: Utils::GridCreator<2>::flow_around_cylinder and
    add_hard_coded_boundary_condition are host-side and run; the solver construction needs the device"""
    _build("fluid_cylinder_mpi")
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE_CYL, os.path.join(golden_dir, "ins_cylinder_2d.prm")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_cylinder_driver_reference_golden(golden_dir):
    if os.environ.get("IFEM_CPU_EMULATION"):
        pytest.skip("compiled drivers link the product library: not replayable on the emulated device")
    _build("fluid_cylinder_mpi")
    r = subprocess.run([EXE_CYL, os.path.join(golden_dir, "ins_cylinder_2d.prm")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout


def test_cpp_contact_driver_compiles_and_fails_loudly_without_gpu(golden_dir):
    """tests/fsi_contact_model_mpi through the facade (SCnsIM + SharedLinearElasticity + MPI::FSI with a penetration
    criterion, GridTools::shift, Tensor<1, dim>): the host side compiles and runs; the solver construction needs the device"""
    _build("fsi_contact_model_mpi")
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "fsi_contact_model_mpi")
    r = subprocess.run([exe, os.path.join(golden_dir, "fsi_contact_model_2d.prm")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.parametrize("name,args", [("fluid_cylinder_mpi_insimex", ["ins_cylinder_2d.prm"]),
                                       ("fluid_supg_insim_mpi", ["wall", "supg_ins_plane_wall_driven_2d.prm"])])
def test_cpp_new_solver_drivers_compile_and_fail_loudly_without_gpu(golden_dir, name, args):
    """Fluid::MPI::InsIMEX / SUPGInsIM through the facade (PETScWrappers::MPI::Vector::l2_norm, the serial Vector copy the
    drivers sort): the host side compiles and runs; the solver construction needs the device"""
    _build(name)
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    exe = os.path.join(ROOT, "tests", "cpp", "_build", name)
    argv = [a if not a.endswith(".prm") else os.path.join(golden_dir, a) for a in args]
    r = subprocess.run([exe] + argv, capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr
