"""TEST INFRASTRUCTURE - NOT PRODUCT CODE. Replays `-m gpu` parity tests on a machine without a GPU: builds the emulated
library (build_emulated.py), installs it as the library handle of the Python mirror FOR THIS PROCESS ONLY, and runs pytest with
the arguments given. The product loader (openifem_b200/_lib.py) has no switch for this - the handle is replaced from outside.
Usage:  python tests/cpu_emul/run_tests.py tests/test_zz_insimex_gpu.py -q -m gpu"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import build_emulated  # noqa: E402

path = build_emulated.build()
import openifem_b200._lib as product_loader  # noqa: E402

handle = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
handle.ifem_last_error.restype = ctypes.c_char_p
product_loader._lib = handle
os.environ["IFEM_CPU_EMULATION"] = "1"  # lets tests that spawn compiled drivers skip themselves

import pytest  # noqa: E402

sys.exit(pytest.main(sys.argv[1:]))
