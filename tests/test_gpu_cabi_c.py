"""The C ABI driven from plain C (no Python / torch in the process): examples/c_abi_demo.c."""
import os
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_c_host_program(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "vietnamese_qa_system_b200")
    exe = str(tmp_path / "c_abi_demo")
    cmd = [gcc, "-O2", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", libdir, "-lvqa_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK")
    assert "status -1" in r.stdout            # VQA_E_INVALID for k = 1000, reported, not aborted
