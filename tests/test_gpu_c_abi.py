"""The C ABI from plain C (no Python on the data path): compile tests/c_abi_smoke.c against include/dfcsr_b200.h,
link libdfcsr_b200.so + cudart, run it on the GPU."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, source):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not shutil.which("gcc") or not os.path.isdir(os.path.join(cuda, "include")):
        pytest.skip("no C toolchain / CUDA headers")
    exe = str(tmp_path / source.replace(".c", ""))
    lib_dir = os.path.join(ROOT, "pydfcsr_b200")
    cmd = ["gcc", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", source), "-o", exe, "-L", lib_dir, "-l:libdfcsr_b200.so",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{cuda}/lib64"]
    subprocess.check_call(cmd)
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)


def test_plain_c_caller(tmp_path):
    """Deposit (NGP, CIC, two-shard fixed point) and statistics from plain C."""
    out = _build_and_run(tmp_path, "c_abi_smoke.c")
    assert out.returncode == 0 and "C ABI OK" in out.stdout, out.stdout + out.stderr


def test_plain_c_caller_wake_path(tmp_path):
    """K3 -> K4 (both mappings, whole mesh and dealt out over peer grids) -> K5 from plain C."""
    out = _build_and_run(tmp_path, "c_abi_wake.c")
    assert out.returncode == 0 and "C ABI WAKE OK" in out.stdout, out.stdout + out.stderr
