"""The C ABI from plain C (no Python on the data path): compile tests/c_abi_smoke.c against include/dfcsr_b200.h,
link libdfcsr_b200.so + cudart, run it on the GPU."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller(tmp_path):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not shutil.which("gcc") or not os.path.isdir(os.path.join(cuda, "include")):
        pytest.skip("no C toolchain / CUDA headers")
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.join(ROOT, "pydfcsr_b200")
    cmd = ["gcc", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", lib_dir, "-l:libdfcsr_b200.so",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{cuda}/lib64"]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "C ABI OK" in out.stdout, out.stdout + out.stderr
