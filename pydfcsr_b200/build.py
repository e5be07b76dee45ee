"""Build libdfcsr_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdfcsr_b200.so")
SOURCES = ["api.cu", "beam.cu", "deposit.cu", "make_df.cu", "history.cu", "wake.cu", "sgolay2d.cu", "track.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dfcsr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libdfcsr_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    # measurement helpers (not part of the library): the L1 load-bandwidth probe used by bench.py and the
    # shared-memory atomic probe behind the deposit design
    for probe in ("l1_probe", "atoms_probe"):
        probe_src = os.path.join(HERE, "..", "tools", probe + ".cu")
        if os.path.exists(probe_src):
            subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o",
                            os.path.join(HERE, probe), probe_src], capture_output=True, text=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
