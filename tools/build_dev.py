"""Developer build: libdfcsr_b200_dev.so = the product sources compiled with -DDFCSR_DEV_VARIANTS (extra K4 template
instantiations selectable per launch with DFCSR_WAKE_CFG).  Use it with DFCSR_LIB=pydfcsr_b200/libdfcsr_b200_dev.so;
the product library (pydfcsr_b200/build.py) contains neither the variants nor any getenv."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pydfcsr_b200"))
import build as product  # noqa: E402

out = os.path.join(ROOT, "pydfcsr_b200", "libdfcsr_b200_dev.so")
cmd = [os.environ.get("NVCC", "nvcc")] + product.NVCC_FLAGS + ["-DDFCSR_DEV_VARIANTS", "-o", out] + \
      [os.path.join(product.CSRC, s) for s in product.SOURCES]
res = subprocess.run(cmd, capture_output=True, text=True)
sys.stderr.write(res.stdout + res.stderr)
sys.exit(res.returncode)
