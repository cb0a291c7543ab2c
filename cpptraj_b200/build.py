"""Build cpptraj_b200/libb200rmsd.so (CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "b200_rmsd.cu")]
DEPS = SRC + [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc"))) if f.endswith((".cuh", ".h"))] + [
    os.path.join(os.path.dirname(HERE), "include", "b200_rmsd.h"), os.path.join(os.path.dirname(HERE), "include", "b200_rmsd_debug.h")]
OUT = os.environ.get("B200_RMSD_LIB_OUT") or os.path.join(HERE, "libb200rmsd.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-I", os.path.join(os.path.dirname(HERE), "include"), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-cudart", "static",
         "-Xptxas", "-v"]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    extra = os.environ.get("B200_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + extra + ["-o", OUT] + SRC
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode:
        sys.stdout.write(r.stdout)
    if r.returncode:
        raise RuntimeError("nvcc failed (%d): %s" % (r.returncode, " ".join(cmd)))
    if os.environ.get("B200_NO_PTXAS_LOG"):   # (variant builds: the committed log stays that of the product)
        return OUT
    with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as fh:
        fh.write("".join(l for l in r.stdout.splitlines(True) if "Compile time" not in l))   # (deterministic: no timings)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
