"""Build libhsbp.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhsbp.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [
        os.path.join(os.path.dirname(HERE), "include", "hsbp.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++20", "-lineinfo",
           "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC,-O2",
           "-shared", "-cudart", "shared", "-o", LIB, os.path.join(CSRC, "hsbp.cu"), "-ldl", "-lcublas", "-lcusolver"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libhsbp.so")
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
