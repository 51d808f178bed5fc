"""Build librmp2_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["rmp2_kernels.cu", "rmp2_api.cu"]
HEADERS = ["rmp2_tables.h", "rmp2_leaves.cuh", "rmp2_step.cuh", "rmp2_launch.h",
           os.path.join("..", "..", "include", "rmp2_b200.h")]
OUT = os.path.join(CSRC, "librmp2_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, f)) <= t for f in SOURCES + HEADERS)


def build(force=False, verbose=True):
    if up_to_date() and not force:
        if verbose:
            print(f"[rmp2_b200] {OUT} is up to date")
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("RMP2_NVCC_EXTRA", "").split() + ["-o", OUT] + SOURCES
    if verbose:
        print("[rmp2_b200]", " ".join(cmd))
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building librmp2_b200.so")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
