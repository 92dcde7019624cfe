"""Build libb2048.so (CUDA kernels + C-ABI) in-tree for sm_100a.  nvcc cross-compiles without a GPU."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libb2048.so")
SOURCES = [os.path.join(CSRC, "b2048_kernels.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "b2048_device.cuh"), os.path.join(ROOT, "include", "b2048.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared"]


def up_to_date():
    return os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(p) for p in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + SOURCES + ["-o", SO]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
