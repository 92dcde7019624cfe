"""Build libb2048.so (CUDA kernels + C-ABI) in-tree for sm_100a.  nvcc cross-compiles without a GPU.

The n-tuple agent kernels are templates on the tuple size; each size is its own translation unit
(b2048_agent_inst.cu with -DB2048_N=n) so that the five sizes compile in parallel."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libb2048.so")
COMMON = os.path.join(CSRC, "b2048_kernels.cu")
AGENT = os.path.join(CSRC, "b2048_agent_inst.cu")
SIZES = (2, 3, 4, 5, 6)
DEPS = [COMMON, AGENT] + [os.path.join(CSRC, f) for f in ("b2048_device.cuh", "b2048_host.cuh", "b2048_agent.cuh")] + \
       [os.path.join(ROOT, "include", "b2048.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def up_to_date():
    return os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(p) for p in DEPS)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    extra = (["-Xptxas=-v"] if verbose else []) + os.environ.get("B2048_NVCC_EXTRA", "").split()
    jobs = [([nvcc] + NVCC_FLAGS + extra + ["-c", COMMON, "-o", os.path.join(OBJ, "common.o")])]
    for n in SIZES:
        jobs.append([nvcc] + NVCC_FLAGS + extra + [f"-DB2048_N={n}", "-c", AGENT, "-o", os.path.join(OBJ, f"agent{n}.o")])
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        logs = list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, "common.o")] + [os.path.join(OBJ, f"agent{n}.o") for n in SIZES]
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", SO])
    if verbose:
        print("\n".join(logs))
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
