"""Lab builds: compile a VARIANT of libb2048.so with extra -D flags into 2048_b200/build/lab/<name>/libb2048.so
(only the translation units named; the other objects come from the in-tree build), for A/B runs with
`profiles/run_*.py --so <path>`.   usage: python profiles/lab_build.py NAME [common|agent4|...]... -- -DFLAG=1 ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2048_b200"))
import build as B  # noqa: E402


def main():
    args = sys.argv[1:]
    name = args[0]
    split = args.index("--") if "--" in args else len(args)
    units, flags = args[1:split] or ["common"], args[split + 1:]
    B.build()
    out = os.path.join(B.OBJ, "lab", name)
    os.makedirs(out, exist_ok=True)
    objs = []
    for u in ["common"] + [f"agent{n}" for n in B.SIZES]:
        src_obj = os.path.join(B.OBJ, u + ".o")
        if u in units:
            dst = os.path.join(out, u + ".o")
            if u == "common":
                cmd = ["nvcc"] + B.NVCC_FLAGS + flags + ["-c", B.COMMON, "-o", dst]
            else:
                cmd = ["nvcc"] + B.NVCC_FLAGS + flags + [f"-DB2048_N={u[5:]}", "-c", B.AGENT, "-o", dst]
            subprocess.run(cmd, check=True)
            objs.append(dst)
        else:
            objs.append(src_obj)
    so = os.path.join(out, "libb2048.so")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", so],
                   check=True)
    print(so)


if __name__ == "__main__":
    main()
