"""Summarise a B2048_PERSIST_TLOG file: per-phase cycles of the persistent trainer (mean / max over CTAs)."""
import sys
import numpy as np


def show(hdr, rows):
    a = np.array(rows, dtype=np.float64)[:, 2:]
    names = hdr.split("per step:")[1].split("(")[0].split()
    print(hdr.strip())
    print("   mean over CTAs/steps: " + "  ".join(f"{n} {v:.0f}" for n, v in zip(names, a.mean(0))) + f"  | sum {a.sum(1).mean():.0f} cycles")
    print("   max over CTAs (mean over steps): " + "  ".join(f"{n} {v:.0f}" for n, v in zip(names, a.reshape(-1, 16, a.shape[1]).mean(1).max(0))))


rows, hdr = [], None
for ln in open(sys.argv[1]):
    if ln.startswith("#"):
        if rows:
            show(hdr, rows)
        rows, hdr = [], ln
    else:
        rows.append([float(x) for x in ln.split()])
if rows:
    show(hdr, rows)
