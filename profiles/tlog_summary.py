"""Summarise a B2048_PERSIST_TLOG file: per-phase cycles of the persistent trainer (mean / max over CTAs)."""
import sys
import numpy as np
rows, hdr = [], None
for ln in open(sys.argv[1]):
    if ln.startswith("#"):
        if rows:
            a = np.array(rows, dtype=np.float64); rows = []
            print(hdr.strip()); print("   mean over CTAs/steps: A %.0f  B %.0f  bar1 %.0f  apply %.0f  bar2 %.0f  | sum %.0f cycles" % (*a[:, 2:].mean(0), a[:, 2:].sum(1).mean()))
            print("   max  over CTAs (mean over steps): A %.0f  B %.0f  bar1 %.0f  apply %.0f  bar2 %.0f" % tuple(a[:, 2:].reshape(-1, 16, 5).mean(1).max(0)))
        hdr = ln
    else:
        rows.append([float(x) for x in ln.split()])
if rows:
    a = np.array(rows, dtype=np.float64)
    print(hdr.strip()); print("   mean over CTAs/steps: A %.0f  B %.0f  bar1 %.0f  apply %.0f  bar2 %.0f  | sum %.0f cycles" % (*a[:, 2:].mean(0), a[:, 2:].sum(1).mean()))
    print("   max  over CTAs (mean over steps): A %.0f  B %.0f  bar1 %.0f  apply %.0f  bar2 %.0f" % tuple(a[:, 2:].reshape(-1, 16, 5).mean(1).max(0)))
