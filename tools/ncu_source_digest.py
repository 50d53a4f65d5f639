#!/usr/bin/env python
"""Reduce `ncu --page source --csv` (stdin) to the columns that explain these kernels, quote-aware:
address, SASS, stall samples, warp instructions executed, shared-memory wavefronts (actual / ideal)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
want = ["Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
        "L1 Wavefronts Shared Excessive"]
idx = [hdr.index(w) for w in want if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] for i in idx])
for r in rows[hdr_i + 1:]:
    if len(r) > max(idx):
        out.writerow([r[i] for i in idx])
