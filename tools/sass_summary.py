#!/usr/bin/env python
"""Static evidence from the build (no GPU needed): registers / spills per instantiation of the fused pass (ptxas -v
logs) and the instruction mix of its main loop per particle (cuobjdump -sass), for the lane-private cubic kernels.
Usage: python tools/sass_summary.py > profiles/<round>_sass_summary.txt   (after `make -C vlasovmethods.jl_b200/csrc`)"""
import re
import subprocess
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "vlasovmethods.jl_b200" / "csrc" / "build"
MODES = {0: "deposit", 1: "push+deposit", 2: "drift+deposit"}
VARS = {0: "lane-private", 1: "match", 2: "atomic", 3: "xor", 5: "limb-atomic"}


def parse_name(sym):
    m = re.search(r"k_vp_passILi(\d)ELi(\d)ELi(\d)ELi(\d)ELb(\d)ELb(\d)ELi(\d+)ELb(\d)E", sym)
    return tuple(int(g) for g in m.groups()) if m else None


def main():
    log = (BUILD / "vm_pass_k4.ptxas.log").read_text()
    rows = []
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers",
                         log, re.S):
        key = parse_name(m.group(1))
        if key:
            rows.append((key, int(m.group(5)), int(m.group(3)), m.group(1)))
    print("k_vp_pass<K=4, VAR, MODE, U, SPLIT, POW2, MAXT, REPG>: registers and spills (ptxas -v)\n")
    print(f"{'variant':13s} {'mode':14s} {'pairs':>5s} {'split':>5s} {'pow2':>4s} {'maxT':>5s} {'16x table':>9s} {'regs':>5s} {'spill B':>7s}")
    for key, regs, spill, _ in sorted(rows):
        k, var, mode, u, split, pow2, maxt, repg = key
        print(f"{VARS[var]:13s} {MODES[mode]:14s} {u:5d} {split:5d} {pow2:4d} {maxt:5d} {repg:9d} {regs:5d} {spill:7d}")

    print("\nMain loop of the lane-private fused pass (SPLIT = 0, POW2 = 1): SASS instructions per particle\n")
    obj = BUILD / "vm_pass_k4.o"
    for key, regs, spill, sym in sorted(rows):
        k, var, mode, u, split, pow2, maxt, repg = key
        if not (var == 0 and mode == 1 and split == 0 and pow2 == 1):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", sym, str(obj)], capture_output=True, text=True).stdout
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", sass)]
        # the particle loop = the shortest backward-branch body that holds the 16-byte streaming stores
        best = None
        for a, t in ins:
            mm = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
            if mm and int(mm.group(1), 16) < a:
                lo = int(mm.group(1), 16)
                body = [x for x in ins if lo <= x[0] <= a]
                if any("STG.E.EF.128" in x for _, x in body) and (best is None or len(body) < len(best)):
                    best = body
        body = best
        particles = 2 * u
        c = Counter()
        for _, t in body:
            op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
            c[op] += 1
        tot = sum(c.values())
        fp64 = sum(v for o, v in c.items() if o in ("DFMA", "DADD", "DMUL", "DSETP"))
        print(f"pairs={u} maxT={maxt} 16x-table={repg}: {tot / particles:6.1f} instructions/particle, fp64 {fp64 / particles:5.1f}, "
              f"LDS {c['LDS'] / particles:4.1f}, STS {c['STS'] / particles:4.1f}, LDG {c['LDG'] / particles:4.2f}, STG {c['STG'] / particles:4.2f}")


if __name__ == "__main__":
    main()
