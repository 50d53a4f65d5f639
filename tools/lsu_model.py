#!/usr/bin/env python
"""Monte-Carlo model of the shared-memory wavefronts the deposit / gather of one warp of particles costs, per
deposit variant and mesh size (no GPU needed).  A 64-bit warp access is served per half-warp; a half-warp needs as
many wavefronts as the most loaded bank pair has DISTINCT addresses.  The L1TEX data pipe retires one wavefront per
clock per SM, so wavefronts per warp-particle x 21 125 warp-particles per SM (1e8 particles, 148 SMs) is a time floor.
Checked against ncu (shared wavefronts per warp of particles = l1tex__data_pipe_lsu_wavefronts_mem_shared / 3.125 M):
n_h = 16 lane-private: model 22.0 (gather 6 + read-modify-write 16), measured 24.4; n_h = 32 with the plain gather
table: model 27.8, measured 31.2 (profiles/r01d_ncu_vp_pass_nh32_before_raw.csv); n_h = 64 with the 16-fold table:
model 22.0, measured 24.4 -- the measured extra 2.4-3.4 is the zero-fill, flush and table load of every launch."""
import argparse

import numpy as np

rng = np.random.default_rng(1)


def wavefronts(addr):
    """addr: (trials, 32) 8-byte word addresses of one warp instruction -> mean wavefronts per instruction."""
    tot = 0.0
    for half in (addr[:, :16], addr[:, 16:]):
        bank = half % 16
        w = np.zeros(half.shape[0], dtype=np.int64)
        for b in range(16):
            sel = np.where(bank == b, half, -1)
            sel.sort(axis=1)
            distinct = (np.diff(sel, axis=1) != 0).sum(axis=1) + 1 - (sel == -1).any(axis=1)
            w = np.maximum(w, distinct)
        tot += w.mean()
    return tot


def model(n, K, variant, R, trials):
    rows = n + K - 1
    cell = rng.integers(0, n, size=(trials, 32))
    lane = np.arange(32)[None, :]
    out = {}
    # gather: K-1 reads of the table at cell + j
    out["gather_plain"] = sum(wavefronts(cell + j) for j in range(K - 1))
    out["gather_16x"] = sum(wavefronts((cell + j) * 16 + lane % 16) for j in range(K - 1))
    # read-modify-write of K rows: 2 accesses (LDS + STS) per row
    if variant == "priv":
        rmw = sum(wavefronts((cell + j) * 32 + lane) for j in range(K))
    else:  # R replicas per warp, lane -> replica lane % R; colliding lanes are merged first, so addresses are distinct
        rmw = sum(wavefronts((cell + j) * R + lane % R) for j in range(K))
    out["rmw"] = 2 * rmw
    out["rows_bytes_per_warp"] = rows * (32 if variant == "priv" else R) * 8
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=4000)
    ap.add_argument("--order", type=int, default=4)
    args = ap.parse_args()
    K = args.order
    print(f"{'n_h':>5s} {'variant':>10s} {'KB/warp':>8s} {'warps/SM':>8s} {'gather':>7s} {'gather16x':>9s} {'rmw':>6s} "
          f"{'total':>6s} {'LSU floor ms (1e8 particles)':>28s}")
    for n in (16, 32, 64, 128, 256, 512, 1024):
        for variant, R in (("priv", 32), ("rep16", 16), ("rep8", 8), ("rep4", 4), ("rep2", 2), ("rep1", 1)):
            m = model(n, K, variant, R, args.trials)
            kb = m["rows_bytes_per_warp"] / 1024
            warps = int(min(32, (226 - (n + K) * 8 / 1024) // kb)) if kb > 0 else 0
            if warps < 1:
                continue
            g = m["gather_plain"]
            total = g + m["rmw"] + 8          # + streaming loads/stores of the fused pass
            ms = total * 21125 / 1.965e9 * 1e3
            print(f"{n:5d} {variant:>10s} {kb:8.1f} {warps:8d} {g:7.1f} {m['gather_16x']:9.1f} {m['rmw']:6.1f} {total:6.1f} {ms:28.3f}")


if __name__ == "__main__":
    main()
