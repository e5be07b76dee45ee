"""Where K4's in-grid samples are and how their history cells are shared (CPU, oracle state of the bench workload).

For a spread of observation points: per quadrature rectangle the in-grid fraction, how many distinct (t', z) cells
the 32 s' nodes of one warp-step touch, how often a sample's cell equals that of the SAME (x', s') sample of the
neighbouring observation point in x (same s, hence same x' and s' nodes), and how far the cell coordinates move from one
s' node to the next.  Used to decide which data-sharing mapping is worth building.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import dfcsr_oracle as O  # noqa: E402


def cells(s, x, sc, lat, hist, xp, sp):
    def orbit(q):
        return [O.interp1d(q, tab, lat.min_s, lat.delta_s) for tab in
                (lat.coords[:, 0], lat.coords[:, 1], lat.n_vec[:, 0], lat.n_vec[:, 1])]
    X0, Y0, nx0, ny0 = (v[0] for v in orbit(np.array([s])))
    X1, Y1, nx1, ny1 = orbit(sp)
    rx = X0 - X1 + x * nx0 - xp * nx1
    ry = Y0 - Y1 + x * ny0 - xp * ny1
    r = np.sqrt(rx ** 2 + ry ** 2)
    t_ret = sc.t - r
    ut = (t_ret - hist.min_x) / hist.delta_x
    uy = (xp - hist.min_y) / hist.delta_y
    uz = ((sp - t_ret) - hist.min_z) / hist.delta_z
    T, X, Z = hist.shape
    ok = (ut > -1) & (ut < T) & (uy > -1) & (uy < X) & (uz > -1) & (uz < Z)
    return ut, uz, ok


def main():
    xm, zm, sc, lat, hist = bench.oracle_state(bench.WORKLOAD)
    T, X, Z = hist.shape
    print(f"history {hist.shape}, dt {hist.delta_x:.3e} dx {hist.delta_y:.3e} dz {hist.delta_z:.3e}; sigma_x {sc.sigma_x:.3e} "
          f"sigma_z {sc.sigma_z:.3e} slope {sc.slope0:.3f} t {sc.t}")
    nzm = 64
    tot = np.zeros((4, 8))
    for ix in (4, 20, 32, 44, 58):
        for iz in (4, 20, 32, 44, 58):
            k, k2 = ix * nzm + iz, (ix + 1) * nzm + iz
            s, x, xb = sc.t + zm[k], xm[k], xm[k2]
            for r, (xa, xe, n_x, sa, se, n_s) in enumerate(O.wake_regions(s, x, sc)):
                xn, sn = np.linspace(xa, xe, n_x), np.linspace(sa, se, n_s)
                xg, sg = np.meshgrid(xn, sn, indexing="ij")
                ut, uz, ok = cells(s, x, sc, lat, hist, xg.ravel(), sg.ravel())
                utb, uzb, okb = cells(s, xb, sc, lat, hist, xg.ravel(), sg.ravel())
                ut, uz, ok = ut.reshape(xg.shape), uz.reshape(xg.shape), ok.reshape(xg.shape)
                utb, uzb, okb = utb.reshape(xg.shape), uzb.reshape(xg.shape), okb.reshape(xg.shape)
                t0, z0 = np.floor(ut).astype(int), np.floor(uz).astype(int)
                t0b, z0b = np.floor(utb).astype(int), np.floor(uzb).astype(int)
                same = ok & okb & (t0 == t0b) & (z0 == z0b)
                # warp-steps: 32 consecutive s' nodes of one x' node
                nwarp = (n_s + 31) // 32
                steps = active = distinct = pair_all = 0
                for w in range(nwarp):
                    sl = slice(32 * w, min(32 * w + 32, n_s))
                    o = ok[:, sl]
                    act = o.any(axis=1)
                    active += act.sum()
                    steps += n_x
                    key = (t0[:, sl] * (Z + 2) + z0[:, sl])
                    for i in np.nonzero(act)[0]:
                        distinct += len(np.unique(key[i][o[i]]))
                    # a paired warp-step can share its loads when every in-grid lane of A has the same cell in B
                    pa = (same[:, sl] | ~(o | okb[:, sl])).all(axis=1) & act
                    pair_all += pa.sum()
                d_uz = np.abs(np.diff(uz, axis=1))[ok[:, 1:] & ok[:, :-1]]
                d_ut = np.abs(np.diff(ut, axis=1))[ok[:, 1:] & ok[:, :-1]]
                tot[r] += [ok.size, ok.sum(), same.sum(), steps, active, distinct, pair_all,
                           0]
                if ix == 32 and iz == 32:
                    print(f"  centre point region {r}: in-grid {ok.mean():.3f}, |d uz| per node median {np.median(d_uz) if d_uz.size else 0:.3f} "
                          f"p90 {np.quantile(d_uz, .9) if d_uz.size else 0:.3f}; |d ut| median {np.median(d_ut) if d_ut.size else 0:.4f}")
    print("region  samples  in-grid  frac   same-cell-as-x-neighbour  active-warp-steps/all  cells-per-active-step  lane-util  steps-all-lanes-same")
    for r in range(3):
        n, ing, same, steps, active, distinct, pall, _ = tot[r]
        print(f"{r}  {n:.3e} {ing:.3e} {ing / n:.3f}   {same / max(ing, 1):.3f}   {active / steps:.3f}   {distinct / max(active, 1):.2f}   "
              f"{ing / max(active * 32, 1):.3f}   {pall / max(active, 1):.3f}")
    print(f"share of in-grid samples per region: {tot[:3, 1] / tot[:3, 1].sum()}")
    print(f"share of active warp-steps per region: {tot[:3, 4] / tot[:3, 4].sum()}")


if __name__ == "__main__":
    main()
