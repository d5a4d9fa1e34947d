"""numpy model of K1's row sort + wedge test (csrc/k1_pairtest.cu) on the config-4 generator: per view pair, the
fraction of pair tests a warp of 32 sorted rows still has to evaluate, the spread over the warps, and what a tile
barrier costs when the eight warps of a CTA have different numbers of survivors.  CPU only; DESIGN.md 4.1 quotes it
(14 - 21 % per pair predicted, 16.4 % measured on the GPU; barrier imbalance x 1.45).

    python tools/k1_wedge_model.py [n_views]
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_wedge_rule import _rule, _sorted_order  # noqa: E402  (the same restatement the CPU test checks)


def fundamental(vs, vt):
    def skew(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    R = vt.R @ vs.R.T
    t = vt.t - R @ vs.t
    return np.linalg.inv(vt.K).T @ (skew(t) @ R) @ np.linalg.inv(vs.K)


def main():
    scene_mod = importlib.import_module("3dline-slam_b200.scene")
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    sc = scene_mod.make_scene("c4", n_views=nv)
    rows = []
    for i in range(0, nv, max(1, nv // 6)):
        vs = sc.views[i]
        for j in vs.neighbors[::7]:
            vt = sc.views[j]
            F = fundamental(vs, vt)
            S, T = vs.segs.astype(np.float64), vt.segs.astype(np.float64)
            xb = float(np.max(np.abs(T[:, [0, 2]]) + np.abs(T[:, [1, 3]])))
            _, klo, khi = _rule(F, S, T, xb, np.arange(len(S)), np.float32)
            order = _sorted_order(klo, khi)
            skip, _, _ = _rule(F, S, T, xb, order, np.float32)
            run = ~skip[order]                                  # rows in sorted order
            nw = len(S) // 32
            per_warp = run[:nw * 32].reshape(nw, 32, -1)[:, 0, :]   # a warp's rows share their survivors
            frac = per_warp.mean(axis=1)
            tiles = per_warp.shape[1] // 512
            surv = per_warp[:, :tiles * 512].reshape(nw, tiles, 512).sum(axis=2)
            cost = 74.0 * 16 + 45.0 * surv                      # instructions per tile (bench.py's SASS counts)
            nc = nw // 8
            blk = cost[:nc * 8].reshape(nc, 8, tiles)
            imbalance = (blk.max(axis=1).sum() * 8) / blk.sum()
            rows.append((i, j, frac.mean(), np.quantile(frac, [0.1, 0.5, 0.9, 1.0]), imbalance))
            print("pair %3d -> %3d   evaluated %.3f   warps 10/50/90/100 %%: %s   tile-barrier cost x %.2f"
                  % (i, j, frac.mean(), np.round(rows[-1][3], 2), imbalance))
    print("mean over the pairs: evaluated %.3f, barrier x %.2f" % (np.mean([r[2] for r in rows]), np.mean([r[4] for r in rows])))


if __name__ == "__main__":
    main()
