"""Developer aid: per-row cycle counters of the scoring wavefront (L3D_WF_DEBUG)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["L3D_WF_DEBUG"] = "gpurun_out/wf_debug.bin"
api = importlib.import_module("3dline-slam_b200.api"); scene = importlib.import_module("3dline-slam_b200.scene")
sc = scene.make_scene("c2")
l3 = api.run_scene(sc, reconstruct=False)
l3.matchImages(5.0, 10.0, 10, 0.25, 10, -1.0)
d = np.fromfile("gpurun_out/wf_debug.bin", dtype=np.uint32).reshape(-1, 4)
N = 1000
for v in (0, 1, 10, 25, 49):
    r = d[v * N:(v + 1) * N]
    tot = r[:, 0].astype(np.int64) + r[:, 1]
    i = int(np.argmax(tot))
    T = r[:, 3].astype(np.int64); T[T == 0xffffffff] = -1
    print("view %2d: gather cyc mean %6.0f max %6d | score cyc mean %6.0f max %7d | m mean %.1f max %d | flagged pairs mean %.0f max %d | slowest row: m=%d gather=%d score=%d T=%d"
          % (v, r[:, 0].mean(), r[:, 0].max(), r[:, 1].mean(), r[:, 1].max(), r[:, 2].mean(), r[:, 2].max(), T.mean(), T.max(), r[i, 2], r[i, 0], r[i, 1], T[i]))
    order = np.argsort(-tot)[:5]
    print("   top rows (m, gather, score, T):", [(int(r[j, 2]), int(r[j, 0]), int(r[j, 1]), int(T[j])) for j in order])
print(l3.timings())
