"""Driver for compute-sanitizer (memcheck / racecheck): every kernel family of the library on small inputs --
the batch path on the tiny scene (K0-K4, K2's fast path and its row-kernel fallback via duplicated segments,
the sparse matrix, the collinear table, the 3-D line tail), a sharded run of three shards in one process
(export / import / all-to-all records / adopt kernels) and a key-frame stream with deleted views.

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

api = importlib.import_module("3dline-slam_b200.api")
scene = importlib.import_module("3dline-slam_b200.scene")
sharding = importlib.import_module("3dline-slam_b200.sharding")
import stream_utils

sc = scene.make_scene("tiny")
# duplicated segments -> equal overlaps -> K2's rank kernel hands rows to the literal row kernel
for v in sc.views[:3]:
    v.segs[1::7] = v.segs[0::7][:len(v.segs[1::7])]
l3 = api.run_scene(sc, keep_scored=True)
e, s = l3.sparse_matrix(False, 1.0)
e2, s2 = l3.sparse_matrix(True, 2.0)
lines = l3.get3Dlines(3)
t = api.Context().find_collinear(sc.views[0].segs[:131], 2.0)
for knn in (0, 40):      # no pruning / more than the insertion network holds
    l3.matchImages(5.0, 10.0, 4, 0.25, knn, -1.0)
    l3.reconstruct3Dlines()
shards = []
for r in range(3):
    s3 = api.Line3D("", False, sc.max_image_width)
    s3.shard = (r, 3)
    s3.load_scene(sc)
    shards.append(s3)
grp = sharding.LocalGroup(shards)
grp.run(sc.params)
st = scene.make_stream(n_keyframes=9, n_seg=150, window=5, nbrs=4, jitter=0.2, n_world=500, cull_every=3)
l, calls = stream_utils.cuda_driver(api, st)
scene.drive_stream(st, **calls)
print("sanitize_run ok: edges %d, lines %d, collinear %d, stream entries %d, shard edges %d" %
      (len(e), len(lines), int(t.sum()), l.counts()["num_entries"], shards[0].counts()["num_edges"]))
