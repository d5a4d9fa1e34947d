#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ (run here, in the authoring container):

  c1_nvm_scene.npz   BASELINE config 1 inputs: the 25 cameras + world-point visibility of the
                     reference's own NVM dump (/root/reference/vsfm_result.nvm, written by
                     System::SaveKeyFrameDataForLine3D, src/System.cc:459-535), 2-D segments
                     synthesised (the images are not shipped), and the visual neighbours the oracle
                     derives from the world points (Line3D::findVisualNeighborsFromWPs,
                     src/line3D.cc:723-843).  /root/reference does not exist on the GPU box, hence
                     the committed copy.
  c1_nvm_wps.npz     the world-point ids every camera of that dump observes (input of the
                     neighbors_by_worldpoints mode; with them the product must choose exactly the
                     neighbours frozen in c1_nvm_scene.npz)
  c1_nvm_expected.npz / tiny_expected.npz
                     outputs of the oracle on those inputs (pairs, filtered lists, hypotheses,
                     affinity edges, cluster ids), the regression pin for both the oracle (CPU test)
                     and the CUDA path (GPU test).

  stream_expected.npz
                     per-cycle sha256 digests and sizes of the oracle's outputs on the key-frame
                     stream tests/stream_utils.GOLDEN_STREAM (incremental mode, BASELINE config 3's
                     call sequence: delete / add / re-pose / match / reconstruct per cycle).

The reference has no golden vectors for this path, so these are oracle
outputs, not reference outputs: parity stays "unpinned" in the sense of DESIGN.md section 1.
"""
import hashlib
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
scene_mod = importlib.import_module("3dline-slam_b200.scene")
import oracle_py


def save_scene(path, sc):
    d = dict(max_image_width=sc.max_image_width, n=len(sc.views))
    for i, v in enumerate(sc.views):
        d["cam_%d" % i] = v.cam_id
        d["K_%d" % i], d["R_%d" % i], d["t_%d" % i] = v.K, v.R, v.t
        d["wh_%d" % i] = np.array([v.width, v.height])
        d["md_%d" % i] = v.median_depth
        d["segs_%d" % i] = v.segs
        d["nb_%d" % i] = np.asarray(v.neighbors, np.uint32)
    np.savez_compressed(path, **d)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def expected(o, sc):
    """Compact expectation: small arrays in full, large ones as SHA-256 of their raw bytes."""
    e = o.entries()
    ij, w = o.edges()
    d = dict(pairs=o.pairs(), n_entries=len(e), entries_sha=digest(e), entry_keys=np.stack(
                 [e["src_cam"], e["src_seg"], e["tgt_cam"], e["tgt_seg"]], axis=1).astype(np.uint32),
             entry_scores=e["score"].copy(), n_edges=len(w), edges_ij_sha=digest(ij), edges_w_sha=digest(w),
             edges_head_ij=ij[:512].copy(), edges_head_w=w[:512].copy(), local2global_sha=digest(o.local2global()),
             cluster_ids=o.cluster_ids(), med_scene_depth_lines=np.float32(o.med_scene_depth_lines()))
    for v in sc.views:
        off, rec = o.lists(v.cam_id, 1)
        d["filt_n_%d" % v.cam_id] = len(rec)
        d["filt_off_sha_%d" % v.cam_id], d["filt_rec_sha_%d" % v.cam_id] = digest(off), digest(rec)
        d["k_%d" % v.cam_id] = np.float32(o.view_info(v.cam_id)["k"])
        d["md_%d" % v.cam_id] = np.float32(o.view_info(v.cam_id)["median_depth"])
    return d


def main():
    # ---- C1: NVM cameras, neighbours from world points ----
    sc = scene_mod.scene_from_nvm("/root/reference/vsfm_result.nvm", n_seg=300, n_world=260)
    o = oracle_py.OracleLine3D(sc.max_image_width, True)
    o.load_scene(sc)
    p = sc.params
    o.match_images(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
    # the world-point lists themselves (neighbors_by_worldpoints input of the product / of the oracle)
    np.savez_compressed(os.path.join(HERE, "c1_nvm_wps.npz"),
                        **{"wps_%d" % v.cam_id: np.asarray(v.worldpoints, np.uint32) for v in sc.views})
    for v in sc.views:                       # freeze the world-point neighbours as explicit lists
        v.neighbors = o.neighbors(v.cam_id)
        v.worldpoints = None
    sc.neighbors_by_worldpoints = False
    sc.views = [v for v in sc.views if len(v.neighbors) > 0]
    ids = {v.cam_id for v in sc.views}
    for v in sc.views:
        v.neighbors = [n for n in v.neighbors if n in ids]
    save_scene(os.path.join(HERE, "c1_nvm_scene.npz"), sc)
    o2 = oracle_py.run_scene(sc)
    np.savez_compressed(os.path.join(HERE, "c1_nvm_expected.npz"), **expected(o2, sc))
    print("c1: views", len(sc.views), "pairs", len(o2.pairs()), "entries", len(o2.entries()), "edges", len(o2.edges()[1]),
          "clusters", len(set(o2.cluster_ids().tolist())))
    # ---- tiny synthetic scene ----
    st = scene_mod.make_scene("tiny")
    ot = oracle_py.run_scene(st)
    np.savez_compressed(os.path.join(HERE, "tiny_expected.npz"), **expected(ot, st))
    print("tiny: pairs", len(ot.pairs()), "entries", len(ot.entries()), "edges", len(ot.edges()[1]))
    # ---- key-frame stream (incremental mode) ----
    sys.path.insert(0, os.path.dirname(HERE))
    import stream_utils
    stream = scene_mod.make_stream(**stream_utils.GOLDEN_STREAM)
    o3, calls = stream_utils.oracle_driver(oracle_py, stream)
    dig, sizes = stream_utils.stream_digests(stream, o3, calls)
    np.savez_compressed(os.path.join(HERE, "stream_expected.npz"), digests=dig, sizes=sizes)
    print("stream: cycles", len(dig), "entries/edges/ids per cycle", sizes.tolist())


if __name__ == "__main__":
    main()
