"""ctypes binding of the CPU oracle (oracle/libl3d_oracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REC_DTYPE = np.dtype([("tgt_cam", "<u4"), ("tgt_seg", "<u4"), ("overlap", "<f4"), ("score", "<f4"),
                      ("d_p1", "<f4"), ("d_p2", "<f4"), ("d_q1", "<f4"), ("d_q2", "<f4"),
                      ("flags", "<u4")])
ENTRY_DTYPE = np.dtype([("src_cam", "<u4"), ("src_seg", "<u4"), ("tgt_cam", "<u4"), ("tgt_seg", "<u4"),
                        ("overlap", "<f4"), ("score", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"),
                        ("d_q1", "<f4"), ("d_q2", "<f4"), ("length", "<f4"), ("pad", "<u4"),
                        ("P1", "<f8", 3), ("P2", "<f8", 3), ("dir", "<f8", 3)])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libl3d_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("l3d_oracle.cpp", "detmath.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    build_ref()
    return so


REF_ROOT = "/root/reference"
REF_CLUSTERING = os.path.join(_HERE, "_ref", "libref_clustering.so")


REF_LINE3D = {"det": os.path.join(_HERE, "_ref", "libref_line3d_det.so"),
              "libm": os.path.join(_HERE, "_ref", "libref_line3d_libm.so"),
              "omp": os.path.join(_HERE, "_ref", "libref_line3d_omp.so")}   # timing only: order depends on the schedule


def build_ref() -> bool:
    """oracle/_ref: the reference's own sources compiled in place (only where /root/reference exists, i.e. in
    the authoring container; the GPU box uses the prebuilt files that travelled with the repo): the graph
    clustering, and the Line3D++ path itself (src/line3D.cc + src/view.cc against oracle/standin/)."""
    if os.path.exists(os.path.join(REF_ROOT, "src", "clustering.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(REF_CLUSTERING)


def ref_cluster(ij, w, n, c=3.0):
    """The reference's L3DPP::performClustering itself (oracle/_ref); None if the library is not there."""
    if not os.path.exists(REF_CLUSTERING) and not build_ref():
        return None
    L = C.CDLL(REF_CLUSTERING)
    ij = np.ascontiguousarray(ij, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(max(n, 1), dtype=np.int32)
    L.ref_cluster.restype = C.c_int
    L.ref_cluster(_p(ij), _p(w), len(w), int(n), C.c_float(c), _p(out))
    return out[:n]


def _bind_common(L):
    """argtypes of the entry points the restatement and the compiled reference share."""
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_set_snapshot.argtypes = [C.c_void_p, C.c_int]
    L.orc_add_image.restype = C.c_int
    L.orc_add_image.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                C.c_uint, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.orc_update_image.restype = C.c_int
    L.orc_update_image.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int]
    L.orc_delete_image.restype = C.c_int
    L.orc_delete_image.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_begin_cycle.argtypes = [C.c_void_p]
    L.orc_match_images.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint, C.c_float, C.c_int, C.c_float]
    L.orc_reconstruct.argtypes = [C.c_void_p]
    L.orc_reconstruct_collin.argtypes = [C.c_void_p, C.c_float]
    L.orc_num_pairs.restype = C.c_int
    L.orc_num_pairs.argtypes = [C.c_void_p]
    L.orc_get_pairs.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_pair_tests.restype = C.c_uint64
    L.orc_pair_tests.argtypes = [C.c_void_p]
    L.orc_list_total.restype = C.c_uint64
    L.orc_list_total.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
    L.orc_get_lists.restype = C.c_int
    L.orc_get_lists.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_num_entries.restype = C.c_int
    L.orc_num_entries.argtypes = [C.c_void_p]
    L.orc_get_entries.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_num_edges.restype = C.c_int
    L.orc_num_edges.argtypes = [C.c_void_p]
    L.orc_get_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_num_local.restype = C.c_int
    L.orc_num_local.argtypes = [C.c_void_p]
    L.orc_get_local2global.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_get_cluster_ids.restype = C.c_int
    L.orc_get_cluster_ids.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_get_view_info.restype = C.c_int
    L.orc_get_view_info.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    L.orc_get_neighbors.restype = C.c_int
    L.orc_get_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    L.orc_med_scene_depth_lines.restype = C.c_float
    L.orc_med_scene_depth_lines.argtypes = [C.c_void_p]
    return L


_REF_LIBS = {}


def ref_lib(kind="det"):
    """The reference's Line3D++ sources compiled here (oracle/_ref/libref_line3d_<kind>.so); None if absent."""
    if kind not in _REF_LIBS:
        path = REF_LINE3D[kind]
        if not os.path.exists(path):
            build_ref()
        if not os.path.exists(path):
            _REF_LIBS[kind] = None
        else:
            L = _bind_common(C.CDLL(path))
            L.ref_lines3D_counts.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_lines3D_get.argtypes = [C.c_void_p] * 6
            L.ref_save_txt.argtypes = [C.c_void_p, C.c_char_p]
            _REF_LIBS[kind] = L
    return _REF_LIBS[kind]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_snapshot.argtypes = [C.c_void_p, C.c_int]
        L.orc_add_image.restype = C.c_int
        L.orc_add_image.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                    C.c_uint, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_update_image.restype = C.c_int
        L.orc_update_image.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_void_p, C.c_int]
        L.orc_delete_image.restype = C.c_int
        L.orc_delete_image.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_begin_cycle.argtypes = [C.c_void_p]
        L.orc_match_images.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint, C.c_float, C.c_int,
                                       C.c_float]
        L.orc_reconstruct.argtypes = [C.c_void_p]
        L.orc_num_pairs.restype = C.c_int
        L.orc_num_pairs.argtypes = [C.c_void_p]
        L.orc_get_pairs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_pair_tests.restype = C.c_uint64
        L.orc_pair_tests.argtypes = [C.c_void_p]
        L.orc_list_total.restype = C.c_uint64
        L.orc_list_total.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.orc_get_lists.restype = C.c_int
        L.orc_get_lists.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_num_entries.restype = C.c_int
        L.orc_num_entries.argtypes = [C.c_void_p]
        L.orc_get_entries.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_num_edges.restype = C.c_int
        L.orc_num_edges.argtypes = [C.c_void_p]
        L.orc_get_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_num_local.restype = C.c_int
        L.orc_num_local.argtypes = [C.c_void_p]
        L.orc_get_local2global.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_cluster_ids.restype = C.c_int
        L.orc_get_cluster_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_view_info.restype = C.c_int
        L.orc_get_view_info.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_get_neighbors.restype = C.c_int
        L.orc_get_neighbors.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.orc_med_scene_depth_lines.restype = C.c_float
        L.orc_med_scene_depth_lines.argtypes = [C.c_void_p]
        L.orc_get_translation.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_timers.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_kat_expf.restype = C.c_float
        L.orc_kat_expf.argtypes = [C.c_float]
        L.orc_kat_acos.restype = C.c_double
        L.orc_kat_acos.argtypes = [C.c_double]
        L.orc_kat_acosf.restype = C.c_float
        L.orc_kat_acosf.argtypes = [C.c_float]
        L.orc_kat_sin.restype = C.c_double
        L.orc_kat_sin.argtypes = [C.c_double]
        L.orc_kat_mutual_overlap.restype = C.c_float
        L.orc_kat_mutual_overlap.argtypes = [C.c_void_p]
        L.orc_kat_fundamental.argtypes = [C.c_void_p] * 7
        L.orc_kat_inverse3.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_kat_angle.restype = C.c_float
        L.orc_kat_angle.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_kat_dist_point_line.restype = C.c_float
        L.orc_kat_dist_point_line.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_kat_cluster.restype = C.c_int
        L.orc_kat_cluster.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_match_only.restype = C.c_int
        L.orc_match_only.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_int] + [C.c_void_p] * 5
        L.orc_score_packed.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleLine3D:
    """Mirror of the reference's Line3D call sequence (addImage / UpdataImage / deleteImage /
    matchImages / reconstruct3Dlines) on the CPU oracle."""

    TIMER_NAMES = ("match", "orient", "score", "inverse", "filter", "update", "affinity", "cluster",
                   "match_images", "reconstruct")

    def __init__(self, max_img_width: int, neighbors_by_worldpoints: bool = False, threads: int = 0,
                 snapshot: bool = True, library=None):
        self.L = library if library is not None else lib()
        self.L.orc_set_threads(threads)
        self.h = self.L.orc_create(int(max_img_width), int(bool(neighbors_by_worldpoints)))
        self.L.orc_set_snapshot(self.h, int(snapshot))
        self.nseg = {}

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_image(self, cam_id, K, R, t, width, height, median_depth, wps_or_nbrs, segs):
        K, R, t = _f64(K), _f64(R), _f64(t)
        wn = np.ascontiguousarray(wps_or_nbrs, dtype=np.uint32)
        segs = np.ascontiguousarray(segs, dtype=np.float32)
        self.nseg[int(cam_id)] = segs.shape[0]
        return self.L.orc_add_image(self.h, int(cam_id), _p(K), _p(R), _p(t), int(width), int(height),
                                    float(median_depth), _p(wn), wn.size, _p(segs), segs.shape[0])

    def update_image(self, cam_id, R, t, median_depth, wps_or_nbrs):
        R, t = _f64(R), _f64(t)
        wn = np.ascontiguousarray(wps_or_nbrs, dtype=np.uint32)
        return self.L.orc_update_image(self.h, int(cam_id), _p(R), _p(t), float(median_depth), _p(wn),
                                       wn.size)

    def delete_image(self, cam_id):
        return self.L.orc_delete_image(self.h, int(cam_id))

    def begin_cycle(self):
        self.L.orc_begin_cycle(self.h)

    def load_scene(self, scene):
        for v in scene.views:
            wn = v.worldpoints if scene.neighbors_by_worldpoints else v.neighbors
            rc = self.add_image(v.cam_id, v.K, v.R, v.t, v.width, v.height, v.median_depth, wn, v.segs)
            assert rc == 0, rc
        for v in scene.views:  # L3DPPing.cpp:190-205: UpdataImage for all active key frames
            wn = v.worldpoints if scene.neighbors_by_worldpoints else v.neighbors
            self.update_image(v.cam_id, v.R, v.t, v.median_depth, wn)

    def match_images(self, sigma_p=5.0, sigma_a=10.0, num_neighbors=10, epipolar_overlap=0.25, knn=10,
                     const_reg_depth=-1.0):
        self.L.orc_match_images(self.h, sigma_p, sigma_a, int(num_neighbors), epipolar_overlap, int(knn),
                                const_reg_depth)

    def reconstruct(self, collinearity_t=-1.0):
        """Line3D::reconstruct3Dlines up to the clustering; collinearity_t > 0 adds the links to collinear
        segments (src/line3D.cc:2328-2396), which the CUDA product does not build yet."""
        if collinearity_t > 1e-12:
            self.L.orc_reconstruct_collin(self.h, C.c_float(collinearity_t))
        else:
            self.L.orc_reconstruct(self.h)

    # ---- results ----
    def pairs(self):
        n = self.L.orc_num_pairs(self.h)
        out = np.zeros((n, 2), dtype=np.uint32)
        if n:
            self.L.orc_get_pairs(self.h, _p(out))
        return out

    def pair_tests(self):
        return int(self.L.orc_pair_tests(self.h))

    def lists(self, cam_id, which):
        """which=0: lists right after scoring (pre-filter); 1: current (filtered) lists."""
        n = int(self.L.orc_list_total(self.h, int(cam_id), which))
        rows = self.nseg[int(cam_id)]
        off = np.zeros(rows + 1, dtype=np.uint32)
        rec = np.zeros(max(n, 1), dtype=REC_DTYPE)
        rc = self.L.orc_get_lists(self.h, int(cam_id), which, _p(off), _p(rec))
        assert rc == 0
        return off, rec[:n]

    def entries(self):
        n = self.L.orc_num_entries(self.h)
        out = np.zeros(max(n, 1), dtype=ENTRY_DTYPE)
        if n:
            self.L.orc_get_entries(self.h, _p(out))
        return out[:n]

    def edges(self):
        n = self.L.orc_num_edges(self.h)
        ij = np.zeros((max(n, 1), 2), dtype=np.int32)
        w = np.zeros(max(n, 1), dtype=np.float32)
        if n:
            self.L.orc_get_edges(self.h, _p(ij), _p(w))
        return ij[:n], w[:n]

    def local2global(self):
        n = self.L.orc_num_local(self.h)
        out = np.zeros((max(n, 1), 2), dtype=np.uint32)
        if n:
            self.L.orc_get_local2global(self.h, _p(out))
        return out[:n]

    def cluster_ids(self):
        n = self.L.orc_num_local(self.h)
        out = np.zeros(max(n, 1), dtype=np.int32)
        m = self.L.orc_get_cluster_ids(self.h, _p(out))
        return out[:m]

    def view_info(self, cam_id):
        Cc = np.zeros(3)
        kmm = np.zeros(3, dtype=np.float32)
        rc = self.L.orc_get_view_info(self.h, int(cam_id), _p(Cc), _p(kmm))
        assert rc == 0
        return dict(C=Cc, k=kmm[0], median_depth=kmm[1], median_sigma=kmm[2])

    def match_camera(self, cam_id):
        """(RtKinv 3x3, translated centre) exactly as the last match_images used them (test hook)."""
        M = np.zeros(9)
        Cc = np.zeros(3)
        rc = self.L.orc_get_match_camera(self.h, int(cam_id), _p(M), _p(Cc))
        assert rc == 0
        return M.reshape(3, 3), Cc

    def fundamental(self, src, tgt):
        """F cached by the last match_images for the ordered pair (src, tgt), or None (test hook)."""
        F = np.zeros(9)
        if self.L.orc_get_fundamental(self.h, int(src), int(tgt), _p(F)) != 0:
            return None
        return F.reshape(3, 3)

    def neighbors(self, cam_id, cap=256):
        out = np.zeros(cap, dtype=np.uint32)
        n = self.L.orc_get_neighbors(self.h, int(cam_id), _p(out), cap)
        return out[:max(n, 0)].tolist()

    def med_scene_depth_lines(self):
        return float(self.L.orc_med_scene_depth_lines(self.h))

    def match_only(self, src, tgt, epi_overlap=0.25, knn=10):
        """translate + F + matchingCPU(src,tgt) only; returns (F, RtKinv_src, RtKinv_tgt, C_src, C_tgt)."""
        F, Ms, Mt = np.zeros(9), np.zeros(9), np.zeros(9)
        Cs, Ct = np.zeros(3), np.zeros(3)
        rc = self.L.orc_match_only(self.h, int(src), int(tgt), float(epi_overlap), int(knn), _p(F), _p(Ms), _p(Mt),
                                   _p(Cs), _p(Ct))
        assert rc == 0
        return F, Ms, Mt, Cs, Ct

    def timers(self):
        t = np.zeros(10)
        self.L.orc_get_timers(self.h, _p(t))
        return dict(zip(self.TIMER_NAMES, t.tolist()))


class RefLine3D(OracleLine3D):
    """The same driver on the REFERENCE's own Line3D class (src/line3D.cc + src/view.cc compiled into oracle/_ref,
    see oracle/ref_line3d_wrap.cpp).  kind: "det" (deterministic libm replacements, bit-comparable with the
    restatement) or "libm" (glibc)."""

    def __init__(self, max_img_width, neighbors_by_worldpoints=False, kind="det"):
        L = ref_lib(kind)
        if L is None:
            raise RuntimeError("oracle/_ref/libref_line3d_%s.so is not built" % kind)
        super().__init__(max_img_width, neighbors_by_worldpoints, 0, True, library=L)

    def lines3D(self):
        """Final 3-D lines: list of dicts(segs (k,2,3), residuals (r,2), ref_view)."""
        cnt = np.zeros(3, dtype=np.uint32)
        self.L.ref_lines3D_counts(self.h, _p(cnt))
        n, ns, nr = (int(x) for x in cnt)
        so, ro = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
        segs, res, rv = np.zeros((max(ns, 1), 2, 3)), np.zeros((max(nr, 1), 2), np.uint32), np.zeros(max(n, 1), np.uint32)
        self.L.ref_lines3D_get(self.h, _p(so), _p(segs), _p(ro), _p(res), _p(rv))
        return [dict(segs=segs[so[i]:so[i + 1]].copy(), residuals=res[ro[i]:ro[i + 1]].copy(), ref_view=int(rv[i]))
                for i in range(n)]

    def save_txt(self, folder):
        self.L.ref_save_txt(self.h, folder.encode())


def run_scene_ref(scene, kind="det", reconstruct=True):
    o = RefLine3D(scene.max_image_width, scene.neighbors_by_worldpoints, kind)
    o.load_scene(scene)
    p = scene.params
    o.match_images(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"], p["const_reg_depth"])
    if reconstruct:
        o.reconstruct()
    return o


def run_scene(scene, threads=0, snapshot=True, reconstruct=True):
    o = OracleLine3D(scene.max_image_width, scene.neighbors_by_worldpoints, threads, snapshot)
    o.load_scene(scene)
    p = scene.params
    o.match_images(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                   p["const_reg_depth"])
    if reconstruct:
        o.reconstruct()
    return o


def kat_cluster(ij, w, n):
    """The oracle's restatement of performClustering + the id read-out of clusterSegments (c = 3)."""
    ij = np.ascontiguousarray(ij, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(max(n, 1), dtype=np.int32)
    lib().orc_kat_cluster(_p(ij), _p(w), len(w), int(n), _p(out))
    return out[:n]


def sparse_matrix(ij, w, n, norm=1.0, sort_by_row=False):
    """SparseMatrix::SparseMatrix (reference src/sparsematrix.cc:8-61): (entries float4[E], start_indices int[n])."""
    ij = np.ascontiguousarray(ij, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    ent = np.zeros((max(len(w), 1), 4), dtype=np.float32)
    st = np.zeros(max(n, 1), dtype=np.int32)
    lib().orc_sparse_matrix(_p(ij), _p(w), len(w), int(n), C.c_float(norm), int(bool(sort_by_row)), _p(ent), _p(st))
    return ent[:len(w)], st[:n]


def find_collinear(lines, dist_t):
    """View::findCollinCPU (reference src/view.cc:238-293): (n,n) int8, [r,c] = 1 iff c is collinear to r."""
    lines = np.ascontiguousarray(lines, dtype=np.float32).reshape(-1, 4)
    n = lines.shape[0]
    out = np.zeros((n, n), dtype=np.int8)
    lib().orc_find_collinear(_p(lines), n, C.c_float(dist_t), _p(out))
    return out


def score_packed(lines, matches, ranges, regs_tgt, RtKinv, Cc, two_sigA_sqr, k, min_sim=0.5):
    """Oracle restatement of scoringCPU's new-match branch over scoringGPU's packed buffers."""
    L = lib()
    lines = np.ascontiguousarray(lines, dtype=np.float32)
    matches = np.ascontiguousarray(matches, dtype=np.float32)
    ranges = np.ascontiguousarray(ranges, dtype=np.int32)
    regs = np.ascontiguousarray(regs_tgt, dtype=np.float32)
    M, Cc = _f64(RtKinv), _f64(Cc)
    out = np.zeros(matches.shape[0], dtype=np.float32)
    L.orc_score_packed(_p(lines), lines.shape[0], _p(matches), matches.shape[0], _p(ranges), _p(regs), _p(M), _p(Cc),
                       float(two_sigA_sqr), float(k), float(min_sim), _p(out))
    return out
