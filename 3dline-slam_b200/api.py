"""ctypes binding of libl3dpp_b200.so and a Python mirror of the reference's Line3D call sequence.

`Line3D` keeps the method names and argument meaning of the reference class
(include/line3D.h:71-479: addImage / UpdataImage / matchImages / reconstruct3Dlines) so that
the parity tests read like a Line3D++ driver; every compute call goes through the C ABI
(include/l3dpp_b200.h).  There is no CPU fallback: without the built extension or without a
CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libl3dpp_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "l3dpp_b200.h")
_LIB = None

MATCH_DTYPE = np.dtype([("src_cam", "<u4"), ("src_seg", "<u4"), ("tgt_cam", "<u4"), ("tgt_seg", "<u4"),
                        ("overlap", "<f4"), ("score", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"),
                        ("d_q1", "<f4"), ("d_q2", "<f4"), ("flags", "<u4")])
REC_DTYPE = np.dtype([("tgt_cam", "<u4"), ("tgt_seg", "<u4"), ("overlap", "<f4"), ("score", "<f4"),
                      ("d_p1", "<f4"), ("d_p2", "<f4"), ("d_q1", "<f4"), ("d_q2", "<f4"), ("flags", "<u4")])
ENTRY_DTYPE = np.dtype([("src_cam", "<u4"), ("src_seg", "<u4"), ("tgt_cam", "<u4"), ("tgt_seg", "<u4"),
                        ("overlap", "<f4"), ("score", "<f4"), ("d_p1", "<f4"), ("d_p2", "<f4"),
                        ("d_q1", "<f4"), ("d_q2", "<f4"), ("length", "<f4"), ("pad", "<u4"),
                        ("P1", "<f8", 3), ("P2", "<f8", 3), ("dir", "<f8", 3)])

T_NAMES = ("prep", "pairtest", "exact", "score", "affinity", "total", "k1_kernel", "k1_launches")


class View(C.Structure):
    _fields_ = [("cam_id", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("num_segs", C.c_uint32),
                ("K", C.c_double * 9), ("R", C.c_double * 9), ("t", C.c_double * 3), ("median_depth", C.c_float)]


class Params(C.Structure):
    _fields_ = [("sigma_p", C.c_float), ("sigma_a", C.c_float), ("num_neighbors", C.c_uint32),
                ("epipolar_overlap", C.c_float), ("knn", C.c_int32), ("const_reg_depth", C.c_float),
                ("max_image_width", C.c_int32), ("filter_mode", C.c_int32), ("keep_scored", C.c_int32),
                ("shard_rank", C.c_int32), ("shard_world", C.c_int32)]


class Counts(C.Structure):
    _fields_ = [("pair_tests", C.c_uint64), ("candidates", C.c_uint64), ("forward_matches", C.c_uint64),
                ("scored_entries", C.c_uint64), ("sim_evals", C.c_uint64), ("filtered_entries", C.c_uint64),
                ("pair_tests_run", C.c_uint64),
                ("num_views", C.c_uint32), ("num_pairs", C.c_uint32), ("num_pairs_local", C.c_uint32),
                ("num_entries", C.c_uint32), ("num_edges", C.c_uint32), ("num_local_ids", C.c_uint32),
                ("num_clusters", C.c_uint32), ("gpu_launches", C.c_uint32)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class L3DError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile the extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean"])
    subprocess.check_call(["make", "-C", _HERE, "-s", "-j8"])
    return LIB_PATH


def declared_symbols():
    """Every function name declared in include/l3dpp_b200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(l3d_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise L3DError("libl3dpp_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
        L.l3d_last_error.restype = C.c_char_p
        L.l3d_version.restype = C.c_char_p
        L.l3d_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
        L.l3d_ctx_destroy.argtypes = [vp]
        L.l3d_ctx_destroy.restype = None
        L.l3d_ctx_set_stream.argtypes = [vp, vp]
        L.l3d_match_lines.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp, vp, vp, u32, u32, f32, i32, i32, i32, vp, u64,
                                      C.POINTER(u64), vp]
        L.l3d_score_matches.argtypes = [vp, vp, u32, vp, u32, vp, vp, vp, vp, vp, f32, f32, f32]
        L.l3d_scene_begin.argtypes = [vp]
        L.l3d_scene_add_view.argtypes = [vp, C.POINTER(View), vp, vp, u32]
        L.l3d_scene_commit.argtypes = [vp]
        L.l3d_scene_set.argtypes = [vp, vp, u32, vp, vp, vp]
        L.l3d_scene_set_wps.argtypes = [vp, vp, u32, vp, vp, vp]
        L.l3d_scene_add_view_wps.argtypes = [vp, C.POINTER(View), vp, vp, u32]
        L.l3d_neighbors_from_worldpoints.argtypes = [vp, u32, vp, vp, u32, vp, vp]
        L.l3d_get_neighbors.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
        L.l3d_match_images.argtypes = [vp, C.POINTER(Params)]
        L.l3d_match_stage12.argtypes = [vp, C.POINTER(Params)]
        L.l3d_match_stage3.argtypes = [vp]
        L.l3d_affinity.argtypes = [vp]
        L.l3d_affinity_sparse.argtypes = [vp, C.c_int, f32, vp, vp, u32, u32]
        L.l3d_get_sparse_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.l3d_find_collinear.argtypes = [vp, vp, u32, f32, vp, u64]
        L.l3d_cluster.argtypes = [vp]
        L.l3d_lines3D.argtypes = [vp, u32]
        L.l3d_get_lines3D_counts.argtypes = [vp, vp]
        L.l3d_get_lines3D.argtypes = [vp, vp, vp, vp, vp, vp]
        L.l3d_save_lines3D_txt.argtypes = [vp, C.c_char_p]
        L.l3d_cluster_edges.argtypes = [vp, vp, u32, u32, vp]
        L.l3d_get_counts.argtypes = [vp, C.POINTER(Counts)]
        L.l3d_reset_counters.argtypes = [vp]
        L.l3d_get_timings.argtypes = [vp, vp, u32]
        L.l3d_get_pairs.argtypes = [vp, vp, u32]
        L.l3d_get_view_lists.argtypes = [vp, u32, C.c_int, vp, vp, u64, C.POINTER(u64)]
        L.l3d_get_entries.argtypes = [vp, vp, u32]
        L.l3d_get_edges.argtypes = [vp, vp, vp, u32]
        L.l3d_get_local2global.argtypes = [vp, vp, u32]
        L.l3d_get_cluster_ids.argtypes = [vp, vp, u32]
        L.l3d_get_view_info.argtypes = [vp, u32, vp, vp]
        L.l3d_get_med_scene_depth_lines.argtypes = [vp, vp]
        L.l3d_score_build.argtypes = [vp]
        L.l3d_score_fold.argtypes = [vp]
        L.l3d_affinity_edges.argtypes = [vp]
        L.l3d_affinity_ids.argtypes = [vp]
        L.l3d_shard_blob_size.argtypes = [vp, C.c_int, C.POINTER(u64)]
        L.l3d_shard_export.argtypes = [vp, C.c_int, vp, u64, C.c_int]
        L.l3d_shard_import.argtypes = [vp, C.c_int, vp, u64, C.c_int, vp, C.c_int]
        L.l3d_shard_export_hdr.argtypes = [vp, C.c_int, vp, u64]
        L.l3d_shard_forward_plan.argtypes = [vp, vp, vp]
        L.l3d_shard_forward_pack.argtypes = [vp, vp, u64, C.c_int]
        L.l3d_shard_forward_unpack.argtypes = [vp, vp, u64, C.c_int]
        L.l3d_shard_import_hdr.argtypes = [vp, C.c_int, vp, u64, C.c_int, vp, C.POINTER(C.c_int)]
        L.l3d_stream_begin.argtypes = [vp, C.c_int]
        L.l3d_stream_begin_cycle.argtypes = [vp]
        L.l3d_stream_add_image.argtypes = [vp, C.POINTER(View), vp, vp, u32]
        L.l3d_stream_delete_image.argtypes = [vp, u32]
        L.l3d_stream_update_image.argtypes = [vp, u32, vp, vp, f32, vp, u32]
        L.l3d_test_expf.argtypes = [vp, vp, vp, u32]
        L.l3d_test_acos.argtypes = [vp, vp, vp, u32]
        L.l3d_bench_fp32_peak.argtypes = [vp, vp]
        L.l3d_bench_fp64_peak.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One l3d_ctx (one GPU)."""

    def __init__(self, device: int = -1, stream: int = 0):
        self.L = lib()
        h = C.c_void_p()
        self._ck(self.L.l3d_ctx_create(C.byref(h), device))
        self.h = h
        if stream:
            self.set_stream(stream)

    def _ck(self, rc):
        if rc != 0:
            raise L3DError("l3dpp_b200 error %d: %s" % (rc, self.L.l3d_last_error().decode()))

    def set_stream(self, stream: int):
        self._ck(self.L.l3d_ctx_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            self.L.l3d_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- cudawrapper level ----
    def match_lines(self, lines_src, lines_tgt, F, RtKinv_src, RtKinv_tgt, C_src, C_tgt, src_cam, tgt_cam,
                    epi_overlap, knn, max_image_width, filter_mode=0):
        ls = np.ascontiguousarray(lines_src, dtype=np.float32)
        lt = np.ascontiguousarray(lines_tgt, dtype=np.float32)
        F, Ms, Mt, Cs, Ct = map(_f64, (F, RtKinv_src, RtKinv_tgt, C_src, C_tgt))
        cap = max(1, ls.shape[0] * max(knn, 1)) if knn > 0 else max(1, ls.shape[0] * 16)
        while True:
            out = np.zeros(cap, dtype=MATCH_DTYPE)
            off = np.zeros(ls.shape[0] + 1, dtype=np.uint32)
            n = C.c_uint64(0)
            rc = self.L.l3d_match_lines(self.h, _p(ls), ls.shape[0], _p(lt), lt.shape[0], _p(F), _p(Ms), _p(Mt),
                                        _p(Cs), _p(Ct), int(src_cam), int(tgt_cam), float(epi_overlap), int(knn),
                                        int(max_image_width), int(filter_mode), _p(out), cap, C.byref(n), _p(off))
            if rc == -4 and n.value > cap:
                cap = int(n.value)
                continue
            self._ck(rc)
            return out[:n.value], off

    def score_matches(self, lines, matches, ranges, regularizers_tgt, RtKinv, Cc, two_sigA_sqr, k,
                      min_similarity=0.5):
        lines = np.ascontiguousarray(lines, dtype=np.float32)
        matches = np.ascontiguousarray(matches, dtype=np.float32)
        ranges = np.ascontiguousarray(ranges, dtype=np.int32)
        regs = np.ascontiguousarray(regularizers_tgt, dtype=np.float32)
        M, Cc = _f64(RtKinv), _f64(Cc)
        scores = np.zeros(matches.shape[0], dtype=np.float32)
        self._ck(self.L.l3d_score_matches(self.h, _p(lines), lines.shape[0], _p(matches), matches.shape[0], _p(ranges),
                                          _p(scores), _p(regs), _p(M), _p(Cc), float(two_sigA_sqr), float(k),
                                          float(min_similarity)))
        return scores

    def find_collinear(self, lines, dist_t, row_stride=None):
        """find_collinear_segments_GPU (include/cudawrapper.h:84-86): (n, n) int8 table."""
        lines = np.ascontiguousarray(lines, dtype=np.float32).reshape(-1, 4)
        n = lines.shape[0]
        stride = int(row_stride or n)
        buf = np.zeros((max(n, 1), stride), dtype=np.int8)
        self._ck(self.L.l3d_find_collinear(self.h, _p(lines), n, float(dist_t), _p(buf), stride))
        return buf[:n, :n]

    # ---- device math hooks ----
    def test_expf(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.zeros_like(x)
        self._ck(self.L.l3d_test_expf(self.h, _p(x), _p(y), x.size))
        return y

    def test_acos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._ck(self.L.l3d_test_acos(self.h, _p(x), _p(y), x.size))
        return y

    def fp32_peak_tflops(self):
        v = C.c_float(0)
        self._ck(self.L.l3d_bench_fp32_peak(self.h, C.byref(v)))
        return float(v.value)

    def fp64_peak_tflops(self):
        v = C.c_float(0)
        self._ck(self.L.l3d_bench_fp64_peak(self.h, C.byref(v)))
        return float(v.value)


class Line3D:
    """Mirror of L3DPP::Line3D for the matching / scoring / affinity path on one B200."""

    def __init__(self, output_folder: str = "", load_segments: bool = False, max_img_width: int = -1,
                 max_line_segments: int = 3000, neighbors_by_worldpoints: bool = False, use_GPU: bool = True,
                 device: int = -1, stream: int = 0):
        if not use_GPU:
            raise L3DError("l3dpp-b200 has no CPU path (use_GPU must be True)")
        self.neighbors_by_worldpoints = bool(neighbors_by_worldpoints)
        self.ctx = Context(device, stream)
        self.L = self.ctx.L
        self.h = self.ctx.h
        self.max_img_width = int(max_img_width)
        self.max_line_segments = int(max_line_segments)
        self._views = {}
        self._packed = None
        self._dirty = True
        self.params = None
        self.filter_mode = 0
        self.keep_scored = False
        self.shard = (0, 1)

    def _ck(self, rc):
        self.ctx._ck(rc)

    # Line3D::addImage (src/line3D.cc:117-227); line segments must be given (detection is upstream)
    def addImage(self, camID, image_size, K, R, t, median_depth, wps_or_neighbors, line_segments):
        w, h = image_size
        segs = np.ascontiguousarray(line_segments, dtype=np.float32).reshape(-1, 4)
        if camID in self._views:
            raise L3DError("camera ID [%d] already in use!" % camID)
        self._views[int(camID)] = dict(w=int(w), h=int(h), K=_f64(K).reshape(9), R=_f64(R).reshape(9),
                                       t=_f64(t).reshape(3), md=float(median_depth),
                                       nb=np.ascontiguousarray(wps_or_neighbors, dtype=np.uint32), segs=segs)
        self._dirty = True

    # Line3D::UpdataImage (src/line3D.cc:433-487)
    def UpdataImage(self, camID, R, t, median_depth, wps_or_neighbors):
        v = self._views.get(int(camID))
        if v is None:
            return
        v["R"], v["t"], v["md"] = _f64(R).reshape(9), _f64(t).reshape(3), float(median_depth)
        v["nb"] = np.ascontiguousarray(wps_or_neighbors, dtype=np.uint32)
        self._dirty = True

    def load_scene(self, scene):
        wn = (lambda v: v.worldpoints) if self.neighbors_by_worldpoints else (lambda v: v.neighbors)
        for v in scene.views:
            self.addImage(v.cam_id, (v.width, v.height), v.K, v.R, v.t, v.median_depth, wn(v), v.segs)
        for v in scene.views:
            self.UpdataImage(v.cam_id, v.R, v.t, v.median_depth, wn(v))

    def neighbors(self, cam_id, cap=1024):
        """Visual neighbours used by the last matchImages (camera ids, ascending)."""
        out = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint32(0)
        self._ck(self.L.l3d_get_neighbors(self.h, int(cam_id), _p(out), cap, C.byref(n)))
        return out[:n.value].tolist()

    def _pack(self):
        """Host-side packing of the scene for l3d_scene_set (kept until a view changes)."""
        cams = list(self._views.items())
        arr = (View * max(len(cams), 1))()
        for i, (cam, v) in enumerate(cams):
            vv = arr[i]
            vv.cam_id, vv.width, vv.height, vv.num_segs = cam, v["w"], v["h"], v["segs"].shape[0]
            vv.K[:] = v["K"].tolist()
            vv.R[:] = v["R"].tolist()
            vv.t[:] = v["t"].tolist()
            vv.median_depth = v["md"]
        segs = np.ascontiguousarray(np.concatenate([v["segs"] for _, v in cams]) if cams else np.zeros((0, 4), np.float32))
        nb = np.ascontiguousarray(np.concatenate([v["nb"] for _, v in cams]) if cams else np.zeros(0, np.uint32))
        cnt = np.array([v["nb"].size for _, v in cams], dtype=np.uint32)
        self._packed = (arr, len(cams), segs, nb.astype(np.uint32), cnt)

    def upload(self):
        """Host -> device transfer of the scene tables (what DataArray::upload does per view)."""
        if self._dirty or self._packed is None:
            self._pack()
        arr, n, segs, nb, cnt = self._packed
        fn = self.L.l3d_scene_set_wps if self.neighbors_by_worldpoints else self.L.l3d_scene_set
        self._ck(fn(self.h, C.cast(arr, C.c_void_p), n, _p(segs), _p(nb), _p(cnt)))
        self._dirty = False

    def _params(self, sigma_position, sigma_angle, num_neighbors, epipolar_overlap, kNN, const_regularization_depth):
        p = Params()
        p.sigma_p, p.sigma_a, p.num_neighbors = sigma_position, sigma_angle, int(num_neighbors)
        p.epipolar_overlap, p.knn, p.const_reg_depth = epipolar_overlap, int(kNN), const_regularization_depth
        p.max_image_width = self.max_img_width
        p.filter_mode = int(self.filter_mode)
        p.keep_scored = int(self.keep_scored)
        p.shard_rank, p.shard_world = self.shard
        return p

    # Line3D::matchImages (src/line3D.cc:496-640)
    def matchImages(self, sigma_position=2.5, sigma_angle=10.0, num_neighbors=10, epipolar_overlap=0.25, kNN=10,
                    const_regularization_depth=-1.0):
        if self._dirty:
            self.upload()
        p = self._params(sigma_position, sigma_angle, num_neighbors, epipolar_overlap, kNN,
                         const_regularization_depth)
        self._ck(self.L.l3d_match_images(self.h, C.byref(p)))

    def match_stage12(self, sigma_position=2.5, sigma_angle=10.0, num_neighbors=10, epipolar_overlap=0.25, kNN=10,
                      const_regularization_depth=-1.0):
        if self._dirty:
            self.upload()
        p = self._params(sigma_position, sigma_angle, num_neighbors, epipolar_overlap, kNN,
                         const_regularization_depth)
        self._ck(self.L.l3d_match_stage12(self.h, C.byref(p)))

    def match_stage3(self):
        self._ck(self.L.l3d_match_stage3(self.h))

    # Line3D::reconstruct3Dlines up to and including clustering (src/line3D.cc:2018-2118)
    def reconstruct3Dlines(self, visibility_t=3, perform_diffusion=False, collinearity_t=-1.0, use_CERES=False):
        if perform_diffusion or use_CERES or collinearity_t > 1e-12:
            raise L3DError("diffusion / CERES / collinearity are disabled in the reference configuration "
                           "and not part of this path")
        self._ck(self.L.l3d_affinity(self.h))
        self._ck(self.L.l3d_cluster(self.h))

    def affinity(self):
        self._ck(self.L.l3d_affinity(self.h))

    # Line3D::get3Dlines (src/line3D.cc:2924-2933) for the current reconstruction: the cluster -> 3-D line tail
    def get3Dlines(self, visibility_t=3):
        """list of dict(segs (k,2,3) float64, residuals (r,2) uint32 (cam, seg), ref_view)."""
        self._ck(self.L.l3d_lines3D(self.h, int(visibility_t)))
        cnt = np.zeros(3, dtype=np.uint32)
        self._ck(self.L.l3d_get_lines3D_counts(self.h, _p(cnt)))
        n, ns, nr = (int(x) for x in cnt)
        so, ro = np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32)
        segs, res = np.zeros((max(ns, 1), 2, 3)), np.zeros((max(nr, 1), 2), np.uint32)
        rv = np.zeros(max(n, 1), np.uint32)
        self._ck(self.L.l3d_get_lines3D(self.h, _p(so), _p(segs), _p(ro), _p(res), _p(rv)))
        return [dict(segs=segs[so[i]:so[i + 1]].copy(), residuals=res[ro[i]:ro[i + 1]].copy(), ref_view=int(rv[i]))
                for i in range(n)]

    # Line3D::save3DLinesAsTXT (src/line3D.cc:3122-3178); needs get3Dlines() first
    def save3DLinesAsTXT(self, path):
        self._ck(self.L.l3d_save_lines3D_txt(self.h, str(path).encode()))

    def sparse_matrix(self, sort_by_row=False, normalization_factor=1.0):
        """SparseMatrix(A_, n, norm, sort_by_row) (src/sparsematrix.cc:8-61): (entries (E,4) f32, start_indices (n,) i32)."""
        c = self.counts()
        ent = np.zeros((max(c["num_edges"], 1), 4), dtype=np.float32)
        st = np.zeros(max(c["num_local_ids"], 1), dtype=np.int32)
        self._ck(self.L.l3d_affinity_sparse(self.h, int(bool(sort_by_row)), float(normalization_factor), _p(ent), _p(st),
                                            ent.shape[0], st.size))
        return ent[:c["num_edges"]], st[:c["num_local_ids"]]

    # ---- results ----
    def counts(self):
        c = Counts()
        self._ck(self.L.l3d_get_counts(self.h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self._ck(self.L.l3d_reset_counters(self.h))

    def timings(self):
        t = np.zeros(len(T_NAMES), dtype=np.float32)
        self._ck(self.L.l3d_get_timings(self.h, _p(t), t.size))
        return dict(zip(T_NAMES, t.tolist()))

    def pairs(self):
        n = self.counts()["num_pairs"]
        out = np.zeros((max(n, 1), 2), dtype=np.uint32)
        self._ck(self.L.l3d_get_pairs(self.h, _p(out), max(n, 1)))
        return out[:n]

    def lists(self, cam_id, which):
        nseg = self._views[int(cam_id)]["segs"].shape[0]
        off = np.zeros(nseg + 1, dtype=np.uint32)
        n = C.c_uint64(0)
        rc = self.L.l3d_get_view_lists(self.h, int(cam_id), which, _p(off), None, 0, C.byref(n))
        if rc not in (0, -4):
            self._ck(rc)
        rec = np.zeros(max(int(n.value), 1), dtype=REC_DTYPE)
        self._ck(self.L.l3d_get_view_lists(self.h, int(cam_id), which, _p(off), _p(rec), rec.size, C.byref(n)))
        return off, rec[:n.value]

    def entries(self):
        n = self.counts()["num_entries"]
        out = np.zeros(max(n, 1), dtype=ENTRY_DTYPE)
        self._ck(self.L.l3d_get_entries(self.h, _p(out), max(n, 1)))
        return out[:n]

    def edges(self):
        n = self.counts()["num_edges"]
        ij = np.zeros((max(n, 1), 2), dtype=np.int32)
        w = np.zeros(max(n, 1), dtype=np.float32)
        self._ck(self.L.l3d_get_edges(self.h, _p(ij), _p(w), max(n, 1)))
        return ij[:n], w[:n]

    def local2global(self):
        n = self.counts()["num_local_ids"]
        out = np.zeros((max(n, 1), 2), dtype=np.uint32)
        self._ck(self.L.l3d_get_local2global(self.h, _p(out), max(n, 1)))
        return out[:n]

    def cluster_ids(self):
        n = self.counts()["num_local_ids"]
        out = np.zeros(max(n, 1), dtype=np.int32)
        self._ck(self.L.l3d_get_cluster_ids(self.h, _p(out), max(n, 1)))
        return out[:n]

    def view_info(self, cam_id):
        Cc = np.zeros(3)
        kmm = np.zeros(3, dtype=np.float32)
        self._ck(self.L.l3d_get_view_info(self.h, int(cam_id), _p(Cc), _p(kmm)))
        return dict(C=Cc, k=kmm[0], median_depth=kmm[1], median_sigma=kmm[2])

    def med_scene_depth_lines(self):
        v = C.c_float(0)
        self._ck(self.L.l3d_get_med_scene_depth_lines(self.h, C.byref(v)))
        return float(v.value)

    # ---- multi-GPU: phases and exchanges of a sharded run (include/l3dpp_b200.h) ----
    def score_build(self):
        self._ck(self.L.l3d_score_build(self.h))

    def score_fold(self):
        self._ck(self.L.l3d_score_fold(self.h))

    def affinity_edges(self):
        self._ck(self.L.l3d_affinity_edges(self.h))

    def affinity_ids(self):
        self._ck(self.L.l3d_affinity_ids(self.h))

    def shard_blob_size(self, kind):
        n = C.c_uint64(0)
        self._ck(self.L.l3d_shard_blob_size(self.h, int(kind), C.byref(n)))
        return int(n.value)

    def shard_export(self, kind, ptr, cap_bytes, device_ptr):
        self._ck(self.L.l3d_shard_export(self.h, int(kind), C.c_void_p(ptr), cap_bytes, int(device_ptr)))

    def shard_import(self, kind, ptr, stride_bytes, world, sizes, device_ptr):
        sz = np.ascontiguousarray(sizes, dtype=np.uint64)
        self._ck(self.L.l3d_shard_import(self.h, int(kind), C.c_void_p(ptr), stride_bytes, int(world), _p(sz),
                                         int(device_ptr)))


    def shard_forward_plan(self):
        """(send, recv): records this rank sends to / receives from every peer (uint64 arrays of world entries)."""
        world = max(self.shard[1], 1)
        send, recv = np.zeros(world, dtype=np.uint64), np.zeros(world, dtype=np.uint64)
        self._ck(self.L.l3d_shard_forward_plan(self.h, _p(send), _p(recv)))
        return send, recv

    def shard_forward_pack(self, ptr, cap_bytes, device_ptr):
        self._ck(self.L.l3d_shard_forward_pack(self.h, C.c_void_p(ptr), int(cap_bytes), int(device_ptr)))

    def shard_forward_unpack(self, ptr, nbytes, device_ptr):
        self._ck(self.L.l3d_shard_forward_unpack(self.h, C.c_void_p(ptr), int(nbytes), int(device_ptr)))

    def shard_export_hdr(self, kind, ptr, stride_bytes):
        self._ck(self.L.l3d_shard_export_hdr(self.h, int(kind), C.c_void_p(ptr), stride_bytes))

    def shard_import_hdr(self, kind, ptr, stride_bytes, world):
        """Returns (redo, sizes): redo=True means nothing was imported (a blob did not fit)."""
        sz = np.zeros(world, dtype=np.uint64)
        redo = C.c_int(0)
        self._ck(self.L.l3d_shard_import_hdr(self.h, int(kind), C.c_void_p(ptr), stride_bytes, int(world), _p(sz),
                                             C.byref(redo)))
        return bool(redo.value), sz


class Line3DStream(Line3D):
    """The same mirror driven the way L3DPPing::Run (src/L3DPPing.cpp:98-236) drives its Line3D object:
    per cycle beginCycle(), deleteImage()*, addImage()*, UpdataImage()* for every current view, then
    matchImages() and reconstruct3Dlines().  The context keeps matched_ / processed_ / the filtered
    lists with their scores from cycle to cycle (include/l3dpp_b200.h, l3d_stream_*)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._ck(self.L.l3d_stream_begin(self.h, int(self.neighbors_by_worldpoints)))
        self._dirty = False

    def beginCycle(self):
        self._ck(self.L.l3d_stream_begin_cycle(self.h))

    def addImage(self, camID, image_size, K, R, t, median_depth, wps_or_neighbors, line_segments):
        w, h = image_size
        segs = np.ascontiguousarray(line_segments, dtype=np.float32).reshape(-1, 4)
        nb = np.ascontiguousarray(wps_or_neighbors, dtype=np.uint32)
        vv = View()
        vv.cam_id, vv.width, vv.height, vv.num_segs = int(camID), int(w), int(h), segs.shape[0]
        vv.K[:] = _f64(K).reshape(9).tolist()
        vv.R[:] = _f64(R).reshape(9).tolist()
        vv.t[:] = _f64(t).reshape(3).tolist()
        vv.median_depth = float(median_depth)
        self._ck(self.L.l3d_stream_add_image(self.h, C.byref(vv), _p(segs), _p(nb), nb.size))
        self._views[int(camID)] = dict(segs=segs)

    def deleteImage(self, camID):
        rc = self.L.l3d_stream_delete_image(self.h, int(camID))
        return rc == 0

    def UpdataImage(self, camID, R, t, median_depth, wps_or_neighbors):
        nb = np.ascontiguousarray(wps_or_neighbors, dtype=np.uint32)
        R, t = _f64(R).reshape(9), _f64(t).reshape(3)
        self._ck(self.L.l3d_stream_update_image(self.h, int(camID), _p(R), _p(t), float(median_depth), _p(nb), nb.size))

    def upload(self):
        pass


X_FORWARD, X_PROGRAMS, X_HYPOTHESES, X_EDGES = 0, 1, 2, 3


def _pack_views(views):
    arr = (View * max(len(views), 1))()
    for i, v in enumerate(views):
        vv = arr[i]
        vv.cam_id, vv.width, vv.height, vv.num_segs = v.cam_id, v.width, v.height, v.segs.shape[0]
        vv.K[:] = _f64(v.K).reshape(9).tolist()
        vv.R[:] = _f64(v.R).reshape(9).tolist()
        vv.t[:] = _f64(v.t).reshape(3).tolist()
        vv.median_depth = v.median_depth
    return arr


def neighbors_from_worldpoints(scene, num_neighbors):
    """Host-only (no GPU needed): {cam_id: [neighbour cam ids]} chosen from the views' world-point lists
    like Line3D::findVisualNeighborsFromWPs (reference src/line3D.cc:723-843)."""
    L = lib()
    arr = _pack_views(scene.views)
    wps = np.ascontiguousarray(np.concatenate([np.asarray(v.worldpoints, dtype=np.uint32) for v in scene.views]))
    cnt = np.array([len(v.worldpoints) for v in scene.views], dtype=np.uint32)
    nn = max(int(num_neighbors), 2)
    out = np.zeros((len(scene.views), nn), dtype=np.uint32)
    oc = np.zeros(len(scene.views), dtype=np.uint32)
    rc = L.l3d_neighbors_from_worldpoints(C.cast(arr, C.c_void_p), len(scene.views), _p(wps), _p(cnt), nn, _p(out), _p(oc))
    if rc:
        raise L3DError(L.l3d_last_error().decode())
    return {v.cam_id: out[i, :oc[i]].tolist() for i, v in enumerate(scene.views)}


def cluster_edges(ij, w, n):
    """Stand-alone F-H clustering (host; src/clustering.cc:7-48)."""
    L = lib()
    ij = np.ascontiguousarray(ij, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(max(n, 1), dtype=np.int32)
    rc = L.l3d_cluster_edges(_p(ij), _p(w), w.size, int(n), _p(out))
    if rc:
        raise L3DError(L.l3d_last_error().decode())
    return out[:n]


def result_digest(l3, cam_ids):
    """sha256 over everything the path hands to the clustering and to the 3-D line tail: the filtered
    match lists of every view (matches_ after filterMatches), the hypotheses (estimated_position3D_),
    A_ (pairs and weights, reference order) and local2global_.  Equal digests on every rank of a sharded
    run, and across different numbers of ranks, mean bit-identical results."""
    import hashlib
    h = hashlib.sha256()
    for cam in cam_ids:
        off, rec = l3.lists(cam, 1)
        h.update(off.tobytes())
        h.update(rec.tobytes())
    ent = l3.entries()
    for name in ENTRY_DTYPE.names:
        if name != "pad":
            h.update(np.ascontiguousarray(ent[name]).tobytes())
    ij, w = l3.edges()
    h.update(ij.tobytes())
    h.update(w.tobytes())
    h.update(l3.local2global().tobytes())
    return h.hexdigest()


def run_scene(scene, filter_mode=0, keep_scored=False, reconstruct=True, device=-1, stream=0):
    l3 = Line3D("", False, scene.max_image_width, 3000, scene.neighbors_by_worldpoints, True, device, stream)
    l3.filter_mode = filter_mode
    l3.keep_scored = keep_scored
    l3.load_scene(scene)
    p = scene.params
    l3.matchImages(p["sigma_p"], p["sigma_a"], p["num_neighbors"], p["epipolar_overlap"], p["knn"],
                   p["const_reg_depth"])
    if reconstruct:
        l3.reconstruct3Dlines()
    return l3
