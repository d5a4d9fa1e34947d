"""GPU: the cudawrapper-level C ABI entry points (host buffers in/out) against the oracle, and the
error behaviour / edge cases of the Line3D-level entry points."""
import numpy as np
import pytest

from parity_utils import assert_struct_equal, compare_full

pytestmark = pytest.mark.gpu


def _two_view_scene(scene_mod, seed=3, n_seg=220):
    sc = scene_mod.make_scene("tiny", seed=seed, n_views=4, n_seg=n_seg, nbrs=3)
    return sc, sc.views[1], sc.views[2]


@pytest.mark.parametrize("knn", [10, 2, -1])
def test_match_lines_equals_matchingCPU(api, oracle, scene_mod, knn):
    """l3d_match_lines = the drop-in for L3DPP::match_lines_GPU: Line3D::matchingCPU's list."""
    sc, va, vb = _two_view_scene(scene_mod)
    o = oracle.OracleLine3D(sc.max_image_width, False)
    o.load_scene(sc)
    F, Ms, Mt, Cs, Ct = o.match_only(va.cam_id, vb.cam_id, 0.25, knn)
    off_o, rec_o = o.lists(va.cam_id, 1)          # matches_[src] right after matchingCPU
    ctx = api.Context()
    got, off_g = ctx.match_lines(va.segs, vb.segs, F, Ms, Mt, Cs, Ct, va.cam_id, vb.cam_id, 0.25, knn,
                                 sc.max_image_width)
    assert (off_g == off_o).all()
    assert len(got) == len(rec_o) > 0
    assert (got["tgt_seg"] == rec_o["tgt_seg"]).all() and (got["tgt_cam"] == vb.cam_id).all()
    rows = np.repeat(np.arange(len(off_o) - 1), np.diff(off_o.astype(np.int64)))
    assert (got["src_seg"] == rows).all()
    for a, b in (("overlap", "overlap"), ("d_p1", "d_p1"), ("d_p2", "d_p2"), ("d_q1", "d_q1"), ("d_q2", "d_q2")):
        assert (got[a].view(np.uint32) == rec_o[b].view(np.uint32)).all(), a
    # the FP32 pre-filter must not change anything
    got2, _ = ctx.match_lines(va.segs, vb.segs, F, Ms, Mt, Cs, Ct, va.cam_id, vb.cam_id, 0.25, knn,
                              sc.max_image_width, filter_mode=1)
    assert got2.tobytes() == got.tobytes()


@pytest.mark.parametrize("eps", [1e-3, 0.05])
def test_match_lines_with_a_full_rank_F(api, oracle, scene_mod, eps):
    """A matrix that is NOT a fundamental matrix (rank 3: its "epipolar lines" do not meet in one point) is legal
    input to match_lines_GPU.  K1's wedge test assumes concurrent lines, so it must switch itself off for such a
    pair: the pre-filtered run and the all-pairs-exact run agree."""
    sc, va, vb = _two_view_scene(scene_mod, n_seg=400)
    o = oracle.OracleLine3D(sc.max_image_width, False)
    o.load_scene(sc)
    F, Ms, Mt, Cs, Ct = o.match_only(va.cam_id, vb.cam_id, 0.25, 10)
    rng = np.random.default_rng(5)
    F3 = np.asarray(F, dtype=np.float64).reshape(3, 3) + eps * np.abs(F).max() * rng.standard_normal((3, 3))
    assert abs(np.linalg.det(F3)) > 1e-9 * np.abs(F3).max() ** 3
    ctx = api.Context()
    a, off_a = ctx.match_lines(va.segs, vb.segs, F3, Ms, Mt, Cs, Ct, va.cam_id, vb.cam_id, 0.25, 10, sc.max_image_width)
    b, off_b = ctx.match_lines(va.segs, vb.segs, F3, Ms, Mt, Cs, Ct, va.cam_id, vb.cam_id, 0.25, 10, sc.max_image_width,
                               filter_mode=1)
    assert (off_a == off_b).all() and a.tobytes() == b.tobytes()


def test_match_lines_empty_and_capacity(api, scene_mod):
    sc, va, vb = _two_view_scene(scene_mod)
    ctx = api.Context()
    I = np.eye(3)
    out, off = ctx.match_lines(np.zeros((0, 4), np.float32), vb.segs, I, I, I, np.zeros(3), np.ones(3), 1, 2, 0.25, 10, 640)
    assert len(out) == 0 and len(off) == 1
    with pytest.raises(api.L3DError):
        ctx.match_lines(va.segs, vb.segs, I, I, I, np.zeros(3), np.ones(3), 7, 7, 0.25, 10, 640)  # same camera id


def test_score_matches_equals_scoringCPU(api, oracle, scene_mod):
    """l3d_score_matches = the drop-in for L3DPP::score_matches_GPU, on buffers packed the way
    Line3D::scoringGPU packs them (sorted by target camera/segment, src/line3D.cc:1579-1623)."""
    sc = scene_mod.make_scene("tiny", seed=11)
    orc = oracle.run_scene(sc)
    v = sc.views[3]
    off, rec = orc.lists(v.cam_id, 0)
    info = orc.view_info(v.cam_id)
    n = len(v.segs)
    matches, ranges, regs = [], np.full((n, 2), -1, np.int32), []
    rng = np.random.default_rng(0)
    for i in range(n):
        r = rec[off[i]:off[i + 1]]
        if len(r) == 0:
            continue
        order = np.lexsort((r["tgt_seg"], r["tgt_cam"]))       # sortMatchesByIDs
        ranges[i] = (len(matches), len(matches) + len(r) - 1)
        for e in r[order]:
            matches.append((float(i), float(e["tgt_cam"]), e["d_p1"], e["d_p2"]))
            regs.append((abs(rng.normal(0.03, 0.01)), abs(rng.normal(0.03, 0.01))))
    matches = np.asarray(matches, np.float32)
    regs = np.asarray(regs, np.float32)
    # any consistent camera works for the entry point: use K^-1 with R = I, C = 0
    RtKinv = np.linalg.inv(v.K)
    Cc = np.array([0.1, -0.2, 0.3])
    exp = oracle.score_packed(v.segs, matches, ranges, regs, RtKinv, Cc, 200.0, float(info["k"]), 0.5)
    got = api.Context().score_matches(v.segs, matches, ranges, regs, RtKinv, Cc, 200.0, float(info["k"]), 0.5)
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()
    assert (got > 0.75).sum() > 0
    # a lower truncation threshold disables the early exits: still identical
    exp2 = oracle.score_packed(v.segs, matches, ranges, regs, RtKinv, Cc, 200.0, float(info["k"]), 0.25)
    got2 = api.Context().score_matches(v.segs, matches, ranges, regs, RtKinv, Cc, 200.0, float(info["k"]), 0.25)
    assert (got2.view(np.uint32) == exp2.view(np.uint32)).all()


def test_ragged_views_and_degenerate_segments(api, oracle, scene_mod):
    """Different segment counts per view, zero-length and out-of-image segments, a neighbour id that
    does not exist."""
    sc = scene_mod.make_scene("tiny", seed=21, n_views=6, n_seg=150, nbrs=3)
    for i, v in enumerate(sc.views):
        s = v.segs[:150 - 17 * i].copy()
        s[5] = (100.0, 100.0, 100.0, 100.0)            # zero length
        s[6] = (-50.0, 20.0, 700.0, 500.0)              # leaves the image
        s[7, 2:] = s[7, :2] + 0.25                      # shorter than one pixel
        v.segs = s
    sc.views[2].neighbors = sc.views[2].neighbors + [999]   # unknown camera: ignored (src/line3D.cc:612)
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, keep_scored=True)
    compare_full(l3, orc, sc)


def test_no_matches_at_all(api, oracle, scene_mod):
    """A bounds window of 1 px rejects every pair: empty lists, no hypotheses, no edges."""
    sc = scene_mod.make_scene("tiny", seed=5, n_views=4, n_seg=60, nbrs=2)
    sc.max_image_width = 1
    orc = oracle.run_scene(sc)
    l3 = api.run_scene(sc, keep_scored=True)
    sizes = compare_full(l3, orc, sc)
    assert sizes["entries"] == 0 and sizes["edges"] == 0
    assert l3.counts()["forward_matches"] == 0


def test_error_behaviour_mirrors_addImage(api, scene_mod):
    """Argument checks of Line3D::addImage (src/line3D.cc:123-205) surface as errors, not crashes."""
    import ctypes as C
    sc = scene_mod.make_scene("tiny", seed=1, n_views=3, n_seg=40, nbrs=2)
    v = sc.views[0]
    l3 = api.Line3D("", False, 640)
    L, h = l3.L, l3.h
    assert L.l3d_scene_begin(h) == 0

    def add(cam, w, hh, segs, nb):
        vv = api.View()
        vv.cam_id, vv.width, vv.height, vv.num_segs = cam, w, hh, len(segs)
        vv.K[:] = v.K.ravel().tolist(); vv.R[:] = v.R.ravel().tolist(); vv.t[:] = v.t.tolist()
        nb = np.asarray(nb, np.uint32)
        segs = np.ascontiguousarray(segs, np.float32)
        return L.l3d_scene_add_view(h, C.byref(vv), segs.ctypes.data, nb.ctypes.data, nb.size)

    assert add(0, 320, 240, v.segs, [1]) == -1 and b"too small" in L.l3d_last_error()
    assert add(0, 640, 480, v.segs, [1]) == 0
    assert add(0, 640, 480, v.segs, [1]) == -1 and b"already in use" in L.l3d_last_error()
    assert add(1, 640, 480, v.segs, []) == -1 and b"no visual neighbors" in L.l3d_last_error()
    assert add(2, 640, 480, v.segs[:0], [0]) == -1 and b"no line segments" in L.l3d_last_error()
    # state errors
    assert L.l3d_affinity(h) == -3 and b"has not run" in L.l3d_last_error()
    p = api.Params()
    p.sigma_p, p.sigma_a, p.num_neighbors, p.epipolar_overlap, p.knn, p.max_image_width = 5, 10, 10, 0.25, 10, -1
    assert L.l3d_scene_commit(h) == 0
    assert L.l3d_match_images(h, C.byref(p)) == -1 and b"max_image_width" in L.l3d_last_error()
    # explicit neighbour lists and world-point lists cannot be mixed in one scene
    assert L.l3d_scene_begin(h) == 0
    assert add(0, 640, 480, v.segs, [1]) == 0
    vv = api.View()
    vv.cam_id, vv.width, vv.height, vv.num_segs = 1, 640, 480, len(v.segs)
    vv.K[:] = v.K.ravel().tolist(); vv.R[:] = v.R.ravel().tolist(); vv.t[:] = v.t.tolist()
    wps = np.arange(8, dtype=np.uint32)
    segs = np.ascontiguousarray(v.segs, np.float32)
    assert L.l3d_scene_add_view_wps(h, C.byref(vv), segs.ctypes.data, wps.ctypes.data, wps.size) == -1
    assert b"cannot be mixed" in L.l3d_last_error()
    with pytest.raises(api.L3DError):
        api.Line3D("", False, 640, 3000, False, False)   # use_GPU=False: there is no CPU path


def test_find_collinear_equals_findCollinCPU(api, oracle, scene_mod):
    """l3d_find_collinear (drop-in for find_collinear_segments_GPU, include/cudawrapper.h:84-86) against
    the oracle's View::findCollinCPU on a synthetic view (many truly collinear fragments: projected
    world segments split in two) and with a padded row stride."""
    sc = scene_mod.make_scene("tiny")
    rng = np.random.default_rng(5)
    segs = sc.views[0].segs.copy()
    # split the first 60 segments into two collinear fragments with a gap, jittered by a fraction of a pixel
    a, b = segs[:60, :2], segs[:60, 2:]
    frag1 = np.concatenate([a, a + 0.4 * (b - a)], axis=1)
    frag2 = np.concatenate([a + 0.6 * (b - a), b], axis=1) + rng.normal(0, 0.3, size=(60, 4)).astype(np.float32)
    lines = np.concatenate([frag1, frag2, segs[60:]]).astype(np.float32)[:211]     # odd size on purpose
    ctx = api.Context()
    for t in (0.5, 2.0):
        want = oracle.find_collinear(lines, t)
        got = ctx.find_collinear(lines, t)
        assert (got == want).all()
        assert (ctx.find_collinear(lines, t, row_stride=224) == want).all()
    assert want.sum() > 60


def test_cudawrapper_shim_find_collinear_runs(api, oracle, scene_mod, tmp_path):
    """The shim's L3DPP::find_collinear_segments_GPU (include/cudawrapper.h:84-86), compiled as the
    reference would compile it (stand-in DataArray with its 32-byte row padding) and run on the GPU:
    the byte table it leaves in the DataArray equals View::findCollinCPU's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "3dline-slam_b200")
    base = scene_mod.make_scene("tiny").views[1].segs.astype(np.float32)
    a, b = base[:30, :2], base[:30, 2:]           # 30 segments split into two collinear fragments with a gap
    segs = np.ascontiguousarray(np.concatenate([np.concatenate([a, a + 0.4 * (b - a)], axis=1),
                                                np.concatenate([a + 0.6 * (b - a), b], axis=1),
                                                base[30:47]]).astype(np.float32))     # 77: rows padded to 96 bytes
    assert segs.shape == (77, 4)
    (tmp_path / "segs.bin").write_bytes(segs.tobytes())
    src = tmp_path / "main.cpp"
    src.write_text(r'''
#include "%s/shim/cudawrapper_b200.cpp"
#include <cstdio>
int main(int argc, char** argv) {
    const unsigned n = 77;
    L3DPP::DataArray<float4> lines(n, 1);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(lines.dataCPU(0, 0), sizeof(float4), n, f) != n) return 2;
    fclose(f);
    L3DPP::DataArray<char> C(n, n);
    L3DPP::find_collinear_segments_GPU(&C, &lines, 2.0f);
    FILE* o = fopen(argv[2], "wb");
    for (unsigned r = 0; r < n; ++r) fwrite(C.dataCPU(0, r), 1, n, o);   // C(c, r): column c of row r
    fclose(o);
    return 0;
}
''' % pkg)
    exe = tmp_path / "shim_collin"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-I", os.path.join(root, "include"), str(src), "-L", pkg,
                           "-ll3dpp_b200", "-Wl,-rpath," + pkg, "-o", str(exe)])
    subprocess.check_call([str(exe), str(tmp_path / "segs.bin"), str(tmp_path / "out.bin")])
    got = np.frombuffer((tmp_path / "out.bin").read_bytes(), dtype=np.int8).reshape(77, 77)
    want = oracle.find_collinear(segs, 2.0)
    assert (got == want).all() and want.sum() >= 60


def test_cudawrapper_shim_match_lines_runs(api, oracle, scene_mod, tmp_path):
    """The shim's L3DPP::match_lines_GPU_f64 compiled and run: the std::vector<std::list<Match>> it fills
    (what Line3D::matchingGPU hands to the rest of Line3D) equals Line3D::matchingCPU's lists."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "3dline-slam_b200")
    sc, va, vb = _two_view_scene(scene_mod)
    o = oracle.OracleLine3D(sc.max_image_width, False)
    o.load_scene(sc)
    F, Ms, Mt, Cs, Ct = o.match_only(va.cam_id, vb.cam_id, 0.25, 10)
    off_o, rec_o = o.lists(va.cam_id, 1)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([len(va.segs), len(vb.segs)], np.uint32).tobytes())
        f.write(np.ascontiguousarray(va.segs, np.float32).tobytes())
        f.write(np.ascontiguousarray(vb.segs, np.float32).tobytes())
        for a in (F, Ms, Mt, Cs, Ct):
            f.write(np.ascontiguousarray(a, np.float64).tobytes())
    src = tmp_path / "main.cpp"
    src.write_text(r'''
#include "%s/shim/cudawrapper_b200.cpp"
#include <cstdio>
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    unsigned n[2];
    if (!f || fread(n, 4, 2, f) != 2) return 2;
    L3DPP::DataArray<float4> ls(n[0], 1), lt(n[1], 1);
    if (fread(ls.dataCPU(0, 0), 16, n[0], f) != n[0] || fread(lt.dataCPU(0, 0), 16, n[1], f) != n[1]) return 3;
    double M[9 * 3 + 6];
    if (fread(M, 8, 33, f) != 33) return 4;
    fclose(f);
    std::vector<std::list<L3DPP::Match> > matches(n[0]);
    const unsigned got = L3DPP::match_lines_GPU_f64(&ls, &lt, M, M + 9, M + 18, M + 27, M + 30, &matches, %d, %d, 0.25f, 10, %d);
    FILE* o = fopen(argv[2], "wb");
    fwrite(&got, 4, 1, o);
    for (unsigned r = 0; r < n[0]; ++r)
        for (std::list<L3DPP::Match>::const_iterator it = matches[r].begin(); it != matches[r].end(); ++it) {
            const unsigned ids[4] = {it->src_camID_, it->src_segID_, it->tgt_camID_, it->tgt_segID_};
            const float v[6] = {it->overlap_score_, it->score3D_, it->depth_p1_, it->depth_p2_, it->depth_q1_, it->depth_q2_};
            if (it->src_segID_ != r || it->match_orientation_) return 5;
            fwrite(ids, 4, 4, o);
            fwrite(v, 4, 6, o);
        }
    fclose(o);
    return 0;
}
''' % (pkg, va.cam_id, vb.cam_id, sc.max_image_width))
    exe = tmp_path / "shim_match"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-I", os.path.join(root, "include"), str(src), "-L", pkg,
                           "-ll3dpp_b200", "-Wl,-rpath," + pkg, "-o", str(exe)])
    subprocess.check_call([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    raw = (tmp_path / "out.bin").read_bytes()
    n = int(np.frombuffer(raw[:4], np.uint32)[0])
    rec = np.frombuffer(raw[4:], dtype=np.dtype([("ids", "<u4", 4), ("v", "<f4", 6)]))
    assert n == len(rec) == len(rec_o) > 0
    assert (rec["ids"][:, 0] == va.cam_id).all() and (rec["ids"][:, 2] == vb.cam_id).all()
    rows = np.repeat(np.arange(len(off_o) - 1), np.diff(off_o.astype(np.int64)))
    assert (rec["ids"][:, 1] == rows).all() and (rec["ids"][:, 3] == rec_o["tgt_seg"]).all()
    for col, name in ((0, "overlap"), (2, "d_p1"), (3, "d_p2"), (4, "d_q1"), (5, "d_q2")):
        assert (rec["v"][:, col].view(np.uint32) == rec_o[name].view(np.uint32)).all(), name
    assert (rec["v"][:, 1] == 0).all()
    o.close()


_MIRROR_MAIN = r'''
#include "line3d_b200.hpp"
#include <cstdio>
#include <cstdlib>
static FILE* f;
template <typename T> static T rd() { T v; if (fread(&v, sizeof(T), 1, f) != 1) exit(9); return v; }
template <typename T> static std::vector<T> rdv(size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) exit(9); return v; }
static std::list<unsigned int> rdl() { unsigned n = rd<unsigned>(); std::vector<unsigned> v = rdv<unsigned>(n); return std::list<unsigned int>(v.begin(), v.end()); }
template <class L3> static void dump(L3& l3, FILE* o) {
    std::list<L3DPP_B200::CLEdge> A;
    std::vector<std::pair<unsigned int, unsigned int> > l2g;
    l3.getAffinityMatrix(A, l2g);
    unsigned n = (unsigned)A.size(), m = (unsigned)l2g.size();
    fwrite(&n, 4, 1, o); fwrite(&m, 4, 1, o);
    for (std::list<L3DPP_B200::CLEdge>::const_iterator it = A.begin(); it != A.end(); ++it) { fwrite(&it->i_, 4, 1, o); fwrite(&it->j_, 4, 1, o); fwrite(&it->w_, 4, 1, o); }
    for (size_t i = 0; i < l2g.size(); ++i) { fwrite(&l2g[i].first, 4, 1, o); fwrite(&l2g[i].second, 4, 1, o); }
}
int main(int argc, char** argv) {
    f = fopen(argv[1], "rb");
    FILE* o = fopen(argv[2], "wb");
    if (!f || !o) return 2;
    const unsigned stream = rd<unsigned>(), by_wps = rd<unsigned>(), width = rd<unsigned>(), n_cycles = rd<unsigned>();
    const float sp = rd<float>(), sa = rd<float>(); const unsigned nn = rd<unsigned>(); const float eo = rd<float>(); const int knn = rd<int>();
    L3DPP_B200::Line3D batch("", false, (int)width, 3000, by_wps != 0, true);
    L3DPP_B200::Line3DStream st("", false, (int)width, 3000, by_wps != 0, true);
    for (unsigned c = 0; c < n_cycles; ++c) {
        if (stream) st.beginCycle();
        unsigned nd = rd<unsigned>();
        for (unsigned i = 0; i < nd; ++i) { unsigned cam = rd<unsigned>(); if (stream && !st.deleteImage(cam)) return 3; }
        unsigned na = rd<unsigned>();
        for (unsigned i = 0; i < na; ++i) {
            unsigned cam = rd<unsigned>(), w = rd<unsigned>(), h = rd<unsigned>();
            std::vector<double> K = rdv<double>(9), R = rdv<double>(9), t = rdv<double>(3);
            float md = rd<float>();
            std::list<unsigned int> l = rdl();
            unsigned ns = rd<unsigned>();
            std::vector<float> segs = rdv<float>(4 * (size_t)ns);
            if (stream) st.addImage(cam, w, h, K.data(), R.data(), t.data(), md, l, segs);
            else batch.addImage(cam, w, h, K.data(), R.data(), t.data(), md, l, segs);
        }
        unsigned nu = rd<unsigned>();
        for (unsigned i = 0; i < nu; ++i) {
            unsigned cam = rd<unsigned>();
            std::vector<double> R = rdv<double>(9), t = rdv<double>(3);
            float md = rd<float>();
            std::list<unsigned int> l = rdl();
            if (stream) st.UpdataImage(cam, R.data(), t.data(), md, l);
            else batch.UpdataImage(cam, R.data(), t.data(), md, l);
        }
        if (stream) { st.matchImages(sp, sa, nn, eo, knn, -1.0f); st.reconstruct3Dlines(); }
        else { batch.matchImages(sp, sa, nn, eo, knn, -1.0f); batch.reconstruct3Dlines(); }
        if (stream) dump(st, o); else dump(batch, o);
    }
    fclose(o);
    return 0;
}
'''


def _write_cycles(path, stream_mode, by_wps, width, params, cycles):
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", int(stream_mode), int(by_wps), width, len(cycles)))
        f.write(struct.pack("<ffIfi", params["sigma_p"], params["sigma_a"], params["num_neighbors"],
                            params["epipolar_overlap"], params["knn"]))
        for dels, adds, upds in cycles:
            f.write(struct.pack("<I", len(dels)) + np.asarray(dels, np.uint32).tobytes())
            f.write(struct.pack("<I", len(adds)))
            for v, lst in adds:
                f.write(struct.pack("<3I", v.cam_id, v.width, v.height))
                for a in (v.K, v.R, v.t):
                    f.write(np.ascontiguousarray(a, np.float64).tobytes())
                f.write(struct.pack("<f", v.median_depth))
                f.write(struct.pack("<I", len(lst)) + np.asarray(lst, np.uint32).tobytes())
                f.write(struct.pack("<I", len(v.segs)) + np.ascontiguousarray(v.segs, np.float32).tobytes())
            f.write(struct.pack("<I", len(upds)))
            for cam, R, t, md, lst in upds:
                f.write(struct.pack("<I", cam) + np.ascontiguousarray(R, np.float64).tobytes() +
                        np.ascontiguousarray(t, np.float64).tobytes() + struct.pack("<f", md))
                f.write(struct.pack("<I", len(lst)) + np.asarray(lst, np.uint32).tobytes())


def _read_dumps(path, n):
    raw = open(path, "rb").read()
    out, p = [], 0
    for _ in range(n):
        ne, nl = np.frombuffer(raw[p:p + 8], np.uint32)
        p += 8
        e = np.frombuffer(raw[p:p + 12 * ne], dtype=np.dtype([("i", "<i4"), ("j", "<i4"), ("w", "<f4")]))
        p += 12 * int(ne)
        l2g = np.frombuffer(raw[p:p + 8 * nl], np.uint32).reshape(-1, 2)
        p += 8 * int(nl)
        out.append((e, l2g))
    return out


def test_cpp_line3d_mirrors_run(api, oracle, scene_mod, tmp_path):
    """The C++ host mirrors (3dline-slam_b200/host/line3d_b200.hpp) compiled and run on the GPU: the batch
    Line3D on the tiny scene and Line3DStream on a short key-frame stream give the oracle's A_ and id maps."""
    import os
    import subprocess
    import stream_utils
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "3dline-slam_b200")
    src = tmp_path / "main.cpp"
    src.write_text(_MIRROR_MAIN)
    exe = tmp_path / "mirror"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-I", os.path.join(root, "include"), "-I", os.path.join(pkg, "host"),
                           str(src), "-L", pkg, "-ll3dpp_b200", "-Wl,-rpath," + pkg, "-o", str(exe)])
    # batch
    sc = scene_mod.make_scene("tiny")
    _write_cycles(tmp_path / "b.bin", False, False, sc.max_image_width, sc.params,
                  [([], [(v, v.neighbors) for v in sc.views], [])])
    subprocess.check_call([str(exe), str(tmp_path / "b.bin"), str(tmp_path / "b.out")])
    (e, l2g), = _read_dumps(tmp_path / "b.out", 1)
    orc = oracle.run_scene(sc)
    oij, ow = orc.edges()
    assert len(e) == len(ow) > 100 and (e["i"] == oij[:, 0]).all() and (e["j"] == oij[:, 1]).all()
    assert (e["w"].view(np.uint32) == ow.view(np.uint32)).all() and (l2g == orc.local2global()).all()
    orc.close()
    # key-frame stream
    st = scene_mod.make_stream(n_keyframes=9, n_seg=250, window=6, nbrs=4, jitter=0.2, n_world=700)
    cyc = [(c.deletes, [(v, v.worldpoints) for v in c.adds], c.updates) for c in st.cycles]
    _write_cycles(tmp_path / "s.bin", True, True, st.max_image_width, st.params, cyc)
    subprocess.check_call([str(exe), str(tmp_path / "s.bin"), str(tmp_path / "s.out")])
    dumps = _read_dumps(tmp_path / "s.out", len(cyc))
    o, calls = stream_utils.oracle_driver(oracle, st)
    k = [0]

    def on_cycle(ci, cy):
        e, l2g = dumps[ci]
        oij, ow = o.edges()
        assert len(e) == len(ow), ci
        assert (e["i"] == oij[:, 0]).all() and (e["j"] == oij[:, 1]).all() and (e["w"].view(np.uint32) == ow.view(np.uint32)).all()
        assert (l2g == o.local2global()).all()
        k[0] += len(ow)
    scene_mod.drive_stream(st, on_cycle=on_cycle, **calls)
    assert k[0] > 200
    o.close()


def test_cudawrapper_shim_score_matches_runs(api, oracle, scene_mod, tmp_path):
    """The shim's L3DPP::score_matches_GPU (include/cudawrapper.h:74-81) compiled and run on buffers packed
    the way Line3D::scoringGPU packs them; the float RtKinv / C of the reference signature are promoted to
    double by the shim, so the oracle gets the same float-rounded camera."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "3dline-slam_b200")
    sc = scene_mod.make_scene("tiny", seed=11)
    orc = oracle.run_scene(sc)
    v = sc.views[3]
    off, rec = orc.lists(v.cam_id, 0)
    info = orc.view_info(v.cam_id)
    n = len(v.segs)
    matches, ranges, regs = [], np.full((n, 2), -1, np.int32), []
    rng = np.random.default_rng(0)
    for i in range(n):
        r = rec[off[i]:off[i + 1]]
        if len(r) == 0:
            continue
        order = np.lexsort((r["tgt_seg"], r["tgt_cam"]))
        ranges[i] = (len(matches), len(matches) + len(r) - 1)
        for e in r[order]:
            matches.append((float(i), float(e["tgt_cam"]), e["d_p1"], e["d_p2"]))
            regs.append((abs(rng.normal(0.03, 0.01)), abs(rng.normal(0.03, 0.01))))
    matches = np.asarray(matches, np.float32)
    regs = np.asarray(regs, np.float32)
    M32 = np.linalg.inv(v.K).astype(np.float32)
    C32 = np.array([0.1, -0.2, 0.3], np.float32)
    k = np.float32(info["k"])
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n, len(matches)], np.uint32).tobytes())
        f.write(np.ascontiguousarray(v.segs, np.float32).tobytes() + matches.tobytes() + ranges.tobytes() + regs.tobytes())
        f.write(M32.tobytes() + C32.tobytes() + np.array([200.0, k, 0.5], np.float32).tobytes())
    src = tmp_path / "main.cpp"
    src.write_text(r'''
#include "%s/shim/cudawrapper_b200.cpp"
#include <cstdio>
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    unsigned n[2];
    if (!f || fread(n, 4, 2, f) != 2) return 2;
    L3DPP::DataArray<float4> lines(n[0], 1), matches(n[1], 1);
    L3DPP::DataArray<int2> ranges(n[0], 1);
    L3DPP::DataArray<float2> regs(n[1], 1);
    L3DPP::DataArray<float> scores(n[1], 1), M(3, 3);
    if (fread(lines.dataCPU(0, 0), 16, n[0], f) != n[0] || fread(matches.dataCPU(0, 0), 16, n[1], f) != n[1] ||
        fread(ranges.dataCPU(0, 0), 8, n[0], f) != n[0] || fread(regs.dataCPU(0, 0), 8, n[1], f) != n[1]) return 3;
    float m[9], c[3], p[3];
    if (fread(m, 4, 9, f) != 9 || fread(c, 4, 3, f) != 3 || fread(p, 4, 3, f) != 3) return 4;
    for (unsigned r = 0; r < 3; ++r)
        for (unsigned col = 0; col < 3; ++col) M.dataCPU(col, r)[0] = m[3 * r + col];   // (x = col, y = row)
    float3 C; C.x = c[0]; C.y = c[1]; C.z = c[2];
    L3DPP::score_matches_GPU(&lines, &matches, &ranges, &scores, &regs, &M, C, p[0], p[1], p[2]);
    FILE* o = fopen(argv[2], "wb");
    fwrite(scores.dataCPU(0, 0), 4, n[1], o);
    fclose(o);
    return 0;
}
''' % pkg)
    exe = tmp_path / "shim_score"
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-I", os.path.join(root, "include"), str(src), "-L", pkg,
                           "-ll3dpp_b200", "-Wl,-rpath," + pkg, "-o", str(exe)])
    subprocess.check_call([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")])
    got = np.frombuffer((tmp_path / "out.bin").read_bytes(), np.float32)
    exp = oracle.score_packed(v.segs, matches, ranges, regs, M32.astype(np.float64), C32.astype(np.float64), 200.0,
                              float(k), 0.5)
    assert len(got) == len(exp) and (got.view(np.uint32) == exp.view(np.uint32)).all() and (got > 0.75).sum() > 0
    orc.close()
