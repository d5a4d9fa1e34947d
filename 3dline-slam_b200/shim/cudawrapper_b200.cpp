// cudawrapper_b200.cpp -- drop-in definitions of the reference's GPU entry points
// (include/cudawrapper.h:63-89) on top of libl3dpp_b200.so.
//
// Build inside the reference tree with -DL3DPP_CUDA (it then sees the reference's own dataArray.h /
// commons.h / sparsematrix.h); without L3DPP_B200_IN_REFERENCE_TREE a minimal stand-in of those
// types is used so that this file can be compile-checked on its own (tests/test_shim_compiles.py).
//
// The reference signatures carry FLOAT matrices (DataArray<float> F, RtKinv, float3 C): the
// upstream GPU path never was bit-compatible with the CPU path.  The float
// entry points below promote to double; match_lines_GPU_f64 takes the doubles Line3D already has
// (Eigen::Matrix3d F, View::RtKinv(), View::C()) and reproduces Line3D::matchingCPU bit for bit --
// INTEGRATION.md shows the three-line change in Line3D::matchingGPU that calls it.
#include <cstdint>
#include <iostream>
#include <list>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "l3dpp_b200.h"

#ifdef L3DPP_B200_IN_REFERENCE_TREE
#include "cudawrapper.h"
#else
// ---- minimal stand-ins (layout-compatible subset of the reference types) ----
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
namespace L3DPP {
template <class T>
class DataArray {  // include/dataArray.h:94-384 (host side only)
  public:
    DataArray(unsigned int w, unsigned int h) : width_(w), height_(h)
    {
        size_t pitch = w * sizeof(T);
        if (pitch % 32) pitch += 32 - pitch % 32;
        stride_ = pitch / sizeof(T);
        data_.resize(stride_ * h);
    }
    T* dataCPU(unsigned int x = 0, unsigned int y = 0) { return (x < width_ && y < height_) ? &data_[y * stride_ + x] : nullptr; }
    unsigned int width() const { return width_; }
    unsigned int height() const { return height_; }
  private:
    unsigned int width_, height_;
    size_t stride_;
    std::vector<T> data_;
};
struct Match {  // include/commons.h:197-219
    unsigned int src_camID_, src_segID_, tgt_camID_, tgt_segID_;
    float overlap_score_, score3D_, depth_p1_, depth_p2_, depth_q1_, depth_q2_;
    bool match_orientation_, match_valid_;
};
class SparseMatrix;
}  // namespace L3DPP
#endif

namespace L3DPP {

// One l3d_ctx per Line3D object (the contract of include/l3dpp_b200.h: a context is not thread-safe and
// belongs to one Line3D).  The Line3D constructor calls l3dpp_b200_attach(this, max_image_width_) and the
// destructor l3dpp_b200_detach(this); the entry points below act on the context of the
// object attached last on the calling thread.  Without any attach call they fall back to one process-wide
// context -- the behaviour of the upstream global staging buffer -- whose image width must then be given
// with l3dpp_b200_set_max_image_width before the float-signature match_lines_GPU is used.
struct ShimState {
    l3d_ctx* ctx = nullptr;
    int max_image_width = -1;  // Line3D::max_image_width_; -1 = not told yet
};
static std::mutex g_shim_mutex;
static std::map<const void*, ShimState> g_shim_states;
static thread_local const void* g_shim_current = nullptr;

static ShimState* shim_state()
{
    std::lock_guard<std::mutex> lock(g_shim_mutex);
    ShimState& st = g_shim_states[g_shim_current];
    if (!st.ctx && l3d_ctx_create(&st.ctx, -1) != L3D_OK) {
        std::cerr << "l3dpp_b200: " << l3d_last_error() << std::endl;
        st.ctx = nullptr;
        return nullptr;
    }
    return &st;
}
static l3d_ctx* shim_ctx()
{
    ShimState* st = shim_state();
    return st ? st->ctx : nullptr;
}

}  // namespace L3DPP
extern "C" {
void l3dpp_b200_attach(const void* line3d, int max_image_width)
{
    L3DPP::g_shim_current = line3d;
    L3DPP::ShimState* st = L3DPP::shim_state();
    if (st) st->max_image_width = max_image_width;
}
void l3dpp_b200_use(const void* line3d) { L3DPP::g_shim_current = line3d; }
void l3dpp_b200_detach(const void* line3d)
{
    std::lock_guard<std::mutex> lock(L3DPP::g_shim_mutex);
    auto f = L3DPP::g_shim_states.find(line3d);
    if (f != L3DPP::g_shim_states.end()) {
        if (f->second.ctx) l3d_ctx_destroy(f->second.ctx);
        L3DPP::g_shim_states.erase(f);
    }
    if (L3DPP::g_shim_current == line3d) L3DPP::g_shim_current = nullptr;
}
void l3dpp_b200_set_max_image_width(int max_image_width)
{
    L3DPP::ShimState* st = L3DPP::shim_state();
    if (st) st->max_image_width = max_image_width;
}
}
namespace L3DPP {

// 3x3 DataArray<float> is indexed (x=col, y=row): src/line3D.cc:3266-3272, src/view.cc:39-42
static void mat_from_da(DataArray<float>* M, double out[9])
{
    for (unsigned r = 0; r < 3; ++r)
        for (unsigned c = 0; c < 3; ++c) out[3 * r + c] = (double)M->dataCPU(c, r)[0];
}

unsigned int match_lines_GPU_f64(DataArray<float4>* lines_src, DataArray<float4>* lines_tgt, const double F[9],
                                 const double RtKinv_src[9], const double RtKinv_tgt[9], const double C_src[3],
                                 const double C_tgt[3], std::vector<std::list<Match> >* matches,
                                 const unsigned int srcCamID, const unsigned int tgtCamID, const float epi_overlap,
                                 const int kNN, const int max_image_width)
{
    l3d_ctx* ctx = shim_ctx();
    if (!ctx || !lines_src || !lines_tgt || !matches) return 0;
    const uint32_t ns = lines_src->width(), nt = lines_tgt->width();
    // rows are contiguous float4 (x1,y1,x2,y2): src/line3D.cc:188-196
    const float* ps = (const float*)lines_src->dataCPU(0, 0);
    const float* pt = (const float*)lines_tgt->dataCPU(0, 0);
    uint64_t cap = (uint64_t)ns * (kNN > 0 ? (uint64_t)kNN : 32), n = 0;
    std::vector<l3d_match> out(cap ? cap : 1);
    int rc = l3d_match_lines(ctx, ps, ns, pt, nt, F, RtKinv_src, RtKinv_tgt, C_src, C_tgt, srcCamID, tgtCamID,
                             epi_overlap, kNN, max_image_width, 0, out.data(), cap, &n, nullptr);
    if (rc == L3D_ERR_CAPACITY) {
        out.resize(n);
        cap = n;
        rc = l3d_match_lines(ctx, ps, ns, pt, nt, F, RtKinv_src, RtKinv_tgt, C_src, C_tgt, srcCamID, tgtCamID,
                             epi_overlap, kNN, max_image_width, 0, out.data(), cap, &n, nullptr);
    }
    if (rc != L3D_OK) {
        std::cerr << "match_lines_GPU: " << l3d_last_error() << std::endl;  // print and continue, like the reference
        return 0;
    }
    for (uint64_t i = 0; i < n; ++i) {
        const l3d_match& m = out[i];
        Match M;
        M.src_camID_ = m.src_cam; M.src_segID_ = m.src_seg; M.tgt_camID_ = m.tgt_cam; M.tgt_segID_ = m.tgt_seg;
        M.overlap_score_ = m.overlap_score; M.score3D_ = 0.0f;
        M.depth_p1_ = m.depth_p1; M.depth_p2_ = m.depth_p2; M.depth_q1_ = m.depth_q1; M.depth_q2_ = m.depth_q2;
        M.match_orientation_ = false; M.match_valid_ = false;
        if (m.src_seg < matches->size()) (*matches)[m.src_seg].push_back(M);
    }
    return (unsigned int)n;
}

// include/cudawrapper.h:63-71 (float inputs, promoted)
unsigned int match_lines_GPU(DataArray<float4>* lines_src, DataArray<float4>* lines_tgt, DataArray<float>* F,
                             DataArray<float>* RtKinv_src, DataArray<float>* RtKinv_tgt, const float3 C_src,
                             const float3 C_tgt, std::vector<std::list<Match> >* matches,
                             const unsigned int srcCamID, const unsigned int tgtCamID, const float epi_overlap,
                             const int kNN)
{
    double Fd[9], Ms[9], Mt[9];
    mat_from_da(F, Fd);
    mat_from_da(RtKinv_src, Ms);
    mat_from_da(RtKinv_tgt, Mt);
    const double Cs[3] = {C_src.x, C_src.y, C_src.z}, Ct[3] = {C_tgt.x, C_tgt.y, C_tgt.z};
    // the reference signature has no image width, but Line3D::matchingCPU's bounds test needs
    // Line3D::max_image_width_ (src/line3D.cc:1142-1148): it comes from l3dpp_b200_attach /
    // l3dpp_b200_set_max_image_width.  Guessing it would silently drop matches, so refuse instead.
    ShimState* st = shim_state();
    if (!st || st->max_image_width <= 0) {
        std::cerr << "match_lines_GPU: Line3D::max_image_width_ is unknown -- call l3dpp_b200_attach(this, max_image_width_) "
                     "in the Line3D constructor (or l3dpp_b200_set_max_image_width), or use match_lines_GPU_f64" << std::endl;
        return 0;
    }
    return match_lines_GPU_f64(lines_src, lines_tgt, Fd, Ms, Mt, Cs, Ct, matches, srcCamID, tgtCamID, epi_overlap, kNN,
                               st->max_image_width);
}

// include/cudawrapper.h:74-81
void score_matches_GPU(DataArray<float4>* lines, DataArray<float4>* matches, DataArray<int2>* ranges,
                       DataArray<float>* scores, DataArray<float2>* regularizers_tgt, DataArray<float>* RtKinv,
                       const float3 C, const float two_sigA_sqr, const float k, const float min_similarity)
{
    l3d_ctx* ctx = shim_ctx();
    if (!ctx || !lines || !matches || !ranges || !scores || !regularizers_tgt || !RtKinv) return;
    double M[9];
    mat_from_da(RtKinv, M);
    const double Cd[3] = {C.x, C.y, C.z};
    const int rc = l3d_score_matches(ctx, (const float*)lines->dataCPU(0, 0), lines->width(),
                                     (const float*)matches->dataCPU(0, 0), matches->width(),
                                     (const int32_t*)ranges->dataCPU(0, 0), scores->dataCPU(0, 0),
                                     (const float*)regularizers_tgt->dataCPU(0, 0), M, Cd, two_sigA_sqr, k,
                                     min_similarity);
    if (rc != L3D_OK) std::cerr << "score_matches_GPU: " << l3d_last_error() << std::endl;
}

// include/cudawrapper.h:84-86, as View::findCollinGPU calls it (src/view.cc:203-236): C is N x N,
// C(c, r) = 1 iff segment c is collinear to segment r (View::findCollinCPU, src/view.cc:238-293)
void find_collinear_segments_GPU(DataArray<char>* C, DataArray<float4>* lines, const float dist_t)
{
    l3d_ctx* ctx = shim_ctx();
    if (!ctx || !C || !lines || lines->width() == 0) return;
    const unsigned int n = lines->width();
    if (C->width() < n || C->height() < n) {
        std::cerr << "find_collinear_segments_GPU: buffer smaller than " << n << " x " << n << std::endl;
        return;
    }
    const uint64_t stride = n > 1 ? (uint64_t)(C->dataCPU(0, 1) - C->dataCPU(0, 0)) : n;
    if (l3d_find_collinear(ctx, (const float*)lines->dataCPU(0, 0), n, dist_t, C->dataCPU(0, 0), stride) != L3D_OK)
        std::cerr << "l3dpp_b200: " << l3d_last_error() << std::endl;
}
// include/cudawrapper.h:88-89 -- diffusion is off in the reference configuration
// (include/L3DPPing.h:80) and the kernel's source is not part of the reference tree; loud stub.
void replicator_dynamics_diffusion_GPU(SparseMatrix*&, const std::string)
{
    std::cerr << "replicator_dynamics_diffusion_GPU: not part of the l3dpp-b200 path (RDD is off)" << std::endl;
}

}  // namespace L3DPP
