// line3d_b200.hpp -- C++ host-side mirror of L3DPP::Line3D (reference include/line3D.h:71-479)
// for the matching / scoring / affinity path, on top of the C ABI (include/l3dpp_b200.h).
// Same method names, argument meaning and error behaviour (print + return, no exceptions) as the
// reference, with Eigen/OpenCV types replaced by plain arrays.  Header-only; link libl3dpp_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "l3dpp_b200.h"

namespace L3DPP_B200 {

struct CLEdge {  // include/clustering.h:57-61
    int i_, j_;
    float w_;
};

// A_ as the reference's std::list<CLEdge> (include/clustering.h:57-61) and local2global_
inline void read_affinity_matrix(l3d_ctx* ctx, std::list<CLEdge>& A,
                                 std::vector<std::pair<unsigned int, unsigned int>>& local2global)
{
    A.clear();
    local2global.clear();
    l3d_counts c{};
    if (!ctx || l3d_get_counts(ctx, &c) != L3D_OK || c.num_edges == 0) return;
    std::vector<int32_t> ij(2 * (size_t)c.num_edges);
    std::vector<float> w(c.num_edges);
    std::vector<uint32_t> l2g(2 * (size_t)c.num_local_ids);
    if (l3d_get_edges(ctx, ij.data(), w.data(), c.num_edges) != L3D_OK ||
        l3d_get_local2global(ctx, l2g.data(), c.num_local_ids) != L3D_OK) {
        std::cerr << "[L3D++] ERROR: " << l3d_last_error() << std::endl;
        return;
    }
    for (uint32_t e = 0; e < c.num_edges; ++e) A.push_back(CLEdge{ij[2 * e], ij[2 * e + 1], w[e]});
    for (uint32_t i = 0; i < c.num_local_ids; ++i) local2global.push_back({l2g[2 * i], l2g[2 * i + 1]});
}

class Line3D {
  public:
    // Line3D::Line3D (src/line3D.cc:6-74); load_segments / output folder are I/O options of the
    // reference that this path does not use.  use_GPU must be true: there is no CPU path.
    Line3D(const std::string& /*output_folder*/, bool /*load_segments*/ = false, int max_img_width = -1,
           unsigned int max_line_segments = 3000, bool neighbors_by_worldpoints = false, bool use_GPU = true,
           int device = -1)
        : max_image_width_(max_img_width), max_line_segments_(max_line_segments), by_wps_(neighbors_by_worldpoints)
    {
        prefix_ = "[L3D++] ";
        if (!use_GPU) std::cerr << prefix_ << "ERROR: l3dpp-b200 has no CPU path (use_GPU must be true)" << std::endl;
        if (l3d_ctx_create(&ctx_, device) != L3D_OK) {
            std::cerr << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
            ctx_ = nullptr;
        }
    }
    ~Line3D() { l3d_ctx_destroy(ctx_); }
    Line3D(const Line3D&) = delete;
    Line3D& operator=(const Line3D&) = delete;

    // Line3D::addImage (src/line3D.cc:117-227) with pre-detected segments (x1,y1,x2,y2 per row)
    void addImage(unsigned int camID, unsigned int width, unsigned int height, const double K[9], const double R[9],
                  const double t[3], float median_depth, const std::list<unsigned int>& wps_or_neighbors,
                  const std::vector<float>& line_segments_xyxy)
    {
        if (views_.count(camID)) {
            std::cout << prefix_ << "ERROR: camera ID [" << camID << "] already in use!" << std::endl;
            return;
        }
        if (wps_or_neighbors.empty()) {
            std::cout << prefix_ << "ERROR: view [" << camID << (by_wps_ ? "] has no worldpoints!" : "] has no visual neighbors!")
                      << std::endl;
            return;
        }
        V v;
        v.d.cam_id = camID;
        v.d.width = width;
        v.d.height = height;
        v.d.num_segs = (uint32_t)(line_segments_xyxy.size() / 4);
        for (int i = 0; i < 9; ++i) { v.d.K[i] = K[i]; v.d.R[i] = R[i]; }
        for (int i = 0; i < 3; ++i) v.d.t[i] = t[i];
        v.d.median_depth = median_depth;
        v.segs = line_segments_xyxy;
        v.nbrs.assign(wps_or_neighbors.begin(), wps_or_neighbors.end());
        views_[camID] = v;
        dirty_ = true;
    }

    // Line3D::UpdataImage (src/line3D.cc:433-487)
    void UpdataImage(unsigned int camID, const double R[9], const double t[3], float median_depth,
                     const std::list<unsigned int>& wps_or_neighbors)
    {
        auto f = views_.find(camID);
        if (f == views_.end()) return;
        for (int i = 0; i < 9; ++i) f->second.d.R[i] = R[i];
        for (int i = 0; i < 3; ++i) f->second.d.t[i] = t[i];
        f->second.d.median_depth = median_depth;
        f->second.nbrs.assign(wps_or_neighbors.begin(), wps_or_neighbors.end());
        dirty_ = true;
    }

    // Line3D::matchImages (src/line3D.cc:496-640)
    void matchImages(float sigma_position = 2.5f, float sigma_angle = 10.0f, unsigned int num_neighbors = 10,
                     float epipolar_overlap = 0.25f, int kNN = 10, float const_regularization_depth = -1.0f)
    {
        if (!ctx_) return;
        if (views_.empty()) {
            std::cout << prefix_ << "WARNING: no images to match! forgot to add them?" << std::endl;
            return;
        }
        if (dirty_ && !upload()) return;
        l3d_params p{};
        p.sigma_p = sigma_position;
        p.sigma_a = sigma_angle;
        p.num_neighbors = num_neighbors;
        p.epipolar_overlap = epipolar_overlap;
        p.knn = kNN;
        p.const_reg_depth = const_regularization_depth;
        p.max_image_width = max_image_width_;
        if (l3d_match_images(ctx_, &p) != L3D_OK) std::cerr << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
    }

    // Line3D::reconstruct3Dlines up to and including clusterSegments' clustering call
    // (src/line3D.cc:2018-2118, 2502-2516); A_ and the cluster IDs are returned to the caller, the
    // 3-D line tail (get3DlineFromCluster ...) stays in the reference.
    void reconstruct3Dlines(unsigned int /*visibility_t*/ = 3, bool perform_diffusion = false,
                            float collinearity_t = -1.0f, bool use_CERES = false)
    {
        if (!ctx_) return;
        if (perform_diffusion || use_CERES || collinearity_t > 1e-12f)
            std::cout << prefix_ << "ERROR: diffusion / CERES / collinearity are not part of this path" << std::endl;
        if (l3d_affinity(ctx_) != L3D_OK || l3d_cluster(ctx_) != L3D_OK)
            std::cerr << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
    }

    // A_ as the reference's std::list<CLEdge> (include/clustering.h) and the ID maps
    void getAffinityMatrix(std::list<CLEdge>& A, std::vector<std::pair<unsigned int, unsigned int>>& local2global)
    {
        read_affinity_matrix(ctx_, A, local2global);
    }

    size_t numImages() const { return views_.size(); }
    l3d_ctx* context() { return ctx_; }

  private:
    struct V {
        l3d_view d;
        std::vector<float> segs;
        std::vector<uint32_t> nbrs;
    };
    bool upload()
    {
        if (l3d_scene_begin(ctx_) != L3D_OK) return false;
        for (auto& kv : views_)
            if ((by_wps_ ? l3d_scene_add_view_wps : l3d_scene_add_view)(ctx_, &kv.second.d, kv.second.segs.data(),
                                                                        kv.second.nbrs.data(),
                                                                        (uint32_t)kv.second.nbrs.size()) != L3D_OK) {
                std::cout << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
                return false;
            }
        if (l3d_scene_commit(ctx_) != L3D_OK) {
            std::cerr << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
            return false;
        }
        dirty_ = false;
        return true;
    }
    l3d_ctx* ctx_ = nullptr;
    std::map<unsigned int, V> views_;
    bool dirty_ = true;
    int max_image_width_;
    unsigned int max_line_segments_;
    bool by_wps_ = false;
    std::string prefix_;
};

// The same object driven incrementally, the way L3DPPing::Run (src/L3DPPing.cpp:98-236) drives it:
// per cycle beginCycle(), deleteImage() for the culled key frames, addImage() for the new ones,
// UpdataImage() for every current one, matchImages(), reconstruct3Dlines().  The context keeps
// matched_, processed_, the filtered match lists with their scores and the Add / Delete sets
// (l3d_stream_* in include/l3dpp_b200.h).  Errors are printed, never thrown, as in the reference.
class Line3DStream {
  public:
    Line3DStream(const std::string& /*output_folder*/, bool /*load_segments*/ = false, int max_img_width = -1,
                 unsigned int /*max_line_segments*/ = 3000, bool neighbors_by_worldpoints = true, bool use_GPU = true,
                 int device = -1)
        : max_image_width_(max_img_width)
    {
        prefix_ = "[L3D++] ";
        if (!use_GPU) std::cerr << prefix_ << "ERROR: l3dpp-b200 has no CPU path (use_GPU must be true)" << std::endl;
        if (l3d_ctx_create(&ctx_, device) != L3D_OK || l3d_stream_begin(ctx_, neighbors_by_worldpoints ? 1 : 0) != L3D_OK) {
            std::cerr << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
            l3d_ctx_destroy(ctx_);
            ctx_ = nullptr;
        }
    }
    ~Line3DStream() { l3d_ctx_destroy(ctx_); }
    Line3DStream(const Line3DStream&) = delete;
    Line3DStream& operator=(const Line3DStream&) = delete;

    // the resets at the top of L3DPPing::Run's loop body (src/L3DPPing.cpp:98-103)
    void beginCycle() { report(ctx_ ? l3d_stream_begin_cycle(ctx_) : L3D_OK); }

    void addImage(unsigned int camID, unsigned int width, unsigned int height, const double K[9], const double R[9],
                  const double t[3], float median_depth, const std::list<unsigned int>& wps_or_neighbors,
                  const std::vector<float>& line_segments_xyxy)
    {
        if (!ctx_) return;
        l3d_view v{};
        v.cam_id = camID;
        v.width = width;
        v.height = height;
        v.num_segs = (uint32_t)(line_segments_xyxy.size() / 4);
        for (int i = 0; i < 9; ++i) { v.K[i] = K[i]; v.R[i] = R[i]; }
        for (int i = 0; i < 3; ++i) v.t[i] = t[i];
        v.median_depth = median_depth;
        std::vector<uint32_t> l(wps_or_neighbors.begin(), wps_or_neighbors.end());
        report(l3d_stream_add_image(ctx_, &v, line_segments_xyxy.data(), l.data(), (uint32_t)l.size()));
    }
    bool deleteImage(unsigned int camID)
    {
        if (!ctx_) return false;
        return report(l3d_stream_delete_image(ctx_, camID));
    }
    void UpdataImage(unsigned int camID, const double R[9], const double t[3], float median_depth,
                     const std::list<unsigned int>& wps_or_neighbors)
    {
        if (!ctx_) return;
        std::vector<uint32_t> l(wps_or_neighbors.begin(), wps_or_neighbors.end());
        report(l3d_stream_update_image(ctx_, camID, R, t, median_depth, l.data(), (uint32_t)l.size()));
    }
    void matchImages(float sigma_position = 2.5f, float sigma_angle = 10.0f, unsigned int num_neighbors = 10,
                     float epipolar_overlap = 0.25f, int kNN = 10, float const_regularization_depth = -1.0f)
    {
        if (!ctx_) return;
        l3d_params p{};
        p.sigma_p = sigma_position;
        p.sigma_a = sigma_angle;
        p.num_neighbors = num_neighbors;
        p.epipolar_overlap = epipolar_overlap;
        p.knn = kNN;
        p.const_reg_depth = const_regularization_depth;
        p.max_image_width = max_image_width_;
        report(l3d_match_images(ctx_, &p));
    }
    void reconstruct3Dlines(unsigned int /*visibility_t*/ = 3, bool perform_diffusion = false,
                            float collinearity_t = -1.0f, bool use_CERES = false)
    {
        if (!ctx_) return;
        if (perform_diffusion || use_CERES || collinearity_t > 1e-12f)
            std::cout << prefix_ << "ERROR: diffusion / CERES / collinearity are not part of this path" << std::endl;
        if (report(l3d_affinity(ctx_))) report(l3d_cluster(ctx_));
    }
    void getAffinityMatrix(std::list<CLEdge>& A, std::vector<std::pair<unsigned int, unsigned int>>& local2global)
    {
        read_affinity_matrix(ctx_, A, local2global);
    }
    l3d_ctx* context() { return ctx_; }

  private:
    bool report(int rc)
    {
        if (rc != L3D_OK) std::cout << prefix_ << "ERROR: " << l3d_last_error() << std::endl;
        return rc == L3D_OK;
    }
    l3d_ctx* ctx_ = nullptr;
    int max_image_width_;
    std::string prefix_;
};

}  // namespace L3DPP_B200
