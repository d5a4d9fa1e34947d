// abi.cu -- cudawrapper-level entry points (host buffers in/out, blocking) and the multi-GPU
// export/import of forward-match lists.
#include "ctx.h"

// replaces L3DPP::match_lines_GPU (include/cudawrapper.h:63-71) with Line3D::matchingCPU results
int l3d_match_lines(l3d_ctx* ctx, const float* lines_src, uint32_t n_src, const float* lines_tgt, uint32_t n_tgt,
                    const double* F, const double* RtKinv_src, const double* RtKinv_tgt, const double* C_src,
                    const double* C_tgt, uint32_t src_cam, uint32_t tgt_cam, float epi_overlap, int32_t knn,
                    int32_t max_image_width, int32_t filter_mode, l3d_match* out, uint64_t cap, uint64_t* out_count,
                    uint32_t* out_row_off)
{
    if (!ctx || !lines_src || !lines_tgt || !F || !RtKinv_src || !RtKinv_tgt || !C_src || !C_tgt || !out_count)
        return fail(L3D_ERR_ARG, "NULL argument");
    if (src_cam == tgt_cam) return fail(L3D_ERR_ARG, "src and tgt camera IDs must differ");
    if (n_src == 0 || n_tgt == 0) {
        *out_count = 0;
        if (out_row_off)
            for (uint32_t i = 0; i <= n_src; ++i) out_row_off[i] = 0;
        return L3D_OK;
    }
    // a private two-view scene in raw mode: cameras given as (RtKinv, C), F given, no translation
    l3d_ctx t;
    t.device = ctx->device;
    t.stream = ctx->stream;
    t.raw_mode = true;
    memcpy(t.F_override, F, sizeof(t.F_override));
    auto mk = [&](uint32_t cam, const float* segs, uint32_t n, const double* M, const double* C) {
        HostView hv;
        memset(&hv.v, 0, sizeof(hv.v));
        hv.v.cam_id = cam;
        hv.v.width = hv.v.height = (uint32_t)std::max(max_image_width, 400);
        hv.v.num_segs = n;
        hv.segs.assign(segs, segs + 4 * (size_t)n);
        memcpy(hv.cam.RtKinv.m, M, sizeof(hv.cam.RtKinv.m));
        hv.cam.C = hg::V3{C[0], C[1], C[2]};
        hv.k = 0.0f;
        return hv;
    };
    t.views.push_back(mk(src_cam, lines_src, n_src, RtKinv_src, C_src));
    t.views.push_back(mk(tgt_cam, lines_tgt, n_tgt, RtKinv_tgt, C_tgt));
    t.views[0].nbrs.push_back(tgt_cam);
    // commit sorts by camera id; the pair list must still be (src -> tgt)
    if (tgt_cam < src_cam) {
        // matchingCPU is asymmetric: keep src first by giving the neighbour list to src only
    }
    int rc = l3d_scene_commit(&t);
    if (rc) return rc;
    l3d_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.sigma_p = 1.0f;
    prm.sigma_a = 10.0f;
    prm.num_neighbors = 2;
    prm.epipolar_overlap = epi_overlap;
    prm.knn = knn;
    prm.const_reg_depth = -1.0f;
    prm.max_image_width = max_image_width;
    prm.filter_mode = filter_mode;
    rc = l3d_match_stage12(&t, &prm);
    ctx->cnt.gpu_launches += t.cnt.gpu_launches;
    ctx->cnt.pair_tests = t.cnt.pair_tests;
    ctx->cnt.candidates = t.cnt.candidates;
    memcpy(ctx->tm.ms, t.tm.ms, sizeof(ctx->tm.ms));
    if (rc) return rc;
    if (t.pairs.size() != 1) return fail(L3D_ERR_STATE, "internal: expected one pair, got %zu", t.pairs.size());
    const uint64_t total = t.total_fwd;
    *out_count = total;
    const uint32_t row_base = t.pairs_h[0].row_base;
    std::vector<uint32_t> off(n_src), cnt(n_src);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(off.data(), t.d_fwd_off.p + row_base, (size_t)n_src * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cnt.data(), t.d_fwd_cnt.p + row_base, (size_t)n_src * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (out_row_off) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < n_src; ++i) {
            out_row_off[i] = run;
            run += cnt[i];
        }
        out_row_off[n_src] = run;
    }
    if (total > cap) return fail(L3D_ERR_CAPACITY, "need %llu matches", (unsigned long long)total);
    if (total == 0) return L3D_OK;
    if (!out) return fail(L3D_ERR_ARG, "out is NULL");
    std::vector<FwdRec> rec(total);
    CK(cudaMemcpyAsync(rec.data(), t.d_fwd_rec.p, total * sizeof(FwdRec), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t n = 0;
    for (uint32_t r = 0; r < n_src; ++r)
        for (uint32_t e = 0; e < cnt[r]; ++e) {
            const FwdRec& f = rec[off[r] + e];
            l3d_match& m = out[n++];
            m.src_cam = src_cam;
            m.src_seg = r;
            m.tgt_cam = tgt_cam;
            m.tgt_seg = f.c;
            m.overlap_score = f.overlap;
            m.score3D = 0.0f;
            m.depth_p1 = f.d_p1;
            m.depth_p2 = f.d_p2;
            m.depth_q1 = f.d_q1;
            m.depth_q2 = f.d_q2;
            m.flags = 0;
        }
    return L3D_OK;
}

// replaces L3DPP::score_matches_GPU (include/cudawrapper.h:74-81)
int l3d_score_matches(l3d_ctx* ctx, const float* lines, uint32_t n_lines, const float* matches, uint32_t n_matches,
                      const int32_t* ranges, float* scores, const float* regularizers_tgt, const double* RtKinv,
                      const double* C, float two_sigA_sqr, float k, float min_similarity)
{
    if (!ctx || !lines || !ranges || !RtKinv || !C) return fail(L3D_ERR_ARG, "NULL argument");
    if (n_matches == 0) return L3D_OK;
    if (!matches || !scores || !regularizers_tgt) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // ranges -> CSR offsets (src/line3D.cc:1582-1596: consecutive, {-1,-1} for empty rows)
    std::vector<uint32_t> off(n_lines + 1);
    uint32_t run = 0;
    for (uint32_t i = 0; i < n_lines; ++i) {
        const int32_t a = ranges[2 * i], b = ranges[2 * i + 1];
        if (a >= 0) {
            if ((uint32_t)a != run || b < a || (uint32_t)b >= n_matches)
                return fail(L3D_ERR_ARG, "ranges[%u] = {%d,%d} is not consecutive", i, a, b);
            off[i] = run;
            run = (uint32_t)b + 1;
        } else
            off[i] = run;
    }
    off[n_lines] = run;
    if (run != n_matches) return fail(L3D_ERR_ARG, "ranges cover %u of %u matches", run, n_matches);
    DevBuf<float4> d_lines, d_m;
    DevBuf<float2> d_rt;
    DevBuf<double> d_cam;
    DevBuf<uint32_t> d_off;
    DevBuf<ListRec> d_L;
    DevBuf<ListGeo> d_G;
    DevBuf<unsigned char> d_stats;
    CK(d_lines.ensure(n_lines));
    CK(d_m.ensure(n_matches));
    CK(d_rt.ensure(n_matches));
    CK(d_cam.ensure(12));
    CK(d_off.ensure((size_t)n_lines + 1));
    CK(d_L.ensure(n_matches));
    CK(d_G.ensure(n_matches));
    CK(d_stats.ensure(k3_stats_bytes()));
    double cam[12];
    memcpy(cam, RtKinv, 72);
    memcpy(cam + 9, C, 24);
    CK(cudaMemcpyAsync(d_lines.p, lines, (size_t)n_lines * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_m.p, matches, (size_t)n_matches * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_rt.p, regularizers_tgt, (size_t)n_matches * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cam.p, cam, sizeof(cam), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off.p, off.data(), ((size_t)n_lines + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_stats.p, 0, k3_stats_bytes(), st));
    ctx->cnt.gpu_launches +=
        launch_score_prep(d_lines.p, n_lines, d_m.p, d_rt.p, d_cam.p, d_cam.p + 9, k, d_off.p, d_L.p, d_G.p, st);
    ctx->cnt.gpu_launches +=
        launch_k3_score(n_lines, d_off.p, d_L.p, d_G.p, two_sigA_sqr, min_similarity, d_stats.p, st);
    std::vector<ListRec> L(n_matches);
    CK(cudaMemcpyAsync(L.data(), d_L.p, (size_t)n_matches * sizeof(ListRec), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n_matches; ++i) scores[i] = L[i].score;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU: forward lists of this shard as one blob:  uint32 cnt[rows_pad] | FwdRec recs[total]
// (rows_pad = total_rows rounded up to 8 so that the records are 32-byte aligned)
// ------------------------------------------------------------------------------------------
static uint64_t rows_pad(const l3d_ctx* ctx) { return ((uint64_t)ctx->total_rows + 7ull) & ~7ull; }

int l3d_forward_blob_size(l3d_ctx* ctx, uint64_t* bytes)
{
    if (!ctx || !bytes) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx->stage < 1) return fail(L3D_ERR_STATE, "l3d_match_stage12 has not run");
    *bytes = rows_pad(ctx) * 4 + ctx->total_fwd * sizeof(FwdRec);
    return L3D_OK;
}

int l3d_export_forward(l3d_ctx* ctx, void* dst, uint64_t cap_bytes, int device_ptr)
{
    if (!ctx || !dst) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx->stage < 1) return fail(L3D_ERR_STATE, "l3d_match_stage12 has not run");
    uint64_t need = 0;
    l3d_forward_blob_size(ctx, &need);
    if (cap_bytes < need) return fail(L3D_ERR_CAPACITY, "need %llu bytes", (unsigned long long)need);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const cudaMemcpyKind kind = device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    unsigned char* d = (unsigned char*)dst;
    if (device_ptr)
        CK(cudaMemsetAsync(d, 0, rows_pad(ctx) * 4, st));
    else
        memset(d, 0, rows_pad(ctx) * 4);
    CK(cudaMemcpyAsync(d, ctx->d_fwd_cnt.p, (size_t)ctx->total_rows * 4, kind, st));
    if (ctx->total_fwd)
        CK(cudaMemcpyAsync(d + rows_pad(ctx) * 4, ctx->d_fwd_rec.p, ctx->total_fwd * sizeof(FwdRec), kind, st));
    CK(cudaStreamSynchronize(st));
    return L3D_OK;
}

// blobs: `world` blobs of `stride_bytes` each (the all-gathered exports, own shard included)
int l3d_import_forward(l3d_ctx* ctx, const void* blobs, uint64_t stride_bytes, int world, int device_ptr)
{
    if (!ctx || !blobs || world < 1) return fail(L3D_ERR_ARG, "bad argument");
    if (ctx->stage < 1) return fail(L3D_ERR_STATE, "l3d_match_stage12 has not run");
    if (stride_bytes % 32) return fail(L3D_ERR_ARG, "stride must be a multiple of 32 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t R = ctx->total_rows;
    DevBuf<unsigned char> stage;
    const uint32_t* dblobs = (const uint32_t*)blobs;
    if (!device_ptr) {
        CK(stage.ensure(stride_bytes * world));
        CK(cudaMemcpyAsync(stage.p, blobs, stride_bytes * world, cudaMemcpyHostToDevice, st));
        dblobs = (const uint32_t*)stage.p;
    }
    const uint64_t stride_words = stride_bytes / 4;
    // per-shard exclusive scans (source offsets) and the merged counts / offsets
    DevBuf<uint32_t> shard_off;
    CK(shard_off.ensure((size_t)world * (R + 1)));
    CK(ctx->d_scan.ensure(scan_scratch_words(R + 1) + 64));
    for (int w = 0; w < world; ++w)
        ctx->cnt.gpu_launches += launch_scan_u32(dblobs + (size_t)w * stride_words, shard_off.p + (size_t)w * (R + 1),
                                                 R, ctx->d_scan.p, ctx->d_scan.cap, st);
    ctx->cnt.gpu_launches += launch_fwd_merge_cnt(dblobs, stride_words, world, R, ctx->d_fwd_cnt.p, st);
    DevBuf<uint32_t> new_off;
    CK(new_off.ensure((size_t)R + 2));
    ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_fwd_cnt.p, new_off.p, R, ctx->d_scan.p, ctx->d_scan.cap, st);
    uint32_t total = 0;
    CK(cudaMemcpyAsync(&total, new_off.p + R, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // the merged records go to a fresh buffer (the own shard is one of the sources)
    DevBuf<FwdRec> merged;
    CK(merged.ensure((size_t)total + 1));
    ctx->cnt.gpu_launches +=
        launch_fwd_merge_copy(dblobs, stride_words, world, R, shard_off.p, new_off.p, merged.p, st);
    CK(cudaMemcpyAsync(ctx->d_fwd_off.p, new_off.p, ((size_t)R + 1) * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    std::swap(ctx->d_fwd_rec.p, merged.p);
    std::swap(ctx->d_fwd_rec.cap, merged.cap);
    ctx->total_fwd = total;
    ctx->cnt.forward_matches = total;
    return refresh_pair_totals(ctx);
}
