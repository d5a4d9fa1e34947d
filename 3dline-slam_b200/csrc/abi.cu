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
    // a private two-view scene in raw mode: cameras given as (RtKinv, C), F given, no translation.  The
    // scratch context lives inside ctx and keeps its device buffers and pinned read-back area from one
    // call to the next: the shim calls this once per view pair.
    if (!ctx->pair_scratch) {
        ctx->pair_scratch = new l3d_ctx;
        if (cudaMallocHost((void**)&ctx->pair_scratch->rb, 1u << 16) == cudaSuccess) ctx->pair_scratch->rb_cap = 1u << 16;
        else ctx->pair_scratch->rb = nullptr;
    }
    l3d_ctx& t = *ctx->pair_scratch;
    l3d_scene_begin(&t);
    t.device = ctx->device;
    t.stream = ctx->stream;
    t.raw_mode = true;
    t.cnt = l3d_counts{};
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
    // commit sorts by camera id; matchingCPU is asymmetric: only src carries a neighbour list, so the
    // pair list is (src -> tgt) whichever id is smaller
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
// multi-GPU: the four exchange points of a sharded run (one process per GPU; the collective itself
// is torch.distributed / NCCL, see 3dline-slam_b200/sharding.py).  Every rank exports one blob,
// the blobs are all-gathered `stride` bytes apart, every rank imports all of them.
//   FORWARD     u32 cnt[rows_pad8]                                        rows = pair rows of the slice
//               (every rank learns all per-row counts, i.e. the global record numbering).  The RECORDS of a
//               boundary pair -- target view in another slice -- are needed by one rank only, the owner of the
//               target view, so they travel all-to-all right after: l3d_shard_forward_plan / _pack / _unpack.
//   PROGRAMS    u32 nh[rows_pad4] | u32 off[rows_pad4] | uint4 rec[]     rows = segments of the slice
//   HYPOTHESES  ShardHypHdr | u32 filt_cnt[rows_pad4] | u32 filt_off[rows_pad4] | EntryDev e[rows] | ListRec filt[]
//   EDGES       {u32 src, u32 tgt, float w}[]
// ------------------------------------------------------------------------------------------
struct ShardHypHdr {
    unsigned long long sim_evals, scored;
    uint32_t filt_n, num_valid, pad0, pad1;
};
static_assert(sizeof(ShardHypHdr) == 32, "header must be 32 bytes");

namespace l3d {
int launch_hyp_adopt(const void* all, uint64_t stride, int world, const uint32_t* slice_g, const uint32_t* fbase,
                     uint32_t S, EntryDev* entries, uint32_t* filt_cnt, uint32_t* filt_off, ListRec* filt_all,
                     cudaStream_t st);
}

static uint32_t slice_rows(const l3d_ctx* ctx, int kind, int q)
{
    return kind == L3D_X_FORWARD ? ctx->slice_row[q + 1] - ctx->slice_row[q] : ctx->slice_g[q + 1] - ctx->slice_g[q];
}
static uint64_t pad_to(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// fixed part of rank q's blob
static uint64_t blob_fixed_bytes(const l3d_ctx* ctx, int kind, int q)
{
    const uint64_t rows = slice_rows(ctx, kind, q);
    switch (kind) {
    case L3D_X_FORWARD: return pad_to(rows, 8) * 4;
    case L3D_X_PROGRAMS: return 2 * pad_to(rows, 4) * 4;
    case L3D_X_HYPOTHESES: return sizeof(ShardHypHdr) + 2 * pad_to(rows, 4) * 4 + rows * sizeof(EntryDev);
    default: return 0;
    }
}

static int check_phase(l3d_ctx* ctx, int kind)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    switch (kind) {
    case L3D_X_FORWARD:
        if (ctx->stage < 1) return fail(L3D_ERR_STATE, "l3d_match_stage12 has not run");
        break;
    case L3D_X_PROGRAMS:
        if (ctx->stage3_phase < 1) return fail(L3D_ERR_STATE, "l3d_score_build has not run");
        break;
    case L3D_X_HYPOTHESES:
        if (ctx->stage3_phase < 2) return fail(L3D_ERR_STATE, "l3d_score_fold has not run");
        break;
    case L3D_X_EDGES:
        if (ctx->stage4_phase < 1) return fail(L3D_ERR_STATE, "l3d_affinity_edges has not run");
        break;
    default: return fail(L3D_ERR_ARG, "unknown exchange kind %d", kind);
    }
    return L3D_OK;
}

// variable part of this rank's blob (elements); reads the device cursors, so it synchronises
static int blob_var_count(l3d_ctx* ctx, int kind, uint64_t* n)
{
    cudaStream_t st = ctx->stream;
    switch (kind) {
    case L3D_X_FORWARD: *n = 0; return L3D_OK;  // counts only; the records travel all-to-all
    case L3D_X_EDGES: *n = ctx->n_edges_local; return L3D_OK;
    case L3D_X_PROGRAMS:
        for (int attempt = 0;; ++attempt) {
            uint32_t w[2] = {0, 0};  // WfStats::err, WfStats::prog_cursor
            CK(cudaMemcpyAsync(w, ctx->d_stats.p + 24, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (!(w[0] & 4u)) {
                *n = w[1];
                return L3D_OK;
            }
            if (attempt >= 8) return fail(L3D_ERR_CAPACITY, "fold program store overflow");
            int rc = score_rebuild(ctx, w[1]);
            if (rc) return rc;
        }
    case L3D_X_HYPOTHESES: {
        uint32_t w[2] = {0, 0};  // WfStats::filt_cursor, WfStats::err
        CK(cudaMemcpyAsync(w, ctx->d_stats.p + 20, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (w[1] & 2u) return fail(L3D_ERR_CAPACITY, "filtered-match store overflow");
        if (w[1] & 8u) return fail(L3D_ERR_STATE, "internal: scoring dependency wait timed out");
        *n = w[0];
        return L3D_OK;
    }
    }
    return fail(L3D_ERR_ARG, "unknown exchange kind %d", kind);
}

static uint64_t var_elem_bytes(int kind)
{
    switch (kind) {
    case L3D_X_FORWARD: return sizeof(FwdRec);
    case L3D_X_PROGRAMS: return 16;
    case L3D_X_HYPOTHESES: return sizeof(ListRec);
    default: return 12;
    }
}

int l3d_shard_blob_size(l3d_ctx* ctx, int kind, uint64_t* bytes)
{
    int rc = check_phase(ctx, kind);
    if (rc) return rc;
    if (!bytes) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    uint64_t n = 0;
    rc = blob_var_count(ctx, kind, &n);
    if (rc) return rc;
    ctx->xchg_var[kind] = n;
    *bytes = blob_fixed_bytes(ctx, kind, ctx->rank) + n * var_elem_bytes(kind);
    return L3D_OK;
}

// payload of this rank's blob; n_var = elements of the variable part to copy
static int export_payload(l3d_ctx* ctx, int kind, unsigned char* d, uint64_t n, cudaMemcpyKind ck)
{
    cudaStream_t st = ctx->stream;
    const uint32_t rows = slice_rows(ctx, kind, ctx->rank);
    switch (kind) {
    case L3D_X_FORWARD: {
        const uint32_t r0 = ctx->slice_row[ctx->rank];
        if (rows) CK(cudaMemcpyAsync(d, ctx->d_fwd_cnt.p + r0, (size_t)rows * 4, ck, st));
        break;
    }
    case L3D_X_PROGRAMS: {
        const uint32_t g0 = ctx->slice_g[ctx->rank];
        const uint64_t rp = pad_to(rows, 4);
        if (rows) {
            CK(cudaMemcpyAsync(d, ctx->d_prog_nh.p + g0, (size_t)rows * 4, ck, st));
            CK(cudaMemcpyAsync(d + rp * 4, ctx->d_prog_off.p + g0, (size_t)rows * 4, ck, st));
        }
        if (n) CK(cudaMemcpyAsync(d + 2 * rp * 4, ctx->d_prog.p, n * 16, ck, st));
        break;
    }
    case L3D_X_HYPOTHESES: {
        const uint32_t g0 = ctx->slice_g[ctx->rank];
        const uint64_t rp = pad_to(rows, 4);
        // header = the first 32 bytes of WfStats with filt_cursor in the filt_n slot
        CK(cudaMemcpyAsync(d, ctx->d_stats.p, 16, ck, st));
        CK(cudaMemcpyAsync(d + 16, ctx->d_stats.p + 20, 4, ck, st));
        CK(cudaMemcpyAsync(d + 20, ctx->d_stats.p + 16, 4, ck, st));
        unsigned char* p = d + sizeof(ShardHypHdr);
        if (rows) {
            CK(cudaMemcpyAsync(p, ctx->d_filt_cnt.p + g0, (size_t)rows * 4, ck, st));
            CK(cudaMemcpyAsync(p + rp * 4, ctx->d_filt_off.p + g0, (size_t)rows * 4, ck, st));
            CK(cudaMemcpyAsync(p + 2 * rp * 4, ctx->d_entries.p + g0, (size_t)rows * sizeof(EntryDev), ck, st));
        }
        if (n)
            CK(cudaMemcpyAsync(p + 2 * rp * 4 + (size_t)rows * sizeof(EntryDev), ctx->d_filt_rec.p, n * sizeof(ListRec),
                               ck, st));
        break;
    }
    case L3D_X_EDGES:
        if (n) CK(cudaMemcpyAsync(d, ctx->d_edges.p, n * 12, ck, st));
        break;
    }
    return L3D_OK;
}

int l3d_shard_export(l3d_ctx* ctx, int kind, void* dst, uint64_t cap_bytes, int device_ptr)
{
    int rc = check_phase(ctx, kind);
    if (rc) return rc;
    if (!dst) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    const uint64_t n = ctx->xchg_var[kind];
    const uint64_t need = blob_fixed_bytes(ctx, kind, ctx->rank) + n * var_elem_bytes(kind);
    if (cap_bytes < need) return fail(L3D_ERR_CAPACITY, "need %llu bytes", (unsigned long long)need);
    rc = export_payload(ctx, kind, (unsigned char*)dst, n, device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost);
    if (rc) return rc;
    if (!device_ptr) CK(cudaStreamSynchronize(ctx->stream));
    return L3D_OK;
}

// ---- self-describing blobs (device pointers only): {ShardBlobHdr | payload}.  The sender does not
// wait for its cursors: the header is written by the device, the variable part is copied up to what
// the stride holds.  Every receiver reads all headers; a blob that did not fit (or an overflowed
// program store) makes every rank fall back to the size exchange. ----
struct ShardBlobHdr {
    unsigned long long payload_bytes;
    uint32_t flags;  // WfStats::err of the sender
    uint32_t kind;
    unsigned long long pad[2];
};
static_assert(sizeof(ShardBlobHdr) == 32, "blob header must be 32 bytes");

namespace l3d {
int launch_blob_hdr(void* dst, uint64_t fixed_bytes, uint64_t elem_bytes, const uint32_t* n_dev, uint64_t n_imm,
                    const uint32_t* flags_dev, uint32_t kind, cudaStream_t st);
}

int l3d_shard_export_hdr(l3d_ctx* ctx, int kind, void* dst, uint64_t stride_bytes)
{
    int rc = check_phase(ctx, kind);
    if (rc) return rc;
    if (!dst) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    const uint64_t fixed = blob_fixed_bytes(ctx, kind, ctx->rank), eb = var_elem_bytes(kind);
    if (stride_bytes % 32 || stride_bytes < sizeof(ShardBlobHdr))
        return fail(L3D_ERR_CAPACITY, "stride %llu too small", (unsigned long long)stride_bytes);
    // a stride that cannot even hold the fixed part: send the header alone, the receivers fall back
    const bool header_only = stride_bytes < sizeof(ShardBlobHdr) + fixed;
    uint64_t room = header_only ? 0 : (stride_bytes - sizeof(ShardBlobHdr) - fixed) / eb, have = 0;
    const uint32_t* n_dev = nullptr;
    const uint32_t* f_dev = nullptr;
    switch (kind) {
    case L3D_X_FORWARD: have = 0; break;  // counts only
    case L3D_X_EDGES: have = ctx->n_edges_local; break;
    case L3D_X_PROGRAMS:
        have = ctx->prog_cap;
        n_dev = (const uint32_t*)(ctx->d_stats.p + 28);
        f_dev = (const uint32_t*)(ctx->d_stats.p + 24);
        break;
    case L3D_X_HYPOTHESES:
        have = ctx->filt_cap;
        n_dev = (const uint32_t*)(ctx->d_stats.p + 20);
        f_dev = (const uint32_t*)(ctx->d_stats.p + 24);
        break;
    }
    ctx->cnt.gpu_launches += launch_blob_hdr(dst, fixed, eb, n_dev, have, f_dev, (uint32_t)kind, ctx->stream);
    if (header_only) return L3D_OK;
    return export_payload(ctx, kind, (unsigned char*)dst + sizeof(ShardBlobHdr), std::min(room, have),
                          cudaMemcpyDeviceToDevice);
}

// all: `world` blobs `stride_bytes` apart (own blob included); sizes[q] = l3d_shard_blob_size of rank q.
// Device pointers must stay valid until the next exchange of the same kind (PROGRAMS is read in place).
int l3d_shard_import(l3d_ctx* ctx, int kind, const void* all, uint64_t stride_bytes, int world, const uint64_t* sizes,
                     int device_ptr)
{
    int rc = check_phase(ctx, kind);
    if (rc) return rc;
    if (!all || !sizes) return fail(L3D_ERR_ARG, "NULL argument");
    if (world != ctx->world) return fail(L3D_ERR_ARG, "world %d does not match the plan (%d)", world, ctx->world);
    if (stride_bytes % 32) return fail(L3D_ERR_ARG, "stride must be a multiple of 32 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const unsigned char* src = (const unsigned char*)all;
    if (!device_ptr) {  // host blobs (tests): stage them on the device, one buffer per kind
        DevBuf<unsigned char>& stg = ctx->d_xchg_stage[kind];
        CK(stg.ensure(stride_bytes * world));
        CK(cudaMemcpyAsync(stg.p, all, stride_bytes * world, cudaMemcpyHostToDevice, st));
        src = stg.p;
    }
    std::vector<uint64_t> nvar(world);
    for (int q = 0; q < world; ++q) {
        const uint64_t fx = blob_fixed_bytes(ctx, kind, q);
        if (sizes[q] < fx || sizes[q] > stride_bytes || (sizes[q] - fx) % var_elem_bytes(kind))
            return fail(L3D_ERR_ARG, "blob %d has an impossible size (%llu bytes)", q, (unsigned long long)sizes[q]);
        nvar[q] = (sizes[q] - fx) / var_elem_bytes(kind);
    }
    const uint32_t S = ctx->S;
    switch (kind) {
    case L3D_X_FORWARD: {
        const uint32_t R = ctx->total_rows, P = (uint32_t)ctx->pairs.size();
        const uint32_t r0 = ctx->slice_row[ctx->rank], r1 = ctx->slice_row[ctx->rank + 1];
        // the local offsets of this rank's rows, before the canonical offsets replace them
        CK(ctx->d_fwd_off_local.ensure((size_t)R + 1));
        if (r1 > r0)
            CK(cudaMemcpyAsync(ctx->d_fwd_off_local.p + r0, ctx->d_fwd_off.p + r0, (size_t)(r1 - r0) * 4,
                               cudaMemcpyDeviceToDevice, st));
        // per-row counts of every slice (contiguous in pair rows and in rank order) -> canonical offsets
        for (int q = 0; q < world; ++q) {
            const uint32_t rows = slice_rows(ctx, kind, q);
            if (rows)
                CK(cudaMemcpyAsync(ctx->d_fwd_cnt.p + ctx->slice_row[q], src + (uint64_t)q * stride_bytes, (size_t)rows * 4,
                                   cudaMemcpyDeviceToDevice, st));
        }
        CK(ctx->d_scan.ensure(scan_scratch_words(R + 1) + 64));
        ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_fwd_cnt.p, ctx->d_fwd_off.p, R, ctx->d_scan.p, ctx->d_scan.cap, st);
        rc = refresh_pair_totals(ctx);  // synchronises: record totals per pair and overall
        if (rc) return rc;
        const uint64_t total = ctx->pair_total_sum;
        if (total > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "too many forward matches");
        CK(ctx->d_fwd_alt.ensure((size_t)total + 1));
        // this rank's own records: local layout -> canonical layout (the boundary records of the other ranks
        // arrive through l3d_shard_forward_unpack)
        (void)P;
        ctx->cnt.gpu_launches += launch_fwd_move(ctx->d_fwd_cnt.p + r0, r0, r1, ctx->d_fwd_off_local.p, ctx->d_fwd_rec.p,
                                                 ctx->d_fwd_off.p, ctx->d_fwd_alt.p, st);
        std::swap(ctx->d_fwd_rec.p, ctx->d_fwd_alt.p);
        std::swap(ctx->d_fwd_rec.cap, ctx->d_fwd_alt.cap);
        ctx->total_fwd = total;
        ctx->cnt.forward_matches = total;
        ctx->fwd_planned = false;
        return L3D_OK;
    }
    case L3D_X_PROGRAMS: {
        CK(ctx->d_slice_g.ensure(L3D_MAX_WORLD + 1));
        CK(cudaMemcpyAsync(ctx->d_slice_g.p, ctx->slice_g.data(), ((size_t)world + 1) * 4, cudaMemcpyHostToDevice, st));
        if (stride_bytes * (uint64_t)world / 16 > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "fold programs too large");
        ctx->cnt.gpu_launches += launch_k3_adopt_programs(src, stride_bytes, world, ctx->d_slice_g.p, S,
                                                          ctx->d_prog_off.p, ctx->d_prog_nh.p, ctx->d_fwd_score.p, st);
        ctx->prog_all = src;
        return L3D_OK;
    }
    case L3D_X_HYPOTHESES: {
        std::vector<uint32_t> fbase(world + 1, 0);
        for (int q = 0; q < world; ++q) fbase[q + 1] = fbase[q] + (uint32_t)nvar[q];
        CK(ctx->d_slice_g.ensure(2 * (L3D_MAX_WORLD + 1)));
        CK(cudaMemcpyAsync(ctx->d_slice_g.p, ctx->slice_g.data(), ((size_t)world + 1) * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_slice_g.p + L3D_MAX_WORLD + 1, fbase.data(), ((size_t)world + 1) * 4,
                           cudaMemcpyHostToDevice, st));
        CK(ctx->d_filt_all.ensure((size_t)fbase[world] + 1));
        ctx->cnt.gpu_launches += launch_hyp_adopt(src, stride_bytes, world, ctx->d_slice_g.p,
                                                  ctx->d_slice_g.p + L3D_MAX_WORLD + 1, S, ctx->d_entries.p,
                                                  ctx->d_filt_cnt.p, ctx->d_filt_off.p, ctx->d_filt_all.p, st);
        ctx->filt_all = ctx->d_filt_all.p;
        std::vector<ShardHypHdr> hdr(world);
        for (int q = 0; q < world; ++q)
            CK(cudaMemcpyAsync(&hdr[q], src + (uint64_t)q * stride_bytes, sizeof(ShardHypHdr), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->shard_sim_evals = ctx->shard_scored = ctx->shard_filtered = 0;
        for (int q = 0; q < world; ++q) {
            ctx->shard_sim_evals += hdr[q].sim_evals;
            ctx->shard_scored += hdr[q].scored;
            ctx->shard_filtered += hdr[q].filt_n;
        }
        return score_hypotheses_ready(ctx, nullptr, nullptr);
    }
    case L3D_X_EDGES: {
        uint64_t total = 0;
        for (int q = 0; q < world; ++q) total += nvar[q];
        if (total > 0x7ffffff0ull) return fail(L3D_ERR_CAPACITY, "too many edges");
        CK(ctx->d_edges_all.ensure(total * 12 + 16));
        uint64_t at = 0;
        for (int q = 0; q < world; ++q) {
            if (nvar[q])
                CK(cudaMemcpyAsync(ctx->d_edges_all.p + at * 12, src + (uint64_t)q * stride_bytes, nvar[q] * 12,
                                   cudaMemcpyDeviceToDevice, st));
            at += nvar[q];
        }
        ctx->edges_all = ctx->d_edges_all.p;
        ctx->n_edges_all = (uint32_t)total;
        return L3D_OK;
    }
    }
    return fail(L3D_ERR_ARG, "unknown exchange kind %d", kind);
}

// ---- FORWARD records, all-to-all.  After the FORWARD exchange every rank holds all per-row counts, hence the
// canonical record range of every pair (refresh_pair_totals); the records of a pair are contiguous there.
//   plan    send_records[d] / recv_records[q]: records this rank sends to rank d / receives from rank q
//           (0 for itself), summed over the boundary pairs on the host -- no device work
//   pack    this rank's boundary pairs grouped by destination rank (ascending), pairs ascending inside a group
//   unpack  the received pairs, grouped by source rank (ascending) = ascending pair index, into the canonical store
static int owner_of_view(const l3d_ctx* ctx, uint32_t v)
{
    int q = 0;
    while (q + 1 < ctx->world && ctx->slice_view[q + 1] <= v) ++q;
    return q;
}
// copy the record blocks `items` ({src, dst, n}) with one kernel
static int copy_record_blocks(l3d_ctx* ctx, const std::vector<uint4>& items, const FwdRec* src, FwdRec* dst)
{
    if (items.empty()) return L3D_OK;
    cudaStream_t st = ctx->stream;
    std::vector<uint32_t> chunk_item, chunk_first(items.size());
    for (size_t i = 0; i < items.size(); ++i) {
        chunk_first[i] = (uint32_t)chunk_item.size();
        for (uint32_t c = 0; c < (items[i].z + 127u) / 128u; ++c) chunk_item.push_back((uint32_t)i);
    }
    if (chunk_item.empty()) return L3D_OK;
    CK(ctx->d_blk_items.ensure(items.size()));
    CK(ctx->d_blk_chunks.ensure(chunk_item.size() + chunk_first.size()));
    CK(cudaMemcpyAsync(ctx->d_blk_items.p, items.data(), items.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_blk_chunks.p, chunk_item.data(), chunk_item.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_blk_chunks.p + chunk_item.size(), chunk_first.data(), chunk_first.size() * 4,
                       cudaMemcpyHostToDevice, st));
    ctx->cnt.gpu_launches += launch_rec_blocks(ctx->d_blk_items.p, ctx->d_blk_chunks.p, ctx->d_blk_chunks.p + chunk_item.size(),
                                               (uint32_t)chunk_item.size(), src, dst, st);
    // the pageable host vectors are staged by the runtime before cudaMemcpyAsync returns
    return L3D_OK;
}

int l3d_shard_forward_plan(l3d_ctx* ctx, uint64_t* send_records, uint64_t* recv_records)
{
    int rc = check_phase(ctx, L3D_X_FORWARD);
    if (rc) return rc;
    if (!send_records || !recv_records) return fail(L3D_ERR_ARG, "NULL argument");
    const int world = ctx->world;
    if (world > L3D_MAX_WORLD_C) return fail(L3D_ERR_ARG, "world %d too large", world);
    for (int q = 0; q < world; ++q) ctx->fwd_send[q] = ctx->fwd_recv[q] = 0;
    for (const HostPair& hp : ctx->pairs) {
        const int so = owner_of_view(ctx, hp.src), to = owner_of_view(ctx, hp.tgt);
        if (so == to) continue;
        if (so == ctx->rank) ctx->fwd_send[to] += hp.fwd_total;
        if (to == ctx->rank) ctx->fwd_recv[so] += hp.fwd_total;
    }
    for (int q = 0; q < world; ++q) {
        send_records[q] = ctx->fwd_send[q];
        recv_records[q] = ctx->fwd_recv[q];
    }
    ctx->fwd_planned = true;
    return L3D_OK;
}

int l3d_shard_forward_pack(l3d_ctx* ctx, void* dst, uint64_t cap_bytes, int device_ptr)
{
    int rc = check_phase(ctx, L3D_X_FORWARD);
    if (rc) return rc;
    if (!ctx->fwd_planned) return fail(L3D_ERR_STATE, "l3d_shard_forward_plan has not run");
    const int world = ctx->world;
    uint64_t total = 0;
    for (int d = 0; d < world; ++d) total += ctx->fwd_send[d];
    if (total * sizeof(FwdRec) > cap_bytes)
        return fail(L3D_ERR_CAPACITY, "need %llu bytes", (unsigned long long)(total * sizeof(FwdRec)));
    if (!total) return L3D_OK;
    if (!dst) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    FwdRec* out = (FwdRec*)dst;
    DevBuf<FwdRec>& stg = ctx->d_bx_rec;
    if (!device_ptr) {
        CK(stg.ensure((size_t)total + 1));
        out = stg.p;
    }
    // d_fwd_rec holds the canonical layout since the FORWARD import: this rank's pairs are in place there
    std::vector<uint4> items;
    uint64_t at = 0;
    for (int d = 0; d < world; ++d) {
        if (!ctx->fwd_send[d]) continue;
        for (const HostPair& hp : ctx->pairs) {
            if (!hp.fwd_total || owner_of_view(ctx, hp.src) != ctx->rank || owner_of_view(ctx, hp.tgt) != d) continue;
            items.push_back(make_uint4(hp.rec_start, (uint32_t)at, (uint32_t)hp.fwd_total, 0u));
            at += hp.fwd_total;
        }
    }
    rc = copy_record_blocks(ctx, items, ctx->d_fwd_rec.p, out);
    if (rc) return rc;
    if (!device_ptr) {
        CK(cudaMemcpyAsync(dst, stg.p, total * sizeof(FwdRec), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return L3D_OK;
}

int l3d_shard_forward_unpack(l3d_ctx* ctx, const void* src, uint64_t bytes, int device_ptr)
{
    int rc = check_phase(ctx, L3D_X_FORWARD);
    if (rc) return rc;
    if (!ctx->fwd_planned) return fail(L3D_ERR_STATE, "l3d_shard_forward_plan has not run");
    const int world = ctx->world;
    uint64_t total = 0;
    for (int q = 0; q < world; ++q) total += ctx->fwd_recv[q];
    if (bytes != total * sizeof(FwdRec))
        return fail(L3D_ERR_ARG, "expected %llu bytes, got %llu", (unsigned long long)(total * sizeof(FwdRec)),
                    (unsigned long long)bytes);
    if (!total) return L3D_OK;
    if (!src) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const FwdRec* in = (const FwdRec*)src;
    if (!device_ptr) {
        DevBuf<unsigned char>& stg = ctx->d_xchg_stage[L3D_X_FORWARD];
        CK(stg.ensure(bytes));
        CK(cudaMemcpyAsync(stg.p, src, bytes, cudaMemcpyHostToDevice, st));
        in = (const FwdRec*)stg.p;
    }
    // pairs are sorted by source view, so "source rank ascending, pairs ascending" is plain pair order
    std::vector<uint4> items;
    uint64_t at = 0;
    for (const HostPair& hp : ctx->pairs) {
        if (!hp.fwd_total || owner_of_view(ctx, hp.tgt) != ctx->rank || owner_of_view(ctx, hp.src) == ctx->rank) continue;
        items.push_back(make_uint4((uint32_t)at, hp.rec_start, (uint32_t)hp.fwd_total, 0u));
        at += hp.fwd_total;
    }
    return copy_record_blocks(ctx, items, in, ctx->d_fwd_rec.p);
}

// sizes_out[q] = payload bytes of rank q; *redo != 0: some blob did not fit / some program store
// overflowed -- nothing was imported, every rank must repeat the exchange with l3d_shard_blob_size
int l3d_shard_import_hdr(l3d_ctx* ctx, int kind, const void* all, uint64_t stride_bytes, int world, uint64_t* sizes_out,
                         int* redo)
{
    int rc = check_phase(ctx, kind);
    if (rc) return rc;
    if (!all || !sizes_out || !redo) return fail(L3D_ERR_ARG, "NULL argument");
    if (world != ctx->world || world > L3D_MAX_WORLD) return fail(L3D_ERR_ARG, "world %d does not match the plan", world);
    CK(cudaSetDevice(ctx->device));
    ShardBlobHdr hdr[L3D_MAX_WORLD];
    CK(cudaMemcpy2DAsync(hdr, sizeof(ShardBlobHdr), all, stride_bytes, sizeof(ShardBlobHdr), (size_t)world,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *redo = 0;
    for (int q = 0; q < world; ++q) {
        if (hdr[q].kind != (uint32_t)kind) return fail(L3D_ERR_ARG, "blob %d is of kind %u, expected %d", q, hdr[q].kind, kind);
        sizes_out[q] = hdr[q].payload_bytes;
        if (hdr[q].payload_bytes + sizeof(ShardBlobHdr) > stride_bytes) *redo = 1;
        if (kind == L3D_X_PROGRAMS && (hdr[q].flags & 4u)) *redo = 1;
        // a sender whose finish pass overflowed its filtered-match store (2) or timed out on a dependency (8)
        // fails locally; every rank sees the same header, so every rank reports it instead of adopting lists
        // that were cut short and walking into the next collective alone
        if (kind == L3D_X_HYPOTHESES && (hdr[q].flags & 2u))
            return fail(L3D_ERR_CAPACITY, "filtered-match store overflow on rank %d", q);
        if (kind == L3D_X_HYPOTHESES && (hdr[q].flags & 8u))
            return fail(L3D_ERR_STATE, "internal: scoring dependency wait timed out on rank %d", q);
    }
    if (*redo) return L3D_OK;
    return l3d_shard_import(ctx, kind, (const unsigned char*)all + sizeof(ShardBlobHdr), stride_bytes, world, sizes_out, 1);
}
