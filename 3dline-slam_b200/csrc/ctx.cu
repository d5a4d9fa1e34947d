// ctx.cu -- context, host orchestration and the C ABI of libl3dpp_b200.so (include/l3dpp_b200.h).
//
// Host side of Line3D::matchImages / computeMatches / reconstruct3Dlines (src/line3D.cc:496-640,
// 846-930, 2018-2118): camera set-up, translation, pair list, per-batch K1/K2, the
// scoring phases, K4, and the (unchanged) clustering on the host.  No CPU fallback: every
// compute entry point needs a CUDA device.
#include <fstream>
#include <thread>

#include "ctx.h"

static thread_local std::string g_err;
int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

const char* l3d_last_error(void) { return g_err.c_str(); }
const char* l3d_version(void) { return "l3dpp-b200 0.1 (sm_100a)"; }

int l3d_ctx_create(l3d_ctx** out, int device)
{
    if (!out) return fail(L3D_ERR_ARG, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(L3D_ERR_CUDA, "no CUDA device available (%s); libl3dpp_b200 has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(L3D_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    l3d_ctx* c = new l3d_ctx();
    c->device = device;
    if (cudaMallocHost((void**)&c->rb, (size_t)1 << 20) == cudaSuccess) {
        c->rb_cap = (size_t)1 << 20;
    } else {
        delete c;
        return fail(L3D_ERR_CUDA, "cudaMallocHost of the read-back scratch failed");
    }
    *out = c;
    return L3D_OK;
}

void l3d_ctx_destroy(l3d_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->tm.reset();
    delete ctx;
}

int l3d_ctx_set_stream(l3d_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    ctx->stream = (cudaStream_t)cuda_stream;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------
int l3d_scene_begin(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx) leave_translated(ctx);  // a call that failed half-way may have left the cameras shifted
    ctx->views.clear();
    ctx->cam2view.clear();
    ctx->stream_mode = false;
    ctx->by_worldpoints = false;
    ctx->committed = false;
    ctx->stage = 0;
    ctx->pairs.clear();
    return L3D_OK;
}

static int add_view(l3d_ctx* ctx, const l3d_view* view, const float* segs, const uint32_t* nbrs, uint32_t nn, bool wps,
                    bool copy_segs = true);

int l3d_scene_add_view(l3d_ctx* ctx, const l3d_view* view, const float* segs, const uint32_t* nbrs, uint32_t nn)
{
    return add_view(ctx, view, segs, nbrs, nn, false);
}

// neighbors_by_worldpoints = true (Line3D::Line3D, src/line3D.cc:60-75): the list holds the world
// points the view observes; the visual neighbours are chosen by l3d_match_images
int l3d_scene_add_view_wps(l3d_ctx* ctx, const l3d_view* view, const float* segs, const uint32_t* wps, uint32_t nw)
{
    return add_view(ctx, view, segs, wps, nw, true);
}

static int add_view(l3d_ctx* ctx, const l3d_view* view, const float* segs, const uint32_t* nbrs, uint32_t nn, bool wps,
                    bool copy_segs)
{
    if (!ctx || !view) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx->committed) return fail(L3D_ERR_STATE, "scene already committed; call l3d_scene_begin");
    if (!ctx->views.empty() && ctx->by_worldpoints != wps)
        return fail(L3D_ERR_ARG, "explicit neighbour lists and world-point lists cannot be mixed in one scene");
    ctx->by_worldpoints = wps;
    // same argument checks as Line3D::addImage (src/line3D.cc:123-205)
    if (std::max(view->width, view->height) < 400)
        return fail(L3D_ERR_ARG, "image is too small for reliable results: %u px (larger side should be >= 400px)",
                    std::max(view->width, view->height));
    for (auto& hv : ctx->views)
        if (hv.v.cam_id == view->cam_id) return fail(L3D_ERR_ARG, "camera ID [%u] already in use!", view->cam_id);
    if (nn == 0 && !wps) return fail(L3D_ERR_ARG, "view [%u] has no visual neighbors!", view->cam_id);
    if (nn && !nbrs) return fail(L3D_ERR_ARG, "NULL argument");
    if (view->num_segs == 0 || !segs) return fail(L3D_ERR_ARG, "no line segments found in image [%u]!", view->cam_id);
    HostView hv;
    hv.v = *view;
    // l3d_scene_set commits inside the same call: the segments go from the caller's buffer straight into the pinned
    // staging (one host copy instead of two); the staging keeps them readable for the TXT writer
    if (copy_segs) hv.segs.assign(segs, segs + 4 * (size_t)view->num_segs);
    else hv.ext_segs = segs;
    if (wps) hv.wps.assign(nbrs, nbrs + nn);
    else hv.nbrs.assign(nbrs, nbrs + nn);
    hv.cam.init(view->K, view->R, view->t);
    ctx->views.push_back(std::move(hv));
    return L3D_OK;
}

// begin + add_view x n + commit in one call (the views' segments / neighbour lists are concatenated)
static int scene_set(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                     const uint32_t* nbrs_concat, const uint32_t* nbr_counts, bool wps);
int l3d_scene_set(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                  const uint32_t* nbrs_concat, const uint32_t* nbr_counts)
{
    return scene_set(ctx, views, n_views, segs_concat, nbrs_concat, nbr_counts, false);
}
int l3d_scene_set_wps(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                      const uint32_t* wps_concat, const uint32_t* wp_counts)
{
    return scene_set(ctx, views, n_views, segs_concat, wps_concat, wp_counts, true);
}
static int scene_set(l3d_ctx* ctx, const l3d_view* views, uint32_t n_views, const float* segs_concat,
                     const uint32_t* nbrs_concat, const uint32_t* nbr_counts, bool wps)
{
    if (!ctx || !views || !segs_concat || !nbrs_concat || !nbr_counts) return fail(L3D_ERR_ARG, "NULL argument");
    int rc = l3d_scene_begin(ctx);
    if (rc) return rc;
    size_t so = 0, no = 0;
    for (uint32_t i = 0; i < n_views; ++i) {
        rc = add_view(ctx, &views[i], segs_concat + 4 * so, nbrs_concat + no, nbr_counts[i], wps, false);
        if (rc) return rc;
        so += views[i].num_segs;
        no += nbr_counts[i];
    }
    return l3d_scene_commit(ctx);
}

int l3d_scene_commit(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->views.empty()) return fail(L3D_ERR_STATE, "no images to match! forgot to add them?");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::sort(ctx->views.begin(), ctx->views.end(),
              [](const HostView& a, const HostView& b) { return a.v.cam_id < b.v.cam_id; });
    ctx->cam2view.clear();
    uint64_t S = 0;
    for (size_t i = 0; i < ctx->views.size(); ++i) {
        ctx->cam2view[ctx->views[i].v.cam_id] = (uint32_t)i;
        ctx->views[i].seg_off = (uint32_t)S;
        S += ctx->views[i].v.num_segs;
    }
    for (auto& hv : ctx->views) {  // neighbour camera ids -> ascending, unique view indices
        hv.nb_views.clear();
        for (uint32_t cam : hv.nbrs) {
            auto f = ctx->cam2view.find(cam);
            if (f != ctx->cam2view.end()) hv.nb_views.push_back(f->second);
        }
        std::sort(hv.nb_views.begin(), hv.nb_views.end());
        hv.nb_views.erase(std::unique(hv.nb_views.begin(), hv.nb_views.end()), hv.nb_views.end());
    }
    if (S > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "too many segments (%llu)", (unsigned long long)S);
    ctx->S = (uint32_t)S;
    const uint32_t V = (uint32_t)ctx->views.size();
    // pinned staging (kept by the context): one H2D copy of all segments and of their view ids
    const size_t need = std::max<size_t>(S, 1) * (sizeof(float4) + sizeof(uint32_t));
    if (need > ctx->pinned_cap) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_cap = 0;
        CK(cudaMallocHost(&ctx->pinned, need + need / 2));
        ctx->pinned_cap = need + need / 2;
    }
    float4* hseg = (float4*)ctx->pinned;
    uint32_t* seg_view = (uint32_t*)(hseg + std::max<size_t>(S, 1));
    auto stage_views = [&](uint32_t v0, uint32_t v1) {
        for (uint32_t v = v0; v < v1; ++v) {
            HostView& hv = ctx->views[v];
            memcpy(hseg + hv.seg_off, hv.ext_segs ? hv.ext_segs : hv.segs.data(), (size_t)hv.v.num_segs * sizeof(float4));
            hv.ext_segs = nullptr;
            std::fill(seg_view + hv.seg_off, seg_view + hv.seg_off + hv.v.num_segs, v);
        }
    };
    if (S >= (1u << 19) && V >= 8) {  // tens of MB: one core copies at ~10 GB/s, the H2D link takes five times that
        const uint32_t T = 4;
        std::vector<std::thread> th;
        for (uint32_t t = 1; t < T; ++t) th.emplace_back(stage_views, V * t / T, V * (t + 1) / T);
        stage_views(0, V / T);
        for (auto& x : th) x.join();
    } else {
        stage_views(0, V);
    }
    CK(ctx->d_segs.ensure(S));
    CK(ctx->d_seg_view.ensure(S));
    CK(ctx->d_desc.ensure(S));
    CK(ctx->d_rays.ensure(S));
    CK(ctx->d_midray.ensure(3 * S));
    CK(ctx->d_planes.ensure(S));
    CK(ctx->d_v32.ensure(S));
    CK(ctx->d_view_xb.ensure(V));
    CK(ctx->d_views.ensure(V));
    CK(cudaMemcpyAsync(ctx->d_segs.p, hseg, S * sizeof(float4), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_seg_view.p, seg_view, S * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->committed = true;
    ctx->stage = 0;
    ctx->k3_list_max = 0;
    return L3D_OK;
}

// ViewDev table from the host cameras (C is the current, possibly translated, centre)
int upload_views(l3d_ctx* ctx)
{
    const uint32_t V = (uint32_t)ctx->views.size();
    std::vector<ViewDev> vd(V);
    for (uint32_t v = 0; v < V; ++v) {
        const HostView& hv = ctx->views[v];
        ViewDev& d = vd[v];
        d.C[0] = hv.cam.C.x; d.C[1] = hv.cam.C.y; d.C[2] = hv.cam.C.z;
        memcpy(d.RtKinv, hv.cam.RtKinv.m, sizeof(d.RtKinv));
        d.k = hv.k;
        d.median_depth = hv.median_depth;
        d.seg_off = hv.seg_off;
        d.n_seg = hv.v.num_segs;
        d.cam_id = hv.v.cam_id;
        d.xb = 0.0f;
        d.order = v;
        d.needed = (ctx->view_needed.size() == V) ? ctx->view_needed[v] : 1u;
    }
    // pageable source: the runtime stages it before the call returns, no synchronisation needed
    CK(cudaMemcpyAsync(ctx->d_views.p, vd.data(), V * sizeof(ViewDev), cudaMemcpyHostToDevice, ctx->stream));
    return L3D_OK;
}

// Line3D::translate (src/line3D.cc:643-680): per-axis median of the camera centres
void compute_translation(l3d_ctx* ctx)
{
    double* tr[3] = {&ctx->translation.x, &ctx->translation.y, &ctx->translation.z};
    for (int a = 0; a < 3; ++a) {
        *tr[a] = 0.0;
        std::vector<double> c;
        for (auto& hv : ctx->views) {
            if (!hv.current) continue;  // view_order_ holds the current views only (src/line3D.cc:396-430)
            const double val = a == 0 ? hv.cam.C.x : (a == 1 ? hv.cam.C.y : hv.cam.C.z);
            if (std::fabs(val) > 1e-12) c.push_back(val);
        }
        if (!c.empty()) {
            std::sort(c.begin(), c.end());
            *tr[a] = c[c.size() / 2];
        }
    }
}
void apply_translation(l3d_ctx* ctx, double sign)
{
    const hg::V3 tv{sign * ctx->translation.x, sign * ctx->translation.y, sign * ctx->translation.z};
    for (auto& hv : ctx->views)
        if (hv.current) hv.cam.translate(tv);
}
// translate() / untranslate() as a guarded pair (src/line3D.cc:568,637 and :2065,2140).  The host cameras
// are shifted in place like the reference's; `translated` remembers that a shift is pending, so a call
// that fails half-way (capacity error, CUDA error, aborted sharded step) or a repeated stage-1/2 call
// cannot stack a second shift on top of the first: the pending one is undone -- with the translation it
// was made with -- before a new median is taken, and before any getter reads the cameras.
void enter_translated(l3d_ctx* ctx)
{
    leave_translated(ctx);
    compute_translation(ctx);
    apply_translation(ctx, -1.0);
    ctx->translated = true;
}
void leave_translated(l3d_ctx* ctx)
{
    if (!ctx->translated) return;
    apply_translation(ctx, +1.0);
    ctx->translated = false;
}

__global__ void pair_totals_kernel(const PairDev* __restrict__ pairs, uint32_t P, const uint32_t* __restrict__ fwd_off,
                                   const uint32_t* __restrict__ fwd_cnt, uint32_t* __restrict__ out)
{
    // forward records of a pair are contiguous (rows ascending): total = end(last row) - start(first row)
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const PairDev& D = pairs[p];
    if (D.n_src == 0) {
        out[p] = 0;
        return;
    }
    const uint32_t r0 = D.row_base, r1 = D.row_base + D.n_src - 1;
    out[p] = fwd_off[r1] + fwd_cnt[r1] - fwd_off[r0];
}


// per-pair forward-record totals -> host (list capacities, record ranges of the canonical layout)
int refresh_pair_totals(l3d_ctx* ctx)
{
    const uint32_t P = (uint32_t)ctx->pairs.size();
    cudaStream_t st = ctx->stream;
    if (!P) return L3D_OK;
    CK(ctx->stream_mode ? ensure_roomy(ctx->d_scan_tmp, (size_t)std::max(P, ctx->total_tgt_rows) + 2, (size_t)1 << 18)
                         : ctx->d_scan_tmp.ensure((size_t)std::max(P, ctx->total_tgt_rows) + 2));
    pair_totals_kernel<<<(P + 255) / 256, 256, 0, st>>>(ctx->d_pairs.p, P, ctx->d_fwd_off.p, ctx->d_fwd_cnt.p,
                                                        ctx->d_scan_tmp.p);
    ctx->cnt.gpu_launches++;
    std::vector<uint32_t> tot_pageable;
    uint32_t* tot = ctx->rb_at<uint32_t>(l3d_ctx::RB_BIG);
    if (!ctx->rb_fits((size_t)P * 4)) {
        tot_pageable.resize(P);
        tot = tot_pageable.data();
    }
    CK(cudaMemcpyAsync(tot, ctx->d_scan_tmp.p, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t run = 0;
    for (uint32_t p = 0; p < P; ++p) {
        ctx->pairs[p].fwd_total = tot[p];
        ctx->pairs[p].rec_start = (uint32_t)run;
        run += tot[p];
    }
    ctx->pair_total_sum = run;
    return L3D_OK;
}

// batches over local pairs: bounded bit-mask size (fills mask_base / batch_row0 of pairs_h, the
// K1 CTA list and ctx->batches)
void plan_batches(l3d_ctx* ctx)
{
    const uint32_t P = (uint32_t)ctx->pairs.size();
    // Mask words of one batch.  A batch also holds K2's scratch, 72 B per K1 candidate: ~40 B per mask word at the
    // 2 % candidate rate of config 4, ~110 B at the 5 % of config 2.  2 GB of mask (three batches on config 4, 38 GB
    // in use) instead of 512 MB (twelve batches, 21 GB) took 5 ms off an 78 ms step: fewer partial waves per
    // launch and fewer host synchronisations.  L3D_MASK_WORDS_LOG2 overrides (27 ... 31).
    static int words_log2 = -1;
    if (words_log2 < 0) {
        const char* ev = getenv("L3D_MASK_WORDS_LOG2");
        words_log2 = ev ? std::min(31, std::max(20, atoi(ev))) : 29;
    }
    const uint64_t max_words = 1ull << words_log2;
    const uint32_t rows_per_cta = (uint32_t)k1_rows_per_cta();
    ctx->batches.clear();
    ctx->ctas_h.clear();
    Batch cur{};
    bool open = false;
    auto close = [&]() {
        if (open) {
            cur.n_ctas = (uint32_t)ctx->ctas_h.size() - cur.cta0;
            ctx->batches.push_back(cur);
            open = false;
        }
    };
    for (uint32_t p = 0; p < P; ++p) {
        if (!ctx->pairs[p].local) {
            close();  // keep batch rows contiguous
            continue;
        }
        PairDev& d = ctx->pairs_h[p];
        const uint64_t w = (uint64_t)d.words * d.n_src;
        if (open && cur.mask_words + w > max_words) close();
        if (!open) {
            cur = Batch{};
            cur.pair0 = p;
            cur.row0 = d.row_base;
            cur.cta0 = (uint32_t)ctx->ctas_h.size();
            open = true;
        }
        d.mask_base = cur.mask_words;
        d.batch_row0 = cur.row0;
        cur.mask_words += w;
        cur.max_tgt = std::max(cur.max_tgt, d.n_tgt);
        cur.n_rows += d.n_src;
        cur.pair1 = p + 1;
        ctx->pairs[p].batch = (uint32_t)ctx->batches.size();
        for (uint32_t t = 0; t * rows_per_cta < d.n_src; ++t) ctx->ctas_h.push_back(K1Cta{p, t});
    }
    close();

}

// ------------------------------------------------------------------------------------------
// planning: pair list, batches, incident lists
// ------------------------------------------------------------------------------------------
int plan_pairs(l3d_ctx* ctx)
{
    const l3d_params& prm = ctx->prm;
    const uint32_t V = (uint32_t)ctx->views.size();
    // visual neighbours = fixed neighbours that exist (src/line3D.cc:604-616); nb_views is the
    // ascending list of their view indices (view index order == camera id order), made at commit.
    // computeMatches pair order (src/line3D.cc:848-887): (s, t) is skipped iff the pair was already
    // created from the other side, i.e. t < s and s is a neighbour of t.
    if (ctx->by_worldpoints) {  // Line3D::matchImages, src/line3D.cc:604-625 (after translate())
        std::vector<const hg::Camera*> cams(V);
        std::vector<float> md(V);
        std::vector<std::vector<uint32_t>> wps(V), nb;
        for (uint32_t v = 0; v < V; ++v) {
            cams[v] = &ctx->views[v].cam;
            md[v] = ctx->views[v].median_depth;
            wps[v] = ctx->views[v].wps;
        }
        hg::visual_neighbors_from_worldpoints(cams, md, wps, prm.num_neighbors, nb);
        for (uint32_t v = 0; v < V; ++v) ctx->views[v].nb_views = nb[v];
    }
    ctx->pairs.clear();
    for (uint32_t s = 0; s < V; ++s)
        for (uint32_t t : ctx->views[s].nb_views) {
            const std::vector<uint32_t>& nt = ctx->views[t].nb_views;
            if (t < s && std::binary_search(nt.begin(), nt.end(), s)) continue;
            HostPair hp;
            hp.src = s;
            hp.tgt = t;
            hp.batch = 0;
            ctx->pairs.push_back(hp);
        }
    const uint32_t P = (uint32_t)ctx->pairs.size();
    const int world = prm.shard_world > 1 ? prm.shard_world : 1;
    const int rank = world > 1 ? prm.shard_rank : 0;
    ctx->pairs_h.assign(P, PairDev{});
    // shard = a contiguous slice of reference (source) views, balanced by the number of segment-pair
    // tests of the pairs they own; pairs are ordered by source view, so a slice owns a contiguous
    // block of pairs, of pair rows and of segments
    if (world > L3D_MAX_WORLD) return fail(L3D_ERR_ARG, "at most %d shards", L3D_MAX_WORLD);
    std::vector<uint64_t> vtests(V + 1, 0);
    for (uint32_t p = 0; p < P; ++p)
        vtests[ctx->pairs[p].src + 1] += (uint64_t)ctx->views[ctx->pairs[p].src].v.num_segs *
                                         ctx->views[ctx->pairs[p].tgt].v.num_segs;
    for (uint32_t v = 0; v < V; ++v) vtests[v + 1] += vtests[v];
    const uint64_t total_tests = vtests[V];
    ctx->world = world;
    ctx->rank = rank;
    ctx->slice_view.assign(world + 1, V);
    ctx->slice_view[0] = 0;
    for (int q = 1; q < world; ++q) {
        const uint64_t want = total_tests / (uint64_t)world * (uint64_t)q;
        uint32_t v = ctx->slice_view[q - 1];
        // first view whose cumulative test count (up to its middle) passes the q-th share
        while (v < V && vtests[v] + (vtests[v + 1] - vtests[v]) / 2 < want) ++v;
        ctx->slice_view[q] = v;
    }
    auto owner_of = [&](uint32_t v) {
        int q = 0;
        while (q + 1 < world && ctx->slice_view[q + 1] <= v) ++q;
        return q;
    };
    uint64_t row = 0, trow = 0;
    ctx->cnt.pair_tests = 0;
    ctx->cnt.num_pairs_local = 0;
    for (uint32_t p = 0; p < P; ++p) {
        HostPair& hp = ctx->pairs[p];
        hp.local = (owner_of(hp.src) == rank);
        const HostView& vs = ctx->views[hp.src];
        const HostView& vt = ctx->views[hp.tgt];
        PairDev& d = ctx->pairs_h[p];
        if (ctx->raw_mode) memcpy(d.F, ctx->F_override, sizeof(d.F));  // else: below, on several threads
        d.src_view = hp.src;
        d.tgt_view = hp.tgt;
        d.src_off = vs.seg_off;
        d.n_src = vs.v.num_segs;
        d.tgt_off = vt.seg_off;
        d.n_tgt = vt.v.num_segs;
        d.row_base = (uint32_t)row;
        d.tgt_base = (uint32_t)trow;
        d.words = (d.n_tgt + 31) / 32;
        d.emit_inverse = hp.tgt > hp.src ? 1u : 0u;  // !processed_[tgt] (src/line3D.cc:1994)
        // boundary pair: 1 + the rank that owns the target view (it receives the pair's records as inverse candidates)
        d.xflag = (owner_of(hp.tgt) != owner_of(hp.src)) ? (uint32_t)owner_of(hp.tgt) + 1u : 0u;
        row += d.n_src;
        trow += d.n_tgt;
        if (hp.local) {
            ctx->cnt.pair_tests += (uint64_t)d.n_src * d.n_tgt;
            ctx->cnt.num_pairs_local++;
        }
    }
    if (!ctx->raw_mode) {
        // fundamental matrices (src/line3D.cc:1082-1104): five 3x3 products and two inverses per pair; 10^4 pairs
        // of config 4 are a millisecond and more of a step on one core, with the GPU waiting
        auto fill_F = [&](uint32_t p0, uint32_t p1) {
            for (uint32_t p = p0; p < p1; ++p) {
                if (!ctx->pairs[p].local) continue;  // only K1 / K2 read F, and only for this rank's pairs
                const hg::M3 F = hg::fundamental(ctx->views[ctx->pairs[p].src].cam, ctx->views[ctx->pairs[p].tgt].cam);
                memcpy(ctx->pairs_h[p].F, F.m, sizeof(ctx->pairs_h[p].F));
            }
        };
        if (P >= 2048) {
            const uint32_t T = 4;
            std::vector<std::thread> th;
            for (uint32_t t = 1; t < T; ++t) th.emplace_back(fill_F, (uint32_t)((uint64_t)P * t / T), (uint32_t)((uint64_t)P * (t + 1) / T));
            fill_F(0, (uint32_t)((uint64_t)P / T));
            for (auto& x : th) x.join();
        } else {
            fill_F(0, P);
        }
    }
    if (row > 0xfffffff0ull || trow > 0xfffffff0ull)
        return fail(L3D_ERR_CAPACITY, "row index space exhausted (%llu rows)", (unsigned long long)row);
    ctx->total_rows = (uint32_t)row;
    ctx->total_tgt_rows = (uint32_t)trow;
    // views whose per-segment tables this rank reads: both views of every pair incident to its slice
    ctx->view_needed.assign(V, world > 1 ? 0u : 1u);
    if (world > 1) {
        const uint32_t lo = ctx->slice_view[rank], hi = ctx->slice_view[rank + 1];
        for (uint32_t v = lo; v < hi; ++v) ctx->view_needed[v] = 1u;
        for (uint32_t p = 0; p < P; ++p) {
            const uint32_t a = ctx->pairs[p].src, b = ctx->pairs[p].tgt;
            if ((a >= lo && a < hi) || (b >= lo && b < hi)) ctx->view_needed[a] = ctx->view_needed[b] = 1u;
        }
    }
    // slice boundaries in segments and in pair rows
    ctx->slice_g.assign(world + 1, ctx->S);
    ctx->slice_row.assign(world + 1, ctx->total_rows);
    for (int q = 0; q <= world; ++q) {
        const uint32_t v = ctx->slice_view[q];
        ctx->slice_g[q] = v < V ? ctx->views[v].seg_off : ctx->S;
        uint32_t r = ctx->total_rows;
        for (uint32_t p = 0; p < P; ++p)
            if (ctx->pairs[p].src >= v) {
                r = ctx->pairs_h[p].row_base;
                break;
            }
        ctx->slice_row[q] = r;
    }

    plan_batches(ctx);

    // incident pairs per view in list order: inverse blocks (source views ascending = the order
    // their storeInverseMatches ran), then forward blocks (targets ascending)
    ctx->inc_off_h.assign(V + 1, 0);
    ctx->inc_h.clear();
    std::vector<std::vector<uint32_t>> inv_of(V), fwd_of(V);
    for (uint32_t p = 0; p < P; ++p) {
        fwd_of[ctx->pairs[p].src].push_back(p);
        if (ctx->pairs_h[p].emit_inverse) inv_of[ctx->pairs[p].tgt].push_back(p);
    }
    for (uint32_t v = 0; v < V; ++v) {
        ctx->inc_off_h[v] = (uint32_t)ctx->inc_h.size();
        for (uint32_t p : inv_of[v]) ctx->inc_h.push_back(IncDev{p, 1u});  // pairs are sorted by src
        for (uint32_t p : fwd_of[v]) ctx->inc_h.push_back(IncDev{p, 0u});  // and by tgt within a src
    }
    ctx->inc_off_h[V] = (uint32_t)ctx->inc_h.size();
    ctx->cnt.num_pairs = P;
    ctx->cnt.num_views = V;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// stage 1 + 2
// ------------------------------------------------------------------------------------------
int set_params(l3d_ctx* ctx, const l3d_params* params)
{
    if (!params) return fail(L3D_ERR_ARG, "params is NULL");
    // the longest match list of the committed scene is cached between runs (score build); it depends on the
    // matching parameters, so a change invalidates it (a stale, too small value sized the row staging wrongly)
    if (ctx->have_params && (ctx->prm.knn != params->knn || ctx->prm.epipolar_overlap != params->epipolar_overlap ||
                             ctx->prm.num_neighbors != params->num_neighbors || ctx->prm.filter_mode != params->filter_mode ||
                             ctx->prm.max_image_width != params->max_image_width))
        ctx->k3_list_max = 0;
    ctx->prm = *params;
    l3d_params& p = ctx->prm;
    // parameter clamps of Line3D::matchImages (src/line3D.cc:517-536)
    p.num_neighbors = (uint32_t)std::max((int)p.num_neighbors, 2);
    p.sigma_a = (float)std::fmin(std::fabs((double)p.sigma_a), 90.0);
    ctx->two_sigA_sqr = 2.0f * p.sigma_a * p.sigma_a;
    ctx->epi_overlap = (float)std::fmin(std::fabs((double)p.epipolar_overlap), (double)0.99f);
    if (p.sigma_p < 0.0f) {
        // fixed sigma_p in world units (src/line3D.cc:525-530): k = sigma_p / med_scene_depth for every view
        // (View::update_k, include/view.h:138-141).  The reference takes med_scene_depth from
        // const_regularization_depth or, if that is negative, from views_avg_depths_[size/2] -- a map
        // indexed by POSITION as if it were a key (src/line3D.cc:557-565); only the well-defined case is built.
        if (!(p.const_reg_depth > 0.0f))
            return fail(L3D_ERR_ARG, "sigma_p < 0 (metric regulariser) needs const_regularization_depth > 0");
        ctx->fixed3D = true;
        p.sigma_p = std::fabs(p.sigma_p);
    } else {
        ctx->fixed3D = false;
        p.sigma_p = (float)std::fmax((double)0.1f, (double)p.sigma_p);
    }
    if (p.max_image_width <= 0)
        return fail(L3D_ERR_ARG, "max_image_width must be > 0 (the reference's bounds test rejects every pair otherwise)");
    ctx->have_params = true;
    return L3D_OK;
}

// Wait for the stream by polling: the step reads a count back after K1 and after K2 of every batch, and a blocking
// cudaStreamSynchronize puts the thread to sleep -- the wake-up (tens of microseconds) is GPU idle time each time
static cudaError_t spin_sync(cudaStream_t st)
{
    cudaError_t e;
    while ((e = cudaStreamQuery(st)) == cudaErrorNotReady) {
    }
    return e;
}

// per-segment tables (K0), then K1 + K2 over the planned batches -> forward store
int run_stage12_batches(l3d_ctx* ctx)
{
    cudaStream_t st = ctx->stream;
    // stream mode: the scratch sizes change a little from cycle to cycle; reallocate rarely (ensure_roomy)
    const bool roomy = ctx->stream_mode;
    auto grow = [roomy](auto& buf, size_t n, size_t keep = 0, cudaStream_t s2 = 0) {
        return roomy ? ensure_roomy(buf, n, (size_t)1 << 20, keep, s2) : buf.ensure(n, keep, s2);
    };
    const uint32_t V = (uint32_t)ctx->views.size(), P = (uint32_t)ctx->pairs.size(), S = ctx->S;

    // per-segment tables
    cudaEvent_t ev = ctx->tm.begin(L3D_T_PREP, st);
    CK(cudaMemsetAsync(ctx->d_view_xb.p, 0, V * sizeof(float), st));
    ctx->cnt.gpu_launches += launch_k0_prep(ctx->d_segs.p, ctx->d_seg_view.p, ctx->d_views.p, S,
                                            ctx->prm.max_image_width, ctx->d_desc.p, ctx->d_rays.p, ctx->d_midray.p, ctx->d_planes.p, ctx->d_v32.p,
                                            ctx->d_view_xb.p, st);
    ctx->tm.end(ev, st);

    CK(grow(ctx->d_fwd_off, (size_t)ctx->total_rows + 1));
    CK(grow(ctx->d_fwd_cnt, (size_t)ctx->total_rows + 1));
    CK(cudaMemsetAsync(ctx->d_fwd_cnt.p, 0, ((size_t)ctx->total_rows + 1) * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(ctx->d_fwd_off.p, 0, ((size_t)ctx->total_rows + 1) * sizeof(uint32_t), st));
    ctx->total_fwd = 0;
    ctx->cnt.candidates = 0;

    if (P) {
        uint32_t max_rows = 0;
        uint64_t max_words = 0;
        for (auto& b : ctx->batches) {
            max_rows = std::max(max_rows, b.n_rows);
            max_words = std::max(max_words, b.mask_words);
        }
        CK(grow(ctx->d_mask, max_words));
        CK(grow(ctx->d_cand_cnt, (size_t)max_rows + 1));
        CK(grow(ctx->d_cand_off, (size_t)max_rows + 1));
        CK(grow(ctx->d_fin_cnt, (size_t)max_rows + 1));
        CK(grow(ctx->d_fin_off, (size_t)max_rows + 1));
        CK(grow(ctx->d_row_epi, (size_t)max_rows + 1));
        CK(grow(ctx->d_row_epi_nat, (size_t)max_rows + 1));
        CK(grow(ctx->d_row_key, (size_t)max_rows + 1));
        CK(grow(ctx->d_perm, (size_t)max_rows + 1));
        CK(grow(ctx->d_iperm, (size_t)max_rows + 1));
        CK(grow(ctx->d_ncont, (size_t)max_rows + 1));
        CK(grow(ctx->d_fb_rows, (size_t)max_rows + 1));
        CK(grow(ctx->d_row_pair, (size_t)max_rows + 1));
        CK(grow(ctx->d_row_T, (size_t)max_rows + 1));
        CK(ctx->d_k2ctr.ensure(8));
        CK(ctx->d_k1_run.ensure(2));
        CK(cudaMemsetAsync(ctx->d_k1_run.p, 0, 2 * sizeof(unsigned long long), st));
        if (!ctx->n_sm) {
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, ctx->device));
            ctx->n_sm = prop.multiProcessorCount;
        }
        CK(grow(ctx->d_scan, scan_scratch_words(std::max(max_rows, ctx->total_tgt_rows) + 1) + 64));
        CK(grow(ctx->d_ctas, ctx->ctas_h.size()));
        CK(cudaMemcpyAsync(ctx->d_ctas.p, ctx->ctas_h.data(), ctx->ctas_h.size() * sizeof(K1Cta),
                           cudaMemcpyHostToDevice, st));
        CK(grow(ctx->d_pairs, P));
        CK(cudaMemcpyAsync(ctx->d_pairs.p, ctx->pairs_h.data(), P * sizeof(PairDev), cudaMemcpyHostToDevice, st));
    }

    uint32_t rec_base = 0;  // running start of this shard's forward records (local layout)
    for (const Batch& b : ctx->batches) {
        // K1
        cudaEvent_t e1 = ctx->tm.begin(L3D_T_PAIRTEST, st);
        cudaEvent_t ek = ctx->tm.begin(L3D_T_K1_KERNEL, st);
        ctx->cnt.gpu_launches +=
            launch_k1_rowsort(ctx->d_pairs.p, ctx->d_ctas.p + b.cta0, b.n_ctas, ctx->d_segs.p, ctx->d_view_xb.p,
                              ctx->d_row_epi_nat.p, ctx->d_row_epi.p, ctx->d_row_key.p, ctx->d_perm.p, ctx->d_iperm.p, st);
        ctx->cnt.gpu_launches +=
            launch_k1_pairtest(ctx->d_pairs.p, ctx->d_ctas.p + b.cta0, b.n_ctas, ctx->d_desc.p, ctx->d_row_epi.p,
                               ctx->d_row_key.p, ctx->d_perm.p, ctx->d_mask.p, ctx->d_cand_cnt.p, ctx->epi_overlap,
                               ctx->prm.filter_mode, ctx->d_k1_run.p, st);
        ctx->tm.end(ek, st);
        ctx->tm.ms[L3D_T_K1_LAUNCHES] += 1.0f;
        ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_cand_cnt.p, ctx->d_cand_off.p, b.n_rows, ctx->d_scan.p,
                                                 ctx->d_scan.cap, st);
        ctx->tm.end(e1, st);
        CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NCAND), ctx->d_cand_off.p + b.n_rows, sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, st));
        CK(spin_sync(st));
        const uint32_t n_cand = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NCAND);
        ctx->cnt.candidates += n_cand;
        // K2
        CK(grow(ctx->d_heap, (size_t)n_cand + 1));
        CK(grow(ctx->d_cand_rec, (size_t)n_cand + 1));
        CK(grow(ctx->d_fin_rec, (size_t)n_cand + 1));
        cudaEvent_t e2 = ctx->tm.begin(L3D_T_EXACT, st);
        int uses_ncont = 0;
        ctx->cnt.gpu_launches +=
            launch_k2_exact(ctx->d_pairs.p, ctx->d_ctas.p + b.cta0, b.n_ctas, b.n_rows, n_cand, ctx->d_segs.p,
                            ctx->d_rays.p, ctx->d_midray.p, ctx->d_planes.p, ctx->d_v32.p, ctx->d_desc.p, ctx->d_row_epi.p, ctx->d_perm.p, ctx->d_iperm.p,
                            ctx->d_views.p, ctx->d_mask.p, ctx->d_cand_off.p, ctx->d_heap.p, ctx->d_cand_rec.p,
                            ctx->d_fin_rec.p, ctx->d_fin_cnt.p, ctx->d_ncont.p, ctx->d_k2ctr.p, ctx->d_fb_rows.p, ctx->d_row_pair.p, ctx->d_row_T.p,
                            ctx->epi_overlap, ctx->prm.knn, ctx->prm.max_image_width, ctx->raw_mode ? 0 : 1, ctx->n_sm,
                            &uses_ncont, st);
        if (getenv("L3D_K2_DEBUG")) {  // contenders / rows handed to the row kernel, per batch
            uint32_t c2[4] = {0, 0, 0, 0};
            cudaMemcpyAsync(c2, ctx->d_k2ctr.p, sizeof(c2), cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            fprintf(stderr, "[k2] batch rows %u candidates %u contenders %u popped %u fallback rows %u\n", b.n_rows, n_cand, c2[0],
                    c2[2], c2[1]);
        }
        ctx->cnt.gpu_launches +=
            launch_scan_u32(ctx->d_fin_cnt.p, ctx->d_fin_off.p, b.n_rows, ctx->d_scan.p, ctx->d_scan.cap, st);
        CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NFIN), ctx->d_fin_off.p + b.n_rows, sizeof(uint32_t),
                           cudaMemcpyDeviceToHost, st));
        CK(spin_sync(st));
        const uint32_t n_fin = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NFIN);
        if ((uint64_t)rec_base + n_fin > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "too many forward matches");
        CK(grow(ctx->d_fwd_rec, (size_t)rec_base + n_fin, rec_base, st));
        ctx->cnt.gpu_launches +=
            launch_k2_compact(ctx->d_cand_off.p, ctx->d_fin_cnt.p, ctx->d_fin_off.p, rec_base, ctx->d_fin_rec.p,
                              ctx->d_fwd_rec.p, ctx->d_fwd_off.p + b.row0, b.n_rows, uses_ncont ? ctx->d_ncont.p : nullptr, st);
        CK(cudaMemcpyAsync(ctx->d_fwd_cnt.p + b.row0, ctx->d_fin_cnt.p, b.n_rows * sizeof(uint32_t),
                           cudaMemcpyDeviceToDevice, st));
        ctx->tm.end(e2, st);
        rec_base += n_fin;
    }
    ctx->cnt.pair_tests_run = 0;
    if (P) {  // read with the per-pair totals' synchronisation (refresh_pair_totals)
        CK(cudaMemcpyAsync(ctx->rb_at<unsigned long long>(l3d_ctx::RB_K1RUN), ctx->d_k1_run.p, sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, st));
        ctx->k1_run_pending = true;
    }
    ctx->total_fwd = rec_base;
    ctx->local_fwd = rec_base;
    ctx->prog_all = nullptr;
    ctx->filt_all = nullptr;
    ctx->edges_all = nullptr;
    ctx->stage3_phase = ctx->stage4_phase = 0;

    return L3D_OK;
}

int l3d_match_stage12(l3d_ctx* ctx, const l3d_params* params)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (!ctx->committed) return fail(L3D_ERR_STATE, "scene not committed");
    int rc = set_params(ctx, params);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->tm.reset();
    cudaEvent_t ev_total = ctx->tm.begin(L3D_T_TOTAL, st);

    // translate(), spatial regularisers (src/line3D.cc:568-590)
    if (!ctx->raw_mode) {
        enter_translated(ctx);
        for (auto& hv : ctx->views) {
            hv.k = ctx->fixed3D ? ctx->prm.sigma_p / ctx->prm.const_reg_depth : hv.cam.spatial_regularizer(ctx->prm.sigma_p);
            hv.median_depth = 0.0f;
        }
    }
    rc = plan_pairs(ctx);
    if (rc) return rc;
    rc = upload_views(ctx);
    if (rc) return rc;
    rc = run_stage12_batches(ctx);
    if (rc) return rc;
    rc = refresh_pair_totals(ctx);
    if (rc) return rc;
    ctx->tm.end(ev_total, st);
    CK(cudaStreamSynchronize(st));
    ctx->tm.collect();
    ctx->cnt.forward_matches = ctx->total_fwd;
    if (ctx->k1_run_pending) ctx->cnt.pair_tests_run = *ctx->rb_at<unsigned long long>(l3d_ctx::RB_K1RUN);
    ctx->k1_run_pending = false;
    ctx->stage = 1;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// stage 3: scoring
// ------------------------------------------------------------------------------------------
// stage 3: scoring (k3_dataflow.cu) in three phases: build (rows of this rank's view slice),
// fold (all rows, replicated), finish (rows of the slice).  With one rank the phases run back to
// back; with several the fold programs are exchanged between build and fold, and the hypotheses
// between finish and the affinity stage (abi.cu: l3d_shard_export / l3d_shard_import).
// ------------------------------------------------------------------------------------------
static K3Tables k3_tables(l3d_ctx* ctx)
{
    K3Tables t;
    t.views = ctx->d_views.p; t.seg_view = ctx->d_seg_view.p; t.pairs = ctx->d_pairs.p; t.inc = ctx->d_inc.p;
    t.inc_off = ctx->d_inc_off.p; t.rays = ctx->d_rays.p; t.fwd_off = ctx->d_fwd_off.p; t.fwd_cnt = ctx->d_fwd_cnt.p;
    t.fwd_rec = ctx->d_fwd_rec.p; t.fwd_score = ctx->d_fwd_score.p; t.fwd_row = ctx->d_fwd_row.p;
    t.inv_off = ctx->d_inv_off.p; t.inv_fill = ctx->d_inv_fill.p; t.inv_ent = ctx->d_inv_ent.p;
    t.L_off = ctx->d_L_off.p; t.L_f = ctx->d_L_f.p; t.L_meta = ctx->d_L_meta.p; t.L_score = ctx->d_L_score.p;
    const bool big = ctx->k3_big_rows;
    t.L_sib = big ? ctx->d_L_sib.p : nullptr; t.L_dir = big ? ctx->d_L_dir.p : nullptr;
    t.L_reg = big ? ctx->d_L_reg.p : nullptr; t.L_c = big ? ctx->d_L_c.p : nullptr; t.L_h = big ? ctx->d_L_h.p : nullptr;
    t.prog_off = ctx->d_prog_off.p; t.prog_nh = ctx->d_prog_nh.p;
    t.prog = ctx->prog_all ? const_cast<void*>(ctx->prog_all) : (void*)ctx->d_prog.p;
    t.prog_cap = (uint32_t)ctx->prog_cap;
    t.L_cnt = ctx->d_L_cnt.p; t.L_rec = ctx->prm.keep_scored ? ctx->d_L_rec.p : nullptr;
    t.view_max = ctx->d_view_max.p; t.filt_rec = ctx->d_filt_rec.p; t.filt_cap = (uint32_t)ctx->filt_cap;
    t.filt_off = ctx->d_filt_off.p; t.filt_cnt = ctx->d_filt_cnt.p; t.entries = ctx->d_entries.p;
    t.stats = ctx->d_stats.p; t.S = ctx->S; t.maxm = ctx->k3_maxm; t.two_sigA_sqr = ctx->two_sigA_sqr;
    t.g_lo = ctx->slice_g[ctx->rank]; t.g_hi = ctx->slice_g[ctx->rank + 1];
    t.cls = ctx->d_k3_cls.p;
    return t;
}

static int launch_records(l3d_ctx* ctx)
{
    return launch_k3_records(ctx->d_pairs.p, (uint32_t)ctx->pairs.size(), (uint32_t)ctx->total_fwd, ctx->d_fwd_row.p,
                             ctx->d_fwd_rec.p, ctx->d_fwd_score.p, ctx->d_inv_off.p, ctx->d_inv_fill.p, ctx->d_inv_ent.p,
                             ctx->slice_view[ctx->rank], ctx->slice_view[ctx->rank + 1], ctx->stream);
}

// pre-pass + build
int l3d_score_build(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 1) return fail(L3D_ERR_STATE, "l3d_match_stage12 has not run");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t V = (uint32_t)ctx->views.size(), P = (uint32_t)ctx->pairs.size(), S = ctx->S;
    for (uint32_t v = 0; v < V; ++v)
        if (ctx->inc_off_h[v + 1] - ctx->inc_off_h[v] > (uint32_t)k3_wf_max_inc())
            return fail(L3D_ERR_CAPACITY, "view %u takes part in more than %d matched pairs", ctx->views[v].v.cam_id,
                        k3_wf_max_inc());
    ctx->prog_all = nullptr;
    ctx->ev_total3 = ctx->tm.begin(L3D_T_TOTAL, st);
    cudaEvent_t ev = ctx->tm.begin(L3D_T_SCORE, st);
    const size_t F = (size_t)ctx->total_fwd;
    const size_t TR = (size_t)ctx->total_tgt_rows;
    CK(ctx->d_inc.ensure(ctx->inc_h.size() + 1));
    CK(ctx->d_inc_off.ensure((size_t)V + 1));
    if (!ctx->inc_h.empty())
        CK(cudaMemcpyAsync(ctx->d_inc.p, ctx->inc_h.data(), ctx->inc_h.size() * sizeof(IncDev), cudaMemcpyHostToDevice,
                           st));
    CK(cudaMemcpyAsync(ctx->d_inc_off.p, ctx->inc_off_h.data(), ((size_t)V + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(ctx->d_inv_cap.ensure(TR + 1));
    CK(ctx->d_inv_fill.ensure(TR + 1));
    CK(ctx->d_inv_off.ensure(TR + 2));
    CK(ctx->d_inv_ent.ensure(F + 1));
    CK(ctx->d_fwd_row.ensure(F + 1));
    CK(ctx->d_fwd_score.ensure(F + 1));
    CK(cudaMemsetAsync(ctx->d_inv_cap.p, 0, (TR + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_inv_fill.p, 0, (TR + 1) * 4, st));
    CK(ctx->d_scan.ensure(scan_scratch_words((uint32_t)std::max<size_t>(TR, S) + 2) + 64));
    CK(ctx->d_L_ub.ensure((size_t)S + 1));
    CK(ctx->d_L_off.ensure((size_t)S + 2));
    CK(ctx->d_L_cnt.ensure((size_t)S + 1));
    CK(ctx->d_stats.ensure(k3_wf_stats_bytes()));
    CK(cudaMemsetAsync(ctx->d_stats.p, 0, k3_wf_stats_bytes(), st));

    // pre-pass: row of every forward record, the (static) inverse-match slots, the potential lists
    ctx->cnt.gpu_launches += launch_k3_inv_capacity(ctx->d_pairs.p, P, ctx->total_rows, ctx->d_fwd_off.p,
                                                    ctx->d_fwd_cnt.p, ctx->d_fwd_rec.p, ctx->d_fwd_row.p,
                                                    ctx->d_inv_cap.p, ctx->slice_view[ctx->rank],
                                                    ctx->slice_view[ctx->rank + 1], st);
    ctx->cnt.gpu_launches +=
        launch_scan_u32(ctx->d_inv_cap.p, ctx->d_inv_off.p, (uint32_t)TR, ctx->d_scan.p, ctx->d_scan.cap, st);
    ctx->cnt.gpu_launches += launch_records(ctx);
    ctx->cnt.gpu_launches +=
        launch_k3_list_capacity(ctx->d_views.p, ctx->d_seg_view.p, S, ctx->d_inc.p, ctx->d_inc_off.p, ctx->d_pairs.p,
                                ctx->d_fwd_cnt.p, ctx->d_inv_cap.p, ctx->d_L_ub.p, ctx->d_stats.p, st);
    ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_L_ub.p, ctx->d_L_off.p, S, ctx->d_scan.p, ctx->d_scan.cap, st);

    // potential lists: forward records of the view's own pairs + of the pairs pointing at it
    std::vector<uint64_t>& cap = ctx->L_cap_h;
    cap.assign(V, 0);
    for (uint32_t p = 0; p < P; ++p) {
        cap[ctx->pairs[p].src] += ctx->pairs[p].fwd_total;
        if (ctx->pairs_h[p].emit_inverse) cap[ctx->pairs[p].tgt] += ctx->pairs[p].fwd_total;
    }
    const bool keep = ctx->prm.keep_scored != 0;
    ctx->L_base_h.assign(V + 1, 0);
    uint64_t Lcap = 0;
    for (uint32_t v = 0; v < V; ++v) {
        ctx->L_base_h[v] = Lcap;  // == L_off[first segment of v]
        Lcap += cap[v];
    }
    ctx->L_total = Lcap;
    if (Lcap > 0xfffffff0ull || 2 * F > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "match lists too large");
    CK(ctx->d_L_f.ensure(Lcap + 1));
    CK(ctx->d_L_meta.ensure(Lcap + 1));
    CK(ctx->d_L_score.ensure(Lcap + 1));
    CK(cudaMemsetAsync(ctx->d_L_score.p, 0, (Lcap + 1) * 4, st));
    if (keep) CK(ctx->d_L_rec.ensure(Lcap + 1));
    // rows are staged in shared memory up to maxm entries (an inverse block is not bounded by kNN:
    // any number of source segments may match the same target segment): read the longest list back
    // (the value is kept while the same committed scene is re-matched: a longer row than expected
    // makes the build kernel flag the program store as overflowed and the retry reads the new maximum)
    uint32_t maxm = (uint32_t)k3_max_staged();
    bool big_rows = true;
    if (ctx->k3_list_max == 0 || ctx->force_list_max) {
        CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_DEVMAX), ctx->d_stats.p + 36, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t dev_max = *ctx->rb_at<uint32_t>(l3d_ctx::RB_DEVMAX);  // WfStats::max_list
        ctx->k3_list_max = std::max(dev_max, 1u);
        ctx->force_list_max = false;
    }
    big_rows = ctx->k3_list_max > maxm;
    maxm = std::max<uint32_t>(std::min(ctx->k3_list_max, maxm), 1u);
    if (const char* ov = getenv("L3D_K3_MAXM")) {  // test hook: force the long-row path
        maxm = (uint32_t)std::max(1, atoi(ov));
        big_rows = true;
    }
    if (big_rows) {
        CK(ctx->d_L_sib.ensure((Lcap + 1) * k3_sib_bytes()));
        CK(ctx->d_L_dir.ensure(3 * (Lcap + 1)));
        CK(ctx->d_L_reg.ensure(Lcap + 1));
        CK(ctx->d_L_c.ensure(Lcap + S + 2));
        CK(ctx->d_L_h.ensure(Lcap + S + 2));
    }
    if (getenv("L3D_K3_DEBUG")) {  // histogram of the potential-list lengths (debugging aid)
        std::vector<uint32_t> ub(S);
        CK(cudaMemcpyAsync(ub.data(), ctx->d_L_ub.p, (size_t)S * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t edges[] = {0, 16, 32, 64, 128, 256, 512, 1024, 2048, 0xffffffffu};
        uint64_t n[10] = {0}, sm[10] = {0}, sq[10] = {0};
        for (uint32_t g = 0; g < S; ++g) {
            int b = 0;
            while (ub[g] > edges[b]) ++b;
            n[b]++; sm[b] += ub[g]; sq[b] += (uint64_t)ub[g] * ub[g];
        }
        for (int b = 0; b < 10; ++b)
            if (n[b]) fprintf(stderr, "[k3] len<=%u rows %llu sum %llu sumsq %llu\n", edges[b], (unsigned long long)n[b],
                              (unsigned long long)sm[b], (unsigned long long)sq[b]);
    }
    ctx->k3_maxm = maxm;
    ctx->k3_big_rows = big_rows;
    ctx->filt_cap = 2 * F + 1;
    CK(ctx->d_filt_rec.ensure(ctx->filt_cap));
    CK(ctx->d_filt_off.ensure((size_t)S + 1));
    CK(ctx->d_filt_cnt.ensure((size_t)S + 1));
    CK(ctx->d_view_max.ensure((size_t)V + 1));
    CK(cudaMemsetAsync(ctx->d_view_max.p, 0, ((size_t)V + 1) * 4, st));
    CK(ctx->d_small.ensure(16));
    CK(cudaMemsetAsync(ctx->d_small.p, 0, 16 * 4, st));
    CK(ctx->d_entries.ensure((size_t)S + 1));
    CK(ctx->d_prog_off.ensure((size_t)S + 1));
    CK(ctx->d_prog_nh.ensure((size_t)S + 1));
    const uint32_t my_rows = ctx->slice_g[ctx->rank + 1] - ctx->slice_g[ctx->rank];
    CK(ctx->d_k3_cls.ensure(k3_class_words(my_rows)));
    if (ctx->prog_cap == 0) ctx->prog_cap = std::max<uint64_t>(24ull * my_rows + 1024, 4096);
    if (ctx->prog_cap > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "fold programs too large");
    CK(ctx->d_prog.ensure(ctx->prog_cap * 16));
    int cuerr = 0;
    const int nl = launch_k3_build(k3_tables(ctx), st, &cuerr);
    if (nl < 0) return fail(L3D_ERR_CUDA, "launch of the build kernel failed: %s", cudaGetErrorString((cudaError_t)cuerr));
    ctx->cnt.gpu_launches += nl;
    ctx->tm.end(ev, st);
    ctx->stage3_phase = 1;
    return L3D_OK;
}

// the fold programs did not fit: grow the store and build again (the record kernel resets the
// scores and inverse slots the aborted pass left behind)
int score_rebuild(l3d_ctx* ctx, uint32_t needed_units)
{
    cudaStream_t st = ctx->stream;
    const uint32_t V = (uint32_t)ctx->views.size();
    {   // the longest list may have grown past the cached value: re-read it and re-size the staging
        uint32_t dev_max = 0;
        CK(cudaMemcpyAsync(&dev_max, ctx->d_stats.p + 36, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (dev_max > ctx->k3_list_max) {
            ctx->k3_list_max = dev_max;
            const uint32_t cap = (uint32_t)k3_max_staged();
            ctx->k3_maxm = std::min(dev_max, cap);
            if (dev_max > cap && !ctx->k3_big_rows) {
                const uint64_t Lcap = ctx->L_total;
                CK(ctx->d_L_sib.ensure((Lcap + 1) * k3_sib_bytes()));
                CK(ctx->d_L_dir.ensure(3 * (Lcap + 1)));
                CK(ctx->d_L_reg.ensure(Lcap + 1));
                CK(ctx->d_L_c.ensure(Lcap + ctx->S + 2));
                CK(ctx->d_L_h.ensure(Lcap + ctx->S + 2));
                ctx->k3_big_rows = true;
            }
        }
    }
    ctx->prog_cap = std::max<uint64_t>(2 * ctx->prog_cap, (uint64_t)needed_units + 1024);
    if (ctx->prog_cap > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "fold programs too large");
    CK(ctx->d_prog.ensure(ctx->prog_cap * 16));
    cudaEvent_t ev = ctx->tm.begin(L3D_T_SCORE, st);
    CK(cudaMemsetAsync(ctx->d_stats.p, 0, k3_wf_stats_bytes(), st));
    CK(cudaMemsetAsync(ctx->d_inv_fill.p, 0, ((size_t)ctx->total_tgt_rows + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_L_score.p, 0, (ctx->L_total + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_view_max.p, 0, ((size_t)V + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_small.p, 0, 16 * 4, st));
    ctx->prog_all = nullptr;
    ctx->cnt.gpu_launches += launch_records(ctx);
    int cuerr = 0;
    const int nl = launch_k3_build(k3_tables(ctx), st, &cuerr);
    if (nl < 0) return fail(L3D_ERR_CUDA, "launch of the build kernel failed: %s", cudaGetErrorString((cudaError_t)cuerr));
    ctx->cnt.gpu_launches += nl;
    ctx->tm.end(ev, st);
    return L3D_OK;
}

// fold (all rows) + finish (rows of the slice)
int l3d_score_fold(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 1 || ctx->stage3_phase < 1) return fail(L3D_ERR_STATE, "l3d_score_build has not run");
    if (ctx->world > 1 && !ctx->prog_all) return fail(L3D_ERR_STATE, "the fold programs have not been exchanged");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    cudaEvent_t ev = ctx->tm.begin(L3D_T_SCORE, st);
    int cuerr = 0;
    const K3Tables t = k3_tables(ctx);
    int nl = launch_k3_fold(t, st, &cuerr);
    if (nl < 0) return fail(L3D_ERR_CUDA, "launch of the fold kernel failed: %s", cudaGetErrorString((cudaError_t)cuerr));
    ctx->cnt.gpu_launches += nl;
    ctx->cnt.gpu_launches += launch_k3_finish(t, st);
    ctx->tm.end(ev, st);
    ctx->stage3_phase = 2;
    return L3D_OK;
}

// hypothesis index, per-view median depths, counters: needs the hypotheses of every view
int score_hypotheses_ready(l3d_ctx* ctx, bool* prog_overflow, uint32_t* prog_needed)
{
    cudaStream_t st = ctx->stream;
    const uint32_t V = (uint32_t)ctx->views.size(), S = ctx->S;
    cudaEvent_t ev = ctx->tm.begin(L3D_T_SCORE, st);
    // estimated_position3D_ index (canonical order = global segment order) and median depths
    CK(ctx->d_has.ensure((size_t)S + 1));
    CK(ctx->d_entry_idx.ensure((size_t)S + 2));
    ctx->cnt.gpu_launches += launch_k4_has(ctx->d_entries.p, S, ctx->d_has.p, st);
    ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_has.p, ctx->d_entry_idx.p, S, ctx->d_scan.p, ctx->d_scan.cap, st);
    ctx->cnt.gpu_launches += launch_k4_median(ctx->d_views.p, V, ctx->d_entries.p, ctx->d_small.p + 2, st);
    // read-backs into the pinned scratch (ctx.h); the view table falls back to pageable memory when it is large
    uint32_t* small = ctx->rb_at<uint32_t>(l3d_ctx::RB_SMALL);
    unsigned char* acc = ctx->rb_at<unsigned char>(l3d_ctx::RB_STATS);
    static_assert(l3d_ctx::RB_NEDGES - l3d_ctx::RB_STATS >= 256, "stats slot");
    if (k3_wf_stats_bytes() > 256) return fail(L3D_ERR_STATE, "internal: stats larger than their read-back slot");
    std::vector<ViewDev> vd_pageable;
    ViewDev* vd = ctx->rb_at<ViewDev>(l3d_ctx::RB_BIG);
    if (!ctx->rb_fits((size_t)V * sizeof(ViewDev))) {
        vd_pageable.resize(V);
        vd = vd_pageable.data();
    }
    CK(cudaMemcpyAsync(small, ctx->d_small.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NENT), ctx->d_entry_idx.p + S, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(acc, ctx->d_stats.p, k3_wf_stats_bytes(), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(vd, ctx->d_views.p, V * sizeof(ViewDev), cudaMemcpyDeviceToHost, st));
    ctx->tm.end(ev, st);
    if (ctx->ev_total3) ctx->tm.end(ctx->ev_total3, st);
    ctx->ev_total3 = nullptr;
    CK(cudaStreamSynchronize(st));
    const uint32_t n_entries = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NENT);
    const unsigned long long* a64 = (const unsigned long long*)acc;
    const uint32_t* a32 = (const uint32_t*)(acc + 16);
    if (prog_overflow) {
        *prog_overflow = (a32[2] & 4u) != 0;
        *prog_needed = a32[3];
        if (*prog_overflow) return L3D_OK;
    }
    ctx->tm.collect();
    if (ctx->world > 1) {  // the slices' counters were summed when the hypotheses were imported
        ctx->cnt.sim_evals = ctx->shard_sim_evals;
        ctx->cnt.scored_entries = ctx->shard_scored;
        ctx->cnt.filtered_entries = ctx->shard_filtered;
    } else {
        ctx->cnt.sim_evals = a64[0];
        ctx->cnt.scored_entries = a64[1];
        ctx->cnt.filtered_entries = a32[1];  // filt_cursor
    }
    if (a32[2] & 2u) return fail(L3D_ERR_CAPACITY, "filtered-match store overflow");
    if (a32[2] & 8u) return fail(L3D_ERR_STATE, "internal: scoring dependency wait timed out");
    if (small[2]) return fail(L3D_ERR_CAPACITY, "more than 8192 hypotheses in one view (median-depth kernel)");
    ctx->cnt.num_entries = n_entries;
    for (uint32_t v = 0; v < V; ++v) {
        ctx->views[v].median_depth = vd[v].median_depth;
        ctx->views[v].median_sigma = ctx->views[v].k * vd[v].median_depth;  // view.h:122-135
    }
    // update_Matches_and_Estimated_position3D (src/line3D.cc:1857-1908) re-triangulates the best
    // matches with unchanged poses: the identical call, hence identical depths -- nothing to do.
    leave_translated(ctx);  // untranslate() (src/line3D.cc:637); a no-op in raw mode
    ctx->stage = 2;
    ctx->stage3_phase = 3;
    return L3D_OK;
}

int l3d_match_stage3(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->world > 1)
        return fail(L3D_ERR_STATE, "sharded run: use l3d_score_build / l3d_shard_* / l3d_score_fold");
    int rc = l3d_score_build(ctx);
    if (rc) return rc;
    for (int attempt = 0;; ++attempt) {
        rc = l3d_score_fold(ctx);
        if (rc) return rc;
        bool overflow = false;
        uint32_t needed = 0;
        rc = score_hypotheses_ready(ctx, &overflow, &needed);
        if (rc) return rc;
        if (!overflow) return L3D_OK;
        if (attempt >= 8) return fail(L3D_ERR_CAPACITY, "fold program store overflow");
        rc = score_rebuild(ctx, needed);
        if (rc) return rc;
    }
}

int l3d_match_images(l3d_ctx* ctx, const l3d_params* params)
{
    if (ctx && ctx->stream_mode) return stream_match_images(ctx, params);
    int rc = l3d_match_stage12(ctx, params);
    if (rc) return rc;
    return l3d_match_stage3(ctx);  // stage timers keep accumulating until the next stage12
}

// ------------------------------------------------------------------------------------------
// stage 4: affinity (edges of this rank's slice, then ids over all edges) + clustering
// ------------------------------------------------------------------------------------------
int l3d_affinity_edges(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 2) return fail(L3D_ERR_STATE, "l3d_match_images has not run");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t S = ctx->S;
    ctx->cnt.num_edges = 0;
    ctx->cnt.num_local_ids = 0;
    ctx->cluster_ids.clear();
    ctx->n_edges_local = 0;
    ctx->n_edges_all = 0;
    ctx->edges_all = nullptr;
    if (ctx->cnt.num_entries == 0) {  // "no clusterable segments" (src/line3D.cc:2028-2034)
        ctx->stage = 3;
        ctx->stage4_phase = 1;
        return L3D_OK;
    }
    // translate() again (src/line3D.cc:2065); only the camera centres move
    enter_translated(ctx);
    // the cluster -> 3-D line tail (l3d_lines3D) runs in this frame too (src/line3D.cc:2115-2140): keep the cameras
    ctx->lines_ready = false;
    ctx->tail_translation = ctx->translation;
    ctx->tail_views.resize(ctx->views.size());
    for (size_t v = 0; v < ctx->views.size(); ++v) {
        const HostView& hv = ctx->views[v];
        TailView& tv = ctx->tail_views[v];
        memcpy(tv.K, hv.cam.K.m, sizeof(tv.K));
        memcpy(tv.R, hv.cam.R.m, sizeof(tv.R));
        tv.t[0] = hv.cam.t.x; tv.t[1] = hv.cam.t.y; tv.t[2] = hv.cam.t.z;
        tv.C[0] = hv.cam.C.x; tv.C[1] = hv.cam.C.y; tv.C[2] = hv.cam.C.z;
        // View::View (src/view.cc:28-33): diagonal_ = sqrtf(float(w*w + h*h)), min_line_length_ = diagonal_ * 0.005f
        const float diag = sqrtf((float)(hv.v.width * hv.v.width + hv.v.height * hv.v.height));
        tv.min_line_length = diag * 0.005f;
        tv.cam_id = hv.v.cam_id;
    }
    // median scene depth of the lines (src/line3D.cc:2074-2091)
    std::vector<float> sd;
    for (auto& hv : ctx->views)
        if (hv.median_depth > 1e-12 && hv.current) sd.push_back(hv.median_depth);  // views in view_order_ only
    if (!sd.empty()) {
        std::sort(sd.begin(), sd.end());
        ctx->med_scene_depth_lines = sd[sd.size() / 2];
    } else
        ctx->med_scene_depth_lines = 0.0f;

    ctx->ev_total4 = ctx->tm.begin(L3D_T_TOTAL, st);
    cudaEvent_t ev = ctx->tm.begin(L3D_T_AFFINITY, st);
    const ListRec* filt = ctx->filt_all ? ctx->filt_all : ctx->d_filt_rec.p;
    // stream mode: the filtered lists of a cycle live at per-view bases inside an arena of st_f_extent records
    const size_t nf = ctx->stream_mode ? (size_t)ctx->st_f_extent : (size_t)ctx->cnt.filtered_entries;
    const uint32_t g_lo = ctx->slice_g[ctx->rank], g_hi = ctx->slice_g[ctx->rank + 1];
    if (ctx->stream_mode) {  // growing tables: reallocate rarely (ctx.h: ensure_roomy)
        CK(ensure_roomy(ctx->d_filt_sim, nf + 1, (size_t)1 << 21));
        CK(ensure_roomy(ctx->d_E_cnt, (size_t)S + 1, (size_t)1 << 18));
        CK(ensure_roomy(ctx->d_E_off, (size_t)S + 2, (size_t)1 << 18));
        CK(ensure_roomy(ctx->d_edges, (nf + 1) * k4_edge_bytes(), ((size_t)1 << 21) * k4_edge_bytes()));
        CK(ensure_roomy(ctx->d_first_touch, (size_t)S + 1, (size_t)1 << 18));
    }
    CK(ctx->d_filt_sim.ensure(nf + 1));
    CK(ctx->d_E_cnt.ensure((size_t)S + 1));
    CK(ctx->d_E_off.ensure((size_t)S + 2));
    CK(cudaMemsetAsync(ctx->d_E_cnt.p, 0, ((size_t)S + 1) * 4, st));
    CK(ctx->d_tests.ensure(2));
    CK(cudaMemsetAsync(ctx->d_tests.p, 0, 16, st));
    ctx->cnt.gpu_launches +=
        launch_k4_edges_count(ctx->d_views.p, ctx->d_seg_view.p, ctx->d_entries.p, S, ctx->d_filt_off.p,
                              ctx->d_filt_cnt.p, filt, ctx->two_sigA_sqr, ctx->med_scene_depth_lines,
                              ctx->d_filt_sim.p, ctx->d_E_cnt.p, ctx->d_tests.p, g_lo, g_hi, st);
    ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_E_cnt.p, ctx->d_E_off.p, S, ctx->d_scan.p, ctx->d_scan.cap, st);
    // every filtered list entry yields at most one edge: size the store from that bound and read the
    // count back while the write kernel runs
    CK(ctx->d_edges.ensure((nf + 1) * k4_edge_bytes()));
    CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NEDGES), ctx->d_E_off.p + S, 4, cudaMemcpyDeviceToHost, st));
    ctx->cnt.gpu_launches +=
        launch_k4_edges_write(ctx->d_views.p, S, ctx->d_filt_off.p, ctx->d_filt_cnt.p, filt, ctx->d_filt_sim.p,
                              ctx->d_E_off.p, ctx->d_edges.p, g_lo, g_hi, st);
    ctx->tm.end(ev, st);
    CK(cudaStreamSynchronize(st));
    const uint32_t n_edges = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NEDGES);
    ctx->n_edges_local = n_edges;
    ctx->stage4_phase = 1;
    return L3D_OK;
}

int l3d_affinity_ids(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 2 || ctx->stage4_phase < 1) return fail(L3D_ERR_STATE, "l3d_affinity_edges has not run");
    if (ctx->stage == 3 && ctx->cnt.num_entries == 0) return L3D_OK;
    if (ctx->world > 1 && !ctx->edges_all && ctx->stage4_phase < 2)
        return fail(L3D_ERR_STATE, "the edges have not been exchanged");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool roomy = ctx->stream_mode;  // see run_stage12_batches
    auto grow = [roomy](auto& buf, size_t n) { return roomy ? ensure_roomy(buf, n, (size_t)1 << 20) : buf.ensure(n); };
    const uint32_t S = ctx->S;
    const uint32_t n_edges = ctx->world > 1 ? ctx->n_edges_all : ctx->n_edges_local;
    const void* edges = ctx->world > 1 ? ctx->edges_all : (const void*)ctx->d_edges.p;
    cudaEvent_t ev = ctx->tm.begin(L3D_T_AFFINITY, st);
    uint32_t n_local = 0;
    if (n_edges) {
        CK(grow(ctx->d_first_touch, (size_t)S + 1));
        CK(cudaMemsetAsync(ctx->d_first_touch.p, 0xff, ((size_t)S + 1) * 4, st));
        CK(grow(ctx->d_flags, 2 * (size_t)n_edges + 1));
        CK(grow(ctx->d_flag_scan, 2 * (size_t)n_edges + 2));
        CK(grow(ctx->d_A_ij, 2 * (size_t)n_edges));
        CK(grow(ctx->d_A_w, 2 * (size_t)n_edges));
        CK(ctx->d_l2g.ensure(2 * (size_t)n_edges));
        CK(grow(ctx->d_scan, scan_scratch_words(2 * n_edges + 2) + 64));
        ctx->cnt.gpu_launches +=
            launch_k4_ids(edges, n_edges, ctx->d_first_touch.p, ctx->d_flags.p, ctx->d_flag_scan.p, ctx->d_scan.p,
                          ctx->d_scan.cap, ctx->d_A_ij.p, ctx->d_A_w.p, ctx->d_l2g.p, st);
        CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NLOCAL), ctx->d_flag_scan.p + 2 * (size_t)n_edges, 4,
                           cudaMemcpyDeviceToHost, st));
    }
    ctx->tm.end(ev, st);
    if (ctx->ev_total4) ctx->tm.end(ctx->ev_total4, st);
    ctx->ev_total4 = nullptr;
    CK(cudaStreamSynchronize(st));
    ctx->tm.collect();
    ctx->cnt.num_edges = 2 * n_edges;  // A_ holds both directions
    if (n_edges) n_local = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NLOCAL);
    ctx->cnt.num_local_ids = n_local;
    leave_translated(ctx);  // untranslate() (src/line3D.cc:2140)
    ctx->stage = 3;
    ctx->stage4_phase = 2;
    return L3D_OK;
}

int l3d_affinity(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->world > 1)
        return fail(L3D_ERR_STATE, "sharded run: use l3d_affinity_edges / l3d_shard_* / l3d_affinity_ids");
    int rc = l3d_affinity_edges(ctx);
    if (rc) return rc;
    return l3d_affinity_ids(ctx);
}

// drop-in for L3DPP::find_collinear_segments_GPU (include/cudawrapper.h:84-86): host arrays in and out,
// blocking.  buffer[r * row_stride_bytes + c] = 1 iff segment c is collinear to segment r.
int l3d_find_collinear(l3d_ctx* ctx, const float* lines_xyxy, uint32_t n, float dist_t, char* buffer,
                       uint64_t row_stride_bytes)
{
    if (!ctx || (n && (!lines_xyxy || !buffer))) return fail(L3D_ERR_ARG, "NULL argument");
    if (n == 0) return L3D_OK;
    if (row_stride_bytes < n) return fail(L3D_ERR_ARG, "row stride %llu < %u", (unsigned long long)row_stride_bytes, n);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->d_collin_lines.ensure(n));  // own staging: the resident scene / stream tables stay untouched
    CK(ctx->d_collin.ensure((size_t)n * n));
    CK(cudaMemcpyAsync(ctx->d_collin_lines.p, lines_xyxy, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
    ctx->cnt.gpu_launches += launch_k5_collinear(ctx->d_collin_lines.p, n, dist_t, ctx->d_collin.p, n, st);
    CK(cudaMemcpy2DAsync(buffer, row_stride_bytes, ctx->d_collin.p, n, n, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return L3D_OK;
}

// SparseMatrix::SparseMatrix (src/sparsematrix.cc:8-61) of A_, built on the device and kept there for a
// device-side consumer (l3d_get_sparse_device); the host copies are optional.
int l3d_affinity_sparse(l3d_ctx* ctx, int sort_by_row, float normalization_factor, float* entries_xyzw,
                        int32_t* start_indices, uint32_t cap_entries, uint32_t cap_rows)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 3) return fail(L3D_ERR_STATE, "l3d_affinity has not run");
    if (!(normalization_factor != 0.0f)) return fail(L3D_ERR_ARG, "normalization_factor must not be 0");
    const uint32_t E = ctx->cnt.num_edges, n = ctx->cnt.num_local_ids;
    ctx->sparse_ready = false;
    if ((entries_xyzw && cap_entries < E) || (start_indices && cap_rows < n))
        return fail(L3D_ERR_CAPACITY, "need %u entries and %u start indices", E, n);
    if (!E || !n) return L3D_OK;  // SparseMatrix with entries_ == NULL (src/sparsematrix.cc:18-19)
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(ctx->d_sp_hist.ensure((size_t)n + 2));
    CK(ctx->d_sp_off.ensure((size_t)n + 2));
    CK(ctx->d_sp_fill.ensure((size_t)n + 2));
    CK(ctx->d_sp_tmp.ensure(E));
    CK(ctx->d_sp_entries.ensure(E));
    CK(ctx->d_sp_start.ensure(n));
    CK(ctx->d_scan.ensure(scan_scratch_words(n + 2) + 64));
    ctx->cnt.gpu_launches += launch_k4_sparse(ctx->d_A_ij.p, ctx->d_A_w.p, E, n, sort_by_row ? 1 : 0, normalization_factor,
                                              ctx->d_sp_hist.p, ctx->d_sp_off.p, ctx->d_sp_fill.p, ctx->d_sp_tmp.p,
                                              ctx->d_scan.p, ctx->d_scan.cap, ctx->d_sp_entries.p, ctx->d_sp_start.p, st);
    if (entries_xyzw)
        CK(cudaMemcpyAsync(entries_xyzw, ctx->d_sp_entries.p, (size_t)E * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (start_indices)
        CK(cudaMemcpyAsync(start_indices, ctx->d_sp_start.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->sparse_ready = true;
    return L3D_OK;
}

int l3d_get_sparse_device(l3d_ctx* ctx, const void** entries_float4, const void** start_indices_int)
{
    if (!ctx || !entries_float4 || !start_indices_int) return fail(L3D_ERR_ARG, "NULL argument");
    if (!ctx->sparse_ready) return fail(L3D_ERR_STATE, "l3d_affinity_sparse has not run (or A_ is empty)");
    *entries_float4 = ctx->d_sp_entries.p;
    *start_indices_int = ctx->d_sp_start.p;
    return L3D_OK;
}

// Host-only: the visual neighbours Line3D::matchImages would choose from world-point lists
// (Line3D::findVisualNeighborsFromWPs, src/line3D.cc:723-843, after Line3D::translate).  No device needed.
int l3d_neighbors_from_worldpoints(const l3d_view* views, uint32_t n_views, const uint32_t* wps_concat,
                                   const uint32_t* wp_counts, uint32_t num_neighbors, uint32_t* out_cam_ids,
                                   uint32_t* out_counts)
{
    if (!views || !wps_concat || !wp_counts || !out_cam_ids || !out_counts) return fail(L3D_ERR_ARG, "NULL argument");
    // Line3D::matchImages raises num_neighbors to 2 (src/line3D.cc:531-536); the output rows are laid out with the
    // caller's stride, so a smaller value cannot hold what would be chosen
    if (num_neighbors < 2) return fail(L3D_ERR_ARG, "num_neighbors must be >= 2 (got %u)", num_neighbors);
    {
        std::set<uint32_t> ids;
        for (uint32_t i = 0; i < n_views; ++i)
            if (!ids.insert(views[i].cam_id).second) return fail(L3D_ERR_ARG, "camera ID [%u] already in use!", views[i].cam_id);
    }
    l3d_ctx t;
    size_t no = 0;
    for (uint32_t i = 0; i < n_views; ++i) {
        HostView hv;
        hv.v = views[i];
        hv.wps.assign(wps_concat + no, wps_concat + no + wp_counts[i]);
        no += wp_counts[i];
        hv.cam.init(views[i].K, views[i].R, views[i].t);
        t.views.push_back(std::move(hv));
    }
    std::sort(t.views.begin(), t.views.end(), [](const HostView& a, const HostView& b) { return a.v.cam_id < b.v.cam_id; });
    compute_translation(&t);
    apply_translation(&t, -1.0);
    std::vector<const hg::Camera*> cams(n_views);
    std::vector<float> md(n_views, 0.0f);  // View::median_depth_ starts at 0 (src/view.cc:13)
    std::vector<std::vector<uint32_t>> wps(n_views), nb;
    for (uint32_t v = 0; v < n_views; ++v) {
        cams[v] = &t.views[v].cam;
        wps[v] = t.views[v].wps;
    }
    hg::visual_neighbors_from_worldpoints(cams, md, wps, (unsigned)num_neighbors, nb);
    // output rows follow the order of `views` as given
    for (uint32_t i = 0; i < n_views; ++i) {
        uint32_t v = 0;
        while (t.views[v].v.cam_id != views[i].cam_id) ++v;
        out_counts[i] = (uint32_t)nb[v].size();
        for (size_t k = 0; k < nb[v].size(); ++k) out_cam_ids[(size_t)i * num_neighbors + k] = t.views[nb[v][k]].v.cam_id;
    }
    return L3D_OK;
}

// neighbours of a view as used by the last l3d_match_images (camera ids, ascending)
int l3d_get_neighbors(l3d_ctx* ctx, uint32_t cam_id, uint32_t* out, uint32_t cap, uint32_t* count)
{
    if (!ctx || !count) return fail(L3D_ERR_ARG, "NULL argument");
    auto f = ctx->cam2view.find(cam_id);
    if (f == ctx->cam2view.end()) return fail(L3D_ERR_ARG, "unknown camera %u", cam_id);
    const std::vector<uint32_t>& nb = ctx->views[f->second].nb_views;
    *count = (uint32_t)nb.size();
    if (cap < nb.size()) return fail(L3D_ERR_CAPACITY, "need %zu ids", nb.size());
    for (size_t k = 0; k < nb.size(); ++k) out[k] = ctx->views[nb[k]].v.cam_id;
    return L3D_OK;
}

// Felzenszwalb-Huttenlocher clustering, src/clustering.cc:7-48 + include/universe.h:59-117
int l3d_cluster_edges(const int32_t* ij, const float* w, uint32_t ne, uint32_t n, int32_t* out)
{
    if ((ne && (!ij || !w)) || (n && !out)) return fail(L3D_ERR_ARG, "NULL argument");
    struct El {
        int rank, id, size;
    };
    std::vector<El> el(n);
    for (uint32_t i = 0; i < n; ++i) el[i] = El{0, (int)i, 1};
    auto find = [&](int x) {
        int y = x;
        while (y != el[y].id) y = el[y].id;
        el[x].id = y;
        return y;
    };
    if (ne) {
        std::vector<uint32_t> order(ne);
        for (uint32_t i = 0; i < ne; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return w[a] < w[b]; });
        const float c = 3.0f;
        std::vector<float> thr(n, c);
        for (uint32_t o : order) {
            const int i = ij[2 * o], j = ij[2 * o + 1];
            if (i < 0 || j < 0 || (uint32_t)i >= n || (uint32_t)j >= n) return fail(L3D_ERR_ARG, "edge out of range");
            int a = find(i), b = find(j);
            if (a != b && w[o] <= thr[a] && w[o] <= thr[b]) {
                if (el[a].rank > el[b].rank) {
                    el[b].id = a;
                    el[a].size += el[b].size;
                } else {
                    el[a].id = b;
                    el[b].size += el[a].size;
                    if (el[a].rank == el[b].rank) el[b].rank++;
                }
                a = find(a);
                thr[a] = w[o] + c / (float)el[a].size;
            }
        }
    }
    for (uint32_t i = 0; i < n; ++i) out[i] = find((int)i);
    return L3D_OK;
}

int l3d_cluster(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 3) return fail(L3D_ERR_STATE, "l3d_affinity has not run");
    const uint32_t ne = ctx->cnt.num_edges, n = ctx->cnt.num_local_ids;
    ctx->cluster_ids.assign(n, 0);
    ctx->cnt.num_clusters = 0;
    if (ne == 0) {
        ctx->stage = 4;
        return L3D_OK;
    }
    std::vector<int32_t> ij(2 * (size_t)ne);
    std::vector<float> w(ne);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ij.data(), ctx->d_A_ij.p, (size_t)ne * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(w.data(), ctx->d_A_w.p, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = l3d_cluster_edges(ij.data(), w.data(), ne, n, ctx->cluster_ids.data());
    if (rc) return rc;
    std::set<int32_t> uniq(ctx->cluster_ids.begin(), ctx->cluster_ids.end());
    ctx->cnt.num_clusters = (uint32_t)uniq.size();
    ctx->stage = 4;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// cluster -> 3-D line tail (k6_lines3d.cu)
// ------------------------------------------------------------------------------------------
// replaces the rest of Line3D::reconstruct3Dlines after the clustering (src/line3D.cc:2115-2141):
// clusterSegments' bookkeeping (:2502-2575: segments per cluster root in ascending local id, clusters in order of
// first appearance, kept if seen by >= visibility_t cameras) on the host, then get3DlineFromCluster,
// findCollinearSegments_return, filterTinySegments and the translation back, one thread per cluster on the device
int l3d_lines3D(l3d_ctx* ctx, uint32_t visibility_t)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 4) return fail(L3D_ERR_STATE, "l3d_cluster has not run");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->l3_seg_off.assign(1, 0);
    ctx->l3_res_off.assign(1, 0);
    ctx->l3_res.clear();
    ctx->l3_ref_cam.clear();
    ctx->l3_segs.clear();
    ctx->lines_ready = true;
    const uint32_t n = ctx->cnt.num_local_ids;
    if (n == 0 || ctx->cnt.num_edges == 0) return L3D_OK;
    visibility_t = std::max(visibility_t, 3u);  // src/line3D.cc:2040
    // local id -> global segment
    std::vector<uint32_t> g(n);
    CK(cudaMemcpyAsync(g.data(), ctx->d_l2g.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    auto view_of = [&](uint32_t gs) {
        uint32_t lo = 0, hi = (uint32_t)ctx->views.size();
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) / 2;
            if (ctx->views[mid].seg_off <= gs) lo = mid; else hi = mid;
        }
        return lo;
    };
    // clusters in order of first appearance, members in ascending local id
    std::vector<int32_t> slot_of(n, -1);
    std::vector<uint32_t> cnt;
    std::vector<int32_t> root_of_slot;
    for (uint32_t i = 0; i < n; ++i) {
        const int32_t root = ctx->cluster_ids[i];
        if (slot_of[root] < 0) {
            slot_of[root] = (int32_t)cnt.size();
            cnt.push_back(0);
            root_of_slot.push_back(root);
        }
        ++cnt[slot_of[root]];
    }
    std::vector<uint32_t> off(cnt.size() + 1, 0);
    for (size_t c = 0; c < cnt.size(); ++c) off[c + 1] = off[c] + cnt[c];
    std::vector<uint32_t> mem(n), fill(off.begin(), off.end() - 1);
    for (uint32_t i = 0; i < n; ++i) mem[fill[slot_of[ctx->cluster_ids[i]]]++] = g[i];
    // clusters seen by >= visibility_t cameras
    std::vector<uint32_t> v_off(1, 0), v_mem;
    std::set<uint32_t> cams;
    for (size_t c = 0; c < cnt.size(); ++c) {
        cams.clear();
        for (uint32_t k = off[c]; k < off[c + 1]; ++k) cams.insert(view_of(mem[k]));
        if (cams.size() < visibility_t) continue;
        v_mem.insert(v_mem.end(), mem.begin() + off[c], mem.begin() + off[c + 1]);
        v_off.push_back((uint32_t)v_mem.size());
    }
    const uint32_t ncl = (uint32_t)v_off.size() - 1, M = (uint32_t)v_mem.size();
    if (ncl == 0) return L3D_OK;
    CK(ctx->d_t_cl_off.ensure((size_t)ncl + 1));
    CK(ctx->d_t_members.ensure(M));
    CK(ctx->d_t_views.ensure(ctx->tail_views.size()));
    CK(ctx->d_t_L.ensure(6 * (size_t)M));
    CK(ctx->d_t_LC.ensure(6 * (size_t)M));
    CK(ctx->d_t_pts.ensure(6 * (size_t)M));
    CK(ctx->d_t_dist.ensure(2 * (size_t)M));
    CK(ctx->d_t_ord.ensure(2 * (size_t)M));
    CK(ctx->d_t_ok.ensure(M));
    CK(ctx->d_t_camtab.ensure(2 * (size_t)M));
    CK(ctx->d_t_out_n.ensure(ncl));
    CK(ctx->d_t_out_ref.ensure(ncl));
    CK(ctx->d_t_out_seg.ensure(6 * (size_t)M));
    CK(cudaMemcpyAsync(ctx->d_t_cl_off.p, v_off.data(), ((size_t)ncl + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_t_members.p, v_mem.data(), (size_t)M * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_t_views.p, ctx->tail_views.data(), ctx->tail_views.size() * sizeof(TailView),
                       cudaMemcpyHostToDevice, st));
    const double t3[3] = {ctx->tail_translation.x, ctx->tail_translation.y, ctx->tail_translation.z};
    ctx->cnt.gpu_launches +=
        launch_k6_lines3d(ncl, ctx->d_t_cl_off.p, ctx->d_t_members.p, ctx->d_entries.p, ctx->d_segs.p, ctx->d_rays.p,
                          ctx->d_seg_view.p, ctx->d_t_views.p, t3, ctx->d_t_L.p, ctx->d_t_LC.p, ctx->d_t_pts.p,
                          ctx->d_t_dist.p, ctx->d_t_ord.p, ctx->d_t_ok.p, ctx->d_t_camtab.p, ctx->d_t_out_n.p,
                          ctx->d_t_out_ref.p, ctx->d_t_out_seg.p, st);
    std::vector<uint32_t> out_n(ncl), out_ref(ncl);
    std::vector<double> out_seg(6 * (size_t)M);
    CK(cudaMemcpyAsync(out_n.data(), ctx->d_t_out_n.p, (size_t)ncl * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_ref.data(), ctx->d_t_out_ref.p, (size_t)ncl * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out_seg.data(), ctx->d_t_out_seg.p, 6 * (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // lines3D_: the clusters that kept at least one segment, in cluster order
    for (uint32_t c = 0; c < ncl; ++c) {
        if (!out_n[c]) continue;
        const size_t base = 6 * (size_t)v_off[c];
        ctx->l3_segs.insert(ctx->l3_segs.end(), out_seg.begin() + base, out_seg.begin() + base + 6 * (size_t)out_n[c]);
        ctx->l3_seg_off.push_back((uint32_t)(ctx->l3_segs.size() / 6));
        for (uint32_t k = v_off[c]; k < v_off[c + 1]; ++k) {
            const uint32_t v = view_of(v_mem[k]);
            ctx->l3_res.push_back(ctx->views[v].v.cam_id);
            ctx->l3_res.push_back(v_mem[k] - ctx->views[v].seg_off);
        }
        ctx->l3_res_off.push_back((uint32_t)(ctx->l3_res.size() / 2));
        ctx->l3_ref_cam.push_back(out_ref[c] == 0xffffffffu ? 0u : ctx->views[out_ref[c]].v.cam_id);
    }
    return L3D_OK;
}

int l3d_get_lines3D_counts(l3d_ctx* ctx, uint32_t* counts3)
{
    if (!ctx || !counts3) return fail(L3D_ERR_ARG, "NULL argument");
    if (!ctx->lines_ready) return fail(L3D_ERR_STATE, "l3d_lines3D has not run");
    counts3[0] = (uint32_t)ctx->l3_ref_cam.size();
    counts3[1] = (uint32_t)(ctx->l3_segs.size() / 6);
    counts3[2] = (uint32_t)(ctx->l3_res.size() / 2);
    return L3D_OK;
}

int l3d_get_lines3D(l3d_ctx* ctx, uint32_t* seg_off, double* segs6, uint32_t* res_off, uint32_t* res2, uint32_t* ref_cam)
{
    if (!ctx || !seg_off || !segs6 || !res_off || !res2 || !ref_cam) return fail(L3D_ERR_ARG, "NULL argument");
    if (!ctx->lines_ready) return fail(L3D_ERR_STATE, "l3d_lines3D has not run");
    memcpy(seg_off, ctx->l3_seg_off.data(), ctx->l3_seg_off.size() * 4);
    memcpy(res_off, ctx->l3_res_off.data(), ctx->l3_res_off.size() * 4);
    if (!ctx->l3_segs.empty()) memcpy(segs6, ctx->l3_segs.data(), ctx->l3_segs.size() * 8);
    if (!ctx->l3_res.empty()) memcpy(res2, ctx->l3_res.data(), ctx->l3_res.size() * 4);
    if (!ctx->l3_ref_cam.empty()) memcpy(ref_cam, ctx->l3_ref_cam.data(), ctx->l3_ref_cam.size() * 4);
    return L3D_OK;
}

// Line3D::save3DLinesAsTXT (src/line3D.cc:3122-3178): per line "k  k x (P1 P2)  r  r x (camID segID x1 y1 x2 y2)",
// std::ofstream default formatting
int l3d_save_lines3D_txt(l3d_ctx* ctx, const char* path)
{
    if (!ctx || !path) return fail(L3D_ERR_ARG, "NULL argument");
    if (!ctx->lines_ready) return fail(L3D_ERR_STATE, "l3d_lines3D has not run");
    if (ctx->l3_ref_cam.empty()) return fail(L3D_ERR_STATE, "no 3D lines to save!");
    std::ofstream file(path);
    if (!file) return fail(L3D_ERR_ARG, "cannot open %s", path);
    for (size_t i = 0; i < ctx->l3_ref_cam.size(); ++i) {
        const uint32_t s0 = ctx->l3_seg_off[i], s1 = ctx->l3_seg_off[i + 1];
        file << (size_t)(s1 - s0) << " ";
        for (uint32_t s = s0; s < s1; ++s) {
            const double* p = &ctx->l3_segs[6 * (size_t)s];
            file << p[0] << " " << p[1] << " " << p[2] << " ";
            file << p[3] << " " << p[4] << " " << p[5] << " ";
        }
        const uint32_t r0 = ctx->l3_res_off[i], r1 = ctx->l3_res_off[i + 1];
        file << (size_t)(r1 - r0) << " ";
        for (uint32_t r = r0; r < r1; ++r) {
            const uint32_t cam = ctx->l3_res[2 * (size_t)r], seg = ctx->l3_res[2 * (size_t)r + 1];
            file << cam << " " << seg << " ";
            const HostView& hv = ctx->views[ctx->cam2view[cam]];
            // batch scenes keep their segments in the pinned staging of the commit, stream scenes per view
            const float* c = hv.segs.empty() ? (const float*)ctx->pinned + 4 * ((size_t)hv.seg_off + seg) : &hv.segs[4 * (size_t)seg];
            file << c[0] << " " << c[1] << " ";
            file << c[2] << " " << c[3] << " ";
        }
        file << std::endl;
    }
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------
int l3d_get_counts(l3d_ctx* ctx, l3d_counts* out)
{
    if (!ctx || !out) return fail(L3D_ERR_ARG, "NULL argument");
    *out = ctx->cnt;
    return L3D_OK;
}
int l3d_reset_counters(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    ctx->cnt.gpu_launches = 0;
    return L3D_OK;
}
int l3d_get_timings(l3d_ctx* ctx, float* ms, uint32_t n)
{
    if (!ctx || !ms) return fail(L3D_ERR_ARG, "NULL argument");
    for (uint32_t i = 0; i < n && i < L3D_T_COUNT; ++i) ms[i] = ctx->tm.ms[i];
    return L3D_OK;
}
int l3d_get_pairs(l3d_ctx* ctx, uint32_t* out, uint32_t cap)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (cap < ctx->pairs.size()) return fail(L3D_ERR_CAPACITY, "need %zu pairs", ctx->pairs.size());
    for (size_t p = 0; p < ctx->pairs.size(); ++p) {
        out[2 * p] = ctx->views[ctx->pairs[p].src].v.cam_id;
        out[2 * p + 1] = ctx->views[ctx->pairs[p].tgt].v.cam_id;
    }
    return L3D_OK;
}

int l3d_get_view_lists(l3d_ctx* ctx, uint32_t cam_id, int which, uint32_t* row_off, l3d_list_rec* recs, uint64_t cap,
                       uint64_t* out_count)
{
    if (!ctx || !row_off) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx->stage < 2) return fail(L3D_ERR_STATE, "l3d_match_images has not run");
    auto f = ctx->cam2view.find(cam_id);
    if (f == ctx->cam2view.end()) return fail(L3D_ERR_ARG, "unknown camera %u", cam_id);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t v = f->second;
    const HostView& hv = ctx->views[v];
    const uint32_t N = hv.v.num_segs;
    std::vector<uint32_t> off(N + 1, 0), cnt(N, 0);
    const ListRec* src = nullptr;
    uint64_t region = 0;
    if (which == 0 && ctx->stream_mode) {
        // stream mode: the cycle's working lists; entries to cameras deleted this cycle are flagged dead
        // (Line3D::updateMatch, src/line3D.cc:1016-1055) and skipped here
        CK(cudaMemcpyAsync(off.data(), ctx->d_L_off.p + hv.seg_off, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cnt.data(), ctx->d_L_cnt.p + hv.seg_off, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint32_t lo = 0xffffffffu, hi = 0;
        for (uint32_t i = 0; i < N; ++i)
            if (cnt[i]) {
                lo = std::min(lo, off[i]);
                hi = std::max(hi, off[i] + cnt[i]);
            }
        if (lo == 0xffffffffu) lo = hi = 0;
        std::vector<ListRec> tmp(hi - lo);
        if (hi > lo) {
            CK(cudaMemcpyAsync(tmp.data(), ctx->d_st_W_rec.p + lo, (size_t)(hi - lo) * sizeof(ListRec),
                               cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        uint64_t total = 0;
        for (uint32_t i = 0; i < N; ++i) {
            row_off[i] = (uint32_t)total;
            for (uint32_t e = 0; e < cnt[i]; ++e)
                if (!(tmp[(size_t)off[i] - lo + e].flags & 4u)) ++total;
        }
        row_off[N] = (uint32_t)total;
        if (out_count) *out_count = total;
        if (total > cap) return fail(L3D_ERR_CAPACITY, "need %llu records", (unsigned long long)total);
        if (total == 0) return L3D_OK;
        if (!recs) return fail(L3D_ERR_ARG, "recs is NULL");
        uint64_t n = 0;
        for (uint32_t i = 0; i < N; ++i)
            for (uint32_t e = 0; e < cnt[i]; ++e) {
                const ListRec& L = tmp[(size_t)off[i] - lo + e];
                if (L.flags & 4u) continue;
                l3d_list_rec& o = recs[n++];
                o.tgt_cam = ctx->views[L.tgt_view].v.cam_id;
                o.tgt_seg = L.tgt_seg;
                o.overlap_score = L.overlap;
                o.score3D = L.score;
                o.depth_p1 = L.d_p1;
                o.depth_p2 = L.d_p2;
                o.depth_q1 = L.d_q1;
                o.depth_q2 = L.d_q2;
                o.flags = L.flags & 1u;
            }
        return L3D_OK;
    }
    if (which == 0) {
        if (!ctx->prm.keep_scored) return fail(L3D_ERR_STATE, "keep_scored was not set");
        if (v < ctx->slice_view[ctx->rank] || v >= ctx->slice_view[ctx->rank + 1])
            return fail(L3D_ERR_STATE, "the scored lists of camera %u are held by another rank", cam_id);
        CK(cudaMemcpyAsync(off.data(), ctx->d_L_off.p + hv.seg_off, ((size_t)N + 1) * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cnt.data(), ctx->d_L_cnt.p + hv.seg_off, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t o0 = off[0];
        for (uint32_t i = 0; i <= N; ++i) off[i] -= o0;
        src = ctx->d_L_rec.p + o0;  // rows live at their L_off
        region = off[N];
    } else {
        CK(cudaMemcpyAsync(off.data(), ctx->d_filt_off.p + hv.seg_off, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cnt.data(), ctx->d_filt_cnt.p + hv.seg_off, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        uint32_t lo = 0xffffffffu, hi = 0;
        for (uint32_t i = 0; i < N; ++i)
            if (cnt[i]) {
                lo = std::min(lo, off[i]);
                hi = std::max(hi, off[i] + cnt[i]);
            }
        if (lo == 0xffffffffu) lo = hi = 0;
        for (uint32_t i = 0; i < N; ++i) off[i] = cnt[i] ? off[i] - lo : 0;
        src = (ctx->filt_all ? ctx->filt_all : ctx->d_filt_rec.p) + lo;
        region = hi - lo;
    }
    uint64_t total = 0;
    for (uint32_t i = 0; i < N; ++i) {
        row_off[i] = (uint32_t)total;
        total += cnt[i];
    }
    row_off[N] = (uint32_t)total;
    if (out_count) *out_count = total;
    if (total > cap) return fail(L3D_ERR_CAPACITY, "need %llu records", (unsigned long long)total);
    if (total == 0) return L3D_OK;
    if (!recs) return fail(L3D_ERR_ARG, "recs is NULL");
    std::vector<ListRec> tmp(region);
    CK(cudaMemcpyAsync(tmp.data(), src, region * sizeof(ListRec), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t n = 0;
    for (uint32_t i = 0; i < N; ++i)
        for (uint32_t e = 0; e < cnt[i]; ++e) {
            const ListRec& L = tmp[(size_t)off[i] + e];
            l3d_list_rec& o = recs[n++];
            o.tgt_cam = ctx->views[L.tgt_view].v.cam_id;
            o.tgt_seg = L.tgt_seg;
            o.overlap_score = L.overlap;
            o.score3D = L.score;
            o.depth_p1 = L.d_p1;
            o.depth_p2 = L.d_p2;
            o.depth_q1 = L.d_q1;
            o.depth_q2 = L.d_q2;
            o.flags = L.flags & 1u;
        }
    return L3D_OK;
}

int l3d_get_entries(l3d_ctx* ctx, l3d_entry* out, uint32_t cap)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 2) return fail(L3D_ERR_STATE, "l3d_match_images has not run");
    if (cap < ctx->cnt.num_entries) return fail(L3D_ERR_CAPACITY, "need %u entries", ctx->cnt.num_entries);
    if (ctx->cnt.num_entries == 0) return L3D_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<EntryDev> e(ctx->S);
    CK(cudaMemcpyAsync(e.data(), ctx->d_entries.p, (size_t)ctx->S * sizeof(EntryDev), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    uint32_t n = 0;
    for (uint32_t v = 0; v < ctx->views.size(); ++v) {
        const HostView& hv = ctx->views[v];
        for (uint32_t i = 0; i < hv.v.num_segs; ++i) {
            const EntryDev& E = e[hv.seg_off + i];
            if (!E.has) continue;
            l3d_entry& o = out[n++];
            o.src_cam = hv.v.cam_id;
            o.src_seg = i;
            o.tgt_cam = ctx->views[E.tgt_view].v.cam_id;
            o.tgt_seg = E.tgt_seg;
            o.overlap_score = E.overlap;
            o.score3D = E.score;
            o.depth_p1 = E.d_p1; o.depth_p2 = E.d_p2; o.depth_q1 = E.d_q1; o.depth_q2 = E.d_q2;
            o.length = E.length;
            o.pad = 0;
            memcpy(o.P1, E.P1, 24);
            memcpy(o.P2, E.P2, 24);
            memcpy(o.dir, E.dir, 24);
        }
    }
    return L3D_OK;
}

int l3d_get_edges(l3d_ctx* ctx, int32_t* ij, float* w, uint32_t cap)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 3) return fail(L3D_ERR_STATE, "l3d_affinity has not run");
    const uint32_t ne = ctx->cnt.num_edges;
    if (cap < ne) return fail(L3D_ERR_CAPACITY, "need %u edges", ne);
    if (!ne) return L3D_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ij, ctx->d_A_ij.p, (size_t)ne * sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(w, ctx->d_A_w.p, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return L3D_OK;
}

int l3d_get_local2global(l3d_ctx* ctx, uint32_t* cam_seg, uint32_t cap)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 3) return fail(L3D_ERR_STATE, "l3d_affinity has not run");
    const uint32_t n = ctx->cnt.num_local_ids;
    if (cap < n) return fail(L3D_ERR_CAPACITY, "need %u ids", n);
    if (!n) return L3D_OK;
    CK(cudaSetDevice(ctx->device));
    // (camera id, segment) of every id on the device, one copy into the caller's buffer
    CK(ctx->d_l2g_cs.ensure(n));
    ctx->cnt.gpu_launches += launch_l2g_camseg(ctx->d_l2g.p, n, ctx->d_seg_view.p, ctx->d_views.p, ctx->d_l2g_cs.p, ctx->stream);
    CK(cudaMemcpyAsync(cam_seg, ctx->d_l2g_cs.p, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return L3D_OK;
}

int l3d_get_cluster_ids(l3d_ctx* ctx, int32_t* out, uint32_t cap)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx->stage < 4) return fail(L3D_ERR_STATE, "l3d_cluster has not run");
    if (cap < ctx->cluster_ids.size()) return fail(L3D_ERR_CAPACITY, "need %zu ids", ctx->cluster_ids.size());
    if (!ctx->cluster_ids.empty()) memcpy(out, ctx->cluster_ids.data(), ctx->cluster_ids.size() * 4);
    return L3D_OK;
}

int l3d_get_view_info(l3d_ctx* ctx, uint32_t cam_id, double* C, float* kmm)
{
    if (!ctx || !C || !kmm) return fail(L3D_ERR_ARG, "NULL argument");
    auto f = ctx->cam2view.find(cam_id);
    if (f == ctx->cam2view.end()) return fail(L3D_ERR_ARG, "unknown camera %u", cam_id);
    const HostView& hv = ctx->views[f->second];
    C[0] = hv.cam.C.x; C[1] = hv.cam.C.y; C[2] = hv.cam.C.z;
    if (ctx->translated && hv.current) {  // between two stages of a step: report what untranslate() will restore
        C[0] += ctx->translation.x; C[1] += ctx->translation.y; C[2] += ctx->translation.z;
    }
    kmm[0] = hv.k;
    kmm[1] = hv.median_depth;
    kmm[2] = hv.median_sigma;
    return L3D_OK;
}

int l3d_get_med_scene_depth_lines(l3d_ctx* ctx, float* out)
{
    if (!ctx || !out) return fail(L3D_ERR_ARG, "NULL argument");
    *out = ctx->med_scene_depth_lines;
    return L3D_OK;
}

// ------------------------------------------------------------------------------------------
// device-math test hooks and the FP32 peak probe
// ------------------------------------------------------------------------------------------
int l3d_test_expf(l3d_ctx* ctx, const float* x, float* y, uint32_t n)
{
    if (!ctx || !x || !y) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    DevBuf<float> dx, dy;
    CK(dx.ensure(n));
    CK(dy.ensure(n));
    CK(cudaMemcpyAsync(dx.p, x, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cnt.gpu_launches += launch_test_expf(dx.p, dy.p, n, ctx->stream);
    CK(cudaMemcpyAsync(y, dy.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return L3D_OK;
}
int l3d_test_acos(l3d_ctx* ctx, const double* x, double* y, uint32_t n)
{
    if (!ctx || !x || !y) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    DevBuf<double> dx, dy;
    CK(dx.ensure(n));
    CK(dy.ensure(n));
    CK(cudaMemcpyAsync(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cnt.gpu_launches += launch_test_acos(dx.p, dy.p, n, ctx->stream);
    CK(cudaMemcpyAsync(y, dy.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return L3D_OK;
}
int l3d_bench_fp32_peak(l3d_ctx* ctx, float* tflops)
{
    if (!ctx || !tflops) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    DevBuf<float> sink;
    CK(sink.ensure(1 << 20));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 0.0f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, ctx->stream);
        ctx->cnt.gpu_launches += launch_fp32_peak(sink.p, blocks, iters, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        // 16 independent FMA chains per thread, 2 flop per FMA
        const double flop = (double)blocks * 256.0 * (double)iters * 16.0 * 2.0;
        if (rep > 0 && ms > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return L3D_OK;
}

// the FP64 (DFMA) pipe peak, measured the same way: the denominator of the K2 roofline
int l3d_bench_fp64_peak(l3d_ctx* ctx, float* tflops)
{
    if (!ctx || !tflops) return fail(L3D_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    DevBuf<double> sink;
    CK(sink.ensure(1 << 19));
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 0.0f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, ctx->stream);
        ctx->cnt.gpu_launches += launch_fp64_peak(sink.p, blocks, iters, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, a, b);
        // 8 independent DFMA chains per thread, 2 flop per FMA
        const double flop = (double)blocks * 256.0 * (double)iters * 8.0 * 2.0;
        if (rep > 0 && ms > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return L3D_OK;
}
