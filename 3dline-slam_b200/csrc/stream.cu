// stream.cu -- the incremental (key-frame stream) mode of the path behind the C ABI: what
// L3DPPing::Run (src/L3DPPing.cpp:98-236) drives on a Line3D object between two reconstructions --
// deleteImage (src/line3D.cc:396-430), addImage (src/line3D.cc:117-227), UpdataImage
// (src/line3D.cc:433-487) -- and the state Line3D::matchImages keeps from one cycle to the next:
// matched_ (a view pair is matched once), processed_ (inverse matches are stored only into views
// that have not been scored yet), the filtered match lists with their scores, Add_camID_ /
// Delete_camID_ (the score deltas of Line3D::scoringCPU, src/line3D.cc:1439-1512).
//
// Device side: the view table only grows (a deleted view keeps its segments and its last pose, as
// views_ does); K0/K1/K2 run over the NEW pairs of the cycle; the scoring walk is per view
// (k3_stream.cu); K4 and the clustering are the batch mode's.
#include "ctx.h"

#include <chrono>


int l3d_stream_begin(l3d_ctx* ctx, int neighbors_by_worldpoints)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx) leave_translated(ctx);  // a call that failed half-way may have left the cameras shifted
    int rc = l3d_scene_begin(ctx);
    if (rc) return rc;
    ctx->stream_mode = true;
    ctx->by_worldpoints = neighbors_by_worldpoints != 0;
    ctx->st_add.clear();
    ctx->st_del.clear();
    ctx->st_matched.clear();
    ctx->st_cycle = 0;
    ctx->S = 0;
    ctx->cnt = l3d_counts{};
    return L3D_OK;
}

// the resets L3DPPing::Run performs before it deletes / adds / updates (src/L3DPPing.cpp:98-103)
int l3d_stream_begin_cycle(l3d_ctx* ctx)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (!ctx->stream_mode) return fail(L3D_ERR_STATE, "not in stream mode (l3d_stream_begin)");
    for (auto& hv : ctx->views) hv.wps.clear();  // views2worldpoints_ / worldpoints2views_
    ctx->st_add.clear();
    ctx->st_del.clear();
    return L3D_OK;
}

// Line3D::addImage (src/line3D.cc:117-227).  As in the reference's fork the list is only checked for
// emptiness here; world points / neighbours are registered by l3d_stream_update_image.
int l3d_stream_add_image(l3d_ctx* ctx, const l3d_view* view, const float* segs, const uint32_t* wps_or_nbrs,
                         uint32_t n_list)
{
    if (!ctx || !view) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx) leave_translated(ctx);  // a call that failed half-way may have left the cameras shifted
    if (!ctx->stream_mode) return fail(L3D_ERR_STATE, "not in stream mode (l3d_stream_begin)");
    if (std::max(view->width, view->height) < 400)
        return fail(L3D_ERR_ARG, "image is too small for reliable results: %u px (larger side should be >= 400px)",
                    std::max(view->width, view->height));
    auto f = ctx->cam2view.find(view->cam_id);
    if (f != ctx->cam2view.end()) {
        if (ctx->views[f->second].current) return fail(L3D_ERR_ARG, "camera ID [%u] already in use!", view->cam_id);
        return fail(L3D_ERR_ARG, "camera ID [%u] was deleted; the stream mode does not re-use camera ids", view->cam_id);
    }
    if (!ctx->views.empty() && ctx->views.back().v.cam_id > view->cam_id)
        return fail(L3D_ERR_ARG, "stream mode: camera ids must be added in ascending order (%u after %u)", view->cam_id,
                    ctx->views.back().v.cam_id);
    if (n_list == 0)
        return fail(L3D_ERR_ARG, ctx->by_worldpoints ? "view [%u] has no worldpoints!" : "view [%u] has no visual neighbors!",
                    view->cam_id);
    (void)wps_or_nbrs;
    if (view->num_segs == 0 || !segs) return fail(L3D_ERR_ARG, "no line segments found in image [%u]!", view->cam_id);
    if ((uint64_t)ctx->S + view->num_segs > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "too many segments");
    HostView hv;
    hv.v = *view;
    hv.segs.assign(segs, segs + 4 * (size_t)view->num_segs);
    hv.cam.init(view->K, view->R, view->t);
    hv.seg_off = ctx->S;
    const uint32_t idx = (uint32_t)ctx->views.size();
    ctx->S += view->num_segs;
    ctx->cam2view[view->cam_id] = idx;
    ctx->views.push_back(std::move(hv));
    ctx->st_add.insert(idx);
    return L3D_OK;
}

// Line3D::deleteImage (src/line3D.cc:396-430)
int l3d_stream_delete_image(l3d_ctx* ctx, uint32_t cam_id)
{
    if (!ctx) return fail(L3D_ERR_ARG, "ctx is NULL");
    if (ctx) leave_translated(ctx);  // a call that failed half-way may have left the cameras shifted
    if (!ctx->stream_mode) return fail(L3D_ERR_STATE, "not in stream mode (l3d_stream_begin)");
    auto f = ctx->cam2view.find(cam_id);
    if (f == ctx->cam2view.end() || !ctx->views[f->second].current)
        return fail(L3D_ERR_ARG, "camera ID [%u] non_existent!", cam_id);
    HostView& hv = ctx->views[f->second];
    hv.current = false;
    hv.processed = false;
    hv.nb_views.clear();  // visual_neighbors_.erase
    hv.wps.clear();
    hv.num_wps = 0;
    ctx->st_del.insert(f->second);
    return L3D_OK;
}

// Line3D::UpdataImage (src/line3D.cc:433-487): new pose (View::UpdateView, src/view.cc:62-87) and the
// world points / neighbours of the cycle
int l3d_stream_update_image(l3d_ctx* ctx, uint32_t cam_id, const double* R, const double* t, float median_depth,
                            const uint32_t* wps_or_nbrs, uint32_t n_list)
{
    if (!ctx || !R || !t) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx) leave_translated(ctx);  // a call that failed half-way may have left the cameras shifted
    if (!ctx->stream_mode) return fail(L3D_ERR_STATE, "not in stream mode (l3d_stream_begin)");
    (void)median_depth;  // only seeds View::initial_median_depth_, which the path never reads
    auto f = ctx->cam2view.find(cam_id);
    if (f == ctx->cam2view.end()) return L3D_OK;  // unknown ids are ignored (src/line3D.cc:439)
    HostView& hv = ctx->views[f->second];
    hv.cam.update(R, t);
    memcpy(hv.v.R, R, sizeof(hv.v.R));
    memcpy(hv.v.t, t, sizeof(hv.v.t));
    if (!hv.current) return fail(L3D_ERR_ARG, "Can not find camID[%u] in views_reserved_", cam_id);
    if (n_list && !wps_or_nbrs) return fail(L3D_ERR_ARG, "NULL argument");
    if (ctx->by_worldpoints) {
        if (n_list == 0) return fail(L3D_ERR_ARG, "view [%u] has no worldpoints!", cam_id);
        hv.wps.assign(wps_or_nbrs, wps_or_nbrs + n_list);  // processWPlist (src/line3D.cc:230-241)
        hv.num_wps = n_list;
    } else {
        hv.nbrs.assign(wps_or_nbrs, wps_or_nbrs + n_list);  // setVisualNeighbors (src/line3D.cc:244-247)
        hv.has_fixed = true;
    }
    return L3D_OK;
}

// reservation floors of the growing tables (elements): 256 Ki segments, 1 Ki views, 2 Mi list entries
static const size_t kSegFloor = 1u << 18, kViewFloor = 1u << 10, kListFloor = 1u << 21;

template <typename T>
static void swap_buf(DevBuf<T>& a, DevBuf<T>& b)
{
    std::swap(a.p, b.p);
    std::swap(a.cap, b.cap);
}

// Line3D::matchImages (src/line3D.cc:496-640) on the persistent state
int stream_match_images(l3d_ctx* ctx, const l3d_params* params)
{
    int rc = set_params(ctx, params);
    if (rc) return rc;
    if (ctx->prm.shard_world > 1) return fail(L3D_ERR_ARG, "the stream mode runs on one GPU (replicas only)");
    const uint32_t V = (uint32_t)ctx->views.size();
    std::vector<uint32_t> cur;
    for (uint32_t v = 0; v < V; ++v)
        if (ctx->views[v].current) cur.push_back(v);
    if (V == 0) return fail(L3D_ERR_STATE, "no images to match! forgot to add them?");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->tm.reset();
    cudaEvent_t ev_total = ctx->tm.begin(L3D_T_TOTAL, st);
    const uint32_t S = ctx->S;

    // L3D_STREAM_TRACE=1: host wall time of the sections of a cycle (stderr)
    static const bool trace = getenv("L3D_STREAM_TRACE") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = tnow();
    std::string tr;
    auto lap = [&](const char* name) {
        if (!trace) return;
        const auto t = tnow();
        char buf[64];
        snprintf(buf, sizeof(buf), " %s %.0fus", name, std::chrono::duration<double, std::micro>(t - t_prev).count());
        tr += buf;
        t_prev = t;
    };
    // ---- new segments -> device (the tables of the views already there stay) ----
    uint32_t S_old = 0;
    for (auto& hv : ctx->views)
        if (hv.uploaded) S_old = hv.seg_off + hv.v.num_segs;
    CK(ensure_roomy(ctx->d_segs, S, kSegFloor, S_old, st));
    CK(ensure_roomy(ctx->d_seg_view, S, kSegFloor, S_old, st));
    CK(ensure_roomy(ctx->d_filt_off, (size_t)S + 1, kSegFloor, S_old, st));
    CK(ensure_roomy(ctx->d_filt_cnt, (size_t)S + 1, kSegFloor, S_old, st));
    if (S > S_old) {
        CK(cudaMemsetAsync(ctx->d_filt_off.p + S_old, 0, (size_t)(S - S_old) * 4, st));
        CK(cudaMemsetAsync(ctx->d_filt_cnt.p + S_old, 0, (size_t)(S - S_old) * 4, st));
    }
    std::vector<uint32_t> sv;
    for (uint32_t v = 0; v < V; ++v) {
        HostView& hv = ctx->views[v];
        if (hv.uploaded) continue;
        CK(cudaMemcpyAsync(ctx->d_segs.p + hv.seg_off, hv.segs.data(), (size_t)hv.v.num_segs * sizeof(float4),
                           cudaMemcpyHostToDevice, st));
        sv.assign(hv.v.num_segs, v);
        CK(cudaMemcpyAsync(ctx->d_seg_view.p + hv.seg_off, sv.data(), (size_t)hv.v.num_segs * 4, cudaMemcpyHostToDevice,
                           st));
        CK(cudaStreamSynchronize(st));  // sv is reused
        hv.uploaded = true;
    }
    CK(ensure_roomy(ctx->d_desc, S, kSegFloor));
    CK(ensure_roomy(ctx->d_rays, S, kSegFloor));
    CK(ensure_roomy(ctx->d_midray, 3 * (size_t)S, 3 * kSegFloor));
    CK(ensure_roomy(ctx->d_planes, S, kSegFloor));
    CK(ensure_roomy(ctx->d_v32, S, kSegFloor));
    CK(ensure_roomy(ctx->d_view_xb, V, kViewFloor));
    CK(ensure_roomy(ctx->d_views, V, kViewFloor));
    CK(ensure_roomy(ctx->d_entries, (size_t)S + 1, kSegFloor));
    CK(ensure_roomy(ctx->d_L_off, (size_t)S + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_L_cnt, (size_t)S + 1, kSegFloor));
    CK(ensure_roomy(ctx->d_has, (size_t)S + 1, kSegFloor));
    CK(ensure_roomy(ctx->d_entry_idx, (size_t)S + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_view_max, (size_t)V + 1, kViewFloor));
    CK(ensure_roomy(ctx->d_st_view_total, (size_t)V + 1, kViewFloor));
    CK(ctx->d_small.ensure(16));
    CK(ensure_roomy(ctx->d_scan, scan_scratch_words(S + 2) + 64, kSegFloor));

    lap("upload");
    // ---- translate(), spatial regularisers of the current views (src/line3D.cc:568-590) ----
    enter_translated(ctx);
    // untranslate() (src/line3D.cc:637) on every way out; a cycle that fails half-way leaves matched_ and the
    // lists in an undefined state -- the caller starts over with l3d_stream_begin
    struct Untranslate {
        l3d_ctx* c;
        ~Untranslate() { leave_translated(c); }
    } untranslate_on_exit{ctx};
    for (uint32_t v : cur)
        ctx->views[v].k = ctx->fixed3D ? ctx->prm.sigma_p / ctx->prm.const_reg_depth
                                       : ctx->views[v].cam.spatial_regularizer(ctx->prm.sigma_p);

    lap("translate+k");
    // ---- visual neighbours (src/line3D.cc:598-620) ----
    if (ctx->by_worldpoints) {
        std::vector<const hg::Camera*> cams(cur.size());
        std::vector<float> md(cur.size());
        std::vector<std::vector<uint32_t>> wps(cur.size()), nb;
        for (size_t i = 0; i < cur.size(); ++i) {
            cams[i] = &ctx->views[cur[i]].cam;
            md[i] = ctx->views[cur[i]].median_depth;
            wps[i] = ctx->views[cur[i]].wps;
        }
        hg::visual_neighbors_from_worldpoints(cams, md, wps, ctx->prm.num_neighbors, nb);
        for (size_t i = 0; i < cur.size(); ++i) {
            std::vector<uint32_t>& out = ctx->views[cur[i]].nb_views;
            out.clear();
            for (uint32_t j : nb[i]) out.push_back(cur[j]);  // cur is ascending, so is the image
        }
    } else {
        for (uint32_t v : cur) {
            HostView& hv = ctx->views[v];
            if (!hv.has_fixed) {  // no fixed list: the world-point route finds nothing and clears the set
                hv.nb_views.clear();
                continue;
            }
            if (!hv.nb_views.empty()) continue;  // filled once (src/line3D.cc:606)
            for (uint32_t cam : hv.nbrs) {
                auto f = ctx->cam2view.find(cam);  // views_ keeps deleted views
                if (f != ctx->cam2view.end()) hv.nb_views.push_back(f->second);
            }
            std::sort(hv.nb_views.begin(), hv.nb_views.end());
            hv.nb_views.erase(std::unique(hv.nb_views.begin(), hv.nb_views.end()), hv.nb_views.end());
        }
    }

    lap("neighbours");
    // ---- new pairs in computeMatches order (src/line3D.cc:846-887): matched_ is never cleared ----
    ctx->pairs.clear();
    for (uint32_t s : cur)
        for (uint32_t t : ctx->views[s].nb_views) {
            const std::pair<uint32_t, uint32_t> key(std::min(s, t), std::max(s, t));
            if (ctx->st_matched.count(key)) continue;
            ctx->st_matched.insert(key);
            HostPair hp;
            hp.src = s;
            hp.tgt = t;
            hp.batch = 0;
            hp.local = true;
            ctx->pairs.push_back(hp);
        }
    const uint32_t P = (uint32_t)ctx->pairs.size();
    ctx->pairs_h.assign(P, PairDev{});
    uint64_t row = 0, trow = 0;
    ctx->cnt.pair_tests = 0;
    for (uint32_t p = 0; p < P; ++p) {
        const HostPair& hp = ctx->pairs[p];
        const HostView& vs = ctx->views[hp.src];
        const HostView& vt = ctx->views[hp.tgt];
        PairDev& d = ctx->pairs_h[p];
        const hg::M3 F = hg::fundamental(vs.cam, vt.cam);
        memcpy(d.F, F.m, sizeof(d.F));
        d.src_view = hp.src;
        d.tgt_view = hp.tgt;
        d.src_off = vs.seg_off;
        d.n_src = vs.v.num_segs;
        d.tgt_off = vt.seg_off;
        d.n_tgt = vt.v.num_segs;
        d.row_base = (uint32_t)row;
        d.tgt_base = (uint32_t)trow;
        d.words = (d.n_tgt + 31) / 32;
        d.emit_inverse = 0u;
        d.xflag = 0u;
        row += d.n_src;
        trow += d.n_tgt;
        ctx->cnt.pair_tests += (uint64_t)d.n_src * d.n_tgt;
    }
    if (row > 0xfffffff0ull || trow > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "row index space exhausted");
    ctx->total_rows = (uint32_t)row;
    ctx->total_tgt_rows = (uint32_t)trow;
    ctx->world = 1;
    ctx->rank = 0;
    ctx->slice_view.assign(2, V);
    ctx->slice_view[0] = 0;
    ctx->slice_g.assign(2, S);
    ctx->slice_g[0] = 0;
    ctx->slice_row.assign(2, ctx->total_rows);
    ctx->slice_row[0] = 0;
    // K0 refreshes the per-segment tables of the current views (their poses moved) and of the deleted views
    // that were matched against after their deletion (explicit-neighbour mode only: later cycles still
    // re-triangulate hypotheses whose target they are); the tables of the other deleted views are dead
    for (uint32_t p = 0; p < P; ++p)
        if (!ctx->views[ctx->pairs[p].tgt].current) ctx->views[ctx->pairs[p].tgt].paired_after_delete = true;
    ctx->view_needed.assign(V, 0u);
    for (uint32_t v = 0; v < V; ++v)
        ctx->view_needed[v] = (ctx->views[v].current || ctx->views[v].paired_after_delete) ? 1u : 0u;
    ctx->cnt.num_pairs = ctx->cnt.num_pairs_local = P;
    ctx->cnt.num_views = (uint32_t)cur.size();
    plan_batches(ctx);
    rc = upload_views(ctx);
    if (rc) return rc;

    lap("pairs+views");
    // ---- K0 (all views: the poses moved), K1 + K2 over the new pairs ----
    rc = run_stage12_batches(ctx);
    if (rc) return rc;
    rc = refresh_pair_totals(ctx);
    if (rc) return rc;
    ctx->cnt.forward_matches = ctx->total_fwd;
    if (ctx->k1_run_pending) ctx->cnt.pair_tests_run = *ctx->rb_at<unsigned long long>(l3d_ctx::RB_K1RUN);
    ctx->k1_run_pending = false;
    const size_t F = (size_t)ctx->total_fwd;
    CK(ensure_roomy(ctx->d_fwd_score, F + 1, kListFloor));
    CK(cudaMemsetAsync(ctx->d_fwd_score.p, 0, (F + 1) * sizeof(float), st));

    lap("K0-K2");
    // ---- the walk over the current views (src/line3D.cc:848-930) ----
    cudaEvent_t ev = ctx->tm.begin(L3D_T_SCORE, st);
    std::vector<std::vector<uint32_t>> in_of(V), out_of(V);
    for (uint32_t p = 0; p < P; ++p) {
        const HostPair& hp = ctx->pairs[p];
        out_of[hp.src].push_back(p);
        const HostView& vt = ctx->views[hp.tgt];
        // storeInverseMatches (src/line3D.cc:1994): the target has not been scored yet, neither in an
        // earlier cycle nor earlier in this walk
        if (vt.current && !vt.processed && hp.tgt > hp.src) in_of[hp.tgt].push_back(p);
    }
    // descriptors: per view its outgoing pairs (targets ascending), per NEW view its incoming pairs
    std::vector<StreamPair> sp;
    std::vector<uint32_t> in0(V, 0), vout0(V, 0), vnout(V, 0);
    std::vector<uint64_t> vcap(V, 0), vin(V, 0);
    for (uint32_t v : cur) {
        const HostView& hv = ctx->views[v];
        uint64_t cap = hv.filt_total;
        auto desc = [&](uint32_t p, uint32_t other) {
            const PairDev& d = ctx->pairs_h[p];
            StreamPair q;
            q.rec_start = ctx->pairs[p].rec_start;
            q.rec_cnt = (uint32_t)ctx->pairs[p].fwd_total;
            q.row_base = d.row_base;
            q.n_src = d.n_src;
            q.other = other;
            q.pad = 0;
            cap += q.rec_cnt;
            return q;
        };
        in0[v] = (uint32_t)sp.size();
        for (uint32_t p : in_of[v]) {
            sp.push_back(desc(p, ctx->pairs[p].src));
            vin[v] += ctx->pairs[p].fwd_total;
        }
        vout0[v] = (uint32_t)sp.size();
        vnout[v] = (uint32_t)out_of[v].size();
        for (uint32_t p : out_of[v]) sp.push_back(desc(p, ctx->pairs[p].tgt));
        vcap[v] = cap;
    }
    // steps: every view scored in an earlier cycle receives no inverse matches any more, so those views
    // do not feed each other and go in ONE step; each new view follows in ascending camera id (it
    // receives the inverse matches of the views before it, src/line3D.cc:1994)
    struct Step {
        std::vector<uint32_t> views;
        uint32_t row0 = 0, n = 0;
        uint64_t cap = 0, base = 0;
    };
    std::vector<Step> steps;
    {
        Step old;
        for (uint32_t v : cur)
            if (ctx->views[v].processed) old.views.push_back(v);
        if (!old.views.empty()) steps.push_back(old);
        for (uint32_t v : cur)
            if (!ctx->views[v].processed) {
                Step one;
                one.views.push_back(v);
                steps.push_back(one);
            }
    }
    std::vector<uint32_t> row_g;
    uint64_t extent = 0;
    uint32_t max_n = 0;
    uint64_t max_in = 0;
    for (Step& sx : steps) {
        sx.row0 = (uint32_t)row_g.size();
        for (uint32_t v : sx.views) {
            const HostView& hv = ctx->views[v];
            for (uint32_t i = 0; i < hv.v.num_segs; ++i) row_g.push_back(hv.seg_off + i);
            sx.cap += vcap[v];
            max_in = std::max(max_in, vin[v]);
        }
        sx.n = (uint32_t)row_g.size() - sx.row0;
        sx.base = extent;
        extent += sx.cap;
        max_n = std::max(max_n, sx.n);
    }
    if (extent > 0xfffffff0ull) return fail(L3D_ERR_CAPACITY, "too many list entries (%llu)", (unsigned long long)extent);
    ctx->st_w_extent = ctx->st_f_extent = extent;
    swap_buf(ctx->d_filt_rec, ctx->d_st_filt_old);  // last cycle's filtered lists become the persisted input
    CK(ensure_roomy(ctx->d_filt_rec, extent + 1, kListFloor));
    CK(ensure_roomy(ctx->d_st_W_rec, extent + 1, kListFloor));
    CK(ensure_roomy(ctx->d_st_W_row, extent + 1, kListFloor));
    CK(ensure_roomy(ctx->d_st_W_geo, (extent + 1) * sizeof(ListGeo), kListFloor * sizeof(ListGeo)));
    CK(ctx->d_st_pairs.ensure((sp.size() + 1) * sizeof(StreamPair)));
    CK(ensure_roomy(ctx->d_st_vflag, (size_t)V + 1, kViewFloor));
    CK(ctx->d_st_stats.ensure(stream_stats_bytes()));
    CK(ensure_roomy(ctx->d_st_row_g, row_g.size() + 1, kSegFloor));
    CK(ensure_roomy(ctx->d_st_vout, 2 * (size_t)V + 2, 2 * kViewFloor));
    CK(ensure_roomy(ctx->d_st_I_cnt, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_I_off, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_I_fill, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_I_key, (size_t)max_in + 1, kListFloor));
    CK(ensure_roomy(ctx->d_st_W_cnt, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_W_off, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_F_cnt, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_F_off, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_st_best, (size_t)max_n + 2, kSegFloor));
    CK(ensure_roomy(ctx->d_scan, scan_scratch_words(std::max(S, max_n) + 2) + 64, kSegFloor));
    std::vector<unsigned char> vflag(V + 1, 0);
    for (uint32_t v : ctx->st_add) vflag[v] |= 1u;
    for (uint32_t v : ctx->st_del) vflag[v] |= 2u;
    std::vector<uint32_t> vout(2 * (size_t)V + 2, 0);
    for (uint32_t v = 0; v < V; ++v) {
        vout[v] = vout0[v];
        vout[V + v] = vnout[v];
    }
    if (!sp.empty())
        CK(cudaMemcpyAsync(ctx->d_st_pairs.p, sp.data(), sp.size() * sizeof(StreamPair), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_st_vflag.p, vflag.data(), (size_t)V + 1, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_st_vout.p, vout.data(), vout.size() * 4, cudaMemcpyHostToDevice, st));
    if (!row_g.empty())
        CK(cudaMemcpyAsync(ctx->d_st_row_g.p, row_g.data(), row_g.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // the sources are locals
    CK(cudaMemsetAsync(ctx->d_st_stats.p, 0, stream_stats_bytes(), st));
    CK(cudaMemsetAsync(ctx->d_view_max.p, 0, ((size_t)V + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_st_view_total.p, 0, ((size_t)V + 1) * 4, st));
    CK(cudaMemsetAsync(ctx->d_small.p, 0, 16 * 4, st));
    CK(cudaMemsetAsync(ctx->d_entries.p, 0, ((size_t)S + 1) * sizeof(EntryDev), st));  // estimated_position3D_.clear()
    CK(cudaMemsetAsync(ctx->d_L_cnt.p, 0, ((size_t)S + 1) * 4, st));
    for (uint32_t v : ctx->st_del) {  // the lists of a deleted view are never read again
        const HostView& hv = ctx->views[v];
        CK(cudaMemsetAsync(ctx->d_filt_cnt.p + hv.seg_off, 0, (size_t)hv.v.num_segs * 4, st));
        ctx->views[v].filt_total = 0;
    }
    for (const Step& sx : steps) {
        const bool single_new = sx.views.size() == 1 && !ctx->views[sx.views[0]].processed;
        const uint32_t v0 = sx.views[0];
        StreamStepArgs a{};
        a.n = sx.n;
        a.n_in = single_new ? vout0[v0] - in0[v0] : 0u;
        a.in0 = in0[v0];
        a.in_total = single_new ? (uint32_t)vin[v0] : 0u;
        a.w_base = a.f_base = (uint32_t)sx.base;
        a.w_cap = a.f_cap = (uint32_t)sx.cap;
        a.row_g = ctx->d_st_row_g.p + sx.row0;
        a.seg_view = ctx->d_seg_view.p;
        a.pairs = ctx->d_st_pairs.p;
        a.vout0 = ctx->d_st_vout.p;
        a.vnout = ctx->d_st_vout.p + V;
        a.fwd_rec = ctx->d_fwd_rec.p; a.fwd_score = ctx->d_fwd_score.p;
        a.fwd_off = ctx->d_fwd_off.p; a.fwd_cnt = ctx->d_fwd_cnt.p;
        a.I_off = ctx->d_st_I_off.p; a.I_cnt = ctx->d_st_I_cnt.p; a.I_fill = ctx->d_st_I_fill.p;
        a.I_key = ctx->d_st_I_key.p;
        a.scan = ctx->d_scan.p; a.scan_words = ctx->d_scan.cap;
        a.filt_off = ctx->d_filt_off.p; a.filt_cnt = ctx->d_filt_cnt.p;
        a.filt_old = ctx->d_st_filt_old.p; a.filt_new = ctx->d_filt_rec.p;
        a.W_cnt = ctx->d_st_W_cnt.p; a.W_off = ctx->d_st_W_off.p; a.W_rec = ctx->d_st_W_rec.p;
        a.W_row = ctx->d_st_W_row.p; a.W_geo = (ListGeo*)ctx->d_st_W_geo.p;
        a.L_off = ctx->d_L_off.p; a.L_cnt = ctx->d_L_cnt.p;
        a.views = ctx->d_views.p; a.rays = ctx->d_rays.p; a.midray = ctx->d_midray.p; a.vflag = ctx->d_st_vflag.p;
        a.view_max = ctx->d_view_max.p; a.F_cnt = ctx->d_st_F_cnt.p; a.F_off = ctx->d_st_F_off.p;
        a.best_e = ctx->d_st_best.p; a.entries = ctx->d_entries.p; a.view_total = ctx->d_st_view_total.p;
        a.stats = ctx->d_st_stats.p; a.two_sigA_sqr = ctx->two_sigA_sqr;
        ctx->cnt.gpu_launches += launch_stream_step(a, st);
    }
    for (uint32_t v : cur) ctx->views[v].processed = true;

    lap("walk(launch)");
    // ---- view medians from the hypotheses as filterMatches stored them, then
    // update_Matches_and_Estimated_position3D, then the index of estimated_position3D_ ----
    ctx->cnt.gpu_launches += launch_k4_median(ctx->d_views.p, V, ctx->d_entries.p, ctx->d_small.p + 2, st);
    ctx->cnt.gpu_launches += launch_stream_update_entries(S, ctx->d_seg_view.p, ctx->d_views.p, ctx->d_rays.p,
                                                          ctx->d_planes.p, ctx->d_entries.p, st);
    ctx->cnt.gpu_launches += launch_k4_has(ctx->d_entries.p, S, ctx->d_has.p, st);
    ctx->cnt.gpu_launches += launch_scan_u32(ctx->d_has.p, ctx->d_entry_idx.p, S, ctx->d_scan.p, ctx->d_scan.cap, st);
    // read-backs into the pinned scratch (ctx.h); pageable fallback when the view table outgrows it
    uint32_t* small = ctx->rb_at<uint32_t>(l3d_ctx::RB_SMALL);
    unsigned char* stats = ctx->rb_at<unsigned char>(l3d_ctx::RB_STATS);
    std::vector<unsigned char> big_pageable;
    unsigned char* big = ctx->rb_at<unsigned char>(l3d_ctx::RB_BIG);
    const size_t vd_bytes = (size_t)V * sizeof(ViewDev), vt_bytes = (size_t)V * 4;
    if (!ctx->rb_fits(vd_bytes + vt_bytes)) {
        big_pageable.resize(vd_bytes + vt_bytes);
        big = big_pageable.data();
    }
    const ViewDev* vd = reinterpret_cast<const ViewDev*>(big);
    const uint32_t* vtot = reinterpret_cast<const uint32_t*>(big + vd_bytes);
    CK(cudaMemcpyAsync(small, ctx->d_small.p, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ctx->rb_at<uint32_t>(l3d_ctx::RB_NENT), ctx->d_entry_idx.p + S, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(big, ctx->d_views.p, vd_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(big + vd_bytes, ctx->d_st_view_total.p, vt_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(stats, ctx->d_st_stats.p, stream_stats_bytes(), cudaMemcpyDeviceToHost, st));
    ctx->tm.end(ev, st);
    ctx->tm.end(ev_total, st);
    CK(cudaStreamSynchronize(st));
    lap("finish+sync");
    if (trace) fprintf(stderr, "[l3d stream cycle %u]%s\n", ctx->st_cycle, tr.c_str());
    ctx->tm.collect();
    const uint32_t n_entries = *ctx->rb_at<uint32_t>(l3d_ctx::RB_NENT);
    const unsigned long long* s64 = (const unsigned long long*)stats;
    const uint32_t* s32 = (const uint32_t*)(stats + 16);
    if (s32[1]) return fail(L3D_ERR_CAPACITY, "internal: stream list arena overflow (%u)", s32[1]);
    if (small[2]) return fail(L3D_ERR_CAPACITY, "more than 8192 hypotheses in one view (median-depth kernel)");
    ctx->cnt.sim_evals = s64[0];
    ctx->cnt.scored_entries = s64[1];
    ctx->cnt.filtered_entries = s32[2];
    ctx->cnt.num_entries = n_entries;
    for (uint32_t v : cur) {
        HostView& hv = ctx->views[v];
        hv.median_depth = vd[v].median_depth;
        hv.median_sigma = hv.k * vd[v].median_depth;  // view.h:122-135
        hv.filt_total = vtot[v];
    }
    ctx->prog_all = nullptr;
    ctx->filt_all = nullptr;
    ctx->edges_all = nullptr;
    ctx->stage = 2;
    ctx->stage3_phase = 3;
    ctx->stage4_phase = 0;
    ctx->st_cycle++;
    return L3D_OK;
}
