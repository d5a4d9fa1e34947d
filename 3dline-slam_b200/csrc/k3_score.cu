// k3_score.cu -- stand-alone scoring kernel behind l3d_score_matches (the drop-in for
// L3DPP::score_matches_GPU, include/cudawrapper.h:74-81): Line3D::scoringCPU's new-match branch
// (src/line3D.cc:1513-1547) over caller-provided lists (exact TU).  The resident pipeline uses the
// data-flow kernels (k3_dataflow.cu) instead.
#include "internal.h"
#include "score_core.cuh"

namespace l3d {

struct ScoreStats {
    unsigned long long sim_evals;
    unsigned long long scored;
    uint32_t max_score_ord;
    uint32_t num_valid;
};

// one warp per row; lanes own matches M, siblings are walked in list order (broadcast loads)
__global__ void __launch_bounds__(256) k3_score_kernel(uint32_t n_rows, const uint32_t* __restrict__ L_off,
                                                       ListRec* __restrict__ L_rec,
                                                       const ListGeo* __restrict__ L_geo, float two_sigA_sqr,
                                                       float min_sim, float dotcut,
                                                       ScoreStats* __restrict__ stats)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const uint32_t b = L_off[warp];
    const uint32_t m = L_off[warp + 1] - b;
    if (m == 0) return;
    unsigned long long evals = 0;
    // exp(x) <= 0.4966 for x < -0.70: with min_sim >= 0.5 the result would be truncated to 0
    // anyway; for smaller thresholds the early exits are disabled
    const float xcut = (min_sim >= 0.5f) ? -0.70f : -3.0e38f;
    const float pcut = (min_sim >= 0.5f) ? 0.5f : -1.0f;
    for (uint32_t base = 0; base < m; base += 32) {
        const uint32_t me = base + lane;
        const bool act = me < m;
        ListRec M;
        ListGeo G;
        if (act) {
            M = L_rec[b + me];
            G = L_geo[b + me];
        } else {
            M.tgt_view = 0xffffffffu; M.d_p1 = M.d_p2 = 0.0f;
            G.dir[0] = G.dir[1] = G.dir[2] = 0.0; G.reg1 = G.reg2 = 1.0f; G.length = 0.0f;
        }
        const D3 dirM = d3(G.dir[0], G.dir[1], G.dir[2]);
        const bool Mvalid = !(G.length < 1e-12);
        float score = 0.0f, stored = 0.0f;
        bool in_run = false;
        for (uint32_t j = 0; j < m; ++j) {
            const ListRec M2 = L_rec[b + j];
            const ListGeo G2 = L_geo[b + j];
            if (G2.run) in_run = false;
            if (!act || M2.tgt_view == M.tgt_view) continue;
            ++evals;
            Sib s2;
            s2.d_p1 = M2.d_p1;
            s2.d_p2 = M2.d_p2;
            s2.cam = M2.tgt_view;
            s2.flags = (G2.length < 1e-12) ? 0u : 2u;
            const float sim = sim_for_scoring(M.d_p1, M.d_p2, G.reg1, G.reg2, Mvalid, dirM, s2, G2.dir, two_sigA_sqr,
                                              min_sim, xcut, pcut, dotcut);
            if (in_run) {
                if (sim > stored) {
                    score = fs(score, stored);
                    score = fa(score, sim);
                    stored = sim;
                }
            } else {
                score = fa(score, sim);
                stored = sim;
                in_run = true;
            }
        }
        if (act) L_rec[b + me].score = score;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, d);
    if (lane == 0) {
        atomicAdd(&stats->sim_evals, evals);
        atomicAdd(&stats->scored, (unsigned long long)m);
    }
}

int launch_k3_score(uint32_t n_rows, const uint32_t* L_off, ListRec* L_rec, const ListGeo* L_geo, float two_sigA_sqr,
                    float min_sim, void* stats, cudaStream_t st)
{
    if (!n_rows) return 0;
    const uint32_t warps_per_block = 8;
    k3_score_kernel<<<(n_rows + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(
        n_rows, L_off, L_rec, L_geo, two_sigA_sqr, min_sim, score_dotcut(two_sigA_sqr, min_sim), (ScoreStats*)stats);
    return 1;
}

size_t k3_stats_bytes() { return sizeof(ScoreStats); }

}  // namespace l3d
