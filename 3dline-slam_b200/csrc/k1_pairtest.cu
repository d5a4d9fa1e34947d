// k1_pairtest.cu -- K1: the N x M epipolar-overlap test of one view pair in FP32, with a
// certified guard band, as a conservative pre-filter for the exact kernel (k2_exact.cu).
//
// Replaces the inner loop of Line3D::matchingCPU (src/line3D.cc:1124-1158): for source segment r
// with epipolar lines e1 = F p1, e2 = F p2 and target segment c on the line q1 + s u, the points
// the reference intersects (l2 x e1, l2 x e2, normalised) are q1 + s1 u and q1 + s2 u with
//        s_i = -(e_i . (q1x,q1y,1)) / (e_i.xy . u),
// the bounds test (src/line3D.cc:1142-1148) is s_i in [slo,shi] and Line3D::mutualOverlap
// (src/line3D.cc:1283-1362) is (min(hi,L)-max(lo,0)) / (max(hi,L)-min(lo,0)) on {0,L,s1,s2}.
// A pair is REJECTED only if it fails by more than the rounding-error bound D of the FP32
// evaluation (derivation in DESIGN.md section 4.1); everything else is a candidate and is
// re-evaluated in the reference's double sequence.  Surviving pairs are written as one bit per
// test, layout [pair][word = c/32][row r] so that both this kernel's stores and K2's loads are
// coalesced; per-row candidate counts come out of the same pass (popc).
//
// One CTA = 256 source rows of one pair; target descriptors (32 B each) are streamed through
// shared memory in 512-entry tiles with TMA bulk copies (cp.async.bulk) on a two-stage
// mbarrier pipeline; every lane reads the same descriptor (shared-memory broadcast).
#include "internal.h"
#include "tma.cuh"

namespace l3d {

static constexpr int K1_ROWS = 256;
static constexpr int K1_TILE = 512;  // descriptors per stage (16 KB)

// one MUFU.RCP, <= 1 ulp (covered by the guard band).  The .ftz form skips the range fix-up code of
// rcp.approx.f32 (six more instructions per reciprocal): a denormal denominator gives +-inf here,
// and every |D| < 1e-20 is forced to be a candidate anyway.
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct RowEpi {
    float A1, B1, C1, A2, B2, C2;  // epipolar lines of both endpoints, |(A,B)| = 1
    float cN;                      // numerator error bound (the larger of the two lines')
    bool degenerate;
};

// one pair test; returns true if the pair may be a match (candidate)
//
// Every test is a rejection ("reject if ..."), so NaN never rejects.  A denominator close to 0
// needs no special case: the error bound grows like 1/|D| -- for |D| below ~10 u the bound exceeds
// |s| itself (cD/|D| > 3) and the numerator part exceeds every image coordinate (cN/|D| > 800 px),
// so nothing can be rejected; D = 0 (or a denormal, flushed by the .ftz reciprocal) gives inf/NaN in
// every comparison.  The reference's "outer distance < 1 px => overlap 0" rule is left to the exact
// kernel: it cannot fire for segments longer than a pixel.
__device__ __forceinline__ bool pair_candidate(const RowEpi& e, const float4 d0, const float4 d1, float thr,
                                               float k2thr)
{
    const float cD = 32.0f * 5.9604645e-08f;
    const float N1 = fmaf(e.A1, d0.x, fmaf(e.B1, d0.y, e.C1));
    const float D1 = fmaf(e.A1, d0.z, e.B1 * d0.w);
    const float N2 = fmaf(e.A2, d0.x, fmaf(e.B2, d0.y, e.C2));
    const float D2 = fmaf(e.A2, d0.z, e.B2 * d0.w);
    const float r1 = rcp_approx(D1);
    const float r2 = rcp_approx(D2);
    const float s1 = -N1 * r1;
    const float s2 = -N2 * r2;
    const float lo = fminf(s1, s2), hi = fmaxf(s1, s2);
    // one bound for both intersection parameters: (cN + cD max|s|) max|1/D|
    const float D = fmaf(fmaxf(fabsf(lo), fabsf(hi)), cD, e.cN) * fmaxf(fabsf(r1), fabsf(r2));
    const float L = d1.x, slo = d1.y, shi = d1.z, g = d1.w;
    bool rej = (hi - D > shi) | (lo + D < slo);
    const float inner = fminf(hi, L) - fmaxf(lo, 0.0f);
    const float outer = fmaxf(hi, L) - fminf(lo, 0.0f);
    const float margin = fmaf(-thr, outer, inner);
    const float G = fmaf(D, k2thr, fmaf(outer, 1.0e-6f, g));
    rej |= (margin < -G);
    return !rej;
}

__global__ void __launch_bounds__(K1_ROWS) k1_pairtest_kernel(const PairDev* __restrict__ pairs,
                                                              const K1Cta* __restrict__ ctas,
                                                              const float4* __restrict__ segs,
                                                              const SegDesc* __restrict__ desc,
                                                              const float* __restrict__ view_xb,
                                                              uint32_t* __restrict__ mask,
                                                              uint32_t* __restrict__ cand_cnt,
                                                              RowEpi32* __restrict__ row_epi, float thr,
                                                              int filter_mode)
{
    __shared__ __align__(128) float4 tile[2][K1_TILE * 2];
    __shared__ __align__(8) uint64_t bars[2];

    const K1Cta cta = ctas[blockIdx.x];
    const PairDev& P = pairs[cta.pair];
    const uint32_t n_src = P.n_src, n_tgt = P.n_tgt;
    const uint32_t r = cta.tile * K1_ROWS + threadIdx.x;
    const bool row_ok = r < n_src;
    const SegDesc* __restrict__ tdesc = desc + P.tgt_off;

    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t ntiles = (n_tgt + K1_TILE - 1) / K1_TILE;
    if (threadIdx.x == 0 && ntiles > 0) {
        const uint32_t cnt = min((uint32_t)K1_TILE, n_tgt);
        mbar_expect_tx(&bars[0], cnt * 32u);
        tma_load_1d(&tile[0][0], tdesc, cnt * 32u, &bars[0]);
    }

    // ---- per-row set-up: epipolar lines in double, normalised, rounded to float ----
    RowEpi e;
    e.degenerate = false;
    e.A1 = e.B1 = e.C1 = e.A2 = e.B2 = e.C2 = 0.0f;
    e.cN = 0.0f;
    float nmin = 0.0f;
    if (row_ok) {
        const float4 sg = segs[P.src_off + r];
        const double p1x = sg.x, p1y = sg.y, p2x = sg.z, p2y = sg.w;
        const double* F = P.F;
        const double a1 = F[0] * p1x + F[1] * p1y + F[2], b1 = F[3] * p1x + F[4] * p1y + F[5],
                     c1 = F[6] * p1x + F[7] * p1y + F[8];
        const double a2 = F[0] * p2x + F[1] * p2y + F[2], b2 = F[3] * p2x + F[4] * p2y + F[5],
                     c2 = F[6] * p2x + F[7] * p2y + F[8];
        const double n1 = sqrt(a1 * a1 + b1 * b1), n2 = sqrt(a2 * a2 + b2 * b2);
        e.A1 = (float)(a1 / n1); e.B1 = (float)(b1 / n1); e.C1 = (float)(c1 / n1);
        e.A2 = (float)(a2 / n2); e.B2 = (float)(b2 / n2); e.C2 = (float)(c2 / n2);
        const float xb = view_xb[P.tgt_view];
        const float u8 = 8.0f * 5.9604645e-08f;
        e.cN = u8 * (xb + fmaxf(fabsf(e.C1), fabsf(e.C2)));
        const float chk = e.A1 + e.B1 + e.C1 + e.A2 + e.B2 + e.C2;
        e.degenerate = !(fabsf(chk) < 3.0e38f);  // NaN or Inf anywhere
        nmin = (float)fmin(n1, n2);
        // K2's FP32 ranking re-uses the row's lines; a degenerate row carries NaN bounds (nothing is certified)
        RowEpi32 re;
        re.A1 = e.A1; re.B1 = e.B1; re.C1 = e.C1; re.A2 = e.A2; re.B2 = e.B2; re.C2 = e.C2;
        re.cN = e.degenerate ? __int_as_float(0x7fc00000) : e.cN;
        re.nmin = nmin;
        row_epi[P.row_base - P.batch_row0 + r] = re;
    }
    const float k2thr = 2.0f * (1.0f + thr);
    const bool all_pass = (filter_mode != 0) | e.degenerate;

    uint32_t total = 0;
    uint32_t* __restrict__ mrow = mask + P.mask_base + r;

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t buf = t & 1;
        if (threadIdx.x == 0 && t + 1 < ntiles) {
            const uint32_t nb = buf ^ 1;
            const uint32_t base = (t + 1) * K1_TILE;
            const uint32_t cnt = min((uint32_t)K1_TILE, n_tgt - base);
            mbar_expect_tx(&bars[nb], cnt * 32u);
            tma_load_1d(&tile[nb][0], tdesc + base, cnt * 32u, &bars[nb]);
        }
        mbar_wait(&bars[buf], (t >> 1) & 1);

        const uint32_t base = t * K1_TILE;
        const uint32_t cnt = min((uint32_t)K1_TILE, n_tgt - base);
        const float4* __restrict__ tl = tile[buf];
        if (row_ok) {
            const uint32_t nwords = (cnt + 31) / 32;
            for (uint32_t w = 0; w < nwords; ++w) {
                const uint32_t j0 = w * 32;
                const uint32_t nj = min(32u, cnt - j0);
                uint32_t bits = 0;
                if (all_pass) {
                    bits = (nj == 32) ? 0xffffffffu : ((1u << nj) - 1u);
                } else if (nj == 32) {
#pragma unroll 8
                    for (uint32_t j = 0; j < 32; ++j) {
                        const float4 d0 = tl[2 * (j0 + j)], d1 = tl[2 * (j0 + j) + 1];
                        bits |= (pair_candidate(e, d0, d1, thr, k2thr) ? 1u : 0u) << j;
                    }
                } else {
                    for (uint32_t j = 0; j < nj; ++j) {
                        const float4 d0 = tl[2 * (j0 + j)], d1 = tl[2 * (j0 + j) + 1];
                        bits |= (pair_candidate(e, d0, d1, thr, k2thr) ? 1u : 0u) << j;
                    }
                }
                mrow[(size_t)((base >> 5) + w) * n_src] = bits;
                total += __popc(bits);
            }
        }
        __syncthreads();  // everyone is done with tile[buf] before it is refilled
    }
    if (row_ok) cand_cnt[P.row_base - P.batch_row0 + r] = total;
}

int launch_k1_pairtest(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, const float4* segs,
                       const SegDesc* desc, const float* view_xb, uint32_t* mask, uint32_t* cand_cnt,
                       RowEpi32* row_epi, float thr, int filter_mode, cudaStream_t st)
{
    if (n_ctas == 0) return 0;
    k1_pairtest_kernel<<<n_ctas, K1_ROWS, 0, st>>>(pairs, ctas, segs, desc, view_xb, mask, cand_cnt, row_epi,
                                                    thr, filter_mode);
    return 1;
}

int k1_rows_per_cta() { return K1_ROWS; }

}  // namespace l3d
