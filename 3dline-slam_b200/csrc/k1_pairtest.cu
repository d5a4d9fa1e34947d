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
// test, layout [pair][word = c/32][rho] (rho = position of the source row in the sorted order below) so that both
// this kernel's stores and K2's loads are coalesced; per-row candidate counts come out of the same pass (popc).
//
// Two kernels per batch:
//   k1_rowsort_kernel   sorts the rows of every pair by the direction of their epipolar lines, so that the 32 rows
//                       of a warp span a narrow wedge through the epipole;
//   k1_pairtest_kernel  one CTA = 256 consecutive rows of the sorted order of one pair; target descriptors (32 B
//                       each) are streamed through shared memory in 512-entry tiles with TMA bulk copies
//                       (cp.async.bulk) on a two-stage mbarrier pipeline; per 32 targets the warp first tests the
//                       targets against its wedge (one target per lane) and then evaluates only the ones that reach
//                       into it (every lane the same descriptor: shared-memory broadcast).  On BASELINE config 4 one
//                       test in six is evaluated.
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------
// row order.  All epipolar lines of a pair pass through one point (the epipole), and a target segment can only
// overlap a source row if it reaches into the wedge between the row's two lines.  The rows of a pair are
// therefore SORTED by the direction of their lines before the test kernel runs: the 32 rows of a warp then span a
// narrow wedge (their "hull"), and a target segment that lies outside that wedge by more than half a pixel is
// skipped for the whole warp (k1_pairtest_kernel below).  The order only decides which tests are skipped -- any
// permutation gives the same mask -- so the sort keys may be as coarse as they like; what the skip test needs
// to be sound is stated next to it.
//
// One CTA sorts one chunk of up to KS_N rows of one pair:
//   * per row, in double: the two normalised epipolar lines (src/line3D.cc:1113-1121), with their signs chosen
//     so that the normals of all lines of the pair lie in one half plane (n.g >= 0, g = the direction at right
//     angles to the one from the epipole towards the image: no line that crosses the image comes near n.g = 0
//     unless the epipole is inside the image); the direction key: t = n.g' (the sine of the angle to g), monotone
//     in the direction over that half plane, for an epipole near the image; the signed distance of the middle of
//     the image to the line (= R t, R the distance of the epipole) for a far one, which also orders the
//     PARALLEL epipolar lines of a rectified stereo pair, where every angle is the same;
//   * sort key: width class (log2 of the row's wedge against the mean, 8 classes) above the quantised wedge
//     centre -- wide rows would otherwise widen the hull of every warp they land in; the centres run forwards in
//     even classes and backwards in odd ones;
//   * bitonic sort in shared memory; the rows' line records, keys and the permutation (both ways) are written
//     in sorted ("rho") order.  K1's mask and K2's front kernel work in rho order; everything else keeps the
//     natural row order through perm / iperm.
// ------------------------------------------------------------------------------------------
static constexpr int KS_N = 4096;  // rows per sort chunk = 16 K1 tiles
static constexpr int KS_THREADS = 512;

__global__ void __launch_bounds__(KS_THREADS, 2) k1_rowsort_kernel(const PairDev* __restrict__ pairs,
                                                                 const K1Cta* __restrict__ ctas,
                                                                 const float4* __restrict__ segs,
                                                                 const float* __restrict__ view_xb,
                                                                 RowEpi32* __restrict__ epi_nat,
                                                                 RowEpi32* __restrict__ epi_rho,
                                                                 float2* __restrict__ key_rho,
                                                                 uint32_t* __restrict__ perm,
                                                                 uint32_t* __restrict__ iperm, int sort_on)
{
    extern __shared__ __align__(16) unsigned char ks_raw[];  // 64 KB: above the static limit
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(ks_raw);
    float2* skey = reinterpret_cast<float2*>(ks_raw + sizeof(unsigned long long) * KS_N);
    __shared__ float red_lo[KS_THREADS / 32], red_hi[KS_THREADS / 32], red_w[KS_THREADS / 32];
    __shared__ uint32_t red_n[KS_THREADS / 32];
    __shared__ double s_g[6];
    const K1Cta cta = ctas[blockIdx.x];
    if (cta.tile % (KS_N / K1_ROWS)) return;  // one sort CTA per chunk: the K1 CTA list is reused as its grid
    const PairDev& P = pairs[cta.pair];
    const uint32_t r0 = (cta.tile / (KS_N / K1_ROWS)) * KS_N;
    const uint32_t n = min((uint32_t)KS_N, P.n_src - r0);
    const uint32_t lbase = P.row_base - P.batch_row0 + r0;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* F = P.F;
    if (tid == 0) {
        // the epipole e (e^T F = 0) is at right angles to every column of F: the cross product of the two columns
        // that give the longest one
        const double c0[3] = {F[0], F[3], F[6]}, c1[3] = {F[1], F[4], F[7]}, c2[3] = {F[2], F[5], F[8]};
        const double* cols[3] = {c0, c1, c2};
        double best[3] = {0, 0, 0}, bn = -1.0;
        for (int a = 0; a < 3; ++a) {
            const double* u = cols[a];
            const double* v = cols[(a + 1) % 3];
            const double x = u[1] * v[2] - u[2] * v[1], y = u[2] * v[0] - u[0] * v[2], z = u[0] * v[1] - u[1] * v[0];
            const double nn = x * x + y * y + z * z;
            if (nn > bn) { bn = nn; best[0] = x; best[1] = y; best[2] = z; }
        }
        // direction from the epipole towards the middle of the image, d = c e_w - e_xy (sign-free), g at right angles
        const double cx = 0.25 * (double)view_xb[P.tgt_view], cy = cx;  // xb bounds |x| + |y|
        double dx = cx * best[2] - best[0], dy = cy * best[2] - best[1];
        const double dn = sqrt(dx * dx + dy * dy);
        // distance of the epipole from the middle of the image, in pixels: |c e_w - e_xy| / |e_w|
        const double R = dn / fabs(best[2]);
        if (dn > 0.0 && dn < 1e300) { dx /= dn; dy /= dn; } else { dx = 1.0; dy = 0.0; }
        s_g[0] = -dy;
        s_g[1] = dx;
        // which key orders the lines of the pencil: near epipole -- the sine of the angle to g; far epipole (also at
        // infinity: parallel epipolar lines, a rectified stereo pair) -- the signed distance of the middle of the
        // image to the line, which is R times that sine and stays exact when the angles no longer tell the lines
        // apart.  (The distance alone would fail the other way round: an epipole at the middle of the image.)
        s_g[3] = (R <= 4.0 * (double)view_xb[P.tgt_view]) ? 0.0 : 1.0;  // NaN / inf: far
        s_g[4] = cx;
        s_g[5] = cy;
        // The wedge test needs the epipolar lines of the pair to pass through ONE point, i.e. a rank-2 F: true to
        // 1e-16 for a matrix made from two cameras in double, not guaranteed for a matrix handed in through
        // match_lines_GPU.  If the third column is not at right angles to the epipole to 1e-9, the rows keep their
        // order and get NaN keys: no warp of this pair skips anything.
        double worst = 0.0;
        const double nb = sqrt(bn);
        for (int a = 0; a < 3; ++a) {
            const double* u = cols[a];
            const double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            const double r = fabs(u[0] * best[0] + u[1] * best[1] + u[2] * best[2]);
            if (nu > 0.0 && nb > 0.0) worst = fmax(worst, r / (nu * nb));
        }
        s_g[2] = (bn > 0.0 && worst <= 1e-9) ? 1.0 : 0.0;
    }
    __syncthreads();
    const double gx = s_g[0], gy = s_g[1];
    const bool concurrent = s_g[2] != 0.0, far = s_g[3] != 0.0;
    const double kcx = s_g[4], kcy = s_g[5];
    const float xb = view_xb[P.tgt_view];
    const float pinf = __int_as_float(0x7f800000);
    float clo = pinf, chi = -pinf, wsum = 0.0f;
    uint32_t nval = 0;
    for (uint32_t i = tid; i < n; i += KS_THREADS) {
        const float4 sg = segs[P.src_off + r0 + i];
        const double p1x = sg.x, p1y = sg.y, p2x = sg.z, p2y = sg.w;
        double a1 = F[0] * p1x + F[1] * p1y + F[2], b1 = F[3] * p1x + F[4] * p1y + F[5], c1 = F[6] * p1x + F[7] * p1y + F[8];
        double a2 = F[0] * p2x + F[1] * p2y + F[2], b2 = F[3] * p2x + F[4] * p2y + F[5], c2 = F[6] * p2x + F[7] * p2y + F[8];
        const double n1 = sqrt(a1 * a1 + b1 * b1), n2 = sqrt(a2 * a2 + b2 * b2);
        a1 /= n1; b1 /= n1; c1 /= n1;
        a2 /= n2; b2 /= n2; c2 /= n2;
        // s = -N / D is unchanged by the sign of a line: choose it so that n.g >= 0
        if (a1 * gx + b1 * gy < 0.0) { a1 = -a1; b1 = -b1; c1 = -c1; }
        if (a2 * gx + b2 * gy < 0.0) { a2 = -a2; b2 = -b2; c2 = -c2; }
        // direction keys: n.g' with g' = (-gy, gx), or the distance of the image centre (see above)
        const double t1 = far ? a1 * kcx + b1 * kcy + c1 : b1 * gx - a1 * gy;
        const double t2 = far ? a2 * kcx + b2 * kcy + c2 : b2 * gx - a2 * gy;
        RowEpi32 re;
        re.A1 = (float)a1; re.B1 = (float)b1; re.C1 = (float)c1;
        re.A2 = (float)a2; re.B2 = (float)b2; re.C2 = (float)c2;
        const float u8 = 8.0f * 5.9604645e-08f;
        const float cN = u8 * (xb + fmaxf(fabsf(re.C1), fabsf(re.C2)));
        const float chk = re.A1 + re.B1 + re.C1 + re.A2 + re.B2 + re.C2;
        const bool degenerate = !(fabsf(chk) < 3.0e38f);  // NaN or Inf anywhere
        // K2's FP32 ranking re-uses the row's lines; a degenerate row carries NaN bounds (nothing is certified)
        re.cN = degenerate ? __int_as_float(0x7fc00000) : cN;
        re.nmin = (float)fmin(n1, n2);
        epi_nat[lbase + i] = re;
        float2 k = make_float2((float)t1, (float)t2);
        if (degenerate || !concurrent) k = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        skey[i] = k;
        if (!degenerate) {
            const float c = 0.5f * (k.x + k.y);
            clo = fminf(clo, c);
            chi = fmaxf(chi, c);
            wsum += fabsf(k.x - k.y);
            ++nval;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        clo = fminf(clo, __shfl_xor_sync(0xffffffffu, clo, d));
        chi = fmaxf(chi, __shfl_xor_sync(0xffffffffu, chi, d));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
        nval += __shfl_xor_sync(0xffffffffu, nval, d);
    }
    if (lane == 0) { red_lo[warp] = clo; red_hi[warp] = chi; red_w[warp] = wsum; red_n[warp] = nval; }
    __syncthreads();
    clo = red_lo[0]; chi = red_hi[0]; wsum = red_w[0]; nval = red_n[0];
    for (int w = 1; w < KS_THREADS / 32; ++w) {
        clo = fminf(clo, red_lo[w]);
        chi = fmaxf(chi, red_hi[w]);
        wsum += red_w[w];
        nval += red_n[w];
    }
    const float wmean = nval ? wsum / (float)nval : 0.0f;
    const float cscale = (chi > clo) ? (float)((1u << 29) - 1u) / (chi - clo) : 0.0f;
    uint32_t np2 = 32;
    while (np2 < n) np2 <<= 1;
    for (uint32_t i = tid; i < np2; i += KS_THREADS) {
        unsigned long long v = ~0ull;
        if (i < n) {
            const float2 k = skey[i];
            uint32_t key = 0xffffffffu;  // degenerate rows last
            if (k.x == k.x && sort_on) {
                const float w = fabsf(k.x - k.y), c = 0.5f * (k.x + k.y);
                int cls = 0;
                if (wmean > 0.0f && w > 0.0f) cls = min(7, max(0, (int)floorf(__log2f(w / wmean)) + 3));
                uint32_t q = (uint32_t)fminf(fmaxf((c - clo) * cscale, 0.0f), (float)((1u << 29) - 1u));
                // serpentine: odd classes run backwards, so the warp that straddles two classes sees neighbouring
                // centres (a jump from one end of the range to the other made its wedge cover everything, and that
                // one warp then held up its CTA at every tile barrier)
                if (cls & 1) q = ((1u << 29) - 1u) - q;
                key = ((uint32_t)cls << 29) | q;
            } else if (!sort_on) {
                key = 0u;  // ties keep the row index order (the index is the low word)
            }
            v = ((unsigned long long)key << 32) | i;
        }
        sk[i] = v;
    }
    __syncthreads();
    if (sort_on)
        for (uint32_t k = 2; k <= np2; k <<= 1)
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = tid; t < (np2 >> 1); t += KS_THREADS) {
                    const uint32_t i = 2 * t - (t & (j - 1));
                    const unsigned long long a = sk[i], b = sk[i + j];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        sk[i] = b;
                        sk[i + j] = a;
                    }
                }
                __syncthreads();
            }
    for (uint32_t i = tid; i < n; i += KS_THREADS) {
        const uint32_t idx = (uint32_t)(sk[i] & 0xffffffffull);
        perm[lbase + i] = r0 + idx;
        iperm[lbase + idx] = r0 + i;
        epi_rho[lbase + i] = epi_nat[lbase + idx];  // written by this CTA before the barriers above
        key_rho[lbase + i] = skey[idx];
    }
}

// Hull test of one target segment against the wedge of a warp's rows, hull lines (Al, Bl, Cl) and (Ah, Bh, Ch)
// (the rows' lines with the smallest / largest direction key; every other line of the warp is a combination
// a Hl + b Hh with a, b >= 0, a + b >= 1, because all lines pass through the epipole and their normals lie in one
// half plane).  With f(x) the signed distance of x to a line: if f_l and f_h are both > m (or both < -m) at both
// end points of the target segment, every line e of the warp has |f_e| >= m on the whole segment, so neither
// of a row's two intersection parameters lies in [-m, L + m]; if in addition the slopes D_l and D_h (and with
// them D_e, which lies between them) have one sign, both parameters lie on the SAME side: inner = min(hi, L) -
// max(lo, 0) <= -m, the reference's overlap is 0 and the pair is no match.  m = 0.5 px: the float roundings of
// the lines, of N, D and L D (each below 1e-2 px on image coordinates, see pair_candidate) and the lines not
// being exactly concurrent after rounding are far below it.  Every comparison is false for NaN: NaN never skips.
__device__ __forceinline__ bool hull_skips(float Al, float Bl, float Cl, float Ah, float Bh, float Ch, const float4 d0,
                                           float L)
{
    const float Nl = fmaf(Al, d0.x, fmaf(Bl, d0.y, Cl)), Dl = fmaf(Al, d0.z, Bl * d0.w);
    const float Nh = fmaf(Ah, d0.x, fmaf(Bh, d0.y, Ch)), Dh = fmaf(Ah, d0.z, Bh * d0.w);
    const float Nl2 = fmaf(L, Dl, Nl), Nh2 = fmaf(L, Dh, Nh);
    const float mn = fminf(fminf(Nl, Nl2), fminf(Nh, Nh2)), mx = fmaxf(fmaxf(Nl, Nl2), fmaxf(Nh, Nh2));
    const bool outside = (mn > 0.5f) | (mx < -0.5f);
    const bool slopes = (Dl * Dh > 0.0f) & (fminf(fabsf(Dl), fabsf(Dh)) > 1.0e-5f);
    return outside & slopes;
}

// K1_UNROLL: independent tests in flight per lane
template <int K1_UNROLL>
__global__ void __launch_bounds__(K1_ROWS) k1_pairtest_kernel(const PairDev* __restrict__ pairs,
                                                              const K1Cta* __restrict__ ctas,
                                                              const SegDesc* __restrict__ desc,
                                                              const RowEpi32* __restrict__ epi_rho,
                                                              const float2* __restrict__ key_rho,
                                                              const uint32_t* __restrict__ perm,
                                                              uint32_t* __restrict__ mask,
                                                              uint32_t* __restrict__ cand_cnt, float thr,
                                                              int filter_mode, int hull_on,
                                                              unsigned long long* __restrict__ tests_run)
{
    __shared__ __align__(128) float4 tile[2][K1_TILE * 2];
    __shared__ __align__(8) uint64_t bars[2];

    const K1Cta cta = ctas[blockIdx.x];
    const PairDev& P = pairs[cta.pair];
    const uint32_t n_src = P.n_src, n_tgt = P.n_tgt;
    const uint32_t rho = cta.tile * K1_ROWS + threadIdx.x;  // position in the sorted row order
    const bool row_ok = rho < n_src;
    const uint32_t lane = threadIdx.x & 31;
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

    // ---- the row's lines (k1_rowsort_kernel) and the hull of the warp ----
    RowEpi e;
    e.degenerate = false;
    e.A1 = e.B1 = e.C1 = e.A2 = e.B2 = e.C2 = 0.0f;
    e.cN = 0.0f;
    uint32_t lrow = 0;
    const float pinf = __int_as_float(0x7f800000);
    float klo = pinf, khi = -pinf;
    bool first_lo = true;
    if (row_ok) {
        const uint32_t lrho = P.row_base - P.batch_row0 + rho;
        const RowEpi32 re = epi_rho[lrho];
        e.A1 = re.A1; e.B1 = re.B1; e.C1 = re.C1; e.A2 = re.A2; e.B2 = re.B2; e.C2 = re.C2;
        e.degenerate = !(re.cN == re.cN);
        e.cN = e.degenerate ? 0.0f : re.cN;
        lrow = P.row_base - P.batch_row0 + perm[lrho];
        const float2 k = key_rho[lrho];
        first_lo = k.x <= k.y;
        klo = fminf(k.x, k.y);
        khi = fmaxf(k.x, k.y);
    }
    const float k2thr = 2.0f * (1.0f + thr);
    const bool all_pass = (filter_mode != 0) | e.degenerate;
    // hull lines: the line with the smallest key and the one with the largest key among the warp's rows
    float wlo = klo, whi = khi;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        wlo = fminf(wlo, __shfl_xor_sync(0xffffffffu, wlo, d));
        whi = fmaxf(whi, __shfl_xor_sync(0xffffffffu, whi, d));
    }
    const uint32_t own_lo = __ffs(__ballot_sync(0xffffffffu, row_ok && klo == wlo)) - 1;
    const uint32_t own_hi = __ffs(__ballot_sync(0xffffffffu, row_ok && khi == whi)) - 1;
    // no skipping unless every row of the warp has finite keys.  (All normals of a pair lie in one half plane, so
    // the lines between the two hull lines are non-negative combinations of them whatever the width of the wedge.)
    const bool hull_ok = hull_on && !__any_sync(0xffffffffu, row_ok && (all_pass || !(khi - klo >= 0.0f))) &&
                         own_lo < 32u && own_hi < 32u;
    const float lA = first_lo ? e.A1 : e.A2, lB = first_lo ? e.B1 : e.B2, lC = first_lo ? e.C1 : e.C2;
    const float hA = first_lo ? e.A2 : e.A1, hB = first_lo ? e.B2 : e.B1, hC = first_lo ? e.C2 : e.C1;
    const float Al = __shfl_sync(0xffffffffu, lA, own_lo & 31u), Bl = __shfl_sync(0xffffffffu, lB, own_lo & 31u),
                Cl = __shfl_sync(0xffffffffu, lC, own_lo & 31u);
    const float Ah = __shfl_sync(0xffffffffu, hA, own_hi & 31u), Bh = __shfl_sync(0xffffffffu, hB, own_hi & 31u),
                Ch = __shfl_sync(0xffffffffu, hC, own_hi & 31u);

    uint32_t total = 0;
    unsigned long long n_run = 0;  // targets this warp tested
    uint32_t* __restrict__ mrow = mask + P.mask_base + rho;

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
        const uint32_t nwords = (cnt + 31) / 32;
        for (uint32_t w = 0; w < nwords; ++w) {  // warp-uniform
            const uint32_t j0 = w * 32;
            const uint32_t nj = min(32u, cnt - j0);
            const uint32_t full = (nj == 32) ? 0xffffffffu : ((1u << nj) - 1u);
            const float4* __restrict__ tw = tl + 2 * j0;
            // which of the word's 32 targets the warp has to test: lane j looks at target j0 + j.  The ballot is a
            // warp-uniform value: the loop below, its bit positions and its shared-memory addresses run on the
            // uniform datapath.  (A queue of survivors across the words of a tile, read back per test, cost 34
            // instructions per test to put each result bit into its word -- the compiler could not know that the
            // queued indices were uniform.)
            uint32_t need = full;
            if (hull_ok) {
                bool skip = true;
                if (lane < nj) skip = hull_skips(Al, Bl, Cl, Ah, Bh, Ch, tw[2 * lane], tw[2 * lane + 1].x);
                need = __ballot_sync(0xffffffffu, !skip);
            }
            n_run += __popc(need);
            uint32_t bits = 0;
            while (need) {  // K1_UNROLL independent tests per iteration; a short tail repeats the last target
                uint32_t jj[K1_UNROLL];
                bool c[K1_UNROLL];
#pragma unroll
                for (int u = 0; u < K1_UNROLL; ++u) {
                    jj[u] = (u == 0 || need) ? (uint32_t)__ffs(need) - 1u : jj[u > 0 ? u - 1 : 0];
                    need &= need - 1u;
                }
#pragma unroll
                for (int u = 0; u < K1_UNROLL; ++u) c[u] = pair_candidate(e, tw[2 * jj[u]], tw[2 * jj[u] + 1], thr, k2thr);
#pragma unroll
                for (int u = 0; u < K1_UNROLL; ++u) bits |= (c[u] ? 1u : 0u) << jj[u];
            }
            if (all_pass) bits = full;
            if (row_ok) {
                mrow[(size_t)((base >> 5) + w) * n_src] = bits;
                total += __popc(bits);
            }
        }
        __syncthreads();  // everyone is done with tile[buf] before it is refilled
    }
    if (row_ok) cand_cnt[lrow] = total;
    const uint32_t live = __popc(__ballot_sync(0xffffffffu, row_ok));
    if (lane == 0 && live) atomicAdd(tests_run, n_run * live);
}

int launch_k1_rowsort(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, const float4* segs,
                      const float* view_xb, RowEpi32* epi_nat, RowEpi32* epi_rho, float2* key_rho, uint32_t* perm,
                      uint32_t* iperm, cudaStream_t st)
{
    if (n_ctas == 0) return 0;
    static int sort_on = -1;
    if (sort_on < 0) {
        const char* ev = getenv("L3D_K1_SORT");  // 0: keep the natural row order (A/B runs)
        sort_on = ev ? atoi(ev) : 1;
    }
    const size_t smem = (sizeof(unsigned long long) + sizeof(float2)) * KS_N;
    cudaFuncSetAttribute(k1_rowsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k1_rowsort_kernel<<<n_ctas, KS_THREADS, smem, st>>>(pairs, ctas, segs, view_xb, epi_nat, epi_rho, key_rho, perm, iperm,
                                                      sort_on);
    return 1;
}

int launch_k1_pairtest(const PairDev* pairs, const K1Cta* ctas, uint32_t n_ctas, const SegDesc* desc,
                       const RowEpi32* epi_rho, const float2* key_rho, const uint32_t* perm, uint32_t* mask,
                       uint32_t* cand_cnt, float thr, int filter_mode, unsigned long long* tests_run,
                       cudaStream_t st)
{
    if (n_ctas == 0) return 0;
    static int hull_on = -1;
    if (hull_on < 0) {
        const char* ev = getenv("L3D_K1_HULL");  // 0: test every pair (A/B runs)
        hull_on = ev ? atoi(ev) : 1;
    }
    static int unroll = -1;
    if (unroll < 0) {
        const char* ev = getenv("L3D_K1_UNROLL");  // tuning hook
        unroll = ev ? atoi(ev) : 2;
    }
#define K1_LAUNCH(U) \
    k1_pairtest_kernel<U><<<n_ctas, K1_ROWS, 0, st>>>(pairs, ctas, desc, epi_rho, key_rho, perm, mask, cand_cnt, thr, \
                                                      filter_mode, hull_on, tests_run)
    if (unroll >= 4) K1_LAUNCH(4);
    else if (unroll >= 2) K1_LAUNCH(2);
    else K1_LAUNCH(1);
#undef K1_LAUNCH
    return 1;
}

int k1_rows_per_cta() { return K1_ROWS; }

}  // namespace l3d
