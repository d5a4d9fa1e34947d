// scan.cu -- exclusive prefix sum of uint32 (out[0..n-1] exclusive, out[n] = total).
// Hierarchical reduce / scan / apply; used for CSR offsets everywhere in the pipeline.
#include "internal.h"

namespace l3d {

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0u;
        const uint32_t si = warp_incl_scan(s, lane);
        if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) tot = si;
    }
    __syncthreads();
    const uint32_t r = inc - v + wsum[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                                   uint32_t* __restrict__ sums, uint32_t n)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// single block: exclusive scan of n values (any n), out[n] = total
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_kernel(const uint32_t* __restrict__ in,
                                                                   uint32_t* __restrict__ out, uint32_t n)
{
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += SCAN_TILE) {
        uint32_t v[SCAN_ITEMS];
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const uint32_t idx = base + threadIdx.x * SCAN_ITEMS + i;
            v[i] = (idx < n) ? in[idx] : 0u;
            s += v[i];
        }
        uint32_t tot;
        uint32_t ex = block_excl_scan(s, &tot) + carry;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            const uint32_t idx = base + threadIdx.x * SCAN_ITEMS + i;
            if (idx < n) out[idx] = ex;
            ex += v[i];
        }
        carry += tot;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                  uint32_t* __restrict__ out,
                                                                  const uint32_t* __restrict__ block_off,
                                                                  uint32_t n, uint32_t nblocks)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t idx = base + (size_t)threadIdx.x * SCAN_ITEMS + i;
        v[i] = (idx < n) ? in[idx] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + block_off[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        const size_t idx = base + (size_t)threadIdx.x * SCAN_ITEMS + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
    if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) out[n] = block_off[nblocks];
}

size_t scan_scratch_words(uint32_t n)
{
    size_t total = 0;
    size_t m = n;
    while (m > (size_t)SCAN_TILE) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        total += 2 * (m + 1);  // block sums + their scan
    }
    return total + 8;
}

int launch_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch, size_t scratch_words,
                    cudaStream_t st)
{
    (void)scratch_words;
    if (n <= (uint32_t)SCAN_TILE * 4) {  // small: one block is cheapest (one launch)
        scan_single_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n);
        return 1;
    }
    const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* sums = scratch;
    uint32_t* sums_scan = scratch + (nb + 1);
    int launches = 0;
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, sums, n);
    ++launches;
    launches += launch_scan_u32(sums, sums_scan, nb, scratch + 2 * (size_t)(nb + 1), 0, st);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, out, sums_scan, n, nb);
    ++launches;
    return launches;
}

}  // namespace l3d
