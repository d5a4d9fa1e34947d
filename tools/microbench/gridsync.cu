// micro-benchmark: cost of cooperative_groups grid.sync() versus grid size on B200
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k(int iters, unsigned* sink)
{
    cg::grid_group g = cg::this_grid();
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        g.sync();
    }
    if (acc == 0xdeadbeef) *sink = acc;
}
// hand-rolled barrier: one atomic per CTA on a monotonically increasing counter
__global__ void k2(int iters, unsigned* bar, unsigned* sink)
{
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += i;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            const unsigned target = (unsigned)(i + 1) * gridDim.x;
            while (*(volatile unsigned*)bar < target) {}
            __threadfence();
        }
        __syncthreads();
    }
    if (acc == 0xdeadbeef) *sink = acc;
}
int main()
{
    unsigned *sink, *bar;
    cudaMalloc(&sink, 4);
    cudaMalloc(&bar, 4);
    int iters = 200;
    for (int threads : {64, 256}) {
        for (int per_sm : {1, 2, 4, 8, 12}) {
            int grid = 148 * per_sm;
            if (threads == 256 && per_sm > 8) continue;
            void* args[] = {&iters, &sink};
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(threads), args, 0, 0);
            cudaEventRecord(a);
            cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(grid), dim3(threads), args, 0, 0);
            cudaEventRecord(b);
            cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, a, b);
            cudaMemset(bar, 0, 4);
            void* args2[] = {&iters, &bar, &sink};
            cudaEventRecord(a);
            cudaError_t e2 = cudaLaunchCooperativeKernel((void*)k2, dim3(grid), dim3(threads), args2, 0, 0);
            cudaEventRecord(b);
            cudaDeviceSynchronize();
            float ms2 = 0; cudaEventElapsedTime(&ms2, a, b);
            printf("threads %4d grid %5d: cg grid.sync %.2f us/sync (%s)   atomic barrier %.2f us/sync (%s)\n", threads, grid,
                   1e3 * ms / iters, cudaGetErrorString(e), 1e3 * ms2 / iters, cudaGetErrorString(e2));
        }
    }
    return 0;
}
