// micro-benchmark of the persistent kernel's device-wide barrier (csrc/nem_persist.cuh pk_grid_sync)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/pk_barrier_bench profiles/pk_barrier_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
static __device__ __forceinline__ void pk_grid_sync(unsigned *bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *vgen = bar + 1;
        const unsigned gen = *vgen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1u) { bar[0] = 0u; __threadfence(); atomicAdd(bar + 1, 1u); }
        else { while (*vgen == gen) { } }
        __threadfence();
    }
    __syncthreads();
}
__global__ void k_mine(unsigned *bar, int reps) { for (int r = 0; r < reps; r++) pk_grid_sync(bar, gridDim.x); }
__global__ void k_cg(int reps) { cg::grid_group g = cg::this_grid(); for (int r = 0; r < reps; r++) g.sync(); }
int main() {
    unsigned *bar; cudaMalloc(&bar, 16); cudaMemset(bar, 0, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int reps = 2000;
    for (int grid : {16, 40, 148, 296}) for (int threads : {512}) {
        void *a1[] = {&bar, &reps}; void *a2[] = {&reps};
        float ms1, ms2;
        cudaLaunchCooperativeKernel((void *)k_mine, dim3(grid), dim3(threads), a1, 0, 0);
        cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_mine, dim3(grid), dim3(threads), a1, 0, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        cudaLaunchCooperativeKernel((void *)k_cg, dim3(grid), dim3(threads), a2, 0, 0);
        cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_cg, dim3(grid), dim3(threads), a2, 0, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1);
        printf("grid %3d x %d: own barrier %.2f us, cooperative_groups grid.sync %.2f us  (%s)\n", grid, threads,
               ms1 * 1e3 / reps, ms2 * 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
