// Random-access roofs on B200 as a function of footprint: what bounds GFPush's table updates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ra random_access.cu && /tmp/ra
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

// mode 0: atomicAdd(double) with return; 1: RED (no return); 2: 16B ld.cg + st (settle); 3: 4B gather
template <int MODE, int UNROLL>
__global__ void k(double *tab, uint64_t slots, uint64_t per_thread, uint64_t region_slots, double *sink) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // region_slots: each CTA confines itself to its own region (models per-CTA slabs); 0 = whole table
    const uint64_t base = region_slots ? ((uint64_t)blockIdx.x * region_slots) % slots : 0;
    const uint64_t span = region_slots ? region_slots : slots;
    double acc = 0;
    for (uint64_t i = 0; i < per_thread; i += UNROLL) {
        uint64_t idx[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; q++) idx[q] = base + mix(tid * per_thread + i + q) % span;
#pragma unroll
        for (int q = 0; q < UNROLL; q++) {
            if (MODE == 0) acc += atomicAdd(&tab[idx[q]], 1.0);
            else if (MODE == 1) atomicAdd(&tab[idx[q]], 1.0);
            else if (MODE == 2) { double2 *s = reinterpret_cast<double2 *>(tab) + (idx[q] >> 1); double2 t = __ldcg(s); t.y += t.x; t.x = 0; *s = t; acc += t.y; }
            else acc += __ldg(reinterpret_cast<const float *>(tab) + idx[q] * 2);
        }
    }
    if (acc == 12345.678) *sink = acc;
}

template <int MODE>
void run(const char *name, double *tab, uint64_t bytes, uint64_t region_bytes, int block, int ctas_per_sm, double *sink) {
    const uint64_t slots = bytes / 8;
    const int grid = 148 * ctas_per_sm;
    const uint64_t per_thread = 2048;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, 4><<<grid, block>>>(tab, slots, 256, region_bytes / 8, sink);
    cudaEventRecord(a);
    k<MODE, 4><<<grid, block>>>(tab, slots, per_thread, region_bytes / 8, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)grid * block * per_thread;
    printf("%-22s footprint %8.0f MB region %7.1f MB  blk %4d x%d : %7.2f G ops/s  (%.2f ms)\n", name, bytes / 1e6,
           region_bytes / 1e6, block, ctas_per_sm, ops / ms / 1e6, ms);
}

int main() {
    double *tab, *sink;
    const uint64_t maxb = 32ull << 30;
    cudaMalloc(&tab, maxb); cudaMalloc(&sink, 8);
    cudaMemset(tab, 0, maxb);
    for (uint64_t mb : {16ull, 64ull, 128ull, 256ull, 512ull, 1024ull, 4096ull, 16384ull, 32768ull}) {
        uint64_t bytes = mb << 20;
        run<0>("atomicAdd f64 ret", tab, bytes, 0, 512, 2, sink);
        run<1>("red f64", tab, bytes, 0, 512, 2, sink);
        run<2>("ld.cg16+st16 RMW", tab, bytes, 0, 512, 2, sink);
        run<3>("gather 4B", tab, bytes, 0, 512, 2, sink);
    }
    printf("-- per-CTA regions (slab model): total = 296 regions\n");
    for (uint64_t rmb : {1ull, 4ull, 16ull, 64ull}) {
        run<0>("atomicAdd f64 ret", tab, 296ull * (rmb << 20), rmb << 20, 512, 2, sink);
        run<2>("ld.cg16+st16 RMW", tab, 296ull * (rmb << 20), rmb << 20, 512, 2, sink);
    }
    printf("-- occupancy sweep at 4 GB\n");
    for (int c : {1, 2, 4}) for (int blk : {256, 512, 1024}) if (blk * c <= 2048) run<0>("atomicAdd f64 ret", tab, 4096ull << 20, 0, blk, c, sink);
    return 0;
}
