// dsmem_hash.cu -- the smem_hash.cu experiment with the table DISTRIBUTED over a thread-block cluster (sm_100a):
// G CTAs x 16384 slots {key, fp64 residue}, every CTA inserts/accumulates random ids into the whole table through
// distributed shared memory (remote atomicCAS on the key, remote CAS-loop add on the residue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_hash dsmem_hash.cu && ./dsmem_hash
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

constexpr int kSlots = 16384;

template <int P>
__global__ void __launch_bounds__(1024, 1) k(int edges_per_cta, int distinct, int rounds, unsigned long long *out) {
    extern __shared__ double s_val[];
    int *s_key = reinterpret_cast<int *>(s_val + kSlots);
    __shared__ unsigned spills;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned G = cl.num_blocks(), rank = cl.block_rank();
    const unsigned mask = G * kSlots - 1;
    for (int i = threadIdx.x; i < kSlots; i += blockDim.x) { s_val[i] = 0.0; s_key[i] = -1; }
    if (threadIdx.x == 0) spills = 0;
    cl.sync();
    const unsigned cluster_id = blockIdx.x / G;
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        for (int e = threadIdx.x; e < edges_per_cta; e += blockDim.x) {
            const unsigned id = mix(mix(cluster_id * 7919u + r) + (mix((rank * edges_per_cta + e) * 2654435761u + r) % (unsigned)distinct)) % 2400000u;
            unsigned h = (id * 2654435761u) >> 7 & mask;
            int slot = -1;
            for (int p = 0; p < P; p++, h = (h + 1) & mask) {
                int *kp = cl.map_shared_rank(s_key, h >> 14) + (h & (kSlots - 1));
                int kk = *kp;
                if (kk == -1) { kk = atomicCAS(kp, -1, (int)id); if (kk == -1) { slot = h; break; } }
                if (kk == (int)id) { slot = h; break; }
            }
            if (slot >= 0) atomicAdd(cl.map_shared_rank(s_val, (unsigned)slot >> 14) + (slot & (kSlots - 1)), 1e-3);
            else atomicAdd(&spills, 1u);
        }
        cl.sync();
        for (int i = threadIdx.x; i < kSlots; i += blockDim.x) { if (s_key[i] != -1) { s_key[i] = -1; s_val[i] = 0.0; } }
        cl.sync();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { atomicAdd(out, (unsigned long long)(t1 - t0)); atomicAdd(out + 1, (unsigned long long)spills); }
}

template <int P>
void run(int G, int edges_total, int distinct) {
    const int rounds = 32;
    unsigned long long *out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    size_t smem = (size_t)kSlots * 12;
    cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (G > 8) cudaFuncSetAttribute(k<P>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    cfg.gridDim = dim3(G);
    cudaOccupancyMaxActiveClusters(&nclusters, k<P>, &cfg);
    if (nclusters <= 0) { printf("G=%d not schedulable\n", G); return; }
    cfg.gridDim = dim3(nclusters * G);
    cudaLaunchKernelEx(&cfg, k<P>, edges_total / G, distinct, rounds, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / (nclusters * G) / rounds;
    printf("G %2d (%3d clusters) probe %2d  edges %7d distinct %7d (load %.2f): %9.0f cycles/round (%.1f us), %.3f edges/clk/SM, spills/round/cluster %.0f  %s\n",
           G, nclusters, P, edges_total, distinct, (double)distinct / (G * kSlots), cyc, cyc / 1965.0, edges_total / G / cyc,
           (double)h[1] / nclusters / rounds, cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    // one "wide level" per round: Reddit-like on 1-2 CTAs, MAG-like on 2, Amazon2M-like on 8/16
    run<4>(1, 16250, 13000);
    run<4>(2, 16250, 13000);
    run<4>(2, 26000, 19000);
    run<4>(4, 60000, 45000);
    run<4>(8, 170000, 134000);
    run<8>(8, 170000, 134000);
    run<4>(16, 170000, 134000);
    run<8>(16, 170000, 134000);
    return 0;
}
