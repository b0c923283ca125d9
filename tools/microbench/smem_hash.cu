// smem_hash.cu -- cost of one GFPush "wide level" on a shared-memory {key, fp64 residue} table (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_hash smem_hash.cu && ./smem_hash
// One 1024-thread CTA per SM; per round every CTA inserts/accumulates `edges` random node ids (a fraction `dup`
// of them repeats of earlier ids) into a 16384-slot open-addressed table with linear probing (probe limit P),
// then clears the table by a scan -- what expand + settle of one level would do to the table.  Reports SM cycles
// per round and the spill count (ids that found no slot within P probes).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int P>
__global__ void __launch_bounds__(1024, 1) k(int slots, int edges, int distinct, int rounds, unsigned long long *out) {
    extern __shared__ double s_val[];
    int *s_key = reinterpret_cast<int *>(s_val + slots);
    __shared__ unsigned spills;
    const unsigned mask = slots - 1;
    for (int i = threadIdx.x; i < slots; i += blockDim.x) { s_val[i] = 0.0; s_key[i] = -1; }
    if (threadIdx.x == 0) spills = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        for (int e = threadIdx.x; e < edges; e += blockDim.x) {
            // `distinct` different ids per round, drawn so that repeats happen
            const unsigned id = mix(mix(blockIdx.x * 7919u + r) + (mix(e * 2654435761u + r) % (unsigned)distinct)) % 233000u;
            unsigned h = (id * 2654435761u) >> 7 & mask;
            int slot = -1;
            for (int p = 0; p < P; p++, h = (h + 1) & mask) {
                int kk = s_key[h];
                if (kk == -1) { kk = atomicCAS(s_key + h, -1, (int)id); if (kk == -1) { slot = h; break; } }
                if (kk == (int)id) { slot = h; break; }
            }
            if (slot >= 0) atomicAdd(s_val + slot, 1e-3);
            else atomicAdd(&spills, 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < slots; i += blockDim.x) { if (s_key[i] != -1) { s_key[i] = -1; s_val[i] = 0.0; } }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { atomicAdd(out, (unsigned long long)(t1 - t0)); atomicAdd(out + 1, (unsigned long long)spills); }
}

template <int P>
void run(int sms, int edges, int distinct) {
    const int slots = 16384, rounds = 64;
    unsigned long long *out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    size_t smem = (size_t)slots * 12;
    cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<P><<<sms, 1024, smem>>>(slots, edges, distinct, rounds, out);
    unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("probe limit %3d  edges %6d distinct %6d (load %.2f): %8.0f cycles/round (%.2f us), %.3f edges/clk/SM, spills/round %.1f  %s\n",
           P, edges, distinct, (double)distinct / slots, (double)h[0] / sms / rounds, (double)h[0] / sms / rounds / 1965.0,
           edges / ((double)h[0] / sms / rounds), (double)h[1] / sms / rounds, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int distinct : {4000, 8000, 11000, 13000, 15000}) {
        int edges = distinct * 5 / 4;
        run<4>(sms, edges, distinct);
        run<8>(sms, edges, distinct);
        run<16>(sms, edges, distinct);
        run<64>(sms, edges, distinct);
    }
    return 0;
}
