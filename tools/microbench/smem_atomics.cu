// smem_atomics.cu -- shared-memory atomic throughput on B200 (sm_100a), random addresses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
// One 1024-thread CTA per SM, table of `slots` entries in dynamic shared memory, `iters` operations per
// thread at pseudo-random slots.  Reports operations/s per SM and per GPU.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int slots, int iters, unsigned long long *sink) {
    extern __shared__ unsigned long long sm64[];
    int *keys = reinterpret_cast<int *>(sm64 + slots);
    for (int i = threadIdx.x; i < slots; i += blockDim.x) { sm64[i] = 0; keys[i] = -1; }
    __syncthreads();
    unsigned s = mix(blockIdx.x * 1024u + threadIdx.x + 1u);
    unsigned long long acc = 0;
    for (int i = 0; i < iters; i++) {
        s = mix(s + i);
        const unsigned j = s % (unsigned)slots;
        if (MODE == 0) acc += atomicCAS(keys + j, -1, (int)(s >> 8));                      // 32-bit CAS
        if (MODE == 1) acc += (unsigned long long)atomicAdd(reinterpret_cast<double *>(sm64) + j, 1.0);  // fp64 add (CAS loop)
        if (MODE == 2) acc += atomicAdd(sm64 + j, 1ull);                                    // u64 add
        if (MODE == 3) acc += atomicAdd(reinterpret_cast<unsigned *>(keys) + j, 1u);        // u32 add
        if (MODE == 4) atomicAdd(sm64 + j, 1ull);                                           // u64 add, result unused
        if (MODE == 5) atomicAdd(reinterpret_cast<double *>(sm64) + j, 1.0);                // fp64 add, result unused
        if (MODE == 6) { sm64[j] += 1; }                                                    // plain LDS+STS RMW (racy; cost reference)
        if (MODE == 7) {                                                                    // find-or-claim + fp64 add, what a push does
            int kk = atomicCAS(keys + j, -1, (int)j);
            if (kk == -1 || kk == (int)j) atomicAdd(reinterpret_cast<double *>(sm64) + j, 1.0);
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
}

template <int MODE>
void run(const char *name, int slots, int sms) {
    const int iters = 4096;
    unsigned long long *sink; cudaMalloc(&sink, 8);
    size_t smem = (size_t)slots * 12;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms, 1024, smem>>>(slots, 64, sink);
    cudaEventRecord(a);
    k<MODE><<<sms, 1024, smem>>>(slots, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)sms * 1024.0 * iters;
    printf("%-44s slots %6d : %7.2f G ops/s GPU, %6.3f ops/clk/SM @1.965GHz (%.3f ms) %s\n", name, slots, ops / ms / 1e6,
           ops / ms / 1e6 / sms / 1.965, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int slots : {2048, 16384}) {
        run<0>("atomicCAS u32 (ret)", slots, sms);
        run<3>("atomicAdd u32 (ret)", slots, sms);
        run<2>("atomicAdd u64 (ret)", slots, sms);
        run<4>("atomicAdd u64 (no ret)", slots, sms);
        run<1>("atomicAdd f64 (ret, CAS loop)", slots, sms);
        run<5>("atomicAdd f64 (no ret, CAS loop)", slots, sms);
        run<6>("plain LDS+STS u64 RMW (reference)", slots, sms);
        run<7>("CAS key + f64 add (one push)", slots, sms);
    }
    return 0;
}
