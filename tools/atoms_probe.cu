// atoms_probe.cu — shared-memory atomic update rates on one GPU (what bounds the CIC deposit).
// Each thread issues `iters` atomic updates to pseudo-random cells of a 96 KB shared-memory table.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_probe tools/atoms_probe.cu && ./atoms_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kWords = 24576;      // 96 KB of u32 / 12288 u64 or fp64 cells

__device__ __forceinline__ unsigned lcg(unsigned s) { return s * 1664525u + 1013904223u; }

template <int kKind>
__global__ void __launch_bounds__(1024, 1) probe(int iters, unsigned long long* out) {
    extern __shared__ unsigned table[];
    for (int c = threadIdx.x; c < kWords; c += blockDim.x) table[c] = 0;
    __syncthreads();
    unsigned s = (blockIdx.x * 1024u + threadIdx.x) * 2654435761u + 12345u;
    unsigned acc = 0;
    double* dt = reinterpret_cast<double*>(table);
    unsigned long long* lt = reinterpret_cast<unsigned long long*>(table);
    for (int k = 0; k < iters; ++k) {
        s = lcg(s);
        const unsigned cell = (s >> 8) % (kWords / 2);
        if (kKind == 0) atomicAdd(table + cell, s & 0xffu);                          // u32, result unused
        if (kKind == 1) acc += atomicAdd(table + cell, s & 0xffu);                   // u32, result used
        if (kKind == 2) atomicAdd(dt + cell, (double)(s & 0xffu));                   // fp64 (CAS loop)
        if (kKind == 3) atomicAdd(lt + cell, (unsigned long long)(s & 0xffu));       // u64
        if (kKind == 4) {                                                            // 64-bit fixed point as 2 x u32 with carry
            const unsigned lo = s, hi = s & 0xfu;
            const unsigned old = atomicAdd(table + 2 * cell, lo);
            atomicAdd(table + 2 * cell + 1, hi + (old + lo < old ? 1u : 0u));
        }
        if (kKind == 5) atomicAdd(reinterpret_cast<float*>(table) + cell, (float)(s & 0xffu));   // fp32
    }
    __syncthreads();
    unsigned long long sum = acc;
    for (int c = threadIdx.x; c < kWords; c += blockDim.x) sum += table[c];
    if (sum == 0x1234567ull) out[0] = sum;
}

template <int kKind>
static void run(const char* name, int sms, double mhz) {
    unsigned long long* out;
    cudaMalloc(&out, 8);
    const int iters = 2048;
    const size_t smem = kWords * 4;
    cudaFuncSetAttribute(probe<kKind>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    probe<kKind><<<sms, 1024, smem>>>(iters, out);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) probe<kKind><<<sms, 1024, smem>>>(iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    const double updates = (double)sms * 1024 * iters;
    printf("%-28s %8.3f ms  %7.2f G updates/s  %6.2f updates/clk/SM (at %.0f MHz)  err=%s\n", name, ms,
           updates / ms * 1e-6, updates / (ms * 1e-3) / sms / (mhz * 1e6), mhz, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("u32 add (no result)", p.multiProcessorCount, mhz);
    run<1>("u32 add (result used)", p.multiProcessorCount, mhz);
    run<2>("fp64 add (CAS loop)", p.multiProcessorCount, mhz);
    run<3>("u64 add", p.multiProcessorCount, mhz);
    run<4>("2 x u32 fixed point + carry", p.multiProcessorCount, mhz);
    run<5>("fp32 add", p.multiProcessorCount, mhz);
    return 0;
}
