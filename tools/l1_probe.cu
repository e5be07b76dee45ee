// l1_probe.cu — micro-benchmark: L1 -> register-file load bandwidth of one B200 (the unit that bounds the
// wake kernel).  Every thread issues independent 16-byte loads from an L1-resident window.
//   pattern 0: warp-broadcast (all lanes read the same 16 B, as the wake kernel's voxel loads mostly do)
//   pattern 1: lane-contiguous (32 lanes x 16 B = 512 contiguous bytes)
//   pattern 2: 48-byte-strided lanes (neighbouring voxels)
// Prints achieved GB/s counted as bytes delivered to registers (32 lanes x 16 B per load instruction).
#include <cstdio>
#include <cuda_runtime.h>

template <int PATTERN>
__global__ void __launch_bounds__(256) probe(const double2* __restrict__ buf, int window, int iters, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double2* base = buf + (size_t)blockIdx.x * window;
    double acc = 0.0;
    int off = warp * 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int idx;
            if (PATTERN == 0) idx = (off + u * 5) & (window - 1);
            else if (PATTERN == 1) idx = (off + u * 37 + lane) & (window - 1);
            else idx = (off + u * 41 + lane * 3) & (window - 1);
            double2 v = __ldg(base + idx);
            acc += v.x + v.y;
        }
        off = (off + 11) & (window - 1);
    }
    if (acc == 123.456) out[0] = acc;
}

template <int PATTERN>
static double run(const double2* buf, int window, double* out, int ctas) {
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<PATTERN><<<ctas, 256>>>(buf, window, 64, out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<PATTERN><<<ctas, 256>>>(buf, window, iters, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    double bytes = (double)ctas * 256 * (double)iters * 8 * 16;
    return bytes / (ms * 1e-3) / 1e9;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int window = 1024;                 // 16 KB per CTA: L1-resident (power of two: index math is one AND)
    const int ctas = sms * 4;
    double2* buf; double* out;
    cudaMalloc(&buf, (size_t)ctas * window * sizeof(double2));
    cudaMemset(buf, 0, (size_t)ctas * window * sizeof(double2));
    cudaMalloc(&out, 8);
    double g0 = run<0>(buf, window, out, ctas), g1 = run<1>(buf, window, out, ctas), g2 = run<2>(buf, window, out, ctas);
    double nominal = 128.0 * sms * (khz * 1e3) / 1e9;
    printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"l1_broadcast_gbs\": %.1f, \"l1_contiguous_gbs\": %.1f, \"l1_strided48_gbs\": %.1f, "
           "\"nominal_128B_per_clk_gbs\": %.1f}\n", sms, khz / 1e3, g0, g1, g2, nominal);
    return 0;
}
