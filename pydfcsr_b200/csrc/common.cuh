// common.cuh — shared helpers for libdfcsr_b200 (sm_100a).  Internal; the ABI is include/dfcsr_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dfcsr_b200.h"

namespace dfcsr {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);
void count_launch(int n);   // kernels launched by this library (dfcsr_launch_count)

// deposit.cu: both stages of the fixed-point deposit of one GPU, limits and max|px| taken from device memory
int deposit_one_gpu_from_device(const double* d_x, const double* d_z, const double* d_px, long long n, int nx, int nz,
                                const double* d_lim, const unsigned long long* d_wslot, long long* d_q, double* d_count,
                                double* d_vxsum, unsigned long long* d_count_max, cudaStream_t st);

#define DFCSR_CUDA_OK(expr)                                             \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return ::dfcsr::cuda_fail(_e, #expr);    \
    } while (0)

#define DFCSR_REQUIRE(cond, msg)                                        \
    do {                                                                \
        if (!(cond)) {                                                  \
            ::dfcsr::set_error("%s: %s", __func__, msg);                \
            return DFCSR_ERR_INVALID;                                   \
        }                                                               \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- numpy.linspace on the device -------------------------------------------------------------
// numpy builds node i as fl(fl(i*step) + start) with step = (stop-start)/(n-1) and overwrites the
// last node with `stop`.  The products/sums are written with explicit round-to-nearest intrinsics
// so that ptxas cannot contract them into an FMA: node values are then bit-identical to numpy's.
struct Axis {
    double start, stop, step;
    int n;
};

__host__ __device__ inline Axis make_axis(double start, double stop, int n) {
    Axis a;
    a.start = start;
    a.stop = stop;
    a.n = n;
    a.step = (n > 1) ? (stop - start) / (double)(n - 1) : 0.0;
    return a;
}

__device__ __forceinline__ double axis_node(const Axis& a, int i) {
    if (i >= a.n - 1 && a.n > 1) return a.stop;
    return __dadd_rn(__dmul_rn((double)i, a.step), a.start);
}

// ---- the reference's uniform-grid cell rule (interp3D.py:30-52, interp1D.py:21-34) ---------------
// u = (v - min)/delta; i0 = int(u) truncates toward zero; the last node clamps (i1 = i0);
// valid iff i0 >= 0 and i1 < n  <=>  -1 < u < n (NaN fails).  For u in (-1,0): i0 = 0 and the
// fraction is negative (linear extrapolation) — kept.
__device__ __forceinline__ bool cell_valid(double u, int n) { return (u > -1.0) && (u < (double)n); }

__device__ __forceinline__ void cell_split(double u, int n, int& i0, int& i1, double& frac) {
    i0 = __double2int_rz(u);
    i1 = (i0 == n - 1) ? i0 : i0 + 1;
    frac = u - (double)i0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace dfcsr
