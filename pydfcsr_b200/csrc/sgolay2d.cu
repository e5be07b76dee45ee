// sgolay2d.cu — true 2-D Savitzky-Golay smoothing / derivative stencils.
//
// Restates sgolay2d(z, window_size, order, derivative) of SGolay_filter.py:3-81 (dead code in the
// reference's run loop, deposit.py:187,194,227,232, offered here as a stand-alone operator): the input is
// extended by window/2 samples on every side with the reference's odd reflection
//     low sides : b - |z(mirror) - b|        high sides : b + |z(mirror) - b|        (b = border sample)
// (SGolay_filter.py:36-65; the top-right and bottom-left corners are reflected from the already extended
// right / bottom bands) and convolved ('valid') with window x window kernels from a least-squares fit.
//
// Device layout: a CTA stages the (32 + window - 1) x (32 + window - 1) extended input patch of its 32 x 32
// output tile in shared memory — the border rule is evaluated while staging, so the extended array never
// exists in HBM — together with up to three kernels, and every thread produces 4 outputs per kernel.
// TMA is not used: its out-of-bounds fill is a constant, not this data-dependent reflection.
// Bound: HBM, 8 B read + 8 B x n_kernels written per cell; grids here are <= 2000^2, i.e. microseconds.
#include "common.cuh"

namespace dfcsr {

constexpr int kSgTile = 32;
constexpr int kSgRows = 8;          // thread block 32 x 8, four output rows per thread
constexpr int kSgMaxWindow = 25;
constexpr int kSgMaxKernels = 3;

struct SgIn {
    const double* z;
    int rows, cols, half;
};

__device__ __forceinline__ double sg_at(const SgIn& in, int i, int j) { return in.z[(long long)i * in.cols + j]; }

// right band of row i (0 <= i < rows), c = 0.. half-1 columns past the last one
__device__ __forceinline__ double sg_right(const SgIn& in, int i, int c) {
    const double b = sg_at(in, i, in.cols - 1);
    return b + fabs(sg_at(in, i, in.cols - 2 - c) - b);
}

// bottom band of column j, r = 0.. half-1 rows past the last one
__device__ __forceinline__ double sg_bottom(const SgIn& in, int j, int r) {
    const double b = sg_at(in, in.rows - 1, j);
    return b + fabs(sg_at(in, in.rows - 2 - r, j) - b);
}

// value of the extended array at (i, j), -half <= i < rows + half, same for j
__device__ double sg_extended(const SgIn& in, int i, int j) {
    const int H = in.rows, W = in.cols;
    const bool top = i < 0, bot = i >= H, left = j < 0, right = j >= W;
    if (!top && !bot) {
        if (!left && !right) return sg_at(in, i, j);
        if (left) {
            const double b = sg_at(in, i, 0);
            return b - fabs(sg_at(in, i, -j) - b);
        }
        return sg_right(in, i, j - W);
    }
    if (top) {
        if (!left && !right) {
            const double b = sg_at(in, 0, j);
            return b - fabs(sg_at(in, -i, j) - b);
        }
        if (left) {
            const double b = sg_at(in, 0, 0);
            return b - fabs(sg_at(in, -i, -j) - b);
        }
        const double b = sg_right(in, 0, j - W);                  // SGolay_filter.py:60-61
        return b - fabs(sg_right(in, -i, j - W) - b);
    }
    const int r = i - H;
    if (!left && !right) return sg_bottom(in, j, r);
    if (right) {
        const double b = sg_at(in, H - 1, W - 1);
        return b + fabs(sg_at(in, H - 2 - r, W - 2 - (j - W)) - b);
    }
    const double b = sg_bottom(in, 0, r);                         // SGolay_filter.py:63-64
    return b - fabs(sg_bottom(in, -j, r) - b);
}

template <int kKernels>
__global__ void __launch_bounds__(kSgTile * kSgRows)
sgolay2d_kernel(SgIn in, int window, const double* __restrict__ kernels, double* __restrict__ out) {
    extern __shared__ double sg_smem[];
    constexpr int kPer = kSgTile / kSgRows;                        // output rows per thread
    const int span = kSgTile + window - 1;
    const int pitch = span | 1;                                    // odd pitch: rows start in different banks
    double* patch = sg_smem;                                       // span x pitch
    double* ker = sg_smem + span * pitch;                          // kKernels x window x window
    const int tid = threadIdx.y * kSgTile + threadIdx.x;
    const int i0 = blockIdx.y * kSgTile, j0 = blockIdx.x * kSgTile;
    for (int k = tid; k < kKernels * window * window; k += kSgTile * kSgRows) ker[k] = kernels[k];
    for (int k = tid; k < span * span; k += kSgTile * kSgRows) {
        const int a = k / span, b = k - a * span;
        const int i = i0 + a - in.half, j = j0 + b - in.half;
        patch[a * pitch + b] = (i < in.rows + in.half && j < in.cols + in.half) ? sg_extended(in, i, j) : 0.0;
    }
    __syncthreads();
    // 'valid' convolution: out(i,j) = sum_ab K(a,b) * Z(i + w-1-a, j + w-1-b), Z the extended array.
    // One kernel value feeds the thread's four rows, one patch value feeds all kernels.
    const int w1 = window - 1, ww = window * window;
    const int li = threadIdx.y, lj = threadIdx.x;
    double acc[kKernels][kPer];
#pragma unroll
    for (int q = 0; q < kKernels; ++q)
#pragma unroll
        for (int rr = 0; rr < kPer; ++rr) acc[q][rr] = 0.0;
    for (int a = 0; a < window; ++a) {
        const double* row = patch + (li + w1 - a) * pitch + lj + w1;
        const double* ka = ker + a * window;
        for (int b = 0; b < window; ++b) {
            double kv[kKernels], zv[kPer];
#pragma unroll
            for (int q = 0; q < kKernels; ++q) kv[q] = ka[q * ww + b];
#pragma unroll
            for (int rr = 0; rr < kPer; ++rr) zv[rr] = row[rr * kSgRows * pitch - b];
#pragma unroll
            for (int q = 0; q < kKernels; ++q)
#pragma unroll
                for (int rr = 0; rr < kPer; ++rr) acc[q][rr] = fma(kv[q], zv[rr], acc[q][rr]);
        }
    }
#pragma unroll
    for (int q = 0; q < kKernels; ++q)
#pragma unroll
        for (int rr = 0; rr < kPer; ++rr) {
            const int i = i0 + li + rr * kSgRows, j = j0 + lj;
            if (i < in.rows && j < in.cols) out[((long long)q * in.rows + i) * in.cols + j] = acc[q][rr];
        }
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_sgolay2d(const double* d_z, int32_t rows, int32_t cols, int32_t window,
                              const double* d_kernels, int32_t n_kernels, double* d_out, void* stream) {
    DFCSR_REQUIRE(d_z && d_kernels && d_out, "null pointer");
    DFCSR_REQUIRE(window >= 1 && (window & 1) && window <= kSgMaxWindow, "window must be odd and at most 25");
    DFCSR_REQUIRE(rows >= window && cols >= window, "array smaller than the window");
    DFCSR_REQUIRE(n_kernels >= 1 && n_kernels <= kSgMaxKernels, "1 to 3 kernels");
    SgIn in;
    in.z = d_z;
    in.rows = rows;
    in.cols = cols;
    in.half = window / 2;
    const int span = kSgTile + window - 1;
    const size_t smem = ((size_t)span * (span | 1) + (size_t)n_kernels * window * window) * sizeof(double);
    dim3 grid((cols + kSgTile - 1) / kSgTile, (rows + kSgTile - 1) / kSgTile);
    dim3 block(kSgTile, kSgRows);
    cudaStream_t st = as_stream(stream);
    if (n_kernels == 1) sgolay2d_kernel<1><<<grid, block, smem, st>>>(in, window, d_kernels, d_out);
    else if (n_kernels == 2) sgolay2d_kernel<2><<<grid, block, smem, st>>>(in, window, d_kernels, d_out);
    else sgolay2d_kernel<3><<<grid, block, smem, st>>>(in, window, d_kernels, d_out);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
