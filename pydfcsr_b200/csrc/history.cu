// history.cu — K3: bilinear re-gridding of one raw density-function record into a voxel slice of the
// device-resident history ring, plus SoA<->AoS converters used to import/export oracle histories.
//
// Replaces DF_tracker.DF_interp (deposit.py:296-309), i.e. scipy RegularGridInterpolator(method=
// 'linear', bounds_error=False, fill_value=f) evaluated on meshgrid(x_grid_interp, z_grid_interp),
// called five times per step (deposit.py:328-332) or 5*T times on a rebuild (deposit.py:379-390),
// and the np.array(deque) re-stack of build_interpolant (deposit.py:422-426): the slice is written
// in place into its ring slot, so nothing is re-stacked or copied per step.
//
// scipy semantics restated (scipy/interpolate/_rgi.py:446-477,635-642):
//   per axis  i = clip(searchsorted(g, q, 'right') - 1, 0, n-2),  y = (q - g[i]) / (g[i+1] - g[i])
//   value     F[i,j](1-yx)(1-yz) + F[i,j+1](1-yx)yz + F[i+1,j]yx(1-yz) + F[i+1,j+1]yx yz
//   fill      where q < g[0] or q > g[-1] on either axis.
// Products/sums use explicit round-to-nearest intrinsics (no FMA contraction) so that, given
// identical source fields, the slice is bit-identical to scipy's.
//
// Bound: HBM write of X*Z*48 B per slice (the source, <= 300x300x5 doubles, stays in L2).
#include <limits.h>
#include "common.cuh"

namespace dfcsr {

constexpr int kRegridThreads = 256;   // threads per block; a block owns kRegridRows whole rows and strides along z
constexpr int kRegridRows = 2;        // rows (x) per block

struct AxisCell {
    int i;
    double y;
    int outside;
};

// searchsorted(g, q, 'right') - 1 on linspace nodes: arithmetic guess, then exact fix-up against the
// bit-exact node values.
__device__ inline AxisCell locate(const Axis& g, double q) {
    AxisCell c;
    const int n = g.n;
    double g0 = g.start, gl = g.stop;
    c.outside = (q < g0) || (q > gl) || !(q == q);
    int i = 0;
    if (g.step > 0.0 && q == q) {
        double guess = floor((q - g0) / g.step);
        i = (guess < 0.0) ? 0 : ((guess > (double)(n - 2)) ? n - 2 : (int)guess);
        while (i > 0 && axis_node(g, i) > q) --i;              // need g[i] <= q
        while (i < n - 2 && axis_node(g, i + 1) <= q) ++i;     // and q < g[i+1]
    }
    c.i = i;
    double a = axis_node(g, i), b = axis_node(g, i + 1);
    c.y = __ddiv_rn(__dsub_rn(q, a), __dsub_rn(b, a));
    return c;
}

// one voxel to a slice: 48-byte fp64 record or 32-byte fp32 record
template <bool kF32>
__device__ __forceinline__ void store_voxel(void* slice, size_t cell, const double (&v)[5]) {
    if (kF32) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(slice) + cell * DFCSR_VOXEL_FLOATS);
        dst[0] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
        dst[1] = make_float4((float)v[4], 0.f, 0.f, 0.f);
    } else {
        double2* dst = reinterpret_cast<double2*>(reinterpret_cast<double*>(slice) + cell * DFCSR_VOXEL_DOUBLES);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
        dst[2] = make_double2(v[4], 0.0);
    }
}

// Does a voxel count for the row support?  Any non-zero (or NaN) density / density gradient, or a velocity field that is
// not finite: the reference turns 0 * inf into NaN, so such a sample must be evaluated, not skipped.  In the fp32 format
// the test is made on the values as stored.
template <bool kF32>
__device__ __forceinline__ bool voxel_supports(const double (&v)[5]) {
    if (kF32) {
        const float a = (float)v[0], b = (float)v[1], c = (float)v[2], d = (float)v[3], e = (float)v[4];
        return !(a == 0.f) || !(b == 0.f) || !(c == 0.f) || !isfinite(d) || !isfinite(e);
    }
    return !(v[0] == 0.0) || !(v[1] == 0.0) || !(v[2] == 0.0) || !isfinite(v[3]) || !isfinite(v[4]);
}

template <bool kF32>
__device__ __forceinline__ void load_voxel(const void* slice, size_t cell, double (&v)[5]) {
    if (kF32) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(slice) + cell * DFCSR_VOXEL_FLOATS);
        float4 a = src[0], b = src[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x;
    } else {
        const double2* src = reinterpret_cast<const double2*>(reinterpret_cast<const double*>(slice) + cell * DFCSR_VOXEL_DOUBLES);
        double2 a = src[0], b = src[1], d = src[2];
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = d.x;
    }
}

// One block re-grids kRegridRows whole rows: every thread resolves the source cell of its columns (z), the rows'
// source cells come from shared memory.  Because a block sees its rows completely, it also knows their row support
// (hull of the voxels with non-zero density or density gradient, or non-finite velocity fields): one store per row,
// no second pass over the slice.
template <bool kF32>
__global__ void __launch_bounds__(kRegridThreads)
regrid_kernel(const double* __restrict__ src, Axis sx, Axis sz, Axis dx, Axis dz, double fill_vx_x,
              const double* __restrict__ fill_ptr, void* __restrict__ slice, int* __restrict__ support) {
    __shared__ AxisCell rows[kRegridRows];
    __shared__ int hull[kRegridRows][2][kRegridThreads / 32];
    if (fill_ptr) fill_vx_x = __ldg(fill_ptr);
    const int row0 = blockIdx.x * kRegridRows;
    if (threadIdx.x < kRegridRows && row0 + threadIdx.x < dx.n)
        rows[threadIdx.x] = locate(sx, axis_node(dx, row0 + threadIdx.x));
    __syncthreads();
    const size_t plane = (size_t)sx.n * sz.n;
    const int rmax = min(kRegridRows, dx.n - row0);
    int lo[kRegridRows], hi[kRegridRows];
#pragma unroll
    for (int r = 0; r < kRegridRows; ++r) { lo[r] = INT_MAX; hi[r] = -1; }
    for (int col = threadIdx.x; col < dz.n; col += kRegridThreads) {
        const AxisCell cz = locate(sz, axis_node(dz, col));
        const double wz1 = cz.y, wz0 = __dsub_rn(1.0, cz.y);
#pragma unroll
        for (int r = 0; r < kRegridRows; ++r) {
            if (r >= rmax) break;
            const AxisCell cx = rows[r];
            double out[5];
            if (cx.outside || cz.outside) {
                out[0] = out[1] = out[2] = out[3] = 0.0;
                out[4] = fill_vx_x;
            } else {
                const double wx1 = cx.y, wx0 = __dsub_rn(1.0, cx.y);
                const size_t o = (size_t)cx.i * sz.n + cz.i;
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    // scipy's evaluate_linear_2d order: ((F * wx) * wz), corners accumulated in sequence
                    const double* p = src + f * plane + o;
                    double v = __dmul_rn(__dmul_rn(__ldg(p), wx0), wz0);
                    v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(p + 1), wx0), wz1));
                    v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(p + sz.n), wx1), wz0));
                    v = __dadd_rn(v, __dmul_rn(__dmul_rn(__ldg(p + sz.n + 1), wx1), wz1));
                    out[f] = v;
                }
            }
            store_voxel<kF32>(slice, (size_t)(row0 + r) * dz.n + col, out);
            if (voxel_supports<kF32>(out)) { lo[r] = min(lo[r], col); hi[r] = max(hi[r], col); }
        }
    }
    if (!support) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < kRegridRows; ++r) {
        const int l = __reduce_min_sync(0xffffffffu, lo[r]), h = __reduce_max_sync(0xffffffffu, hi[r]);
        if (lane == 0) { hull[r][0][warp] = l; hull[r][1][warp] = h; }
    }
    __syncthreads();
    if (threadIdx.x < rmax) {
        int l = INT_MAX, h = -1;
        for (int w = 0; w < kRegridThreads / 32; ++w) { l = min(l, hull[threadIdx.x][0][w]); h = max(h, hull[threadIdx.x][1][w]); }
        support[2 * (row0 + threadIdx.x)] = l;
        support[2 * (row0 + threadIdx.x) + 1] = h;
    }
}

template <bool kF32>
__global__ void pack_kernel(const double* __restrict__ fields, long long cells, void* __restrict__ slice) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells;
         c += (long long)gridDim.x * blockDim.x) {
        double v[5];
#pragma unroll
        for (int f = 0; f < 5; ++f) v[f] = fields[f * cells + c];
        store_voxel<kF32>(slice, (size_t)c, v);
    }
}

template <bool kF32>
__global__ void unpack_kernel(const void* __restrict__ slice, long long cells, double* __restrict__ fields) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells;
         c += (long long)gridDim.x * blockDim.x) {
        double v[5];
        load_voxel<kF32>(slice, (size_t)c, v);
#pragma unroll
        for (int f = 0; f < 5; ++f) fields[f * cells + c] = v[f];
    }
}

static inline unsigned blocks_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_history_regrid(const double* d_fields, dfcsr_axis src_x, dfcsr_axis src_z,
                                    dfcsr_axis dst_x, dfcsr_axis dst_z, double fill_vx_x,
                                    const double* d_fill_vx_x, int32_t format, void* d_slice, int32_t* d_row_support,
                                    void* stream) {
    DFCSR_REQUIRE(d_fields && d_slice, "null pointer");
    DFCSR_REQUIRE(src_x.n >= 2 && src_z.n >= 2 && dst_x.n >= 1 && dst_z.n >= 1, "axes too short");
    Axis sx = make_axis(src_x.start, src_x.stop, src_x.n), sz = make_axis(src_z.start, src_z.stop, src_z.n);
    Axis dx = make_axis(dst_x.start, dst_x.stop, dst_x.n), dz = make_axis(dst_z.start, dst_z.stop, dst_z.n);
    const unsigned grid = (unsigned)((dx.n + kRegridRows - 1) / kRegridRows);
    DFCSR_REQUIRE(format == DFCSR_VOXEL_F64 || format == DFCSR_VOXEL_F32, "unknown voxel format");
    if (format == DFCSR_VOXEL_F32)
        regrid_kernel<true><<<grid, kRegridThreads, 0, as_stream(stream)>>>(d_fields, sx, sz, dx, dz, fill_vx_x, d_fill_vx_x, d_slice, d_row_support);
    else
        regrid_kernel<false><<<grid, kRegridThreads, 0, as_stream(stream)>>>(d_fields, sx, sz, dx, dz, fill_vx_x, d_fill_vx_x, d_slice, d_row_support);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

// Row support of a voxel slice: per transverse row the hull [z_lo, z_hi] of voxels with a non-zero density or density
// gradient (dfcsr_history.d_row_support).  One warp per row, lanes stride along z.
template <bool kF32>
__global__ void __launch_bounds__(256)
row_support_kernel(const void* __restrict__ slice, int X, int Z, int* __restrict__ support) {
    const int lane = threadIdx.x & 31;
    const int warps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
    for (int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5); row < X; row += warps) {
        int lo = INT_MAX, hi = -1;
        for (int z = lane; z < Z; z += 32) {
            double v[5];
            load_voxel<kF32>(slice, (size_t)row * Z + z, v);
            const bool nz = voxel_supports<kF32>(v);
            if (nz) { lo = min(lo, z); hi = max(hi, z); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { support[2 * row] = lo; support[2 * row + 1] = hi; }
    }
}

extern "C" int dfcsr_history_row_support(const void* d_slice, int32_t X, int32_t Z, int32_t format, int32_t* d_support,
                                         void* stream) {
    DFCSR_REQUIRE(d_slice && d_support && X > 0 && Z > 0, "bad argument");
    DFCSR_REQUIRE(format == DFCSR_VOXEL_F64 || format == DFCSR_VOXEL_F32, "unknown voxel format");
    const long long threads = (long long)X * 32;
    if (format == DFCSR_VOXEL_F32) row_support_kernel<true><<<blocks_for(threads, 256), 256, 0, as_stream(stream)>>>(d_slice, X, Z, d_support);
    else row_support_kernel<false><<<blocks_for(threads, 256), 256, 0, as_stream(stream)>>>(d_slice, X, Z, d_support);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_history_pack(const double* d_fields, int32_t X, int32_t Z, int32_t format, void* d_slice, void* stream) {
    DFCSR_REQUIRE(d_fields && d_slice && X > 0 && Z > 0, "bad argument");
    DFCSR_REQUIRE(format == DFCSR_VOXEL_F64 || format == DFCSR_VOXEL_F32, "unknown voxel format");
    long long cells = (long long)X * Z;
    if (format == DFCSR_VOXEL_F32) pack_kernel<true><<<blocks_for(cells, 256), 256, 0, as_stream(stream)>>>(d_fields, cells, d_slice);
    else pack_kernel<false><<<blocks_for(cells, 256), 256, 0, as_stream(stream)>>>(d_fields, cells, d_slice);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_history_unpack(const void* d_slice, int32_t X, int32_t Z, int32_t format, double* d_fields, void* stream) {
    DFCSR_REQUIRE(d_fields && d_slice && X > 0 && Z > 0, "bad argument");
    DFCSR_REQUIRE(format == DFCSR_VOXEL_F64 || format == DFCSR_VOXEL_F32, "unknown voxel format");
    long long cells = (long long)X * Z;
    if (format == DFCSR_VOXEL_F32) unpack_kernel<true><<<blocks_for(cells, 256), 256, 0, as_stream(stream)>>>(d_slice, cells, d_fields);
    else unpack_kernel<false><<<blocks_for(cells, 256), 256, 0, as_stream(stream)>>>(d_slice, cells, d_fields);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
