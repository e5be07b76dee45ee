// deposit.cu — K1: particle-to-grid deposition.
//
// Replaces histogram_cic_2d (deposit.py:42-87), which the reference calls twice per step with
// identical indices (weights 1 and px, deposit.py:172-182): here ONE pass over (x, z, px) deposits
// both weights.  Semantics kept: bin spacing (end-start)/nbins, c = (q-start)*inv_spacing,
// i = floor(c), S_low = 1-(c-i), and each of the four corner updates guarded separately by
// 0 <= index < nbins, so particles near the edge deposit partially.
//
// Two code paths
//   private : each CTA owns a full copy of both grids in shared memory (fits for the reference's
//             hard-coded 100x100 branch, deposit.py:164-167: 2*100*100*8 B = 160 KB of the 227 KB),
//             lanes that hit the same lower-left cell are combined with a warp match before the
//             shared-memory atomics, and non-zero cells are flushed with fp64 L2 reductions;
//   direct  : fp64 reductions (RED.ADD.F64) straight into the L2-resident global grids, for grids
//             that do not fit in shared memory (YAML grids such as 300x300 or 64x2048).
// fp64 atomics make the summation order run-dependent (last-bit differences only).
// Bound: HBM read of 24 B / particle.
//
// NGP (nearest grid point) is an extension with no reference counterpart:
// int64 counts, bit-exact for any order.
#include "common.cuh"

namespace dfcsr {

struct DepGrid {
    int nx, nz;
    double x_start, inv_dx, z_start, inv_dz;
};

struct CicSample {
    int i, j;
    double a_lo, b_lo;
};

__device__ __forceinline__ CicSample cic_cell(const DepGrid& g, double x, double z) {
    CicSample s;
    double cx = (x - g.x_start) * g.inv_dx;
    double cz = (z - g.z_start) * g.inv_dz;
    double fx = floor(cx), fz = floor(cz);
    // clamp far-away particles before the int conversion (they fail every guard anyway)
    fx = fmin(fmax(fx, -2.0), (double)g.nx + 1.0);
    fz = fmin(fmax(fz, -2.0), (double)g.nz + 1.0);
    s.i = (int)fx;
    s.j = (int)fz;
    s.a_lo = 1.0 - (cx - fx);
    s.b_lo = 1.0 - (cz - fz);
    if (!(cx == cx) || !(cz == cz)) { s.i = -2; s.j = -2; }
    return s;
}

__global__ void __launch_bounds__(256)
cic_direct_kernel(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ px,
                  long long n, DepGrid g, double* __restrict__ count, double* __restrict__ vxsum) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        CicSample s = cic_cell(g, x[p], z[p]);
        const double w = px[p];
        const double a_hi = 1.0 - s.a_lo, b_hi = 1.0 - s.b_lo;
        const bool i0 = (s.i >= 0) && (s.i < g.nx), i1 = (s.i + 1 >= 0) && (s.i + 1 < g.nx);
        const bool j0 = (s.j >= 0) && (s.j < g.nz), j1 = (s.j + 1 >= 0) && (s.j + 1 < g.nz);
        const long long o = (long long)s.i * g.nz + s.j;
        if (i0 && j0) { double c = s.a_lo * s.b_lo; atomicAdd(count + o, c); atomicAdd(vxsum + o, w * c); }
        if (i0 && j1) { double c = s.a_lo * b_hi;   atomicAdd(count + o + 1, c); atomicAdd(vxsum + o + 1, w * c); }
        if (i1 && j0) { double c = a_hi * s.b_lo;   atomicAdd(count + o + g.nz, c); atomicAdd(vxsum + o + g.nz, w * c); }
        if (i1 && j1) { double c = a_hi * b_hi;     atomicAdd(count + o + g.nz + 1, c); atomicAdd(vxsum + o + g.nz + 1, w * c); }
    }
}

// Block-private tiles: both grids live in dynamic shared memory.
__global__ void __launch_bounds__(1024, 1)
cic_private_kernel(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ px,
                   long long n, DepGrid g, double* __restrict__ count, double* __restrict__ vxsum) {
    extern __shared__ double tile[];
    const int cells = g.nx * g.nz;
    double* t_count = tile;
    double* t_vx = tile + cells;
    for (int c = threadIdx.x; c < 2 * cells; c += blockDim.x) tile[c] = 0.0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (n + stride - 1) / stride;   // uniform trip count: the warp match needs all lanes
    for (long long rnd = 0; rnd < rounds; ++rnd) {
        const long long p = rnd * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live = p < n;
        CicSample s;
        s.i = -2; s.j = -2; s.a_lo = 0.0; s.b_lo = 0.0;
        double w = 0.0;
        if (live) {
            s = cic_cell(g, x[p], z[p]);
            w = px[p];
        }
        const double a_hi = 1.0 - s.a_lo, b_hi = 1.0 - s.b_lo;
        double c00 = s.a_lo * s.b_lo, c01 = s.a_lo * b_hi, c10 = a_hi * s.b_lo, c11 = a_hi * b_hi;
        double v00 = w * c00, v01 = w * c01, v10 = w * c10, v11 = w * c11;
        // warp aggregation: lanes sharing the lower-left cell form a group; the lowest lane of the
        // group collects the group's eight partial weights and issues the atomics.
        const int key = live ? (s.i + 2) * (g.nz + 4) + (s.j + 2) : -1 - (int)(threadIdx.x & 31);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        const int lane = threadIdx.x & 31;
        if (peers != (1u << lane)) {
            // all lanes of a group run the same shuffles (same mask, same trip count); only the
            // leader keeps the sums
            unsigned rest = peers & ~(1u << leader);
            const int steps = __popc(rest);
            for (int k = 0; k < steps; ++k) {
                int src = __ffs(rest) - 1;
                rest &= rest - 1;
                double t;
                t = __shfl_sync(peers, c00, src); if (lane == leader) c00 += t;
                t = __shfl_sync(peers, c01, src); if (lane == leader) c01 += t;
                t = __shfl_sync(peers, c10, src); if (lane == leader) c10 += t;
                t = __shfl_sync(peers, c11, src); if (lane == leader) c11 += t;
                t = __shfl_sync(peers, v00, src); if (lane == leader) v00 += t;
                t = __shfl_sync(peers, v01, src); if (lane == leader) v01 += t;
                t = __shfl_sync(peers, v10, src); if (lane == leader) v10 += t;
                t = __shfl_sync(peers, v11, src); if (lane == leader) v11 += t;
            }
        }
        if (live && lane == leader) {
            const bool i0 = (s.i >= 0) && (s.i < g.nx), i1 = (s.i + 1 >= 0) && (s.i + 1 < g.nx);
            const bool j0 = (s.j >= 0) && (s.j < g.nz), j1 = (s.j + 1 >= 0) && (s.j + 1 < g.nz);
            const int o = s.i * g.nz + s.j;
            if (i0 && j0) { atomicAdd(t_count + o, c00); atomicAdd(t_vx + o, v00); }
            if (i0 && j1) { atomicAdd(t_count + o + 1, c01); atomicAdd(t_vx + o + 1, v01); }
            if (i1 && j0) { atomicAdd(t_count + o + g.nz, c10); atomicAdd(t_vx + o + g.nz, v10); }
            if (i1 && j1) { atomicAdd(t_count + o + g.nz + 1, c11); atomicAdd(t_vx + o + g.nz + 1, v11); }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        double a = t_count[c], b = t_vx[c];
        if (a != 0.0) atomicAdd(count + c, a);
        if (b != 0.0) atomicAdd(vxsum + c, b);
    }
}

__global__ void __launch_bounds__(256)
ngp_kernel(const double* __restrict__ x, const double* __restrict__ z, long long n, DepGrid g,
           unsigned long long* __restrict__ count) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double cx = __dadd_rn(__dmul_rn(__dsub_rn(x[p], g.x_start), g.inv_dx), 0.5);
        double cz = __dadd_rn(__dmul_rn(__dsub_rn(z[p], g.z_start), g.inv_dz), 0.5);
        double fx = floor(cx), fz = floor(cz);
        if (fx >= 0.0 && fx < (double)g.nx && fz >= 0.0 && fz < (double)g.nz)
            atomicAdd(count + (long long)fx * g.nz + (long long)fz, 1ULL);
    }
}

static DepGrid make_grid(int nx, double xs, double xe, int nz, double zs, double ze) {
    DepGrid g;
    g.nx = nx;
    g.nz = nz;
    g.x_start = xs;
    g.z_start = zs;
    g.inv_dx = 1.0 / ((xe - xs) / (double)nx);   // deposit.py:55-58
    g.inv_dz = 1.0 / ((ze - zs) / (double)nz);
    return g;
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_deposit_cic(const double* d_x, const double* d_z, const double* d_px, int64_t n,
                                 int32_t nx, double x_start, double x_end, int32_t nz, double z_start,
                                 double z_end, double* d_count, double* d_vxsum, int32_t mode, void* stream) {
    DFCSR_REQUIRE(d_count && d_vxsum && (n == 0 || (d_x && d_z && d_px)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n >= 0, "bad sizes");
    DFCSR_REQUIRE((long long)nx * nz < (1LL << 30), "grid too large");
    cudaStream_t st = as_stream(stream);
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_count, 0, cells * sizeof(double), st));
    DFCSR_CUDA_OK(cudaMemsetAsync(d_vxsum, 0, cells * sizeof(double), st));
    if (n == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, x_start, x_end, nz, z_start, z_end);
    const size_t smem = 2 * cells * sizeof(double);
    const bool fits = smem <= 200 * 1024;
    if (mode == 0) mode = (fits && n >= (4LL << 20)) ? 1 : 2;
    if (mode == 1) {
        if (!fits) {
            set_error("dfcsr_deposit_cic: %dx%d grid does not fit the shared-memory tile path", nx, nz);
            return DFCSR_ERR_UNSUPPORTED;
        }
        DFCSR_CUDA_OK(cudaFuncSetAttribute(cic_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long want = (n + 1024LL * 32 - 1) / (1024LL * 32);
        unsigned blocks = (unsigned)(want < 148 ? (want < 1 ? 1 : want) : 148);
        cic_private_kernel<<<blocks, 1024, smem, st>>>(d_x, d_z, d_px, n, g, d_count, d_vxsum);
    count_launch(1);
    } else if (mode == 2) {
        long long want = (n + 255) / 256;
        unsigned blocks = (unsigned)(want < 148LL * 8 ? want : 148LL * 8);
        cic_direct_kernel<<<blocks, 256, 0, st>>>(d_x, d_z, d_px, n, g, d_count, d_vxsum);
    count_launch(1);
    } else {
        DFCSR_REQUIRE(false, "mode must be 0, 1 or 2");
    }
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_deposit_ngp(const double* d_x, const double* d_z, int64_t n, int32_t nx, double x_start,
                                 double x_end, int32_t nz, double z_start, double z_end, int64_t* d_count,
                                 void* stream) {
    DFCSR_REQUIRE(d_count && (n == 0 || (d_x && d_z)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n >= 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    DFCSR_CUDA_OK(cudaMemsetAsync(d_count, 0, (size_t)nx * nz * sizeof(int64_t), st));
    if (n == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, x_start, x_end, nz, z_start, z_end);
    long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 8 ? want : 148LL * 8);
    ngp_kernel<<<blocks, 256, 0, st>>>(d_x, d_z, n, g, reinterpret_cast<unsigned long long*>(d_count));
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
