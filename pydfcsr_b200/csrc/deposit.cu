// deposit.cu — K1: particle-to-grid deposition.
//
// Replaces histogram_cic_2d (deposit.py:42-87), which the reference calls twice per step with
// identical indices (weights 1 and px, deposit.py:172-182): here ONE pass over (x, z, px) deposits
// both weights.  Semantics kept: bin spacing (end-start)/nbins, c = (q-start)*inv_spacing,
// i = floor(c), S_low = 1-(c-i), and each of the four corner updates guarded separately by
// 0 <= index < nbins, so particles near the edge deposit partially.
//
// Code paths
//   tile   : every CTA keeps a private copy of a centred sub-rectangle of BOTH grids in shared memory
//            (up to 12288 cells = 192 KB: the whole grid for the reference's hard-coded 100x100 branch,
//            deposit.py:164-167; the central ~+-1.9 sigma, i.e. ~88 % of a Gaussian bunch, for a 300x300
//            YAML grid).  Updates inside the tile are shared-memory fp64 atomics, optionally after a
//            warp match that merges lanes hitting the same cell; the few updates outside go straight
//            to L2 (fp64 RED).  Non-zero tile cells are flushed with fp64 L2 reductions at the end.
//   direct : fp64 reductions (RED.ADD.F64) into the L2-resident global grids for every update.
// fp64 atomics make the summation order run-dependent (last-bit differences only).
//
// Bound: 24 B / particle of HBM reads, but the unit that saturates first is the atomic path: 8 fp64
// updates per particle at ~1 shared-memory CAS update per clock per SM (shared memory has no native
// fp64/fp32 add on sm_100a: ATOMS.CAST.SPIN) — see DESIGN.md §4.
//
// NGP (nearest grid point) is an extension with no reference counterpart:
// int64 counts, bit-exact for any order.
#include "common.cuh"

namespace dfcsr {

constexpr int kTileCells = 12288;      // 2 grids x 8 B x 12288 = 192 KB of the 227 KB per CTA
constexpr int kTileThreads = 1024;
constexpr long long kParticlesPerCta = 8192;

struct DepGrid {
    int nx, nz;
    double x_start, inv_dx, z_start, inv_dz;
};

struct Tile {
    int i0, j0, ni, nj;   // sub-rectangle [i0, i0+ni) x [j0, j0+nj) of the grid
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

// (A {count, vx} pair update with one 128-bit shared CAS, ATOMS.CAS.128, was measured: the retry loop is
// 3-4x slower than two ATOMS.CAST.SPIN.64 on the hot central cells, so it is not used.)
// one corner update: shared-memory tile if the cell is inside it, L2 reduction otherwise
__device__ __forceinline__ void corner(const DepGrid& g, const Tile& t, double* t_count, double* t_vx,
                                       double* __restrict__ count, double* __restrict__ vxsum, int i, int j,
                                       double c, double v) {
    if (i < 0 || i >= g.nx || j < 0 || j >= g.nz) return;     // the reference's per-corner guards
    const int ti = i - t.i0, tj = j - t.j0;
    if (ti >= 0 && ti < t.ni && tj >= 0 && tj < t.nj) {
        const int o = ti * t.nj + tj;
        atomicAdd(t_count + o, c);
        atomicAdd(t_vx + o, v);
    } else {
        const long long o = (long long)i * g.nz + j;
        atomicAdd(count + o, c);
        atomicAdd(vxsum + o, v);
    }
}

template <bool kAggregate>
__global__ void __launch_bounds__(kTileThreads, 1)
cic_tile_kernel(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ px,
                long long n, DepGrid g, Tile t, double* __restrict__ count, double* __restrict__ vxsum) {
    extern __shared__ double tile[];
    const int cells = t.ni * t.nj;
    double* t_count = tile;
    double* t_vx = tile + cells;
    for (int c = threadIdx.x; c < 2 * cells; c += blockDim.x) tile[c] = 0.0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (n + stride - 1) / stride;   // uniform trip count: the warp match needs all lanes
    const int lane = threadIdx.x & 31;
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
        bool owner = live;
        if (kAggregate) {
            // warp aggregation: lanes sharing the lower-left cell form a group; the lowest lane of the
            // group collects the group's eight partial weights and issues the atomics.  All lanes of a
            // group run the same shuffles (same mask, same trip count).
            const int key = live ? (s.i + 2) * (g.nz + 4) + (s.j + 2) : -1 - lane;
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(peers) - 1;
            owner = live && (lane == leader);
            if (peers != (1u << lane)) {
                unsigned rest = peers & ~(1u << leader);
                const int steps = __popc(rest);
                for (int k = 0; k < steps; ++k) {
                    int src = __ffs(rest) - 1;
                    rest &= rest - 1;
                    double u;
                    u = __shfl_sync(peers, c00, src); if (lane == leader) c00 += u;
                    u = __shfl_sync(peers, c01, src); if (lane == leader) c01 += u;
                    u = __shfl_sync(peers, c10, src); if (lane == leader) c10 += u;
                    u = __shfl_sync(peers, c11, src); if (lane == leader) c11 += u;
                    u = __shfl_sync(peers, v00, src); if (lane == leader) v00 += u;
                    u = __shfl_sync(peers, v01, src); if (lane == leader) v01 += u;
                    u = __shfl_sync(peers, v10, src); if (lane == leader) v10 += u;
                    u = __shfl_sync(peers, v11, src); if (lane == leader) v11 += u;
                }
            }
        }
        if (owner) {
            corner(g, t, t_count, t_vx, count, vxsum, s.i, s.j, c00, v00);
            corner(g, t, t_count, t_vx, count, vxsum, s.i, s.j + 1, c01, v01);
            corner(g, t, t_count, t_vx, count, vxsum, s.i + 1, s.j, c10, v10);
            corner(g, t, t_count, t_vx, count, vxsum, s.i + 1, s.j + 1, c11, v11);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const int ti = c / t.nj, tj = c - ti * t.nj;
        const long long o = (long long)(t.i0 + ti) * g.nz + (t.j0 + tj);
        double a = t_count[c], b = t_vx[c];
        if (a != 0.0) atomicAdd(count + o, a);
        if (b != 0.0) atomicAdd(vxsum + o, b);
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

// NGP with a block-private int32 tile in shared memory (native ATOMS.ADD, no CAS loop): per-CTA counts cannot
// overflow 32 bits (a CTA sees far fewer than 2^31 particles); cells outside the centred tile and the final
// flush use 64-bit L2 atomics.  Integer adds commute: the result is bit-exact and run-to-run reproducible.
constexpr int kNgpTileCells = 49152;   // 192 KB of int32

__global__ void __launch_bounds__(kTileThreads, 1)
ngp_tile_kernel(const double* __restrict__ x, const double* __restrict__ z, long long n, DepGrid g, Tile t,
                unsigned long long* __restrict__ count) {
    extern __shared__ int ntile[];
    const int cells = t.ni * t.nj;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) ntile[c] = 0;
    __syncthreads();
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double cx = __dadd_rn(__dmul_rn(__dsub_rn(x[p], g.x_start), g.inv_dx), 0.5);
        double cz = __dadd_rn(__dmul_rn(__dsub_rn(z[p], g.z_start), g.inv_dz), 0.5);
        double fx = floor(cx), fz = floor(cz);
        if (fx >= 0.0 && fx < (double)g.nx && fz >= 0.0 && fz < (double)g.nz) {
            const int i = (int)fx, j = (int)fz;
            const int ti = i - t.i0, tj = j - t.j0;
            if (ti >= 0 && ti < t.ni && tj >= 0 && tj < t.nj) atomicAdd(ntile + ti * t.nj + tj, 1);
            else atomicAdd(count + (long long)i * g.nz + j, 1ULL);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const int v = ntile[c];
        if (v) {
            const int ti = c / t.nj, tj = c - ti * t.nj;
            atomicAdd(count + (long long)(t.i0 + ti) * g.nz + (t.j0 + tj), (unsigned long long)v);
        }
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

// centred sub-rectangle with the grid's aspect ratio and at most max_cells cells
static Tile make_tile(int nx, int nz, int max_cells = kTileCells) {
    const int kTileCells = max_cells;
    Tile t;
    if ((long long)nx * nz <= kTileCells) {
        t.i0 = 0; t.j0 = 0; t.ni = nx; t.nj = nz;
        return t;
    }
    double f = sqrt((double)kTileCells / ((double)nx * (double)nz));
    int ni = (int)(nx * f);
    ni = ni < 1 ? 1 : (ni > nx ? nx : ni);
    int nj = kTileCells / ni;
    nj = nj > nz ? nz : nj;
    t.ni = ni; t.nj = nj;
    t.i0 = (nx - ni) / 2;
    t.j0 = (nz - nj) / 2;
    return t;
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_deposit_cic(const double* d_x, const double* d_z, const double* d_px, int64_t n,
                                 int32_t nx, double x_start, double x_end, int32_t nz, double z_start,
                                 double z_end, double* d_count, double* d_vxsum, int32_t mode, void* stream) {
    DFCSR_REQUIRE(d_count && d_vxsum && (n == 0 || (d_x && d_z && d_px)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n >= 0, "bad sizes");
    DFCSR_REQUIRE((long long)nx * nz < (1LL << 30), "grid too large");
    DFCSR_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0 (auto), 1 (tile + warp match), 2 (direct) or 3 (tile)");
    cudaStream_t st = as_stream(stream);
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_count, 0, cells * sizeof(double), st));
    DFCSR_CUDA_OK(cudaMemsetAsync(d_vxsum, 0, cells * sizeof(double), st));
    if (n == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, x_start, x_end, nz, z_start, z_end);
    if (mode == 0) mode = (n >= 65536) ? 3 : 2;
    if (mode == 2) {
        long long want = (n + 255) / 256;
        unsigned blocks = (unsigned)(want < 148LL * 8 ? want : 148LL * 8);
        cic_direct_kernel<<<blocks, 256, 0, st>>>(d_x, d_z, d_px, n, g, d_count, d_vxsum);
    } else {
        Tile t = make_tile(nx, nz);
        const size_t smem = (size_t)2 * t.ni * t.nj * sizeof(double);
        long long want = (n + kParticlesPerCta - 1) / kParticlesPerCta;
        unsigned blocks = (unsigned)(want < 148 ? (want < 1 ? 1 : want) : 148);
#define DFCSR_LAUNCH_TILE(AGG)                                                                               \
    do {                                                                                                     \
        DFCSR_CUDA_OK(cudaFuncSetAttribute(cic_tile_kernel<AGG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                      \
        cic_tile_kernel<AGG><<<blocks, kTileThreads, smem, st>>>(d_x, d_z, d_px, n, g, t, d_count, d_vxsum);  \
    } while (0)
        if (mode == 1) DFCSR_LAUNCH_TILE(true);
        else DFCSR_LAUNCH_TILE(false);
#undef DFCSR_LAUNCH_TILE
    }
    count_launch(1);
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
    if (n >= 65536) {
        Tile t = make_tile(nx, nz, kNgpTileCells);
        const size_t smem = (size_t)t.ni * t.nj * sizeof(int);
        long long want = (n + kParticlesPerCta - 1) / kParticlesPerCta;
        unsigned blocks = (unsigned)(want < 148 ? want : 148);
        DFCSR_CUDA_OK(cudaFuncSetAttribute(ngp_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ngp_tile_kernel<<<blocks, kTileThreads, smem, st>>>(d_x, d_z, n, g, t, reinterpret_cast<unsigned long long*>(d_count));
    } else {
        long long want = (n + 255) / 256;
        unsigned blocks = (unsigned)(want < 148LL * 8 ? want : 148LL * 8);
        ngp_kernel<<<blocks, 256, 0, st>>>(d_x, d_z, n, g, reinterpret_cast<unsigned long long*>(d_count));
    }
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
