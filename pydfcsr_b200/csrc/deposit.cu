// deposit.cu — K1: particle-to-grid deposition.
//
// Replaces histogram_cic_2d (deposit.py:42-87), which the reference calls twice per step with
// identical indices (weights 1 and px, deposit.py:172-182): here ONE pass over (x, z, px) deposits
// both weights.  Semantics kept: bin spacing (end-start)/nbins, c = (q-start)*inv_spacing,
// i = floor(c), S_low = 1-(c-i), and each of the four corner updates guarded separately by
// 0 <= index < nbins, so particles near the edge deposit partially.
//
// Code paths
//   fixed-point (default, modes 4/5): 64-bit integer accumulation, bit-reproducible; see below.
//   fp64 tile (modes 1/3): every CTA keeps a private copy of a centred sub-rectangle of BOTH grids in shared
//            memory (up to 14464 cells = 226 KB: the whole grid for the reference's hard-coded 100x100 branch,
//            deposit.py:164-167; the central ~+-2 sigma, i.e. ~91 % of a Gaussian bunch, for a 300x300
//            YAML grid).  Updates inside the tile are shared-memory fp64 atomics, optionally after a
//            warp match that merges lanes hitting the same cell; the few updates outside go straight
//            to L2 (fp64 RED).  Non-zero tile cells are flushed with fp64 L2 reductions at the end.
//   fp64 direct (mode 2): fp64 reductions (RED.ADD.F64) into the L2-resident global grids for every update.
// fp64 atomics make the summation order run-dependent (last-bit differences only).
//
// Bound: 24 B / particle of HBM reads, but the unit that saturates first is the shared-memory atomic path
// (8 updates per particle): fp64 adds are CAS loops (ATOMS.CAST.SPIN.64, ~1.4 updates/clk/SM on a bunch),
// the fixed-point limbs use the native 32-bit ATOMS.ADD (~3.3 updates/clk/SM) — see DESIGN.md §4.
//
// NGP (nearest grid point) is an extension with no reference counterpart:
// int64 counts, bit-exact for any order.
#include <atomic>
#include <cmath>
#include <cstring>
#include <math_constants.h>

#include "common.cuh"

namespace dfcsr {

constexpr int kTileCells = 14464;      // 2 grids x 8 B x 14464 = 226 KB of the 227 KB per CTA
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

// ------------------------------------------------------------------------------------------------------
// Fixed-point path (the default).  Shared memory on sm_100a has a native 32-bit integer add (ATOMS.ADD,
// 8.5 updates/clk/SM measured with tools/atoms_probe.cu) but fp64/u64 adds are CAS loops (1.8/clk/SM, less
// on hot cells).  Each cell is therefore a 64-bit two's-complement fixed-point number held as two 32-bit
// limbs: add the low limb (the returned old value gives the carry), then add high limb + carry
// (4.1 updates/clk/SM).  Integer adds commute, so the result is bit-reproducible from run to run and
// identical on every rank, unlike fp64 atomics.
//
// Scales (powers of two, so the scaling itself is exact):
//   tile   : a CTA sees at most `chunk` particles and a particle adds at most 1 (count) or max|px| (velocity
//            sum) to one cell, so F_t = 62 - ceil(log2(chunk+1)) fraction bits cannot overflow;
//   global : same with chunk -> n, F_g <= F_t; a tile is shifted (rounded) to F_g when it is flushed with
//            64-bit integer L2 reductions into the output buffers, which a last kernel converts to fp64 in place.
// max|px| comes from a reduction pass over px (absmax_kernel) into a device scratch slot.
// Norm-wise error of a cell sum: ~4 * 2^-62 / (largest cell's share of the bunch), e.g. 1e-14 on 300x300.
constexpr int kScratchSlots = 64;
__device__ unsigned long long g_absmax[kScratchSlots];

__global__ void __launch_bounds__(256)
absmax_kernel(const double* __restrict__ v, long long n, unsigned long long* __restrict__ slot) {
    double m = 0.0;
    bool bad = false;
#pragma unroll 8
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const double a = fabs(v[p]);
        bad |= !(a == a);
        m = fmax(m, a);
    }
    m = warp_max(m);
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0)     // non-negative doubles order like their bit patterns; NaN sorts above +inf
        atomicMax(slot, bad ? 0x7ff8000000000000ULL : (unsigned long long)__double_as_longlong(m));
}

struct FixedScales {
    int f_tile, f_glob;          // fraction bits of the count grid in the tile / in the global buffer
};

struct VxScale {
    double tile, glob;           // 2^(f_tile - e), 2^(f_glob - e) with max|px| < 2^e; 0 if px is all zero or not finite
};

// device twin of absmax_bits(): a statistic that is negative, infinite or NaN means "not finite" (vxsum becomes NaN)
__device__ __forceinline__ unsigned long long finite_or_nan_bits(unsigned long long bits) {
    const double v = __longlong_as_double((long long)bits);
    return (v >= 0.0 && v < CUDART_INF) ? bits : 0x7ff8000000000000ULL;
}

__device__ __forceinline__ VxScale vx_scale(unsigned long long wbits, FixedScales fs) {
    const double wmax = __longlong_as_double((long long)wbits);
    VxScale s;
    s.tile = 0.0;
    s.glob = 0.0;
    if (wmax > 0.0 && wmax < CUDART_INF) {
        int e = ilogb(wmax) + 1;
        e = e < -900 ? -900 : e;
        s.tile = scalbn(1.0, fs.f_tile - e);
        s.glob = scalbn(1.0, fs.f_glob - e);
    }
    return s;
}

// low limbs of all cells first, then the high limbs: consecutive cells fall in consecutive banks for both
__device__ __forceinline__ void fixed_add_shared(unsigned* lo_limb, unsigned* hi_limb, long long q) {
    const unsigned lo = (unsigned)q, hi = (unsigned)((unsigned long long)q >> 32);
    const unsigned old = atomicAdd(lo_limb, lo);
    atomicAdd(hi_limb, hi + ((unsigned)(old + lo) < lo ? 1u : 0u));
}

// Same update with the conversion done by the adder: rn(c * scale) + 1.5 * 2^52 puts the integer (|q| <= 2^50)
// into the mantissa, so the limbs are the two words of the sum minus the constant's high word.  One DFMA
// and one IADD instead of DMUL + F2I.S64 + shifts.
constexpr double kMagic = 6755399441055744.0;          // 2^52 + 2^51 = 0x4338000000000000
__device__ __forceinline__ void fixed_add_shared_fma(unsigned* lo_limb, unsigned* hi_limb, double c, double scale) {
    const double m = __fma_rn(c, scale, kMagic);
    const unsigned lo = (unsigned)__double2loint(m), hi = (unsigned)__double2hiint(m) - 0x43380000u;
    const unsigned old = atomicAdd(lo_limb, lo);
    atomicAdd(hi_limb, hi + ((unsigned)(old + lo) < lo ? 1u : 0u));
}

template <bool kTile>
__device__ __forceinline__ void fixed_corner(const DepGrid& g, const Tile& t, unsigned* tile, int cells,
                                             unsigned long long* __restrict__ count, unsigned long long* __restrict__ vxsum,
                                             int i, int j, double c, double v, double sc_t, double sc_g,
                                             double sv_t, double sv_g) {
    if (i < 0 || i >= g.nx || j < 0 || j >= g.nz) return;     // the reference's per-corner guards
    if (kTile) {
        const int ti = i - t.i0, tj = j - t.j0;
        if (ti >= 0 && ti < t.ni && tj >= 0 && tj < t.nj) {
            const int o = ti * t.nj + tj;
            fixed_add_shared(tile + o, tile + 2 * cells + o, __double2ll_rn(c * sc_t));
            fixed_add_shared(tile + cells + o, tile + 3 * cells + o, __double2ll_rn(v * sv_t));
            return;
        }
    }
    const long long o = (long long)i * g.nz + j;
    atomicAdd(count + o, (unsigned long long)__double2ll_rn(c * sc_g));
    atomicAdd(vxsum + o, (unsigned long long)__double2ll_rn(v * sv_g));
}

template <bool kTile>
__global__ void __launch_bounds__(kTile ? kTileThreads : 256, kTile ? 1 : 4)
cic_fixed_kernel(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ px,
                 long long n, DepGrid g, Tile t, FixedScales fs, const unsigned long long* __restrict__ wslot,
                 unsigned long long wbits_value, unsigned long long* __restrict__ count,
                 unsigned long long* __restrict__ vxsum, const double* __restrict__ lim) {
    extern __shared__ unsigned ftile[];
    if (lim) {      // grid limits {x_lo, x_hi, z_lo, z_hi} from device memory (dfcsr_get_df_from_stats): make_grid on the device
        g.x_start = lim[0];
        g.z_start = lim[2];
        g.inv_dx = __ddiv_rn(1.0, __ddiv_rn(__dsub_rn(lim[1], lim[0]), (double)g.nx));
        g.inv_dz = __ddiv_rn(1.0, __ddiv_rn(__dsub_rn(lim[3], lim[2]), (double)g.nz));
    }
    const int cells = kTile ? t.ni * t.nj : 0;
    if (kTile) {
        for (int c = threadIdx.x; c < 4 * cells; c += blockDim.x) ftile[c] = 0u;
        __syncthreads();
    }
    const VxScale vs = vx_scale(wslot ? (lim ? finite_or_nan_bits(*wslot) : *wslot) : wbits_value, fs);
    const double sc_t = scalbn(1.0, fs.f_tile), sc_g = scalbn(1.0, fs.f_glob);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double xn = 0.0, zn = 0.0, wn = 0.0;
    if (p < n) { xn = x[p]; zn = z[p]; wn = px[p]; }
    while (p < n) {
        const double xc = xn, zc = zn, w = wn;
        p += stride;
        if (p < n) { xn = x[p]; zn = z[p]; wn = px[p]; }      // next particle in flight while this one deposits
        const CicSample s = cic_cell(g, xc, zc);
        const double a_hi = 1.0 - s.a_lo, b_hi = 1.0 - s.b_lo;
        const double c00 = s.a_lo * s.b_lo, c01 = s.a_lo * b_hi, c10 = a_hi * s.b_lo, c11 = a_hi * b_hi;
        if (kTile) {
            const int ti = s.i - t.i0, tj = s.j - t.j0;
            if (ti >= 0 && ti + 1 < t.ni && tj >= 0 && tj + 1 < t.nj) {     // all four corners inside the tile
                unsigned* cc = ftile + ti * t.nj + tj;       // count: low limbs, +2*cells high; velocity sum at +cells
                unsigned* vv = cc + cells;
                const int hi = 2 * cells;
                fixed_add_shared_fma(cc, cc + hi, c00, sc_t);
                fixed_add_shared_fma(cc + 1, cc + hi + 1, c01, sc_t);
                fixed_add_shared_fma(cc + t.nj, cc + hi + t.nj, c10, sc_t);
                fixed_add_shared_fma(cc + t.nj + 1, cc + hi + t.nj + 1, c11, sc_t);
                fixed_add_shared_fma(vv, vv + hi, w * c00, vs.tile);
                fixed_add_shared_fma(vv + 1, vv + hi + 1, w * c01, vs.tile);
                fixed_add_shared_fma(vv + t.nj, vv + hi + t.nj, w * c10, vs.tile);
                fixed_add_shared_fma(vv + t.nj + 1, vv + hi + t.nj + 1, w * c11, vs.tile);
                continue;
            }
        }
        fixed_corner<kTile>(g, t, ftile, cells, count, vxsum, s.i, s.j, c00, w * c00, sc_t, sc_g, vs.tile, vs.glob);
        fixed_corner<kTile>(g, t, ftile, cells, count, vxsum, s.i, s.j + 1, c01, w * c01, sc_t, sc_g, vs.tile, vs.glob);
        fixed_corner<kTile>(g, t, ftile, cells, count, vxsum, s.i + 1, s.j, c10, w * c10, sc_t, sc_g, vs.tile, vs.glob);
        fixed_corner<kTile>(g, t, ftile, cells, count, vxsum, s.i + 1, s.j + 1, c11, w * c11, sc_t, sc_g, vs.tile, vs.glob);
    }
    if (kTile) {
        __syncthreads();
        const int shift = fs.f_tile - fs.f_glob;
        const long long half = shift > 0 ? (1LL << (shift - 1)) : 0;
        // every CTA starts its sweep somewhere else, so that the CTAs do not queue on the same L2 lines
        const int start = (int)(((long long)blockIdx.x * 2 * cells) / gridDim.x);
        for (int k = threadIdx.x; k < 2 * cells; k += blockDim.x) {
            const int c = k + start < 2 * cells ? k + start : k + start - 2 * cells;
            const long long v64 = (long long)(((unsigned long long)ftile[2 * cells + c] << 32) | ftile[c]);
            const long long q = (v64 + half) >> shift;                     // round to the global scale
            if (q != 0) {
                const int cc = c < cells ? c : c - cells;
                const int ti = cc / t.nj, tj = cc - ti * t.nj;
                const long long o = (long long)(t.i0 + ti) * g.nz + (t.j0 + tj);
                atomicAdd((c < cells ? count : vxsum) + o, (unsigned long long)q);
            }
        }
    }
}

// int64 fixed point -> fp64, in place
__global__ void __launch_bounds__(256)
cic_fixed_finish(long long cells, FixedScales fs, const unsigned long long* __restrict__ wslot,
                 double* __restrict__ count, double* __restrict__ vxsum) {
    const unsigned long long wbits = *wslot;
    const VxScale vs = vx_scale(wbits, fs);
    const double inv_c = scalbn(1.0, -fs.f_glob);
    const double wmax = __longlong_as_double((long long)wbits);
    const bool finite = wmax < CUDART_INF;      // false for inf and NaN
    const double inv_v = vs.glob > 0.0 ? 1.0 / vs.glob : 0.0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        count[c] = (double)__double_as_longlong(count[c]) * inv_c;
        vxsum[c] = finite ? (double)__double_as_longlong(vxsum[c]) * inv_v : CUDART_NAN;
    }
}

// The all-reduce of a particle-sharded deposit, fused with the conversion: every rank reads the (2, cells) int64 grids of
// ALL ranks through NVLink peer mappings, adds them (integer adds: exact, any order) and converts to fp64.  Each rank
// ends up with identical bits, and with the bits a single rank holding all particles would have produced.
struct PeerQ {
    const long long* q[DFCSR_MAX_PEERS];
    int n;
};

__global__ void __launch_bounds__(256)
cic_fixed_finish_peers(long long cells, FixedScales fs, unsigned long long wbits, PeerQ peers,
                       double* __restrict__ count, double* __restrict__ vxsum, unsigned long long* __restrict__ count_max,
                       const unsigned long long* __restrict__ wslot) {
    if (wslot) wbits = finite_or_nan_bits(*wslot);      // max|px| from device memory (bit pattern of the statistic)
    const VxScale vs = vx_scale(wbits, fs);
    const double inv_c = scalbn(1.0, -fs.f_glob);
    const double wmax = __longlong_as_double((long long)wbits);
    const bool finite = wmax < CUDART_INF;      // false for inf and NaN
    const double inv_v = vs.glob > 0.0 ? 1.0 / vs.glob : 0.0;
    double cmax = 0.0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
        long long a = 0, b = 0;
#pragma unroll
        for (int p = 0; p < DFCSR_MAX_PEERS; ++p) {
            if (p < peers.n) {
                a += peers.q[p][c];
                b += peers.q[p][cells + c];
            }
        }
        const double cv = (double)a * inv_c;
        count[c] = cv;
        vxsum[c] = finite ? (double)b * inv_v : CUDART_NAN;
        cmax = fmax(cmax, cv);
    }
    if (count_max) {          // max(count) for dfcsr_make_df (deposit.py:183): non-negative doubles order like their bit patterns
        cmax = warp_max(cmax);
        if ((threadIdx.x & 31) == 0 && cmax > 0.0) atomicMax(count_max, (unsigned long long)__double_as_longlong(cmax));
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

static int bits_for(long long v) {      // smallest b with 2^b > v
    int b = 0;
    while ((1LL << b) <= v) ++b;
    return b;
}

static std::atomic<unsigned> g_next_slot{0};

// Scales of the fixed-point grids.  ONE scale for the block-private tiles and the output buffers: every particle's
// contribution is rounded once, to a grid of 2^-f that depends on the TOTAL particle count only, and everything after
// that is exact integer addition -- so the result does not depend on how the particles are cut into CTAs, launches or
// GPUs (a particle-sharded deposit summed over ranks gives the bits of the single-GPU deposit).
static FixedScales fixed_scales(long long n_total) {
    FixedScales fs;
    fs.f_glob = 62 - bits_for(n_total);
    fs.f_glob = fs.f_glob > 50 ? 50 : fs.f_glob;          // |q| <= 2^50 for the mantissa trick of fixed_add_shared_fma
    fs.f_tile = fs.f_glob;
    return fs;
}

// fixed-point deposit of n particles into the int64 grids c64 / v64 (not zeroed here)
static int deposit_fixed_accumulate(const double* d_x, const double* d_z, const double* d_px, long long n, long long n_total,
                                    const DepGrid& g, const unsigned long long* d_wslot, unsigned long long wbits,
                                    unsigned long long* c64, unsigned long long* v64, bool tiled, cudaStream_t st,
                                    const double* d_lim = nullptr) {
    const FixedScales fs = fixed_scales(n_total);
    if (tiled) {
        Tile t = make_tile(g.nx, g.nz);
        const size_t smem = (size_t)2 * t.ni * t.nj * sizeof(double);
        long long want = (n + kParticlesPerCta - 1) / kParticlesPerCta;
        unsigned blocks = (unsigned)(want < 148 ? (want < 1 ? 1 : want) : 148);
        DFCSR_CUDA_OK(cudaFuncSetAttribute(cic_fixed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cic_fixed_kernel<true><<<blocks, kTileThreads, smem, st>>>(d_x, d_z, d_px, n, g, t, fs, d_wslot, wbits, c64, v64, d_lim);
    } else {
        Tile t = {0, 0, 0, 0};
        long long want = (n + 255) / 256;
        unsigned blocks = (unsigned)(want < 148LL * 4 ? want : 148LL * 4);
        cic_fixed_kernel<false><<<blocks, 256, 0, st>>>(d_x, d_z, d_px, n, g, t, fs, d_wslot, wbits, c64, v64, d_lim);
    }
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

// absmax(px) -> fixed-point deposit -> in-place conversion to fp64 (3 launches)
static int deposit_fixed(const double* d_x, const double* d_z, const double* d_px, long long n, const DepGrid& g,
                         double* d_count, double* d_vxsum, bool tiled, cudaStream_t st) {
    unsigned long long* slots = nullptr;
    DFCSR_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&slots), g_absmax));
    unsigned long long* slot = slots + (g_next_slot.fetch_add(1) % kScratchSlots);
    DFCSR_CUDA_OK(cudaMemsetAsync(slot, 0, sizeof(unsigned long long), st));
    {
        long long want = (n + 2047) / 2048;
        unsigned blocks = (unsigned)(want < 148LL * 8 ? (want < 1 ? 1 : want) : 148LL * 8);
        absmax_kernel<<<blocks, 256, 0, st>>>(d_px, n, slot);
        count_launch(1);
    }
    int rc = deposit_fixed_accumulate(d_x, d_z, d_px, n, n, g, slot, 0ull, reinterpret_cast<unsigned long long*>(d_count),
                                      reinterpret_cast<unsigned long long*>(d_vxsum), tiled, st);
    if (rc) return rc;
    {
        const long long cells = (long long)g.nx * g.nz;
        long long want = (cells + 255) / 256;
        unsigned blocks = (unsigned)(want < 148LL * 4 ? want : 148LL * 4);
        cic_fixed_finish<<<blocks, 256, 0, st>>>(cells, fixed_scales(n), slot, d_count, d_vxsum);
        count_launch(1);
    }
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

// bit pattern of max|px| as the kernels expect it (NaN / inf / negative input: "not finite" => vxsum becomes NaN)
static unsigned long long absmax_bits(double absmax_px) {
    if (!(absmax_px >= 0.0) || std::isinf(absmax_px)) return 0x7ff8000000000000ULL;
    unsigned long long b;
    memcpy(&b, &absmax_px, sizeof(b));
    return b;
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_deposit_cic(const double* d_x, const double* d_z, const double* d_px, int64_t n,
                                 int32_t nx, double x_start, double x_end, int32_t nz, double z_start,
                                 double z_end, double* d_count, double* d_vxsum, int32_t mode, void* stream) {
    DFCSR_REQUIRE(d_count && d_vxsum && (n == 0 || (d_x && d_z && d_px)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n >= 0, "bad sizes");
    DFCSR_REQUIRE((long long)nx * nz < (1LL << 30), "grid too large");
    DFCSR_REQUIRE(mode >= 0 && mode <= 5,
                  "mode must be 0 (auto), 1 (tile + warp match), 2 (direct), 3 (tile), 4 (fixed-point tile) or 5 (fixed-point direct)");
    cudaStream_t st = as_stream(stream);
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_count, 0, cells * sizeof(double), st));
    DFCSR_CUDA_OK(cudaMemsetAsync(d_vxsum, 0, cells * sizeof(double), st));
    if (n == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, x_start, x_end, nz, z_start, z_end);
    if (mode == 0) mode = (n >= 65536) ? 4 : 5;
    if (mode >= 4) return deposit_fixed(d_x, d_z, d_px, n, g, d_count, d_vxsum, mode == 4, st);   // counts its own launches
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

extern "C" int dfcsr_deposit_cic_q(const double* d_x, const double* d_z, const double* d_px, int64_t n_local,
                                   int64_t n_total, int32_t nx, double x_start, double x_end, int32_t nz, double z_start,
                                   double z_end, double absmax_px, int64_t* d_q, void* stream) {
    DFCSR_REQUIRE(d_q && (n_local == 0 || (d_x && d_z && d_px)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n_local >= 0 && n_total >= n_local && n_total >= 1, "bad sizes");
    DFCSR_REQUIRE((long long)nx * nz < (1LL << 30), "grid too large");
    cudaStream_t st = as_stream(stream);
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_q, 0, 2 * cells * sizeof(int64_t), st));
    if (n_local == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, x_start, x_end, nz, z_start, z_end);
    unsigned long long* q = reinterpret_cast<unsigned long long*>(d_q);
    return deposit_fixed_accumulate(d_x, d_z, d_px, n_local, n_total, g, nullptr, absmax_bits(absmax_px), q, q + cells,
                                    n_local >= 65536, st);
}

extern "C" int dfcsr_deposit_cic_finish(const uint64_t* h_peer_q, int32_t n_peers, int32_t nx, int32_t nz, int64_t n_total,
                                        double absmax_px, double* d_count, double* d_vxsum, uint64_t* d_count_max,
                                        void* stream) {
    DFCSR_REQUIRE(h_peer_q && n_peers >= 1 && n_peers <= DFCSR_MAX_PEERS && d_count && d_vxsum, "bad argument");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n_total >= 1, "bad sizes");
    PeerQ pq;
    pq.n = n_peers;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) {
        pq.q[p] = p < n_peers ? reinterpret_cast<const long long*>(static_cast<uintptr_t>(h_peer_q[p])) : nullptr;
        DFCSR_REQUIRE(p >= n_peers || pq.q[p] != nullptr, "null peer grid");
    }
    const long long cells = (long long)nx * nz;
    long long want = (cells + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 4 ? want : 148LL * 4);
    if (d_count_max) DFCSR_CUDA_OK(cudaMemsetAsync(d_count_max, 0, sizeof(uint64_t), as_stream(stream)));
    cic_fixed_finish_peers<<<blocks, 256, 0, as_stream(stream)>>>(cells, fixed_scales(n_total), absmax_bits(absmax_px), pq,
                                                                  d_count, d_vxsum,
                                                                  reinterpret_cast<unsigned long long*>(d_count_max), nullptr);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

// Library-internal (common.cuh): the two stages of the fixed-point deposit of ONE GPU with the grid limits and max|px| read
// from device memory -- d_lim = {x_lo, x_hi, z_lo, z_hi}, d_wslot = the max|px| statistic -- so that they can be enqueued
// before the host has seen the statistics (dfcsr_get_df_from_stats).
namespace dfcsr {
int deposit_one_gpu_from_device(const double* d_x, const double* d_z, const double* d_px, long long n, int nx, int nz,
                                const double* d_lim, const unsigned long long* d_wslot, long long* d_q, double* d_count,
                                double* d_vxsum, unsigned long long* d_count_max, cudaStream_t st) {
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_q, 0, 2 * cells * sizeof(long long), st));
    DepGrid g = make_grid(nx, 0.0, 1.0, nz, 0.0, 1.0);          // dimensions only: the kernel takes the limits from d_lim
    unsigned long long* q = reinterpret_cast<unsigned long long*>(d_q);
    if (n > 0) {
        int rc = deposit_fixed_accumulate(d_x, d_z, d_px, n, n, g, d_wslot, 0ull, q, q + cells, n >= 65536, st, d_lim);
        if (rc) return rc;
    }
    PeerQ pq;
    pq.n = 1;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) pq.q[p] = p == 0 ? d_q : nullptr;
    long long want = ((long long)cells + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 4 ? want : 148LL * 4);
    DFCSR_CUDA_OK(cudaMemsetAsync(d_count_max, 0, sizeof(unsigned long long), st));
    cic_fixed_finish_peers<<<blocks, 256, 0, st>>>((long long)cells, fixed_scales(n), 0ull, pq, d_count, d_vxsum, d_count_max,
                                                   d_wslot);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
}  // namespace dfcsr

// The two stages of a particle-sharded deposit with the grid limits (d_limits = {x_lo, x_hi, z_lo, z_hi}) and max|px|
// (the statistic in d_stats) read from device memory: for DF_tracker.prefetch_DF with particle shards.
extern "C" int dfcsr_deposit_cic_q_dev(const double* d_x, const double* d_z, const double* d_px, int64_t n_local,
                                       int64_t n_total, int32_t nx, int32_t nz, const double* d_limits,
                                       const double* d_stats, int64_t* d_q, void* stream) {
    DFCSR_REQUIRE(d_q && d_limits && d_stats && (n_local == 0 || (d_x && d_z && d_px)), "null pointer");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n_local >= 0 && n_total >= n_local && n_total >= 1, "bad sizes");
    DFCSR_REQUIRE((long long)nx * nz < (1LL << 30), "grid too large");
    cudaStream_t st = as_stream(stream);
    const size_t cells = (size_t)nx * nz;
    DFCSR_CUDA_OK(cudaMemsetAsync(d_q, 0, 2 * cells * sizeof(int64_t), st));
    if (n_local == 0) return DFCSR_OK;
    DepGrid g = make_grid(nx, 0.0, 1.0, nz, 0.0, 1.0);          // dimensions only
    unsigned long long* q = reinterpret_cast<unsigned long long*>(d_q);
    return deposit_fixed_accumulate(d_x, d_z, d_px, n_local, n_total, g,
                                    reinterpret_cast<const unsigned long long*>(d_stats + DFCSR_S_ABSMAX_PX), 0ull, q, q + cells,
                                    n_local >= 65536, st, d_limits);
}

extern "C" int dfcsr_deposit_cic_finish_dev(const uint64_t* h_peer_q, int32_t n_peers, int32_t nx, int32_t nz,
                                            int64_t n_total, const double* d_stats, double* d_count, double* d_vxsum,
                                            uint64_t* d_count_max, void* stream) {
    DFCSR_REQUIRE(h_peer_q && n_peers >= 1 && n_peers <= DFCSR_MAX_PEERS && d_count && d_vxsum && d_stats, "bad argument");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && n_total >= 1, "bad sizes");
    PeerQ pq;
    pq.n = n_peers;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) {
        pq.q[p] = p < n_peers ? reinterpret_cast<const long long*>(static_cast<uintptr_t>(h_peer_q[p])) : nullptr;
        DFCSR_REQUIRE(p >= n_peers || pq.q[p] != nullptr, "null peer grid");
    }
    const long long cells = (long long)nx * nz;
    long long want = (cells + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 4 ? want : 148LL * 4);
    if (d_count_max) DFCSR_CUDA_OK(cudaMemsetAsync(d_count_max, 0, sizeof(uint64_t), as_stream(stream)));
    cic_fixed_finish_peers<<<blocks, 256, 0, as_stream(stream)>>>(
        cells, fixed_scales(n_total), 0ull, pq, d_count, d_vxsum, reinterpret_cast<unsigned long long*>(d_count_max),
        reinterpret_cast<const unsigned long long*>(d_stats + DFCSR_S_ABSMAX_PX));
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
