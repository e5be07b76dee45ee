// beam.cu — A14 beam scalars and K5 kick application.
//
// dfcsr_beam_stats replaces the O(Np) numpy reductions the reference performs on the host every
// step (two passes: moments about the first particle, then the statistics that need them): np.std / np.mean / np.polyfit(z, x, 1) in Beam.update_status (beams.py:88-98,137-156,
// 201-215), the slice test of DF_tracker.get_DF (deposit.py:147-159) and the statistics of
// x_transform used by get_CSR_mesh (CSR.py:368-374).  Three two-level deterministic reductions
// (fixed grid, fixed summation order -> bitwise reproducible); only 16 doubles go back to the host.
//
// dfcsr_apply_kick replaces Beam.apply_wakes (beams.py:108-131): bilinear samples
// (RegularGridInterpolator, fill 0) of the two wake grids at (x - polyval(slope, z), z).
// Bound: HBM, 48 B / particle (read x, z, px, pz; write px, pz).
#include <math_constants.h>
#include "common.cuh"

namespace dfcsr {

// ---- reductions that do not depend on how the particles are distributed over GPUs ------------------------------
// The particle index space [0, n_total) is cut into DFCSR_STAT_BLOCKS contiguous chunks of C = ceil(n_total / 1024)
// particles.  One CTA reduces one chunk in a fixed order (thread t takes particles t, t + 256, ... of the chunk, then
// a fixed warp butterfly, then the warps in order) and publishes one row of a (1024, NV) table; the table is summed
// in a fixed order by one CTA.  Which GPU reduced which chunk is irrelevant: a run on N GPUs, each holding whole
// chunks, produces the same bits as a run on one.  With several ranks the rows are stored straight into the tables of
// all ranks (NVLink peer memory) and a barrier + dfcsr_*_final follow; a single rank lets the last CTA do the final sum.
constexpr int kStatThreads = 256;
constexpr int kStatBlocks = DFCSR_STAT_BLOCKS;
constexpr int kStatVals = 8;
constexpr int kCovVals = 27;

struct StatWorkspace {
    double partial[2][kStatBlocks][kStatVals];
    unsigned int ticket[4];   // must be zero before the first call; every pass leaves its ticket at zero
};

struct CovWorkspace {
    double partial[kStatBlocks][kCovVals];
    unsigned int ticket[4];
};

struct PeerTables {
    double* tab[DFCSR_MAX_PEERS];
    int n;
};

__device__ __forceinline__ double absmax_merge(double m, double a) {      // NaN is sticky (fmax would drop it)
    return (a != a || m != m) ? CUDART_NAN : fmax(m, a);
}

// block totals of v[0..NV) in thread 0; entry kMaxIdx (if >= 0) is a maximum, all others are sums
template <int NV, int kMaxIdx>
__device__ __forceinline__ void block_total(double (&v)[NV]) {
    __shared__ double sm[kStatThreads / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (k == kMaxIdx) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] = absmax_merge(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
        } else {
            v[k] = warp_sum(v[k]);
        }
    }
    __syncthreads();                       // sm may still be read by a previous call
    if (lane == 0)
        for (int k = 0; k < NV; ++k) sm[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < NV; ++k) {
            double s = sm[0][k];
            for (int w = 1; w < kStatThreads / 32; ++w) s = (k == kMaxIdx) ? absmax_merge(s, sm[w][k]) : s + sm[w][k];
            v[k] = s;
        }
    }
}

// fixed-order total of a (kStatBlocks, NV) table; result in thread 0 (called by all 256 threads of ONE block)
template <int NV, int kMaxIdx>
__device__ __forceinline__ void table_total(const double* table, double (&tot)[NV]) {
    // all loads first (L2-coherent ld.cg: the rows were written by other CTAs, possibly of other GPUs), then the adds in a
    // fixed order -- a load-add-load-add chain would serialise 4 NV L2 round trips per thread
    constexpr int kPer = kStatBlocks / kStatThreads;
    static_assert(kPer * kStatThreads == kStatBlocks, "table rows must divide evenly among the threads");
    double p[kPer][NV];
#pragma unroll
    for (int i = 0; i < kPer; ++i)
#pragma unroll
        for (int k = 0; k < NV; ++k) p[i][k] = __ldcg(table + (size_t)(threadIdx.x + i * kStatThreads) * NV + k);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = p[0][k];
#pragma unroll
        for (int i = 1; i < kPer; ++i) s = (k == kMaxIdx) ? absmax_merge(s, p[i][k]) : s + p[i][k];
        tot[k] = s;
    }
    block_total<NV, kMaxIdx>(tot);
}

// publish one table row: local table and, with several ranks, every peer's table
template <int NV>
__device__ __forceinline__ void publish_row(const double (&v)[NV], int row, double* table, const PeerTables& peers) {
    for (int k = 0; k < NV; ++k) table[(size_t)row * NV + k] = v[k];
#pragma unroll
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p)
        if (p < peers.n && peers.tab[p] != table)
            for (int k = 0; k < NV; ++k) peers.tab[p][(size_t)row * NV + k] = v[k];
}

// single-rank mode: the last CTA to publish its row does the final sum.  true in thread 0 of that CTA, with tot filled.
template <int NV, int kMaxIdx>
__device__ __forceinline__ bool last_block_total(unsigned int* ticket, const double* table, double (&tot)[NV]) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    table_total<NV, kMaxIdx>(table, tot);
    if (threadIdx.x == 0) *ticket = 0u;
    return threadIdx.x == 0;
}

struct Centre3 { double x, z, p; };

__device__ __forceinline__ void finish_moments(const double (&tot)[kStatVals], long long n, Centre3 c, bool have_pz,
                                               bool have_px, double* __restrict__ stats) {
    const double inv_n = 1.0 / (double)n;
    const double ex = tot[0] * inv_n, ez = tot[1] * inv_n, ep = tot[5] * inv_n;
    const double vxx = fmax(tot[2] * inv_n - ex * ex, 0.0);
    const double vzz = fmax(tot[3] * inv_n - ez * ez, 0.0);
    const double cxz = tot[4] * inv_n - ex * ez;
    const double mx = c.x + ex, mz = c.z + ez;
    stats[DFCSR_S_MEAN_X] = mx;
    stats[DFCSR_S_MEAN_Z] = mz;
    stats[DFCSR_S_SIGMA_X] = sqrt(vxx);     // np.std: population (ddof = 0)
    stats[DFCSR_S_SIGMA_Z] = sqrt(vzz);
    const double slope = cxz / vzz;           // least-squares line x = slope z + b (np.polyfit(z, x, 1))
    stats[DFCSR_S_SLOPE] = slope;
    stats[DFCSR_S_INTERCEPT] = mx - slope * mz;
    stats[DFCSR_S_MEAN_PZ] = have_pz ? c.p + ep : 0.0;
    stats[DFCSR_S_SIGMA_PZ] = have_pz ? sqrt(fmax(tot[6] * inv_n - ep * ep, 0.0)) : 0.0;
    stats[DFCSR_S_N] = (double)n;
    stats[DFCSR_S_ABSMAX_PX] = have_px ? tot[7] : -1.0;
    stats[14] = 0.0;
    stats[15] = 0.0;
}

__device__ __forceinline__ void finish_residuals(const double (&tot)[kStatVals], long long n, double* __restrict__ stats) {
    const double mx = stats[DFCSR_S_MEAN_X], mz = stats[DFCSR_S_MEAN_Z];
    double me = tot[0] / (double)n;
    stats[DFCSR_S_MEAN_XT] = me;
    stats[DFCSR_S_SIGMA_XT] = sqrt(fmax(tot[1] / (double)n - me * me, 0.0));
    double c = tot[4];
    double ms = tot[2] / c;
    stats[DFCSR_S_SLICE_SIGMA_X] = sqrt(fmax(tot[3] / c - ms * ms, 0.0));
    stats[DFCSR_S_SLICE_COUNT] = c;
    stats[DFCSR_S_SIGMA_X] = sqrt(tot[5] / (double)n);
    stats[DFCSR_S_SIGMA_Z] = sqrt(tot[6] / (double)n);
    const double slope2 = tot[7] / tot[6];
    stats[DFCSR_S_SLOPE] = slope2;
    stats[DFCSR_S_INTERCEPT] = mx - slope2 * mz;
}

// Pass A: first and second moments in ONE read of (x, z[, pz][, px]).  Sums are taken about a centre the caller supplies
// (the previous step's means: E[d^2] - E[d]^2 then loses at most ~1 digit even for a bunch far from the origin; pass B
// recomputes the second moments about the true means anyway).  Entry 7 is max|px| for the fixed-point deposit.
// Pointers are those of this rank's shard; CTA b reduces chunk first_block + b, i.e. local particles [b C, (b+1) C).
__global__ void __launch_bounds__(kStatThreads)
stats_moments(const double* __restrict__ x, const double* __restrict__ z, const double* __restrict__ pz,
              const double* __restrict__ px, long long n_local, long long chunk, long long n_total, int first_block,
              Centre3 c, double* __restrict__ table, PeerTables peers, unsigned int* ticket, double* __restrict__ stats) {
    double v[kStatVals];
#pragma unroll
    for (int k = 0; k < kStatVals; ++k) v[k] = 0.0;
    const long long lo = (long long)blockIdx.x * chunk;
    const long long hi = (lo + chunk < n_local) ? lo + chunk : n_local;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += kStatThreads) {
        double dx = x[i] - c.x, dz = z[i] - c.z;
        v[0] += dx;
        v[1] += dz;
        v[2] = fma(dx, dx, v[2]);
        v[3] = fma(dz, dz, v[3]);
        v[4] = fma(dx, dz, v[4]);
        if (pz) {
            double dp = pz[i] - c.p;
            v[5] += dp;
            v[6] = fma(dp, dp, v[6]);
        }
        if (px) v[7] = absmax_merge(v[7], fabs(px[i]));
    }
    block_total<kStatVals, 7>(v);
    if (threadIdx.x == 0) publish_row<kStatVals>(v, first_block + blockIdx.x, table, peers);
    if (ticket) {
        double tot[kStatVals];
        if (last_block_total<kStatVals, 7>(ticket, table, tot)) finish_moments(tot, n_total, c, pz != nullptr, px != nullptr, stats);
    }
}

// Pass B: statistics that need pass A's results: x_transform = x - polyval(slope, z) (CSR.py:368-372) and the
// central slice |z| < 0.1 sigma_z of DF_tracker.get_DF (deposit.py:157-159).
__global__ void __launch_bounds__(kStatThreads)
stats_residuals(const double* __restrict__ x, const double* __restrict__ z, long long n_local, long long chunk,
                long long n_total, int first_block, double* __restrict__ table, PeerTables peers, unsigned int* ticket,
                double* __restrict__ stats) {
    const double mx = stats[DFCSR_S_MEAN_X], mz = stats[DFCSR_S_MEAN_Z];
    const double slope = stats[DFCSR_S_SLOPE];
    const double cut = 0.1 * stats[DFCSR_S_SIGMA_Z];
    double v[kStatVals];
#pragma unroll
    for (int k = 0; k < kStatVals; ++k) v[k] = 0.0;
    const long long lo = (long long)blockIdx.x * chunk;
    const long long hi = (lo + chunk < n_local) ? lo + chunk : n_local;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += kStatThreads) {
        double zi = z[i];
        double dx = x[i] - mx, dz = zi - mz;
        double e = dx - slope * dz;          // x_transform up to the (rounding-level) constant term
        v[0] += e;
        v[1] = fma(e, e, v[1]);
        v[5] = fma(dx, dx, v[5]);            // second moments about the true means: two-pass accuracy, like
        v[6] = fma(dz, dz, v[6]);            // np.std / np.polyfit (pass A's one-pass values lose ~2 digits,
        v[7] = fma(dx, dz, v[7]);            // enough to shift the deposit grid by 1e-13 of a cell)
        if (fabs(zi) < cut) {                // deposit.py:157: |z|, not |z - mean z|
            v[2] += dx;
            v[3] = fma(dx, dx, v[3]);
            v[4] += 1.0;
        }
    }
    block_total<kStatVals, -1>(v);
    if (threadIdx.x == 0) publish_row<kStatVals>(v, first_block + blockIdx.x, table, peers);
    if (ticket) {
        double tot[kStatVals];
        if (last_block_total<kStatVals, -1>(ticket, table, tot)) finish_residuals(tot, n_total, stats);
    }
}

// final sums after the cross-rank exchange (one CTA); pass 0 = moments, 1 = residuals
__global__ void __launch_bounds__(kStatThreads)
stats_final_kernel(int pass, const double* __restrict__ table, long long n_total, Centre3 c, int have_pz, int have_px,
                   double* __restrict__ stats) {
    double tot[kStatVals];
    if (pass == 0) {
        table_total<kStatVals, 7>(table, tot);
        if (threadIdx.x == 0) finish_moments(tot, n_total, c, have_pz != 0, have_px != 0, stats);
    } else {
        table_total<kStatVals, -1>(table, tot);
        if (threadIdx.x == 0) finish_residuals(tot, n_total, stats);
    }
}

// ---- 6 x 6 covariance of the phase-space coordinates (Twiss / dispersion statistics, twiss.py:2-71) ------
// One read of the six coordinate arrays: sums and upper-triangle products about a caller-supplied centre, reduced like
// the passes above.  out[0..5] = means, out[6..26] = covariance (i <= j, row-major over the upper triangle) with
// np.cov's normalisation 1/(n-1).
struct SixPtr {
    const double* q[6];
};

struct Centre6 { double o[6]; };

__device__ __forceinline__ void finish_cov(const double (&tot)[kCovVals], long long n, const Centre6& c, double* __restrict__ out) {
    const double inv_n = 1.0 / (double)n, inv_n1 = 1.0 / (double)(n - 1);
    for (int a = 0; a < 6; ++a) out[a] = c.o[a] + tot[a] * inv_n;
    int k = 6;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b, ++k) out[k] = (tot[k] - tot[a] * tot[b] * inv_n) * inv_n1;
}

__global__ void __launch_bounds__(kStatThreads)
cov6_kernel(SixPtr c, long long n_local, long long chunk, long long n_total, int first_block, Centre6 ctr,
            double* __restrict__ table, PeerTables peers, unsigned int* ticket, double* __restrict__ out) {
    double v[kCovVals];
#pragma unroll
    for (int k = 0; k < kCovVals; ++k) v[k] = 0.0;
    const long long lo = (long long)blockIdx.x * chunk;
    const long long hi = (lo + chunk < n_local) ? lo + chunk : n_local;
    for (long long i = lo + threadIdx.x; i < hi; i += kStatThreads) {
        double d[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) d[a] = c.q[a][i] - ctr.o[a];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            v[a] += d[a];
#pragma unroll
            for (int b = a; b < 6; ++b) {
                const int k = 6 + a * 6 - (a * (a - 1)) / 2 + (b - a);      // upper triangle, row-major
                v[k] = fma(d[a], d[b], v[k]);
            }
        }
    }
    block_total<kCovVals, -1>(v);
    if (threadIdx.x == 0) publish_row<kCovVals>(v, first_block + blockIdx.x, table, peers);
    if (ticket) {
        double tot[kCovVals];
        if (last_block_total<kCovVals, -1>(ticket, table, tot)) finish_cov(tot, n_total, ctr, out);
    }
}

__global__ void __launch_bounds__(kStatThreads)
cov_final_kernel(const double* __restrict__ table, long long n_total, Centre6 ctr, double* __restrict__ out) {
    double tot[kCovVals];
    table_total<kCovVals, -1>(table, tot);
    if (threadIdx.x == 0) finish_cov(tot, n_total, ctr, out);
}

// ---- K5 ---------------------------------------------------------------------------------------
struct Cell1 {
    int i;
    double y;
    bool outside;
};

// RegularGridInterpolator's per-axis search on linspace nodes (see history.cu::locate).  This kernel is
// bandwidth-bound at 48 B / particle, so the two fp64 divisions of the textbook form are replaced by
// multiplications with 1/step (differences at the 1e-16 level; the kick gate is 1e-10).
__device__ __forceinline__ Cell1 locate1(const Axis& g, double inv_step, double q) {
    Cell1 c;
    c.outside = (q < g.start) || (q > g.stop) || !(q == q);
    int i = 0;
    if (!c.outside) {
        double guess = floor((q - g.start) * inv_step);
        i = (guess < 0.0) ? 0 : ((guess > (double)(g.n - 2)) ? g.n - 2 : (int)guess);
        if (i > 0 && axis_node(g, i) > q) --i;                  // the guess is off by at most one node
        else if (i < g.n - 2 && axis_node(g, i + 1) <= q) ++i;
    }
    c.i = i;
    c.y = (q - axis_node(g, i)) * inv_step;
    return c;
}

__global__ void __launch_bounds__(256)
apply_kick_kernel(const double* __restrict__ x, const double* __restrict__ z, double* __restrict__ px,
                  double* __restrict__ pz, long long n, double slope, double intercept,
                  const double* __restrict__ dE, const double* __restrict__ kick, Axis ax, Axis az,
                  double factor, int transverse_on, double inv_sx, double inv_sz) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const double zp = z[p];
        const double xt = __dsub_rn(x[p], __dadd_rn(__dmul_rn(slope, zp), intercept));   // x - polyval(slope, z)
        Cell1 cx = locate1(ax, inv_sx, xt), cz = locate1(az, inv_sz, zp);
        if (cx.outside || cz.outside) continue;   // fill_value = 0: nothing to add
        const double wx0 = 1.0 - cx.y, wz0 = 1.0 - cz.y;
        const double w00 = wx0 * wz0, w01 = wx0 * cz.y, w10 = cx.y * wz0, w11 = cx.y * cz.y;
        const size_t o = (size_t)cx.i * az.n + cz.i;
        double e = __ldg(dE + o) * w00 + __ldg(dE + o + 1) * w01 + __ldg(dE + o + az.n) * w10 + __ldg(dE + o + az.n + 1) * w11;
        pz[p] += factor * e;
        if (transverse_on) {
            double k = __ldg(kick + o) * w00 + __ldg(kick + o + 1) * w01 + __ldg(kick + o + az.n) * w10 + __ldg(kick + o + az.n + 1) * w11;
            px[p] += factor * k;
        }
    }
}

// ---- linear transfer map (first-order stand-in for track_element, beams.py:101-102) ----
// v <- M v for every particle, in place: 96 B / particle, one pass instead of a dozen elementwise launches.
struct Map6 {
    double m[36];
};

__global__ void __launch_bounds__(256)
track_linear_kernel(double* __restrict__ x, double* __restrict__ px, double* __restrict__ y, double* __restrict__ py,
                    double* __restrict__ z, double* __restrict__ pz, long long n, Map6 M) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const double v[6] = {x[p], px[p], y[p], py[p], z[p], pz[p]};
        double o[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) a = fma(M.m[6 * r + c], v[c], a);
            o[r] = a;
        }
        x[p] = o[0]; px[p] = o[1]; y[p] = o[2]; py[p] = o[3]; z[p] = o[4]; pz[p] = o[5];
    }
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_track_linear(double* d_x, double* d_px, double* d_y, double* d_py, double* d_z, double* d_pz,
                                  int64_t n, const double* h_matrix, void* stream) {
    DFCSR_REQUIRE(h_matrix && n >= 0, "null matrix or negative count");
    if (n == 0) return DFCSR_OK;
    DFCSR_REQUIRE(d_x && d_px && d_y && d_py && d_z && d_pz, "null pointer");
    Map6 M;
    for (int k = 0; k < 36; ++k) M.m[k] = h_matrix[k];
    long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 64 ? want : 148LL * 64);
    track_linear_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_x, d_px, d_y, d_py, d_z, d_pz, n, M);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int64_t dfcsr_beam_stats_workspace(void) { return (int64_t)sizeof(StatWorkspace); }

extern "C" int64_t dfcsr_beam_cov_workspace(void) { return (int64_t)sizeof(CovWorkspace); }

extern "C" int64_t dfcsr_stat_chunk(int64_t n_total) {
    return n_total <= 0 ? 1 : (n_total + kStatBlocks - 1) / kStatBlocks;
}

static int peer_tables(const uint64_t* h_peer_tables, int32_t n_peers, PeerTables& pt) {
    DFCSR_REQUIRE(n_peers >= 0 && n_peers <= DFCSR_MAX_PEERS && (n_peers == 0 || h_peer_tables), "bad peer list");
    pt.n = n_peers;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p)
        pt.tab[p] = p < n_peers ? reinterpret_cast<double*>(static_cast<uintptr_t>(h_peer_tables[p])) : nullptr;
    for (int p = 0; p < n_peers; ++p) DFCSR_REQUIRE(pt.tab[p] != nullptr, "null peer table");
    return DFCSR_OK;
}

// a shard must consist of whole chunks (the last chunk of the bunch may be short)
static int check_shard(int64_t n_local, int64_t n_total, int32_t first_block, int32_t n_blocks) {
    const int64_t chunk = dfcsr_stat_chunk(n_total);
    DFCSR_REQUIRE(n_total >= 2 && n_local >= 0, "need at least two particles");
    DFCSR_REQUIRE(first_block >= 0 && n_blocks >= 0 && first_block + n_blocks <= kStatBlocks, "bad block range");
    const int64_t lo = (int64_t)first_block * chunk, hi = (int64_t)(first_block + n_blocks) * chunk;
    const int64_t expect = (hi < n_total ? hi : n_total) - (lo < n_total ? lo : n_total);
    DFCSR_REQUIRE(n_local == expect, "shard size does not match its block range (shards are whole chunks of dfcsr_stat_chunk(n_total) particles)");
    return DFCSR_OK;
}

static Centre3 centre3(const double* h) {
    Centre3 c;
    c.x = h ? h[0] : 0.0; c.z = h ? h[1] : 0.0; c.p = h ? h[2] : 0.0;
    return c;
}

extern "C" int dfcsr_beam_stats(const double* d_x, const double* d_z, const double* d_pz, const double* d_px, int64_t n,
                                const double* h_centre, double* d_stats, void* d_workspace, void* stream) {
    DFCSR_REQUIRE(d_x && d_z && d_stats && d_workspace, "null pointer");
    DFCSR_REQUIRE(n >= 2, "need at least two particles");
    cudaStream_t st = as_stream(stream);
    StatWorkspace* ws = reinterpret_cast<StatWorkspace*>(d_workspace);
    const long long chunk = dfcsr_stat_chunk(n);
    PeerTables none;
    none.n = 0;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) none.tab[p] = nullptr;
    stats_moments<<<kStatBlocks, kStatThreads, 0, st>>>(d_x, d_z, d_pz, d_px, n, chunk, n, 0, centre3(h_centre),
                                                         &ws->partial[0][0][0], none, &ws->ticket[0], d_stats);
    stats_residuals<<<kStatBlocks, kStatThreads, 0, st>>>(d_x, d_z, n, chunk, n, 0, &ws->partial[1][0][0], none,
                                                           &ws->ticket[1], d_stats);
    count_launch(2);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_beam_stats_partial(int32_t pass, const double* d_x, const double* d_z, const double* d_pz,
                                        const double* d_px, int64_t n_local, int64_t n_total, int32_t first_block,
                                        int32_t n_blocks, const double* h_centre, double* d_stats, double* d_table,
                                        const uint64_t* h_peer_tables, int32_t n_peers, void* stream) {
    DFCSR_REQUIRE((pass == 0 || pass == 1) && d_table && d_stats, "bad argument");
    DFCSR_REQUIRE(n_local == 0 || (d_x && d_z), "null pointer");
    int rc = check_shard(n_local, n_total, first_block, n_blocks);
    if (rc) return rc;
    PeerTables pt;
    rc = peer_tables(h_peer_tables, n_peers, pt);
    if (rc) return rc;
    if (n_blocks == 0) return DFCSR_OK;
    const long long chunk = dfcsr_stat_chunk(n_total);
    cudaStream_t st = as_stream(stream);
    if (pass == 0)
        stats_moments<<<(unsigned)n_blocks, kStatThreads, 0, st>>>(d_x, d_z, d_pz, d_px, n_local, chunk, n_total, first_block,
                                                                   centre3(h_centre), d_table, pt, nullptr, d_stats);
    else
        stats_residuals<<<(unsigned)n_blocks, kStatThreads, 0, st>>>(d_x, d_z, n_local, chunk, n_total, first_block, d_table, pt,
                                                                     nullptr, d_stats);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_beam_stats_final(int32_t pass, const double* d_table, int64_t n_total, const double* h_centre,
                                      int32_t have_pz, int32_t have_px, double* d_stats, void* stream) {
    DFCSR_REQUIRE((pass == 0 || pass == 1) && d_table && d_stats && n_total >= 2, "bad argument");
    stats_final_kernel<<<1, kStatThreads, 0, as_stream(stream)>>>(pass, d_table, n_total, centre3(h_centre), have_pz, have_px, d_stats);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

static Centre6 centre6(const double* h) {
    Centre6 c;
    for (int a = 0; a < 6; ++a) c.o[a] = h ? h[a] : 0.0;
    return c;
}

extern "C" int dfcsr_beam_cov(const double* d_x, const double* d_px, const double* d_y, const double* d_py,
                              const double* d_z, const double* d_pz, int64_t n, const double* h_centre, double* d_out,
                              void* d_workspace, void* stream) {
    DFCSR_REQUIRE(d_x && d_px && d_y && d_py && d_z && d_pz && d_out && d_workspace, "null pointer");
    DFCSR_REQUIRE(n >= 2, "need at least two particles");
    SixPtr c;
    c.q[0] = d_x; c.q[1] = d_px; c.q[2] = d_y; c.q[3] = d_py; c.q[4] = d_z; c.q[5] = d_pz;
    CovWorkspace* ws = static_cast<CovWorkspace*>(d_workspace);
    PeerTables none;
    none.n = 0;
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) none.tab[p] = nullptr;
    cov6_kernel<<<kStatBlocks, kStatThreads, 0, as_stream(stream)>>>(c, n, dfcsr_stat_chunk(n), n, 0, centre6(h_centre),
                                                                     &ws->partial[0][0], none, &ws->ticket[0], d_out);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_beam_cov_partial(const double* d_x, const double* d_px, const double* d_y, const double* d_py,
                                      const double* d_z, const double* d_pz, int64_t n_local, int64_t n_total,
                                      int32_t first_block, int32_t n_blocks, const double* h_centre, double* d_table,
                                      const uint64_t* h_peer_tables, int32_t n_peers, void* stream) {
    DFCSR_REQUIRE(d_table, "null table");
    DFCSR_REQUIRE(n_local == 0 || (d_x && d_px && d_y && d_py && d_z && d_pz), "null pointer");
    int rc = check_shard(n_local, n_total, first_block, n_blocks);
    if (rc) return rc;
    PeerTables pt;
    rc = peer_tables(h_peer_tables, n_peers, pt);
    if (rc) return rc;
    if (n_blocks == 0) return DFCSR_OK;
    SixPtr c;
    c.q[0] = d_x; c.q[1] = d_px; c.q[2] = d_y; c.q[3] = d_py; c.q[4] = d_z; c.q[5] = d_pz;
    cov6_kernel<<<(unsigned)n_blocks, kStatThreads, 0, as_stream(stream)>>>(c, n_local, dfcsr_stat_chunk(n_total), n_total, first_block,
                                                                            centre6(h_centre), d_table, pt, nullptr, nullptr);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_beam_cov_final(const double* d_table, int64_t n_total, const double* h_centre, double* d_out, void* stream) {
    DFCSR_REQUIRE(d_table && d_out && n_total >= 2, "bad argument");
    cov_final_kernel<<<1, kStatThreads, 0, as_stream(stream)>>>(d_table, n_total, centre6(h_centre), d_out);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_apply_kick(const double* d_x, const double* d_z, double* d_px, double* d_pz, int64_t n,
                                double slope, double intercept, const double* d_dE, const double* d_kick,
                                dfcsr_axis x_axis, dfcsr_axis z_axis, double step_size, double init_energy,
                                int32_t transverse_on, void* stream) {
    DFCSR_REQUIRE(n >= 0, "negative particle count");
    DFCSR_REQUIRE(d_dE && d_kick && (n == 0 || (d_x && d_z && d_px && d_pz)), "null pointer");
    DFCSR_REQUIRE(x_axis.n >= 2 && z_axis.n >= 2, "wake mesh needs at least 2 nodes per axis");
    DFCSR_REQUIRE(x_axis.stop > x_axis.start && z_axis.stop > z_axis.start, "wake mesh axes must be increasing");
    if (n == 0) return DFCSR_OK;
    Axis ax = make_axis(x_axis.start, x_axis.stop, x_axis.n), az = make_axis(z_axis.start, z_axis.stop, z_axis.n);
    const double factor = step_size * 1e6 / init_energy;   // beams.py:110,117
    // one particle per thread up to 148 x 64 CTAs (then grid-stride): the kernel is a pure stream, so memory-level
    // parallelism comes from resident threads, not from a per-thread loop
    long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 64 ? want : 148LL * 64);
    apply_kick_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_x, d_z, d_px, d_pz, n, slope, intercept, d_dE,
                                                           d_kick, ax, az, factor, transverse_on, 1.0 / ax.step,
                                                           1.0 / az.step);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

// ---- small results to pinned host memory without the copy engine ---------------------------------------------
// The 16 statistics of a pass are what the host waits for between two steps.  Sent with cudaMemcpyAsync they queue on the
// device-to-host copy engine behind any bulk download in flight on another stream (measured: +0.15 ms per step in the
// end-to-end loop of bench.py, whose 16 MB result download runs concurrently).  A one-warp kernel that stores them
// straight into mapped pinned memory does not.
__global__ void mirror_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

extern "C" int dfcsr_mirror_to_host(const double* d_src, double* h_dst, int32_t n, void* stream) {
    DFCSR_REQUIRE(d_src && h_dst && n > 0, "bad argument");
    double* mapped = nullptr;
    const cudaError_t e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&mapped), h_dst, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();      // not sticky: clear it so that the caller's fallback (cudaMemcpyAsync) starts clean
        return cuda_fail(e, "cudaHostGetDevicePointer (h_dst must be mapped pinned memory)");
    }
    mirror_kernel<<<1, 32, 0, as_stream(stream)>>>(d_src, mapped, n);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
