// make_df.cu — K2: deposit grids -> the five smoothed density-function fields.
//
// Replaces the numpy/scipy body of DF_tracker.get_DF after the two CIC calls (deposit.py:183-235):
//   thr = max(count)/velocity_threhold ; vx[count>thr] /= count                         :183-184
//   density, vx <- savgol_filter along axis 0 then axis 1 (mode='interp' edge fit)       :195-199
//   density /= trapz(trapz(density, x, axis=0), z)                                       :201-202
//   vx[density <= thr] = 0            (thr still in raw-count units: kept)               :204
//   density_x, density_z = np.gradient(density, x, z) ; vx_x = np.gradient(vx, x)[0]     :212-213
//   density_x, density_z, vx_x <- savgol o savgol                                        :215-224
//   thr2 = max(density)/velocity_threhold*8 ; vx_x[density<thr2] = mean(vx_x[density>thr2])  :233-235
//
// scipy's savgol_filter(mode='interp') is a symmetric FIR in the interior plus a fixed
// (window//2 x window) polynomial edge operator on the first/last `window` samples
// (scipy/signal/_savitzky_golay.py:244-258,261); both are computed on the
// host in fp64 for the (window, order) pair and handed in.
//
// Three tile kernels + one elementwise kernel, all staged through shared memory (no cluster / grid barriers):
//   A  per 32 x 32 output tile: the count and normalised-velocity tiles with their filter halos are loaded ONCE into
//      shared memory, smoothed along x and then along z there, and the raw smoothed density / velocity go out with the
//      tile's trapezoid sum and maximum; the last tile to finish adds the tile partials in tile order;
//   B  per tile: normalised density and masked velocity with halo (+1 for np.gradient), the three gradients, their two
//      smoothing passes, all in shared memory; partial sums for the masked mean of vx_x;
//   C  elementwise: vx_x below the second threshold is replaced by the masked mean.
// max(count), which everything depends on, comes with the deposit (dfcsr_deposit_cic_finish) or from a reduction launch.
// Every output element is produced by the same sequence of operations as scipy / numpy would apply to it (same taps,
// same order), and every reduction has a fixed order: bitwise reproducible.  100 x 100: 16 CTAs; 300 x 300: 100 CTAs.
// Latency-bound by construction (a few microseconds per launch); a roofline fraction is not meaningful.
#include <limits.h>
#include <math_constants.h>
#include <atomic>
#include "common.cuh"

namespace dfcsr {

constexpr int kTile = 32;         // output tile edge
constexpr int kDfThreads = 512;    // 16 warps: rows of a tile by warp, columns by lane
constexpr int kMaxWindow = 33;
constexpr int kPartials = 4;

struct DfOps {   // device-resident Savitzky-Golay operators (uploaded once per (window, order) by the caller)
    const double* taps;      // [window]
    const double* edge_lo;   // [window/2][window]
    const double* edge_hi;   // [window/2][window]
};

struct DfHeader {
    unsigned long long cmax_bits;    // max(count) as the bit pattern of a non-negative double (when reduced here)
    unsigned int ticket[2];
    double pad[6];
};

struct DfParams {
    const double* count;
    const double* vxsum;
    Axis ax, az;
    int window;
    DfOps ops;
    double velocity_threshold;
    double* fields;
    double* scalars;
    DfHeader* hdr;
    const unsigned long long* cmax_bits;   // where max(count) is found (the header's slot or the deposit's)
    double* partial;                       // [2][tiles][kPartials]
    double* t0;                            // raw smoothed density plane
    double* t1;                            // raw smoothed velocity plane
    int tiles_x, tiles_z;
    const double* lim;                     // {x_lo, x_hi, z_lo, z_hi} in device memory, or nullptr: the axes above are final
};

// dfcsr_get_df_from_stats: the axis end points arrive in device memory (the launch was enqueued before the host saw the
// statistics); numpy.linspace's step is rebuilt from them exactly as make_axis does on the host
__device__ __forceinline__ void axes_from_device(DfParams& P) {
    if (P.lim) {
        P.ax = make_axis(P.lim[0], P.lim[1], P.ax.n);
        P.az = make_axis(P.lim[2], P.lim[3], P.az.n);
    }
}

// grid limits of DF_tracker.get_DF from the statistics vector (deposit.py:160-171: mean -+ lim * sigma), with the
// roundings of the Python expression `mean - lim * sigma` (one multiplication, one addition, no contraction)
__global__ void df_limits_kernel(const double* __restrict__ stats, double xlim, double zlim, double* __restrict__ lim) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double mx = stats[DFCSR_S_MEAN_X], mz = stats[DFCSR_S_MEAN_Z];
        const double hx = __dmul_rn(xlim, stats[DFCSR_S_SIGMA_X]), hz = __dmul_rn(zlim, stats[DFCSR_S_SIGMA_Z]);
        lim[0] = __dsub_rn(mx, hx);
        lim[1] = __dadd_rn(mx, hx);
        lim[2] = __dsub_rn(mz, hz);
        lim[3] = __dadd_rn(mz, hz);
    }
}

// Savitzky-Golay along one axis for element i of a line of n samples; `at(k)` fetches sample k of the line.
// kW > 0: compile-time window (fully unrolled taps); kW = 0: runtime window.  Same operations in the same order either way.
template <int kW, typename Fetch>
__device__ __forceinline__ double sg_line(const double* taps, const double* edge_lo, const double* edge_hi, int window_rt, int n,
                                          int i, Fetch at) {
    const int window = kW > 0 ? kW : window_rt;
    const int half = window >> 1;
    double acc = 0.0;
    if (i < half) {
        const double* e = edge_lo + i * window;
#pragma unroll
        for (int k = 0; k < window; ++k) acc = fma(e[k], at(k), acc);
    } else if (i >= n - half) {
        const double* e = edge_hi + (i - (n - half)) * window;
#pragma unroll
        for (int k = 0; k < window; ++k) acc = fma(e[k], at(n - window + k), acc);
    } else {
#pragma unroll
        for (int k = 0; k < window; ++k) acc = fma(taps[k], at(i - half + k), acc);
    }
    return acc;
}

// first sample of the filter support of element i (interior: i - half; clamped at both ends)
__device__ __forceinline__ int sg_start(int i, int n, int window) {
    const int s = i - (window >> 1);
    return s < 0 ? 0 : (s > n - window ? n - window : s);
}

// np.gradient along one axis with coordinates (numpy/lib/_function_base_impl.py): second-order
// interior (uniform formula only if all node differences are bit-identical), first-order edges.
template <typename Fetch>
__device__ __forceinline__ double grad_line(const Axis& a, bool uniform, int i, Fetch at) {
    const int n = a.n;
    if (i == 0) return (at(1) - at(0)) / (axis_node(a, 1) - axis_node(a, 0));
    if (i == n - 1) return (at(n - 1) - at(n - 2)) / (axis_node(a, n - 1) - axis_node(a, n - 2));
    if (uniform) {
        double h = axis_node(a, 1) - axis_node(a, 0);
        return (at(i + 1) - at(i - 1)) / (2.0 * h);
    }
    double xm = axis_node(a, i - 1), x0 = axis_node(a, i), xp = axis_node(a, i + 1);
    double h_lo = x0 - xm, h_hi = xp - x0;
    double ca = -(h_hi) / (h_lo * (h_lo + h_hi));
    double cb = (h_hi - h_lo) / (h_lo * h_hi);
    double cc = h_lo / (h_hi * (h_lo + h_hi));
    return __dadd_rn(__dadd_rn(__dmul_rn(ca, at(i - 1)), __dmul_rn(cb, at(i))), __dmul_rn(cc, at(i + 1)));
}

// np.gradient's three-point coefficients of node i (the divisions are per node, not per grid cell): value =
// ca f(lo) + cb f(mid) + cc f(hi) with (lo, mid, hi) = (i-1, i, i+1), or the one-sided / uniform variants of grad_line
// expressed in the same form.  kind: 0 = three-term rounded sum (non-uniform interior), 1 = (f(hi) - f(lo)) / den.
struct GradCoef {
    double ca, cb, cc;     // kind 0: coefficients; kind 1: cc = denominator
    int lo, hi, kind;
};

__device__ __forceinline__ GradCoef grad_coef(const Axis& a, bool uniform, int i) {
    GradCoef g;
    const int n = a.n;
    g.ca = g.cb = 0.0;
    if (i == 0) { g.lo = 0; g.hi = 1; g.kind = 1; g.cc = axis_node(a, 1) - axis_node(a, 0); return g; }
    if (i == n - 1) { g.lo = n - 2; g.hi = n - 1; g.kind = 1; g.cc = axis_node(a, n - 1) - axis_node(a, n - 2); return g; }
    g.lo = i - 1; g.hi = i + 1;
    if (uniform) { g.kind = 1; g.cc = 2.0 * (axis_node(a, 1) - axis_node(a, 0)); return g; }
    const double xm = axis_node(a, i - 1), x0 = axis_node(a, i), xp = axis_node(a, i + 1);
    const double h_lo = x0 - xm, h_hi = xp - x0;
    g.kind = 0;
    g.ca = -(h_hi) / (h_lo * (h_lo + h_hi));
    g.cb = (h_hi - h_lo) / (h_lo * h_hi);
    g.cc = h_lo / (h_hi * (h_lo + h_hi));
    return g;
}

__device__ __forceinline__ double grad_apply(const GradCoef& g, double f_lo, double f_mid, double f_hi) {
    if (g.kind == 1) return (f_hi - f_lo) / g.cc;
    return __dadd_rn(__dadd_rn(__dmul_rn(g.ca, f_lo), __dmul_rn(g.cb, f_mid)), __dmul_rn(g.cc, f_hi));
}

__device__ __forceinline__ bool axis_uniform(const Axis& a) {
    // (np.diff(nodes) == diff[0]).all(), evaluated redundantly by every CTA
    double h0 = axis_node(a, 1) - axis_node(a, 0);
    int ok = 1;
    for (int i = threadIdx.x; i < a.n - 1; i += blockDim.x)
        ok &= ((axis_node(a, i + 1) - axis_node(a, i)) == h0);
    return __syncthreads_and(ok) != 0;
}

__device__ __forceinline__ double trapz_weight(const Axis& a, int i) {
    double w = 0.0;
    if (i > 0) w += axis_node(a, i) - axis_node(a, i - 1);
    if (i < a.n - 1) w += axis_node(a, i + 1) - axis_node(a, i);
    return 0.5 * w;
}

// block totals of NV values (sum, or max where is_max) in thread 0
template <int NV>
__device__ __forceinline__ void cta_total(double (&v)[NV], const bool (&is_max)[NV]) {
    __shared__ double sm[kDfThreads / 32][kPartials];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = is_max[k] ? warp_max(v[k]) : warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
        for (int k = 0; k < NV; ++k) sm[warp][k] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < NV; ++k) {
            double s = sm[0][k];
            for (int w = 1; w < kDfThreads / 32; ++w) s = is_max[k] ? fmax(s, sm[w][k]) : s + sm[w][k];
            v[k] = s;
        }
    }
}

// publish this tile's partials; true (in every thread) in the last tile to arrive, which may then read all partials
__device__ __forceinline__ bool publish_and_ticket(const double* v, int nv, double* partial, int tile, int tiles, unsigned int* ticket) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        for (int k = 0; k < nv; ++k) partial[(size_t)tile * kPartials + k] = v[k];
        __threadfence();
        const bool last = (atomicAdd(ticket, 1u) == (unsigned)tiles - 1u);
        if (last) { __threadfence(); *ticket = 0u; }
        s_last = last ? 1 : 0;
    }
    __syncthreads();
    return s_last != 0;
}

// fixed-order total of the tile partials by the whole CTA: thread t takes tiles t, t + 256, ... (loads first, L2-coherent),
// then the fixed block reduction; result in thread 0
template <int NV>
__device__ __forceinline__ void partial_total(const double* partial, int tiles, const bool (&is_max)[NV], double (&tot)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = is_max[k] ? -CUDART_INF : 0.0;
    for (int t = threadIdx.x; t < tiles; t += kDfThreads) {
        double p[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) p[k] = __ldcg(partial + (size_t)t * kPartials + k);
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] = is_max[k] ? fmax(tot[k], p[k]) : tot[k] + p[k];
    }
    cta_total<NV>(tot, is_max);
}

struct TileGeom {
    int i0, i1, j0, j1;        // output tile (inclusive)
    int r0, r1, c0, c1;        // rows / columns of the smoothing input region (inclusive)
    int rw, cw;                // region extents
};

__device__ __forceinline__ TileGeom tile_geom(const DfParams& P) {
    TileGeom g;
    const int tx = blockIdx.x / P.tiles_z, tz = blockIdx.x - tx * P.tiles_z;
    g.i0 = tx * kTile; g.i1 = min(g.i0 + kTile, P.ax.n) - 1;
    g.j0 = tz * kTile; g.j1 = min(g.j0 + kTile, P.az.n) - 1;
    g.r0 = sg_start(g.i0, P.ax.n, P.window); g.r1 = sg_start(g.i1, P.ax.n, P.window) + P.window - 1;
    g.c0 = sg_start(g.j0, P.az.n, P.window); g.c1 = sg_start(g.j1, P.az.n, P.window) + P.window - 1;
    g.rw = g.r1 - g.r0 + 1;
    g.cw = g.c1 - g.c0 + 1;
    return g;
}

// operators to shared memory: taps[w], edge_lo[h][w], edge_hi[h][w]
__device__ __forceinline__ void load_ops(const DfOps& ops, int window, double* s_taps, double* s_lo, double* s_hi) {
    const int half = window >> 1;
    for (int k = threadIdx.x; k < window; k += kDfThreads) s_taps[k] = ops.taps[k];
    for (int k = threadIdx.x; k < half * window; k += kDfThreads) { s_lo[k] = ops.edge_lo[k]; s_hi[k] = ops.edge_hi[k]; }
}

// smooth the region `in` (rows r0.., columns c0..; extents rw x cw) along x into `tmp` (tile rows x cw), then along z
// into out(i, j) for the tile; `emit(i, j, value)` consumes the result.  Rows go by warp, columns by lane.
template <int kW, typename Emit>
__device__ __forceinline__ void smooth_tile_w(const DfParams& P, const TileGeom& g, const double* in, double* tmp,
                                              const double* s_taps, const double* s_lo, const double* s_hi, Emit emit) {
    const int nx = P.ax.n, nz = P.az.n, w = P.window;
    const int th = g.i1 - g.i0 + 1, tw = g.j1 - g.j0 + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kDfThreads / 32;
    for (int ti = warp; ti < th; ti += kWarps)
        for (int cc = lane; cc < g.cw; cc += 32)
            tmp[ti * g.cw + cc] = sg_line<kW>(s_taps, s_lo, s_hi, w, nx, g.i0 + ti, [&](int k) { return in[(k - g.r0) * g.cw + cc]; });
    __syncthreads();
    for (int ti = warp; ti < th; ti += kWarps)
        for (int tj = lane; tj < tw; tj += 32) {
            const double v = sg_line<kW>(s_taps, s_lo, s_hi, w, nz, g.j0 + tj, [&](int k) { return tmp[ti * g.cw + (k - g.c0)]; });
            emit(g.i0 + ti, g.j0 + tj, v);
        }
    __syncthreads();
}

template <typename Emit>
__device__ __forceinline__ void smooth_tile(const DfParams& P, const TileGeom& g, const double* in, double* tmp,
                                            const double* s_taps, const double* s_lo, const double* s_hi, Emit emit) {
    if (P.window == 9) smooth_tile_w<9>(P, g, in, tmp, s_taps, s_lo, s_hi, emit);        // chicane_config.yaml:16
    else if (P.window == 5) smooth_tile_w<5>(P, g, in, tmp, s_taps, s_lo, s_hi, emit);   // the hard-coded branch, deposit.py:167
    else smooth_tile_w<0>(P, g, in, tmp, s_taps, s_lo, s_hi, emit);
}

// max(count) when the deposit did not deliver it
__global__ void __launch_bounds__(kDfThreads)
count_max_kernel(const double* __restrict__ count, int cells, unsigned long long* slot) {
    double m = 0.0;
    for (int c = blockIdx.x * kDfThreads + threadIdx.x; c < cells; c += gridDim.x * kDfThreads) m = fmax(m, count[c]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(slot, (unsigned long long)__double_as_longlong(m));
}

// ---- phase A -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDfThreads)
make_df_phase_a(DfParams P) {
    extern __shared__ double smem[];
    axes_from_device(P);
    const TileGeom g = tile_geom(P);
    const int nz = P.az.n;
    const int half = P.window >> 1;
    double* s_taps = smem;
    double* s_lo = s_taps + kMaxWindow;
    double* s_hi = s_lo + half * P.window;
    double* in_c = s_hi + half * P.window;
    double* in_v = in_c + g.rw * g.cw;
    double* tmp = in_v + g.rw * g.cw;
    double* wx = tmp + kTile * g.cw;            // trapezoid weights of the tile's rows and columns
    double* wz = wx + kTile;
    load_ops(P.ops, P.window, s_taps, s_lo, s_hi);
    if (threadIdx.x < kTile) wx[threadIdx.x] = trapz_weight(P.ax, min(g.i0 + (int)threadIdx.x, P.ax.n - 1));
    else if (threadIdx.x < 2 * kTile) wz[threadIdx.x - kTile] = trapz_weight(P.az, min(g.j0 + (int)threadIdx.x - kTile, nz - 1));
    const double cmax = __longlong_as_double((long long)*P.cmax_bits);
    const double thr = cmax / P.velocity_threshold;
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int rr = warp; rr < g.rw; rr += kDfThreads / 32)
            for (int cc = lane; cc < g.cw; cc += 32) {
                const size_t o = (size_t)(g.r0 + rr) * nz + (g.c0 + cc);
                const double cn = P.count[o], vs = P.vxsum[o];
                in_c[rr * g.cw + cc] = cn;
                in_v[rr * g.cw + cc] = (cn > thr) ? vs / cn : vs;            // deposit.py:184
            }
    }
    __syncthreads();
    double v[2] = {0.0, -CUDART_INF};
    smooth_tile(P, g, in_c, tmp, s_taps, s_lo, s_hi, [&](int i, int j, double d) {
        P.t0[(size_t)i * nz + j] = d;
        v[0] = fma(wx[i - g.i0] * wz[j - g.j0], d, v[0]);
        v[1] = fmax(v[1], d);
    });
    smooth_tile(P, g, in_v, tmp, s_taps, s_lo, s_hi, [&](int i, int j, double wv) { P.t1[(size_t)i * nz + j] = wv; });
    const bool mx[2] = {false, true};
    cta_total<2>(v, mx);
    const int tiles = P.tiles_x * P.tiles_z;
    if (publish_and_ticket(v, 2, P.partial, blockIdx.x, tiles, &P.hdr->ticket[0])) {
        double tot[2];
        partial_total<2>(P.partial, tiles, mx, tot);
        if (threadIdx.x != 0) return;
        const double dsum = tot[0];
        const double dmax = tot[1] / dsum;               // max of the normalised density
        P.scalars[0] = cmax;
        P.scalars[1] = thr;
        P.scalars[2] = dsum;
        P.scalars[3] = dmax;
        P.scalars[7] = dmax / P.velocity_threshold * 8.0;
    }
}

// ---- phase B -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDfThreads)
make_df_phase_b(DfParams P) {
    extern __shared__ double smem[];
    axes_from_device(P);
    const TileGeom g = tile_geom(P);
    const int nx = P.ax.n, nz = P.az.n;
    const int cells = nx * nz;
    const int half = P.window >> 1;
    // region of normalised density / masked velocity: the smoothing region widened by one node for np.gradient
    const int er0 = max(g.r0 - 1, 0), er1 = min(g.r1 + 1, nx - 1), ec0 = max(g.c0 - 1, 0), ec1 = min(g.c1 + 1, nz - 1);
    const int ecw = ec1 - ec0 + 1, erw = er1 - er0 + 1;
    double* s_taps = smem;
    double* s_lo = s_taps + kMaxWindow;
    double* s_hi = s_lo + half * P.window;
    double* dn = s_hi + half * P.window;       // erw x ecw
    double* vm = dn + erw * ecw;
    double* grad = vm + erw * ecw;             // rw x cw
    double* tmp = grad + g.rw * g.cw;          // tile rows x cw
    GradCoef* gcx = reinterpret_cast<GradCoef*>(tmp + (g.i1 - g.i0 + 1) * g.cw);     // rw entries
    GradCoef* gcz = gcx + g.rw;                                                      // cw entries
    load_ops(P.ops, P.window, s_taps, s_lo, s_hi);
    const bool uni_x = axis_uniform(P.ax);
    const bool uni_z = axis_uniform(P.az);
    for (int e = threadIdx.x; e < g.rw + g.cw; e += kDfThreads) {
        if (e < g.rw) gcx[e] = grad_coef(P.ax, uni_x, g.r0 + e);
        else gcz[e - g.rw] = grad_coef(P.az, uni_z, g.c0 + (e - g.rw));
    }
    const double thr = P.scalars[1], dsum = P.scalars[2], thr2 = P.scalars[7];
    double* density = P.fields + (size_t)DFCSR_DENSITY * cells;
    double* density_x = P.fields + (size_t)DFCSR_DENSITY_X * cells;
    double* density_z = P.fields + (size_t)DFCSR_DENSITY_Z * cells;
    double* vx = P.fields + (size_t)DFCSR_VX * cells;
    double* vx_x = P.fields + (size_t)DFCSR_VX_X * cells;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kDfThreads / 32;
    for (int rr = warp; rr < erw; rr += kWarps)
        for (int cc = lane; cc < ecw; cc += 32) {
            const int i = er0 + rr, j = ec0 + cc;
            const size_t o = (size_t)i * nz + j;
            const double d = P.t0[o] / dsum;                 // deposit.py:202
            const double wv = (d <= thr) ? 0.0 : P.t1[o];    // deposit.py:204 (thr in raw-count units: kept)
            dn[rr * ecw + cc] = d;
            vm[rr * ecw + cc] = wv;
            if (i >= g.i0 && i <= g.i1 && j >= g.j0 && j <= g.j1) { density[o] = d; vx[o] = wv; }
        }
    __syncthreads();
    double v[4] = {0.0, 0.0, 0.0, 0.0};        // sum / count of vx_x where density > thr2; sum where == thr2; count where <
    for (int f = 0; f < 3; ++f) {
        // gradient field f on the smoothing region: d(density)/dx, d(density)/dz, d(vx)/dx   (deposit.py:212-213)
        for (int rr = warp; rr < g.rw; rr += kWarps)
            for (int cc = lane; cc < g.cw; cc += 32) {
                const int i = g.r0 + rr, j = g.c0 + cc;
                const double* src = (f == 2) ? vm : dn;
                double gv;
                if (f == 1) {
                    const GradCoef c = gcz[cc];
                    const double* row = src + (i - er0) * ecw - ec0;
                    gv = grad_apply(c, row[c.lo], row[j], row[c.hi]);
                } else {
                    const GradCoef c = gcx[rr];
                    const double* col = src + (j - ec0) - er0 * ecw;
                    gv = grad_apply(c, col[c.lo * ecw], col[i * ecw], col[c.hi * ecw]);
                }
                grad[rr * g.cw + cc] = gv;
            }
        __syncthreads();
        double* out = (f == 0) ? density_x : ((f == 1) ? density_z : vx_x);
        smooth_tile(P, g, grad, tmp, s_taps, s_lo, s_hi, [&](int i, int j, double sv) {
            out[(size_t)i * nz + j] = sv;
            if (f == 2) {
                const double d = dn[(i - er0) * ecw + (j - ec0)];
                if (d > thr2) { v[0] += sv; v[1] += 1.0; }
                else if (d < thr2) v[3] += 1.0;
                else v[2] += sv;                          // == thr2 (or NaN): keeps its own value
            }
        });
    }
    const bool mx[4] = {false, false, false, false};
    cta_total<4>(v, mx);
    const int tiles = P.tiles_x * P.tiles_z;
    double* partial_b = P.partial + (size_t)tiles * kPartials;
    if (publish_and_ticket(v, 4, partial_b, blockIdx.x, tiles, &P.hdr->ticket[1])) {
        double tot[4];
        partial_total<4>(partial_b, tiles, mx, tot);
        if (threadIdx.x != 0) return;
        const double s_gt = tot[0], n_gt = tot[1], s_eq = tot[2], n_lt = tot[3];
        const double mmean = s_gt / n_gt;                // deposit.py:235
        P.scalars[5] = mmean;
        P.scalars[6] = n_gt;
        // mean of vx_x after the fill of phase C = fill value of the re-gridding (deposit.py:332)
        P.scalars[4] = (n_lt > 0.0 ? (s_gt + s_eq) + n_lt * mmean : (s_gt + s_eq)) / (double)cells;
    }
}

// ---- phase C: vx_x[density < thr2] = mean(vx_x[density > thr2])  (deposit.py:233-235) -------------------------------
__global__ void __launch_bounds__(kDfThreads)
make_df_phase_c(DfParams P) {
    const int cells = P.ax.n * P.az.n;
    const double thr2 = P.scalars[7], mmean = P.scalars[5];
    const double* density = P.fields + (size_t)DFCSR_DENSITY * cells;
    double* vx_x = P.fields + (size_t)DFCSR_VX_X * cells;
    for (int c = blockIdx.x * kDfThreads + threadIdx.x; c < cells; c += gridDim.x * kDfThreads)
        if (density[c] < thr2) vx_x[c] = mmean;
}

}  // namespace dfcsr

using namespace dfcsr;

static long long df_tiles(int nx, int nz) { return (long long)((nx + kTile - 1) / kTile) * ((nz + kTile - 1) / kTile); }

extern "C" int64_t dfcsr_make_df_workspace(int32_t nx, int32_t nz) {
    if (nx < 1 || nz < 1) return 0;
    return (int64_t)sizeof(DfHeader) + (int64_t)2 * df_tiles(nx, nz) * kPartials * (int64_t)sizeof(double) +
           (int64_t)2 * nx * nz * (int64_t)sizeof(double);
}

static int make_df_impl(const double* d_count, const double* d_vxsum, dfcsr_axis x_axis, dfcsr_axis z_axis,
                        int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                        double velocity_threshold, const uint64_t* d_count_max, double* d_fields, double* d_scalars,
                        void* d_workspace, void* stream, const double* d_lim) {
    DFCSR_REQUIRE(d_count && d_vxsum && d_fields && d_scalars && d_workspace, "null device pointer");
    DFCSR_REQUIRE(d_taps && (window < 3 || (d_edge_lo && d_edge_hi)), "null operator pointer");
    DFCSR_REQUIRE(window >= 1 && (window & 1), "window must be odd and positive");
    if (window > kMaxWindow) {
        set_error("dfcsr_make_df: filter_window %d exceeds the supported %d", window, kMaxWindow);
        return DFCSR_ERR_UNSUPPORTED;
    }
    DFCSR_REQUIRE(x_axis.n >= window && z_axis.n >= window && x_axis.n >= 2 && z_axis.n >= 2,
                  "grid smaller than the filter window");
    DFCSR_REQUIRE((long long)x_axis.n * z_axis.n < (1LL << 30), "grid too large");
    cudaStream_t st = as_stream(stream);
    const long long cells = (long long)x_axis.n * z_axis.n;
    const long long tiles = df_tiles(x_axis.n, z_axis.n);
    char* ws = reinterpret_cast<char*>(d_workspace);
    DfParams P;
    P.count = d_count;
    P.vxsum = d_vxsum;
    P.ax = make_axis(x_axis.start, x_axis.stop, x_axis.n);
    P.az = make_axis(z_axis.start, z_axis.stop, z_axis.n);
    P.window = window;
    P.ops.taps = d_taps;
    P.ops.edge_lo = d_edge_lo;
    P.ops.edge_hi = d_edge_hi;
    P.velocity_threshold = velocity_threshold;
    P.fields = d_fields;
    P.scalars = d_scalars;
    P.hdr = reinterpret_cast<DfHeader*>(ws);
    P.partial = reinterpret_cast<double*>(ws + sizeof(DfHeader));
    P.t0 = P.partial + 2 * tiles * kPartials;
    P.t1 = P.t0 + cells;
    P.tiles_x = (x_axis.n + kTile - 1) / kTile;
    P.tiles_z = (z_axis.n + kTile - 1) / kTile;
    P.lim = d_lim;
    DFCSR_CUDA_OK(cudaMemsetAsync(P.hdr, 0, sizeof(DfHeader), st));      // tickets and the max slot
    int launches = 3;
    if (d_count_max) {
        P.cmax_bits = reinterpret_cast<const unsigned long long*>(d_count_max);
    } else {
        P.cmax_bits = &P.hdr->cmax_bits;
        long long want = (cells + kDfThreads - 1) / kDfThreads;
        count_max_kernel<<<(unsigned)(want < 64 ? want : 64), kDfThreads, 0, st>>>(d_count, (int)cells, &P.hdr->cmax_bits);
        ++launches;
    }
    const int half = window >> 1;
    const int reg = kTile + 2 * half;                  // largest smoothing-region edge
    const size_t ops_words = kMaxWindow + 2 * (size_t)half * window;
    const size_t smem_a = (ops_words + 2 * (size_t)reg * reg + (size_t)kTile * reg + 2 * kTile) * sizeof(double);
    const size_t smem_b = (ops_words + 2 * (size_t)(reg + 2) * (reg + 2) + (size_t)reg * reg + (size_t)kTile * reg) * sizeof(double) +
                          2 * (size_t)reg * sizeof(GradCoef);
    // opt in to more than 48 KB of dynamic shared memory only when a window needs it, and only once per size (the call
    // costs several microseconds, as much as one of these kernels)
    static std::atomic<size_t> granted_a{48 * 1024}, granted_b{48 * 1024};
    if (smem_a > granted_a.load()) {
        DFCSR_CUDA_OK(cudaFuncSetAttribute(make_df_phase_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
        granted_a.store(smem_a);
    }
    if (smem_b > granted_b.load()) {
        DFCSR_CUDA_OK(cudaFuncSetAttribute(make_df_phase_b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        granted_b.store(smem_b);
    }
    make_df_phase_a<<<(unsigned)tiles, kDfThreads, smem_a, st>>>(P);
    make_df_phase_b<<<(unsigned)tiles, kDfThreads, smem_b, st>>>(P);
    {
        long long want = (cells + kDfThreads - 1) / kDfThreads;
        make_df_phase_c<<<(unsigned)(want < 148LL * 4 ? want : 148LL * 4), kDfThreads, 0, st>>>(P);
    }
    count_launch(launches);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_make_df(const double* d_count, const double* d_vxsum, dfcsr_axis x_axis, dfcsr_axis z_axis,
                             int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                             double velocity_threshold, const uint64_t* d_count_max, double* d_fields, double* d_scalars,
                             void* d_workspace, void* stream) {
    return make_df_impl(d_count, d_vxsum, x_axis, z_axis, window, d_taps, d_edge_lo, d_edge_hi, velocity_threshold,
                        d_count_max, d_fields, d_scalars, d_workspace, stream, nullptr);
}

// DF_tracker.get_DF enqueued BEFORE the host has the statistics (include/dfcsr_b200.h): limits kernel, both deposit
// stages and the three density-function kernels, all reading the grid limits and max|px| from device memory
extern "C" int dfcsr_get_df_from_stats(const double* d_x, const double* d_z, const double* d_px, int64_t n,
                                       const double* d_stats, double xlim, double zlim, int32_t nx, int32_t nz,
                                       double* d_limits, int64_t* d_q, double* d_count, double* d_vxsum,
                                       uint64_t* d_count_max, int32_t window, const double* d_taps,
                                       const double* d_edge_lo, const double* d_edge_hi, double velocity_threshold,
                                       double* d_fields, double* d_scalars, void* d_workspace, void* stream) {
    DFCSR_REQUIRE(d_stats && d_limits && d_q && d_count && d_vxsum && d_count_max, "null scratch / output pointer");
    DFCSR_REQUIRE(n >= 1 && d_x && d_z && d_px, "the particle arrays (with px) are required");
    DFCSR_REQUIRE(nx >= 1 && nz >= 1 && (long long)nx * nz < (1LL << 30), "bad grid");
    cudaStream_t st = as_stream(stream);
    df_limits_kernel<<<1, 32, 0, st>>>(d_stats, xlim, zlim, d_limits);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    int rc = deposit_one_gpu_from_device(d_x, d_z, d_px, n, nx, nz, d_limits,
                                         reinterpret_cast<const unsigned long long*>(d_stats + DFCSR_S_ABSMAX_PX),
                                         reinterpret_cast<long long*>(d_q), d_count, d_vxsum,
                                         reinterpret_cast<unsigned long long*>(d_count_max), st);
    if (rc) return rc;
    dfcsr_axis xa, za;                     // dimensions only: the kernels take the end points from d_limits
    xa.start = 0.0; xa.stop = 1.0; xa.n = nx; xa._pad = 0;
    za.start = 0.0; za.stop = 1.0; za.n = nz; za._pad = 0;
    return make_df_impl(d_count, d_vxsum, xa, za, window, d_taps, d_edge_lo, d_edge_hi, velocity_threshold, d_count_max,
                        d_fields, d_scalars, d_workspace, stream, d_limits);
}

extern "C" int dfcsr_df_limits(const double* d_stats, double xlim, double zlim, double* d_limits, void* stream) {
    DFCSR_REQUIRE(d_stats && d_limits, "null pointer");
    df_limits_kernel<<<1, 32, 0, as_stream(stream)>>>(d_stats, xlim, zlim, d_limits);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}

extern "C" int dfcsr_make_df_dev(const double* d_count, const double* d_vxsum, int32_t nx, int32_t nz, const double* d_limits,
                                 int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                                 double velocity_threshold, const uint64_t* d_count_max, double* d_fields, double* d_scalars,
                                 void* d_workspace, void* stream) {
    DFCSR_REQUIRE(d_limits, "null limits pointer");
    dfcsr_axis xa, za;                     // dimensions only: the kernels take the end points from d_limits
    xa.start = 0.0; xa.stop = 1.0; xa.n = nx; xa._pad = 0;
    za.start = 0.0; za.stop = 1.0; za.n = nz; za._pad = 0;
    return make_df_impl(d_count, d_vxsum, xa, za, window, d_taps, d_edge_lo, d_edge_hi, velocity_threshold, d_count_max,
                        d_fields, d_scalars, d_workspace, stream, d_limits);
}

// DF_tracker.get_DF on one GPU in one call (include/dfcsr_b200.h): the three stages back to back on one stream
extern "C" int dfcsr_get_df(const double* d_x, const double* d_z, const double* d_px, int64_t n, dfcsr_axis x_axis,
                            dfcsr_axis z_axis, double absmax_px, int64_t* d_q, double* d_count, double* d_vxsum,
                            uint64_t* d_count_max, int32_t window, const double* d_taps, const double* d_edge_lo,
                            const double* d_edge_hi, double velocity_threshold, double* d_fields, double* d_scalars,
                            void* d_workspace, void* stream) {
    DFCSR_REQUIRE(d_q && d_count && d_vxsum && d_count_max, "null scratch / output pointer");
    int rc = dfcsr_deposit_cic_q(d_x, d_z, d_px, n, n, x_axis.n, x_axis.start, x_axis.stop, z_axis.n, z_axis.start, z_axis.stop,
                                 absmax_px, d_q, stream);
    if (rc) return rc;
    const uint64_t self = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(d_q));
    rc = dfcsr_deposit_cic_finish(&self, 1, x_axis.n, z_axis.n, n, absmax_px, d_count, d_vxsum, d_count_max, stream);
    if (rc) return rc;
    return dfcsr_make_df(d_count, d_vxsum, x_axis, z_axis, window, d_taps, d_edge_lo, d_edge_hi, velocity_threshold,
                         d_count_max, d_fields, d_scalars, d_workspace, stream);
}
