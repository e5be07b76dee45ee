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
// One launch: a single thread-block cluster (8 CTAs x 512 threads, distributed over 8 SMs).  The
// grids are at most a few hundred KB and live in L2; the ten dependent phases are separated by
// cluster barriers (release/acquire at cluster scope) instead of ten kernel launches, and the
// grid-wide reductions (max, trapezoid sum, masked mean) are two-level and fixed-order, hence
// bitwise reproducible.  This kernel is latency-bound (~10 barriers); a roofline fraction is not
// meaningful for it.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace dfcsr {

constexpr int kDfCtas = 8;        // CTAs of the cluster launch (grids up to kClusterCells cells)
constexpr int kDfMaxCtas = 64;    // CTAs of the cooperative launch used for larger grids
constexpr int kClusterCells = 16384;
constexpr int kDfThreads = 512;
constexpr int kMaxWindow = 33;
constexpr int kPartials = 4;

struct DfOps {   // device-resident Savitzky-Golay operators (uploaded once per (window, order) by the caller)
    const double* taps;      // [window]
    const double* edge_lo;   // [window/2][window]
    const double* edge_hi;   // [window/2][window]
};

struct DfWorkspace {
    double partial[8][kDfMaxCtas][kPartials];
    // followed by 6 scratch planes of nx*nz doubles
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
    DfWorkspace* ws;
    double* scratch;
};

// Savitzky-Golay along one axis for element (i, j); `stride` is the element stride of the filtered
// axis, `n` its length; `at(k)` fetches the k-th sample of the line through (i, j).
template <typename Fetch>
__device__ __forceinline__ double sg_line(const DfOps& ops, int window, int n, int i, Fetch at) {
    const int half = window >> 1;
    double acc = 0.0;
    if (i < half) {
        const double* e = ops.edge_lo + i * window;
        for (int k = 0; k < window; ++k) acc = fma(e[k], at(k), acc);
    } else if (i >= n - half) {
        const double* e = ops.edge_hi + (i - (n - half)) * window;
        for (int k = 0; k < window; ++k) acc = fma(e[k], at(n - window + k), acc);
    } else {
        for (int k = 0; k < window; ++k) acc = fma(ops.taps[k], at(i - half + k), acc);
    }
    return acc;
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

// block-level reduction of NV values (sum or max per slot) -> partial[slot][cta]
template <int NV>
__device__ __forceinline__ void cta_reduce(double (&v)[NV], const bool (&is_max)[NV], double (*out)[kPartials], int cta) {
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
            out[cta][k] = s;
        }
    }
}

__device__ __forceinline__ double combine(double (*p)[kPartials], int slot, bool is_max) {
    double s = ((volatile double*)&p[0][slot])[0];
    for (int c = 1; c < (int)gridDim.x; ++c) {
        double t = ((volatile double*)&p[c][slot])[0];
        s = is_max ? fmax(s, t) : s + t;
    }
    return s;
}

struct ClusterSync {   // 8 CTAs on 8 SMs of one GPC: hardware cluster barrier (release/acquire at cluster scope)
    __device__ __forceinline__ void sync() { cg::this_cluster().sync(); }
};
struct GridSync {      // cooperative launch: grid-wide barrier for grids too large for one cluster
    __device__ __forceinline__ void sync() { cg::this_grid().sync(); }
};

template <typename Sync>
__device__ __forceinline__ void make_df_body(const DfParams& P, Sync cluster) {
    const int cta = blockIdx.x;
    const int nx = P.ax.n, nz = P.az.n;
    const int cells = nx * nz;
    const int tid = cta * kDfThreads + threadIdx.x;
    const int nthreads = (int)gridDim.x * kDfThreads;
    const DfOps ops = P.ops;
    const int window = P.window;
    double* T0 = P.scratch;
    double* T1 = T0 + cells;
    double* T2 = T1 + cells;
    double* U0 = T2 + cells;
    double* U1 = U0 + cells;
    double* U2 = U1 + cells;
    double* density = P.fields + (size_t)DFCSR_DENSITY * cells;
    double* density_x = P.fields + (size_t)DFCSR_DENSITY_X * cells;
    double* density_z = P.fields + (size_t)DFCSR_DENSITY_Z * cells;
    double* vx = P.fields + (size_t)DFCSR_VX * cells;
    double* vx_x = P.fields + (size_t)DFCSR_VX_X * cells;
    const double* count = P.count;
    const double* vxsum = P.vxsum;

    const bool uni_x = axis_uniform(P.ax);
    const bool uni_z = axis_uniform(P.az);

    // P0: max(count)
    {
        double v[1] = {-INFINITY};
        for (int c = tid; c < cells; c += nthreads) v[0] = fmax(v[0], count[c]);
        const bool mx[1] = {true};
        cta_reduce<1>(v, mx, P.ws->partial[0], cta);
    }
    cluster.sync();
    const double cmax = combine(P.ws->partial[0], 0, true);
    const double thr = cmax / P.velocity_threshold;

    // P1: Savitzky-Golay along axis 0 (x) of count and of the normalised velocity
    for (int c = tid; c < cells; c += nthreads) {
        const int i = c / nz, j = c - i * nz;
        T0[c] = sg_line(ops, window, nx, i, [&](int k) { return count[k * nz + j]; });
        T1[c] = sg_line(ops, window, nx, i, [&](int k) {
            double cn = count[k * nz + j], vs = vxsum[k * nz + j];
            return (cn > thr) ? vs / cn : vs;
        });
    }
    cluster.sync();

    // P2: along axis 1 (z); trapezoid normalisation and max of the smoothed density
    {
        double v[2] = {0.0, -INFINITY};
        for (int c = tid; c < cells; c += nthreads) {
            const int i = c / nz, j = c - i * nz;
            double d = sg_line(ops, window, nz, j, [&](int k) { return T0[i * nz + k]; });
            double w = sg_line(ops, window, nz, j, [&](int k) { return T1[i * nz + k]; });
            density[c] = d;
            vx[c] = w;
            v[0] = fma(trapz_weight(P.ax, i) * trapz_weight(P.az, j), d, v[0]);
            v[1] = fmax(v[1], d);
        }
        const bool mx[2] = {false, true};
        cta_reduce<2>(v, mx, P.ws->partial[1], cta);
    }
    cluster.sync();
    const double dsum = combine(P.ws->partial[1], 0, false);
    const double dmax = combine(P.ws->partial[1], 1, true) / dsum;   // max of the normalised density

    // P3: normalise, zero the velocity where the (normalised) density is below the raw-count threshold
    for (int c = tid; c < cells; c += nthreads) {
        double d = density[c] / dsum;
        density[c] = d;
        if (d <= thr) vx[c] = 0.0;
    }
    cluster.sync();

    // P4: gradients
    for (int c = tid; c < cells; c += nthreads) {
        const int i = c / nz, j = c - i * nz;
        T0[c] = grad_line(P.ax, uni_x, i, [&](int k) { return density[k * nz + j]; });
        T1[c] = grad_line(P.az, uni_z, j, [&](int k) { return density[i * nz + k]; });
        T2[c] = grad_line(P.ax, uni_x, i, [&](int k) { return vx[k * nz + j]; });
    }
    cluster.sync();

    // P5: smooth the three gradients along axis 0
    for (int c = tid; c < cells; c += nthreads) {
        const int i = c / nz, j = c - i * nz;
        U0[c] = sg_line(ops, window, nx, i, [&](int k) { return T0[k * nz + j]; });
        U1[c] = sg_line(ops, window, nx, i, [&](int k) { return T1[k * nz + j]; });
        U2[c] = sg_line(ops, window, nx, i, [&](int k) { return T2[k * nz + j]; });
    }
    cluster.sync();

    // P6: ... and along axis 1; masked sum of vx_x over cells above the second threshold
    const double thr2 = dmax / P.velocity_threshold * 8.0;
    {
        double v[2] = {0.0, 0.0};
        for (int c = tid; c < cells; c += nthreads) {
            const int i = c / nz, j = c - i * nz;
            density_x[c] = sg_line(ops, window, nz, j, [&](int k) { return U0[i * nz + k]; });
            density_z[c] = sg_line(ops, window, nz, j, [&](int k) { return U1[i * nz + k]; });
            double g = sg_line(ops, window, nz, j, [&](int k) { return U2[i * nz + k]; });
            vx_x[c] = g;
            if (density[c] > thr2) { v[0] += g; v[1] += 1.0; }
        }
        const bool mx[2] = {false, false};
        cta_reduce<2>(v, mx, P.ws->partial[2], cta);
    }
    cluster.sync();
    const double msum = combine(P.ws->partial[2], 0, false);
    const double mcnt = combine(P.ws->partial[2], 1, false);
    const double mmean = msum / mcnt;

    // P7: fill vx_x below the threshold with the masked mean; total mean = re-gridding fill value
    {
        double v[1] = {0.0};
        for (int c = tid; c < cells; c += nthreads) {
            double g = vx_x[c];
            if (density[c] < thr2) { g = mmean; vx_x[c] = g; }
            v[0] += g;
        }
        const bool mx[1] = {false};
        cta_reduce<1>(v, mx, P.ws->partial[3], cta);
    }
    cluster.sync();
    if (cta == 0 && threadIdx.x == 0) {
        P.scalars[0] = cmax;
        P.scalars[1] = thr;
        P.scalars[2] = dsum;
        P.scalars[3] = dmax;
        P.scalars[4] = combine(P.ws->partial[3], 0, false) / (double)cells;
        P.scalars[5] = mmean;
        P.scalars[6] = mcnt;
        P.scalars[7] = thr2;
    }
}

__global__ void __cluster_dims__(kDfCtas, 1, 1) __launch_bounds__(kDfThreads, 1) make_df_kernel(DfParams P) {
    make_df_body(P, ClusterSync());
}

__global__ void __launch_bounds__(kDfThreads, 1) make_df_kernel_grid(DfParams P) { make_df_body(P, GridSync()); }

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int64_t dfcsr_make_df_workspace(int32_t nx, int32_t nz) {
    if (nx < 1 || nz < 1) return 0;
    return (int64_t)sizeof(DfWorkspace) + (int64_t)6 * nx * nz * (int64_t)sizeof(double);
}

extern "C" int dfcsr_make_df(const double* d_count, const double* d_vxsum, dfcsr_axis x_axis, dfcsr_axis z_axis,
                             int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                             double velocity_threshold, double* d_fields, double* d_scalars, void* d_workspace,
                             void* stream) {
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
    DfWorkspace* ws = reinterpret_cast<DfWorkspace*>(d_workspace);
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
    P.ws = ws;
    P.scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(d_workspace) + sizeof(DfWorkspace));
    const long long cells = (long long)x_axis.n * z_axis.n;
    if (cells <= kClusterCells) {
        make_df_kernel<<<kDfCtas, kDfThreads, 0, st>>>(P);
    } else {
        long long want = (cells + 2047) / 2048;
        int ctas = (int)(want < kDfCtas ? kDfCtas : (want > kDfMaxCtas ? kDfMaxCtas : want));
        void* args[] = {&P};
        DFCSR_CUDA_OK(cudaLaunchCooperativeKernel((const void*)make_df_kernel_grid, dim3(ctas), dim3(kDfThreads), args, 0, st));
    }
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
