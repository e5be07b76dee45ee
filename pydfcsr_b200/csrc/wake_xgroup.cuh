// wake_xgroup.cuh — K4, second mapping: one warp lane per OBSERVATION point, 32 observation points that share s.
// Included by wake.cu (uses its device helpers); same reference lines: CSR.py:397-451, 454-602, 605-782,
// interp3D.py:18-66, interp1D.py:13-36.
//
// Why a second mapping.  Without chirp band (|slope| <= 1, CSR.py:480) the quadrature nodes of an observation point
// (s, x) depend on s only: x' nodes are centred on x0 = (s - t) tan(theta) and s' nodes on s (CSR.py:482-520).  The
// mesh get_CSR_mesh builds is a tensor grid (CSR.py:380-389), so all observation points of one z row share their
// (x', s') nodes, and tools/k4_sample_stats.py shows (bench workload) that the SAME sample of neighbouring observation
// points falls into the same (t', z) history cell in 98 % (far rectangles) / 71 % (near rectangle) of the cases, and
// that along s' a sample's cell moves by 0.001 (middle rectangle) to 0.09 (far rectangle) z cells per node.  The
// point-per-CTA kernel (lane = s' node) cannot use either fact: every lane fetches 8 voxels x 40 B for every sample,
// and the L1 -> register path (320 B per lane-sample) is its binding unit (profiles/k4_r2_bench.txt).
//
// Mapping here: a GROUP = up to 32 mesh points with the same z index and consecutive x indices; lane = point.  A warp
// walks one x' node through the s' nodes IN SEQUENCE.  Per step all lanes share the node record (broadcast reads), the two
// history rows and the transverse fraction, and they look at nearly the same place of the (t', z) plane: the warp keeps
// the transverse-blended history nodes of a 2 x 16 window around that place in shared memory (one node per lane when the
// window is refilled) and a sample reads its four corners from there -- 12 shared-memory loads and a 4-corner blend
// instead of 24 global loads and an 8-corner blend.  The window follows the sweep (a refill every ~25-80 steps); a lane
// that is outside it (near the observer r depends strongly on x: 6 % of the warp-steps on the bench workload) blends its
// corners from global memory with the same operations, so the bits do not depend on the path.  The sweep is software-
// pipelined: the dependent geometry chain of node j + 1 is issued next to the independent integrand algebra of node j.
// (First version: four corners per lane cached in registers, reloaded when the lane changed cell -- that divergent path
// ran in 53 % of the steps with 7 of 32 lanes active; profiles/k4_r2_xgroup_variants.txt has all the steps.)
//
// Work split and summation order (bitwise independent of the launch geometry and of the rank split):
//   * the pruned x' nodes of all rectangles form one list; UNIT u = nodes [u U, (u+1) U) of it; a warp accumulates a
//     unit per lane in node order, s' order, and stores the unit's 2 x 32 sums to the group's partial table;
//   * the warps of all CTAs that serve a group (any number) draw units from the group's queue (a global counter), most
//     expensive rectangle first; which warp computes a unit does not matter, its sums land in the unit's slot;
//   * every finished unit is counted; the warp that finishes the group's LAST unit adds the unit sums in unit order and
//     writes the 32 points (to the local arrays and/or, over NVLink, to every rank's grid).  No CTA-wide barrier after
//     set-up: a warp that finds the queue empty just exits.
// U depends on the global mesh and the integration parameters only (xgroup_plan), never on the split.
#pragma once
// (included INSIDE namespace dfcsr of wake.cu, after the shared device helpers)

constexpr int kXRec = 12;          // base_x, base_y, n'x, n'y, tau'x, tau'y, kappa, s', w_s, n-n' (2), n.tau': 96 B, six LDS.128
constexpr int kXWin = 16;          // z nodes of a warp's window of transverse-blended history nodes (x 2 slices = 32 = one per lane)
constexpr int kXThreads = 256;
constexpr int kXWarps = kXThreads / 32;

struct XGroupShared {
    Region reg[kMaxRegions];
    double X0, Y0, nx, ny, tx, ty;   // orbit, normal, tangent at s (shared by the group)
    double s, x_mid, x_half;         // group's s; centre and half width of its x values (s' bracket)
    int nreg;
    int jlo[kMaxRegions], jhi[kMaxRegions];
    int node_base[kMaxRegions + 1];  // prefix of pruned x' nodes per rectangle
    int skip;                        // 1: the group's queue was already empty when this CTA arrived
};

struct XGroupArgs {
    long long group_first, group_stride;   // group k of the launch is global group group_first + k * group_stride
    int ngroups;                           // groups of this launch (the grid is ngroups x CTAs-per-group)
    int unit_nodes;                        // U
    int max_units;                         // rows of a group's partial table
    double* partials;                      // [groups of the launch][max_units][64]
    unsigned int* tickets;                 // [groups of the launch][2] = {units handed out, units finished}; zeroed by the launcher
};

// the s'-only constants of every node of every rectangle as 96-byte records with base = R0(s) - R0(s') instead of the
// point's C = base + x n(s) (CSR.py:645: the lanes add their own x n(s)), and the bracket of the s' nodes that can reach
// the history grid for ANY x' of the pruned range and ANY x of the group: r(x, x') = |base + x n - x' n'| differs from
// r(x_mid, x') by at most |x - x_mid| |n| (triangle inequality), so the point-kernel bracket at x_mid, widened by the
// group's half width, is conservative.  The exact per-sample test stays in the sweep.
__device__ __forceinline__ void fill_node_records_x(const HistDev& H, const LatDev& L, const XGroupShared& sh, double t,
                                                    int nz, int nzp, double* tab, int* jlo, int* jhi) {
    const double hn = sh.x_half * sqrt(sh.nx * sh.nx + sh.ny * sh.ny) * (1.0 + 1e-9);
    if (threadIdx.x < kXRec) tab[(size_t)sh.nreg * nzp * kXRec + threadIdx.x] = 0.0;   // pad record: the pipelined sweep reads one node ahead
    for (int n = threadIdx.x; n < sh.nreg * nzp; n += (int)blockDim.x) {
        const int r = n / nzp, jj = n - r * nzp;
        const Axis sa = sh.reg[r].sa;
        const double sp = axis_node(sa, jj);                      // clamps past the last node
        const double sp_prev = (jj > 0) ? axis_node(sa, jj - 1) : sp;
        const double sp_next = axis_node(sa, jj + 1);
        double v[6];
        lattice_at(L, sp, v);
        const double bx = sub_rn(sh.X0, v[0]), by = sub_rn(sh.Y0, v[1]);
        double* o = tab + (size_t)n * kXRec;
        o[0] = bx; o[1] = by; o[2] = v[2]; o[3] = v[3]; o[4] = v[4]; o[5] = v[5];
        o[6] = curvature_at(L, sp); o[7] = sp;
        o[8] = (jj < nz) ? 0.5 * ((sp_next - sp) + (sp - sp_prev)) : 0.0;
        o[9] = sh.nx - v[2];                                      // n(s) - n(s') and n(s) . tau(s') (CSR.py:757-760)
        o[10] = sh.ny - v[3];
        o[11] = add_rn(mul_rn(sh.nx, v[4]), mul_rn(sh.ny, v[5]));
        if (jj < nz && sh.reg[r].ilo <= sh.reg[r].ihi) {
            const double Cx = bx + sh.x_mid * sh.nx, Cy = by + sh.x_mid * sh.ny;
            const double xa = axis_node(sh.reg[r].xa, sh.reg[r].ilo), xb = axis_node(sh.reg[r].xa, sh.reg[r].ihi);
            const double ax = Cx - xa * v[2], ay = Cy - xa * v[3];
            const double cx = Cx - xb * v[2], cy = Cy - xb * v[3];
            const double ra = sqrt(ax * ax + ay * ay), rb = sqrt(cx * cx + cy * cy);
            double r_max = fmax(ra, rb), r_min = fmin(ra, rb);
            const double nn = v[2] * v[2] + v[3] * v[3];
            const double xs = (Cx * v[2] + Cy * v[3]) / nn;
            if (!(xs <= fmin(xa, xb)) && !(xs >= fmax(xa, xb))) {     // foot point inside (or undecidable)
                const double fx = Cx - xs * v[2], fy = Cy - xs * v[3];
                r_min = fmin(r_min, sqrt(fx * fx + fy * fy));
            }
            r_max = r_max * (1.0 + 1e-12) + hn;
            r_min = fmax(0.0, r_min * (1.0 - 1e-12) - hn);
            const double m = 1e-3;
            const double ut_lo = ((t - r_max) - H.min_t) * H.inv_dt, ut_hi = ((t - r_min) - H.min_t) * H.inv_dt;
            const double uz_lo = ((sp - (t - r_min)) - H.min_z) * H.inv_dz, uz_hi = ((sp - (t - r_max)) - H.min_z) * H.inv_dz;
            const bool outside = (fmax(ut_lo, ut_hi) <= -1.0 - m) || (fmin(ut_lo, ut_hi) >= (double)H.T + m) ||
                                 (fmax(uz_lo, uz_hi) <= -1.0 - m) || (fmin(uz_lo, uz_hi) >= (double)H.Z + m);
            const bool certain = (ut_lo == ut_lo) && (ut_hi == ut_hi) && (uz_lo == uz_lo) && (uz_hi == uz_hi);   // no NaN
            if (!(outside && certain)) {
                atomicMin(jlo + r, jj);
                atomicMax(jhi + r, jj);
            }
        }
    }
}

// shared-memory accesses by 32-bit shared-space address (volatile: kept in program order with each other)
__device__ __forceinline__ void lds128(unsigned a, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds64(unsigned a, double& x) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a)); }
__device__ __forceinline__ void sts128(unsigned a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void sts64(unsigned a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }

// one history voxel of both transverse rows, blended along the transverse axis (same operations as yblend_zrun)
template <bool kF32>
__device__ __forceinline__ void yblend_node(const char* __restrict__ pa, const char* __restrict__ pb, double wy0, double yd,
                                            double (&out)[5]) {
    if (kF32) {
        const float w0 = (float)wy0, w1 = (float)yd;
        const float4 a = __ldg(reinterpret_cast<const float4*>(pa)), b = __ldg(reinterpret_cast<const float4*>(pb));
        const float ca = __ldg(reinterpret_cast<const float*>(pa) + 4), cb = __ldg(reinterpret_cast<const float*>(pb) + 4);
        out[0] = (double)fmaf(w1, b.x, w0 * a.x); out[1] = (double)fmaf(w1, b.y, w0 * a.y);
        out[2] = (double)fmaf(w1, b.z, w0 * a.z); out[3] = (double)fmaf(w1, b.w, w0 * a.w);
        out[4] = (double)fmaf(w1, cb, w0 * ca);
    } else {
        const double2* qa = reinterpret_cast<const double2*>(pa);
        const double2* qb = reinterpret_cast<const double2*>(pb);
        const double2 a0 = __ldg(qa), a1 = __ldg(qa + 1), b0 = __ldg(qb), b1 = __ldg(qb + 1);
        const double ca = __ldg(reinterpret_cast<const double*>(pa) + 4), cb = __ldg(reinterpret_cast<const double*>(pb) + 4);
        out[0] = fma(yd, b0.x, wy0 * a0.x); out[1] = fma(yd, b0.y, wy0 * a0.y);
        out[2] = fma(yd, b1.x, wy0 * a1.x); out[3] = fma(yd, b1.y, wy0 * a1.y);
        out[4] = fma(yd, cb, wy0 * ca);
    }
}

// Moves the warp's 2 x kXWin window of transverse-blended history nodes so that it holds the cells (t0, z0) of all lanes
// with `ok` -- if they fit one window (same slice pair, z spread <= kXWin - 2) -- and refills it: lane -> (slice tw or its
// successor, node zw + 0..15), clamped into the grid (clamped copies are never read: a valid cell ends at node Z - 1).
// Returns (tw << 32) | zw, or -1 when the lanes do not fit (they then gather from global memory).  Warp-uniform.
template <bool kF32>
__device__ __noinline__ long long xg_move_window(int HT, int HZ, int Hhead, int Hcap, bool ok, int t0, int z0, const char* row0,
                                                 const char* row1, double wy0, double yd, unsigned tile_s, unsigned slice_bytes) {
    constexpr int VB = kF32 ? DFCSR_VOXEL_FLOATS * 4 : DFCSR_VOXEL_DOUBLES * 8;
    const int lane = threadIdx.x & 31;
    const int zmin = __reduce_min_sync(0xffffffffu, ok ? z0 : INT_MAX);
    const int zmax = __reduce_max_sync(0xffffffffu, ok ? z0 : INT_MIN);
    const int tmin = __reduce_min_sync(0xffffffffu, ok ? t0 : INT_MAX);
    const int tmax = __reduce_max_sync(0xffffffffu, ok ? t0 : INT_MIN);
    if (tmin != tmax || zmax - zmin > kXWin - 2) return -1;
    const int tw = tmin;
    const int zw = max(0, min(zmin - ((kXWin - 2 - (zmax - zmin)) >> 1), HZ - kXWin));
    const int ti = lane >> 4, zn = min(zw + (lane & (kXWin - 1)), HZ - 1);
    int sl = Hhead + ((ti && tw != HT - 1) ? tw + 1 : tw);
    sl -= (sl >= Hcap) ? Hcap : 0;
    const size_t o = (size_t)((unsigned long long)(unsigned)sl * slice_bytes + (unsigned)zn * (unsigned)VB);
    double y[5];
    yblend_node<kF32>(row0 + o, row1 + o, wy0, yd, y);
    __syncwarp();
    sts128(tile_s + lane * 48, y[0], y[1]);
    sts128(tile_s + lane * 48 + 16, y[2], y[3]);
    sts64(tile_s + lane * 48 + 32, y[4]);
    __syncwarp();
    return ((long long)tw << 32) | (long long)(unsigned)zw;
}

// Same for lanes that carry several cells: [tlo, thi] x [zlo, zhi] per lane.
template <bool kF32>
__device__ __noinline__ long long xg_move_window_range(int HT, int HZ, int Hhead, int Hcap, bool ok, int tlo, int thi, int zlo,
                                                       int zhi, const char* row0, const char* row1, double wy0, double yd,
                                                       unsigned tile_s, unsigned slice_bytes) {
    constexpr int VB = kF32 ? DFCSR_VOXEL_FLOATS * 4 : DFCSR_VOXEL_DOUBLES * 8;
    const int lane = threadIdx.x & 31;
    const int zmin = __reduce_min_sync(0xffffffffu, ok ? zlo : INT_MAX);
    const int zmax = __reduce_max_sync(0xffffffffu, ok ? zhi : INT_MIN);
    const int tmin = __reduce_min_sync(0xffffffffu, ok ? tlo : INT_MAX);
    const int tmax = __reduce_max_sync(0xffffffffu, ok ? thi : INT_MIN);
    if (tmin != tmax || zmax - zmin > kXWin - 2) return -1;
    const int tw = tmin;
    const int zw = max(0, min(zmin - ((kXWin - 2 - (zmax - zmin)) >> 1), HZ - kXWin));
    const int ti = lane >> 4, zn = min(zw + (lane & (kXWin - 1)), HZ - 1);
    int sl = Hhead + ((ti && tw != HT - 1) ? tw + 1 : tw);
    sl -= (sl >= Hcap) ? Hcap : 0;
    const size_t o = (size_t)((unsigned long long)(unsigned)sl * slice_bytes + (unsigned)zn * (unsigned)VB);
    double y[5];
    yblend_node<kF32>(row0 + o, row1 + o, wy0, yd, y);
    __syncwarp();
    sts128(tile_s + lane * 48, y[0], y[1]);
    sts128(tile_s + lane * 48 + 16, y[2], y[3]);
    sts64(tile_s + lane * 48 + 32, y[4]);
    __syncwarp();
    return ((long long)tw << 32) | (long long)(unsigned)zw;
}

// the two results of the group's points (CSR.py:588-589), x-major flattening of the mesh (CSR.py:382-389)
__device__ __forceinline__ void xgroup_store(const MeshSrc& M, const dfcsr_wake_params& wp, const PeerOut& peers,
                                             double* out_dE, double* out_kick, int ix, int iz, bool lane_valid,
                                             double z, double xk) {
    const double v_dE = -wp.csr_scaling * z;
    const double v_kick = wp.csr_scaling * xk;
    const long long idx = (long long)ix * M.mz.n + iz;
    if (!lane_valid) return;
    if (out_dE) out_dE[idx] = v_dE;
    if (out_kick) out_kick[idx] = v_kick;
#pragma unroll
    for (int p = 0; p < DFCSR_MAX_PEERS; ++p) {
        if (p < peers.n) {
            peers.grid[p][idx] = v_dE;
            peers.grid[p][peers.n_total + idx] = v_kick;
        }
    }
}

template <bool kF32, bool kPipe, int kT = kXThreads, int kB = 2, int kUnroll = 1>
__global__ void __launch_bounds__(kT, kB)
wake_xgroup_kernel(HistDev H, LatDev L, dfcsr_wake_params wp, MeshSrc M, XGroupArgs A, double* __restrict__ out_dE,
                   double* __restrict__ out_kick, unsigned long long* counters, const PeerOut peers) {
    constexpr int VB = kF32 ? DFCSR_VOXEL_FLOATS * 4 : DFCSR_VOXEL_DOUBLES * 8;   // bytes per voxel
    __shared__ XGroupShared sh;
    // dynamic shared memory: one 2 x kXWin window of 48-byte transverse-blended history nodes per warp, then the node
    // table [nreg * nzp + 1][kXRec] (16-byte aligned records)
    extern __shared__ double2 xg_smem[];
    double2* const node_tab2 = xg_smem + (kT / 32) * (2 * kXWin * 3);
    double* const node_tab = reinterpret_cast<double*>(node_tab2);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // CTAs are numbered chunk-major: the first wave spreads over all groups, and a CTA that starts later joins the
    // group after the one its predecessor joined
    const int gl = (int)(blockIdx.x % (unsigned)A.ngroups);           // group of this launch
    const long long g = A.group_first + (long long)gl * A.group_stride;
    const int ngx = (M.mx.n + 31) >> 5;
    const int iz = (int)(g / ngx), gx = (int)(g - (long long)iz * ngx);
    const int ix = gx * 32 + lane;
    const bool lane_valid = ix < M.mx.n;
    const int nz = wp.nz;
    const int nzp = (nz + 31) & ~31;
    // the warp's window as a 32-bit shared-space address, accessed through ld/st.shared below (a generic pointer would be
    // re-derived from the thread and CTA ids at every step once registers are short)
    unsigned tile_s = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double*>(xg_smem) + warp * (2 * kXWin * 6));
    asm volatile("mov.u32 %0, %0;" : "+r"(tile_s));        // opaque: keep it in a register instead of recomputing it

    // ---- set-up 1: the group's s, its rectangles and the list of pruned x' nodes ---------------------
    const double zz = axis_node(M.mz, iz);
    const double shift = __dadd_rn(__dmul_rn(M.slope, zz), M.intercept);
    if (threadIdx.x == 0) {
        const double s = wp.t + zz;                       // CSR.py:412
        const int ix_hi = min(gx * 32 + 31, M.mx.n - 1);
        const double xa = __dadd_rn(axis_node(M.mx, gx * 32), shift), xb = __dadd_rn(axis_node(M.mx, ix_hi), shift);
        sh.s = s;
        sh.x_mid = 0.5 * (xa + xb);
        sh.x_half = 0.5 * fabs(xb - xa);
        int nreg;
        build_regions(wp, H, s, sh.x_mid, sh.reg, nreg);   // |slope0| <= 1 (launcher): the rectangles do not depend on x
        sh.nreg = nreg;
        int base = 0;
        for (int r = 0; r < nreg; ++r) {
            sh.node_base[r] = base;
            base += max(0, sh.reg[r].ihi - sh.reg[r].ilo + 1);
        }
        for (int r = nreg; r <= kMaxRegions; ++r) sh.node_base[r] = base;
        for (int r = 0; r < kMaxRegions; ++r) { sh.jlo[r] = INT_MAX; sh.jhi[r] = -1; }
        double v[6];
        lattice_at(L, s, v);
        sh.X0 = v[0]; sh.Y0 = v[1]; sh.nx = v[2]; sh.ny = v[3]; sh.tx = v[4]; sh.ty = v[5];
        // nothing left to do for this group?  (a late CTA of a group whose queue has run dry)
        const int total = base, nun = (total + A.unit_nodes - 1) / A.unit_nodes;
        sh.skip = (nun > 0 && *reinterpret_cast<volatile unsigned int*>(A.tickets + 2 * gl) >= (unsigned)nun) ? 1 : 0;
    }
    __syncthreads();
    if (sh.skip) return;
    const int nreg = sh.nreg;
    fill_node_records_x(H, L, sh, wp.t, nz, nzp, node_tab, sh.jlo, sh.jhi);

    // ---- per-lane constants of the lane's observation point (CSR.py:608-613, 645, 716-717) -----------
    const double x_obs = __dadd_rn(axis_node(M.mx, min(ix, M.mx.n - 1)), shift);
    const double Pt = wp.t, Pnx = sh.nx, Pny = sh.ny;
    double Pvx, Pvy;
    {
        double f[5];
        const double ut = (wp.t - H.min_t) * H.inv_dt, uy = (x_obs - H.min_x) * H.inv_dx;
        const double uz = ((sh.s - wp.t) - H.min_z) * H.inv_dz;
        const double vx = gather5<kF32>(H, ut, uy, uz, f) ? f[3] : 0.0;
        Pvx = add_rn(sh.tx, mul_rn(vx, Pnx));              // vs*tau + vx*n with vs = 1
        Pvy = add_rn(sh.ty, mul_rn(vx, Pny));
    }
    const double xnx = mul_rn(x_obs, Pnx), xny = mul_rn(x_obs, Pny);
    __syncthreads();

    const char* const ring = reinterpret_cast<const char*>(H.ring);
    const unsigned slice_bytes = (unsigned)H.slice_elems * (kF32 ? 4u : 8u);   // < 2^32, checked by the launcher
    const unsigned row_bytes = (unsigned)H.Z * (unsigned)VB;
    const int total_nodes = sh.node_base[nreg];
    const int U = A.unit_nodes;
    const int nunits = (total_nodes + U - 1) / U;
    double* const gpart = A.partials + (size_t)gl * A.max_units * 64;
    unsigned long long n_in = 0;
    const bool counting = counters != nullptr;      // sample accounting is optional (bench / tests)
    const int Zm1 = H.Z - 1, Tm1 = H.T - 1;

    unsigned int* const q_next = A.tickets + 2 * gl;
    unsigned int* const q_done = A.tickets + 2 * gl + 1;
    if (nunits == 0) {                               // no x' node can reach the grid: the wakes of the group are zero
        if (blockIdx.x < (unsigned)A.ngroups && warp == 0) xgroup_store(M, wp, peers, out_dE, out_kick, ix, iz, lane_valid, 0.0, 0.0);
        return;
    }
    for (;;) {
        unsigned int qi = 0;
        if (lane == 0) qi = atomicAdd(q_next, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= (unsigned)nunits) break;
        const int u = nunits - 1 - (int)qi;          // the near rectangle (last in the list, a reload at every step) first
        double acc_z = 0.0, acc_x = 0.0;
        const int n_end = min(total_nodes, (u + 1) * U);
        for (int n = u * U; n < n_end; ++n) {
            int r = 0;
            while (r + 1 < nreg && n >= sh.node_base[r + 1]) ++r;
            const int i = sh.reg[r].ilo + (n - sh.node_base[r]);
            const Axis xa = sh.reg[r].xa;
            const double xv = axis_node(xa, i);
            const double uy = (xv - H.min_x) * H.inv_dx;
            if (!cell_valid(uy, H.X)) continue;               // warp-uniform
            int y0, y1;
            double yd;
            cell_split(uy, H.X, y0, y1, yd);
            const double wy0 = 1.0 - yd;
            const double x_prev = (i > 0) ? axis_node(xa, i - 1) : xv;
            const double x_next = axis_node(xa, i + 1);
            const double wx = 0.5 * ((x_next - xv) + (xv - x_prev));
            const char* const row0 = ring + (size_t)((unsigned)y0 * (unsigned long long)row_bytes);
            const char* const row1 = ring + (size_t)((unsigned)y1 * (unsigned long long)row_bytes);
            const int j_lo = sh.jlo[r], j_hi = sh.jhi[r];
            if (j_lo > j_hi) continue;                          // no s' node of this rectangle reaches the grid
            const double2* rec = node_tab2 + (size_t)(r * nzp + j_lo) * (kXRec / 2);
            // per-warp window of transverse-blended history nodes in shared memory: slices (tw, tw+1) x z nodes [zw, zw+16)
            int tw = INT_MIN, zw = 0;
            // 1 + x' kappa and its reciprocal change only where the curvature does (piecewise constant along s')
            double kprev = 0.0, scale = 1.0, rscale = 1.0;
            unsigned n_node = 0;

            // geometry of the sample at node record `q` (CSR.py:645-650): r - r', r^2 and, for ordinary exponents, r and
            // 1/r -- branch-free, so that the compiler may interleave this dependent chain with independent work
            auto geometry = [&](const double2* q, double& rx, double& ry, double& r2, double& rr, double& ir) -> bool {
                const double2 r01 = q[0], r23 = q[1];
                const double Cx = add_rn(r01.x, xnx), Cy = add_rn(r01.y, xny);   // (R0(s) - R0(s')) + x n(s)
                rx = sub_rn(Cx, mul_rn(xv, r23.x));                             // reference rounding order
                ry = sub_rn(Cy, mul_rn(xv, r23.y));
                r2 = add_rn(mul_rn(rx, rx), mul_rn(ry, ry));
                return sqrt_pair_fast(r2, rr, ir);
            };
            // fractional cell coordinates of the retarded point (CSR.py:648-650, interp3D.py:30-36)
            auto locate = [&](const double2* q, double rr, double& ut, double& uz) {
                const double t_ret = Pt - rr;
                ut = (t_ret - H.min_t) * H.inv_dt;
                uz = ((q[3].y - t_ret) - H.min_z) * H.inv_dz;
            };
            // exceptional exponents (r = 0, inf, NaN): the library's sqrt / rsqrt instead of the fused fast path
            auto fixup = [&](const double2* q, bool fast, double r2, double& ir, double& ut, double& uz) {
                if (!fast) {
                    ir = rsqrt(r2);
                    locate(q, __dsqrt_rn(r2), ut, uz);
                }
            };
            // The five fields at (ut, uz).  All lanes of a step look at nearly the same place of the (t', z) plane (same x'
            // and s' node, observation points 0.1 sigma_x apart), and the place moves slowly along the sweep: the warp keeps
            // the transverse-blended nodes of a 2 x 16 window around it in shared memory (one node per lane when the window
            // is refilled: 6 loads from the two history rows, 5 blends) and a sample is 4 corners x 40 B read from there.
            // When the lanes that need a new window fit into one (same slice pair, z spread <= 14 cells), the window is
            // moved (warp-uniform); a lane that is still outside (near the observer the spread exceeds the window) blends its
            // corners from global memory with the same operations, so the bits do not depend on the path taken.
            auto gather = [&](double ut, double uz, bool ok, double (&fld)[5]) {
                int t0 = ok ? __double2int_rz(ut) : tw;
                int z0 = ok ? __double2int_rz(uz) : zw;
                const double td = ut - (double)t0;
                double zd = uz - (double)z0;
                if (z0 == Zm1) { z0 = Zm1 - 1; zd = 1.0; }      // clamp cell: same voxel, weight exactly 1
                bool inwin = (t0 == tw) && ((unsigned)(z0 - zw) <= (unsigned)(kXWin - 2));
                if (__any_sync(0xffffffffu, ok && !inwin)) {
                    // rare (a window lasts ~25-80 steps): kept out of line so that none of it is hoisted into the step
                    const long long moved = xg_move_window<kF32>(H.T, H.Z, H.head, H.cap, ok, t0, z0, row0, row1, wy0, yd, tile_s, slice_bytes);
                    if (moved >= 0) {
                        tw = (int)(moved >> 32);
                        zw = (int)(moved & 0xffffffffll);
                        inwin = ok;
                    }
                }
                const double wt0 = 1.0 - td, wz0 = 1.0 - zd;
                const double w00 = wt0 * wz0, w01 = wt0 * zd, w10 = td * wz0, w11 = td * zd;
                double Y[4][5];
                if (inwin || !ok) {
                    const unsigned p = tile_s + (inwin ? (unsigned)(z0 - zw) * 48u : 0u);
                    lds128(p, Y[0][0], Y[0][1]);           lds128(p + 16, Y[0][2], Y[0][3]);           lds64(p + 32, Y[0][4]);
                    lds128(p + 48, Y[1][0], Y[1][1]);      lds128(p + 64, Y[1][2], Y[1][3]);           lds64(p + 80, Y[1][4]);
                    lds128(p + 768, Y[2][0], Y[2][1]);     lds128(p + 784, Y[2][2], Y[2][3]);          lds64(p + 800, Y[2][4]);
                    lds128(p + 816, Y[3][0], Y[3][1]);     lds128(p + 832, Y[3][2], Y[3][3]);          lds64(p + 848, Y[3][4]);
                } else {
                    int s0 = H.head + t0;
                    s0 -= (s0 >= H.cap) ? H.cap : 0;
                    int s1 = s0 + 1;
                    s1 = (s1 == H.cap) ? 0 : s1;
                    s1 = (t0 == Tm1) ? s0 : s1;
                    const unsigned zoff = (unsigned)z0 * (unsigned)VB;
                    const size_t o0 = (size_t)((unsigned long long)(unsigned)s0 * slice_bytes + zoff);
                    const size_t o1 = (size_t)((unsigned long long)(unsigned)s1 * slice_bytes + zoff);
                    yblend_zrun<kF32>(row0 + o0, row1 + o0, wy0, yd, Y[0], Y[1]);
                    yblend_zrun<kF32>(row0 + o1, row1 + o1, wy0, yd, Y[2], Y[3]);
                }
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    fld[q] = fma(w11, Y[3][q], fma(w10, Y[2][q], fma(w01, Y[1][q], w00 * Y[0][q])));
            };
            // integrand algebra (CSR.py:713-775), same operation order as integrand_algebra(), and the quadrature sums
            auto algebra = [&](const double2* q, const double (&fld)[5], double rx, double ry, double ir, bool ok) {
                const double2 r23 = q[1], r45 = q[2], r89 = q[4], rab = q[5];
                const double nxp = r23.x, nyp = r23.y, txp = r45.x, typ = r45.y, ws = r89.x;
                const double gz = div_by(fld[2], scale, rscale);   // rho_z / scale (exactly rho_z where kappa = 0)
                const double dnx = r89.y, dny = rab.x, q2 = rab.y;
                const double rho = fld[0], rho_x = fld[1], vxr = fld[3], vxx = fld[4];
                const double vrx = add_rn(txp, mul_rn(vxr, nxp));            // velocity_ret
                const double vry = add_rn(typ, mul_rn(vxr, nyp));
                const double gxx = add_rn(mul_rn(rho_x, nxp), mul_rn(gz, txp));   // nabla_density_ret
                const double gyy = add_rn(mul_rn(rho_x, nyp), mul_rn(gz, typ));
                const double dot = add_rn(mul_rn(Pvx, vrx), mul_rn(Pvy, vry));  // part1
                const double ax = mul_rn(sub_rn(Pvx, mul_rn(dot, vrx)), gxx);
                const double ay = mul_rn(sub_rn(Pvy, mul_rn(dot, vry)), gyy);
                const double num1 = mul_rn(scale, add_rn(ax, ay));
                const double num2 = mul_rn(mul_rn(mul_rn(-scale, dot), rho), vxx);
                const double Iz = add_rn(mul_rn(num1, ir), mul_rn(num2, ir));
                const double q1 = add_rn(mul_rn(rx, dnx), mul_rn(ry, dny));   // (r - r').(n - n')
                const double drho = sub_rn(-add_rn(mul_rn(vrx, gxx), mul_rn(vry, gyy)), mul_rn(rho, vxx));
                const double sq1 = mul_rn(scale, q1);
                const double ir2 = mul_rn(ir, ir);
                const double w1 = mul_rn(mul_rn(sq1, mul_rn(ir2, ir)), rho);
                const double w2 = mul_rn(mul_rn(sq1, ir2), drho);
                const double w3 = mul_rn(mul_rn(mul_rn(-scale, q2), ir), drho);
                const double Ix = add_rn(add_rn(w1, w2), w3);
                const double w = ws * wx;
                if (ok) {
                    acc_z = fma(w, Iz, acc_z);
                    acc_x = fma(w, Ix, acc_x);
                }
            };
            auto in_grid = [&](double ut, double uz) {               // interp3D.py:30-52
                return lane_valid && (ut > -1.0) && (ut < H.Td) && (uz > -1.0) && (uz < H.Zd);
            };
            auto new_scale = [&](double kappa) {                     // warp-uniform, a few times per x' node
                if (kappa != kprev) {
                    kprev = kappa;
                    scale = add_rn(1.0, mul_rn(xv, kappa));
                    rscale = rcp_newton(scale);
                }
            };

            if (kPipe) {
                // Software pipeline: a GPU warp issues in order, and the geometry of a sample is one long dependent fp64
                // chain (orbit difference -> r^2 -> rsqrt refinement -> r -> cell coordinates).  The geometry of node
                // j + 1 is therefore issued in the same straight-line block as the integrand algebra of node j, which is
                // independent of it; the gather of j + 1 (the only divergent part) follows.
                double rx, ry, r2, rr, ir, ut, uz, fld[5];
                bool fast = geometry(rec, rx, ry, r2, rr, ir);
                locate(rec, rr, ut, uz);
                fixup(rec, fast, r2, ir, ut, uz);
                bool ok = in_grid(ut, uz);
                if (counting) n_node += (unsigned)__popc(__ballot_sync(0xffffffffu, ok));
                new_scale(rec[3].x);
                gather(ut, uz, ok, fld);
#pragma unroll kUnroll
                for (int j = j_lo; j <= j_hi; ++j, rec += kXRec / 2) {
                    double rxn, ryn, irn;
                    // one straight-line block: the dependent chain of node j + 1 (the table is padded by one record)
                    // next to the independent algebra of node j
                    fast = geometry(rec + kXRec / 2, rxn, ryn, r2, rr, irn);
                    locate(rec + kXRec / 2, rr, ut, uz);
                    algebra(rec, fld, rx, ry, ir, ok);
                    fixup(rec + kXRec / 2, fast, r2, irn, ut, uz);
                    ok = in_grid(ut, uz) && (j < j_hi);
                    if (counting) n_node += (unsigned)__popc(__ballot_sync(0xffffffffu, ok));
                    new_scale(rec[kXRec / 2 + 3].x);
                    gather(ut, uz, ok, fld);
                    rx = rxn; ry = ryn; ir = irn;
                }
            } else {
                for (int j = j_lo; j <= j_hi; ++j, rec += kXRec / 2) {
                    double rx, ry, r2, rr, ir, ut, uz, fld[5];
                    const bool fast = geometry(rec, rx, ry, r2, rr, ir);
                    locate(rec, rr, ut, uz);
                    fixup(rec, fast, r2, ir, ut, uz);
                    const bool ok = in_grid(ut, uz);
                    const unsigned okm = __ballot_sync(0xffffffffu, ok);
                    if (okm == 0u) continue;
                    n_node += (unsigned)__popc(okm);
                    gather(ut, uz, ok, fld);
                    new_scale(rec[3].x);
                    algebra(rec, fld, rx, ry, ir, ok);
                }
            }
            n_in += n_node;
        }
        __stcg(gpart + (size_t)u * 64 + lane, acc_z);
        __stcg(gpart + (size_t)u * 64 + 32 + lane, acc_x);
        __threadfence();
        __syncwarp();
        unsigned int done = 0;
        if (lane == 0) done = atomicAdd(q_done, 1u);
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == (unsigned)nunits - 1u) {
            // this was the group's last unit: add the unit sums in unit order (loads first, adds in order)
            __threadfence();
            double z = 0.0, xk = 0.0;
            int v = 0;
            for (; v + 4 <= nunits; v += 4) {
                const double* p = gpart + (size_t)v * 64 + lane;
                const double a0 = __ldcg(p), a1 = __ldcg(p + 64), a2 = __ldcg(p + 128), a3 = __ldcg(p + 192);
                const double b0 = __ldcg(p + 32), b1 = __ldcg(p + 96), b2 = __ldcg(p + 160), b3 = __ldcg(p + 224);
                z += a0; z += a1; z += a2; z += a3;
                xk += b0; xk += b1; xk += b2; xk += b3;
            }
            for (; v < nunits; ++v) {
                z += __ldcg(gpart + (size_t)v * 64 + lane);
                xk += __ldcg(gpart + (size_t)v * 64 + 32 + lane);
            }
            xgroup_store(M, wp, peers, out_dE, out_kick, ix, iz, lane_valid, z, xk);
            if (lane == 0 && counters) {
                unsigned long long full = 0;
                for (int r = 0; r < nreg; ++r) full += (unsigned long long)sh.reg[r].xa.n * (unsigned long long)nz;
                atomicAdd(counters + 1, full * (unsigned long long)min(32, M.mx.n - gx * 32));
            }
        }
    }
    if (counters && lane == 0 && n_in) {
        atomicAdd(counters + 0, n_in);
        atomicAdd(counters + 2, n_in);
    }
}

#ifdef DFCSR_DEV_VARIANTS
// ---- MEASURED ALTERNATIVE, developer builds only: the same mapping with kP observation points per lane ----------------
// Result on the bench launch (B200): 1.50 ms (128 threads x 3 CTAs) / 1.59 ms (192 x 2) / 1.95 ms (256 x 1) against 1.47 ms
// for one point per lane at 16 warps per SM; bitwise the same grids.  Twice the instruction-level parallelism per warp at
// three quarters of the warps buys nothing: the kernel is not waiting on its own dependent chains.
// A group is 32 kP consecutive points of a mesh row; lane l carries points l, l + 32, ... of it.  The kP samples of a lane
// at a step share the node record, x' n', 1 + x' kappa, the quadrature weight, the warp's window and all control flow; their
// dependent fp64 chains are independent, so every warp offers the scheduler kP times the instruction-level parallelism (the
// single-point kernel leaves the fp64 pipe 37 % idle waiting on dependent results).  Every point's own arithmetic and
// summation order are exactly those of wake_xgroup_kernel: the two kernels give the same bits, and kP is a pure performance
// choice of the plan.  Workspace per group: max_units x (64 kP) doubles.
template <bool kF32, int kP, int kT, int kB>
__global__ void __launch_bounds__(kT, kB)
wake_xgroup_kernel_mp(HistDev H, LatDev L, dfcsr_wake_params wp, MeshSrc M, XGroupArgs A, double* __restrict__ out_dE,
                      double* __restrict__ out_kick, unsigned long long* counters, const PeerOut peers) {
    constexpr int VB = kF32 ? DFCSR_VOXEL_FLOATS * 4 : DFCSR_VOXEL_DOUBLES * 8;   // bytes per voxel
    constexpr int GW = 32 * kP;                                                    // points per group
    __shared__ XGroupShared sh;
    extern __shared__ double2 xg_smem[];
    double2* const node_tab2 = xg_smem + (kT / 32) * (2 * kXWin * 3);
    double* const node_tab = reinterpret_cast<double*>(node_tab2);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int gl = (int)(blockIdx.x % (unsigned)A.ngroups);
    const long long g = A.group_first + (long long)gl * A.group_stride;
    const int ngx = (M.mx.n + GW - 1) / GW;
    const int iz = (int)(g / ngx), gx = (int)(g - (long long)iz * ngx);
    int ix[kP];
    bool lane_valid[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) { ix[p] = gx * GW + p * 32 + lane; lane_valid[p] = ix[p] < M.mx.n; }
    const int nz = wp.nz;
    const int nzp = (nz + 31) & ~31;
    unsigned tile_s = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double*>(xg_smem) + warp * (2 * kXWin * 6));
    asm volatile("mov.u32 %0, %0;" : "+r"(tile_s));

    const double zz = axis_node(M.mz, iz);
    const double shift = __dadd_rn(__dmul_rn(M.slope, zz), M.intercept);
    if (threadIdx.x == 0) {
        const double s = wp.t + zz;                       // CSR.py:412
        const int ix_hi = min(gx * GW + GW - 1, M.mx.n - 1);
        const double xa = __dadd_rn(axis_node(M.mx, gx * GW), shift), xb = __dadd_rn(axis_node(M.mx, ix_hi), shift);
        sh.s = s;
        sh.x_mid = 0.5 * (xa + xb);
        sh.x_half = 0.5 * fabs(xb - xa);
        int nreg;
        build_regions(wp, H, s, sh.x_mid, sh.reg, nreg);
        sh.nreg = nreg;
        int base = 0;
        for (int r = 0; r < nreg; ++r) {
            sh.node_base[r] = base;
            base += max(0, sh.reg[r].ihi - sh.reg[r].ilo + 1);
        }
        for (int r = nreg; r <= kMaxRegions; ++r) sh.node_base[r] = base;
        for (int r = 0; r < kMaxRegions; ++r) { sh.jlo[r] = INT_MAX; sh.jhi[r] = -1; }
        double v[6];
        lattice_at(L, s, v);
        sh.X0 = v[0]; sh.Y0 = v[1]; sh.nx = v[2]; sh.ny = v[3]; sh.tx = v[4]; sh.ty = v[5];
        const int total = base, nun = (total + A.unit_nodes - 1) / A.unit_nodes;
        sh.skip = (nun > 0 && *reinterpret_cast<volatile unsigned int*>(A.tickets + 2 * gl) >= (unsigned)nun) ? 1 : 0;
    }
    __syncthreads();
    if (sh.skip) return;
    const int nreg = sh.nreg;
    fill_node_records_x(H, L, sh, wp.t, nz, nzp, node_tab, sh.jlo, sh.jhi);

    const double Pt = wp.t;
    double Pvx[kP], Pvy[kP], xnx[kP], xny[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) {
        const double x_obs = __dadd_rn(axis_node(M.mx, min(ix[p], M.mx.n - 1)), shift);
        double f[5];
        const double ut = (wp.t - H.min_t) * H.inv_dt, uy = (x_obs - H.min_x) * H.inv_dx;
        const double uz = ((sh.s - wp.t) - H.min_z) * H.inv_dz;
        const double vx = gather5<kF32>(H, ut, uy, uz, f) ? f[3] : 0.0;
        Pvx[p] = add_rn(sh.tx, mul_rn(vx, sh.nx));            // vs*tau + vx*n with vs = 1
        Pvy[p] = add_rn(sh.ty, mul_rn(vx, sh.ny));
        xnx[p] = mul_rn(x_obs, sh.nx);
        xny[p] = mul_rn(x_obs, sh.ny);
    }
    __syncthreads();

    const char* const ring = reinterpret_cast<const char*>(H.ring);
    const unsigned slice_bytes = (unsigned)H.slice_elems * (kF32 ? 4u : 8u);
    const unsigned row_bytes = (unsigned)H.Z * (unsigned)VB;
    const int total_nodes = sh.node_base[nreg];
    const int U = A.unit_nodes;
    const int nunits = (total_nodes + U - 1) / U;
    double* const gpart = A.partials + (size_t)gl * A.max_units * (64 * kP);
    unsigned long long n_in = 0;
    const bool counting = counters != nullptr;
    const int Zm1 = H.Z - 1, Tm1 = H.T - 1;
    unsigned int* const q_next = A.tickets + 2 * gl;
    unsigned int* const q_done = A.tickets + 2 * gl + 1;
    if (nunits == 0) {
        if (blockIdx.x < (unsigned)A.ngroups && warp == 0) {
#pragma unroll
            for (int p = 0; p < kP; ++p) xgroup_store(M, wp, peers, out_dE, out_kick, ix[p], iz, lane_valid[p], 0.0, 0.0);
        }
        return;
    }
    for (;;) {
        unsigned int qi = 0;
        if (lane == 0) qi = atomicAdd(q_next, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= (unsigned)nunits) break;
        const int u = nunits - 1 - (int)qi;
        double acc_z[kP], acc_x[kP];
#pragma unroll
        for (int p = 0; p < kP; ++p) { acc_z[p] = 0.0; acc_x[p] = 0.0; }
        const int n_end = min(total_nodes, (u + 1) * U);
        for (int n = u * U; n < n_end; ++n) {
            int r = 0;
            while (r + 1 < nreg && n >= sh.node_base[r + 1]) ++r;
            const int i = sh.reg[r].ilo + (n - sh.node_base[r]);
            const Axis xa = sh.reg[r].xa;
            const double xv = axis_node(xa, i);
            const double uy = (xv - H.min_x) * H.inv_dx;
            if (!cell_valid(uy, H.X)) continue;               // warp-uniform
            int y0, y1;
            double yd;
            cell_split(uy, H.X, y0, y1, yd);
            const double wy0 = 1.0 - yd;
            const double x_prev = (i > 0) ? axis_node(xa, i - 1) : xv;
            const double x_next = axis_node(xa, i + 1);
            const double wx = 0.5 * ((x_next - xv) + (xv - x_prev));
            const char* const row0 = ring + (size_t)((unsigned)y0 * (unsigned long long)row_bytes);
            const char* const row1 = ring + (size_t)((unsigned)y1 * (unsigned long long)row_bytes);
            const int j_lo = sh.jlo[r], j_hi = sh.jhi[r];
            if (j_lo > j_hi) continue;
            const double2* rec = node_tab2 + (size_t)(r * nzp + j_lo) * (kXRec / 2);
            int tw = INT_MIN, zw = 0;
            double kprev = 0.0, scale = 1.0, rscale = 1.0;
            unsigned n_node = 0;

            // geometry + cell coordinates of the kP samples at node record q (fast path of the fused sqrt, branch-free)
            auto geometry = [&](const double2* q, double (&rx)[kP], double (&ry)[kP], double (&r2)[kP], double (&ir)[kP],
                                double (&ut)[kP], double (&uz)[kP], bool (&fast)[kP]) {
                const double2 r01 = q[0], r23 = q[1];
                const double sp = q[3].y;
                const double xnp = mul_rn(xv, r23.x), ynp = mul_rn(xv, r23.y);        // x' n'(s'): shared by the lane's points
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const double Cx = add_rn(r01.x, xnx[p]), Cy = add_rn(r01.y, xny[p]);
                    rx[p] = sub_rn(Cx, xnp);
                    ry[p] = sub_rn(Cy, ynp);
                    r2[p] = add_rn(mul_rn(rx[p], rx[p]), mul_rn(ry[p], ry[p]));
                    double rr;
                    fast[p] = sqrt_pair_fast(r2[p], rr, ir[p]);
                    const double t_ret = Pt - rr;
                    ut[p] = (t_ret - H.min_t) * H.inv_dt;
                    uz[p] = ((sp - t_ret) - H.min_z) * H.inv_dz;
                }
            };
            auto fixup = [&](const double2* q, const bool (&fast)[kP], const double (&r2)[kP], double (&ir)[kP], double (&ut)[kP],
                             double (&uz)[kP]) {
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    if (!fast[p]) {                        // exceptional exponents (r = 0, inf, NaN): library path
                        ir[p] = rsqrt(r2[p]);
                        const double t_ret = Pt - __dsqrt_rn(r2[p]);
                        ut[p] = (t_ret - H.min_t) * H.inv_dt;
                        uz[p] = ((q[3].y - t_ret) - H.min_z) * H.inv_dz;
                    }
                }
            };
            auto gather = [&](const double (&ut)[kP], const double (&uz)[kP], const bool (&ok)[kP], double (&fld)[kP][5]) {
                int t0[kP], z0[kP];
                double td[kP], zd[kP];
                bool inwin[kP];
                bool need = false, any_ok = false;
                int tlo = INT_MAX, thi = INT_MIN, zlo = INT_MAX, zhi = INT_MIN;
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    t0[p] = ok[p] ? __double2int_rz(ut[p]) : tw;
                    z0[p] = ok[p] ? __double2int_rz(uz[p]) : zw;
                    td[p] = ut[p] - (double)t0[p];
                    zd[p] = uz[p] - (double)z0[p];
                    if (z0[p] == Zm1) { z0[p] = Zm1 - 1; zd[p] = 1.0; }
                    inwin[p] = (t0[p] == tw) && ((unsigned)(z0[p] - zw) <= (unsigned)(kXWin - 2));
                    need = need || (ok[p] && !inwin[p]);
                }
                if (__any_sync(0xffffffffu, need)) {
#pragma unroll
                    for (int p = 0; p < kP; ++p) {
                        if (ok[p]) {
                            any_ok = true;
                            tlo = min(tlo, t0[p]); thi = max(thi, t0[p]); zlo = min(zlo, z0[p]); zhi = max(zhi, z0[p]);
                        }
                    }
                    const long long moved = xg_move_window_range<kF32>(H.T, H.Z, H.head, H.cap, any_ok, tlo, thi, zlo, zhi, row0, row1,
                                                                       wy0, yd, tile_s, slice_bytes);
                    if (moved >= 0) {
                        tw = (int)(moved >> 32);
                        zw = (int)(moved & 0xffffffffll);
#pragma unroll
                        for (int p = 0; p < kP; ++p) inwin[p] = ok[p];
                    }
                }
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const double wt0 = 1.0 - td[p], wz0 = 1.0 - zd[p];
                    const double w00 = wt0 * wz0, w01 = wt0 * zd[p], w10 = td[p] * wz0, w11 = td[p] * zd[p];
                    double Y[4][5];
                    if (inwin[p] || !ok[p]) {
                        const unsigned a = tile_s + (inwin[p] ? (unsigned)(z0[p] - zw) * 48u : 0u);
                        lds128(a, Y[0][0], Y[0][1]);           lds128(a + 16, Y[0][2], Y[0][3]);           lds64(a + 32, Y[0][4]);
                        lds128(a + 48, Y[1][0], Y[1][1]);      lds128(a + 64, Y[1][2], Y[1][3]);           lds64(a + 80, Y[1][4]);
                        lds128(a + 768, Y[2][0], Y[2][1]);     lds128(a + 784, Y[2][2], Y[2][3]);          lds64(a + 800, Y[2][4]);
                        lds128(a + 816, Y[3][0], Y[3][1]);     lds128(a + 832, Y[3][2], Y[3][3]);          lds64(a + 848, Y[3][4]);
                    } else {
                        int s0 = H.head + t0[p];
                        s0 -= (s0 >= H.cap) ? H.cap : 0;
                        int s1 = s0 + 1;
                        s1 = (s1 == H.cap) ? 0 : s1;
                        s1 = (t0[p] == Tm1) ? s0 : s1;
                        const unsigned zoff = (unsigned)z0[p] * (unsigned)VB;
                        const size_t o0 = (size_t)((unsigned long long)(unsigned)s0 * slice_bytes + zoff);
                        const size_t o1 = (size_t)((unsigned long long)(unsigned)s1 * slice_bytes + zoff);
                        yblend_zrun<kF32>(row0 + o0, row1 + o0, wy0, yd, Y[0], Y[1]);
                        yblend_zrun<kF32>(row0 + o1, row1 + o1, wy0, yd, Y[2], Y[3]);
                    }
#pragma unroll
                    for (int q = 0; q < 5; ++q)
                        fld[p][q] = fma(w11, Y[3][q], fma(w10, Y[2][q], fma(w01, Y[1][q], w00 * Y[0][q])));
                }
            };
            auto algebra = [&](const double2* q, const double (&fld)[kP][5], const double (&rx)[kP], const double (&ry)[kP],
                               const double (&ir)[kP], const bool (&ok)[kP]) {
                const double2 r23 = q[1], r45 = q[2], r89 = q[4], rab = q[5];
                const double nxp = r23.x, nyp = r23.y, txp = r45.x, typ = r45.y, ws = r89.x;
                const double dnx = r89.y, dny = rab.x, q2 = rab.y;
                const double w = ws * wx;
                const double msq2 = mul_rn(-scale, q2);
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    const double gz = div_by(fld[p][2], scale, rscale);
                    const double rho = fld[p][0], rho_x = fld[p][1], vxr = fld[p][3], vxx = fld[p][4];
                    const double vrx = add_rn(txp, mul_rn(vxr, nxp));
                    const double vry = add_rn(typ, mul_rn(vxr, nyp));
                    const double gxx = add_rn(mul_rn(rho_x, nxp), mul_rn(gz, txp));
                    const double gyy = add_rn(mul_rn(rho_x, nyp), mul_rn(gz, typ));
                    const double dot = add_rn(mul_rn(Pvx[p], vrx), mul_rn(Pvy[p], vry));
                    const double ax = mul_rn(sub_rn(Pvx[p], mul_rn(dot, vrx)), gxx);
                    const double ay = mul_rn(sub_rn(Pvy[p], mul_rn(dot, vry)), gyy);
                    const double num1 = mul_rn(scale, add_rn(ax, ay));
                    const double num2 = mul_rn(mul_rn(mul_rn(-scale, dot), rho), vxx);
                    const double Iz = add_rn(mul_rn(num1, ir[p]), mul_rn(num2, ir[p]));
                    const double q1 = add_rn(mul_rn(rx[p], dnx), mul_rn(ry[p], dny));
                    const double drho = sub_rn(-add_rn(mul_rn(vrx, gxx), mul_rn(vry, gyy)), mul_rn(rho, vxx));
                    const double sq1 = mul_rn(scale, q1);
                    const double ir2 = mul_rn(ir[p], ir[p]);
                    const double w1 = mul_rn(mul_rn(sq1, mul_rn(ir2, ir[p])), rho);
                    const double w2 = mul_rn(mul_rn(sq1, ir2), drho);
                    const double w3 = mul_rn(mul_rn(msq2, ir[p]), drho);
                    const double Ix = add_rn(add_rn(w1, w2), w3);
                    if (ok[p]) {
                        acc_z[p] = fma(w, Iz, acc_z[p]);
                        acc_x[p] = fma(w, Ix, acc_x[p]);
                    }
                }
            };
            auto new_scale = [&](double kappa) {
                if (kappa != kprev) {
                    kprev = kappa;
                    scale = add_rn(1.0, mul_rn(xv, kappa));
                    rscale = rcp_newton(scale);
                }
            };
            auto in_grid = [&](const double (&ut)[kP], const double (&uz)[kP], bool live, bool (&ok)[kP]) {
                bool any = false;
#pragma unroll
                for (int p = 0; p < kP; ++p) {
                    ok[p] = live && lane_valid[p] && (ut[p] > -1.0) && (ut[p] < H.Td) && (uz[p] > -1.0) && (uz[p] < H.Zd);
                    any = any || ok[p];
                    if (counting) n_node += (unsigned)__popc(__ballot_sync(0xffffffffu, ok[p]));
                }
                return any;
            };

            double rx[kP], ry[kP], r2[kP], ir[kP], ut[kP], uz[kP], fld[kP][5];
            bool fast[kP], ok[kP];
            geometry(rec, rx, ry, r2, ir, ut, uz, fast);
            fixup(rec, fast, r2, ir, ut, uz);
            in_grid(ut, uz, true, ok);
            new_scale(rec[3].x);
            gather(ut, uz, ok, fld);
            for (int j = j_lo; j <= j_hi; ++j, rec += kXRec / 2) {
                double rxn[kP], ryn[kP], irn[kP];
                geometry(rec + kXRec / 2, rxn, ryn, r2, irn, ut, uz, fast);        // node j + 1 (the table is padded by one record)
                algebra(rec, fld, rx, ry, ir, ok);                                  // node j
                fixup(rec + kXRec / 2, fast, r2, irn, ut, uz);
                in_grid(ut, uz, j < j_hi, ok);
                new_scale(rec[kXRec / 2 + 3].x);
                gather(ut, uz, ok, fld);
#pragma unroll
                for (int p = 0; p < kP; ++p) { rx[p] = rxn[p]; ry[p] = ryn[p]; ir[p] = irn[p]; }
            }
            n_in += n_node;
        }
        double* const row = gpart + (size_t)u * (64 * kP);
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            __stcg(row + p * 64 + lane, acc_z[p]);
            __stcg(row + p * 64 + 32 + lane, acc_x[p]);
        }
        __threadfence();
        __syncwarp();
        unsigned int done = 0;
        if (lane == 0) done = atomicAdd(q_done, 1u);
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == (unsigned)nunits - 1u) {
            __threadfence();
#pragma unroll 1
            for (int p = 0; p < kP; ++p) {
                double z = 0.0, xk = 0.0;
                int v = 0;
                const double* base = gpart + p * 64 + lane;
                for (; v + 4 <= nunits; v += 4) {
                    const double* q = base + (size_t)v * (64 * kP);
                    const double a0 = __ldcg(q), a1 = __ldcg(q + 64 * kP), a2 = __ldcg(q + 128 * kP), a3 = __ldcg(q + 192 * kP);
                    const double b0 = __ldcg(q + 32), b1 = __ldcg(q + 64 * kP + 32), b2 = __ldcg(q + 128 * kP + 32), b3 = __ldcg(q + 192 * kP + 32);
                    z += a0; z += a1; z += a2; z += a3;
                    xk += b0; xk += b1; xk += b2; xk += b3;
                }
                for (; v < nunits; ++v) {
                    z += __ldcg(base + (size_t)v * (64 * kP));
                    xk += __ldcg(base + (size_t)v * (64 * kP) + 32);
                }
                xgroup_store(M, wp, peers, out_dE, out_kick, gx * GW + p * 32 + lane, iz, gx * GW + p * 32 + lane < M.mx.n, z, xk);
            }
            if (lane == 0 && counters) {
                unsigned long long full = 0;
                for (int r = 0; r < nreg; ++r) full += (unsigned long long)sh.reg[r].xa.n * (unsigned long long)nz;
                atomicAdd(counters + 1, full * (unsigned long long)min(GW, M.mx.n - gx * GW));
            }
        }
    }
    if (counters && lane == 0 && n_in) {
        atomicAdd(counters + 0, n_in);
        atomicAdd(counters + 2, n_in);
    }
}
#endif  // DFCSR_DEV_VARIANTS
