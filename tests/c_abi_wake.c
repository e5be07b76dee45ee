/* Plain-C caller of the WAKE half of libdfcsr_b200.so (K3 -> K4 -> K5): header + cudart + device pointers only.
 * Built and run by tests/test_gpu_c_abi.py on the GPU box.
 *   K3  dfcsr_history_regrid        three analytic density-function records into a 4-slot ring (deposit.py:312-390)
 *   K4  dfcsr_wake_grid             one CTA per observation point                       (CSR.py:397-451, 454-782)
 *       dfcsr_wake_xgroup_plan + dfcsr_wake_grid_xgroups   one lane per point of a mesh row; whole mesh, then dealt out
 *                                    to three "ranks" through dfcsr's peer-grid interface: the SAME BITS are required
 *       dfcsr_wake_grid_peers       the point kernel through the same interface
 *   K5  dfcsr_apply_kick            the wake back onto particles (beams.py:108-131)
 * The lattice is a 0.3 m drift followed by a 0.7 m, 0.09 rad bend built like lattice.py:19-62 (arc about the
 * instantaneous centre); the bunch is a Gaussian with analytic gradients.  Checks: both K4 mappings agree to 1e-12 of the
 * mesh maximum, any split gives the single launch's bits, the kick of a particle at a mesh node is the node's wake. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dfcsr_b200.h"

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "FAIL %s: %s\n", #x, dfcsr_last_error()); return 1; } } while (0)
#define CU(x) do { if ((x) != cudaSuccess) { fprintf(stderr, "CUDA FAIL %s\n", #x); return 1; } } while (0)

static dfcsr_axis axis(double a, double b, int n) { dfcsr_axis x; x.start = a; x.stop = b; x.n = n; x._pad = 0; return x; }

int main(void) {
    if (dfcsr_abi_version() != DFCSR_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    const double sx = 1.0e-4, sz = 2.0e-4, t_now = 0.6;
    /* ---- lattice table: drift 0.3 m, then a bend (lattice.py:19-62) ---- */
    const int ns = 400, ne = 2;
    const double L0 = 0.3, L1 = 0.7, ang = 0.09;
    double* lat = (double*)malloc(sizeof(double) * ns * DFCSR_LATTICE_DOUBLES);
    double th = 0.0, X = 0.0, Y = 0.0;
    for (int k = 0; k < ns; ++k) {
        const double s = (L0 + L1) * k / (ns - 1), sp = (L0 + L1) * (k > 0 ? k - 1 : 0) / (ns - 1), ds = s - sp;
        if (k > 0) {
            if (s > L0) {
                const double phi = ds / L1 * ang, rad = L1 / ang, cx = X - rad * sin(th), cy = Y + rad * cos(th);
                X = cx + rad * sin(phi + th); Y = cy - rad * cos(phi + th); th += phi;
            } else { X += ds * cos(th); Y += ds * sin(th); }
        }
        double* r = lat + (size_t)k * DFCSR_LATTICE_DOUBLES;
        r[0] = X; r[1] = Y; r[2] = sin(th); r[3] = -cos(th); r[4] = cos(th); r[5] = sin(th);
    }
    const double rho[2] = {0.0, ang / L1}, dist[2] = {L0, L0 + L1};
    double *d_lat, *d_rho, *d_dist;
    CU(cudaMalloc((void**)&d_lat, sizeof(double) * ns * DFCSR_LATTICE_DOUBLES)); CU(cudaMalloc((void**)&d_rho, 16)); CU(cudaMalloc((void**)&d_dist, 16));
    CU(cudaMemcpy(d_lat, lat, sizeof(double) * ns * DFCSR_LATTICE_DOUBLES, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_rho, rho, 16, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_dist, dist, 16, cudaMemcpyHostToDevice));
    dfcsr_lattice lt; lt.d_table = d_lat; lt.ns = ns; lt.n_elements = ne; lt.min_s = 0.0; lt.delta_s = (L0 + L1) / (ns - 1);
    lt.d_rho = d_rho; lt.d_distance = d_dist;

    /* ---- K3: one analytic record (5 fields on a 60 x 70 source grid) re-gridded into three slots of a 4-slot ring ---- */
    const int SX = 60, SZ = 70, HX = 80, HZ = 90, CAP = 4, T = 3, HEAD = 2;
    double* f = (double*)malloc(sizeof(double) * 5 * SX * SZ);
    for (int i = 0; i < SX; ++i) for (int j = 0; j < SZ; ++j) {
        const double x = -5 * sx + 10 * sx * i / (SX - 1), z = -5 * sz + 10 * sz * j / (SZ - 1);
        const double g = exp(-0.5 * x * x / (sx * sx) - 0.5 * z * z / (sz * sz));
        f[(0 * SX + i) * SZ + j] = g; f[(1 * SX + i) * SZ + j] = -x / (sx * sx) * g; f[(2 * SX + i) * SZ + j] = -z / (sz * sz) * g;
        f[(3 * SX + i) * SZ + j] = 0.01 * x / sx; f[(4 * SX + i) * SZ + j] = 0.01 / sx;
    }
    double* d_f; void* d_ring; int32_t* d_sup;
    const int64_t slice = (int64_t)HX * HZ * DFCSR_VOXEL_DOUBLES;
    CU(cudaMalloc((void**)&d_f, sizeof(double) * 5 * SX * SZ)); CU(cudaMemcpy(d_f, f, sizeof(double) * 5 * SX * SZ, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_ring, sizeof(double) * slice * CAP)); CU(cudaMemset(d_ring, 0, sizeof(double) * slice * CAP));
    CU(cudaMalloc((void**)&d_sup, sizeof(int32_t) * CAP * HX * 2));
    const dfcsr_axis s_x = axis(-5 * sx, 5 * sx, SX), s_z = axis(-5 * sz, 5 * sz, SZ);
    const dfcsr_axis h_x = axis(-5.5 * sx, 5.5 * sx, HX), h_z = axis(-5.5 * sz, 5.5 * sz, HZ);
    for (int k = 0; k < T; ++k) {
        const int slot = (HEAD + k) % CAP;
        CHECK(dfcsr_history_regrid(d_f, s_x, s_z, h_x, h_z, 0.01 / sx, NULL, DFCSR_VOXEL_F64,
                                   (double*)d_ring + slice * slot, d_sup + (size_t)slot * HX * 2, NULL));
    }
    dfcsr_history h; memset(&h, 0, sizeof(h));
    h.d_ring = d_ring; h.slice_elems = slice; h.cap = CAP; h.head = HEAD; h.T = T; h.X = HX; h.Z = HZ; h.format = DFCSR_VOXEL_F64;
    h.min_t = 0.4; h.delta_t = 0.1; h.min_x = h_x.start; h.delta_x = (h_x.stop - h_x.start) / (HX - 1);
    h.min_z = h_z.start; h.delta_z = (h_z.stop - h_z.start) / (HZ - 1); h.d_row_support = d_sup;

    /* ---- K4 on a 48 x 5 mesh (rows of 32 + 16 lanes): point kernel, x-groups (whole mesh / three "ranks" over peer grids) ---- */
    dfcsr_wake_params wp; memset(&wp, 0, sizeof(wp));
    wp.t = t_now; wp.sigma_x = sx; wp.sigma_z = sz; wp.slope0 = 0.0; wp.mean_x = 0.0; wp.formation_window = 0.25;
    wp.csr_scaling = 8.98755e3 * 1e-9; wp.nx = 48; wp.nz = 56; wp.skip_mode = DFCSR_SKIP_OFF;
    const int MX = 48, MZ = 5, N = MX * MZ;
    const dfcsr_axis m_x = axis(-3 * sx, 3 * sx, MX), m_z = axis(-3 * sz, 3 * sz, MZ);
    double *d_a, *d_b, *d_g[3];
    CU(cudaMalloc((void**)&d_a, 16 * N)); CU(cudaMalloc((void**)&d_b, 16 * N));
    for (int p = 0; p < 3; ++p) { CU(cudaMalloc((void**)&d_g[p], 16 * N)); CU(cudaMemset(d_g[p], 0xff, 16 * N)); }
    unsigned long long* d_cnt; CU(cudaMalloc((void**)&d_cnt, 24)); CU(cudaMemset(d_cnt, 0, 24));
    CHECK(dfcsr_wake_grid(&h, &lt, &wp, m_x, m_z, 0.0, 0.0, 0, N, d_a, d_a + N, d_cnt, NULL));
    dfcsr_xgroup_plan plan;
    CHECK(dfcsr_wake_xgroup_plan(&h, &wp, m_x, m_z, &plan));
    if (plan.n_groups != 2 * MZ || plan.group_points != 32) { fprintf(stderr, "plan: %lld groups\n", (long long)plan.n_groups); return 1; }
    void* d_ws; const int64_t ws_bytes = plan.n_groups * plan.workspace_bytes_per_group;
    CU(cudaMalloc(&d_ws, (size_t)ws_bytes));
    CHECK(dfcsr_wake_grid_xgroups(&h, &lt, &wp, m_x, m_z, 0.0, 0.0, 0, plan.n_groups, 1, d_b, d_b + N, NULL, 0, d_ws, ws_bytes, NULL, NULL));
    uint64_t peers[3];
    for (int p = 0; p < 3; ++p) peers[p] = (uint64_t)(uintptr_t)d_g[p];
    for (int r = 0; r < 3; ++r) {
        const int64_t cnt = (plan.n_groups - r + 2) / 3;
        CHECK(dfcsr_wake_grid_xgroups(&h, &lt, &wp, m_x, m_z, 0.0, 0.0, r, cnt, 3, NULL, NULL, peers, 3, d_ws, ws_bytes, NULL, NULL));
    }
    CU(cudaDeviceSynchronize());
    double *a = (double*)malloc(16 * N), *b = (double*)malloc(16 * N), *g = (double*)malloc(16 * N);
    unsigned long long cnt[3];
    CU(cudaMemcpy(a, d_a, 16 * N, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(b, d_b, 16 * N, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cnt, d_cnt, 24, cudaMemcpyDeviceToHost));
    double mde = 0, mk = 0, ede = 0, ek = 0;
    for (int k = 0; k < N; ++k) {
        mde = fmax(mde, fabs(a[k])); mk = fmax(mk, fabs(a[N + k]));
        ede = fmax(ede, fabs(a[k] - b[k])); ek = fmax(ek, fabs(a[N + k] - b[N + k]));
    }
    int ok = mde > 0 && mk > 0 && ede < 1e-12 * mde && ek < 1e-12 * mk && cnt[0] > 0 && cnt[0] < cnt[1] &&
             cnt[1] == (unsigned long long)N * 4ull * wp.nx * wp.nz;
    for (int p = 0; p < 3; ++p) {
        CU(cudaMemcpy(g, d_g[p], 16 * N, cudaMemcpyDeviceToHost));
        ok = ok && memcmp(g, b, 16 * N) == 0;
    }
    /* the point kernel through the same peer interface, points dealt out round-robin */
    for (int p = 0; p < 3; ++p) CU(cudaMemset(d_g[p], 0xff, 16 * N));
    for (int r = 0; r < 3; ++r)
        CHECK(dfcsr_wake_grid_peers(&h, &lt, &wp, m_x, m_z, 0.0, 0.0, r, (N - r + 2) / 3, 3, peers, 3, NULL, NULL));
    CU(cudaDeviceSynchronize());
    for (int p = 0; p < 3; ++p) {
        CU(cudaMemcpy(g, d_g[p], 16 * N, cudaMemcpyDeviceToHost));
        ok = ok && memcmp(g, a, 16 * N) == 0;
    }
    printf("max|dE| %.6e max|kick| %.6e  x-groups vs point kernel: %.2e %.2e (of the maximum)  in-grid %llu of %llu\n", mde, mk,
           ede / mde, ek / mk, cnt[0], cnt[1]);

    /* ---- K5: particles sitting exactly on mesh nodes receive the nodes' wakes (beams.py:108-131) ---- */
    double hxp[3], hzp[3], hpx[3] = {0, 0, 0}, hpz[3] = {0, 0, 0};
    const int pick[3][2] = {{3, 1}, {20, 2}, {45, 3}};
    for (int p = 0; p < 3; ++p) {
        hxp[p] = m_x.start + (m_x.stop - m_x.start) * pick[p][0] / (MX - 1);
        hzp[p] = m_z.start + (m_z.stop - m_z.start) * pick[p][1] / (MZ - 1);
    }
    double *d_px, *d_pz, *d_xp, *d_zp;
    CU(cudaMalloc((void**)&d_xp, 24)); CU(cudaMalloc((void**)&d_zp, 24)); CU(cudaMalloc((void**)&d_px, 24)); CU(cudaMalloc((void**)&d_pz, 24));
    CU(cudaMemcpy(d_xp, hxp, 24, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_zp, hzp, 24, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_px, 0, 24)); CU(cudaMemset(d_pz, 0, 24));
    const double step = 0.1, e0 = 5.0e9;
    CHECK(dfcsr_apply_kick(d_xp, d_zp, d_px, d_pz, 3, 0.0, 0.0, d_b, d_b + N, m_x, m_z, step, e0, 1, NULL));
    CU(cudaMemcpy(hpx, d_px, 24, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(hpz, d_pz, 24, cudaMemcpyDeviceToHost));
    for (int p = 0; p < 3; ++p) {
        const int k = pick[p][0] * MZ + pick[p][1];
        const double want_pz = step * b[k] * 1e6 / e0, want_px = step * b[N + k] * 1e6 / e0;
        ok = ok && fabs(hpz[p] - want_pz) <= 1e-9 * fabs(want_pz) + 1e-300 && fabs(hpx[p] - want_px) <= 1e-9 * fabs(want_px) + 1e-300;
    }
    printf(ok ? "C ABI WAKE OK\n" : "C ABI WAKE FAIL\n");
    return ok ? 0 : 2;
}
