/* Plain-C caller of libdfcsr_b200.so: proves the boundary needs nothing but the header, cudart and device
 * pointers (no Python, no torch).  Built and run by tests/test_gpu_c_abi.py on the GPU box.
 * Deposits 1e5 particles with NGP and CIC, reduces the beam statistics, checks conservation; then repeats statistics
 * and deposit the way two ranks holding half of the particles each would (dfcsr_beam_stats_partial / _final,
 * dfcsr_deposit_cic_q / _finish) and requires the SAME BITS as the single-GPU calls. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dfcsr_b200.h"

#define CHECK(x) do { if ((x) != 0) { fprintf(stderr, "FAIL %s: %s\n", #x, dfcsr_last_error()); return 1; } } while (0)

int main(void) {
    const int64_t n = 100000;
    const int nx = 64, nz = 96;
    double *hx = malloc(n * sizeof(double)), *hz = malloc(n * sizeof(double)), *hp = malloc(n * sizeof(double));
    unsigned long long st = 88172645463325252ULL;
    for (int64_t i = 0; i < n; ++i) {          /* xorshift uniforms in (-1, 1): no libm dependence on the data */
        double u[3];
        for (int k = 0; k < 3; ++k) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; u[k] = (double)(st >> 11) / 9007199254740992.0; }
        hx[i] = (2.0 * u[0] - 1.0) * 1e-4; hz[i] = (2.0 * u[1] - 1.0) * 3e-4; hp[i] = (2.0 * u[2] - 1.0) * 1e-6;
    }
    double *dx, *dz, *dp, *dcount, *dvx, *dstats; long long *dngp; void* ws;
    cudaMalloc((void**)&dx, n * 8); cudaMalloc((void**)&dz, n * 8); cudaMalloc((void**)&dp, n * 8);
    cudaMalloc((void**)&dcount, nx * nz * 8); cudaMalloc((void**)&dvx, nx * nz * 8); cudaMalloc((void**)&dngp, nx * nz * 8);
    cudaMalloc((void**)&dstats, DFCSR_STATS_DOUBLES * 8);
    cudaMalloc(&ws, (size_t)dfcsr_beam_stats_workspace()); cudaMemset(ws, 0, (size_t)dfcsr_beam_stats_workspace());
    cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dz, hz, n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dp, hp, n * 8, cudaMemcpyHostToDevice);
    if (dfcsr_abi_version() != DFCSR_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    CHECK(dfcsr_deposit_ngp(dx, dz, n, nx, -1.2e-4, 1.2e-4, nz, -3.6e-4, 3.6e-4, (int64_t*)dngp, NULL));
    CHECK(dfcsr_deposit_cic(dx, dz, dp, n, nx, -1.2e-4, 1.2e-4, nz, -3.6e-4, 3.6e-4, dcount, dvx, 0, NULL));
    CHECK(dfcsr_beam_stats(dx, dz, dp, dp, n, NULL, dstats, ws, NULL));
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "CUDA error\n"); return 1; }
    long long* hngp = malloc(nx * nz * 8); double* hcount = malloc(nx * nz * 8); double hs[DFCSR_STATS_DOUBLES];
    cudaMemcpy(hngp, dngp, nx * nz * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hcount, dcount, nx * nz * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, dstats, sizeof(hs), cudaMemcpyDeviceToHost);
    long long tot = 0; double ctot = 0.0;
    for (int c = 0; c < nx * nz; ++c) { tot += hngp[c]; ctot += hcount[c]; }
    double mx = 0.0; for (int64_t i = 0; i < n; ++i) mx += hx[i]; mx /= (double)n;
    printf("ngp_total %lld cic_total %.9f mean_x %.6e ref_mean_x %.6e sigma_z %.6e n %.0f\n", tot, ctot, hs[DFCSR_S_MEAN_X], mx,
           hs[DFCSR_S_SIGMA_Z], hs[DFCSR_S_N]);
    int ok = (tot == n) && fabs(ctot - (double)n) < 1e-6 && fabs(hs[DFCSR_S_MEAN_X] - mx) < 1e-18 + 1e-12 * fabs(mx) &&
             hs[DFCSR_S_N] == (double)n && fabs(hs[DFCSR_S_SIGMA_Z] - 3e-4 / sqrt(3.0)) < 3e-6;
    /* two "ranks" on this GPU: chunks [0, 512) and [512, 1024) of the particle index space, one shared table */
    {
        const int64_t chunk = dfcsr_stat_chunk(n);
        const int64_t n0 = 512 * chunk < n ? 512 * chunk : n, n1 = n - n0;
        double *dtab, *dstats2, *dcount2, *dvx2; long long *dq0, *dq1;
        cudaMalloc((void**)&dtab, DFCSR_STAT_BLOCKS * 8 * 8); cudaMalloc((void**)&dstats2, DFCSR_STATS_DOUBLES * 8);
        cudaMemset(dstats2, 0, DFCSR_STATS_DOUBLES * 8);
        cudaMalloc((void**)&dcount2, nx * nz * 8); cudaMalloc((void**)&dvx2, nx * nz * 8);
        cudaMalloc((void**)&dq0, 2 * nx * nz * 8); cudaMalloc((void**)&dq1, 2 * nx * nz * 8);
        for (int pass = 0; pass < 2; ++pass) {
            CHECK(dfcsr_beam_stats_partial(pass, dx, dz, dp, dp, n0, n, 0, 512, NULL, dstats2, dtab, NULL, 0, NULL));
            CHECK(dfcsr_beam_stats_partial(pass, dx + n0, dz + n0, dp + n0, dp + n0, n1, n, 512, 512, NULL, dstats2, dtab, NULL, 0, NULL));
            CHECK(dfcsr_beam_stats_final(pass, dtab, n, NULL, 1, 1, dstats2, NULL));
        }
        double hs2[DFCSR_STATS_DOUBLES];
        cudaMemcpy(hs2, dstats2, sizeof(hs2), cudaMemcpyDeviceToHost);
        ok = ok && memcmp(hs, hs2, 14 * sizeof(double)) == 0;
        const double amax = hs[DFCSR_S_ABSMAX_PX];
        CHECK(dfcsr_deposit_cic_q(dx, dz, dp, n0, n, nx, -1.2e-4, 1.2e-4, nz, -3.6e-4, 3.6e-4, amax, (int64_t*)dq0, NULL));
        CHECK(dfcsr_deposit_cic_q(dx + n0, dz + n0, dp + n0, n1, n, nx, -1.2e-4, 1.2e-4, nz, -3.6e-4, 3.6e-4, amax, (int64_t*)dq1, NULL));
        uint64_t peers[2] = {(uint64_t)(uintptr_t)dq0, (uint64_t)(uintptr_t)dq1};
        CHECK(dfcsr_deposit_cic_finish(peers, 2, nx, nz, n, amax, dcount2, dvx2, NULL, NULL));
        double* hcount2 = malloc(nx * nz * 8); double* hvx = malloc(nx * nz * 8); double* hvx2 = malloc(nx * nz * 8);
        cudaMemcpy(hcount2, dcount2, nx * nz * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hvx, dvx, nx * nz * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hvx2, dvx2, nx * nz * 8, cudaMemcpyDeviceToHost);
        const int same = memcmp(hcount, hcount2, nx * nz * 8) == 0 && memcmp(hvx, hvx2, nx * nz * 8) == 0;
        printf("two-shard statistics bitwise %d, two-shard deposit bitwise %d\n", memcmp(hs, hs2, 14 * sizeof(double)) == 0, same);
        ok = ok && same;
    }
    /* error path: a NULL output pointer is reported, not dereferenced */
    ok = ok && dfcsr_deposit_ngp(dx, dz, n, nx, -1.0, 1.0, nz, -1.0, 1.0, NULL, NULL) == DFCSR_ERR_INVALID;
    puts(ok ? "C ABI OK" : "C ABI FAILED");
    return ok ? 0 : 1;
}
