// track.cu — particle transport through one lattice element, on the device, in place.
//
// Replaces the Bmad-X call of the reference's tracking loop, Beam.track (beams.py:101-106):
//     self.particle = track_element(self.particle, element)
// for the element types CSR2D.get_bmadx_element builds (CSR.py:146-199): Drift, SBend (with the FRINGE_AT variants the
// step splitting uses), Quadrupole, Sextupole.  Bmad-X itself is a third-party package that is not vendored in the
// reference (environment.yml:24, unpinned git HEAD); its maps are restated here from the published algorithms of
// Bmad (D. Sagan, "The Bmad Reference Manual", tracking chapters) that Bmad-X transcribes:
//   drift        exact:  x += L px/pl, y += L py/pl, z += L (beta/beta_ref - 1/pl) with pl = sqrt(1 - (px^2+py^2)/(1+pz)^2),
//                beta/beta_ref - 1 evaluated without cancellation (sqrt_one);
//   sbend        hard-edge "linear_edge" kicks  px += g tan(e) x, py -= g tan(e) y  at the ends FRINGE_AT selects, and the
//                EXACT solution of the motion in the body (uniform field matched to the reference curvature g), written so
//                that every division by g is taken analytically (valid down to g = 0, where it is the exact drift);
//   quadrupole   Bmad's quad_mat2_calc with the momentum-dependent strength k1/(1+pz), its second-order path-length
//                terms and low_energy_z_correction;
//   sextupole    thin kick of integrated strength K2 L between two exact half drifts (stand-in: the reference holds no
//                known answer for it).
// The three single-particle known answers the reference holds (test/test_BmadX_tracking.ipynb cells 25, 28, 31:
// 1e-3 * ones(6) at p0c = 4e7 through Drift(L=1), SBend(L=0.1, G=0.5, E2=0.1), Quadrupole(L=0.1, K1=10)) are reproduced
// to <= 3e-13 relative (tests/test_tracking.py on the host restatement, tests/test_gpu_kernels.py on this kernel).
// Bound: HBM, 96 B / particle (six coordinates read and written).
#include <math_constants.h>
#include "common.cuh"

namespace dfcsr {

struct TrackParams {
    int kind;                 // dfcsr_element_kind
    int fringe_in, fringe_out;
    int n_step;
    double L, g, e1, e2, k1, k2;
    double p0c, mc2;
};

__host__ __device__ inline double sqrt_one(double x) { return x / (sqrt(1.0 + x) + 1.0); }      // sqrt(1 + x) - 1

__device__ __forceinline__ double asin_over(double u) {       // asin(u) / u
    if (fabs(u) < 1e-4) {
        const double u2 = u * u;
        return 1.0 + u2 * (1.0 / 6.0 + u2 * (3.0 / 40.0 + u2 * (15.0 / 336.0)));
    }
    return asin(u) / u;
}

// beta / beta_ref - 1 for relative momentum deviation pz
__device__ __forceinline__ double beta_ratio_minus_one(double pz, double p0c, double mc2) {
    const double P = 1.0 + pz;
    return sqrt_one((mc2 * mc2 * (2.0 * pz + pz * pz)) / ((p0c * P) * (p0c * P) + mc2 * mc2));
}

__device__ __forceinline__ void drift_exact(double& x, double px, double& y, double py, double& z, double pz, double L,
                                            double p0c, double mc2) {
    const double P = 1.0 + pz;
    const double Px = px / P, Py = py / P;
    const double Pxy2 = Px * Px + Py * Py;
    const double Pl = sqrt(1.0 - Pxy2);
    x = x + L * Px / Pl;
    y = y + L * Py / Pl;
    z = z + L * (beta_ratio_minus_one(pz, p0c, mc2) + sqrt_one(-Pxy2) / Pl);
}

// Exact map of a sector bend body: curvature g, field matched to it, length L.  S = sin(theta)/g and
// C = (1 - cos(theta))/g are passed in (finite for g -> 0).
__device__ __forceinline__ void bend_body_exact(double& x, double& px, double& y, double py, double& z, double pz, double L,
                                                double g, double ct, double st, double S, double C, double p0c, double mc2) {
    const double P = 1.0 + pz;
    const double pt2 = P * P - py * py;                 // in-plane momentum squared
    const double ps = sqrt(pt2 - px * px);
    const double a = ps - 1.0 - g * x;
    const double dpx_g = -px * C + a * S;               // (px_f - px) / g
    const double pxf = px + g * dpx_g;
    const double psf = sqrt(pt2 - pxf * pxf);
    const double dps_g = -(pxf + px) / (psf + ps) * dpx_g;     // (ps_f - ps) / g
    const double xf = x * ct + (ps - 1.0) * C + px * S + dps_g;
    const double D = (px * dps_g - ps * dpx_g) / pt2;          // sin(asin(px/pt) - asin(px_f/pt)) / g
    const double path_over_P = L + D * asin_over(D * g);       // particle path length / P
    y = y + py * path_over_P;
    z = z + L * beta_ratio_minus_one(pz, p0c, mc2) - pz * L - P * D * asin_over(D * g);
    x = xf;
    px = pxf;
}

// Bmad quad_mat2_calc: 2 x 2 matrix and path-length coefficients of one transverse plane
__device__ __forceinline__ void quad_mat2(double k1, double len, double rel_p, double& a11, double& a12, double& a21,
                                          double& c1, double& c2, double& c3) {
    const double sqrt_k = sqrt(fabs(k1) + 2.220446049250313e-16);
    const double sk_l = sqrt_k * len;
    double cx, sx;
    if (k1 > 0.0) { cx = cosh(sk_l); sx = sinh(sk_l) / sqrt_k; }
    else { cx = cos(sk_l); sx = sin(sk_l) / sqrt_k; }
    a11 = cx;
    a12 = sx / rel_p;
    a21 = k1 * sx * rel_p;
    c1 = k1 * (-cx * sx + len) / 4.0;
    c2 = -k1 * sx * sx / (2.0 * rel_p);
    c3 = -(cx * sx + len) / (4.0 * rel_p * rel_p);
}

__device__ __forceinline__ double low_energy_z_correction(double pz, double p0c, double mc2, double ds) {
    const double e_tot = sqrt(p0c * p0c + mc2 * mc2);
    const double beta0 = p0c / e_tot;
    const double ev = mc2 * (beta0 * pz) * (beta0 * pz);
    if (ev < 3e-7 * e_tot) {
        const double m2 = (mc2 / e_tot) * (mc2 / e_tot), b2 = beta0 * beta0;
        return ds * pz * (1.0 - 3.0 * (pz * b2) / 2.0 + pz * pz * b2 * (2.0 * b2 - m2 / 2.0)) * m2;
    }
    const double pc = (1.0 + pz) * p0c;
    const double beta = pc / sqrt(pc * pc + mc2 * mc2);
    return ds * (beta - beta0) / beta0;
}

__global__ void __launch_bounds__(256)
track_element_kernel(double* __restrict__ X, double* __restrict__ PX, double* __restrict__ Y, double* __restrict__ PY,
                     double* __restrict__ Z, double* __restrict__ PZ, long long n, TrackParams T) {
    double ct = 1.0, st = 0.0, S = T.L, C = 0.0, t1 = 0.0, t2 = 0.0;
    if (T.kind == DFCSR_ELEM_SBEND) {
        const double th = T.g * T.L;
        ct = cos(th);
        st = sin(th);
        if (fabs(th) < 1e-7) { S = T.L * (1.0 - th * th / 6.0); C = 0.5 * T.g * T.L * T.L; }
        else { S = st / T.g; const double sh = sin(0.5 * th); C = 2.0 * sh * sh / T.g; }
        t1 = T.fringe_in ? T.g * tan(T.e1) : 0.0;
        t2 = T.fringe_out ? T.g * tan(T.e2) : 0.0;
    }
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        double x = X[p], px = PX[p], y = Y[p], py = PY[p], z = Z[p];
        const double pz = PZ[p];
        if (T.kind == DFCSR_ELEM_DRIFT) {
            drift_exact(x, px, y, py, z, pz, T.L, T.p0c, T.mc2);
        } else if (T.kind == DFCSR_ELEM_SBEND) {
            px = px + t1 * x;
            py = py - t1 * y;
            bend_body_exact(x, px, y, py, z, pz, T.L, T.g, ct, st, S, C, T.p0c, T.mc2);
            px = px + t2 * x;
            py = py - t2 * y;
        } else if (T.kind == DFCSR_ELEM_QUADRUPOLE) {
            const double step = T.L / (double)T.n_step;
            for (int i = 0; i < T.n_step; ++i) {
                const double rel_p = 1.0 + pz;
                const double k1 = T.k1 / rel_p;
                double a11, a12, a21, c1, c2, c3, b11, b12, b21, d1, d2, d3;
                quad_mat2(-k1, step, rel_p, a11, a12, a21, c1, c2, c3);
                quad_mat2(k1, step, rel_p, b11, b12, b21, d1, d2, d3);
                z = z + c1 * x * x + c2 * x * px + c3 * px * px + d1 * y * y + d2 * y * py + d3 * py * py;
                const double xn = a11 * x + a12 * px, pxn = a21 * x + a11 * px;
                const double yn = b11 * y + b12 * py, pyn = b21 * y + b11 * py;
                x = xn; px = pxn; y = yn; py = pyn;
                z = z + low_energy_z_correction(pz, T.p0c, T.mc2, step);
            }
        } else {                                          // sextupole: half drift, thin kick, half drift
            drift_exact(x, px, y, py, z, pz, 0.5 * T.L, T.p0c, T.mc2);
            const double kl = T.k2 * T.L;
            px = px - 0.5 * kl * (x * x - y * y);
            py = py + kl * x * y;
            drift_exact(x, px, y, py, z, pz, 0.5 * T.L, T.p0c, T.mc2);
        }
        X[p] = x; PX[p] = px; Y[p] = y; PY[p] = py; Z[p] = z;
    }
}

}  // namespace dfcsr

using namespace dfcsr;

extern "C" int dfcsr_track_element(double* d_x, double* d_px, double* d_y, double* d_py, double* d_z, double* d_pz,
                                   int64_t n, const dfcsr_element* el, double p0c, double mc2, void* stream) {
    DFCSR_REQUIRE(el && n >= 0, "null element or negative count");
    DFCSR_REQUIRE(el->kind >= DFCSR_ELEM_DRIFT && el->kind <= DFCSR_ELEM_SEXTUPOLE, "unknown element kind");
    DFCSR_REQUIRE(p0c > 0.0 && mc2 > 0.0, "reference momentum and rest energy must be positive");
    DFCSR_REQUIRE(el->kind != DFCSR_ELEM_QUADRUPOLE || (el->n_step >= 1 && el->n_step <= 1000), "quadrupole steps out of range");
    if (n == 0) return DFCSR_OK;
    DFCSR_REQUIRE(d_x && d_px && d_y && d_py && d_z && d_pz, "null pointer");
    TrackParams T;
    T.kind = el->kind;
    T.fringe_in = el->fringe_entrance;
    T.fringe_out = el->fringe_exit;
    T.n_step = el->n_step < 1 ? 1 : el->n_step;
    T.L = el->L; T.g = el->g; T.e1 = el->e1; T.e2 = el->e2; T.k1 = el->k1; T.k2 = el->k2;
    T.p0c = p0c; T.mc2 = mc2;
    long long want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < 148LL * 32 ? want : 148LL * 32);
    track_element_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_x, d_px, d_y, d_py, d_z, d_pz, n, T);
    count_launch(1);
    DFCSR_CUDA_OK(cudaGetLastError());
    return DFCSR_OK;
}
