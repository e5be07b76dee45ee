/*
 * dfcsr_b200.h — C ABI of libdfcsr_b200.so: the B200 (sm_100a) implementation of pyDFCSR's
 * per-step CSR-wake hot path.
 *
 * The reference (slaclab/pyDFCSR) has no FFI: the path is reached through Python method calls
 * (SURVEY.md §8(b)).  Each entry point below replaces the arithmetic behind one of those calls
 * and cites it (paths relative to /root/reference/pyDFCSR_2D/).  The Python shim in
 * pydfcsr_b200/ keeps the reference's method names and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain C: pointers, sizes, PODs.  No torch / C++ types cross the boundary.
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device; h_* is host memory.
 *     The caller owns all memory (any allocator); the library allocates nothing persistent.
 *   - every function returns 0 on success or a negative dfcsr_status; dfcsr_last_error() returns
 *     a thread-local message.  No exception crosses the boundary.
 *   - work is enqueued on the caller's stream (void* = cudaStream_t, NULL = default stream) and
 *     the call returns without synchronising unless stated.
 *   - all floating-point data are IEEE fp64 ("double"), C order.
 *
 * Device layouts
 *   - "field stack" (SoA): 5 separate (X,Z) arrays, order = dfcsr_field.
 *   - "voxel slice" (AoS-6): (X,Z,6) doubles = {density, density_x, density_z, vx, vx_x, 0};
 *     48-byte voxels so that one voxel is three 16-byte vector loads and the two z-neighbours of a
 *     trilinear cell are 96 contiguous bytes.
 *   - "history ring": cap voxel slices, slice k of the window lives in slot (head + k) % cap.
 *   - "lattice table": (ns,6) doubles = {X0, Y0, n_x, n_y, tau_x, tau_y} per sample.
 */
#ifndef DFCSR_B200_H
#define DFCSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFCSR_ABI_VERSION 4
#define DFCSR_VOXEL_DOUBLES 6   /* fp64 voxel: 48 bytes */
#define DFCSR_VOXEL_FLOATS 8    /* fp32 voxel: 32 bytes {density, density_x, density_z, vx, vx_x, 0, 0, 0} */
#define DFCSR_LATTICE_DOUBLES 6
#define DFCSR_STATS_DOUBLES 16
#define DFCSR_DF_SCALARS 8
#define DFCSR_MAX_ELEMENTS 256
#define DFCSR_STAT_BLOCKS 1024      /* chunks of the particle index space in the N-independent reductions */
#define DFCSR_MAX_PEERS 8          /* ranks of one NVLink/NVSwitch box (dfcsr_wake_grid_peers) */

typedef enum dfcsr_status {
    DFCSR_OK = 0,
    DFCSR_ERR_INVALID = -1,   /* bad argument                                  */
    DFCSR_ERR_CUDA = -2,      /* CUDA runtime error (message has the string)   */
    DFCSR_ERR_UNSUPPORTED = -3, /* size outside what the kernels support        */
    DFCSR_ERR_WORKSPACE = -4  /* workspace too small                           */
} dfcsr_status;

/* storage format of the history ring.  F64 is the parity mode (wakes within 1e-10 of the reference);
 * F32 is the optional mixed-precision mode of BASELINE.json north_star: the five fields are STORED and
 * blended in fp32, while geometry, retarded time, cell indices, integrand algebra and the quadrature
 * stay fp64 (wakes within 1e-4, measured ~1e-7). */
typedef enum dfcsr_voxel_format { DFCSR_VOXEL_F64 = 0, DFCSR_VOXEL_F32 = 1 } dfcsr_voxel_format;

typedef enum dfcsr_field {
    DFCSR_DENSITY = 0, DFCSR_DENSITY_X = 1, DFCSR_DENSITY_Z = 2, DFCSR_VX = 3, DFCSR_VX_X = 4
} dfcsr_field;

/* indices into the DFCSR_STATS_DOUBLES block written by dfcsr_beam_stats */
typedef enum dfcsr_stat {
    DFCSR_S_MEAN_X = 0, DFCSR_S_MEAN_Z = 1, DFCSR_S_SIGMA_X = 2, DFCSR_S_SIGMA_Z = 3,
    DFCSR_S_SLOPE = 4, DFCSR_S_INTERCEPT = 5,      /* polyfit(z, x, 1)            */
    DFCSR_S_MEAN_XT = 6, DFCSR_S_SIGMA_XT = 7,     /* x - polyval(slope, z)       */
    DFCSR_S_SLICE_SIGMA_X = 8, DFCSR_S_SLICE_COUNT = 9, /* |z| < 0.1 sigma_z slice */
    DFCSR_S_MEAN_PZ = 10, DFCSR_S_SIGMA_PZ = 11, DFCSR_S_N = 12,
    DFCSR_S_ABSMAX_PX = 13                         /* max |px| (for dfcsr_deposit_cic_q); -1 when px was not passed */
} dfcsr_stat;

/* uniform grid described the way numpy.linspace builds it: node i = i*step + start, last = stop */
typedef struct dfcsr_axis {
    double start;
    double stop;
    int32_t n;
    int32_t _pad;
} dfcsr_axis;

/* device-resident (t', x, z) history published by DF_tracker.build_interpolant (deposit.py:395-426) */
typedef struct dfcsr_history {
    const void* d_ring;        /* cap voxel slices (doubles or floats, see format)   */
    int64_t slice_elems;       /* scalars per slice: X * Z * DFCSR_VOXEL_DOUBLES (or _FLOATS) */
    int32_t cap;               /* slots in the ring                                  */
    int32_t head;              /* slot of the oldest slice in the window             */
    int32_t T, X, Z;           /* window depth and slice shape                       */
    int32_t format;            /* dfcsr_voxel_format                                 */
    double min_t, min_x, min_z;      /* deposit.py:416-418 (min_x / min_y / min_z there) */
    double delta_t, delta_x, delta_z;/* deposit.py:419-421                               */
    const int32_t* d_row_support;    /* (cap, X, 2) int32 or NULL: per (slot, transverse row) the hull [z_lo, z_hi] of
                                        the voxels whose density or density gradient is non-zero or whose velocity
                                        fields are not finite (z_lo > z_hi: none), written by dfcsr_history_regrid
                                        (or dfcsr_history_row_support).  Every term of the integrand carries a
                                        factor rho or grad rho of the retarded point (CSR.py:732-775), so a sample whose
                                        eight voxels lie outside the hulls contributes exactly 0 and K4 skips it
                                        without touching the history.  NULL = unknown, nothing is skipped.      */
} dfcsr_history;

/* reference-orbit tables consumed by the integrand (lattice.py:136-143, CSR.py:619-656) */
typedef struct dfcsr_lattice {
    const double* d_table;     /* (ns, 6) lattice table                              */
    int32_t ns;
    int32_t n_elements;        /* <= DFCSR_MAX_ELEMENTS                              */
    double min_s, delta_s;
    const double* d_rho;       /* (n_elements) curvature per element                 */
    const double* d_distance;  /* (n_elements) cumulative end position per element   */
} dfcsr_lattice;

/* scalars read by get_CSR_wake (CSR.py:456-467, 539; CSR.py:80) */
typedef struct dfcsr_wake_params {
    double t;                  /* beam.position                                      */
    double sigma_x, sigma_z;   /* beam._sigma_x / _sigma_z                           */
    double slope0;             /* beam._slope[0]                                     */
    double mean_x;             /* beam._mean_x                                       */
    double formation_window;   /* n_formation_length * formation_length              */
    double csr_scaling;        /* 8.98755e3 * charge                                 */
    int32_t nx, nz;            /* integration_params.xbins / zbins                   */
    int32_t skip_mode;         /* dfcsr_skip_mode: zero-density skipping policy      */
    int32_t reserved;          /* 0                                                  */
} dfcsr_wake_params;

/* Zero-density skipping in the wake kernel (needs dfcsr_history.d_row_support; never changes a bit of the result):
 * AUTO = on when the history grid is sparse by construction (|slope0| > 1, or grid area > 1.5x the +-5 sigma box of the
 * bunch), where it halves the run time; ON / OFF force it (ON costs ~2-5 % on a bunch that fills its grid). */
typedef enum { DFCSR_SKIP_AUTO = 0, DFCSR_SKIP_ON = 1, DFCSR_SKIP_OFF = 2 } dfcsr_skip_mode;

int dfcsr_abi_version(void);
const char* dfcsr_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench accounting) */
int64_t dfcsr_launch_count(void);

/* ---- A14 beam scalars (beams.py:88-98,137-156,201-215; deposit.py:147-159) --------------------
 * Two reduction passes over (x, z[, pz][, px]); results land in d_stats[DFCSR_STATS_DOUBLES].
 * The reductions do not depend on how the particles are distributed: the index space [0, n) is cut into
 * DFCSR_STAT_BLOCKS contiguous chunks of dfcsr_stat_chunk(n) particles, one CTA reduces one chunk in a fixed order and a
 * (1024, 8) table of chunk totals is summed in a fixed order.  A run that shards the particles over N GPUs in whole
 * chunks (dfcsr_beam_stats_partial on every rank, rows stored into the tables of all ranks over NVLink peer memory, a
 * barrier, dfcsr_beam_stats_final) therefore returns the SAME BITS as dfcsr_beam_stats on one GPU.
 * h_centre: 3 HOST doubles (x, z, pz) about which pass A accumulates, e.g. the previous means (NULL = zeros); it
 * must be the same on every rank.  d_workspace needs dfcsr_beam_stats_workspace() bytes, ZERO-INITIALISED once by the
 * caller (the calls leave it reusable).  d_pz and d_px may be NULL. */
int64_t dfcsr_beam_stats_workspace(void);
int64_t dfcsr_stat_chunk(int64_t n_total);     /* particles per chunk: ceil(n_total / DFCSR_STAT_BLOCKS) */
int dfcsr_beam_stats(const double* d_x, const double* d_z, const double* d_pz, const double* d_px, int64_t n,
                     const double* h_centre, double* d_stats, void* d_workspace, void* stream);
/* pass 0 (moments) or 1 (residuals; reads pass 0's results from d_stats) over this rank's shard: chunks
 * [first_block, first_block + n_blocks) = n_local particles.  Rows go to d_table ((1024, 8) doubles) and to the tables
 * of the n_peers ranks listed in h_peer_tables (HOST array of addresses valid in this process; may include d_table). */
int dfcsr_beam_stats_partial(int32_t pass, const double* d_x, const double* d_z, const double* d_pz, const double* d_px,
                             int64_t n_local, int64_t n_total, int32_t first_block, int32_t n_blocks,
                             const double* h_centre, double* d_stats, double* d_table,
                             const uint64_t* h_peer_tables, int32_t n_peers, void* stream);
int dfcsr_beam_stats_final(int32_t pass, const double* d_table, int64_t n_total, const double* h_centre,
                           int32_t have_pz, int32_t have_px, double* d_stats, void* stream);

/* n doubles from device memory into PINNED host memory (cudaHostAlloc / cudaHostRegister, mapped), stored by a one-warp
 * kernel instead of the device-to-host copy engine, so that a small result the host is waiting for (the 16 statistics
 * of dfcsr_beam_stats) does not queue behind a bulk download running on another stream.  Visible to the host once an
 * event recorded after it on `stream` has completed.  Fails with DFCSR_ERR_CUDA if h_dst is not mapped pinned memory. */
int dfcsr_mirror_to_host(const double* d_src, double* h_dst, int32_t n, void* stream);

/* ---- 6 x 6 phase-space covariance (twiss.py:2-71: np.cov([x, px, pz]) and np.cov([y, py, pz]) per step,
 * CSR.py:837-859) -- one read of the six coordinate arrays, same chunked, distribution-independent reduction.
 * d_out[27]: means of (x, px, y, py, z, pz), then the upper triangle (i <= j, row-major) of the covariance
 * with np.cov's 1/(n-1) normalisation.  h_centre: 6 HOST doubles (NULL = zeros).  d_workspace:
 * dfcsr_beam_cov_workspace() bytes, zero-initialised once.  _partial / _final: as for the statistics, table (1024, 27). */
int64_t dfcsr_beam_cov_workspace(void);
int dfcsr_beam_cov(const double* d_x, const double* d_px, const double* d_y, const double* d_py,
                   const double* d_z, const double* d_pz, int64_t n, const double* h_centre, double* d_out,
                   void* d_workspace, void* stream);
int dfcsr_beam_cov_partial(const double* d_x, const double* d_px, const double* d_y, const double* d_py,
                           const double* d_z, const double* d_pz, int64_t n_local, int64_t n_total,
                           int32_t first_block, int32_t n_blocks, const double* h_centre, double* d_table,
                           const uint64_t* h_peer_tables, int32_t n_peers, void* stream);
int dfcsr_beam_cov_final(const double* d_table, int64_t n_total, const double* h_centre, double* d_out, void* stream);

/* ---- A1 / K1 particle deposition (deposit.py:42-87, called at deposit.py:172,178) --------------
 * One pass deposits both weights (w = 1 and w = px) with CIC on an (nx, nz) grid whose bin spacing
 * is (end - start) / n.  d_count / d_vxsum (nx*nz doubles each) are zeroed by the call.
 * mode 0 = automatic (4 for n >= 65536, else 5);
 * 4 = 64-bit fixed-point block-private shared-memory tile, 5 = 64-bit fixed-point L2 reductions: every particle's
 *     contribution is rounded ONCE to a 2^-f grid that depends on n only, everything after is integer addition:
 *     bit-reproducible from run to run, independent of the launch geometry and of how the particles are split over
 *     GPUs; cell sums within ~1e-13 of the largest cell (a non-finite px makes d_vxsum NaN everywhere and leaves
 *     d_count unaffected);
 * 1 = fp64 tile + warp match, 2 = fp64 L2 reductions, 3 = fp64 tile (summation order, hence the last bits,
 *     vary from run to run). */
int dfcsr_deposit_cic(const double* d_x, const double* d_z, const double* d_px, int64_t n,
                      int32_t nx, double x_start, double x_end,
                      int32_t nz, double z_start, double z_end,
                      double* d_count, double* d_vxsum, int32_t mode, void* stream);

/* NGP counts (no reference counterpart): i = floor((q - start)/spacing + 0.5),
 * +1 iff both indices are in range; int64 counts, bit-exact for any summation order. */
/* The fixed-point deposit in two stages, for particles sharded over ranks (CSR.py replicates the particles on every
 * MPI rank; here K1 may take 1/N of them per GPU):
 * _q      deposits this rank's n_local particles into d_q, a (2, nx*nz) int64 buffer [count | vxsum] (zeroed by the
 *         call) at the fixed-point scales of the WHOLE bunch (n_total particles, absmax_px = max |px| over all of
 *         them = stats[DFCSR_S_ABSMAX_PX] of the statistics pass);
 * _finish adds the buffers of all ranks -- h_peer_q: n_peers HOST entries, addresses valid in this process (NVLink peer
 *         mappings; a single rank passes its own buffer) -- and converts to the fp64 grids of dfcsr_deposit_cic.
 *         d_count_max (may be NULL): receives max(count) as the bit pattern of the double, for dfcsr_make_df.
 * Integer addition is exact: every rank gets the bits of dfcsr_deposit_cic (mode 4/5) over all particles on one GPU.
 * The caller puts a cross-rank barrier between the two stages and before the buffers are written again. */
int dfcsr_deposit_cic_q(const double* d_x, const double* d_z, const double* d_px, int64_t n_local, int64_t n_total,
                        int32_t nx, double x_start, double x_end, int32_t nz, double z_start, double z_end,
                        double absmax_px, int64_t* d_q, void* stream);
int dfcsr_deposit_cic_finish(const uint64_t* h_peer_q, int32_t n_peers, int32_t nx, int32_t nz, int64_t n_total,
                             double absmax_px, double* d_count, double* d_vxsum, uint64_t* d_count_max, void* stream);

int dfcsr_deposit_ngp(const double* d_x, const double* d_z, int64_t n,
                      int32_t nx, double x_start, double x_end,
                      int32_t nz, double z_start, double z_end,
                      int64_t* d_count, void* stream);

/* ---- A2-A4 / K2 density functions (deposit.py:183-235) -----------------------------------------
 * count / vxsum -> the five smoothed fields.  The Savitzky-Golay operators for (window, order) are
 * computed once on the host in fp64 and kept on the device by the caller: d_taps[window],
 * d_edge_lo[half*window], d_edge_hi[half*window] (half = window/2; scipy's mode='interp' edge fit).
 * d_fields: field stack (5, nx, nz).  d_scalars[DFCSR_DF_SCALARS] receives
 *   [0] max(count)  [1] threshold  [2] trapz normalisation  [3] max(density)
 *   [4] mean(vx_x) after the mask fill (= fill value used by the re-gridding, deposit.py:332)
 *   [5] masked mean  [6] number of cells above the second threshold.
 * d_count_max (may be NULL): max(count) as delivered by dfcsr_deposit_cic_finish; NULL = reduced here (one more launch).
 * d_workspace needs dfcsr_make_df_workspace(nx, nz) bytes. */
int64_t dfcsr_make_df_workspace(int32_t nx, int32_t nz);
int dfcsr_make_df(const double* d_count, const double* d_vxsum, dfcsr_axis x_axis, dfcsr_axis z_axis,
                  int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                  double velocity_threshold, const uint64_t* d_count_max, double* d_fields, double* d_scalars,
                  void* d_workspace, void* stream);

/* DF_tracker.get_DF (deposit.py:145-245) on one GPU in ONE call: dfcsr_deposit_cic_q + dfcsr_deposit_cic_finish +
 * dfcsr_make_df back to back on `stream` with the same arguments and the same results (the five kernels are the
 * same; only the host round trips between them are gone: three binding calls cost ~100 us of host time per lattice
 * step, more than the kernels take).  d_q: (2, nx*nz) int64 scratch; d_count / d_vxsum: the deposit grids (outputs,
 * kept for the caller); d_count_max: one uint64 of scratch. */
int dfcsr_get_df(const double* d_x, const double* d_z, const double* d_px, int64_t n, dfcsr_axis x_axis, dfcsr_axis z_axis,
                 double absmax_px, int64_t* d_q, double* d_count, double* d_vxsum, uint64_t* d_count_max,
                 int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                 double velocity_threshold, double* d_fields, double* d_scalars, void* d_workspace, void* stream);

/* The same, enqueued BEFORE the host has read the statistics: the grid limits mean -+ lim * sigma (deposit.py:160-171)
 * and max|px| are taken from d_stats, the DEVICE vector dfcsr_beam_stats is writing on the same stream, by a one-thread
 * kernel that evaluates the Python expression with its roundings; d_limits (4 doubles, caller-owned) receives
 * {x_lo, x_hi, z_lo, z_hi}.  The caller chooses the grid shape (nx, nz, window) -- deposit.py:157-167 derives it from the
 * statistics, so it is a guess (the previous step's) that the caller checks once the statistics have arrived; if it was
 * right, d_fields / d_scalars are bitwise what dfcsr_get_df gives with the host-computed limits, and the GPU has not
 * waited for the host in between. */
int dfcsr_get_df_from_stats(const double* d_x, const double* d_z, const double* d_px, int64_t n, const double* d_stats,
                            double xlim, double zlim, int32_t nx, int32_t nz, double* d_limits, int64_t* d_q,
                            double* d_count, double* d_vxsum, uint64_t* d_count_max, int32_t window, const double* d_taps,
                            const double* d_edge_lo, const double* d_edge_hi, double velocity_threshold, double* d_fields,
                            double* d_scalars, void* d_workspace, void* stream);

/* The stages of dfcsr_get_df_from_stats one by one, for particles sharded over ranks (a cross-rank barrier separates the
 * two deposit stages, as with dfcsr_deposit_cic_q / _finish): dfcsr_df_limits writes {x_lo, x_hi, z_lo, z_hi} from the
 * device statistics; the _dev forms take the limits from d_limits and max|px| from d_stats instead of host scalars. */
int dfcsr_df_limits(const double* d_stats, double xlim, double zlim, double* d_limits, void* stream);
int dfcsr_deposit_cic_q_dev(const double* d_x, const double* d_z, const double* d_px, int64_t n_local, int64_t n_total,
                            int32_t nx, int32_t nz, const double* d_limits, const double* d_stats, int64_t* d_q, void* stream);
int dfcsr_deposit_cic_finish_dev(const uint64_t* h_peer_q, int32_t n_peers, int32_t nx, int32_t nz, int64_t n_total,
                                 const double* d_stats, double* d_count, double* d_vxsum, uint64_t* d_count_max, void* stream);
int dfcsr_make_df_dev(const double* d_count, const double* d_vxsum, int32_t nx, int32_t nz, const double* d_limits,
                      int32_t window, const double* d_taps, const double* d_edge_lo, const double* d_edge_hi,
                      double velocity_threshold, const uint64_t* d_count_max, double* d_fields, double* d_scalars,
                      void* d_workspace, void* stream);

/* ---- A7 / K3 bilinear re-gridding into a history slot (deposit.py:296-309,328-332,379-390) -----
 * Samples the five fields of one raw density-function record (field stack on src axes) on the
 * history grid and writes one voxel slice.  Out-of-source points get 0, or the fill value for vx_x
 * (np.mean(vx_x), deposit.py:332,384): *d_fill_vx_x when that device pointer is non-NULL (e.g.
 * scalars[4] of dfcsr_make_df, so the host never has to read it back), else fill_vx_x.
 * d_row_support (may be NULL): the slot's row-support table (next entry) is written by the same kernel, from the
 * voxels it has just computed, instead of re-reading the slice. */
int dfcsr_history_regrid(const double* d_fields, dfcsr_axis src_x, dfcsr_axis src_z,
                         dfcsr_axis dst_x, dfcsr_axis dst_z, double fill_vx_x, const double* d_fill_vx_x,
                         int32_t format, void* d_slice, int32_t* d_row_support, void* stream);

/* Row support of one voxel slice (see dfcsr_history.d_row_support): d_support[X][2] = {z_lo, z_hi} per row, the
 * first and last z index whose density, d(density)/dx or d(density)/dz is not exactly zero, or whose vx / d(vx)/dx
 * is not finite (NaN counts as non-zero: a sample the reference would turn into NaN is never skipped);
 * {INT32_MAX, -1} for a row without any.  For slices written by dfcsr_history_pack (imported histories);
 * dfcsr_history_regrid fills the table itself. */
int dfcsr_history_row_support(const void* d_slice, int32_t X, int32_t Z, int32_t format, int32_t* d_support, void* stream);

/* field stack (5, X, Z) <-> voxel slice (X, Z, 6): import/export of oracle histories in tests */
int dfcsr_history_pack(const double* d_fields, int32_t X, int32_t Z, int32_t format, void* d_slice, void* stream);
int dfcsr_history_unpack(const void* d_slice, int32_t X, int32_t Z, int32_t format, double* d_fields, void* stream);

/* ---- A10-A12 / K4 wake on the observation mesh (CSR.py:397-451, 454-602, 605-782) --------------
 * For k in [0, count): s = t + d_zmesh[first + k], x = d_xmesh[first + k];
 * d_dE[k], d_kick[k] = get_CSR_wake(s, x).  first/count implement the reference's MPI block split
 * (CSR.py:121-125, 434-445).  d_counters (may be NULL, else THREE counters): [0] += in-grid integrand samples the
 * kernel located (with d_row_support, s' nodes that provably miss the band of non-zero density rows are not swept),
 * [1] += samples the reference evaluates, [2] += in-grid samples whose history voxels were actually gathered
 * (= [0] without d_row_support) -- device-side accounting for the roofline figure. */
int dfcsr_wake_mesh(const dfcsr_history* hist, const dfcsr_lattice* lat, const dfcsr_wake_params* wp,
                    const double* d_xmesh, const double* d_zmesh, int64_t first, int64_t count,
                    double* d_dE, double* d_kick, unsigned long long* d_counters, void* stream);

/* Same, with the observation mesh generated on the device exactly as get_CSR_mesh builds it
 * (CSR.py:380-389): point k = (ix, iz) = divmod(k, z_axis.n); z = linspace node iz of z_axis
 * (CSR_zrange); x = linspace node ix of x_axis (CSR_xrange_transformed) + (slope*z + intercept).
 * No mesh arrays have to be built or uploaded per step. */
int dfcsr_wake_grid(const dfcsr_history* hist, const dfcsr_lattice* lat, const dfcsr_wake_params* wp,
                    dfcsr_axis x_axis, dfcsr_axis z_axis, double slope, double intercept,
                    int64_t first, int64_t count, double* d_dE, double* d_kick,
                    unsigned long long* d_counters, void* stream);

/* K4 fused with the exchange step (comm.Allgatherv x2, CSR.py:447-448).  Instead of filling a local send buffer that a
 * collective then distributes, the kernel stores both results of every observation point of this rank's block
 * straight into the wake grid of EVERY rank through NVLink peer mappings: h_peer_grids holds n_peers HOST entries,
 * entry p = the address, valid in THIS process (CUDA IPC / symmetric-memory mapping; this rank's own entry is its
 * local buffer), of rank p's (2, x_axis.n * z_axis.n) fp64 grid [dE | kick]; point k of the launch is mesh point
 * first + k * stride and lands at that index of both halves.  stride = 1 with the reference's count/displ gives the
 * contiguous blocks of CSR.py:121-125; first = rank, stride = n_ranks deals the points out round-robin, which balances
 * ranks whose blocks would hold different numbers of in-grid samples (the gathered grid is bitwise the same).  The caller orders the launch after the peers' last readers of those grids and publishes
 * completion with a barrier across ranks (pydfcsr_b200/distributed.py: two alternating grids + one barrier per step). */
int dfcsr_wake_grid_peers(const dfcsr_history* hist, const dfcsr_lattice* lat, const dfcsr_wake_params* wp,
                          dfcsr_axis x_axis, dfcsr_axis z_axis, double slope, double intercept,
                          int64_t first, int64_t count, int64_t stride, const uint64_t* h_peer_grids, int32_t n_peers,
                          unsigned long long* d_counters, void* stream);

/* ---- K4, x-group mapping (csrc/wake_xgroup.cuh): same get_CSR_wake results (CSR.py:454-602), other work split ------
 * Without chirp band (|slope0| <= 1, CSR.py:480) the (x', s') quadrature nodes of an observation point depend on its s
 * only, and the mesh of get_CSR_mesh is a tensor grid (CSR.py:380-389): all points of one z row share their nodes.  A
 * GROUP is up to plan.group_points (32) mesh points with the same z index and consecutive x indices (global group
 * g = iz * ceil(nx / group_points) + gx);
 * the kernel gives every point of a group one warp lane, walks the s' nodes in sequence and keeps the transverse-blended
 * corners of each lane's history cell in registers, so most samples need no history load at all (1.3-1.5x faster than
 * the point kernel on a bunch that fills its grid).  The summation order differs from dfcsr_wake_grid (results agree to
 * ~1e-15 relative, both within 1e-10 of the reference) and is a function of the plan only: any split of the groups over
 * launches / ranks gives bitwise the same grid.
 * dfcsr_wake_xgroup_plan: n_groups = 0 when the mapping does not apply to this step (chirp band, a sparse grid that
 * dfcsr_wake_uses_skipping would serve, fewer than 70 % of the lanes carrying a point, a bunch so compressed that the
 * points of a group look at history cells dozens of cells apart, integration zbins too large);
 * the plan depends on the history geometry, the beam scalars and the WHOLE mesh, never on the split. */
typedef struct dfcsr_xgroup_plan {
    int64_t n_groups;                  /* groups of the whole mesh; 0 = use dfcsr_wake_grid                     */
    int32_t unit_nodes;                /* x' nodes per partial sum (fixes the summation order)                  */
    int32_t max_units;                 /* partial sums per group                                                */
    int64_t workspace_bytes_per_group; /* d_workspace of a launch must hold group_count times this              */
    int32_t group_points;              /* mesh points per group (32: one per lane of a warp)                    */
    int32_t reserved;                  /* 0                                                                     */
} dfcsr_xgroup_plan;

int dfcsr_wake_xgroup_plan(const dfcsr_history* hist, const dfcsr_wake_params* wp, dfcsr_axis x_axis, dfcsr_axis z_axis,
                           dfcsr_xgroup_plan* plan);

/* Groups group_first + k * group_stride, k in [0, group_count).  d_dE / d_kick are FULL (x_axis.n * z_axis.n) arrays
 * indexed by mesh point (only the points of the launched groups are written; may be NULL when peers are given);
 * h_peer_grids / n_peers as in dfcsr_wake_grid_peers (NULL / 0: none).  d_workspace: caller-owned scratch (queue
 * counters, cleared by the launch, and partial sums); it must not be shared by launches that may run concurrently.
 * d_counters as in dfcsr_wake_mesh. */
int dfcsr_wake_grid_xgroups(const dfcsr_history* hist, const dfcsr_lattice* lat, const dfcsr_wake_params* wp,
                            dfcsr_axis x_axis, dfcsr_axis z_axis, double slope, double intercept,
                            int64_t group_first, int64_t group_count, int64_t group_stride, double* d_dE, double* d_kick,
                            const uint64_t* h_peer_grids, int32_t n_peers, void* d_workspace, int64_t workspace_bytes,
                            unsigned long long* d_counters, void* stream);

/* Loads the code of the wake kernels now instead of at their first launch (CUDA loads kernels lazily; the two large ones
 * cost tens of milliseconds, which would otherwise land in the first lattice step).  Needs a current CUDA context. */
int dfcsr_wake_preload(void);

/* 1 if the wake launches above would use zero-density skipping for this history and these beam scalars (row-support
 * table present and the grid sparse by construction: |slope0| > 1, the chirp-band branch of CSR.py:480, or a history
 * grid more than 1.5x the +-5 sigma box of the current bunch; wp->skip_mode ON / OFF overrides), else 0; < 0 on error. */
int dfcsr_wake_uses_skipping(const dfcsr_history* hist, const dfcsr_wake_params* wp);

/* get_CSR_wake(s, x, debug=True) (CSR.py:571-572, 599-600): integrands of one point.
 * d_iz / d_ix receive the regions back to back, each (n_x, n_s) row-major like the reference's
 * CSR_integrand_z/x arrays; h_regions[4][6] = {x_lo, x_hi, n_x, s_lo, s_hi, n_s}.  Synchronises. */
int dfcsr_wake_point_debug(const dfcsr_history* hist, const dfcsr_lattice* lat, const dfcsr_wake_params* wp,
                           double s, double x, double* d_iz, double* d_ix, int64_t capacity,
                           double* h_regions, int32_t* h_n_regions, void* stream);

/* ---- A13 / K5 kick application (beams.py:108-131) ----------------------------------------------
 * pz += bilinear(step*dE*1e6/E0)(x_T, z); px += bilinear(step*kick*1e6/E0)(x_T, z) if transverse_on,
 * with x_T = x - (slope*z + intercept), zero outside the mesh.  In place on d_px / d_pz. */
int dfcsr_apply_kick(const double* d_x, const double* d_z, double* d_px, double* d_pz, int64_t n,
                     double slope, double intercept,
                     const double* d_dE, const double* d_kick, dfcsr_axis x_axis, dfcsr_axis z_axis,
                     double step_size, double init_energy, int32_t transverse_on, void* stream);

/* ---- particle transport through one lattice element (beams.py:101-106: track_element(particle, element)) ----------
 * In place on the six coordinate arrays (Bmad-X canonical x, px, y, py, z, pz).  The element types are the ones
 * CSR2D.get_bmadx_element builds (CSR.py:146-199); the maps restate Bmad's: exact drift, sector bend with
 * "linear_edge" hard-edge kicks at the selected ends and the exact body solution, thick quadrupole (quad_mat2_calc +
 * low_energy_z_correction), sextupole as drift-kick-drift (csrc/track.cu).  p0c, mc2: reference momentum and rest
 * energy [eV].  Reproduces the reference's Bmad-X known answers (test/test_BmadX_tracking.ipynb cells 25, 28, 31). */
typedef enum { DFCSR_ELEM_DRIFT = 0, DFCSR_ELEM_SBEND = 1, DFCSR_ELEM_QUADRUPOLE = 2, DFCSR_ELEM_SEXTUPOLE = 3 } dfcsr_element_kind;
typedef struct dfcsr_element {
    int32_t kind;              /* dfcsr_element_kind                                              */
    int32_t fringe_entrance;   /* sbend: apply the entrance edge kick (FRINGE_AT both_ends / entrance_end) */
    int32_t fringe_exit;       /* sbend: apply the exit edge kick (FRINGE_AT both_ends / exit_end)         */
    int32_t n_step;            /* quadrupole: NUM_STEPS (Bmad-X default 1)                        */
    double L;                  /* length [m]                                                      */
    double g, e1, e2;          /* sbend: curvature G [1/m], pole-face angles E1, E2 [rad]         */
    double k1, k2;             /* quadrupole K1 [1/m^2], sextupole K2 [1/m^3]                     */
} dfcsr_element;
int dfcsr_track_element(double* d_x, double* d_px, double* d_y, double* d_py, double* d_z, double* d_pz,
                        int64_t n, const dfcsr_element* el, double p0c, double mc2, void* stream);

/* ---- linear transfer map -------------------------------------------------------
 * v <- M v for every particle, v = (x, px, y, py, z, pz), in place.  h_matrix: 36 HOST doubles, row-major.
 * First-order option (tracking order "first"; the default is dfcsr_track_element): drift, sector bend with
 * pole-face rotations, thick quadrupole (pydfcsr_b200/tracking.py builds the matrices). */
int dfcsr_track_linear(double* d_x, double* d_px, double* d_y, double* d_py, double* d_z, double* d_pz,
                       int64_t n, const double* h_matrix, void* stream);

/* ---- 2-D Savitzky-Golay operator (SGolay_filter.py:3-81) ---------------------
 * sgolay2d(z, window_size, order, derivative): d_z (rows x cols, row-major) is extended by window/2
 * samples per side with the reference's reflection rule (SGolay_filter.py:36-65) and convolved
 * ('valid', i.e. kernel flipped like scipy.signal.fftconvolve) with n_kernels window x window kernels
 * d_kernels[n_kernels][window][window] — the arrays the reference passes to fftconvolve: pinv(A)[0] for
 * smoothing, -pinv(A)[1] ('col'), -pinv(A)[2] ('row').  d_out[n_kernels][rows][cols].
 * window odd, <= 25, rows and cols >= window, n_kernels 1..3.  Not called by CSR2D.run (it is dead code in
 * the reference's run loop too, deposit.py:187,194,227,232). */
int dfcsr_sgolay2d(const double* d_z, int32_t rows, int32_t cols, int32_t window,
                   const double* d_kernels, int32_t n_kernels, double* d_out, void* stream);

/* ---- diagnostics -------------------------------------------------------------------------------------
 * K4 derives r = sqrt(r2) and 1/r from ONE reciprocal-square-root seed (wake.cu: sqrt_pair_fast); the
 * retarded time t - r must be the correctly rounded square root np.sqrt returns (CSR.py:647).  This entry
 * compares that routine bit for bit with the CUDA library's sqrt.rn.f64 and rsqrt on n pseudo-random
 * doubles with all 52 mantissa bits random and binary exponents uniform in [lo_exp, hi_exp):
 * h_mismatch[0] = sqrt results that differ, h_mismatch[1] = reciprocal results that differ.  Synchronises. */
int dfcsr_selftest_sqrt(int64_t n, uint64_t seed, double lo_exp, double hi_exp, uint64_t* h_mismatch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFCSR_B200_H */
