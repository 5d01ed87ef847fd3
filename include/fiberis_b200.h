/*
 * fiberis_b200 — C ABI of the B200-native 1D pressure-diffusion hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json's north_star names: the part of
 * fibeRIS's `fiberis.simulator` that assembles the implicit (backward-Euler) tridiagonal system,
 * solves it once per time step and collects the snapshots.  The reference has no FFI for this path
 * (it is pure Python); the entry points below are what a ctypes binding added to the reference's
 * own modules would call (see INTEGRATION.md for that binding).  Reference paths are relative to
 * /root/reference/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`; the caller owns all
 *     memory (the library allocates nothing that outlives a call);
 *   - arrays of batched systems are model-major: element (m, i) of an (M, nx) array is at
 *     m * nx + i.  A "shared" per-cell array (one mesh for all models) is passed with stride 0;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and returns an fbr_status; FBR_OK only means "enqueued" — device-side findings
 *     (zero pivots, Jacobi non-convergence) are written to the status arrays the call documents;
 *   - on error the thread-local message is available from fbr_last_error(); the library never
 *     aborts or exits;
 *   - boundary-condition codes: 0 = 'None', 1 = 'Dirichlet', 2 = 'Neumann'
 *     (src/fiberis/simulator/core/pds.py:137).
 */
#ifndef FIBERIS_B200_H
#define FIBERIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBR_VERSION 201 /* 0.2.1: the fp32 windowed kernel; 0.2.0: the options struct of the run entries */

typedef enum {
    FBR_OK = 0,
    FBR_ERR_BAD_ARG = -1,      /* -> ValueError on the Python side */
    FBR_ERR_UNSUPPORTED = -2,  /* shape outside what the selected kernel variant covers */
    FBR_ERR_CUDA = -3          /* a CUDA runtime call failed; message has the cudaError string */
} fbr_status;

enum { FBR_BC_NONE = 0, FBR_BC_DIRICHLET = 1, FBR_BC_NEUMANN = 2 };

/* Solver variants for fbr_solve_tridiag_* (north_star: "variant chosen by mesh length"). */
enum {
    FBR_SOLVER_AUTO = 0,
    FBR_SOLVER_THOMAS = 1, /* one thread per system, tiles transposed in shared memory so that
                              global traffic is coalesced (many small systems)        */
    FBR_SOLVER_PCR = 2,    /* one CTA per system, shared-memory parallel cyclic reduction */
    FBR_SOLVER_REG = 3,    /* register-resident: a thread owns 8 consecutive cells, pivots from a scan of
                              2x2 Moebius matrices, the two sweeps as warp scans; one pass over HBM
                              (40 B per unknown), no work array; 1 <= nx <= 4096 (AUTO's choice there) */
    FBR_SOLVER_PIVOT = 4   /* partial pivoting as in the reference's LAPACK solve (the DGTSV algorithm, one thread
                              per system): the fallback for regular matrices whose un-pivoted elimination meets
                              a zero pivot; work: 4 * M * nx doubles */
};

int fbr_version(void);
const char *fbr_last_error(void);

/* dst (cols, rows) = transpose of src (rows, cols), both row-major fp64.  Turns the (time, cell) rows written by the
 * long-mesh run kernels into the (cell, time) array Data2D / pack_result hold (src/fiberis/simulator/core/pds.py:415,
 * src/fiberis/analyzer/Data2D/core2D.py:27) without a host-side transpose; the on-chip kernels write that layout
 * directly (fbr_run_opts.snap_cell_stride). */
int fbr_transpose_f64(const double *src, int64_t rows, int64_t cols, double *dst, void *stream);

/* The path's only collective: all-gather of the ranks' per-member misfit blocks (`count` doubles each, equal
 * on every rank: pad the last block) with a communicator the CALLER created (ncclComm_t passed as void*), asynchronous on
 * `stream`.  The library links no NCCL: ncclAllGather is resolved at first use from the NCCL loaded in the process
 * (FBR_ERR_UNSUPPORTED when there is none).  PDS1D_Ensemble itself gathers through torch.distributed (NCCL backend). */
int fbr_nccl_allgather_f64(const double *send, double *recv, int64_t count, void *nccl_comm, void *stream);

/* Measurement aid: the fp64 issue peak of the current device in lane-level DFMA per second (8 independent chains per
 * thread, 16 warps per SM, timed with CUDA events on `stream`; synchronous).  bench.py's compute-side roofline divides
 * the fp64 operations K3 executes by this figure.  h_dfma_per_s: host; d_sink: 8 bytes of device scratch. */
int fbr_probe_fp64_peak(double *h_dfma_per_s, double *d_sink, void *stream);

/* Device facts the host layer sizes launches with.  Any out pointer may be NULL. */
int fbr_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *smem_per_block_optin,
                    int64_t *l2_bytes, int64_t *hbm_bytes);

/* ------------------------------------------------------------------------------------------
 * K1  fused coefficient assembly.
 * Replaces matrix_builder_1d_single_source / _multi_source
 * (src/fiberis/simulator/solver/matbuilder.py:6-83, 86-167) and build_pml_sigma (:170-196),
 * producing the three diagonals of the reference's dense A instead of the dense matrix:
 * harmonic-mean interface diffusivity, non-uniform 3-point stencil, boundary rows, source rows
 * (which win over boundary rows), PML added to every diagonal entry.  Bit-identical to the NumPy
 * arithmetic of the reference (same operation order, IEEE division, no FMA contraction).
 *   mesh, diffusivity : (M, nx) with stride nx, or one shared (nx) row with stride 0
 *   src_idx           : (n_src) int64 DEVICE array of source rows (may be NULL when n_src == 0)
 *   dl, d, du         : (M, nx) out; dl[m,0] = du[m,nx-1] = 0
 * ------------------------------------------------------------------------------------------ */
int fbr_assemble_f64(const double *mesh, int64_t mesh_stride, const double *diffusivity,
                     int64_t diff_stride, int64_t M, int64_t nx, double dt, int lbc, int rbc,
                     const int64_t *src_idx, int n_src, double pml_thickness, double sigma_max,
                     double *dl, double *d, double *du, void *stream);

/* Zonal ensembles: D_m(x) = zone_value[m, zone_of_cell[x]],
 * gathered on the fly (the (M, nx) diffusivity array is never materialised); mesh is shared.
 * zone_value: (M, n_zones), e.g. 10 ** theta computed by the host; zone_of_cell: (nx) int32. */
int fbr_assemble_zonal_f64(const double *mesh, const double *zone_value, int n_zones,
                           const int32_t *zone_of_cell, int64_t M, int64_t nx, double dt, int lbc,
                           int rbc, const int64_t *src_idx, int n_src, double pml_thickness,
                           double sigma_max, double *dl, double *d, double *du, void *stream);

/* PML ramp alone: sigma (nx) = build_pml_sigma(mesh, pml_thickness, sigma_max)
 * (src/fiberis/simulator/solver/matbuilder.py:170-196); all zeros when pml_thickness == 0. */
int fbr_pml_sigma_f64(const double *mesh, int64_t nx, double pml_thickness, double sigma_max,
                      double *sigma, void *stream);

/* ------------------------------------------------------------------------------------------
 * K1b time-invariant factorisation.  A is constant for a fixed dt (only b changes,
 * src/fiberis/simulator/core/pds.py:287-310), so the elimination is done once:
 *   u_0 = d_0, r_i = 1 / u_i, l_i = dl_i r_{i-1}, u_i = d_i - l_i du_{i-1}
 *   fl = l, fr = r, fg = du r                  (each (M, nx); one division per cell)
 * work: NULL, or fbr_factorize_work_bytes(M, nx) bytes of scratch; with it, long systems
 * (nx >= 512 and nx > 8 M) are factorised in parallel along the mesh (scan of 2x2 Moebius
 * matrices for the pivot in front of every 8-cell chunk, then the plain recurrence inside the
 * chunk); without it one thread walks each model.
 * singular[m] (int32, may be NULL) receives 1 + the first row with a zero/non-finite pivot, or 0.
 * The Python layer maps a non-zero flag to numpy.linalg.LinAlgError, the exception
 * scipy.linalg.solve raises in the reference (src/fiberis/simulator/solver/PDESolver_IMP.py:28).
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_factorize_work_bytes(int64_t M, int64_t nx);
int fbr_factorize_f64(const double *dl, const double *d, const double *du, int64_t M, int64_t nx,
                      double *fl, double *fr, double *fg, int32_t *singular, double *work,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * K3  persistent multi-step integration — replaces the time loop
 * `while taxis[-1] < t_total` of PDS1D_*.solve (src/fiberis/simulator/core/pds.py:287-310)
 * for a fixed dt.  Pressure and the factors stay on chip for all n_steps; per step
 *     b = p^n ; b[0] = 0 unless lbc == None ; b[nx-1] = 0 unless rbc == None ;
 *     b[src_idx[k]] = src_table[n, k]  (k ascending: later duplicates win) ; p^{n+1} = A^-1 b
 * (src/fiberis/simulator/solver/matbuilder.py:45-76).
 *   p0          : (M, nx) with stride nx or shared (nx) with p0_stride 0
 *   src_table   : (n_steps, n_src) value of source k at the START of step n (host: np.interp)
 *   snap        : NULL, or the buffer that receives the state after steps snapshot_every,
 *                 2*snapshot_every, ... of the models listed in snap_models (int64 (n_snap_models),
 *                 no repeated entries; NULL = all M models, n_snap_models must then be M).
 *                 Recorded models cost about twice as much as the others; short lists (<= 16) are
 *                 integrated first so that they never form the tail of the launch.  Sample r of recorded model
 *                 q, cell i goes to snap[r * snap_row_stride + q * snap_model_stride + i] (strides
 *                 in elements), so the caller chooses (time, model, cell) or (model, time, cell)
 *                 order and can reserve a leading row for the initial state by offsetting snap.
 *   p_final     : NULL or (M, nx): state after the last step
 *   obs_idx/obs/obs_w/misfit : fused objective (NULL obs_idx disables it).  Without opts->n_samples the
 *                 observations sit on the step times:
 *                 misfit[m] = sum_{n=1..n_steps} sum_k obs_w[k] (p^n[m, obs_idx[k]] - obs[n, k])^2
 *                 obs: (n_steps + 1, n_obs), row 0 unused; obs_w may be NULL (all ones).
 *                 With opts->n_samples > 0 the observations have their OWN time axis (see fbr_run_opts).
 *   opts        : NULL (all defaults) or a pointer to a struct fbr_run_opts in host memory, read during the call
 * The _f32 variant keeps state and factors in float (fp32 mode of north_star, gated at 1e-5).
 * Step-at-a-time use (n_steps = 1, snap = NULL, no observations, p_final given, 256 < nx <= 4096, p0_stride = nx): served
 * by the register-resident solve kernel with the factors as its rows (csrc/fbr_solve_reg.cu, FACTORS mode) — the next
 * member's rows are prefetched through shared memory: 40 B per cell-update at 0.93 of the measured HBM peak.
 * Limits of this on-chip variant: nx <= fbr_run_max_nx(), n_src <= 253, n_obs <= 64.
 * ------------------------------------------------------------------------------------------ */

/* Options of fbr_run_f64 / fbr_run_f32 / fbr_run_zonal_f64.  A zero-filled struct (with struct_bytes set) means
 * "all defaults"; the library holds no state between calls.
 *
 * Observations on their own time axis — the misfit of the reference's history-matching workflow
 * (src/fiberis/moose/templates/baseline_model_generator.py:394-403 with Data1D.interpolate,
 * src/fiberis/analyzer/Data1D/core1D.py:732-822): the simulated channel is interpolated linearly onto the
 * observed times, the sample at index `base_sample` is subtracted ("initial time calibration", :398), and the
 * squared residuals of samples [first_sample, last_sample) are summed (:401, the reference uses [1:-1]) with one
 * weight per channel (:403).  The host brackets every observed time t_j between two stored states once:
 *     sample_row[j] = n  with taxis[n] <= t_j <= taxis[n + 1]   (-1: t_j outside the simulated range -> NaN,
 *                                                                the reference's np.interp(left=nan, right=nan))
 *     sample_w[j]   = (t_j - taxis[n]) / (taxis[n + 1] - taxis[n])   (0 when t_j == taxis[n]: row n + 1 is not read)
 *     s_j[k] = p^n[k] + sample_w[j] (p^{n+1}[k] - p^n[k]),  p^0 = initial state
 *     misfit[m] = sum_k obs_w[k] sum_{j = first_sample}^{last_sample - 1} ((s_j[k] - s_base[k]) - obs[j, k])^2
 * obs is then (n_samples, n_obs).  row_first_sample: (n_steps + 2) int32, row_first_sample[r] = the first sample
 * whose last needed state (row n + 1, or n when sample_w == 0) is >= r — samples must be ordered in time.
 * Requires base_sample <= first_sample (or base_sample < 0: nothing is subtracted).  Sampled observations are
 * evaluated by the 64-byte-per-thread kernel family only (fp64: any nx <= fbr_run_max_nx(); the call picks it). */
typedef struct fbr_run_opts {
    int32_t struct_bytes;            /* sizeof(fbr_run_opts) of the caller */
    int32_t reserve_ctas;            /* leave this many CTA slots of the persistent grid free: a SMALL fbr_run_* launch
                                        enqueued just before on another stream (the few recorded members of a sweep, whose
                                        snapshots travel to the host while the sweep runs) then stays resident beside it */
    int64_t snap_cell_stride;        /* 0 or 1: the cells of a snapshot row are contiguous.  Otherwise sample r of slot q,
                                        cell i goes to snap[r * snap_row_stride + q * snap_model_stride + i * snap_cell_stride]:
                                        snap_row_stride = 1, snap_cell_stride = n_rows writes the (cell, time) layout of
                                        Data2D / pack_result directly (src/fiberis/simulator/core/pds.py:415) */
    double scan_log2_eps;            /* far scan stages / cross-warp carry chains of a member are skipped when every
                                        multiplier product they carry is below 2^scan_log2_eps in magnitude (decided per
                                        member from its own factors); 0 = default (-80), > 0 = never skip (exact scans) */
    int64_t n_samples;               /* 0: observations on the step times (see above) */
    const int32_t *sample_row;       /* device (n_samples) */
    const double *sample_w;          /* device (n_samples) */
    const int32_t *row_first_sample; /* device (n_steps + 2) */
    int32_t base_sample;             /* < 0: no baseline subtraction */
    int32_t first_sample, last_sample;
    int32_t snap_f64;                /* fp32 variants only, M <= 64, nx >= 512: `snap` points to DOUBLES (same strides, in
                                        elements) and the recorded rows are widened when they are stored — the few audited
                                        members of an fp32 sweep then need no widening pass, which would have to squeeze
                                        past the sweep's persistent grid (was the reserved word: 0 = off) */
    int32_t *scan_mode;              /* NULL, or device (M) int32: receives how each member's sweeps were scanned
                                        (0 exact, 1 neighbour-warp carry, 2 neighbour-warp carry + 4 scan stages);
                                        written by the 64-byte-per-thread kernels only */
    const int32_t *member_order;     /* NULL, or device (M) int32: a permutation of 0 .. M-1, the order in which the
                                        persistent CTAs of a sweep that records nothing take the members (they are handed
                                        out on demand).  Most expensive first — e.g. by decreasing largest diffusivity: long
                                        decay horizons take the slower scans — keeps the costly members out of the last
                                        round (C3: 19.5 -> see DESIGN.md).  A scheduling hint only: results do not depend on
                                        it, kernels without an on-demand schedule ignore it */
} fbr_run_opts;

/* Order for fbr_run_opts.member_order: members ranked by decreasing cost figure (e.g. the largest diffusivity of the
 * member), in half-octave classes — a counting sort in one CTA, a few microseconds for the ensembles it matters for.
 * cost: device (M) fp64; order: device (M) int32 out, a permutation of 0 .. M-1 (no particular order inside a class;
 * non-positive / NaN figures come last).  No reference counterpart: scheduling of the sweep (north_star). */
int fbr_rank_members_f64(const double *cost, int64_t M, int32_t *order, void *stream);

int fbr_run_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                const int64_t *src_idx, int n_src, const double *src_table,
                int64_t snapshot_every, const int64_t *snap_models, int64_t n_snap_models,
                int64_t snap_row_stride, int64_t snap_model_stride, double *snap, double *p_final, const int64_t *obs_idx, int n_obs,
                const double *obs, const double *obs_w, double *misfit, const fbr_run_opts *opts, void *stream);

int fbr_run_f32(const double *fl, const double *fr, const double *fg, const double *p0,
                int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                const int64_t *src_idx, int n_src, const double *src_table,
                int64_t snapshot_every, const int64_t *snap_models, int64_t n_snap_models,
                int64_t snap_row_stride, int64_t snap_model_stride, float *snap, float *p_final, const int64_t *obs_idx, int n_obs,
                const double *obs, const double *obs_w, double *misfit, const fbr_run_opts *opts, void *stream);

/* K3 with ON-CHIP ASSEMBLY: fbr_assemble_zonal_f64 + fbr_factorize_f64 + fbr_run_f64 in one kernel, for
 * ensembles whose members share the mesh and differ in n_zones diffusivity values (BASELINE.json
 * configs[2], [3]).  Each CTA assembles its member's rows with K1's arithmetic (bit-identical diagonals,
 * src/fiberis/simulator/solver/matbuilder.py:21-81), factorises them in registers (scan of 2x2 Moebius
 * matrices, csrc/fbr_factor_reg.cuh) and integrates all steps: no (M, nx) coefficient array ever exists in
 * HBM.  mesh: (nx); zone_value: (M, n_zones); zone_of_cell: (nx) int32, or NULL with n_zones == nx for
 * members that carry a full (nx) diffusivity row each; sigma: NULL or (nx) PML damping
 * (fbr_pml_sigma_f64); singular: NULL or (M) int32 PRESET TO INT32_MAX by the caller, receives
 * min(1 + zero-pivot row) for members whose matrix is singular.  Remaining arguments as fbr_run_f64.
 * 128 < nx <= fbr_run_max_nx(). */
int fbr_run_zonal_f64(const double *mesh, const double *zone_value, int n_zones,
                      const int32_t *zone_of_cell, double dt, const double *sigma, int32_t *singular,
                      const double *p0, int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc,
                      int rbc, const int64_t *src_idx, int n_src, const double *src_table,
                      int64_t snapshot_every, const int64_t *snap_models, int64_t n_snap_models,
                      int64_t snap_row_stride, int64_t snap_model_stride, double *snap, double *p_final,
                      const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                      double *misfit, const fbr_run_opts *opts, void *stream);

/* The same with state and factors kept in float (fp32 mode): the matrices are still assembled and factorised in
 * fp64 on chip and rounded once.  512 <= nx <= fbr_run_max_nx(). */
int fbr_run_zonal_f32(const double *mesh, const double *zone_value, int n_zones,
                      const int32_t *zone_of_cell, double dt, const double *sigma, int32_t *singular,
                      const double *p0, int64_t p0_stride, int64_t M, int64_t nx, int64_t n_steps, int lbc,
                      int rbc, const int64_t *src_idx, int n_src, const double *src_table,
                      int64_t snapshot_every, const int64_t *snap_models, int64_t n_snap_models,
                      int64_t snap_row_stride, int64_t snap_model_stride, float *snap, float *p_final,
                      const int64_t *obs_idx, int n_obs, const double *obs, const double *obs_w,
                      double *misfit, const fbr_run_opts *opts, void *stream);

int64_t fbr_run_max_nx(void);

/* ------------------------------------------------------------------------------------------
 * K3L the same persistent integration for LONG meshes (fbr_run_max_nx() < nx <=
 * fbr_run_long_max_nx(), about 6e5 cells on a B200): the mesh is tiled over co-resident CTAs of
 * one cooperative launch, every cell's pressure and factors stay in registers for all steps, tiles
 * exchange carries through L2 with one grid barrier per sweep.  Models (M, usually 1) are run one
 * after the other.  p0: (M, nx).  snap: sample r of model m, cell i at
 * snap[r * snap_row_stride + m * snap_model_stride + i].  work: fbr_run_long_work_bytes(nx) bytes.
 * No fused misfit on this path.
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_run_long_max_nx(void);
int64_t fbr_run_long_work_bytes(int64_t nx);
int fbr_run_long_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                     int64_t M, int64_t nx, int64_t n_steps, int lbc, int rbc,
                     const int64_t *src_idx, int n_src, const double *src_table,
                     int64_t snapshot_every, int64_t snap_row_stride, int64_t snap_model_stride,
                     double *snap, double *p_final, double *work, void *stream);

/* ------------------------------------------------------------------------------------------
 * K3S the same integration for ONE model whose mesh does not fit on chip (nx >
 * fbr_run_long_max_nx(); BASELINE.json configs[4], 1e7 cells).  State and factors stream from HBM
 * every step through a hierarchy of chunked linear recurrences (thread per 32-cell chunk,
 * chunk-interleaved layout, coalesced), see csrc/fbr_run_stream.cu.  Same argument meaning as
 * fbr_run_long_f64 with M = 1; work: fbr_run_stream_work_bytes(nx) bytes.  Launches O(n_steps)
 * kernels on `stream`.
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_run_stream_work_bytes(int64_t nx);
int fbr_run_stream_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                       int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx,
                       int n_src, const double *src_table, int64_t snapshot_every,
                       int64_t snap_row_stride, double *snap, double *p_final, double *work,
                       void *stream);

/* ------------------------------------------------------------------------------------------
 * K3P the exact long-mesh integration in ONE pass over HBM per step (40 B per cell-update: three factors and p^n
 * in, p^{n+1} out; csrc/fbr_run_pass.cu) — any nx.  Tiles of 512 .. 2048 cells are solved in registers; the forward carry
 * between tiles comes from aggregates the previous launch left behind, the backward carry from a decoupled look-back
 * inside the launch.  Same arguments as fbr_run_stream_f64; work: fbr_run_pass_work_bytes(nx) bytes; one launch per
 * time step plus one setup launch.  Takes the place of K3S behind PDS1D_*.solve for meshes whose decay horizon is too
 * long for the windowed kernel.
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_run_pass_work_bytes(int64_t nx);
int fbr_run_pass_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                     int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                     const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                     double *snap, double *p_final, double *work, void *stream);
/* The same with the tile width chosen by the caller: tile_warps = 2, 4 or 8 warps of 256 cells (0 = 8, what
 * fbr_run_pass_f64 uses).  What couples the tiles costs per tile, a tile's own latency grows with its width: narrow
 * tiles (2) when the decay horizon of the factors (fbr_decay_horizon_f64) is below ~100 cells — the coupling then ends
 * at the neighbouring tile: 1e7 cells in 89 us per step, 4.5 TB/s —, wide tiles (8) on stiff meshes, where the carries
 * fold over tens of tiles.  Results do not depend on the width beyond rounding. */
int fbr_run_pass_tiled_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                           int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx, int n_src,
                           const double *src_table, int64_t snapshot_every, int64_t snap_row_stride,
                           double *snap, double *p_final, int tile_warps, double *work, void *stream);

/* ------------------------------------------------------------------------------------------
 * K3W the same integration for ONE long mesh through overlapping on-chip WINDOWS with time blocking
 * (csrc/fbr_run_window.cu).  The sweeps' multipliers decay geometrically, so a window of 2048 / 4096
 * cells can be advanced s steps on chip and still be exact (to a 2^log2_eps truncation, default
 * 2^-80, far below fp64 rounding) on all but s*horizon cells at each end; windows overlap by those
 * halos.  HBM traffic per cell-update is (32 B * window/owned + 8 B) / s instead of 40 B.
 *
 * fbr_decay_horizon_f64 measures the horizons of a factorisation: horizon[0] (forward, from fl) and
 * horizon[1] (backward, from fg) = the smallest H such that EVERY run of >= H consecutive
 * multipliers has a product below 2^log2_eps (single multipliers may exceed one — the row after a
 * source row, mesh-spacing jumps — which the kernel bounds rigorously); cap + 1 if that cannot be
 * established within cap cells.  horizon: DEVICE int64[2]; work: DEVICE double[4].
 * fbr_run_window_plan reports what the driver will use for a mesh (plan4 = warps per window CTA,
 * steps per block, owned cells per window, windows) or FBR_ERR_UNSUPPORTED when the horizons are too
 * long for windows to pay — callers then use the exact K3L / K3S paths.  HOST pointer plan4.
 * fbr_run_window_f64: arguments as fbr_run_stream_f64 plus the horizons; work:
 * fbr_run_window_work_bytes(nx) bytes (two state buffers).  Blocks never straddle a snapshot row;
 * snapshot rows are written in place and serve as the next block's input.  snapshot_every >= 1 caps
 * the block length (and fixes the window geometry) even when snap is NULL — a caller that streams
 * rows out itself, one call per interval with p_final = the row, gets bit-identical results;
 * snapshot_every = 0 (snap must be NULL) lifts the cap.
 * ------------------------------------------------------------------------------------------ */
int fbr_decay_horizon_f64(const double *fl, const double *fg, int64_t nx, double log2_eps, int cap,
                          int64_t *horizon, double *work, void *stream);
int fbr_run_window_plan(int64_t nx, int64_t horizon_fwd, int64_t horizon_bwd,
                        int64_t snapshot_every, int64_t *plan4);
int64_t fbr_run_window_work_bytes(int64_t nx);
int fbr_run_window_f64(const double *fl, const double *fr, const double *fg, const double *p0,
                       int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx,
                       int n_src, const double *src_table, int64_t snapshot_every,
                       int64_t snap_row_stride, double *snap, double *p_final,
                       int64_t horizon_fwd, int64_t horizon_bwd, double *work, void *stream);
/* fp32 mode of K3W (the reference has no single-precision path: this is the `dtype='float32'` option of this
 * package's solve(), to 1e-5 of the fp64 result).  Same arguments; factors and initial state are the fp64 arrays
 * of fbr_factorize_f64 and are narrowed once into `work`, 16-byte aligned,
 * fbr_run_window_f32_work_bytes(nx) bytes; snapshots and the final state are float.  A thread holds 16 cells, so
 * a window is 4096 / 8192 cells with the same halos: fbr_run_window_plan_f32 reports the geometry. */
int fbr_run_window_plan_f32(int64_t nx, int64_t horizon_fwd, int64_t horizon_bwd,
                            int64_t snapshot_every, int64_t *plan4);
int64_t fbr_run_window_f32_work_bytes(int64_t nx);
int fbr_run_window_f32(const double *fl, const double *fr, const double *fg, const double *p0,
                       int64_t nx, int64_t n_steps, int lbc, int rbc, const int64_t *src_idx,
                       int n_src, const double *src_table, int64_t snapshot_every,
                       int64_t snap_row_stride, float *snap, float *p_final,
                       int64_t horizon_fwd, int64_t horizon_bwd, void *work, void *stream);

/* Running count of kernels this thread has launched through the library (bench.py's
 * gpu_launches is a difference of two readings). */
int fbr_last_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K2  bare batched tridiagonal solve  dl, d, du, b -> x  (no reuse of factors).
 * Replaces solver_implicit(A, b, solver='numpy') = scipy.linalg.solve
 * (src/fiberis/simulator/solver/PDESolver_IMP.py:9-34) for tridiagonal A; used by the adaptive
 * time-step loop where A changes every iteration (src/fiberis/simulator/core/pds.py:312-336).
 * x may alias b.  singular as in fbr_factorize_f64.  work: (M, nx) scratch needed by the THOMAS
 * variant (fbr_solve_tridiag_work_bytes); may be NULL for PCR / REG.  PCR and REG cover 1 <= nx <= 4096;
 * AUTO = REG up to 4096 cells, THOMAS beyond (work must then be given).
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_solve_tridiag_work_bytes(int64_t M, int64_t nx, int variant);
int fbr_solve_tridiag_f64(const double *dl, const double *d, const double *du, const double *b,
                          double *x, int64_t M, int64_t nx, int variant, double *work,
                          int32_t *singular, void *stream);

/* A general dense system A x = b (API completeness of solver_implicit, src/fiberis/simulator/solver/PDESolver_IMP.py:9-34;
 * the simulator itself only builds tridiagonal matrices): Gaussian elimination with partial pivoting by one CTA.
 * Ab: device working copy of the augmented matrix [A | b], row-major n x (n + 1), destroyed; x: (n);
 * singular: NULL or one int32 that receives 1 + the column of an exactly zero pivot column (0: regular). */
int fbr_solve_dense_f64(double *Ab, int64_t n, double *x, int32_t *singular, void *stream);
/* Jacobi iteration on a general dense system (API completeness of solver_explicit,
 * src/fiberis/simulator/solver/PDESolver_EXP.py:5-39: x starts at 0, x_new = (b - R x) / diag(A) with
 * R = A - diag(A), stop when max|x_new - x| < tol).  A: row-major n x n; work: n doubles;
 * iters: iterations used, or -max_iter when the tolerance was not met (x = the last iterate). */
int fbr_jacobi_dense_f64(const double *A, const double *b, double *x, int64_t n, double tol, int max_iter,
                         int32_t *iters, double *work, void *stream);

/* One implicit step, batched (SURVEY 8(b) "fbr_step_*"): p_next = A^-1 rhs(p) with the diagonals of
 * A(dt) from fbr_assemble_*; rhs = p with the boundary / source overrides of
 * src/fiberis/simulator/solver/matbuilder.py:45-76, the solve of PDESolver_IMP.py:28.  Replaces the
 * pair matbuilder.*(pds, dt) / solver_implicit(A, b) of src/fiberis/simulator/core/pds.py:294-303 for
 * one step of M models.  src_val: DEVICE values of this step, (n_src) shared by all models
 * (src_val_stride = 0) or (M, n_src) rows (src_val_stride >= n_src).  p_next may alias p.  variant /
 * work / singular as in fbr_solve_tridiag_f64; up to 4096 cells the overrides are fused into the
 * register-resident solve (one kernel, 40 B of HBM traffic per cell-update). */
int fbr_step_f64(const double *dl, const double *d, const double *du, const double *p, int64_t M,
                 int64_t nx, int lbc, int rbc, const int64_t *src_idx, int n_src, const double *src_val,
                 int64_t src_val_stride, double *p_next, int variant, double *work, int32_t *singular,
                 void *stream);

/* ------------------------------------------------------------------------------------------
 * Device-side adaptive time stepping: the step-doubling loop of PDS1D_*.solve(optimizer=True)
 * (src/fiberis/simulator/core/pds.py:312-336) with the controller of
 * src/fiberis/simulator/optimizer/tso.py:4-50, one CTA per model, no host round trip per
 * iteration.  Sources are sampled on the device with np.interp semantics from their own series:
 * src_taxis / src_data: (n_src, src_width) rows padded to src_width, src_len: (n_src) int32.
 *   state : (M, nx) in/out current pressure;  ctrl: (M, 2) in/out = (t, dt)
 *   tol   : accept threshold; adj_tol: the tol adjust_dt uses (the reference never forwards the
 *           user's tol to adjust_dt, so callers pass 1e-3); safety, p_order, max_dt, min_dt as in tso
 *   snap_models / n_snap_models : the models whose accepted states are RECORDED (int64 list; NULL = all M,
 *           n_snap_models must then be M).  Sweeps record a few audited members only: the others then
 *           write nothing per step and are not stopped by `capacity`.
 *   snap  : (n_snap_models, capacity, nx) accepted states in order; taxis_out: (n_snap_models, capacity)
 *   counts: (M, 4) int64 out = accepted steps, iterations used, done (t >= t_total), 1 + zero-pivot row
 * A member returns when its `capacity` rows are filled (recorded members only), t >= t_total or max_iters
 * iterations ran; call again with the same state / ctrl to continue.  nx <= fbr_run_adaptive_max_nx();
 * no PML.  Implementation: a thread owns 8 cells; A(h) is factorised in registers (pivots from a warp
 * scan of 2x2 Moebius matrices, one division per cell) and applied with the run kernels' scan sweeps.
 * ------------------------------------------------------------------------------------------ */
int64_t fbr_run_adaptive_max_nx(void);
int fbr_run_adaptive_f64(const double *mesh, int64_t mesh_stride, const double *diffusivity,
                         int64_t diff_stride, double *state, int64_t M, int64_t nx, int lbc, int rbc,
                         const int64_t *src_idx, int n_src, const double *src_taxis,
                         const double *src_data, const int32_t *src_len, int src_width, double t0,
                         double t_total, double *ctrl, double tol, double adj_tol, double safety,
                         double p_order, double max_dt, double min_dt, int64_t max_iters,
                         int64_t capacity, double *snap, double *taxis_out,
                         const int64_t *snap_models, int64_t n_snap_models, int64_t *counts,
                         void *stream);

/* Right-hand side of one step, b = p with boundary / source overrides
 * (src/fiberis/simulator/solver/matbuilder.py:45-76).  src_val: (n_src) DEVICE values shared by
 * all models.  b may alias p. */
int fbr_rhs_f64(const double *p, int64_t M, int64_t nx, int lbc, int rbc, const int64_t *src_idx,
                int n_src, const double *src_val, double *b, void *stream);

/* Misfit of STORED rows against observations (SURVEY 8(b) "fbr_misfit_*"; the same quantity K3 sums
 * on the fly through fbr_run_*'s obs arguments — parity unpinned, modelled on
 * src/fiberis/moose/templates/baseline_model_generator.py:394-404):
 *   misfit[r] = sum_{t = first_row}^{n_rows - 1} sum_k obs_w[k] (snap[r, t, obs_idx[k]] - obs[t, k])^2
 * snap[r * snap_model_stride + t * snap_row_stride + i]; obs: (n_rows, n_obs); obs_w: (n_obs) or NULL. */
int fbr_misfit_f64(const double *snap, int64_t n_models, int64_t n_rows, int64_t nx,
                   int64_t snap_row_stride, int64_t snap_model_stride, const int64_t *obs_idx, int n_obs,
                   const double *obs, const double *obs_w, int64_t first_row, double *misfit,
                   void *stream);

/* Scratch sizes through one entry point (the caller owns all device memory): bytes of `work` the call
 * named by `what` needs for M models of nx cells (0 = none); < 0 on an unknown kind. */
enum {
    FBR_WS_FACTORIZE = 0, FBR_WS_SOLVE = 1, FBR_WS_SOLVE_THOMAS = 2, FBR_WS_RUN = 3, FBR_WS_RUN_LONG = 4,
    FBR_WS_RUN_STREAM = 5, FBR_WS_RUN_WINDOW = 6, FBR_WS_RUN_PASS = 7
};
int64_t fbr_workspace_bytes(int what, int64_t M, int64_t nx);

/* ------------------------------------------------------------------------------------------
 * K5  step-doubling error of the adaptive controller: err[m] = ||u1 - u2||_2 / ||u1||_2
 * (src/fiberis/simulator/optimizer/tso.py:18).  u1, u2: (M, nx); err: (M).
 * ------------------------------------------------------------------------------------------ */
int fbr_error_norms_f64(const double *u1, const double *u2, int64_t M, int64_t nx, double *err,
                        void *stream);

/* ------------------------------------------------------------------------------------------
 * "explicit" mode = Jacobi iteration from x = 0 on A x = b, stop when ||dx||_inf < tol or after
 * max_iter sweeps (src/fiberis/simulator/solver/PDESolver_EXP.py:5-39).  iters[m] (int32, may be
 * NULL) receives the sweeps used, or -max_iter when it did not converge (the reference is silent).
 * ------------------------------------------------------------------------------------------ */
int fbr_jacobi_f64(const double *dl, const double *d, const double *du, const double *b, double *x,
                   int64_t M, int64_t nx, double tol, int max_iter, int32_t *iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FIBERIS_B200_H */
