/*
 * asora_b200.h -- C ABI of libasora_b200.so: the B200-native (sm_100a) replacement for pyc2ray's
 * ASORA ray-tracing library and for the per-cell ionisation chemistry pass.
 *
 * Every entry point returns 0 on success and a non-zero code on failure; it never throws and never
 * aborts.  asora_last_error() returns a description of the most recent failure on the calling thread's
 * process (the reference throws C++ exceptions through extern "C" -> std::terminate:
 * src/asora/raytracing.cu:134-139, src/asora/memory.cu:70-75).
 *
 * One context per process, bound to the CUDA device that is current when asora_device_init() is
 * called (same ownership model as the reference's process-global pointers, src/asora/memory.cu:20-29).
 * Not thread-safe; calls are synchronous unless stated otherwise.
 *
 * Grid layout everywhere: float64, flat, index i*N*N + j*N + k of the logical cell (i,j,k)
 * (src/asora/raytracing.cu:30).  Sources: int32[3*NumSrc] interleaved x,y,z, 0-indexed, and
 * float64[NumSrc] fluxes in units of 1e48 photons/s (pyc2ray/utils/sourceutils.py:30-31).
 *
 * Paths cited below are relative to the reference checkout (phirling/pyc2ray).
 */
#ifndef ASORA_B200_H
#define ASORA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the six libasora entry points (src/asora/python_module.cu:153-161) ------------------------- */

/* replaces libasora.device_init(N, num_src_par): python_module.cu:73-82 -> memory.cu:34-80.
 * Allocates the density / ionised-fraction / rate grids for mesh size N on the current device.
 * num_src_par (the reference's source batch size) is accepted for signature compatibility; this
 * implementation keeps column densities on-chip (or in one L2-resident scratch grid) and does not
 * allocate N^3 doubles per in-flight source. */
int asora_device_init(int N, int num_src_par);

/* replaces libasora.device_close(): python_module.cu:87-92 -> memory.cu:119-129.  Frees everything,
 * including both photo tables (the reference leaks the thick table). */
int asora_device_close(void);

/* replaces libasora.density_to_device(ndens, N): python_module.cu:97-109 -> memory.cu:85-88. */
int asora_density_to_device(const double* ndens, int N);

/* replaces libasora.photo_table_to_device(thin, thick, NumTau): python_module.cu:114-128 ->
 * memory.cu:90-98.  NumTau is the number of doubles in each table.  Re-calling replaces the tables
 * (the reference leaks the previous allocation). */
int asora_photo_table_to_device(const double* thin_table, const double* thick_table, int NumTau);

/* replaces libasora.source_data_to_device(pos, flux, NumSrc): python_module.cu:133-148 ->
 * memory.cu:99-114. */
int asora_source_data_to_device(const int32_t* pos, const double* flux, int NumSrc);

/* replaces libasora.do_all_sources(R, coldensh_out, sig, dr, ndens, xh_av, phi_ion, NumSrc, m1,
 * minlogtau, dlogtau, NumTau): python_module.cu:21-68 -> raytracing.cu:79-148.
 * Copies xh_av (host, N^3) to the device, zeroes the rate grid, ray-traces the first NumSrc uploaded
 * sources out to radius R (cells), and copies the summed rates into phi_ion (host, N^3).  The
 * reference ignores its coldensh_out and ndens arguments (raytracing.cu:116), so they are not part
 * of this ABI.  NumTau has the reference's meaning: the upper clamp of the table index
 * (rates.cu:78-79); indices are additionally clamped to the uploaded table length. */
int asora_do_all_sources(double R, double sig, double dr, const double* xh_av, double* phi_ion,
                         int NumSrc, int N, double minlogtau, double dlogtau, int NumTau);

/* The two halves of asora_do_all_sources, for callers that put a collective between the sweep and the download (one
 * process per GPU, sources sharded over the ranks, pyc2ray/evolve.py:360-373,433-437): _begin copies xh_av to the
 * device (or, with xh_av == NULL, takes what ASORA_BUF_XH_AV already holds: a rank that received the grid from a peer
 * GPU instead of from its host) and queues the sweep of the first NumSrc uploaded sources on the context's stream; the caller then reduces
 * ASORA_BUF_PHI_ION over the ranks on that stream (asora_set_stream, asora_device_buffer); _end waits and copies the
 * rates to phi_ion, or only waits when phi_ion is NULL (a rank that does not need the grid on the host). */
int asora_do_all_sources_begin(double R, double sig, double dr, const double* xh_av, int NumSrc, int N,
                               double minlogtau, double dlogtau, int NumTau);
int asora_do_all_sources_end(double* phi_ion);

/* ---- photo-heating rates --------------------------------------------------------------
 * The reference computes phi_heat only in its CPU ray tracer (src/c2ray/photorates.f90:118,124,
 * src/c2ray/raytracing.f90:530,537; f2py arguments heat_thin_table, heat_thick_table, phi_heat of
 * libc2ray.raytracing.do_all_sources) and lists "add heating rate computation to ASORA" as a TODO
 * (pyc2ray/c2ray_base.py:424-426).  These entry points are what that extension of libasora would bind. */

/* Upload the heating tables (radiation/blackbody.py:79-85), same length as the photo tables already on the
 * device.  photo_table_to_device discards them again. */
int asora_heat_table_to_device(const double* heat_thin_table, const double* heat_thick_table, int NumTau);

/* Heating on/off for the device-resident sweeps (asora_raytrace_device): when on, ASORA_BUF_PHI_HEAT receives
 * the sum over sources of heat / nHI next to ASORA_BUF_PHI_ION.  Restrictions while heating is on: sweeps must zero
 * the rate grids (zero_phi != 0; accumulating on top of earlier rates is an error) and the single-source debug path
 * (asora_debug_single_source) is not available.  The thin-cell branch evaluates the heating table at tau_out, the
 * convention of the ASORA ionisation rate (rates.cu:37); the CPU ray tracer uses tau_in (photorates.f90:124), a
 * relative difference below 1e-7 (|tau_out - tau_in| <= 1e-7 in that branch). */
int asora_set_heating(int on);

/* asora_do_all_sources that also returns the photo-heating rates (host, N^3). */
int asora_do_all_sources_heat(double R, double sig, double dr, const double* xh_av, double* phi_ion,
                              double* phi_heat, int NumSrc, int N, double minlogtau, double dlogtau, int NumTau);

/* ---- chemistry half of the boundary (f2py libc2ray.chemistry.global_pass) ------------------------ */

/* replaces libc2ray.chemistry.global_pass(dt,ndens,temp,xh,xh_av,xh_intermed,phi_ion,bh00,albpow,
 * colh0,temph0,abu_c) -> conv_flag: src/c2ray/chemistry.f90:13-48 (+ :53-316).
 * All arrays are host float64[ncell] sharing one memory order (the update is cell-local).
 * xh_av and xh_intermed are updated in place.  *conv_flag receives the number of non-converged
 * cells.  xh, xh_av and xh_intermed may alias each other (pyc2ray/chemistry.py:85,91): all inputs
 * are read before any output is written; xh_av is stored first, xh_intermed last. */
int asora_global_pass(double dt, const double* ndens, const double* temp, const double* xh,
                      double* xh_av, double* xh_intermed, const double* phi_ion, double bh00,
                      double albpow, double colh0, double temph0, double abu_c, int64_t ncell,
                      int* conv_flag);

/* ---- device-resident variants (no host<->device traffic; used by the fused evolve loop) --------- */

/* Named device buffers owned by the context (all float64[N^3] unless noted). */
enum {
    ASORA_BUF_NDENS = 0,
    ASORA_BUF_XH_AV = 1,
    ASORA_BUF_PHI_ION = 2,
    ASORA_BUF_XH = 3,          /* ionised fraction at the start of the time step */
    ASORA_BUF_XH_INTERMED = 4, /* end-of-step ionised fraction of the current iteration */
    ASORA_BUF_TEMP = 5,
    ASORA_BUF_COLDENS = 6,     /* outgoing column density of the last debug sweep */
    ASORA_BUF_PHI_HEAT = 7,    /* photo-heating rates of the last sweep with heating on */
    ASORA_BUF_COUNT = 8
};

/* Device address of a named buffer (allocated on first use), or NULL on error.  The chemistry caches two
 * temperature-only factors per cell (chemistry.f90:257-262) for the contents of ASORA_BUF_TEMP; every library call that
 * can change that buffer (upload, copy, this function) drops the cache.  A caller that keeps the pointer returned for
 * ASORA_BUF_TEMP and writes through it later must call asora_invalidate_temperature() before the next chemistry pass. */
void* asora_device_buffer(int which);
int asora_invalidate_temperature(void);

/* Host -> device / device -> host copy of a named buffer (N^3 doubles). */
int asora_buffer_upload(int which, const double* host);
int asora_buffer_download(int which, double* host);

/* The same for host grids in Fortran order (index i + N*j + N*N*k), which is what pyc2ray's driver classes hold
 * (c2ray_test.py:167-169): the axis reversal to the device layout is done on the GPU instead of by a strided
 * host copy (0.3 s per 250^3 grid in numpy, evolve.py:142-143,240). */
int asora_buffer_upload_f(int which, const double* host_fortran);
int asora_buffer_download_f(int which, double* host_fortran);

/* Host -> device copy of cells [cell_offset, cell_offset + cell_count) of a named buffer; `host` points at the
 * first cell of the WHOLE host grid (the same offset is applied on both sides). */
int asora_buffer_upload_range(int which, const double* host, int64_t cell_offset, int64_t cell_count);

/* Device -> device copy between two named buffers (N^3 doubles), on the context's stream. */
int asora_buffer_copy(int dst, int src);

/* Peer-memory halo exchange of slab-decomposed multi-GPU runs (one process per GPU on one node; replaces the N^3 Reduce + Bcast
 * of evolve.py:433-437,480-497 together with asora_set_active_slab).  ipc_export writes the 64-byte CUDA IPC handle of grid
 * buffer `which`; a neighbouring rank passes it to ipc_open and gets a device pointer to that buffer on the exporting GPU
 * (valid until ipc_close or the exporter's device_close).  peer_halo then does, on this context's stream,
 *   buffer[which][cell_offset .. +cell_count) += peer_buf[cell_offset .. +cell_count)   (add != 0: rates a neighbour computed)
 *   buffer[which][cell_offset .. +cell_count)  = peer_buf[...]                          (add == 0: a neighbour's new xh_av)
 * reading the neighbour's memory over NVLink.  Ordering across ranks (the neighbour has finished writing, nobody overwrites what
 * is still being read) is the caller's: pyc2ray_b200.parallel.SlabHalo brackets the calls with the collectives of its loop. */
int asora_ipc_export(int which, unsigned char* handle64);
int asora_ipc_open(const unsigned char* handle64, void** dev_ptr);
int asora_ipc_close(void* dev_ptr);
int asora_peer_halo(int which, const void* peer_buf, int64_t cell_offset, int64_t cell_count, int add);

/* Ray-trace sources [src_begin, src_begin+src_count) of the uploaded list using the device-resident
 * NDENS and XH_AV buffers; rates are accumulated into PHI_ION, which is zeroed first when
 * zero_phi != 0.  Asynchronous on the context's stream; asora_sync() waits. */
int asora_raytrace_device(double R, double sig, double dr, int src_begin, int src_count,
                          double minlogtau, double dlogtau, int NumTau, int zero_phi);

/* Chemistry pass on the device-resident buffers (NDENS, TEMP, XH, XH_AV, XH_INTERMED, PHI_ION).
 * Outputs: conv_flag, sum(xh_intermed), sum(1 - xh_intermed) (pyc2ray/evolve.py:210-217), reduced
 * deterministically on the device.  Synchronous. */
int asora_global_pass_device(double dt, double bh00, double albpow, double colh0, double temph0,
                             double abu_c, int* conv_flag, double* sum_xh1, double* sum_xh0);

/* Slab-decomposed multi-GPU runs (pyc2ray_b200/evolve.py, decomposition="slab").  The planes i in
 * [x_begin, x_begin + x_count) (periodic in N) are the only ones the following sweeps of this rank can touch
 * (its sources plus the ray-tracing radius): the nHI pre-pass and the zeroing of PHI_ION are restricted to
 * them.  x_count >= N (or 0) restores the whole grid. */
int asora_set_active_slab(int x_begin, int x_count);

/* asora_global_pass_device restricted to cells [cell_offset, cell_offset + cell_count) of the flat grids
 * (one rank's own planes: cell_offset = x * N * N). */
int asora_global_pass_device_range(double dt, double bh00, double albpow, double colh0, double temph0,
                                   double abu_c, int64_t cell_offset, int64_t cell_count, int* conv_flag,
                                   double* sum_xh1, double* sum_xh0);

/* Wait for all work queued on the context's stream. */
int asora_sync(void);

/* Run all subsequent work of the context on a caller-owned CUDA stream (a cudaStream_t passed as an
 * opaque pointer; NULL restores the context's own stream).  Lets a caller order the sweep with its
 * own work -- e.g. an NCCL all-reduce of PHI_ION -- without host synchronisation, and time both with
 * events on one stream. */
int asora_set_stream(void* cuda_stream);

/* ---- diagnostics -------------------------------------------------------------------------------- */

/* Ray-trace ONE uploaded source and return its outgoing column density grid (host, N^3; cells the
 * sweep does not visit are 0) -- for column-density parity tests.  phi_ion (host, N^3) may be NULL. */
int asora_debug_single_source(double R, double sig, double dr, const double* xh_av, int src_index,
                              double minlogtau, double dlogtau, int NumTau, double* coldensh_out,
                              double* phi_ion);

/* Force the sweep variant: 0 = automatic, 1 = shared-memory level sweep (one CTA per source batch, one cell per
 * thread), 2 = grid-cooperative level sweep (whole GPU per source), 3 = mirror-image sweep (one plan entry and up to
 * eight octant images per thread; needs a mirror-symmetric cell set, i.e. q_max <= N/2 - 1 on even meshes),
 * 4 = large-radius sweep (one thread-block cluster per wedge of a source, level buffers in distributed shared memory).
 * Automatic: 3 or 1 while two levels of a source fit the shared memory of a CTA, else 4, else 2.
 * Returns non-zero for unknown values. */
int asora_set_sweep_variant(int variant);

/* Sphere-only sweeps.  The reference visits the whole octahedron(q_max) & cube although only cells inside
 * the R sphere receive a rate (raytracing.cu:315); a cell outside the sphere can only be upstream of cells
 * that are even further out, so its column density never reaches phi_ion.  With sphere_only != 0 those
 * cells are skipped: phi_ion is bit-for-bit what the full sweep produces, the work drops to the rated cells
 * (41 % fewer at R = 30, 48 % fewer at R = 10.76).  Default 0 (visit exactly the reference's cells);
 * asora_last_sweep_stats() then reports the rated cells as `updates`. */
int asora_set_sphere_only(int sphere_only);

/* Deterministic accumulation of the rates.  By default every rated (source, cell) pair adds its rate with one fp64
 * reduction at L2, so phi_ion depends on the arrival order at the 1e-16 level, like the reference's atomicAdd
 * (raytracing.cu:328).  With on != 0 every contribution is split exactly into two integers of a 128-bit fixed-point
 * number (scaled so that 2^17 of the largest possible contribution fit a cell and contributions down to 1e-11 of it keep
 * 53 significant bits) and added with two 64-bit integer reductions: integer sums are associative, so
 * phi_ion is bit-identical from run to run, for every launch shape, split and sweep variant, and for every sharding of
 * the sources that sums the per-rank grids in a fixed order.  Costs a second N^3 grid of 8 bytes per cell and a second
 * reduction per rated cell; the (k,i,j)-ordered z-face copies are not used.  Not available on the single-source debug
 * path.  Default 0. */
int asora_set_deterministic(int on);

/* The reference's -D GREY_NOTABLES build (src/asora/Makefile:9 comment, raytracing.cu:317-318, rates.cu:44-64; Fortran:
 * raytracing.f90:499-501, photorates.f90:13-57): with on != 0 the rate of a cell is the analytic grey-opacity expression
 * strength * 1e48 / Vfact * (exp(-tau_in) - exp(-tau_out)) (thin cells: (tau_out - tau_in) * exp(-tau_in)) instead of the
 * table lookups; the uploaded tables are ignored and need not exist.  A test option, as in the reference: such sweeps run
 * on the grid-cooperative variant, without heating and without the deterministic mode.  Reset by device_close. */
int asora_set_grey_notables(int on);

/* Override the launch shape of the shared-memory sweep: sources per CTA (1 or 2) and, in the low 16 bits of
 * block_threads, threads per CTA (256, 512, 768, 896 or 1024 for one source; 256 for two); 0 =
 * automatic.  Bits 16-18 of block_threads toggle launch options against their automatic choice (copies of the
 * log2 table in shared memory, table gathers through the texture pipe, offsets word fetched one cell ahead);
 * bit 19 toggles the (k,i,j)-ordered grid copies for z-face cells (automatic only for sweeps of >= 5e8 updates);
 * bits 20-23 force the number of parts a source is split into (1, 2, 4, 8).  For tuning and profiling. */
int asora_set_tuning(int sources_per_cta, int block_threads);

/* Override the launch shape of the mirror-image sweep (variant 3): octants per CTA (8, 4, 2), mirror images per thread,
 * images evaluated side by side, threads per CTA (low 16 bits of block_threads); 0 = automatic for that field.  Only
 * instantiated combinations are accepted at launch time (csrc/sweep_octant.cu).  Bits 16-23 of block_threads are
 * profiling knobs: 1 = eight copies of the log2 table, 2 = toggle the (k,i,j)-ordered z-face grid copies against their
 * automatic choice, 8 = no de-duplication of plane cells; probe builds add 4 and 16 (see launch_opts).  For tuning
 * and profiling. */
int asora_set_octant_shape(int octants_per_cta, int images_per_thread, int batch, int block_threads);

/* Override the launch shape of the large-radius sweep (variant 4): CTAs per cluster (1, 2, 4, 8) and threads per CTA
 * (256, 384 or 512 for clusters of 8; 256 or 512 otherwise); 0 = automatic.  For tuning and profiling. */
int asora_set_cluster_shape(int ctas_per_cluster, int block_threads);

/* Number of sweep plans built since the library was loaded (the plans are cached per mesh, radius, cell size and
 * split; tests use this to check that repeated sweeps do not rebuild them). */
int asora_plan_builds(void);

/* The host-side sweep plan exactly as the kernels consume it (csrc/sweep_plan.cu); needs no device.  octant != 0: the
 * positive-octant plan of the mirror-image sweep (parts ignored), else the whole-sweep plan split into `parts`.
 * Returns the number of plan entries (or -1); info[6] = {levels, largest level, lowest offset, offsets per axis,
 * q_max, parts}.  The arrays are filled only when capacity >= the number of entries: path[n], inv_np[n],
 * upstream_slots[4n], offsets[3n] (biased by -info[2]; octant plans: absolute values), flags[n] (octant plans:
 * flags >> 5 = mask of zero offsets), minor_ab[2n], level_start[parts * (levels + 1)], level_mid[3 * levels]
 * (octant plans only).  Any pointer may be NULL.  For tests of the plan construction. */
int64_t asora_plan_export(int N, double R, double dr, int sphere_only, int octant, int parts, int64_t capacity,
                          double* path, double* inv_np, uint16_t* upstream_slots, uint8_t* offsets, uint8_t* flags,
                          uint8_t* minor_ab, int* level_start, int* level_mid, int* info);

/* Statistics of the most recent ray trace: variant used, number of kernel launches, number of
 * (source, cell) updates, q_max, number of Chebyshev levels, device milliseconds (CUDA events on
 * the context's stream around the whole sweep: opacity pre-pass, zeroing, sweep kernel, division pass).
 * Any pointer may be NULL. */
int asora_last_sweep_stats(int* variant, int* launches, int64_t* updates, int* q_max, int* levels,
                           float* kernel_ms);

/* Device milliseconds of the sweep kernel of the most recent ray trace alone (CUDA events on the launching
 * stream right before and after that one launch). */
int asora_last_sweep_kernel_ms(float* kernel_only_ms);

/* Number of cells the sweep visits per source: |octahedron(q_max) & cube| (raytracing.cu:101,
 * 122-123,241).  Pure host arithmetic; needs no device. */
int64_t asora_cells_per_source(int N, double R);

const char* asora_last_error(void);
const char* asora_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ASORA_B200_H */
