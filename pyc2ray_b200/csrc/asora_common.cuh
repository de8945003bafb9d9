// asora_common.cuh -- shared declarations of libasora_b200 (sm_100a).
//
// Vocabulary (follows the reference, phirling/pyc2ray):
//   source      a point emitter at a mesh cell, flux in units of 1e48 photons/s
//   sweep       the short-characteristics pass that propagates HI column density outwards from a
//               source and turns (column in, column out) into a photo-ionisation rate per cell
//   level       all cells at one Chebyshev distance m = max(|di|,|dj|,|dk|) from the source.  Every
//               non-zero interpolation weight of a level-m cell points at a level-(m-1) cell, so a
//               level is the widest set of cells that can be updated concurrently, and a sweep
//               needs min(q_max, N/2)+1 dependent steps instead of the q_max+1 octahedral shells
//               of src/asora/raytracing.cu:198 (proof: DESIGN.md, "Chebyshev levels")
//   plan        the source-independent description of a sweep (cell offsets, interpolation
//               fractions, path lengths, upstream slots), built once per (N, R, dr) on the host and
//               shared by all sources through L2
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#define ASORA_FOURPI 12.566370614359172463991853874177  // src/asora/raytracing.cu:12
#define ASORA_SQRT3 1.73205080757                       // raytracing.cu:14,435
#define ASORA_SQRT2 1.41421356237                       // raytracing.cu:439
#define ASORA_MAX_COLDENSH 2e30                         // raytracing.cu:15
#define ASORA_TAU_PHOTO_LIMIT 1.0e-7                    // src/asora/rates.cu:7
#define ASORA_S_STAR_REF 1e48                           // src/asora/rates.cu:8 (used by the grey-opacity test rates only)

// Most concurrent sources (CTA groups, one scratch grid each) of the grid-cooperative sweep
#define ASORA_GRID_GROUPS_MAX 74

// Plan cell flags
#define PC_RATED 1u   // inside the R_max sphere: dist2/(dr*dr) <= R*R  (raytracing.cu:315)
#define PC_SOURCE 2u  // the source cell itself (raytracing.cu:285-294)
#define PC_DIAG2 4u   // incoming column scaled by sqrt(2) (raytracing.cu:431-441)
#define PC_DIAG3 8u   // incoming column scaled by sqrt(3)
#define PC_ZFACE 16u  // interior of a z face of its level (|dk| = m > |di|, |dj|): consecutive cells step in j, not in k

// One cell of the sweep plan on the host (48 bytes); the device copy keeps 32 of them (two 16-byte streams).
struct __align__(16) PlanCell {
    double wA, wB;   // |minor offset| / |dominant offset| for the two minor axes
    double path;     // path length through the cell in cell units (raytracing.cu:444,489,533)
    double inv_np;   // 1 / ((di^2+dj^2+dk^2) * path): 1/vol_ph = inv_np / (4 pi dr^3) (raytracing.cu:300-307);
                     // 4 pi for the source cell, whose volume is dr^3 (raytracing.cu:292)
    uint16_t nb[4];  // slots, in the previous level, of the 4 upstream cells c1..c4
    uint8_t d[3];    // offset from the source, biased by -plan.lo (index into the per-source wrap tables)
    uint8_t flags;
    uint32_t ab;     // |minor A| | |minor B| << 8: wA = a/m, wB = b/m are rebuilt on the device (m = level = |dominant|)
};
static_assert(sizeof(PlanCell) == 48, "PlanCell must be 48 bytes");

struct SweepPlan {
    int N = 0;
    double R = 0, dr = 0;
    int q_max = 0;
    int nlevels = 0;
    int max_level_cells = 0;
    int lo = 0, side = 0;              // offsets span [lo, lo+side) on every axis
    bool sphere_only = false;          // unrated cells left out (asora_set_sphere_only)
    int parts = 1;                     // independent pieces per source (sweep_plan.cu), one CTA each
    bool octant = false;               // octant plan: positive octant only, applied to its mirror images (build_octant_plan)
    int64_t ncells = 0;                // plan entries (all parts; bounding planes appear once per part)
    std::vector<int> level_start;      // [parts][nlevels+1], absolute entry offsets
    std::vector<int> level_mid;        // octant plans: [3][nlevels] class A / class B boundary for OPT = 8, 4, 2 (sweep_plan.cu)
    std::vector<PlanCell> cells;       // level-major, lexicographic (di,dj,dk) inside a level
    // device copy, split into two 16-byte streams (structure of arrays): a warp's LDG.128 then covers 4
    // contiguous 128-byte lines instead of the lines of a strided array of structures -- the L1/LSU
    // wavefront pipe is the busiest unit of the sweep (profiles/r01c, r01f)
    int4* d_cells = nullptr;           // [2][ncells]: {path,inv_np} | {nb[4],d[3],flags,ab}
    unsigned* d_dwords = nullptr;      // [ncells]: {d[3],flags} once more as a 4-byte stream (one-cell-ahead fetch)
    int* d_level_start = nullptr;
    bool valid = false;
};

// Scalars every sweep kernel needs.
struct SweepParams {
    int N;
    int q_max;
    int last_l, last_r;       // raytracing.cu:122-123
    double R2;                // R*R
    double sig, dr;
    double kpref;             // sigma * dr / (4 pi dr^3): rate prefactor in optical-depth units (sweep_kernels.cu)
    double tau_max;           // sigma * MAX_COLDENSH: no rate beyond this incoming optical depth (raytracing.cu:315)
    double lut_a, lut_b;      // table index = lut_a + lut_b * log2(tau)  (rates.cu:77-78)
    double tau_lo, tau_hi;    // optical depths at which that index reaches 0 (>= 1e-20) and NumTau
    int hi_min;               // high words of a double in [hi_min, hi_min + hi_span) are strictly inside (tau_lo, tau_hi)
    unsigned hi_span;
    double minlogtau, dlogtau;
    int NumTau;               // index clamp as passed by the caller (rates.cu:78-79)
    int ntab;                 // uploaded table length
    const double* nhi;        // ntau = ndens * (1 - xh_av) * sigma * dr, refreshed before every sweep
    double* phi_ion;
    double* phi_heat;         // photo-heating rates, or null (heating off)
    const double2* thick;     // {T[i], T[i+1]-T[i]} pairs of the uploaded tables, one allocation of 4 x ntab entries:
    const double2* thin;      // thick, thin, heat thick, heat thin (the heating half is zero until uploaded)
    cudaTextureObject_t tex_pairs;  // both pair tables as int4 texels: the same 4 x ntab entries
    const double2* log2_tab;  // 256 x {1/c_j, log2 c_j}, c_j the centre of mantissa bin j
    const int* src_pos;
    const double* src_flux;
    int src_begin, src_count;
    int sphere_only;          // grid-cooperative variant: skip cells outside the R sphere
    int grey;                 // analytic grey-opacity rates (the reference's -D GREY_NOTABLES build), grid-cooperative variant
    unsigned zface_offset;    // != 0: z-face cells use the (k,i,j)-ordered copies of nhi / phi this many doubles behind them
    // deterministic accumulation (asora_set_deterministic): phi_ion / phi_heat then hold the high parts and det_lo /
    // det_lo_heat the low parts of 128-bit fixed-point sums, see deposit_rate in sweep_device.cuh; det_scale == 0: off
    long long* det_lo;
    long long* det_lo_heat;
    double det_scale;         // 2^s: a rate times this is the fixed-point value
    double det_scale_heat;    // the same for the heating rates (their tables have their own magnitude)
    double* coldens_out;      // optional N^3 grid receiving outgoing optical depths (debug) or
                              // the L2-resident scratch of the grid-cooperative variant
};

// ---- fp64 helpers shared by the device code ----------------------------------------------------------
#ifdef __CUDACC__
// 1/x for a normal, finite x: the MUFU.RCP64H seed reads the upper 32 bits of x (relative error < 2^-19) and one
// cubic step r (1 + e + e^2), e = 1 - x r, brings it below 2^-57; the result is within one ulp.
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    return fma(r, e, r);
}

// a/b with a final residual correction (<= 1 ulp for normal operands); used by the chemistry
__device__ __forceinline__ double fast_div(double a, double b)
{
    const double r = fast_rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

#endif

// host-side helpers implemented in sweep_plan.cu
int asora_qmax(int N, double R);
int64_t asora_count_cells(int N, double R);
// upload = false: host-side plan only (plan.cells, level_start, level_mid), no device needed
int sweep_plan_octant_level_cells(int N, double R, double dr, bool sphere_only);
bool build_sweep_plan(SweepPlan& plan, int N, double R, double dr, bool sphere_only, int parts, std::string& err,
                      bool upload = true);
int64_t asora_count_rated_cells(int N, double R, double dr);
bool build_octant_plan(SweepPlan& plan, int N, double R, double dr, bool sphere_only, std::string& err, bool upload = true);
void free_sweep_plan(SweepPlan& plan);

// launchers implemented in sweep_kernels.cu
size_t sweep_smem_bytes(const SweepPlan& plan, int sources_per_cta, int log2_copies);
cudaError_t launch_sweep_smem(const SweepPlan& plan, const SweepParams& p, int sources_per_cta, int block,
                              int opts, cudaStream_t stream, int* launches);
// sweep_octant.cu: the mirror-image sweep.  `noct` octants per CTA (8, 4, 2: a source is split over 8/noct CTAs),
// `opt` mirror images per thread of which `batch` are evaluated side by side, `block` threads; cudaErrorNotSupported
// for combinations that are not instantiated (sweep_octant_shape_ok).
size_t sweep_octant_smem_bytes(const SweepPlan& plan, int noct, int log2_copies, bool zface);
cudaError_t launch_sweep_octant(const SweepPlan& plan, const SweepParams& p, int noct, int opt, int batch, int block, int opts,
                                cudaStream_t stream, int* launches);
int sweep_octant_shape_ok(int noct, int opt, int batch, int block);  // 0 no, 1 yes, 2 yes incl. z-face copies
// sweep_cluster.cu: the large-radius sweep, one cluster of 2^logc CTAs per wedge of a source (24 wedges per source)
int sweep_cluster_levels(const SweepParams& p);
size_t sweep_cluster_smem_bytes(int nlevels, int logc);
cudaError_t launch_sweep_cluster(const SweepParams& p, int logc, int block, cudaStream_t stream, int* launches, int* levels);
int sweep_grid_groups(const SweepParams& p, int max_groups, int* total_ctas_out, int* group_ctas_out);
cudaError_t launch_sweep_grid(const SweepParams& p, int ngroups, unsigned* counters, cudaStream_t stream,
                              int* launches, int* levels);

cudaError_t launch_prepare_nhi(const double* ndens, const double* xh_av, double* ntau, double sig_dr, int64_t ncell,
                               cudaStream_t stream);
cudaError_t launch_peer_halo(double* dst, const double* src, int64_t n, bool add, cudaStream_t stream);
cudaError_t launch_finish_phi(double* phi, const double* ntau, const double* keep, int64_t ncell, cudaStream_t stream);
cudaError_t launch_prepare_nhi_transposed(const double* ndens, const double* xh_av, double* ntau, double* ntau_t, double sig_dr,
                                          int N, cudaStream_t stream);
cudaError_t launch_finish_phi_transposed(double* phi, const double* phi_t, const double* ntau, const double* keep, int N,
                                         cudaStream_t stream);
cudaError_t launch_finish_phi_fixed(double* phi_hi, const long long* lo, const double* ntau, const double* keep, double inv_scale,
                                    int64_t ncell, cudaStream_t stream);
cudaError_t launch_reverse_axes(const double* in, double* out, int N, cudaStream_t stream);
cudaError_t launch_scale_grid(double* grid, double factor, int64_t ncell, cudaStream_t stream);
cudaError_t launch_pair_table(const double* table, double2* pairs, int ntab, cudaStream_t stream);
void host_log2_table(double* tab512);
cudaError_t upload_inv_levels();

// chemistry.cu
cudaError_t launch_temperature_factors(const double* temp, double2* factors, double bh00, double albpow, double colh0,
                                       double temph0, int64_t ncell, cudaStream_t stream);
cudaError_t launch_global_pass(double dt, const double* ndens, const double* temp, const double2* factors,
                               const double* xh, double* xh_av, double* xh_intermed, const double* phi_ion, double bh00,
                               double albpow, double colh0, double temph0, double abu_c, int64_t ncell,
                               int store_av_first, double* d_partials, int* d_iparts, int nblocks_max,
                               int* conv_flag, double* sum1, double* sum0, cudaStream_t stream);
int chemistry_partial_blocks(int64_t ncell);
