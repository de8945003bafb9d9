// asora_api.cu -- the C ABI of libasora_b200.so (include/asora_b200.h): per-process device context,
// host<->device staging, sweep-variant selection and launch.
//
// Replaces src/asora/memory.cu (device state, :20-129), the host driver do_all_sources_gpu
// (src/asora/raytracing.cu:79-148) and the CPython wrappers (src/asora/python_module.cu).
#include "../../include/asora_b200.h"
#include "asora_common.cuh"

#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up under nsys / ncu --nvtx, no-ops otherwise

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <utility>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Context {
    bool init = false;
    int N = 0;
    int64_t ncell = 0;
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    int smem_per_sm = 0;
    bool zface_ok = false;     // 2 N^3 < 2^32: the (k,i,j)-ordered copies of nhi / phi fit the 32-bit cell positions
    cudaStream_t stream = nullptr;      // the stream work is queued on
    cudaStream_t own_stream = nullptr;  // created by device_init; `stream` unless the caller set one
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around a whole sweep: pre-pass, zeroing, sweep kernel, division pass
    cudaEvent_t evk0 = nullptr, evk1 = nullptr; // around the sweep kernel alone
    double* buf[ASORA_BUF_COUNT] = {nullptr};
    double2* thin = nullptr;   // {T[i], T[i+1]-T[i]} pairs (sweep_kernels.cu: photo_lookup)
    double2* thick = nullptr;
    cudaTextureObject_t tex_pairs = 0;  // one allocation: thick pairs, then thin pairs
    int ntab = 0;
    bool heat_tables = false;  // heating half of the pair tables uploaded (asora_heat_table_to_device)
    bool heating = false;      // accumulate ASORA_BUF_PHI_HEAT in the next sweeps (asora_set_heating)
    double* grid_scratch = nullptr;  // ngroups x N^3 column densities of the grid-cooperative sweep
    int grid_scratch_groups = 0;
    int grid_max_groups = 0;
    unsigned* grid_counters = nullptr;
    double* nhi = nullptr;      // ndens * (1 - xh_av) * sigma * dr, rebuilt before every sweep
    double* phi_keep = nullptr; // rates of earlier sweeps while a sweep accumulates on top of them (zero_phi = 0)
    // pageable host buffers: per-thread pinned bounce buffers and streams of host_copy()
    static constexpr int kCopyThreads = 16;   // upper bound; copy_threads() picks how many are used
    static constexpr size_t kBounceBytes = (size_t)4 << 20;
    char* bounce[kCopyThreads][2] = {{nullptr}};
    cudaStream_t copy_stream[kCopyThreads] = {nullptr};
    cudaEvent_t copy_event[kCopyThreads][2] = {{nullptr}};
    double2* log2_tab = nullptr;
    int* src_pos = nullptr;      // as uploaded (positions reduced modulo N)
    double* src_flux = nullptr;
    int* src_pos_sorted = nullptr;   // the same sources in Morton order of their cells: consecutive CTAs
    double* src_flux_sorted = nullptr;  // then sweep neighbouring regions and share ndens/phi lines in L2
    int nsrc = 0;
    std::vector<int32_t> h_src_pos;  // positions of the last upload as passed, and their Morton permutation: an upload of
    std::vector<int> h_src_perm;     // the same positions only refreshes the fluxes
    // sweep plans, most recently used first: (N, R, dr, sphere_only, kind, parts) -> plan.  A time step alternates between
    // at most a few of them (full / sphere-only, split or not), and a rebuild costs several O(side^3) host passes plus a
    // synchronous upload, so the cache holds more than one and remembers the automatic split per key.
    static constexpr int kPlanCache = 6;
    SweepPlan plans[kPlanCache];
    int plan_failed_parts[kPlanCache] = {0};
    struct AutoParts { int N; double R, dr; bool sphere_only; int parts; };
    std::vector<AutoParts> auto_parts;   // parts chosen by the automatic split, per (N, R, dr, sphere_only)
    int plan_builds = 0;                 // number of plans built so far (tests: the cache must hit)
    // temperature factors of the chemistry (chemistry.cu), valid for the TEMP buffer contents and constants below
    double2* chem_factors = nullptr;
    bool chem_factors_valid = false;
    double cf_bh00 = 0, cf_albpow = 0, cf_colh0 = 0, cf_temph0 = 0;
    // chemistry scratch
    double* chem_partials = nullptr;
    int* chem_iparts = nullptr;
    int chem_blocks = 0;
    // host-API chemistry staging (ncell doubles each), allocated on demand
    double* chem_stage[6] = {nullptr};
    int64_t chem_stage_n = 0;
    // stats of the last sweep
    int variant_forced = 0;
    int tune_S = 0, tune_block = 0, tune_opts = 0;
    int tune_parts = 0;
    int oct_noct = 0, oct_opt = 0, oct_batch = 0, oct_block = 0;  // forced shape of the mirror-image sweep (0 = automatic)
    int oct_opts = 0;                                              // its profiling knobs
    int clu_logc = 0, clu_block = 0;  // forced cluster size (log2) and threads of the large-radius sweep (0 = automatic)
    int large_smem_min_sources = 4;   // q_max > 127: fewer sources than this go to the cluster sweep (R = 80, 4 sources: equal)
    int slab_begin = 0, slab_count = 0;  // active planes of a slab-decomposed run (0 = whole grid)
    int sphere_only = 0;
    int deterministic = 0;            // asora_set_deterministic: fixed-point accumulation of the rates
    int grey = 0;                     // asora_set_grey_notables: analytic grey-opacity rates instead of the tables
    long long* det_lo = nullptr;      // low parts of the fixed-point sums (N^3), and of the heating sums
    long long* det_lo_heat = nullptr;
    double flux_max = 0.0;            // largest uploaded source flux
    double table_max = 0.0;           // largest |photo table entry|: bounds one cell's absorbed fraction
    double heat_table_max = 0.0;      // the same for the heating tables
    // parameters of the last sweep, for the lazily evaluated update count; and a one-entry cache of it
    int last_count = 0;
    double last_R = 0, last_dr = 0;
    bool last_sphere_only = false;
    int cells_N = 0;
    double cells_R = 0, cells_dr = 0;
    bool cells_sphere_only = false;
    int64_t cells_per_source = 0;
    int last_variant = 0, last_launches = 0, last_qmax = 0, last_levels = 0;
    int64_t last_updates = 0;
    float last_ms = 0.f;
};

Context g;
std::string g_err;

int fail(const std::string& what)
{
    g_err = what;
    return 1;
}
int fail_cuda(const char* where, cudaError_t e)
{
    g_err = std::string(where) + ": " + cudaGetErrorName(e) + " - " + cudaGetErrorString(e);
    return 2;
}
#define CK(call)                                          \
    do {                                                  \
        cudaError_t _e = (call);                          \
        if (_e != cudaSuccess) return fail_cuda(#call, _e); \
    } while (0)

void free_tables()
{
    if (g.tex_pairs) cudaDestroyTextureObject(g.tex_pairs);
    g.tex_pairs = 0;
    if (g.thick) cudaFree(g.thick);  // one allocation holds all four tables
    g.thin = g.thick = nullptr;
    g.heat_tables = false;
    g.heating = false;
}

// NVTX range for the lifetime of the object (tracing)
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

int need_init()
{
    if (!g.init) return fail("GPU not initialized. Please initialize it by calling device_init(N)");
    return 0;
}

int ensure_buffer(int which)
{
    if (which < 0 || which >= ASORA_BUF_COUNT) return fail("unknown buffer id");
    if (!g.buf[which]) {
        // the rate grids carry a second, (k,i,j)-ordered half for the z-face cells of large sweeps (run_sweep)
        const bool twice = g.zface_ok && (which == ASORA_BUF_PHI_ION || which == ASORA_BUF_PHI_HEAT);
        const size_t n = (size_t)g.ncell * (twice ? 2 : 1);
        CK(cudaMalloc(&g.buf[which], sizeof(double) * n));
        CK(cudaMemsetAsync(g.buf[which], 0, sizeof(double) * n, g.stream));
        // a temperature grid means chemistry passes will follow: their factor grid is allocated with it, so that the
        // first pass of a time step does not pay a 16 N^3-byte cudaMalloc
        if (which == ASORA_BUF_TEMP && !g.chem_factors) CK(cudaMalloc(&g.chem_factors, sizeof(double2) * g.ncell));
    }
    return 0;
}

// Staging threads of host_copy(): ASORA_COPY_THREADS, else half the host threads this process may run on, 4 to 8.
int copy_threads()
{
    static int n = 0;
    if (n == 0) {
        const char* env = std::getenv("ASORA_COPY_THREADS");
        int v = env ? std::atoi(env) : 0;
        if (v <= 0) {
            v = (int)std::thread::hardware_concurrency() / 2;
            v = std::max(4, std::min(8, v));
        }
        n = std::max(1, std::min(Context::kCopyThreads, v));
    }
    return n;
}

// Host <-> device copy of a whole grid.  Pinned host memory goes straight to the copy engine.  Pageable memory
// (what numpy hands over) would be staged by the driver on one thread at ~11 GB/s (measured: 11 ms per 125 MB grid,
// five grids per evolve3D call); here kCopyThreads host threads stage disjoint slices through their own pinned
// bounce buffers and streams, double-buffered, which is limited by PCIe instead.  Synchronous, like the API.
int host_copy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
    Range range(kind == cudaMemcpyHostToDevice ? "asora:h2d" : "asora:d2h");
    const void* host = (kind == cudaMemcpyHostToDevice) ? src : dst;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, host) == cudaSuccess)
        pinned = (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    else
        cudaGetLastError();
    CK(cudaStreamSynchronize(g.stream));
    if (pinned || bytes < ((size_t)8 << 20)) {
        CK(cudaMemcpyAsync(dst, src, bytes, kind, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        return 0;
    }
    const int T = copy_threads();
    const size_t B = Context::kBounceBytes;
    for (int t = 0; t < T; t++) {
        if (!g.copy_stream[t]) CK(cudaStreamCreateWithFlags(&g.copy_stream[t], cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            if (!g.bounce[t][b]) CK(cudaMallocHost(&g.bounce[t][b], B));
            if (!g.copy_event[t][b]) CK(cudaEventCreateWithFlags(&g.copy_event[t][b], cudaEventDisableTiming));
        }
    }
    const int device = g.device;
    cudaError_t errs[Context::kCopyThreads];
    auto worker = [&](int t) {
        cudaError_t e = cudaSetDevice(device);
        // slice of this thread, in whole bounce buffers
        const size_t nchunks = (bytes + B - 1) / B;
        const size_t c0 = nchunks * t / T, c1 = nchunks * (t + 1) / T;
        cudaStream_t st = g.copy_stream[t];
        if (kind == cudaMemcpyHostToDevice) {
            for (size_t c = c0; c < c1 && e == cudaSuccess; c++) {
                const int b = (int)((c - c0) & 1);
                const size_t off = c * B, len = std::min(B, bytes - off);
                if (c - c0 >= 2) e = cudaEventSynchronize(g.copy_event[t][b]);  // bounce buffer free again
                if (e != cudaSuccess) break;
                std::memcpy(g.bounce[t][b], (const char*)src + off, len);
                e = cudaMemcpyAsync((char*)dst + off, g.bounce[t][b], len, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaEventRecord(g.copy_event[t][b], st);
            }
        } else {
            // device -> bounce of chunk c+1 in flight while chunk c is copied out
            auto issue = [&](size_t c) {
                const int b = (int)((c - c0) & 1);
                const size_t off = c * B, len = std::min(B, bytes - off);
                cudaError_t r = cudaMemcpyAsync(g.bounce[t][b], (const char*)src + off, len, cudaMemcpyDeviceToHost, st);
                if (r == cudaSuccess) r = cudaEventRecord(g.copy_event[t][b], st);
                return r;
            };
            if (c0 < c1) e = issue(c0);
            for (size_t c = c0; c < c1 && e == cudaSuccess; c++) {
                const int b = (int)((c - c0) & 1);
                const size_t off = c * B, len = std::min(B, bytes - off);
                if (c + 1 < c1) e = issue(c + 1);
                if (e == cudaSuccess) e = cudaEventSynchronize(g.copy_event[t][b]);
                if (e == cudaSuccess) std::memcpy((char*)dst + off, g.bounce[t][b], len);
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        errs[t] = e;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < T; t++)
        if (errs[t] != cudaSuccess) return fail_cuda("host_copy", errs[t]);
    return 0;
}

int ensure_chem_scratch()
{
    if (!g.chem_partials) {
        g.chem_blocks = chemistry_partial_blocks((int64_t)1 << 40);
        CK(cudaMalloc(&g.chem_partials, sizeof(double) * (2 * g.chem_blocks + 2)));
        CK(cudaMalloc(&g.chem_iparts, sizeof(int) * (g.chem_blocks + 1)));
    }
    return 0;
}

// Temperature factors for the device-resident chemistry: refilled when the TEMP buffer was rewritten or the
// constants changed (whole grid, also in slab-decomposed runs: stale planes are never read).
int ensure_chem_factors(double bh00, double albpow, double colh0, double temph0)
{
    if (!g.chem_factors) CK(cudaMalloc(&g.chem_factors, sizeof(double2) * g.ncell));
    if (g.chem_factors_valid && g.cf_bh00 == bh00 && g.cf_albpow == albpow && g.cf_colh0 == colh0 && g.cf_temph0 == temph0)
        return 0;
    cudaError_t e = launch_temperature_factors(g.buf[ASORA_BUF_TEMP], g.chem_factors, bh00, albpow, colh0, temph0, g.ncell,
                                               g.stream);
    if (e != cudaSuccess) return fail_cuda("temperature_factors_kernel launch", e);
    g.chem_factors_valid = true;
    g.cf_bh00 = bh00;
    g.cf_albpow = albpow;
    g.cf_colh0 = colh0;
    g.cf_temph0 = temph0;
    return 0;
}

// Cached sweep plan for the current mesh: octant = true -> build_octant_plan, else build_sweep_plan(parts).
// Returns nullptr (and sets err) when the plan cannot be built.
SweepPlan* get_plan(int N, double R, double dr, bool sphere_only, bool octant, int parts, std::string& err)
{
    SweepPlan* c = g.plans;
    for (int i = 0; i < Context::kPlanCache; i++) {
        if (c[i].valid && c[i].N == N && c[i].R == R && c[i].dr == dr && c[i].sphere_only == sphere_only &&
            c[i].octant == octant && c[i].parts == parts) {
            std::rotate(c, c + i, c + i + 1);  // most recently used first
            return &c[0];
        }
    }
    free_sweep_plan(c[Context::kPlanCache - 1]);  // evict the least recently used entry
    std::rotate(c, c + Context::kPlanCache - 1, c + Context::kPlanCache);
    g.plan_builds++;
    const bool ok = octant ? build_octant_plan(c[0], N, R, dr, sphere_only, err)
                           : build_sweep_plan(c[0], N, R, dr, sphere_only, parts, err);
    return ok ? &c[0] : nullptr;
}

// Choose and launch the sweep for sources [begin, begin+count).  Inputs already on the device.
int run_sweep(double R, double sig, double dr, int begin, int count, double minlogtau, double dlogtau,
              int NumTau, bool zero_phi, double* coldens_grid)
{
    if (!g.grey && (!g.thin || !g.thick)) return fail("photo tables not on device: call photo_table_to_device first");
    if (g.grey && (g.heating || g.deterministic))
        return fail("grey-opacity test rates (set_grey_notables) exist without heating and without the deterministic mode only");
    if (begin < 0 || count < 0 || begin + count > g.nsrc)
        return fail("source range exceeds the list uploaded with source_data_to_device");
    if (int rc = ensure_buffer(ASORA_BUF_NDENS)) return rc;
    if (int rc = ensure_buffer(ASORA_BUF_XH_AV)) return rc;
    if (int rc = ensure_buffer(ASORA_BUF_PHI_ION)) return rc;
    const int N = g.N;
    if (!g.nhi) CK(cudaMalloc(&g.nhi, sizeof(double) * g.ncell * (g.zface_ok ? 2 : 1)));
    if (!g.log2_tab) {
        double h[512];
        host_log2_table(h);
        CK(cudaMalloc(&g.log2_tab, sizeof(h)));
        CK(cudaMemcpy(g.log2_tab, h, sizeof(h), cudaMemcpyHostToDevice));
        CK(upload_inv_levels());
    }

    SweepParams p;
    p.N = N;
    p.q_max = asora_qmax(N, R);
    p.last_r = N / 2 - 1 + (N % 2);
    p.last_l = -N / 2;
    p.R2 = R * R;
    p.sig = sig;
    p.dr = dr;
    p.kpref = (sig * dr) / (ASORA_FOURPI * (dr * dr * dr));
    p.grey = g.grey;
    p.tau_max = sig * ASORA_MAX_COLDENSH;
    // rates.cu:77-78: index = 1 + (log10(tau) - minlogtau)/dlogtau = lut_a + lut_b * log2(tau)
    p.lut_b = 0.30102999566398119521 / dlogtau;
    p.lut_a = 1.0 - minlogtau / dlogtau;
    // rates.cu:77-78 clamps: tau >= 1e-20, 0 <= index <= NumTau  <=>  tau_lo <= tau <= tau_hi
    // (a hair above the exact bound, 4e-10 of a table bin, so that rounding in the logarithm cannot produce a
    // negative index: table_index takes floor() by a rounded-down addition, which assumes index >= 0)
    p.tau_lo = std::max(1.0e-20, std::pow(10.0, minlogtau - dlogtau)) * (1.0 + 1e-12);
    p.tau_hi = std::pow(10.0, minlogtau + ((double)NumTau - 1.0) * dlogtau);
    {   // fast range test of photo_lookup: high words strictly between those of tau_lo and tau_hi
        uint64_t blo, bhi;
        std::memcpy(&blo, &p.tau_lo, 8);
        std::memcpy(&bhi, &p.tau_hi, 8);
        const int hlo = (int)(blo >> 32) + 1, hhi = (int)(bhi >> 32);  // [hlo, hhi) is safe
        p.hi_min = hlo;
        p.hi_span = hhi > hlo ? (unsigned)(hhi - hlo) : 0u;
    }
    p.minlogtau = minlogtau;
    p.dlogtau = dlogtau;
    p.NumTau = NumTau;
    p.ntab = g.ntab;
    p.nhi = g.nhi;
    p.log2_tab = g.log2_tab;
    p.phi_ion = g.buf[ASORA_BUF_PHI_ION];
    p.phi_heat = nullptr;
    if (g.heating) {
        if (!g.heat_tables) return fail("heating requested but no heating tables on device: call heat_table_to_device first");
        if (coldens_grid) return fail("heating is not available on the single-source debug path");
        if (!zero_phi) return fail("heating needs zero_phi = 1");
        if (int rc = ensure_buffer(ASORA_BUF_PHI_HEAT)) return rc;
        p.phi_heat = g.buf[ASORA_BUF_PHI_HEAT];
    }
    p.det_lo = p.det_lo_heat = nullptr;
    p.det_scale = p.det_scale_heat = 0.0;
    if (g.deterministic) {
        if (coldens_grid) return fail("deterministic accumulation is not available on the single-source debug path");
        // largest possible contribution: strength * kpref * 4 pi (source cell) * (difference of two table entries)
        const double vmax = 2.0 * g.flux_max * p.kpref * ASORA_FOURPI * g.table_max;
        int e2 = 0;
        std::frexp(vmax > 0.0 ? vmax : 1.0, &e2);  // vmax < 2^e2
        p.det_scale = std::ldexp(1.0, 88 - e2);   // contributions stay below 2^88 = 2^42 high parts of 2^46
        if (!g.det_lo) CK(cudaMalloc(&g.det_lo, sizeof(long long) * g.ncell));
        p.det_lo = g.det_lo;
        if (p.phi_heat) {
            if (!g.det_lo_heat) CK(cudaMalloc(&g.det_lo_heat, sizeof(long long) * g.ncell));
            p.det_lo_heat = g.det_lo_heat;
            const double hmax = 2.0 * g.flux_max * p.kpref * ASORA_FOURPI * g.heat_table_max;
            std::frexp(hmax > 0.0 ? hmax : 1.0, &e2);
            p.det_scale_heat = std::ldexp(1.0, 88 - e2);
        }
    }
    p.thin = g.thin;
    p.thick = g.thick;
    p.tex_pairs = g.tex_pairs;
    // a sweep over the whole list may take the sources in any order (phi_ion is a sum)
    const bool whole = (begin == 0 && count == g.nsrc && g.src_pos_sorted != nullptr);
    p.src_pos = whole ? g.src_pos_sorted : g.src_pos;
    p.src_flux = whole ? g.src_flux_sorted : g.src_flux;
    p.src_begin = begin;
    p.src_count = count;
    const bool sphere_only = g.sphere_only && !coldens_grid;  // the debug path keeps the reference's cell set
    p.sphere_only = sphere_only ? 1 : 0;
    p.coldens_out = coldens_grid;

    g.last_qmax = p.q_max;
    g.last_launches = 0;
    // the update count of this sweep is evaluated lazily (asora_last_sweep_stats): counting cells is O(side^3)
    g.last_count = count;
    g.last_R = R;
    g.last_dr = dr;
    g.last_sphere_only = sphere_only;
    g.last_updates = -1;
    g.last_ms = 0.f;

    // Variant selection.  1: shared-memory level sweep, one cell per thread (sweep_kernels.cu); 3: mirror-image sweep,
    // one plan entry and up to eight octant images per thread (sweep_octant.cu), for mirror-symmetric cell sets;
    // 2: grid-cooperative sweep through L2 scratch grids when a level does not fit in shared memory.
    int variant = g.variant_forced;
    if (g.grey) variant = 2;  // the analytic test rates exist in the grid-cooperative sweep only (any mesh, any radius)
    int S = 1, block = 256, opts = 0;
    int noct = 8, opt = 8, batch = 4;
    SweepPlan* plan = nullptr;
    const int hi_cells = std::min(p.q_max, std::max(-p.last_l, p.last_r));
    const bool symmetric = std::min(p.q_max, p.last_r) == std::min(p.q_max, -p.last_l) && hi_cells <= 126;
    const size_t budget = (size_t)g.smem_optin;
    // automatic choice (measured, scripts/octant_probe.py, profiles/r02f_radius_sweep_256.md, bench field): the mirror-image
    // sweep wins on sweeps whose eight octants just fit one SM -- full octahedron, 44 <= q_max <= 60: R = 25 9.7 vs 10.6 ms,
    // R = 30 14.5 vs 17.0 ms, R = 34 20.2 vs 26.8 ms (10^4 sources); sphere only from q_max = 50: R = 30 11.1 vs 11.5 ms,
    // R = 34 15.5 vs 16.6 ms, but R = 25 7.3 vs 7.0 ms.  Smaller radii and larger ones (split sweeps) stay on variant 1
    // (R = 20: 52.1 vs 54.0 ms for 10^5 sources; R = 40: 10.7 vs 11.6 ms).
    bool auto_octant = hi_cells >= (sphere_only ? 50 : 44) && hi_cells <= (sphere_only ? 60 : 126) && !g.heating;
    if (auto_octant && variant == 0 && hi_cells > 60) {
        // quadrant CTAs, two per SM: decided from a cell count, before an octant plan that would not be used is built
        const size_t cells = (size_t)sweep_plan_octant_level_cells(N, R, dr, false);
        const size_t side_o = 2 * (size_t)hi_cells + 1, nl = (size_t)hi_cells + 1;
        const size_t smem2 = 256 * sizeof(double2) + 4 * cells * sizeof(double) + nl * sizeof(double) + 6 * side_o * sizeof(unsigned) +
                             (2 * nl + 1) * sizeof(int);
        if (2 * smem2 + 2048 > (size_t)g.smem_per_sm) auto_octant = false;
    }
    if ((variant == 3 || (variant == 0 && auto_octant)) && symmetric && !coldens_grid) {
        std::string err;
        plan = get_plan(N, R, dr, sphere_only, true, 1, err);
        if (!plan && variant == 3) return fail(err);
        if (plan) {
            // octants per CTA: all eight while their level buffers fit, else half-spaces / quadrants as separate CTAs
            noct = g.oct_noct > 0 ? g.oct_noct : 8;
            while (g.oct_noct == 0 && noct > 2 && sweep_octant_smem_bytes(*plan, noct, 1, false) > budget) noct /= 2;
            if (sweep_octant_smem_bytes(*plan, noct, 1, false) > budget) {
                if (variant == 3) return fail("sweep variant 3 forced but a level does not fit in shared memory");
                plan = nullptr;
            }
            // automatic: while all eight octants fit one CTA, or -- full cell set -- two quadrant CTAs fit one SM (below)
            if (plan && variant == 0 && noct < 8 &&
                (sphere_only || sweep_octant_smem_bytes(*plan, 2, 1, true) * 2 + 2048 > (size_t)g.smem_per_sm))
                plan = nullptr;
        }
        if (plan) {
            // Launch shape (measured on B200, scripts/octant_probe.py; DESIGN.md, "Mirror-image sweep").  Large levels:
            // half-spaces as separate CTAs, two of them per SM, four images per thread evaluated side by side, the next
            // plan entry prefetched, plane cells recomputed (the de-duplicated path costs more than the 3 % of cells it
            // saves at R = 30: 16.3 vs 16.4-17.9 ms); small levels: de-duplication on (24 % of the cells at R = 10).
            const int maxc = plan->max_level_cells;
            int auto_opts = 0;
            if (g.oct_noct == 0 && maxc >= 512 && sweep_octant_smem_bytes(*plan, 4, 1, true) * 2 + 2048 <= (size_t)g.smem_per_sm) noct = 4;
            // beyond that (q_max 61 ... ~84 at 256^3): quadrants, two CTAs per SM (R = 38: 7.6 vs 9.0 ms with variant 1 for 2000
            // sources, R = 44: 11.8 vs 13.7 ms; half-spaces with one CTA per SM: 8.4 / 13.7 ms)
            else if (g.oct_noct == 0 && variant == 0 && noct < 8) noct = 2;
            if (noct == 8) {
                if (maxc >= 512) { opt = 4; batch = 2; block = 512; }
                else if (maxc >= 128) { opt = 4; batch = 2; block = 128; }
                else { opt = 2; batch = 1; block = 128; }
            } else if (noct == 4) {
                opt = 4; batch = 4; block = 256;
                auto_opts = 4 | 8;
            } else {
                opt = 2; batch = 2; block = variant == 0 ? 256 : 192;
                auto_opts = variant == 0 ? (4 | 8) : 8;
            }
            if (g.oct_opt > 0) opt = g.oct_opt;
            if (g.oct_batch > 0) batch = g.oct_batch;
            if (g.oct_block > 0) block = g.oct_block;
            if (!sweep_octant_shape_ok(noct, opt, batch, block)) return fail("mirror-image sweep: launch shape not instantiated");
            // profiling knobs (asora_set_octant_shape; csrc/sweep_octant.cu: launch_opts) replace the automatic options
            opts = (g.oct_opts || g.oct_noct || g.oct_opt || g.oct_batch || g.oct_block) ? g.oct_opts : auto_opts;
            variant = 3;
        }
    }
    if (variant == 3 && !plan) return fail("sweep variant 3 forced but the swept region is not mirror-symmetric");
    if (variant != 2 && variant != 3) {
        const int side = std::min(p.q_max, p.last_r) + std::min(p.q_max, -p.last_l) + 1;
        // Beyond q_max = 127 only the eight-octant split can fit, and only up to R ~ 100 cells at 256^3: counted before a
        // plan of up to N^3 entries is built.  Few sources at such radii are better spread over every SM (variant 4).
        // (measured, 256^3: R = 80, 128 sources 105 G updates/s against 30 with the wedge clusters, R = 100, 64 sources 86
        // against 39; full 128^3 box, 256 sources 129 against 51: a plan entry replaces ~130 instructions of geometry)
        bool large_ok = true;
        if (p.q_max > 127 && side <= 256 && g.tune_parts == 0) {
            large_ok = variant == 1 || count >= g.large_smem_min_sources;
            if (large_ok) {
                const size_t cells = (size_t)sweep_plan_octant_level_cells(N, R, dr, sphere_only);
                large_ok = 256 * sizeof(double2) + 2 * cells * sizeof(double) + 6 * (size_t)side * sizeof(unsigned) +
                           (size_t)(hi_cells + 2) * sizeof(int) <= budget;
            }
        }
        if (side <= 256 && large_ok) {
            // Parts: start from the whole sweep and split (half-spaces, quadrants, octants) only until one
            // source's two level buffers fit in shared memory.  (Measured at R = 30, 256^3: 1 part x 1024
            // threads 20.6 ms, 2 x 512: 21.4, 4 x 256: 21.2, 8 x 256 with two sources: 21.1 -- splitting does
            // not pay by itself, it extends the shared-memory variant to radii of ~65 cells.)  The split found for a
            // (mesh, radius, cell size) is remembered, so that later sweeps go straight to the cached plan.
            int parts = g.tune_parts > 0 ? g.tune_parts : (p.q_max > 127 ? 8 : 1);
            Context::AutoParts* memo = nullptr;
            if (g.tune_parts == 0) {
                for (auto& a : g.auto_parts)
                    if (a.N == N && a.R == R && a.dr == dr && a.sphere_only == sphere_only) memo = &a;
                if (memo) parts = memo->parts;
            }
            for (;;) {
                std::string err;
                plan = get_plan(N, R, dr, sphere_only, false, parts, err);
                if (!plan) {
                    if (variant == 1) return fail(err);
                    break;
                }
                if (g.tune_parts > 0 || parts == 8) break;
                if (sweep_smem_bytes(*plan, 1, 1) <= budget) break;
                parts *= 2;
            }
            if (plan && g.tune_parts == 0 && !memo) {
                if (g.auto_parts.size() >= 16) g.auto_parts.erase(g.auto_parts.begin());
                g.auto_parts.push_back({N, R, dr, sphere_only, plan->parts});
            }
        }
        if (plan) {
            const size_t per_src = sweep_smem_bytes(*plan, 1, 1);
            if (per_src > budget) {
                plan = nullptr;
            } else {
                // Launch shape (measured on B200, scripts/perf_probe4.py): 256 threads while a level is at
                // most a few passes wide, 512 for wider levels when two CTAs fit per SM, 1024 when only one
                // does; two sources per CTA (plan decode and barriers amortised) when both fit next to >= 3
                // resident CTAs.
                const int maxc = plan->max_level_cells;
                // (one CTA per SM: 896 threads at 72 registers beat 1024 at 64, where ptxas spills and delays
                // loads, and 768 / 640 / 512 threads: 17.7 vs 18.3 / 17.8 / 18.6 / 20.5 ms at R = 30)
                block = maxc < 2048 ? 256 : (2 * per_src <= budget ? 512 : 896);
                S = (block == 256 && count >= 8 * g.sm_count && 3 * sweep_smem_bytes(*plan, 2, 1) <= budget) ? 2 : 1;
                if (g.tune_S > 0 && (size_t)g.tune_S * per_src <= budget) S = g.tune_S;
                if (g.tune_block > 0) block = g.tune_block;
                // eight bank-staggered copies of the log2 table (28 KB more) for the long-lived one-per-SM CTAs of large
                // radii; short sweeps do not recover the cost of filling them (R = 10.76: 1.20 vs 1.13 ms)
                opts = (block > 512 && sweep_smem_bytes(*plan, S, 8) <= budget) ? 1 : 0;
                // table gathers through the texture pipe: always (R = 30: 18.9 -> 18.2 ms; R = 10.76: 1.25 -> 1.21 ms);
                // offsets word one cell ahead: only the small-radius shape gains (R = 10.76, two sources x 256
                // threads: 1.21 -> 1.11 ms; R = 30, 1024 threads: 18.2 -> 18.9 ms) -- scripts/perf_probe6.py
                opts |= 2;
                if (block == 256 || block == 896) opts |= 4;
                opts ^= (g.tune_opts & 7);  // profiling knob: bits 16-18 of set_tuning's block_threads toggle the options
            }
        }
        if (variant == 1 && !plan) return fail("sweep variant 1 forced but a level does not fit in shared memory");
        if (variant == 0) variant = plan ? 1 : 4;
    }
    // Large radii: one cluster per wedge with the level buffers in distributed shared memory (variant 4) while a CTA's
    // share of two levels fits its shared memory (up to ~1000^3 meshes with clusters of 8), else the grid-cooperative
    // sweep through L2 scratch grids (variant 2).
    int clu_logc = 3, clu_block = 384;
    if (variant == 4) {
        p.sphere_only = sphere_only ? 1 : 0;
        const int nl = sweep_cluster_levels(p);
        // Cluster shape (measured, full 256^3 box): few sources want every SM on each of them -- 8 CTAs x 384 threads per
        // wedge, 192 CTAs per source: 0.50 ms for one source (0.54 / 0.74 ms with 8 x 256 / 4 x 256); from a handful of
        // sources on, 4 x 256 keeps three CTAs per SM busy: 59 G updates/s against 54 (8 x 256) and 47 (8 x 384).
        if (count >= 4) { clu_logc = 2; clu_block = 256; }
        if (sweep_cluster_smem_bytes(nl, clu_logc) > budget) { clu_logc = 3; clu_block = 256; }
        if (g.clu_logc > 0) clu_logc = g.clu_logc - 1;
        if (g.clu_block > 0) clu_block = g.clu_block;
        if (sweep_cluster_smem_bytes(nl, clu_logc) > budget) {
            if (g.variant_forced == 4) return fail("sweep variant 4 forced but a level does not fit the cluster's shared memory");
            variant = 2;
        }
    }

    // host-side preparation of the grid-cooperative sweep (kept outside the timed region)
    int groups = 1;
    if (variant == 2) {
        if (!p.coldens_out) {
            // concurrent sources: one scratch grid per group, bounded by 1/4 of the free device memory
            const size_t per = sizeof(double) * (size_t)g.ncell;
            if (g.grid_max_groups == 0) {
                size_t free_b = 0, total_b = 0;
                CK(cudaMemGetInfo(&free_b, &total_b));
                g.grid_max_groups = (int)std::min<size_t>(ASORA_GRID_GROUPS_MAX, std::max<size_t>(1, (free_b / 4) / per));
            }
            groups = sweep_grid_groups(p, g.grid_max_groups, nullptr, nullptr);
            if (groups < 1) return fail("sweep_grid_kernel: no resident CTAs");
            if (groups > g.grid_scratch_groups) {
                if (g.grid_scratch) cudaFree(g.grid_scratch);
                g.grid_scratch = nullptr;
                g.grid_scratch_groups = 0;
                CK(cudaMalloc(&g.grid_scratch, per * groups));
                g.grid_scratch_groups = groups;
            }
            p.coldens_out = g.grid_scratch;
        }
        if (!g.grid_counters) CK(cudaMalloc(&g.grid_counters, sizeof(unsigned) * ASORA_GRID_GROUPS_MAX));
    }

    // Large shared-memory sweeps read the opacity and add the rates of their z-face cells through (k,i,j)-ordered
    // copies of the two grids: a third of all cells, whose warps otherwise touch 32 sectors per gather and per RED
    // (R = 30: 17.5 -> 16.4 ms with two extra transposing passes of 0.1 ms each).  Not for short sweeps, which would not
    // recover the passes, nor for slab-decomposed runs, whose plane ranges are not contiguous in the copies.
    const bool slab_active = g.slab_count > 0 && g.slab_count < N;
    const bool z_shape = (variant == 1 && S == 1 && block >= 768) || (variant == 3 && noct * plan->max_level_cells >= 2048 && !p.phi_heat &&
                                                                         sweep_octant_shape_ok(noct, opt, batch, block) == 2);
    const bool z_possible = z_shape && g.zface_ok && !coldens_grid && !slab_active && !g.deterministic;
    bool use_z = z_possible && plan->nlevels >= 24 &&
                 (double)count * (double)(plan->ncells / plan->parts) * (variant == 3 ? 8.0 : 1.0) >= 5e8;
    if (g.tune_opts & 8 && variant == 1) use_z = !use_z && z_possible;  // profiling knob
    if (g.oct_opts & 2 && variant == 3) use_z = !use_z && z_possible;  // profiling knob (mirror-image sweep)
    p.zface_offset = use_z ? (unsigned)g.ncell : 0u;

    // timed region (asora_last_sweep_stats: kernel_ms): nHI pre-pass, rate-grid zeroing, sweep
    CK(cudaEventRecord(g.ev0, g.stream));
    {
        Range prepass_range("asora:opacity_prepass");
        // whole grid, or the (periodic) range of planes this rank's sweeps can touch: at most two segments
        const int64_t plane = (int64_t)N * N;
        int64_t seg_off[2] = {0, 0}, seg_len[2] = {g.ncell, 0};
        if (g.slab_count > 0 && g.slab_count < N) {
            const int b = ((g.slab_begin % N) + N) % N;
            const int first = std::min(g.slab_count, N - b);
            seg_off[0] = b * plane;
            seg_len[0] = first * plane;
            seg_off[1] = 0;
            seg_len[1] = (g.slab_count - first) * plane;
        }
        if (use_z) {  // whole grid (no slab): opacity and its (k,i,j)-ordered copy in one pass
            cudaError_t e = launch_prepare_nhi_transposed(g.buf[ASORA_BUF_NDENS], g.buf[ASORA_BUF_XH_AV], g.nhi, g.nhi + g.ncell,
                                                          sig * dr, N, g.stream);
            if (e != cudaSuccess) return fail_cuda("prepare_nhi_transposed_kernel launch", e);
            g.last_launches += 1;
            CK(cudaMemsetAsync(g.buf[ASORA_BUF_PHI_ION] + g.ncell, 0, sizeof(double) * g.ncell, g.stream));
            if (p.phi_heat) CK(cudaMemsetAsync(p.phi_heat + g.ncell, 0, sizeof(double) * g.ncell, g.stream));
        }
        for (int sg = 0; sg < 2; sg++) {
            if (seg_len[sg] <= 0) continue;
            cudaError_t e = use_z ? cudaSuccess : launch_prepare_nhi(g.buf[ASORA_BUF_NDENS] + seg_off[sg], g.buf[ASORA_BUF_XH_AV] + seg_off[sg],
                                               g.nhi + seg_off[sg], sig * dr, seg_len[sg], g.stream);
            if (e != cudaSuccess) return fail_cuda("prepare_nhi_kernel launch", e);
            if (!use_z) g.last_launches += 1;
            // the sweep accumulates undivided sums in phi_ion (finish_cell); earlier rates wait in phi_keep
            if (!zero_phi) {
                if (!g.phi_keep) CK(cudaMalloc(&g.phi_keep, sizeof(double) * g.ncell));
                CK(cudaMemcpyAsync(g.phi_keep + seg_off[sg], g.buf[ASORA_BUF_PHI_ION] + seg_off[sg],
                                   sizeof(double) * seg_len[sg], cudaMemcpyDeviceToDevice, g.stream));
            }
            CK(cudaMemsetAsync(g.buf[ASORA_BUF_PHI_ION] + seg_off[sg], 0, sizeof(double) * seg_len[sg], g.stream));
            if (p.phi_heat) CK(cudaMemsetAsync(p.phi_heat + seg_off[sg], 0, sizeof(double) * seg_len[sg], g.stream));
            if (p.det_lo) CK(cudaMemsetAsync(p.det_lo + seg_off[sg], 0, sizeof(long long) * seg_len[sg], g.stream));
            if (p.det_lo_heat) CK(cudaMemsetAsync(p.det_lo_heat + seg_off[sg], 0, sizeof(long long) * seg_len[sg], g.stream));
        }
    }

    {
    Range sweep_range("asora:sweep");
    CK(cudaEventRecord(g.evk0, g.stream));
    if (variant == 1) {
        g.last_levels = plan->nlevels;
        cudaError_t e = launch_sweep_smem(*plan, p, S, block, opts, g.stream, &g.last_launches);
        if (e != cudaSuccess) return fail_cuda("sweep_smem_kernel launch", e);
    } else if (variant == 4) {
        cudaError_t e = launch_sweep_cluster(p, clu_logc, clu_block, g.stream, &g.last_launches, &g.last_levels);
        if (e != cudaSuccess) return fail_cuda("sweep_wedge_kernel launch", e);
    } else if (variant == 3) {
        g.last_levels = plan->nlevels;
        if (p.zface_offset && sweep_octant_smem_bytes(*plan, noct, (opts & 1) ? 8 : 1, true) > budget && noct == 8) opts &= ~1;
        cudaError_t e = launch_sweep_octant(*plan, p, noct, opt, batch, block, opts, g.stream, &g.last_launches);
        if (e != cudaSuccess) return fail_cuda("sweep_octant_kernel launch", e);
    } else {
        cudaError_t e = launch_sweep_grid(p, groups, g.grid_counters, g.stream, &g.last_launches, &g.last_levels);
        if (e != cudaSuccess) return fail_cuda("sweep_grid_kernel launch", e);
    }
    CK(cudaEventRecord(g.evk1, g.stream));
    }
    Range finish_range("asora:finish_phi");
    {   // phi = sum / ntau over the planes the sweep could touch
        const int64_t plane = (int64_t)N * N;
        int64_t seg_off[2] = {0, 0}, seg_len[2] = {g.ncell, 0};
        if (g.slab_count > 0 && g.slab_count < N) {
            const int b = ((g.slab_begin % N) + N) % N;
            const int first = std::min(g.slab_count, N - b);
            seg_off[0] = b * plane;
            seg_len[0] = first * plane;
            seg_len[1] = (g.slab_count - first) * plane;
        }
        if (use_z) {
            cudaError_t e = launch_finish_phi_transposed(g.buf[ASORA_BUF_PHI_ION], g.buf[ASORA_BUF_PHI_ION] + g.ncell, g.nhi,
                                                         zero_phi ? nullptr : g.phi_keep, N, g.stream);
            g.last_launches += 1;
            if (e == cudaSuccess && p.phi_heat) {
                e = launch_finish_phi_transposed(p.phi_heat, p.phi_heat + g.ncell, g.nhi, nullptr, N, g.stream);
                g.last_launches += 1;
            }
            if (e != cudaSuccess) return fail_cuda("finish_phi_transposed_kernel launch", e);
        }
        for (int sg = 0; sg < 2 && !use_z; sg++) {
            if (seg_len[sg] <= 0) continue;
            cudaError_t e;
            if (p.det_lo) {
                e = launch_finish_phi_fixed(g.buf[ASORA_BUF_PHI_ION] + seg_off[sg], p.det_lo + seg_off[sg], g.nhi + seg_off[sg],
                                            zero_phi ? nullptr : g.phi_keep + seg_off[sg], 1.0 / p.det_scale, seg_len[sg], g.stream);
                if (e == cudaSuccess && p.phi_heat) {
                    e = launch_finish_phi_fixed(p.phi_heat + seg_off[sg], p.det_lo_heat + seg_off[sg], g.nhi + seg_off[sg], nullptr,
                                                1.0 / p.det_scale_heat, seg_len[sg], g.stream);
                    g.last_launches += 1;
                }
            } else {
                e = launch_finish_phi(g.buf[ASORA_BUF_PHI_ION] + seg_off[sg], g.nhi + seg_off[sg],
                                      zero_phi ? nullptr : g.phi_keep + seg_off[sg], seg_len[sg], g.stream);
                if (e == cudaSuccess && p.phi_heat) {  // raytracing.f90:530: the heating rate is divided by nHI as well
                    e = launch_finish_phi(p.phi_heat + seg_off[sg], g.nhi + seg_off[sg], nullptr, seg_len[sg], g.stream);
                    g.last_launches += 1;
                }
            }
            if (e != cudaSuccess) return fail_cuda("finish_phi_kernel launch", e);
            g.last_launches += 1;
        }
    }
    if (coldens_grid) {
        // the kernels store optical depths; the debug interface promises column densities
        cudaError_t e = launch_scale_grid(coldens_grid, 1.0 / sig, g.ncell, g.stream);
        if (e != cudaSuccess) return fail_cuda("scale_grid_kernel launch", e);
    }
    CK(cudaEventRecord(g.ev1, g.stream));
    g.last_variant = variant;
    return 0;
}

}  // namespace

extern "C" {

const char* asora_last_error(void) { return g_err.c_str(); }
const char* asora_version(void) { return "asora_b200 0.1 (sm_100a)"; }

int64_t asora_cells_per_source(int N, double R) { return asora_count_cells(N, R); }

int asora_last_sweep_kernel_ms(float* kernel_only_ms)
{
    if (int rc = need_init()) return rc;
    if (!kernel_only_ms) return fail("last_sweep_kernel_ms: null pointer");
    CK(cudaEventSynchronize(g.evk1));
    CK(cudaEventElapsedTime(kernel_only_ms, g.evk0, g.evk1));
    return 0;
}

int asora_device_init(int N, int num_src_par)
{
    (void)num_src_par;
    if (N <= 0 || N > 1600) return fail("device_init: mesh size out of range");
    if (g.init) asora_device_close();
    CK(cudaGetDevice(&g.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g.device));
    if (prop.major < 10) return fail(std::string("device_init: ") + prop.name + " is not an sm_100 class GPU");
    g.sm_count = prop.multiProcessorCount;
    CK(cudaDeviceGetAttribute(&g.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, g.device));
    CK(cudaDeviceGetAttribute(&g.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, g.device));
    g.N = N;
    g.ncell = (int64_t)N * N * N;
    g.zface_ok = 2 * g.ncell < ((int64_t)1 << 32);
    CK(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    CK(cudaEventCreate(&g.ev0));
    CK(cudaEventCreate(&g.ev1));
    CK(cudaEventCreate(&g.evk0));
    CK(cudaEventCreate(&g.evk1));
    g.init = true;
    // the three grids the reference allocates up front (memory.cu:66-68)
    if (int rc = ensure_buffer(ASORA_BUF_NDENS)) return rc;
    if (int rc = ensure_buffer(ASORA_BUF_XH_AV)) return rc;
    if (int rc = ensure_buffer(ASORA_BUF_PHI_ION)) return rc;
    if (int rc = ensure_chem_scratch()) return rc;
    CK(cudaStreamSynchronize(g.stream));
    // the reference prints the device and the allocation (memory.cu:52-59,77-78); ASORA_QUIET=1 silences it
    const char* quiet = getenv("ASORA_QUIET");
    if (!(quiet && quiet[0] == '1')) {
        printf("GPU Device %d: \"%s\" with compute capability %d.%d\n", g.device, prop.name, prop.major, prop.minor);
        printf("Succesfully allocated %g Mb of device memory for grid of size N = %d (column densities stay on-chip)\n",
               3.0 * g.ncell * sizeof(double) / 1e6, N);
        fflush(stdout);
    }
    return 0;
}

int asora_device_close(void)
{
    if (!g.init) return 0;
    cudaStreamSynchronize(g.stream);
    for (int i = 0; i < ASORA_BUF_COUNT; i++) {
        if (g.buf[i]) cudaFree(g.buf[i]);
        g.buf[i] = nullptr;
    }
    free_tables();
    for (int t = 0; t < Context::kCopyThreads; t++) {
        for (int b = 0; b < 2; b++) {
            if (g.bounce[t][b]) cudaFreeHost(g.bounce[t][b]);
            if (g.copy_event[t][b]) cudaEventDestroy(g.copy_event[t][b]);
            g.bounce[t][b] = nullptr;
            g.copy_event[t][b] = nullptr;
        }
        if (g.copy_stream[t]) cudaStreamDestroy(g.copy_stream[t]);
        g.copy_stream[t] = nullptr;
    }
    if (g.nhi) cudaFree(g.nhi);
    if (g.phi_keep) cudaFree(g.phi_keep);
    g.phi_keep = nullptr;
    if (g.det_lo) cudaFree(g.det_lo);
    if (g.det_lo_heat) cudaFree(g.det_lo_heat);
    g.det_lo = g.det_lo_heat = nullptr;
    if (g.grid_scratch) cudaFree(g.grid_scratch);
    if (g.grid_counters) cudaFree(g.grid_counters);
    g.grid_scratch = nullptr;
    g.grid_counters = nullptr;
    g.grid_scratch_groups = 0;
    g.grid_max_groups = 0;
    g.cells_N = 0;
    g.slab_begin = g.slab_count = 0;
    g.grey = 0;
    if (g.log2_tab) cudaFree(g.log2_tab);
    g.nhi = nullptr;
    g.log2_tab = nullptr;
    if (g.src_pos) cudaFree(g.src_pos);
    if (g.src_flux) cudaFree(g.src_flux);
    if (g.src_pos_sorted) cudaFree(g.src_pos_sorted);
    if (g.src_flux_sorted) cudaFree(g.src_flux_sorted);
    g.src_pos_sorted = nullptr;
    g.src_flux_sorted = nullptr;
    if (g.chem_factors) cudaFree(g.chem_factors);
    g.chem_factors = nullptr;
    g.chem_factors_valid = false;
    if (g.chem_partials) cudaFree(g.chem_partials);
    if (g.chem_iparts) cudaFree(g.chem_iparts);
    for (int i = 0; i < 6; i++) {
        if (g.chem_stage[i]) cudaFree(g.chem_stage[i]);
        g.chem_stage[i] = nullptr;
    }
    g.chem_stage_n = 0;
    g.thin = g.thick = nullptr;
    g.src_pos = nullptr;
    g.src_flux = nullptr;
    g.chem_partials = nullptr;
    g.chem_iparts = nullptr;
    g.ntab = g.nsrc = 0;
    g.h_src_pos.clear();
    g.h_src_perm.clear();
    for (int i = 0; i < Context::kPlanCache; i++) free_sweep_plan(g.plans[i]);
    g.auto_parts.clear();
    if (g.ev0) cudaEventDestroy(g.ev0);
    if (g.ev1) cudaEventDestroy(g.ev1);
    if (g.evk0) cudaEventDestroy(g.evk0);
    if (g.evk1) cudaEventDestroy(g.evk1);
    g.evk0 = g.evk1 = nullptr;
    if (g.own_stream) cudaStreamDestroy(g.own_stream);
    g.ev0 = g.ev1 = nullptr;
    g.stream = g.own_stream = nullptr;
    g.init = false;
    return 0;
}

int asora_density_to_device(const double* ndens, int N)
{
    if (int rc = need_init()) return rc;
    if (N != g.N) return fail("density_to_device: N differs from device_init");
    if (!ndens) return fail("density_to_device: null pointer");
    return asora_buffer_upload(ASORA_BUF_NDENS, ndens);
}

int asora_photo_table_to_device(const double* thin_table, const double* thick_table, int NumTau)
{
    if (int rc = need_init()) return rc;
    if (NumTau < 2 || !thin_table || !thick_table) return fail("photo_table_to_device: bad arguments");
    free_tables();
    double* raw = nullptr;
    CK(cudaMalloc(&raw, sizeof(double) * 2 * (size_t)NumTau));
    CK(cudaMalloc(&g.thick, sizeof(double2) * 4 * (size_t)NumTau));  // thick, thin, heat thick, heat thin
    CK(cudaMemsetAsync(g.thick, 0, sizeof(double2) * 4 * (size_t)NumTau, g.stream));
    g.thin = g.thick + NumTau;
    g.heat_tables = false;
    CK(cudaMemcpyAsync(raw, thin_table, sizeof(double) * NumTau, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(raw + NumTau, thick_table, sizeof(double) * NumTau, cudaMemcpyHostToDevice, g.stream));
    cudaError_t e = launch_pair_table(raw, g.thin, NumTau, g.stream);
    if (e == cudaSuccess) e = launch_pair_table(raw + NumTau, g.thick, NumTau, g.stream);
    if (e != cudaSuccess) return fail_cuda("pair_table_kernel launch", e);
    CK(cudaStreamSynchronize(g.stream));
    cudaFree(raw);
    g.ntab = NumTau;
    g.table_max = 0.0;
    for (int i = 0; i < NumTau; i++) g.table_max = std::max(g.table_max, std::max(std::fabs(thin_table[i]) * 1e-7, std::fabs(thick_table[i])));
    {
        cudaResourceDesc rd;
        std::memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = (void*)g.thick;
        rd.res.linear.desc = cudaCreateChannelDesc<int4>();
        rd.res.linear.sizeInBytes = sizeof(double2) * 4 * (size_t)NumTau;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&g.tex_pairs, &rd, &td, nullptr));
    }
    return 0;
}

int asora_heat_table_to_device(const double* heat_thin_table, const double* heat_thick_table, int NumTau)
{
    if (int rc = need_init()) return rc;
    if (!g.thick) return fail("heat_table_to_device: call photo_table_to_device first");
    if (NumTau != g.ntab || !heat_thin_table || !heat_thick_table)
        return fail("heat_table_to_device: the heating tables must have the length of the photo tables");
    double* raw = nullptr;
    CK(cudaMalloc(&raw, sizeof(double) * 2 * (size_t)NumTau));
    CK(cudaMemcpyAsync(raw, heat_thin_table, sizeof(double) * NumTau, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(raw + NumTau, heat_thick_table, sizeof(double) * NumTau, cudaMemcpyHostToDevice, g.stream));
    cudaError_t e = launch_pair_table(raw + NumTau, g.thick + 2 * (size_t)NumTau, NumTau, g.stream);
    if (e == cudaSuccess) e = launch_pair_table(raw, g.thick + 3 * (size_t)NumTau, NumTau, g.stream);
    if (e != cudaSuccess) return fail_cuda("pair_table_kernel launch", e);
    CK(cudaStreamSynchronize(g.stream));
    cudaFree(raw);
    g.heat_table_max = 0.0;
    for (int i = 0; i < NumTau; i++)
        g.heat_table_max = std::max(g.heat_table_max, std::max(std::fabs(heat_thin_table[i]) * 1e-7, std::fabs(heat_thick_table[i])));
    g.heat_tables = true;
    return 0;
}

int asora_set_heating(int on)
{
    if (int rc = need_init()) return rc;
    if (on && !g.heat_tables) return fail("set_heating: no heating tables on device (heat_table_to_device)");
    g.heating = on != 0;
    return 0;
}

int asora_do_all_sources_heat(double R, double sig, double dr, const double* xh_av, double* phi_ion, double* phi_heat,
                              int NumSrc, int N, double minlogtau, double dlogtau, int NumTau)
{
    if (int rc = need_init()) return rc;
    if (N != g.N) return fail("do_all_sources_heat: m1 differs from device_init");
    if (!xh_av || !phi_ion || !phi_heat) return fail("do_all_sources_heat: null pointer");
    if (!g.heat_tables) return fail("do_all_sources_heat: no heating tables on device (heat_table_to_device)");
    const bool was = g.heating;
    g.heating = true;
    if (int rc = host_copy(g.buf[ASORA_BUF_XH_AV], xh_av, sizeof(double) * g.ncell, cudaMemcpyHostToDevice)) {
        g.heating = was;
        return rc;
    }
    const int rc = run_sweep(R, sig, dr, 0, NumSrc, minlogtau, dlogtau, NumTau, true, nullptr);
    g.heating = was;
    if (rc) return rc;
    if (int rc2 = host_copy(phi_ion, g.buf[ASORA_BUF_PHI_ION], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost)) return rc2;
    if (int rc2 = host_copy(phi_heat, g.buf[ASORA_BUF_PHI_HEAT], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost)) return rc2;
    cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1);
    return 0;
}

int asora_source_data_to_device(const int32_t* pos, const double* flux, int NumSrc)
{
    if (int rc = need_init()) return rc;
    if (NumSrc < 0 || (NumSrc > 0 && (!pos || !flux))) return fail("source_data_to_device: bad arguments");
    // The same positions as last time (a simulation keeps its sources over many time steps, evolve3D uploads them every
    // step): only the fluxes are refreshed, in upload order and through the remembered Morton permutation (the sort of 10^5
    // sources and the four allocations were ~10 ms of a 215 ms evolve3D call).
    if (NumSrc > 0 && NumSrc == g.nsrc && g.h_src_pos.size() == 3 * (size_t)NumSrc && g.h_src_perm.size() == (size_t)NumSrc &&
        std::memcmp(g.h_src_pos.data(), pos, sizeof(int32_t) * 3 * (size_t)NumSrc) == 0) {
        std::vector<double> sflux((size_t)NumSrc);
        g.flux_max = 0.0;
        for (int n = 0; n < NumSrc; n++) {
            sflux[n] = flux[g.h_src_perm[n]];
            g.flux_max = std::max(g.flux_max, std::fabs(flux[n]));
        }
        CK(cudaMemcpyAsync(g.src_flux, flux, sizeof(double) * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
        CK(cudaMemcpyAsync(g.src_flux_sorted, sflux.data(), sizeof(double) * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
        CK(cudaStreamSynchronize(g.stream));
        return 0;
    }
    g.h_src_pos.clear();
    g.h_src_perm.clear();
    if (g.src_pos) cudaFree(g.src_pos);
    if (g.src_flux) cudaFree(g.src_flux);
    if (g.src_pos_sorted) cudaFree(g.src_pos_sorted);
    if (g.src_flux_sorted) cudaFree(g.src_flux_sorted);
    g.src_pos = g.src_pos_sorted = nullptr;
    g.src_flux = g.src_flux_sorted = nullptr;
    g.nsrc = 0;
    if (NumSrc == 0) return 0;
    // The sweep is periodic (modulo_gpu, raytracing.cu:270-272): reduce positions to [0,N) once here.
    std::vector<int32_t> wrapped(3 * (size_t)NumSrc);
    for (size_t t = 0; t < wrapped.size(); t++) wrapped[t] = ((pos[t] % g.N) + g.N) % g.N;
    CK(cudaMalloc(&g.src_pos, sizeof(int32_t) * 3 * (size_t)NumSrc));
    CK(cudaMalloc(&g.src_flux, sizeof(double) * (size_t)NumSrc));
    CK(cudaMemcpyAsync(g.src_pos, wrapped.data(), sizeof(int32_t) * 3 * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.src_flux, flux, sizeof(double) * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
    // Morton-ordered copy
    std::vector<std::pair<uint64_t, int>> key((size_t)NumSrc);
    auto spread = [](uint64_t v) {  // 21 bits -> every third bit
        v &= 0x1fffff;
        v = (v | v << 32) & 0x1f00000000ffffULL;
        v = (v | v << 16) & 0x1f0000ff0000ffULL;
        v = (v | v << 8) & 0x100f00f00f00f00fULL;
        v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
        v = (v | v << 2) & 0x1249249249249249ULL;
        return v;
    };
    for (int n = 0; n < NumSrc; n++)
        key[n] = {spread(wrapped[3 * n]) << 2 | spread(wrapped[3 * n + 1]) << 1 | spread(wrapped[3 * n + 2]), n};
    std::sort(key.begin(), key.end());
    std::vector<int32_t> spos(3 * (size_t)NumSrc);
    std::vector<double> sflux((size_t)NumSrc);
    for (int n = 0; n < NumSrc; n++) {
        const int o = key[n].second;
        spos[3 * n] = wrapped[3 * o];
        spos[3 * n + 1] = wrapped[3 * o + 1];
        spos[3 * n + 2] = wrapped[3 * o + 2];
        sflux[n] = flux[o];
    }
    CK(cudaMalloc(&g.src_pos_sorted, sizeof(int32_t) * 3 * (size_t)NumSrc));
    CK(cudaMalloc(&g.src_flux_sorted, sizeof(double) * (size_t)NumSrc));
    CK(cudaMemcpyAsync(g.src_pos_sorted, spos.data(), sizeof(int32_t) * 3 * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemcpyAsync(g.src_flux_sorted, sflux.data(), sizeof(double) * (size_t)NumSrc, cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.nsrc = NumSrc;
    g.flux_max = 0.0;
    for (int n = 0; n < NumSrc; n++) g.flux_max = std::max(g.flux_max, std::fabs(flux[n]));
    g.h_src_pos.assign(pos, pos + 3 * (size_t)NumSrc);   // as passed (before the periodic reduction): compared next time
    g.h_src_perm.resize((size_t)NumSrc);
    for (int n = 0; n < NumSrc; n++) g.h_src_perm[n] = key[n].second;
    return 0;
}

int asora_do_all_sources(double R, double sig, double dr, const double* xh_av, double* phi_ion, int NumSrc,
                         int N, double minlogtau, double dlogtau, int NumTau)
{
    if (int rc = need_init()) return rc;
    if (N != g.N) return fail("do_all_sources: m1 differs from device_init");
    if (!xh_av || !phi_ion) return fail("do_all_sources: null pointer");
    if (int rc = host_copy(g.buf[ASORA_BUF_XH_AV], xh_av, sizeof(double) * g.ncell, cudaMemcpyHostToDevice)) return rc;
    if (int rc = run_sweep(R, sig, dr, 0, NumSrc, minlogtau, dlogtau, NumTau, true, nullptr)) return rc;
    if (int rc = host_copy(phi_ion, g.buf[ASORA_BUF_PHI_ION], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost)) return rc;
    cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1);
    return 0;
}

int asora_do_all_sources_begin(double R, double sig, double dr, const double* xh_av, int NumSrc, int N, double minlogtau,
                               double dlogtau, int NumTau)
{
    if (int rc = need_init()) return rc;
    if (N != g.N) return fail("do_all_sources_begin: m1 differs from device_init");
    // xh_av == NULL: ASORA_BUF_XH_AV already holds the fractions (a rank that received them from a peer over NVLink)
    if (xh_av)
        if (int rc = host_copy(g.buf[ASORA_BUF_XH_AV], xh_av, sizeof(double) * g.ncell, cudaMemcpyHostToDevice)) return rc;
    return run_sweep(R, sig, dr, 0, NumSrc, minlogtau, dlogtau, NumTau, true, nullptr);
}

int asora_do_all_sources_end(double* phi_ion)
{
    if (int rc = need_init()) return rc;
    if (phi_ion) {
        if (int rc = host_copy(phi_ion, g.buf[ASORA_BUF_PHI_ION], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost)) return rc;
    } else {
        CK(cudaStreamSynchronize(g.stream));
    }
    if (g.last_launches > 0) cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1);
    return 0;
}

int asora_debug_single_source(double R, double sig, double dr, const double* xh_av, int src_index,
                              double minlogtau, double dlogtau, int NumTau, double* coldensh_out, double* phi_ion)
{
    if (int rc = need_init()) return rc;
    if (!xh_av || !coldensh_out) return fail("debug_single_source: null pointer");
    if (int rc = ensure_buffer(ASORA_BUF_COLDENS)) return rc;
    CK(cudaMemcpyAsync(g.buf[ASORA_BUF_XH_AV], xh_av, sizeof(double) * g.ncell, cudaMemcpyHostToDevice, g.stream));
    CK(cudaMemsetAsync(g.buf[ASORA_BUF_COLDENS], 0, sizeof(double) * g.ncell, g.stream));
    if (int rc = run_sweep(R, sig, dr, src_index, 1, minlogtau, dlogtau, NumTau, true, g.buf[ASORA_BUF_COLDENS]))
        return rc;
    CK(cudaMemcpyAsync(coldensh_out, g.buf[ASORA_BUF_COLDENS], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost, g.stream));
    if (phi_ion)
        CK(cudaMemcpyAsync(phi_ion, g.buf[ASORA_BUF_PHI_ION], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1);
    return 0;
}

int asora_raytrace_device(double R, double sig, double dr, int src_begin, int src_count, double minlogtau,
                          double dlogtau, int NumTau, int zero_phi)
{
    if (int rc = need_init()) return rc;
    return run_sweep(R, sig, dr, src_begin, src_count, minlogtau, dlogtau, NumTau, zero_phi != 0, nullptr);
}

int asora_sync(void)
{
    if (int rc = need_init()) return rc;
    CK(cudaStreamSynchronize(g.stream));
    if (g.last_launches > 0) cudaEventElapsedTime(&g.last_ms, g.ev0, g.ev1);
    return 0;
}

int asora_set_stream(void* cuda_stream)
{
    if (int rc = need_init()) return rc;
    CK(cudaStreamSynchronize(g.stream));
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own_stream;
    return 0;
}

void* asora_device_buffer(int which)
{
    if (need_init()) return nullptr;
    if (ensure_buffer(which)) return nullptr;
    if (which == ASORA_BUF_TEMP) g.chem_factors_valid = false;  // the caller may write through the pointer
    cudaStreamSynchronize(g.stream);
    return g.buf[which];
}

int asora_buffer_upload(int which, const double* host)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    if (which == ASORA_BUF_TEMP) g.chem_factors_valid = false;
    return host_copy(g.buf[which], host, sizeof(double) * g.ncell, cudaMemcpyHostToDevice);
}

// Fortran-ordered host grids: staged through the opacity scratch (rebuilt before every sweep, so free here)
int asora_buffer_upload_f(int which, const double* host_fortran)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    if (!host_fortran) return fail("buffer_upload_f: null pointer");
    if (!g.nhi) CK(cudaMalloc(&g.nhi, sizeof(double) * g.ncell * (g.zface_ok ? 2 : 1)));
    if (which == ASORA_BUF_TEMP) g.chem_factors_valid = false;
    if (int rc = host_copy(g.nhi, host_fortran, sizeof(double) * g.ncell, cudaMemcpyHostToDevice)) return rc;
    cudaError_t e = launch_reverse_axes(g.nhi, g.buf[which], g.N, g.stream);
    if (e != cudaSuccess) return fail_cuda("reverse_axes_kernel launch", e);
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int asora_buffer_download_f(int which, double* host_fortran)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    if (!host_fortran) return fail("buffer_download_f: null pointer");
    if (!g.nhi) CK(cudaMalloc(&g.nhi, sizeof(double) * g.ncell * (g.zface_ok ? 2 : 1)));
    cudaError_t e = launch_reverse_axes(g.buf[which], g.nhi, g.N, g.stream);
    if (e != cudaSuccess) return fail_cuda("reverse_axes_kernel launch", e);
    return host_copy(host_fortran, g.nhi, sizeof(double) * g.ncell, cudaMemcpyDeviceToHost);
}

int asora_buffer_upload_range(int which, const double* host, int64_t cell_offset, int64_t cell_count)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    if (!host || cell_offset < 0 || cell_count < 0 || cell_offset + cell_count > g.ncell)
        return fail("buffer_upload_range: bad range");
    if (which == ASORA_BUF_TEMP) g.chem_factors_valid = false;
    if (cell_count > 0)
        CK(cudaMemcpyAsync(g.buf[which] + cell_offset, host + cell_offset, sizeof(double) * cell_count,
                           cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return 0;
}

int asora_buffer_download(int which, double* host)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    return host_copy(host, g.buf[which], sizeof(double) * g.ncell, cudaMemcpyDeviceToHost);
}

int asora_set_active_slab(int x_begin, int x_count)
{
    if (int rc = need_init()) return rc;
    if (x_count < 0) return fail("set_active_slab: negative plane count");
    g.slab_begin = x_begin;
    g.slab_count = (x_count >= g.N) ? 0 : x_count;
    return 0;
}

int asora_global_pass_device_range(double dt, double bh00, double albpow, double colh0, double temph0, double abu_c,
                                   int64_t cell_offset, int64_t cell_count, int* conv_flag, double* sum_xh1,
                                   double* sum_xh0)
{
    Range range("asora:chemistry");
    if (int rc = need_init()) return rc;
    if (cell_offset < 0 || cell_count < 0 || cell_offset + cell_count > g.ncell)
        return fail("global_pass_device_range: range outside the grid");
    const int ids[] = {ASORA_BUF_NDENS, ASORA_BUF_TEMP, ASORA_BUF_XH, ASORA_BUF_XH_AV, ASORA_BUF_XH_INTERMED,
                       ASORA_BUF_PHI_ION};
    for (int id : ids)
        if (int rc = ensure_buffer(id)) return rc;
    if (int rc = ensure_chem_scratch()) return rc;
    if (cell_count == 0) {
        if (conv_flag) *conv_flag = 0;
        if (sum_xh1) *sum_xh1 = 0.0;
        if (sum_xh0) *sum_xh0 = 0.0;
        return 0;
    }
    const int64_t o = cell_offset;
    if (int rc = ensure_chem_factors(bh00, albpow, colh0, temph0)) return rc;
    cudaError_t e = launch_global_pass(dt, g.buf[ASORA_BUF_NDENS] + o, g.buf[ASORA_BUF_TEMP] + o, g.chem_factors + o,
                                       g.buf[ASORA_BUF_XH] + o,
                                       g.buf[ASORA_BUF_XH_AV] + o, g.buf[ASORA_BUF_XH_INTERMED] + o,
                                       g.buf[ASORA_BUF_PHI_ION] + o, bh00, albpow, colh0, temph0, abu_c, cell_count, 1,
                                       g.chem_partials, g.chem_iparts, g.chem_blocks, conv_flag, sum_xh1, sum_xh0, g.stream);
    if (e != cudaSuccess) return fail_cuda("global_pass_kernel", e);
    return 0;
}

int asora_buffer_copy(int dst, int src)
{
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(dst)) return rc;
    if (int rc = ensure_buffer(src)) return rc;
    if (dst == ASORA_BUF_TEMP) g.chem_factors_valid = false;
    if (dst != src)
        CK(cudaMemcpyAsync(g.buf[dst], g.buf[src], sizeof(double) * g.ncell, cudaMemcpyDeviceToDevice, g.stream));
    return 0;
}

// ---- peer-memory halo exchange (slab-decomposed multi-GPU runs) ------------------------------------------------------
// A rank exports its grid buffers by CUDA IPC; its neighbours map them and read the halo planes straight over NVLink
// from a kernel, instead of four NCCL send/recv pairs per exchange.
int asora_ipc_export(int which, unsigned char* handle64)
{
    if (int rc = need_init()) return rc;
    if (!handle64) return fail("ipc_export: null handle buffer");
    if (int rc = ensure_buffer(which)) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, g.buf[which]));
    std::memcpy(handle64, &h, sizeof(h));
    return 0;
}

int asora_ipc_open(const unsigned char* handle64, void** dev_ptr)
{
    if (int rc = need_init()) return rc;
    if (!handle64 || !dev_ptr) return fail("ipc_open: null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int asora_ipc_close(void* dev_ptr)
{
    if (!dev_ptr) return 0;
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int asora_peer_halo(int which, const void* peer_buf, int64_t cell_offset, int64_t cell_count, int add)
{
    Range range(add ? "asora:peer_halo_add" : "asora:peer_halo_copy");
    if (int rc = need_init()) return rc;
    if (int rc = ensure_buffer(which)) return rc;
    if (!peer_buf) return fail("peer_halo: null peer buffer");
    if (cell_offset < 0 || cell_count < 0 || cell_offset + cell_count > g.ncell) return fail("peer_halo: range outside the grid");
    if (which == ASORA_BUF_TEMP) g.chem_factors_valid = false;
    cudaError_t e = launch_peer_halo(g.buf[which] + cell_offset, static_cast<const double*>(peer_buf) + cell_offset, cell_count,
                                     add != 0, g.stream);
    if (e != cudaSuccess) return fail_cuda("peer_halo_kernel launch", e);
    return 0;
}

int asora_global_pass_device(double dt, double bh00, double albpow, double colh0, double temph0, double abu_c,
                             int* conv_flag, double* sum_xh1, double* sum_xh0)
{
    Range range("asora:chemistry");
    if (int rc = need_init()) return rc;
    const int ids[] = {ASORA_BUF_NDENS, ASORA_BUF_TEMP, ASORA_BUF_XH, ASORA_BUF_XH_AV, ASORA_BUF_XH_INTERMED,
                       ASORA_BUF_PHI_ION};
    for (int id : ids)
        if (int rc = ensure_buffer(id)) return rc;
    if (int rc = ensure_chem_scratch()) return rc;
    if (int rc = ensure_chem_factors(bh00, albpow, colh0, temph0)) return rc;
    cudaError_t e = launch_global_pass(dt, g.buf[ASORA_BUF_NDENS], g.buf[ASORA_BUF_TEMP], g.chem_factors, g.buf[ASORA_BUF_XH],
                                       g.buf[ASORA_BUF_XH_AV], g.buf[ASORA_BUF_XH_INTERMED], g.buf[ASORA_BUF_PHI_ION],
                                       bh00, albpow, colh0, temph0, abu_c, g.ncell, 1, g.chem_partials,
                                       g.chem_iparts, g.chem_blocks, conv_flag, sum_xh1, sum_xh0, g.stream);
    if (e != cudaSuccess) return fail_cuda("global_pass_kernel", e);
    return 0;
}

// Host-buffer chemistry.  Works without device_init (hydrogenODE is usable stand-alone:
// pyc2ray/chemistry.py:43-97), on whatever device is current.
int asora_global_pass(double dt, const double* ndens, const double* temp, const double* xh, double* xh_av,
                      double* xh_intermed, const double* phi_ion, double bh00, double albpow, double colh0,
                      double temph0, double abu_c, int64_t ncell, int* conv_flag)
{
    Range range("asora:chemistry");
    if (ncell <= 0) return fail("global_pass: ncell must be positive");
    if (!ndens || !temp || !xh || !xh_av || !xh_intermed || !phi_ion) return fail("global_pass: null pointer");
    cudaStream_t st = g.init ? g.stream : (cudaStream_t)0;
    if (g.chem_stage_n < ncell) {
        for (int i = 0; i < 6; i++) {
            if (g.chem_stage[i]) cudaFree(g.chem_stage[i]);
            g.chem_stage[i] = nullptr;
        }
        g.chem_stage_n = 0;
        for (int i = 0; i < 6; i++) CK(cudaMalloc(&g.chem_stage[i], sizeof(double) * ncell));
        g.chem_stage_n = ncell;
    }
    if (int rc = ensure_chem_scratch()) return rc;
    const double* src[6] = {ndens, temp, xh, xh_av, xh_intermed, phi_ion};
    const size_t bytes = sizeof(double) * (size_t)ncell;
    // aliased host arrays share one device copy so the kernel sees the same aliasing
    double* dev[6];
    for (int i = 0; i < 6; i++) {
        dev[i] = g.chem_stage[i];
        for (int j = 0; j < i; j++)
            if (src[j] == src[i]) dev[i] = dev[j];
        if (dev[i] == g.chem_stage[i]) CK(cudaMemcpyAsync(dev[i], src[i], bytes, cudaMemcpyHostToDevice, st));
    }
    int flag = 0;
    cudaError_t e = launch_global_pass(dt, dev[0], dev[1], nullptr, dev[2], dev[3], dev[4], dev[5], bh00, albpow, colh0,
                                       temph0, abu_c, ncell, 1, g.chem_partials, g.chem_iparts, g.chem_blocks,
                                       &flag, nullptr, nullptr, st);
    if (e != cudaSuccess) return fail_cuda("global_pass_kernel", e);
    // chemistry.f90:107-108 write-back; xh_av first, xh_intermed last (see chemistry.cu)
    CK(cudaMemcpyAsync(xh_av, dev[3], bytes, cudaMemcpyDeviceToHost, st));
    if (xh_intermed != xh_av) CK(cudaMemcpyAsync(xh_intermed, dev[4], bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (conv_flag) *conv_flag = flag;
    return 0;
}

int asora_invalidate_temperature(void)
{
    g.chem_factors_valid = false;
    return 0;
}

int asora_set_sweep_variant(int variant)
{
    if (variant < 0 || variant > 4) return fail("set_sweep_variant: unknown variant");
    g.variant_forced = variant;
    return 0;
}

int asora_set_octant_shape(int octants_per_cta, int images_per_thread, int batch, int block_threads)
{
    if (octants_per_cta != 0 || images_per_thread != 0 || batch != 0 || block_threads != 0) {
        const int no = octants_per_cta ? octants_per_cta : 8;
        if (!(no == 8 || no == 4 || no == 2)) return fail("set_octant_shape: octants per CTA must be 8, 4 or 2");
    }
    g.oct_noct = octants_per_cta;
    g.oct_opt = images_per_thread;
    g.oct_batch = batch;
    g.oct_opts = (block_threads >> 16) & 0xff;
    g.oct_block = block_threads & 0xffff;
    return 0;
}

int asora_set_cluster_shape(int ctas_per_cluster, int block_threads)
{
    int logc = 0;
    if (ctas_per_cluster != 0) {
        if (!(ctas_per_cluster == 1 || ctas_per_cluster == 2 || ctas_per_cluster == 4 || ctas_per_cluster == 8))
            return fail("set_cluster_shape: CTAs per cluster must be 1, 2, 4 or 8");
        while ((1 << logc) < ctas_per_cluster) logc++;
        logc += 1;  // stored biased by one: 0 = automatic
    }
    g.clu_logc = logc;
    g.clu_block = block_threads;
    return 0;
}

int asora_plan_builds(void) { return g.plan_builds; }

int64_t asora_plan_export(int N, double R, double dr, int sphere_only, int octant, int parts, int64_t capacity,
                          double* path, double* inv_np, uint16_t* upstream_slots, uint8_t* offsets, uint8_t* flags,
                          uint8_t* minor_ab, int* level_start, int* level_mid, int* info)
{
    SweepPlan plan;
    std::string err;
    const bool ok = octant ? build_octant_plan(plan, N, R, dr, sphere_only != 0, err, false)
                           : build_sweep_plan(plan, N, R, dr, sphere_only != 0, parts, err, false);
    if (!ok) {
        fail(err);
        return -1;
    }
    const int64_t n = (int64_t)plan.cells.size();
    if (info) {
        info[0] = plan.nlevels;
        info[1] = plan.max_level_cells;
        info[2] = plan.lo;
        info[3] = plan.side;
        info[4] = plan.q_max;
        info[5] = plan.parts;
    }
    if (capacity >= n) {
        for (int64_t e = 0; e < n; e++) {
            const PlanCell& c = plan.cells[e];
            if (path) path[e] = c.path;
            if (inv_np) inv_np[e] = c.inv_np;
            if (upstream_slots) for (int t = 0; t < 4; t++) upstream_slots[4 * e + t] = c.nb[t];
            if (offsets) for (int t = 0; t < 3; t++) offsets[3 * e + t] = c.d[t];
            if (flags) flags[e] = c.flags;
            if (minor_ab) {
                minor_ab[2 * e] = (uint8_t)(c.ab & 0xff);
                minor_ab[2 * e + 1] = (uint8_t)((c.ab >> 8) & 0xff);
            }
        }
        if (level_start) std::copy(plan.level_start.begin(), plan.level_start.end(), level_start);
        if (level_mid) std::copy(plan.level_mid.begin(), plan.level_mid.end(), level_mid);
    }
    return n;
}

int asora_set_grey_notables(int on)
{
    if (int rc = need_init()) return rc;
    g.grey = on ? 1 : 0;
    return 0;
}

int asora_set_deterministic(int on)
{
    g.deterministic = on ? 1 : 0;
    return 0;
}

int asora_set_sphere_only(int sphere_only)
{
    g.sphere_only = sphere_only ? 1 : 0;
    return 0;
}

int asora_set_tuning(int sources_per_cta, int block_threads)
{
    // bits 16-18 of block_threads toggle launch options of the shared-memory sweep (profiling knob): 16 = copies of
    // the log2 table, 17 = table gathers through the texture pipe, 18 = offsets word fetched one cell ahead
    g.tune_opts = (block_threads >> 16) & 15;  // bit 19 toggles the z-face copies
    g.tune_parts = (block_threads >> 20) & 15; // bits 20-23: parts per source (1, 2, 4, 8), 0 = automatic
    block_threads &= 0xffff;
    if (!(sources_per_cta == 0 || sources_per_cta == 1 || sources_per_cta == 2))
        return fail("set_tuning: sources_per_cta must be 0, 1 or 2");
    if (block_threads < 0 || block_threads > 1024 || block_threads % 32) return fail("set_tuning: bad block size");
    g.tune_S = sources_per_cta;
    g.tune_block = block_threads;
    return 0;
}

int asora_last_sweep_stats(int* variant, int* launches, int64_t* updates, int* q_max, int* levels, float* kernel_ms)
{
    if (variant) *variant = g.last_variant;
    if (launches) *launches = g.last_launches;
    if (updates) {
        if (g.last_updates < 0) {
            if (!(g.cells_N == g.N && g.cells_R == g.last_R && g.cells_sphere_only == g.last_sphere_only &&
                  (!g.last_sphere_only || g.cells_dr == g.last_dr))) {
                g.cells_per_source = g.last_sphere_only ? asora_count_rated_cells(g.N, g.last_R, g.last_dr)
                                                        : asora_count_cells(g.N, g.last_R);
                g.cells_N = g.N;
                g.cells_R = g.last_R;
                g.cells_dr = g.last_dr;
                g.cells_sphere_only = g.last_sphere_only;
            }
            g.last_updates = (int64_t)g.last_count * g.cells_per_source;
        }
        *updates = g.last_updates;
    }
    if (q_max) *q_max = g.last_qmax;
    if (levels) *levels = g.last_levels;
    if (kernel_ms) *kernel_ms = g.last_ms;
    return 0;
}

}  // extern "C"
