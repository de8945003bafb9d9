// sweep_kernels.cu -- the ASORA sweep for sm_100a: column-density propagation by Chebyshev levels,
// photo-ionisation table lookup and rate accumulation.
//
// Replaces evolve0D_gpu + cinterp_gpu (src/asora/raytracing.cu:155-535) and photoion_rates_gpu +
// photo_lookuptable (src/asora/rates.cu:16-83).  Two variants share the per-cell arithmetic:
//
//   sweep_smem_kernel<S>  one CTA sweeps S sources at once.  The outgoing column densities of the
//                         previous and the current level live in shared memory (2 x S x max level
//                         cells doubles); nothing per-source is ever written to HBM.  All geometry
//                         comes from the plan (sweep_plan.cu), amortised over the S sources.
//   sweep_grid_kernel     cooperative launch, the whole GPU sweeps one source at a time, one
//                         grid-wide barrier per level; column densities go through an N^3 scratch
//                         grid that stays in the 126 MB L2.  Geometry is computed on the fly.  Used
//                         when a level does not fit in shared memory (large radii / full box).
//
// Arithmetic.  The first B200 profile of a literal transcription (profiles/r01a_*) showed ~440
// thread-instructions per source-cell update, dominated by the CUDA library's fp64 division (nine per
// update) and log10 (two per update) call sequences, with the fp64 pipe only 33-40 % busy.  The
// per-cell arithmetic below is the same real-number computation re-associated (now ~185 instructions):
//   * the four weightf() divisions and the normalisation of raytracing.cu:422-428 become one
//     reciprocal of products (interp_coldens);
//   * strength/Vfact (rates.cu:24) is a product with the plan's 1/(n path); phi/nHI (raytracing.cu:324) is the
//     same for every source and is applied once per cell after the sweep (finish_phi_kernel);
//   * log10(tau) -> table index (rates.cu:77-78) is index = a + b*log2(tau) with log2 from a
//     256-entry mantissa table in shared memory and a degree-5 polynomial (|error| < 2 ulp);
//   * T[i0] + r*(T[i1]-T[i0]) reads one 16-byte {T[i], T[i+1]-T[i]} pair (through the texture pipe);
//   * int <-> double conversions by 2^52-offset additions (fp64 pipe) instead of I2F / F2I (XU pipe).
// Results agree with the reference's own expression order to ~1e-13 relative (tests/).
#include "asora_common.cuh"
#include "sweep_device.cuh"

#include <cooperative_groups.h>
#include <cmath>
namespace cg = cooperative_groups;

// 1/m for the Chebyshev levels m = 1..255 (entry 0 is 0): a constant-bank operand the compiler can re-read instead
// of holding it in registers across the cell loop
__constant__ double kInvLevel[256];

cudaError_t upload_inv_levels()
{
    double h[256];
    h[0] = 0.0;
    for (int m = 1; m < 256; m++) h[m] = 1.0 / (double)m;
    return cudaMemcpyToSymbol(kInvLevel, h, sizeof(h));
}

void host_log2_table(double* tab512)
{
    for (int j = 0; j < 256; j++) {
        const long double c = 1.0L + ((long double)j + 0.5L) / 256.0L;
        const double inv = (double)(1.0L / c);
        tab512[2 * j + 0] = inv;
        tab512[2 * j + 1] = (double)(-log2l((long double)inv));  // log2 of the value 1/inv actually used
    }
}

// ---------------------------------------------------------------------------------------------------
// variant 1: shared-memory level sweep, S sources per CTA
// ---------------------------------------------------------------------------------------------------
// One plan cell with its per-source grid positions and opacities already fetched.
template <int S>
struct Fetched {
    double wA, wB, path, inv_np;
    int nb1, nb2, nb3, nb4;
    unsigned flags;
    unsigned pos[S];
    double nhi[S];
};

// Plan cell: two coalesced 16-byte read-only loads, shared by the S sources of the CTA.  The interpolation
// fractions wA = a/m, wB = b/m are rebuilt from the byte-sized offsets: the dominant offset of every cell of
// Chebyshev level m is m itself, so 1/m is a per-level constant and one correction step gives the correctly
// rounded quotient (verified exhaustively for 0 <= a <= m <= 255) -- 8 arithmetic instructions instead of a
// third 512-byte load per warp on the L1 data pipe.
// wrapX/Y/Z: per-source shared-memory tables, indexed by the biased offset, holding the periodic
// cell coordinate already multiplied by its stride (N*N, N, 1): one LDS per axis replaces the
// add / sign-fix / compare / select chain of modulo_gpu (raytracing.cu:24,270-272).  A second set of tables serves
// the cells in the interior of a level's z faces (PC_ZFACE): there consecutive cells step in j, a stride of N doubles in
// the (i,j,k) grids, so a warp's opacity gather and its RED touch 32 sectors each; the second set addresses (k,i,j)-ordered
// copies of the two grids (strides N, 1, N*N plus the offset of the copies), where the same warp touches two runs of 16
// consecutive doubles.  Only the ZF instantiations (one CTA per SM shapes of large radii) carry this.
// PF: `dword` ({d[3], flags}, the third word of the second stream) was fetched one cell ahead from its own 4-byte
// stream, so the opacity gather starts together with the plan loads instead of after them.
template <int S, bool PF, bool ZF>
__device__ __forceinline__ void fetch_cell(Fetched<S>& f, const int4* __restrict__ plan, int ncells, int e, unsigned dword,
                                           double md, double inv_m, const unsigned* __restrict__ wrap_tab, int side,
                                           const double* __restrict__ nhi)
{
    if (PF) {
        const unsigned di = dword & 0xff, dj = (dword >> 8) & 0xff, dk = (dword >> 16) & 0xff;
        const unsigned* wz = ZF ? wrap_tab + ((dword >> 28) & 1u) * (3 * S * side) : wrap_tab;  // PC_ZFACE: second set
#pragma unroll
        for (int s = 0; s < S; s++) {
            const unsigned* w = wz + 3 * s * side;
            f.pos[s] = w[di] + w[side + dj] + w[2 * side + dk];
            f.nhi[s] = __ldg(nhi + f.pos[s]);
        }
    }
    const int4 rb = __ldg(plan + (size_t)ncells + e);
    const int4 ra = __ldg(plan + e);
    f.path = __hiloint2double(ra.y, ra.x);
    f.inv_np = __hiloint2double(ra.w, ra.z);
    f.nb1 = rb.x & 0xffff;
    f.nb2 = (unsigned)rb.x >> 16;
    f.nb3 = rb.y & 0xffff;
    f.nb4 = (unsigned)rb.y >> 16;
    f.flags = ((unsigned)rb.z >> 24) & 0xffu;
    const double da = u2d(rb.w & 0xff), db = u2d((rb.w >> 8) & 0xff);
    const double qa = da * inv_m, qb = db * inv_m;
    f.wA = fma(fma(-md, qa, da), inv_m, qa);
    f.wB = fma(fma(-md, qb, db), inv_m, qb);
    if (!PF) {
        const unsigned di = rb.z & 0xff, dj = (rb.z >> 8) & 0xff, dk = (rb.z >> 16) & 0xff;
        const unsigned* wz = ZF ? wrap_tab + (((unsigned)rb.z >> 28) & 1u) * (3 * S * side) : wrap_tab;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const unsigned* w = wz + 3 * s * side;
            f.pos[s] = w[di] + w[side + dj] + w[2 * side + dk];  // N <= 1600: fits 32 bits
            f.nhi[s] = __ldg(nhi + f.pos[s]);
        }
    }
}

// One level of the sweep for the S sources of a CTA.
template <int S, int BLOCK, int REP, bool DIAG, bool CDOUT, bool TEX, bool PF, bool HEAT, bool ZF, bool DET>
__device__ __forceinline__ void sweep_level(const int4* __restrict__ plan, const unsigned* __restrict__ dwords, int ncells,
                                            int beg, int end, int m, double* __restrict__ cur,
                                            const double* __restrict__ prev, int max_level_cells,
                                            const unsigned* __restrict__ wrap_tab, int side, const double (&sk)[S],
                                            const bool (&live)[S], const SweepParams& p,
                                            const double2* __restrict__ log2_tab)
{
    const double md = (double)m, inv_m = kInvLevel[m];
    int e = beg + threadIdx.x;
    unsigned dword = 0;
    if (PF && e < end) dword = __ldg(dwords + e);
    for (; e < end; e += BLOCK) {
        unsigned dnext = 0;
        if (PF && e + BLOCK < end) dnext = __ldg(dwords + e + BLOCK);
        Fetched<S> c;
        fetch_cell<S, PF, ZF>(c, plan, ncells, e, dword, md, inv_m, wrap_tab, side, p.nhi);
        const int slot = e - beg;
#pragma unroll
        for (int s = 0; s < S; s++) {
            if (!live[s]) continue;
            const double* pv = prev + s * max_level_cells;
            const double cin = interp_coldens<false, DIAG>(pv[c.nb1], pv[c.nb2], pv[c.nb3], pv[c.nb4], c.wA, c.wB, c.flags);
            const double cdho = finish_cell<REP, TEX, HEAT, DET>(cin, c.path, c.inv_np, c.flags, c.nhi[s], sk[s], c.pos[s], p, log2_tab);
            cur[s * max_level_cells + slot] = cdho;
            if (CDOUT) p.coldens_out[c.pos[s]] = cdho;
        }
        dword = dnext;
    }
}

// One CTA per S sources.  Per level every thread updates its cells (plan entry + opacity from global memory,
// four upstream optical depths from the previous level's shared-memory buffer), then a CTA barrier hands
// the level over.  Latency-hiding and staging variants that were measured on B200 and dropped ("Roofline"):
// the whole next cell held in registers across the barrier (-30 %), the 16-byte offsets stream one cell ahead at
// 64 registers (-10...-40 %, spills), prefetch.global.L1 of the next plan entry (-4 %), a split arrive/wait level
// barrier (-12 %), the plan staged through shared memory by per-warp bulk copies (-19 %: fewer stalls, but 20 %
// more instructions).  What is kept: the 4-byte offsets word one cell ahead (PF) where registers allow it.
template <int S, int BLOCK, int MINB, int REP, bool CDOUT, bool TEX, bool PF, bool HEAT, bool ZF, bool DET>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_smem_kernel(const int4* __restrict__ plan, const unsigned* __restrict__ dwords, int ncells,
                  const int* __restrict__ level_start_all,
                                  int nlevels, int max_level_cells, int lo, int side, int parts, SweepParams p)
{
    extern __shared__ double2 sh_raw[];
    double2* log2_all = sh_raw;                                        // 256 entries x REP copies
    double* sh_cd = reinterpret_cast<double*>(sh_raw + 256 * REP);     // [2][S][max_level_cells]
    unsigned* wrap_tab = reinterpret_cast<unsigned*>(sh_cd + (size_t)2 * S * max_level_cells);  // [2][S][3][side]
    const int N = p.N;
    // CTA -> (part of the sweep, group of S sources), part-major: the parts of one source have identical work and
    // would run in lock-step if they were co-resident; CTAs of different sources drift apart, so that one CTA's
    // barrier waits and partially filled last passes overlap another CTA's arithmetic.
    const int ngroups = gridDim.x / parts;
    const int part = blockIdx.x / ngroups;
    const int first = (blockIdx.x - part * ngroups) * S;
    int* level_start = reinterpret_cast<int*>(wrap_tab + (size_t)2 * 3 * S * side);  // [nlevels + 1], shared memory
    for (int t = threadIdx.x; t <= nlevels; t += BLOCK) level_start[t] = __ldg(level_start_all + part * (nlevels + 1) + t);

    for (int t = threadIdx.x; t < 256 * REP; t += BLOCK) log2_all[t] = __ldg(p.log2_tab + t / REP);
    const double2* log2_tab = log2_all + (REP > 1 ? (threadIdx.x & (REP - 1)) : 0);
    // the zero slot (last of every level buffer, sweep_plan.cu: resolve_zero_slot): zero-weight corners and the source cell
    if (threadIdx.x < 2 * S) sh_cd[(size_t)(threadIdx.x + 1) * max_level_cells - 1] = 0.0;

    int i0[S], j0[S], k0[S];
    double sk[S];
    bool live[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        live[s] = (first + s) < p.src_count;
        const int ns = p.src_begin + (live[s] ? first + s : 0);
        i0[s] = p.src_pos[3 * ns + 0];
        j0[s] = p.src_pos[3 * ns + 1];
        k0[s] = p.src_pos[3 * ns + 2];
        sk[s] = p.src_flux[ns] * p.kpref;
    }
#pragma unroll
    for (int s = 0; s < S; s++)
        for (int t = threadIdx.x; t < 3 * side; t += BLOCK) {
            const int axis = t / side, d = t - axis * side + lo;
            const int c0 = axis == 0 ? i0[s] : (axis == 1 ? j0[s] : k0[s]);
            const unsigned NN = (unsigned)N * N, w = (unsigned)wrap(c0 + d, N);
            const unsigned stride = axis == 0 ? NN : (axis == 1 ? (unsigned)N : 1u);
            // (k,i,j) order: i*N + j + k*N*N, behind the (i,j,k) grid
            const unsigned stride_t = axis == 0 ? (unsigned)N : (axis == 1 ? 1u : NN);
            wrap_tab[3 * s * side + t] = w * stride;
            if (ZF) wrap_tab[3 * S * side + 3 * s * side + t] = w * stride_t + (axis == 0 ? p.zface_offset : 0u);
        }
    __syncthreads();

    int beg = level_start[0], end = level_start[1];
    for (int m = 0; m < nlevels; m++) {
        // bounds of the next level, read before this level's work so that nothing but the barrier separates two levels
        const int next_end = level_start[min(m + 2, nlevels)];
        double* cur = sh_cd + (size_t)(m & 1) * S * max_level_cells;
        const double* prev = sh_cd + (size_t)((m & 1) ^ 1) * S * max_level_cells;
        if (m < 2)
            sweep_level<S, BLOCK, REP, true, CDOUT, TEX, PF, HEAT, ZF, DET>(plan, dwords, ncells, beg, end, m, cur, prev, max_level_cells,
                                                             wrap_tab, side, sk, live, p, log2_tab);
        else
            sweep_level<S, BLOCK, REP, false, CDOUT, TEX, PF, HEAT, ZF, DET>(plan, dwords, ncells, beg, end, m, cur, prev, max_level_cells,
                                                              wrap_tab, side, sk, live, p, log2_tab);
        __syncthreads();
        beg = end;
        end = next_end;
    }
}

size_t sweep_smem_bytes(const SweepPlan& plan, int S, int rep)
{
    return (size_t)256 * rep * sizeof(double2) + (size_t)2 * S * plan.max_level_cells * sizeof(double) +
           (size_t)2 * 3 * S * plan.side * sizeof(unsigned) + (size_t)(plan.nlevels + 1) * sizeof(int);
}

template <int S, int BLOCK, int MINB, int REP, bool CDOUT, bool TEX, bool PF, bool HEAT = false, bool ZF = false, bool DET = false>
static cudaError_t launch_smem_t(const SweepPlan& plan, const SweepParams& p, cudaStream_t stream)
{
    const size_t smem = sweep_smem_bytes(plan, S, REP);
    const int grid = ((p.src_count + S - 1) / S) * plan.parts;
    auto kernel = sweep_smem_kernel<S, BLOCK, MINB, REP, CDOUT, TEX, PF, HEAT, ZF, DET>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, BLOCK, smem, stream>>>(plan.d_cells, plan.d_dwords, (int)plan.ncells, plan.d_level_start, plan.nlevels,
                                          plan.max_level_cells, plan.lo, plan.side, plan.parts, p);
    return cudaGetLastError();
}

template <int S, int BLOCK, int MINB>
static cudaError_t launch_smem_opts(const SweepPlan& plan, const SweepParams& p, int opts, cudaStream_t stream)
{
    if (p.coldens_out) return launch_smem_t<S, BLOCK, MINB, 1, true, false, false>(plan, p, stream);  // debug path
    if (p.det_lo) {  // deterministic accumulation: one option set per shape (no z-face copies)
        if (p.phi_heat) return launch_smem_t<S, BLOCK, MINB, 1, false, true, false, true, false, true>(plan, p, stream);
        return launch_smem_t<S, BLOCK, MINB, 1, false, true, false, false, false, true>(plan, p, stream);
    }
    if (p.zface_offset) {  // z-face cells through the (k,i,j)-ordered copies: the one-CTA-per-SM shapes only
        if constexpr (S == 1 && BLOCK >= 768) {
            if (p.phi_heat) return launch_smem_t<S, BLOCK, MINB, 1, false, true, true, true, true>(plan, p, stream);
            return (opts & 1) ? launch_smem_t<S, BLOCK, MINB, 8, false, true, true, false, true>(plan, p, stream)
                              : launch_smem_t<S, BLOCK, MINB, 1, false, true, true, false, true>(plan, p, stream);
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (p.phi_heat) return launch_smem_t<S, BLOCK, MINB, 1, false, true, false, true>(plan, p, stream);  // with heating
    switch (opts & 7) {
        case 0: return launch_smem_t<S, BLOCK, MINB, 1, false, false, false>(plan, p, stream);
        case 1: return launch_smem_t<S, BLOCK, MINB, 8, false, false, false>(plan, p, stream);
        case 2: return launch_smem_t<S, BLOCK, MINB, 1, false, true, false>(plan, p, stream);
        case 3: return launch_smem_t<S, BLOCK, MINB, 8, false, true, false>(plan, p, stream);
        case 4: return launch_smem_t<S, BLOCK, MINB, 1, false, false, true>(plan, p, stream);
        case 5: return launch_smem_t<S, BLOCK, MINB, 8, false, false, true>(plan, p, stream);
        case 6: return launch_smem_t<S, BLOCK, MINB, 1, false, true, true>(plan, p, stream);
        default: return launch_smem_t<S, BLOCK, MINB, 8, false, true, true>(plan, p, stream);
    }
}

// `block` in {256,512,1024}; `opts`: bit 0 eight copies of the log2 table in shared memory, bit 1 table gathers
// through the texture pipe, bit 2 offsets word fetched one cell ahead.
cudaError_t launch_sweep_smem(const SweepPlan& plan, const SweepParams& p, int S, int block, int opts,
                              cudaStream_t stream, int* launches)
{
    if (p.src_count <= 0) return cudaSuccess;
    if (launches) *launches += 1;
#define ASORA_CASE(SS, BB, MB) \
    if (S == SS && block == BB) return launch_smem_opts<SS, BB, MB>(plan, p, opts, stream);
    ASORA_CASE(1, 256, 4)
    ASORA_CASE(1, 512, 2)
    ASORA_CASE(1, 1024, 1)
    ASORA_CASE(2, 256, 3)  // 80 registers: at 64 the two-source body spills and its speed depends on where
    ASORA_CASE(1, 768, 1)
    ASORA_CASE(1, 896, 1)
#undef ASORA_CASE
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------
// variant 2: grid-cooperative level sweep, whole GPU per source
// ---------------------------------------------------------------------------------------------------

// Enumerate the cube shell max(|di|,|dj|,|dk|) == m, k fastest on the x and y faces.  A shell has 24 m^2 + 2 <= 1.6e7
// cells (m <= 800): everything fits 32 bits, and the divisions by the face width are 32-bit unsigned ones.
__device__ __forceinline__ void shell_cell(unsigned t, int m, int& di, int& dj, int& dk)
{
    const unsigned w = 2u * m + 1u, v = 2u * m - 1u;
    const unsigned fx = w * w, fy = v * w, fz = v * v;
    if (t < 2 * fx) {
        const int sgn = (t < fx) ? 1 : -1;
        if (t >= fx) t -= fx;
        const unsigned q = t / w;
        di = sgn * m;
        dj = (int)q - m;
        dk = (int)(t - q * w) - m;
    } else if (t < 2 * fx + 2 * fy) {
        t -= 2 * fx;
        const int sgn = (t < fy) ? 1 : -1;
        if (t >= fy) t -= fy;
        const unsigned q = t / w;
        dj = sgn * m;
        di = (int)q - (m - 1);
        dk = (int)(t - q * w) - m;
    } else {
        t -= 2 * fx + 2 * fy;
        const int sgn = (t < fz) ? 1 : -1;
        if (t >= fz) t -= fz;
        const unsigned q = t / v;
        dk = sgn * m;
        di = (int)q - (m - 1);
        dj = (int)(t - q * v) - (m - 1);
    }
}

// sqrt(x) for a normal positive x: MUFU.RSQ64H seed, two Newton steps on 1/sqrt(x), one Heron correction (<= 1 ulp).
__device__ __forceinline__ double fast_sqrt(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    const double s = x * r;
    return fma(fma(-s, s, x), 0.5 * r, s);
}

__device__ __forceinline__ int isign1(int x) { return x >= 0 ? 1 : -1; }

// Barrier among the `nctas` co-resident CTAs of one group (the launch is cooperative, so all CTAs of the
// grid are resident).  Same structure as cooperative_groups' grid.sync(): CTA barrier, one thread
// publishes with a fence + atomic and spins on the group's monotonically increasing counter, fence,
// CTA barrier.  `epoch` counts the barriers this CTA has passed.
__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned nctas, unsigned& epoch)
{
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        // the counter only ever grows; comparing the wrapped difference keeps the barrier correct after 2^32 arrivals
        // (many sources on few groups at large meshes)
        const unsigned target = epoch * nctas;
        while ((int)(*(volatile unsigned*)counter - target) < 0) {
        }
        __threadfence();
    }
    __syncthreads();
}

// The grid is split into `ngroups` groups of `group_ctas` CTAs; group g sweeps sources g, g+ngroups, ...
// through its own N^3 scratch grid, so that sources whose levels are much narrower than the GPU run
// side by side.  ngroups == 1 is "the whole GPU per source".
template <bool HEAT, bool DET, bool GREY = false>
__global__ void __launch_bounds__(512, 2)
sweep_grid_kernel(SweepParams p, int nlevels, int ngroups, int group_ctas, unsigned* counters)
{
    __shared__ double2 log2_tab[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) log2_tab[t] = __ldg(p.log2_tab + t);
    __syncthreads();
    const int N = p.N;
    const int group = blockIdx.x / group_ctas;
    if (group >= ngroups) return;  // spare CTAs of an uneven split
    const unsigned nthreads = (unsigned)group_ctas * blockDim.x;
    const unsigned tid = (unsigned)(blockIdx.x - group * group_ctas) * blockDim.x + threadIdx.x;
    // integer squared distances this far from R^2 decide the sphere test without evaluating the reference's expression
    const double R2_lo = p.R2 * (1.0 - 1e-12), R2_hi = p.R2 * (1.0 + 1e-12);
    double* __restrict__ slab = p.coldens_out + (size_t)group * N * N * N;
    unsigned* counter = counters + group;
    unsigned epoch = 0;

    for (int sidx = group; sidx < p.src_count; sidx += ngroups) {
        const int ns = p.src_begin + sidx;
        const int i0 = p.src_pos[3 * ns + 0], j0 = p.src_pos[3 * ns + 1], k0 = p.src_pos[3 * ns + 2];
        const double sk = p.src_flux[ns] * p.kpref;
        for (int m = 0; m < nlevels; m++) {
            const unsigned ncell = (m == 0) ? 1u : 24u * m * m + 2u;
            // the dominant offset of every cell of level m is m: one division per level instead of five per cell
            const double md = (double)m, inv_m = (m > 0) ? 1.0 / md : 0.0;
            for (unsigned t = tid; t < ncell; t += nthreads) {
                int di = 0, dj = 0, dk = 0;
                if (m > 0) shell_cell(t, m, di, dj, dk);
                const int ia = abs(di), ja = abs(dj), ka = abs(dk);
                if (ia + ja + ka > p.q_max) continue;
                if (di < p.last_l || di > p.last_r || dj < p.last_l || dj > p.last_r || dk < p.last_l ||
                    dk > p.last_r)
                    continue;
                unsigned flags = 0;
                const double dn = (double)(ia * ia + ja * ja + ka * ka);
                if (m > 0) {
                    bool rated;
                    if (dn <= R2_lo) {
                        rated = true;
                    } else if (dn >= R2_hi) {
                        rated = false;
                    } else {
                        // on the sphere's surface to rounding: the reference's own form decides (raytracing.cu:302-305,315)
                        const double xs = p.dr * (double)di, ys = p.dr * (double)dj, zs = p.dr * (double)dk;
                        const double dist2 = __fma_rn(zs, zs, __fma_rn(ys, ys, __dmul_rn(xs, xs)));
                        rated = dist2 / (p.dr * p.dr) <= p.R2;
                    }
                    if (rated) flags |= PC_RATED;
                    else if (p.sphere_only) continue;
                }
                const int i = wrap(i0 + di, N), j = wrap(j0 + dj, N), k = wrap(k0 + dk, N);
                const size_t pos = ((size_t)i * N + j) * N + k;
                const double nHI_p = __ldg(p.nhi + pos);
                double cin = 0.0, path = 0.5, inv_np = ASORA_FOURPI;
                if (m == 0) {
                    flags = PC_SOURCE | PC_RATED;
                } else {
                    // upstream cell coordinates (periodic): one step towards the source on each axis
                    const int im = wrap(i - isign1(di), N), jm = wrap(j - isign1(dj), N),
                              km = wrap(k - isign1(dk), N);
                    int a, b, c;
                    size_t q1, q2, q3, q4;
                    const size_t NN = (size_t)N * N;
                    if (ka >= ja && ka >= ia) {  // raytracing.cu:394
                        a = ia; b = ja; c = ka;
                        q1 = im * NN + (size_t)jm * N + km; q2 = i * NN + (size_t)jm * N + km;
                        q3 = im * NN + (size_t)j * N + km;  q4 = i * NN + (size_t)j * N + km;
                    } else if (ja >= ia && ja >= ka) {  // raytracing.cu:446
                        a = ia; b = ka; c = ja;
                        q1 = im * NN + (size_t)jm * N + km; q2 = i * NN + (size_t)jm * N + km;
                        q3 = im * NN + (size_t)jm * N + k;  q4 = i * NN + (size_t)jm * N + k;
                    } else {  // raytracing.cu:491
                        a = ja; b = ka; c = ia;
                        q1 = im * NN + (size_t)jm * N + km; q2 = im * NN + (size_t)j * N + km;
                        q3 = im * NN + (size_t)jm * N + k;  q4 = im * NN + (size_t)j * N + k;
                    }
                    // c == m.  wA = a/m, wB = b/m correctly rounded (reciprocal + one correction step, as in the plan-driven
                    // sweep); path = sqrt(1 + (a^2+b^2)/c^2) = sqrt(n)/m (raytracing.cu:444); 1/(n path) by reciprocal
                    const double da = (double)a, db = (double)b;
                    const double qa = da * inv_m, qb = db * inv_m;
                    const double wA = fma(fma(-md, qa, da), inv_m, qa), wB = fma(fma(-md, qb, db), inv_m, qb);
                    const double sn = fast_sqrt(dn), qp = sn * inv_m;
                    path = fma(fma(-md, qp, sn), inv_m, qp);
                    inv_np = fast_rcp(dn * path);
                    if (c == 1 && (a == 1 || b == 1)) flags |= (a == 1 && b == 1) ? PC_DIAG3 : PC_DIAG2;
                    // the scratch grid is written by other SMs: read it at L2 (ld.global.cg), and skip
                    // zero-weight corners, which may never have been written for this source
                    const double c1 = (wA * wB != 0.0) ? __ldcg(slab + q1) : 0.0;
                    const double c2 = (wB * (1.0 - wA) != 0.0) ? __ldcg(slab + q2) : 0.0;
                    const double c3 = (wA * (1.0 - wB) != 0.0) ? __ldcg(slab + q3) : 0.0;
                    const double c4 = ((1.0 - wA) * (1.0 - wB) != 0.0) ? __ldcg(slab + q4) : 0.0;
                    cin = interp_coldens<true, true>(c1, c2, c3, c4, wA, wB, flags);
                }
                const double cdho = finish_cell<1, false, HEAT, DET, GREY>(cin, path, inv_np, flags, nHI_p, sk, pos, p, log2_tab);
                __stcg(slab + pos, cdho);
            }
            group_barrier(counter, (unsigned)group_ctas, epoch);
        }
    }
}

#ifndef ASORA_GRID_MIN_CTAS
#define ASORA_GRID_MIN_CTAS 4
#endif
// Number of concurrent sources (groups) for the grid-cooperative sweep, at most `max_groups` (scratch grids).
int sweep_grid_groups(const SweepParams& p, int max_groups, int* total_ctas_out, int* group_ctas_out)
{
    const int block = 512;
    static int total_cached = 0;  // resident CTAs of the cooperative launch (device property, queried once)
    if (total_cached == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        int per_sm_grey = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_grid_kernel<true, true>, block, 0) != cudaSuccess) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_grey, sweep_grid_kernel<false, false, true>, block, 0) != cudaSuccess) return 0;
        per_sm = min(per_sm, per_sm_grey);
        if (per_sm < 1) return 0;
        total_cached = sms * per_sm;
    }
    const int total = total_cached;
    // A level costs one pass of latency (~3 us) per ceil(cells / threads) plus one barrier (~2 us); more
    // groups amortise the barrier over more sources (measured: scripts/perf_probe3.py), so take as many as
    // there are scratch grids and sources, keeping at least 8 CTAs per group.
    int groups = max(1, min(max_groups, p.src_count));
    while (groups > 1 && total / groups < ASORA_GRID_MIN_CTAS) groups--;
    if (total_ctas_out) *total_ctas_out = total;
    if (group_ctas_out) *group_ctas_out = total / groups;
    return groups;
}

cudaError_t launch_sweep_grid(const SweepParams& p, int ngroups, unsigned* counters, cudaStream_t stream,
                              int* launches, int* levels)
{
    if (p.src_count <= 0) return cudaSuccess;
    const int block = 512;
    int total = 0, group_ctas = 0;
    int check = sweep_grid_groups(p, ngroups, &total, &group_ctas);
    if (check < 1) return cudaErrorLaunchOutOfResources;
    group_ctas = total / ngroups;
    // levels 0..min(q_max, max(|last_l|, last_r))
    int nlevels = min(p.q_max, max(-p.last_l, p.last_r)) + 1;
    // sphere-only: no cell beyond Chebyshev distance floor(R) can be inside the sphere
    if (p.sphere_only && sqrt(p.R2) + 2.0 < (double)nlevels) nlevels = (int)sqrt(p.R2) + 2;
    if (levels) *levels = nlevels;
    cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned) * ngroups, stream);
    if (e != cudaSuccess) return e;
    SweepParams pc = p;
    void* args[] = {(void*)&pc, (void*)&nlevels, (void*)&ngroups, (void*)&group_ctas, (void*)&counters};
    if (launches) *launches += 1;
    void* kernel = pc.det_lo ? (pc.phi_heat ? (void*)sweep_grid_kernel<true, true> : (void*)sweep_grid_kernel<false, true>)
                             : (pc.phi_heat ? (void*)sweep_grid_kernel<true, false> : (void*)sweep_grid_kernel<false, false>);
    if (pc.grey) {
        if (pc.det_lo || pc.phi_heat) return cudaErrorNotSupported;
        kernel = (void*)sweep_grid_kernel<false, false, true>;
    }
    return cudaLaunchCooperativeKernel(kernel, dim3(total), dim3(block), args, 0, stream);
}

// ---------------------------------------------------------------------------------------------------
// small preparation kernels
// ---------------------------------------------------------------------------------------------------

// Cell opacity per unit path in cell units, ntau = ndens * (1 - xh_av) * sigma * dr (raytracing.cu:275-276,311),
// once per sweep instead of once per (source, cell): halves the scattered loads of the sweep and removes the
// sigma / dr multiplications from it.
__global__ void prepare_nhi_kernel(const double* __restrict__ ndens, const double* __restrict__ xh_av,
                                   double* __restrict__ ntau, double sig_dr, int64_t ncell)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (int64_t)gridDim.x * blockDim.x)
        ntau[i] = (ndens[i] * (1.0 - xh_av[i])) * sig_dr;
}

cudaError_t launch_prepare_nhi(const double* ndens, const double* xh_av, double* ntau, double sig_dr, int64_t ncell,
                               cudaStream_t stream)
{
    int64_t blocks = (ncell + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    prepare_nhi_kernel<<<(int)blocks, 256, 0, stream>>>(ndens, xh_av, ntau, sig_dr, ncell);
    return cudaGetLastError();
}

// phi = sum / ntau: the division by the cell's own opacity that every (source, cell) rate shares
// (raytracing.cu:324, finish_cell above).  A cell no source deposited into keeps 0 also where it is fully ionised
// (the reference leaves 0 in unvisited cells and 0/0 in visited ones, SURVEY note N7).  `keep` (optional) holds
// rates of earlier sweeps to add on top.
__global__ void finish_phi_kernel(double* __restrict__ phi, const double* __restrict__ ntau,
                                  const double* __restrict__ keep, int64_t ncell)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (int64_t)gridDim.x * blockDim.x) {
        const double sum = phi[i];
        double v = (sum != 0.0) ? sum / ntau[i] : 0.0;
        if (keep) v += keep[i];
        phi[i] = v;
    }
}

// The division pass of a deterministic sweep: sum = (hi * 2^46 + lo) / 2^s from the two integer grids (the high parts sit
// in the rate grid itself), then as above.
__global__ void finish_phi_fixed_kernel(double* __restrict__ phi_hi, const long long* __restrict__ lo, const double* __restrict__ ntau,
                                        const double* __restrict__ keep, double inv_scale, int64_t ncell)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (int64_t)gridDim.x * blockDim.x) {
        const long long hi = reinterpret_cast<const long long*>(phi_hi)[i];
        const double sum = ((double)hi * (double)(1ll << ASORA_DET_LOW_BITS) + (double)lo[i]) * inv_scale;
        double v = (sum != 0.0) ? sum / ntau[i] : 0.0;
        if (keep) v += keep[i];
        phi_hi[i] = v;
    }
}

cudaError_t launch_finish_phi_fixed(double* phi_hi, const long long* lo, const double* ntau, const double* keep, double inv_scale,
                                    int64_t ncell, cudaStream_t stream)
{
    int64_t blocks = (ncell + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    finish_phi_fixed_kernel<<<(int)blocks, 256, 0, stream>>>(phi_hi, lo, ntau, keep, inv_scale, ncell);
    return cudaGetLastError();
}

cudaError_t launch_finish_phi(double* phi, const double* ntau, const double* keep, int64_t ncell, cudaStream_t stream)
{
    int64_t blocks = (ncell + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    finish_phi_kernel<<<(int)blocks, 256, 0, stream>>>(phi, ntau, keep, ncell);
    return cudaGetLastError();
}

// The opacity grid and its (k,i,j)-ordered copy ntau_t[k*N*N + i*N + j] through a 32x32 shared-memory tile, so that both
// are written with coalesced stores.
__global__ void prepare_nhi_transposed_kernel(const double* __restrict__ ndens, const double* __restrict__ xh_av,
                                              double* __restrict__ ntau, double* __restrict__ ntau_t, double sig_dr, int N)
{
    __shared__ double tile[32][33];
    const int i = blockIdx.z;
    const int j0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, k = k0 + threadIdx.x;
        if (j < N && k < N) {
            const size_t idx = ((size_t)i * N + j) * N + k;
            const double v = (ndens[idx] * (1.0 - xh_av[idx])) * sig_dr;
            ntau[idx] = v;
            tile[r][threadIdx.x] = v;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, j = j0 + threadIdx.x;
        if (j < N && k < N) ntau_t[((size_t)k * N + i) * N + j] = tile[threadIdx.x][r];
    }
}

cudaError_t launch_prepare_nhi_transposed(const double* ndens, const double* xh_av, double* ntau, double* ntau_t, double sig_dr,
                                          int N, cudaStream_t stream)
{
    dim3 grid((N + 31) / 32, (N + 31) / 32, N), block(32, 8);
    prepare_nhi_transposed_kernel<<<grid, block, 0, stream>>>(ndens, xh_av, ntau, ntau_t, sig_dr, N);
    return cudaGetLastError();
}

// finish_phi_kernel for a sweep that accumulated its z-face cells in the (k,i,j)-ordered copy phi_t:
// phi[i][j][k] = (phi[i][j][k] + phi_t[k][i][j]) / ntau[i][j][k] (+ keep).
__global__ void finish_phi_transposed_kernel(double* __restrict__ phi, const double* __restrict__ phi_t,
                                             const double* __restrict__ ntau, const double* __restrict__ keep, int N)
{
    __shared__ double tile[32][33];
    const int i = blockIdx.z;
    const int j0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, j = j0 + threadIdx.x;
        if (j < N && k < N) tile[r][threadIdx.x] = phi_t[((size_t)k * N + i) * N + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, k = k0 + threadIdx.x;
        if (j < N && k < N) {
            const size_t idx = ((size_t)i * N + j) * N + k;
            const double sum = phi[idx] + tile[threadIdx.x][r];
            double v = (sum != 0.0) ? sum / ntau[idx] : 0.0;
            if (keep) v += keep[idx];
            phi[idx] = v;
        }
    }
}

cudaError_t launch_finish_phi_transposed(double* phi, const double* phi_t, const double* ntau, const double* keep, int N,
                                         cudaStream_t stream)
{
    dim3 grid((N + 31) / 32, (N + 31) / 32, N), block(32, 8);
    finish_phi_transposed_kernel<<<grid, block, 0, stream>>>(phi, phi_t, ntau, keep, N);
    return cudaGetLastError();
}

// grid *= factor (debug path: optical depths back to column densities)
__global__ void scale_grid_kernel(double* __restrict__ grid, double factor, int64_t ncell)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (int64_t)gridDim.x * blockDim.x)
        grid[i] *= factor;
}

cudaError_t launch_scale_grid(double* grid, double factor, int64_t ncell, cudaStream_t stream)
{
    int64_t blocks = (ncell + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_grid_kernel<<<(int)blocks, 256, 0, stream>>>(grid, factor, ncell);
    return cudaGetLastError();
}

// {T[i], T[i+1]-T[i]} pairs; the last slope is 0 (index clamp, see photo_lookup)
__global__ void pair_table_kernel(const double* __restrict__ table, double2* __restrict__ pairs, int ntab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntab) {
        const double t0 = table[i];
        const double t1 = (i + 1 < ntab) ? table[i + 1] : t0;
        pairs[i] = make_double2(t0, t1 - t0);
    }
}

cudaError_t launch_pair_table(const double* table, double2* pairs, int ntab, cudaStream_t stream)
{
    pair_table_kernel<<<(ntab + 255) / 256, 256, 0, stream>>>(table, pairs, ntab);
    return cudaGetLastError();
}

// Axis reversal between Fortran order (i fastest) and the device layout (k fastest): for every j the (i,k) plane is
// a 2-D transpose.  in[k*N*N + j*N + i] -> out[i*N*N + j*N + k]; applying it twice is the identity, so the same
// kernel serves both directions.
__global__ void reverse_axes_kernel(const double* __restrict__ in, double* __restrict__ out, int N)
{
    __shared__ double tile[32][33];
    const int j = blockIdx.z;
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, i = i0 + threadIdx.x;
        if (i < N && k < N) tile[r][threadIdx.x] = in[((size_t)k * N + j) * N + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, k = k0 + threadIdx.x;
        if (i < N && k < N) out[((size_t)i * N + j) * N + k] = tile[threadIdx.x][r];
    }
}

cudaError_t launch_reverse_axes(const double* in, double* out, int N, cudaStream_t stream)
{
    dim3 grid((N + 31) / 32, (N + 31) / 32, N), block(32, 8);
    reverse_axes_kernel<<<grid, block, 0, stream>>>(in, out, N);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// halo exchange of a slab-decomposed run over peer memory (NVLink): dst[i] (+)= src[i], src being the same buffer of a
// neighbouring rank's GPU, mapped into this process by CUDA IPC (asora_ipc_open).  16-byte accesses when both are aligned.
// ---------------------------------------------------------------------------------------------------
template <bool ADD>
__global__ void peer_halo_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n)
{
    const int64_t n2 = n >> 1;
    double2* d2 = reinterpret_cast<double2*>(dst);
    const double2* s2 = reinterpret_cast<const double2*>(src);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = s2[i];
        if (ADD) {
            double2 o = d2[i];
            o.x += v.x;
            o.y += v.y;
            d2[i] = o;
        } else {
            d2[i] = v;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = ADD ? dst[n - 1] + src[n - 1] : src[n - 1];
}

cudaError_t launch_peer_halo(double* dst, const double* src, int64_t n, bool add, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) return cudaErrorMisalignedAddress;
    int64_t blocks = (n / 2 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (add)
        peer_halo_kernel<true><<<(int)blocks, 256, 0, stream>>>(dst, src, n);
    else
        peer_halo_kernel<false><<<(int)blocks, 256, 0, stream>>>(dst, src, n);
    return cudaGetLastError();
}
