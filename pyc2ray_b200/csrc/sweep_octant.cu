// sweep_octant.cu -- the mirror-image sweep: one plan entry per thread, applied to up to eight octants.
//
// Replaces evolve0D_gpu + cinterp_gpu (src/asora/raytracing.cu:155-535) and photoion_rates_gpu + photo_lookuptable
// (src/asora/rates.cu:16-83) for sweeps whose cell set is mirror-symmetric (-lo == hi, see build_octant_plan in
// sweep_plan.cu).  Same levels, same per-cell arithmetic (sweep_device.cuh) and same results as sweep_smem_kernel;
// what changes is the work of a thread.  There, a thread fetched and decoded one 32-byte plan entry, rebuilt the
// interpolation fractions and bilinear weights and looked up the periodic cell position for ONE cell: 36 bytes of
// plan traffic through L1 and ~45 of its ~194 instructions per update were spent on data that is the same for the
// eight cells (+-di, +-dj, +-dk) (profiles/r01k).  Here a thread does that once per entry of the positive octant and
// then evaluates the OPT mirror images it is responsible for: per image one position (three adds of wrap-table values
// already in registers), one opacity gather, four upstream optical depths from that image's level buffer, the
// interpolation, the table lookups and the rate deposit (120-140 instructions per update, profiles/r02b).
//
// Two phases per round of BATCH images, both straight-line: phase 1 propagates the optical depth (position, opacity,
// interpolation, store into the level buffer), phase 2 evaluates the rates (two logarithms, two table gathers, one
// RED) of the same images side by side, so that their table-gather latencies overlap.  (A variant that parked phase 2
// in a per-thread shared-memory queue and evaluated it while waiting at the level barrier was measured and dropped:
// DESIGN.md, "Mirror-image sweep".)
//
// Work split.  A CTA sweeps NOCT of the eight octants of one source (NOCT = 8: the whole source; 4, 2: half-spaces
// or quadrants as separate CTAs, several of which then fit on one SM); a thread handles OPT of them (OPT divides
// NOCT; NOCT/OPT warps share an entry).  Octant bits: bit 2 = sign of di, bit 1 = dj, bit 0 = dk (1 = negative), the
// part-index convention of build_sweep_plan.
//
// Plane cells.  An entry with a zero offset is its own mirror image on that axis.  Entries without a zero offset on
// the axes a thread iterates itself ("class A", a prefix of every level) get all OPT images.  For the others
// ("class B") the thread evaluates only the OPT/2 images with a clear bit on one of the zero axes and stores each
// result into the level buffers of both images that border the plane; images that differ in a bit owned by another
// thread or another CTA, and the few cells with two zero offsets, are recomputed.  The rate is deposited once: by the
// image whose bits are clear on every zero axis.
#include "asora_common.cuh"
#include "sweep_device.cuh"

#ifndef ASORA_OCT_TU
#define ASORA_OCT_TU 0
#endif

namespace {

#define ASORA_NO_DEPOSIT 0xffffffffu

// Plan entries are streamed: every CTA reads each entry of a level once.  Loading them without allocating in L1 leaves
// the L1 to the opacity gathers, which neighbouring sources on the same SM do share.
__device__ __forceinline__ int4 load_plan(const int4* __restrict__ ptr)
{
#ifdef ASORA_PLAN_L1_ALLOCATE
    return __ldg(ptr);
#else
    int4 v;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
    return v;
#endif
}

// Phase 2 for NB cells side by side: rates.cu:16-41 + the deposit of raytracing.cu:324-329 (finish_cell_pre in
// sweep_device.cuh is the one-cell form).  pos[u] == ASORA_NO_DEPOSIT: nothing to deposit for that cell.
template <int NB, int REP, bool TEX, bool HEAT, bool DET>
__device__ __forceinline__ void rate_cells(double (&tin)[NB], double (&tout)[NB], const double (&skn)[NB], const unsigned (&pos)[NB],
                                           const SweepParams& p, const double2* __restrict__ log2_tab)
{
    double dtau[NB];
    unsigned out_of_range = 0;
#pragma unroll
    for (int u = 0; u < NB; u++) {
        dtau[u] = tout[u] - tin[u];
        const int h1 = __double2hiint(tin[u]), h2 = __double2hiint(tout[u]);
        out_of_range |= (unsigned)((unsigned)(h1 - p.hi_min) >= p.hi_span) | (unsigned)((unsigned)(h2 - p.hi_min) >= p.hi_span);
    }
    if (__builtin_expect(out_of_range != 0, 0)) {  // the source cell, fully ionised paths, beyond the table
#pragma unroll
        for (int u = 0; u < NB; u++) {
            tin[u] = clamp_tau_slow(tin[u], p.tau_lo, p.tau_hi);
            tout[u] = clamp_tau_slow(tout[u], p.tau_lo, p.tau_hi);
        }
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
        const bool thick = fabs(dtau[u]) > ASORA_TAU_PHOTO_LIMIT;
        const TableIndex a = table_index<REP>(__double2hiint(tin[u]), __double2loint(tin[u]), p, log2_tab);
        const TableIndex b = table_index<REP>(__double2hiint(tout[u]), __double2loint(tout[u]), p, log2_tab);
        const int ib = b.i0 + (thick ? 0 : p.ntab);  // thin table right behind the thick one
        double t_in, t_out, h_in = 0.0, h_out = 0.0;
        if (TEX) {
            const int4 ua = tex1Dfetch<int4>(p.tex_pairs, a.i0);
            const int4 ub = tex1Dfetch<int4>(p.tex_pairs, ib);
            t_in = fma(a.residual, __hiloint2double(ua.w, ua.z), __hiloint2double(ua.y, ua.x));
            t_out = fma(b.residual, __hiloint2double(ub.w, ub.z), __hiloint2double(ub.y, ub.x));
            if (HEAT) {
                const int4 va = tex1Dfetch<int4>(p.tex_pairs, a.i0 + 2 * p.ntab);
                const int4 vb = tex1Dfetch<int4>(p.tex_pairs, ib + 2 * p.ntab);
                h_in = fma(a.residual, __hiloint2double(va.w, va.z), __hiloint2double(va.y, va.x));
                h_out = fma(b.residual, __hiloint2double(vb.w, vb.z), __hiloint2double(vb.y, vb.x));
            }
        } else {
            const double2 ta = __ldg(p.thick + a.i0), tb = __ldg(p.thick + ib);
            t_in = fma(a.residual, ta.y, ta.x);
            t_out = fma(b.residual, tb.y, tb.x);
            if (HEAT) {
                const double2 va = __ldg(p.thick + a.i0 + 2 * p.ntab), vb = __ldg(p.thick + ib + 2 * p.ntab);
                h_in = fma(a.residual, va.y, va.x);
                h_out = fma(b.residual, vb.y, vb.x);
            }
        }
        // rates.cu:28-38: thick cells absorb T(tau_in) - T(tau_out), thin cells dtau * T_thin(tau_out)
        const double absorbed = thick ? (t_in - t_out) : dtau[u] * t_out;
        const bool deposit = pos[u] != ASORA_NO_DEPOSIT;
        if (deposit) deposit_rate<DET>(p.phi_ion, p.det_lo, p.det_scale, pos[u], skn[u] * absorbed);  // RED at L2
        if (HEAT) {  // photorates.f90:118,124 with the table argument convention of rates.cu:37 (tau_out for thin cells)
            const double heated = thick ? (h_in - h_out) : dtau[u] * h_out;
            if (deposit) deposit_rate<DET>(p.phi_heat, p.det_lo_heat, p.det_scale_heat, pos[u], skn[u] * heated);
        }
    }
}

// The NIMG images of one plan entry that a thread evaluates.
//   X[s], Y[s], Z[s]: periodic cell coordinate times its stride for the offset +d (s = 0) and -d (s = 1)
//   CLASS_B: NIMG = OPT / 2 images, the ones with a clear bit `zb` (a zero axis of the entry inside the thread's own
//            bits); every result is stored for the image across that plane as well
//   BATCH:   images evaluated side by side, NIMG / BATCH rounds: the straight-line code of a round keeps ~30 registers
//            per image live, so the batch is what fits the register budget of the launch shape
//   THK:     every optical depth of the previous level is >= 0.6 (level-uniform, octant_level), so max(0.6, tau) of
//            raytracing.cu:33 is tau itself and the weighted mean collapses to prod(tau) / sum_i s_i prod_{j != i} tau_j
//   ok:      cleared when an optical depth below 0.6 is stored (feeds the next level's THK decision)
template <int NIMG, int BATCH, int OPT, bool CLASS_B, int REP, bool DIAG, bool TEX, bool HEAT, bool DET, bool THK>
__device__ __forceinline__ void entry_images(const int4 ra, const int4 rb, double md, double inv_m, int slot, int lmax,
                                             int obase /* first local octant of this thread */, int gbase /* the same, global */,
                                             const unsigned (&X)[2], const unsigned (&Y)[2], const unsigned (&Z)[2],
                                             double* __restrict__ cur, const double* __restrict__ prev, double sk,
                                             const SweepParams& p, const double2* __restrict__ log2_tab, bool& ok)
{
    const double path = __hiloint2double(ra.y, ra.x);
    const double skn = sk * __hiloint2double(ra.w, ra.z);  // strength * kpref / (n path)
    const int nb1 = rb.x & 0xffff, nb2 = (unsigned)rb.x >> 16, nb3 = rb.y & 0xffff, nb4 = (unsigned)rb.y >> 16;
    const unsigned flags = ((unsigned)rb.z >> 24) & 0x1fu;
    const unsigned zmask = (unsigned)rb.z >> 29;  // zero offsets: bit 2 = di, bit 1 = dj, bit 0 = dk
    // interpolation fractions a/m, b/m from the byte-sized offsets: reciprocal of the level + one correction step
    // (correctly rounded for 0 <= a <= m <= 255, as in sweep_kernels.cu: fetch_cell), then the bilinear weights
    const double da = u2d(rb.w & 0xff), db = u2d((rb.w >> 8) & 0xff);
    const double qa = da * inv_m, qb = db * inv_m;
    const double wA = fma(fma(-md, qa, da), inv_m, qa);
    const double wB = fma(fma(-md, qb, db), inv_m, qb);
    const double uA = 1.0 - wA, uB = 1.0 - wB;
    const double s1 = wA * wB, s2 = wB * uA, s3 = wA * uB, s4 = uA * uB;
    double diag = 1.0;
    if (DIAG) diag = (flags & PC_DIAG3) ? ASORA_SQRT3 : ((flags & PC_DIAG2) ? ASORA_SQRT2 : 1.0);
    unsigned zb = 0;
    if (CLASS_B) {
        const unsigned zm = zmask & (unsigned)(OPT - 1);  // != 0 for a class B entry
        zb = (zm & 4u) ? 4u : ((zm & 2u) ? 2u : 1u);
    }
    const bool rated = (flags & PC_RATED) != 0;
    constexpr int NB = BATCH < NIMG ? BATCH : NIMG;
#pragma unroll
    for (int u0 = 0; u0 < NIMG; u0 += NB) {
        // ---- phase 1: propagate the optical depth into every image ---------------------------------------------
        unsigned pos[NB];
        int tt[NB];
        double tin[NB], tout[NB];
        bool overflow = false;
#pragma unroll
        for (int v = 0; v < NB; v++) {
            const int u = u0 + v;
            int t;
            unsigned px, py, pz;
            if (CLASS_B) {  // insert a clear bit at position zb
                const unsigned low = (unsigned)u & (zb - 1u);
                t = (int)((((unsigned)u - low) << 1) | low);
                px = (OPT >= 8 && (t & 4)) ? X[1] : X[0];
                py = (OPT >= 4 && (t & 2)) ? Y[1] : Y[0];
                pz = (OPT >= 2 && (t & 1)) ? Z[1] : Z[0];
            } else {
                t = u;
                px = X[(OPT >= 8) ? ((u >> 2) & 1) : 0];
                py = Y[(OPT >= 4) ? ((u >> 1) & 1) : 0];
                pz = Z[(OPT >= 2) ? (u & 1) : 0];
            }
            tt[v] = t;
            pos[v] = px + py + pz;
            const double ntau = __ldg(p.nhi + pos[v]);
            const double* pv = prev + (obase + t) * lmax;
            const double c1 = pv[nb1], c2 = pv[nb2], c3 = pv[nb3], c4 = pv[nb4];
            double cin;
            if (THK) {
                // all four >= 0.6 (zero-weight corners read the zero slot, kept at 1): c_i w_i = s_i prod(c), sum s_i = 1
                const double m12 = c1 * c2, m34 = c3 * c4;
                const double den = fma(m12, fma(s4, c3, s3 * c4), m34 * fma(s2, c1, s1 * c2));
                overflow |= !(den < 1e300);
                cin = (m12 * m34) * fast_rcp(den);
            } else {
                // interp_weighted (sweep_device.cuh) without its overflow branch, which is taken for the whole round below
                const double m1 = dmax(c1, 0.6), m2 = dmax(c2, 0.6), m3 = dmax(c3, 0.6), m4 = dmax(c4, 0.6);
                const double m12 = m1 * m2, m34 = m3 * m4;
                const double w1 = s1 * (m2 * m34), w2 = s2 * (m1 * m34), w3 = s3 * (m4 * m12), w4 = s4 * (m3 * m12);
                const double den = (w1 + w2) + (w3 + w4);
                overflow |= !(den < 1e300);
                cin = (fma(c1, w1, c2 * w2) + fma(c3, w3, c4 * w4)) * fast_rcp(den);
            }
            if (DIAG) {
                cin *= diag;
                if (flags & PC_SOURCE) cin = 0.0;  // the source cell "interpolates" the zero slot, which need not hold 0
            }
            tin[v] = cin;
            tout[v] = fma(ntau, path, cin);
            ok = ok && tau_is_thick(tout[v]);
        }
        if (__builtin_expect(overflow, 0)) {  // optical depths beyond 1e90: the reference's literal form (interp_weighted)
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const double* pv = prev + (obase + tt[u]) * lmax;
                tin[u] = interp_weighted<DIAG>(pv[nb1], pv[nb2], pv[nb3], pv[nb4], s1, s2, s3, s4, flags);
                if (DIAG && (flags & PC_SOURCE)) tin[u] = 0.0;
                tout[u] = fma(__ldg(p.nhi + pos[u]), path, tin[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < NB; u++) {
            cur[(obase + tt[u]) * lmax + slot] = tout[u];
            if (CLASS_B) cur[(obase + tt[u] + (int)zb) * lmax + slot] = tout[u];
        }

        // ---- phase 2: rates (raytracing.cu:315-329, rates.cu:16-41) ----------------------------------------------
        if (rated) {
            // deposit only from the image that owns the cell (clear bits on all its zero axes) and only while
            // coldensh_in <= MAX_COLDENSH (raytracing.cu:315)
#pragma unroll
            for (int u = 0; u < NB; u++)
                if (!(((zmask & (unsigned)(gbase + tt[u])) == 0) && (tin[u] <= p.tau_max))) pos[u] = ASORA_NO_DEPOSIT;
            // a thick cell wholly beyond the end of the table adds exactly T - T = 0 (both lookups clamp to the same entry)
#pragma unroll
            for (int u = 0; u < NB; u++)
                if (tin[u] >= p.tau_hi && fabs(tout[u] - tin[u]) > ASORA_TAU_PHOTO_LIMIT) pos[u] = ASORA_NO_DEPOSIT;
            double sk_n[NB];
#pragma unroll
            for (int u = 0; u < NB; u++) sk_n[u] = skn;
            rate_cells<NB, REP, TEX, HEAT, DET>(tin, tout, sk_n, pos, p, log2_tab);
        }
    }  // rounds
}

template <int BLOCK, int NOCT, int OPT, int BATCH, int REP, bool DIAG, bool TEX, bool HEAT, bool ZF, bool PF, bool DET, bool THK>
__device__ __forceinline__ void octant_level(const int4* __restrict__ plan, int ncells, int beg, int mid, int end, double md,
                                             double inv_m, double* __restrict__ cur, const double* __restrict__ prev, int lmax,
                                             const unsigned* __restrict__ wrap_tab, int hi, int part, double sk,
                                             const SweepParams& p, const double2* __restrict__ log2_tab, bool& ok)
{
    constexpr int G = NOCT / OPT;                 // warps that share an entry
    static_assert((BLOCK / 32) % G == 0 && BLOCK % 32 == 0, "the warps of a CTA must divide evenly over the image groups");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = (G > 1) ? (warp % G) : 0;
    const int obase = sub * OPT;                  // local octants [obase, obase + OPT)
    const int gbase = part * NOCT + obase;        // global octant bits of image t = 0
    const int side = 2 * hi + 1;
    // sign of the offsets on the axes this thread does not iterate (bit set = negative)
    const int fx = (gbase >> 2) & 1, fy = (gbase >> 1) & 1, fz = gbase & 1;
    int e = beg + (warp / G) * 32 + lane;
    int4 ra_next = make_int4(0, 0, 0, 0), rb_next = make_int4(0, 0, 0, 0);
    if (PF && e < end) {  // PF: the entry of the next iteration is fetched while this one is evaluated
        rb_next = load_plan(plan + (size_t)ncells + e);
        ra_next = load_plan(plan + e);
    }
    for (; e < end; e += BLOCK / G) {
        int4 ra, rb;
        if (PF) {
            ra = ra_next;
            rb = rb_next;
            const int en = e + BLOCK / G;
            if (en < end) {
                rb_next = load_plan(plan + (size_t)ncells + en);
                ra_next = load_plan(plan + en);
            }
        } else {
            rb = load_plan(plan + (size_t)ncells + e);
            ra = load_plan(plan + e);
        }
        const int di = rb.z & 0xff, dj = (rb.z >> 8) & 0xff, dk = (rb.z >> 16) & 0xff;
        // PC_ZFACE: second set of wrap tables, addressing the (k,i,j)-ordered copies of the opacity and rate grids
        const unsigned* w = ZF ? wrap_tab + (((unsigned)rb.z >> 28) & 1u) * (3 * side) : wrap_tab;
        unsigned X[2], Y[2], Z[2];
        if (OPT >= 8) {
            X[0] = w[hi + di];
            X[1] = w[hi - di];
        } else {
            X[0] = X[1] = w[fx ? hi - di : hi + di];
        }
        if (OPT >= 4) {
            Y[0] = w[side + hi + dj];
            Y[1] = w[side + hi - dj];
        } else {
            Y[0] = Y[1] = w[side + (fy ? hi - dj : hi + dj)];
        }
        if (OPT >= 2) {
            Z[0] = w[2 * side + hi + dk];
            Z[1] = w[2 * side + hi - dk];
        } else {
            Z[0] = Z[1] = w[2 * side + (fz ? hi - dk : hi + dk)];
        }
        if (OPT < 2 || e < mid)
            entry_images<OPT, BATCH, OPT, false, REP, DIAG, TEX, HEAT, DET, THK>(ra, rb, md, inv_m, e - beg, lmax, obase, gbase, X, Y, Z, cur,
                                                                             prev, sk, p, log2_tab, ok);
        else
            entry_images<(OPT >= 2 ? OPT / 2 : 1), BATCH, OPT, true, REP, DIAG, TEX, HEAT, DET, THK>(ra, rb, md, inv_m, e - beg, lmax, obase,
                                                                                                gbase, X, Y, Z, cur, prev, sk, p,
                                                                                                log2_tab, ok);
    }
}

// One CTA per (source, group of NOCT octants).  level_bounds_g: [nlevels + 1] level starts, then [3][nlevels] class
// boundaries for OPT = 8, 4, 2 (build_octant_plan).  dedup == 0: every entry is treated as class A (profiling).
template <int BLOCK, int MINB, int NOCT, int OPT, int BATCH, int REP, bool TEX, bool HEAT, bool ZF, bool PF, bool DET>
__global__ void __launch_bounds__(BLOCK, MINB)
sweep_octant_kernel(const int4* __restrict__ plan, int ncells, const int* __restrict__ level_bounds_g, int nlevels, int lmax,
                    int hi, int dedup, SweepParams p)
{
    extern __shared__ double2 sh_raw[];
    double2* log2_all = sh_raw;                                             // 256 entries x REP copies
    double* sh_cd = reinterpret_cast<double*>(sh_raw + 256 * REP);          // [2][NOCT][lmax]
    double* inv_level = sh_cd + (size_t)2 * NOCT * lmax;                    // [nlevels]
    const int side = 2 * hi + 1;
    unsigned* wrap_tab = reinterpret_cast<unsigned*>(inv_level + nlevels);  // [ZF ? 2 : 1][3][side]
    int* level_start = reinterpret_cast<int*>(wrap_tab + (size_t)(ZF ? 2 : 1) * 3 * side);  // [nlevels + 1]
    int* level_mid = level_start + nlevels + 1;                             // [nlevels]
    const int N = p.N;
    constexpr int PARTS = 8 / NOCT;
    // part-major: CTAs of different sources drift apart, the parts of one source would run in lock-step
    const int ngroups = gridDim.x / PARTS;
    const int part = blockIdx.x / ngroups;
    const int ns = p.src_begin + (blockIdx.x - part * ngroups);

    for (int t = threadIdx.x; t <= nlevels; t += BLOCK) level_start[t] = __ldg(level_bounds_g + t);
    constexpr int MIDROW = OPT >= 8 ? 0 : (OPT >= 4 ? 1 : 2);
    for (int t = threadIdx.x; t < nlevels; t += BLOCK) {
        level_mid[t] = (dedup && OPT >= 2) ? __ldg(level_bounds_g + (nlevels + 1) + MIDROW * nlevels + t)
                                           : __ldg(level_bounds_g + t + 1);
        inv_level[t] = t > 0 ? 1.0 / (double)t : 0.0;
    }
    for (int t = threadIdx.x; t < 256 * REP; t += BLOCK) log2_all[t] = __ldg(p.log2_tab + t / REP);
    const double2* log2_tab = log2_all + (REP > 1 ? (threadIdx.x & (REP - 1)) : 0);
    // the zero slot (last of every level buffer, sweep_plan.cu: resolve_zero_slot): zero-weight corners and the source cell
    // Its value only ever meets an exact zero weight.  1 (not 0) lets the thick-level form of the interpolation multiply it in
    // without a max(0.6, .); the deterministic instantiations keep the 0 all sweep variants share (bit-identical sums).
    if (threadIdx.x < 2 * NOCT) sh_cd[(size_t)(threadIdx.x + 1) * lmax - 1] = DET ? 0.0 : 1.0;
    const int i0 = p.src_pos[3 * ns + 0], j0 = p.src_pos[3 * ns + 1], k0 = p.src_pos[3 * ns + 2];
    const double sk = p.src_flux[ns] * p.kpref;
    for (int t = threadIdx.x; t < 3 * side; t += BLOCK) {
        const int axis = t / side, d = t - axis * side - hi;
        const int c0 = axis == 0 ? i0 : (axis == 1 ? j0 : k0);
        const unsigned NN = (unsigned)N * N, w = (unsigned)wrap(c0 + d, N);
        const unsigned stride = axis == 0 ? NN : (axis == 1 ? (unsigned)N : 1u);
        const unsigned stride_t = axis == 0 ? (unsigned)N : (axis == 1 ? 1u : NN);  // (k,i,j) order: i*N + j + k*N*N
        wrap_tab[t] = w * stride;
        if (ZF) wrap_tab[3 * side + t] = w * stride_t + (axis == 0 ? p.zface_offset : 0u);
    }
    __syncthreads();

    int beg = level_start[0], end = level_start[1];
    bool prev_thick = false;
    for (int m = 0; m < nlevels; m++) {
        const int next_end = level_start[min(m + 2, nlevels)];
        const int mid = level_mid[m];
        double* cur = sh_cd + (size_t)(m & 1) * NOCT * lmax;
        const double* prev = sh_cd + (size_t)((m & 1) ^ 1) * NOCT * lmax;
        const double md = (double)m, inv_m = inv_level[m];
        bool ok = true;
        if (m < 2)  // the source cell and its 26 neighbours: the only cells with diagonal factors (raytracing.cu:431-441)
            octant_level<BLOCK, NOCT, OPT, BATCH, REP, true, TEX, HEAT, ZF, PF, DET, false>(plan, ncells, beg, mid, end, md, inv_m, cur, prev,
                                                                                       lmax, wrap_tab, hi, part, sk, p, log2_tab, ok);
        else if (!DET && prev_thick)
            octant_level<BLOCK, NOCT, OPT, BATCH, REP, false, TEX, HEAT, ZF, PF, DET, !DET>(plan, ncells, beg, mid, end, md, inv_m, cur, prev,
                                                                                       lmax, wrap_tab, hi, part, sk, p, log2_tab, ok);
        else
            octant_level<BLOCK, NOCT, OPT, BATCH, REP, false, TEX, HEAT, ZF, PF, DET, false>(plan, ncells, beg, mid, end, md, inv_m, cur, prev,
                                                                                        lmax, wrap_tab, hi, part, sk, p, log2_tab, ok);
        // level barrier; it also tells every thread whether all optical depths of this level reached 0.6
        prev_thick = __syncthreads_and(ok) != 0;
        beg = end;
        end = next_end;
    }
}

template <int BLOCK, int MINB, int NOCT, int OPT, int BATCH, int REP, bool TEX, bool HEAT, bool ZF, bool PF, bool DET = false>
cudaError_t launch_t(const SweepPlan& plan, const SweepParams& p, int dedup, cudaStream_t stream)
{
    const size_t smem = sweep_octant_smem_bytes(plan, NOCT, REP, ZF);
    const int grid = p.src_count * (8 / NOCT);
    auto kernel = sweep_octant_kernel<BLOCK, MINB, NOCT, OPT, BATCH, REP, TEX, HEAT, ZF, PF, DET>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, BLOCK, smem, stream>>>(plan.d_cells, (int)plan.ncells, plan.d_level_start, plan.nlevels, plan.max_level_cells,
                                          (plan.side - 1) / 2, dedup, p);
    return cudaGetLastError();
}

// opts: bit 0 eight bank-staggered copies of the log2 table (large shapes only), bit 2 next plan entry prefetched,
// bit 3 no de-duplication of plane cells
template <int BLOCK, int MINB, int NOCT, int OPT, int BATCH, bool BIG>
cudaError_t launch_opts(const SweepPlan& plan, const SweepParams& p, int opts, cudaStream_t stream)
{
    const int dedup = (opts & 8) ? 0 : 1;
    if (p.det_lo) {  // deterministic accumulation: one option set per shape
        if (p.zface_offset || p.phi_heat) return cudaErrorInvalidValue;
        return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, false, false, false, true>(plan, p, dedup, stream);
    }
    if (p.phi_heat) {
        if (p.zface_offset) return cudaErrorInvalidValue;  // heating sweeps do not use the z-face copies
        return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, true, false, false>(plan, p, dedup, stream);
    }
    if (p.zface_offset) {
        if constexpr (BIG) {
            // copies of the log2 table: eight for one CTA per SM, four where two or more CTAs share an SM's shared memory
            constexpr int REPX = MINB >= 2 ? 4 : 8;
            if ((opts & 5) == 5) return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, REPX, true, false, true, true>(plan, p, dedup, stream);
            if (opts & 4) return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, false, true, true>(plan, p, dedup, stream);
            if (opts & 1) return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, REPX, true, false, true, false>(plan, p, dedup, stream);
            return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, false, true, false>(plan, p, dedup, stream);
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (opts & 4) return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, false, false, true>(plan, p, dedup, stream);
    return launch_t<BLOCK, MINB, NOCT, OPT, BATCH, 1, true, false, false, false>(plan, p, dedup, stream);
}

}  // namespace

// (noct, opt, batch, block, min CTAs per SM, big: z-face copies and log2-table copies available), split over
// translation units (the file is compiled once per ASORA_OCT_TU, pyc2ray_b200/_build.py) to keep the build parallel
#define ASORA_OCT_SHAPES_0(X) X(8, 4, 2, 512, 1, true) X(8, 2, 2, 768, 1, true) X(8, 4, 4, 384, 1, true)
#define ASORA_OCT_SHAPES_1(X) X(4, 4, 4, 256, 2, true) X(4, 2, 2, 384, 2, true) X(4, 4, 2, 256, 2, true)
#define ASORA_OCT_SHAPES_2(X) X(2, 2, 2, 192, 4, true) X(2, 2, 2, 128, 5, true) X(2, 2, 2, 256, 2, true)
#define ASORA_OCT_SHAPES_3(X) X(8, 2, 1, 896, 1, true) X(8, 8, 4, 384, 1, true) X(4, 2, 1, 448, 2, true)
#define ASORA_OCT_SHAPES_4(X) X(8, 2, 1, 128, 7, false) X(8, 4, 2, 128, 4, false) X(8, 4, 1, 128, 6, false)
#define ASORA_OCT_SHAPES_5(X) X(8, 2, 1, 256, 3, false) X(8, 4, 2, 256, 2, false) X(8, 8, 2, 64, 8, false) X(8, 4, 2, 64, 8, false)
#define ASORA_OCT_SHAPES_6(X) X(8, 2, 2, 128, 4, false) X(8, 4, 1, 64, 12, false)
#define ASORA_OCT_SHAPES_7(X) X(2, 2, 2, 160, 4, true) X(2, 2, 2, 224, 3, true) X(4, 2, 2, 320, 2, true)
#define ASORA_OCT_NTU 8
#if ASORA_OCT_TU == 0
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_0(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu0
#elif ASORA_OCT_TU == 1
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_1(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu1
#elif ASORA_OCT_TU == 2
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_2(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu2
#elif ASORA_OCT_TU == 3
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_3(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu3
#elif ASORA_OCT_TU == 4
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_4(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu4
#elif ASORA_OCT_TU == 5
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_5(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu5
#elif ASORA_OCT_TU == 6
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_6(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu6
#elif ASORA_OCT_TU == 7
#define ASORA_OCT_SHAPES(X) ASORA_OCT_SHAPES_7(X)
#define ASORA_OCT_LAUNCH launch_sweep_octant_tu7
#endif

cudaError_t ASORA_OCT_LAUNCH(const SweepPlan& plan, const SweepParams& p, int noct, int opt, int batch, int block, int opts,
                             cudaStream_t stream)
{
#define X(NO, OP, BA, BL, MB, BIG) \
    if (noct == NO && opt == OP && batch == BA && block == BL) return launch_opts<BL, MB, NO, OP, BA, BIG>(plan, p, opts, stream);
    ASORA_OCT_SHAPES(X)
#undef X
    return cudaErrorNotSupported;  // not one of this unit's shapes
}

#if ASORA_OCT_TU == 0
#define ASORA_OCT_DECL(K) cudaError_t launch_sweep_octant_tu##K(const SweepPlan&, const SweepParams&, int, int, int, int, int, cudaStream_t);
ASORA_OCT_DECL(1) ASORA_OCT_DECL(2) ASORA_OCT_DECL(3) ASORA_OCT_DECL(4) ASORA_OCT_DECL(5) ASORA_OCT_DECL(6) ASORA_OCT_DECL(7)
#undef ASORA_OCT_DECL

size_t sweep_octant_smem_bytes(const SweepPlan& plan, int noct, int rep, bool zface)
{
    return (size_t)256 * rep * sizeof(double2) + (size_t)2 * noct * plan.max_level_cells * sizeof(double) +
           (size_t)plan.nlevels * sizeof(double) + (size_t)(zface ? 2 : 1) * 3 * plan.side * sizeof(unsigned) +
           (size_t)(2 * plan.nlevels + 1) * sizeof(int);
}

// 0: not instantiated, 1: instantiated, 2: instantiated with the z-face / log2-copies options
int sweep_octant_shape_ok(int noct, int opt, int batch, int block)
{
#define X(NO, OP, BA, BL, MB, BIG) if (noct == NO && opt == OP && batch == BA && block == BL) return BIG ? 2 : 1;
    ASORA_OCT_SHAPES_0(X) ASORA_OCT_SHAPES_1(X) ASORA_OCT_SHAPES_2(X) ASORA_OCT_SHAPES_3(X)
    ASORA_OCT_SHAPES_4(X) ASORA_OCT_SHAPES_5(X) ASORA_OCT_SHAPES_6(X) ASORA_OCT_SHAPES_7(X)
#undef X
    return 0;
}

cudaError_t launch_sweep_octant(const SweepPlan& plan, const SweepParams& p, int noct, int opt, int batch, int block, int opts,
                                cudaStream_t stream, int* launches)
{
    if (p.src_count <= 0) return cudaSuccess;
    if (!plan.octant) return cudaErrorInvalidValue;
    typedef cudaError_t (*tu_fn)(const SweepPlan&, const SweepParams&, int, int, int, int, int, cudaStream_t);
    const tu_fn fns[ASORA_OCT_NTU] = {launch_sweep_octant_tu0, launch_sweep_octant_tu1, launch_sweep_octant_tu2, launch_sweep_octant_tu3,
                                      launch_sweep_octant_tu4, launch_sweep_octant_tu5, launch_sweep_octant_tu6, launch_sweep_octant_tu7};
    cudaError_t e = cudaErrorNotSupported;
    for (int k = 0; k < ASORA_OCT_NTU && e == cudaErrorNotSupported; k++) e = fns[k](plan, p, noct, opt, batch, block, opts, stream);
    if (e == cudaSuccess && launches) *launches += 1;
    return e;
}
#endif
