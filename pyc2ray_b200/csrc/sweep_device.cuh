// sweep_device.cuh -- per-cell device arithmetic shared by the sweep kernels (sweep_kernels.cu, sweep_octant.cu):
// fp64 helpers, the table-driven logarithm, the photo-ionisation table lookup (src/asora/rates.cu:16-83), the
// short-characteristics interpolation (src/asora/raytracing.cu:405-441) and the rate deposit of one cell
// (raytracing.cu:300-329).  Header-only; every translation unit gets its own copy of the constant-bank data.
#pragma once
#include "asora_common.cuh"

// ---------------------------------------------------------------------------------------------------
// fp64 helpers
// ---------------------------------------------------------------------------------------------------

// Polynomial coefficients live in the constant bank so that DFMA reads them as c[][] operands instead of
// materialising each 64-bit immediate with two UMOVs (6.7 % of the issued instructions in r01b).
// log2(1+r) = r * (K[0] + K[1] r + ... + K[5] r^5), K[k] = (-1)^k / ((k+1) ln 2)

static __constant__ double kLog2Poly[6] = {1.44269504088896341, -0.72134752044448170, 0.48089834696298783,
                                    -0.36067376022224085, 0.28853900817779268, -0.24044917348149391};

// (double)u for 0 <= u < 2^31 without the conversion unit: I2F / F2I run on the 16-lane XU pipe (8 cycles per warp,
// long latency; 9 of them per rated update kept that pipe 16 % busy), a DADD on the fp64 pipe takes 2.
#define ASORA_TWO52 4503599627370496.0
__device__ __forceinline__ double u2d(unsigned u) { return __hiloint2double(0x43300000, (int)u) - ASORA_TWO52; }

// max / min for ordinary (non-NaN) operands: one DSETP and two FSELs.  fmax()/fmin() cost 6-8
// instructions each on sm_100 because of their NaN rules.
// (Written in PTX: nvcc pattern-matches the C++ ternary back into max.f64, which sm_100 emulates with a
// DSETP.MAX + FSEL + SEL + LOP3 NaN-quieting sequence.)
__device__ __forceinline__ double dmax(double a, double b)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double dmin(double a, double b)
{
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

// log2(x), x normal and positive.  tab[j * REP] = {1/c_j, log2 c_j} for the 256 mantissa bins of [1,2).
// REP = 8: the caller passes tab + (lane & 7) and the table is stored eight times, entry j of copy r at
// 16-byte word 8 j + r.  The eight lanes an LDS.128 serves per cycle then always hit eight different bank
// groups; with a single copy the scattered mantissa bins of a warp cost ~11 shared-memory wavefronts per
// lookup instead of 4 (profiles/r01f: 0.67e9 excess wavefronts, the L1 data pipe being the busiest unit).
template <int REP>
__device__ __forceinline__ double fast_log2(int hi, int lo, const double2* __restrict__ tab)
{
    const double e = __hiloint2double(0x43300000, hi >> 20) - (ASORA_TWO52 + 1023.0);  // unbiased exponent, exact
    const double2 t = tab[((hi >> 12) & 0xff) * REP];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t.x, -1.0);  // |r| <= 2^-9
    // (the r^6 term, |K[5] r^6| <= 1.3e-17, is below the rounding of the result and left out)
    double p = fma(r, kLog2Poly[4], kLog2Poly[3]);
    p = fma(r, p, kLog2Poly[2]);
    p = fma(r, p, kLog2Poly[1]);
    p = fma(r, p, kLog2Poly[0]);
    return fma(r, p, e + t.y);
}


// ---------------------------------------------------------------------------------------------------
// per-cell arithmetic
// ---------------------------------------------------------------------------------------------------

// Out-of-range optical depths (tau < tau_lo: only the source cell and fully ionised paths; tau > tau_hi:
// beyond the table): a real call, so that the compiler keeps it out of the predicated fast path.
static __device__ __noinline__ double clamp_tau_slow(double tau, double lo, double hi) { return fmin(fmax(tau, lo), hi); }

// rates.cu:70-83.  The reference clamps tau from below at 1e-20 and the table index to [0, NumTau];
// here tau itself is clamped to [tau_lo, tau_hi], the optical depths at which the index reaches those
// bounds (tau_lo >= 1e-20), so index = lut_a + lut_b*log2(tau) needs no further clamping beyond the
// integer guard against the uploaded table length (reference bug N6: Python callers pass NumTau =
// table length, which lets i1 reach one element past the table; the pair table's last slope is 0).
// The range test is one unsigned compare on the high word (positive doubles order like integers): high
// words in [hi_min, hi_min + hi_span) are strictly inside (tau_lo, tau_hi).
struct TableIndex {
    int i0;
    double residual;
};

template <int REP>
__device__ __forceinline__ TableIndex table_index(int hi, int lo, const SweepParams& p, const double2* __restrict__ log2_tab)
{
    const double real_i = fma(p.lut_b, fast_log2<REP>(hi, lo, log2_tab), p.lut_a);
    // floor and fraction of 0 <= real_i < 2^31 (tau is clamped): adding 2^52 rounding down leaves floor(real_i) in
    // the low word
    const double shifted = __dadd_rd(real_i, ASORA_TWO52);
    TableIndex t;
    t.i0 = __double2loint(shifted);
    t.residual = real_i - (shifted - ASORA_TWO52);
    t.i0 = max(0, min(t.i0, p.ntab - 1));
    return t;
}

// The two lookups of a rated cell, T_thick(tau_in) and T_out(tau_out), with one shared (rarely taken) range
// branch so that their logarithms and table loads overlap.
// TEX: the two 16-byte gathers go through the texture pipe (tex1Dfetch) instead of LDG: a warp's 32 scattered
// table pairs cost ~20 wavefronts on the LSU data pipe, the busiest unit of the sweep.
// HEAT: the photo-heating tables (photorates.f90:118,124) share the index and the fraction of the ionisation
// tables, so heating costs two more gathers and no logarithm.  All four pair tables live in one allocation /
// one texture: thick, thin, heat thick, heat thin, ntab entries each.
template <int REP, bool TEX, bool HEAT>
__device__ __forceinline__ void photo_lookup2(bool thick, double tau_in, double tau_out, const SweepParams& p,
                                              const double2* __restrict__ log2_tab, double& t_in, double& t_out,
                                              double& h_in, double& h_out)
{
    int h1 = __double2hiint(tau_in), h2 = __double2hiint(tau_out);
    if (__builtin_expect(((unsigned)(h1 - p.hi_min) >= p.hi_span) | ((unsigned)(h2 - p.hi_min) >= p.hi_span), 0)) {
        // (selects, not a call: optically thick boxes reach the end of the table within tens of cells, so this path is
        // not rare on large-radius sweeps -- profiles/r02c)
        tau_in = dmin(dmax(tau_in, p.tau_lo), p.tau_hi);
        tau_out = dmin(dmax(tau_out, p.tau_lo), p.tau_hi);
        h1 = __double2hiint(tau_in);
        h2 = __double2hiint(tau_out);
    }
    const TableIndex a = table_index<REP>(h1, __double2loint(tau_in), p, log2_tab);
    const TableIndex b = table_index<REP>(h2, __double2loint(tau_out), p, log2_tab);
    const int ib = b.i0 + (thick ? 0 : p.ntab);  // thin table right behind the thick one
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
        const double2 ta = __ldg(p.thick + a.i0);
        const double2 tb = __ldg(p.thick + ib);
        t_in = fma(a.residual, ta.y, ta.x);
        t_out = fma(b.residual, tb.y, tb.x);
        if (HEAT) {
            const double2 va = __ldg(p.thick + a.i0 + 2 * p.ntab);
            const double2 vb = __ldg(p.thick + ib + 2 * p.ntab);
            h_in = fma(a.residual, va.y, va.x);
            h_out = fma(b.residual, vb.y, vb.x);
        }
    }
}

// raytracing.cu:405-441 with s1..s4 written in terms of the minor-axis fractions (sweep_plan.cu) and
// w_i = s_i / m_i, m_i = max(0.6, c_i sigma) (raytracing.cu:33) multiplied through by m1 m2 m3 m4.
// Corners whose bilinear weight is exactly zero never contribute (the reference multiplies whatever it
// reads by 0: raytracing.cu:416-428, SURVEY note N3).  MASK: the caller may have read stale memory
// for those corners, so force them to 0; the plan-driven sweep instead points them at a live slot.
// The kernels work in optical-depth units: every stored column is tau = sigma * N_HI (the pre-pass folds sigma and
// dr into the per-cell opacity), so c_i * sigma of the reference is the stored value itself.
// DIAG: only the 20 edge / corner neighbours of the source (level 1) carry the sqrt(2), sqrt(3) factors of
// raytracing.cu:431-441; the level loop of the shared-memory sweep compiles them out for levels >= 2.
// interp_weighted: the same with the four bilinear weights s1..s4 given (the mirror-image sweep computes them once per
// plan entry for all images).
template <bool DIAG>
__device__ __forceinline__ double interp_weighted(double c1, double c2, double c3, double c4, double s1, double s2,
                                                  double s3, double s4, unsigned flags)
{
    const double m1 = dmax(c1, 0.6), m2 = dmax(c2, 0.6);
    const double m3 = dmax(c3, 0.6), m4 = dmax(c4, 0.6);
    const double m12 = m1 * m2, m34 = m3 * m4;
    double w1 = s1 * (m2 * m34), w2 = s2 * (m1 * m34);
    double w3 = s3 * (m4 * m12), w4 = s4 * (m3 * m12);
    double den = (w1 + w2) + (w3 + w4);
    double cdensi;
    if (den < 1e300) {
        cdensi = (fma(c1, w1, c2 * w2) + fma(c3, w3, c4 * w4)) * fast_rcp(den);
    } else {  // the products overflowed (optical depths beyond 1e90): the reference's literal form
        w1 = s1 / m1, w2 = s2 / m2, w3 = s3 / m3, w4 = s4 / m4;
        cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
    }
    if (DIAG) {
        if ((flags & (PC_DIAG2 | PC_DIAG3)) != 0) cdensi *= (flags & PC_DIAG3) ? ASORA_SQRT3 : ASORA_SQRT2;
    }
    return cdensi;
}

// Thick levels (mirror-image sweep): when every optical depth of the previous level is >= 0.6 -- a level-uniform fact carried
// through the level barrier --, max(0.6, tau) of raytracing.cu:33 is tau itself, c_i w_i = s_i prod(c) and sum s_i = 1, so the
// weighted mean collapses to prod(c) / sum_i s_i prod_{j != i} c_j: no selects, no numerator (sweep_octant.cu: entry_images).
// 0.6 = 0x3FE3333333333333: a high word above 0x3FE33333 means tau > 0.6 (positive doubles order like integers)
__device__ __forceinline__ bool tau_is_thick(double tau) { return __double2hiint(tau) > 0x3FE33333; }

template <bool MASK, bool DIAG>
__device__ __forceinline__ double interp_coldens(double c1, double c2, double c3, double c4, double wA,
                                                 double wB, unsigned flags)
{
    const double uA = 1.0 - wA, uB = 1.0 - wB;
    const double s1 = wA * wB, s2 = wB * uA, s3 = wA * uB, s4 = uA * uB;
    if (MASK) {
        c1 = (s1 != 0.0) ? c1 : 0.0;
        c2 = (s2 != 0.0) ? c2 : 0.0;
        c3 = (s3 != 0.0) ? c3 : 0.0;
        c4 = (s4 != 0.0) ? c4 : 0.0;
    }
    return interp_weighted<DIAG>(c1, c2, c3, c4, s1, s2, s3, s4, flags);
}

__device__ __forceinline__ int wrap(int i, int N)
{
    i += (i < 0) ? N : 0;
    i -= (i >= N) ? N : 0;
    return i;
}

// Everything after the incoming optical depth is known: raytracing.cu:300-329 + rates.cu:16-41, with
//   ntau     = nHI * sigma * dr            opacity of the cell per unit path in cell units (pre-pass)
//   tau_out  = tau_in + ntau * path        sigma * (coldensh_in + nHI * path * dr)          (raytracing.cu:311)
//   phi      = strength / Vfact * absorbed / nHI = strength * inv_np * kpref * absorbed / ntau,
//              kpref = sigma * dr / (4 pi dr^3)                                   (rates.cu:24, raytracing.cu:324)
// The division by ntau is the same for every source that reaches the cell, so the sweep accumulates
// strength * kpref * inv_np * absorbed and one pass over the grid divides afterwards (finish_phi_kernel):
// sk = strength * kpref.  Returns the outgoing optical depth.
// One rate contribution into the accumulation grid.
//   default        RED.E.ADD.F64, resolved at L2: one fire-and-forget fp64 reduction per rated (source, cell) pair; the sum
//                  depends on the order in which the L2 sees them at the 1e-16 level (like the reference's atomicAdd,
//                  raytracing.cu:328)
//   deterministic  (DET instantiations, asora_set_deterministic) the contribution is split exactly into two integers, value * 2^s = hi * 2^46 + lo with |lo| < 2^46, and added with two 64-bit integer REDs; integer addition is associative, so the sums --
//                  and phi_ion after the division pass -- are bit-identical from run to run and for every launch shape.
//                  s is chosen on the host so that the largest possible contribution stays below 2^88 (asora_api.cu):
//                  a contribution 1e-11 times that still keeps 53 significant bits, and 2^17 contributions fit a cell.
#define ASORA_DET_LOW_BITS 46
template <bool DET>
__device__ __forceinline__ void deposit_rate(double* __restrict__ grid, long long* __restrict__ lo_grid, double det_scale,
                                             size_t pos, double value)
{
    if (!DET) {
        atomicAdd(grid + pos, value);
    } else {
        const double v = value * det_scale;                                        // exact: a power of two
        const long long hi = __double2ll_rz(v * (1.0 / (double)(1ll << ASORA_DET_LOW_BITS)));
        const double rest = v - (double)hi * (double)(1ll << ASORA_DET_LOW_BITS);  // exact, |rest| < 2^46
        atomicAdd(reinterpret_cast<unsigned long long*>(grid) + pos, (unsigned long long)hi);
        atomicAdd(reinterpret_cast<unsigned long long*>(lo_grid) + pos, (unsigned long long)__double2ll_rz(rest));
    }
}

// skn = strength * kpref * inv_np.
// GREY: the reference's -D GREY_NOTABLES build (raytracing.cu:317-318): analytic grey-opacity rates instead of the tables.
template <int REP, bool TEX, bool HEAT, bool DET = false, bool GREY = false>
__device__ __forceinline__ double finish_cell_pre(double tau_in, double path_cells, double skn, unsigned flags,
                                                  double ntau_p, size_t pos, const SweepParams& p,
                                                  const double2* __restrict__ log2_tab)
{
    const double tau_out = fma(ntau_p, path_cells, tau_in);
    if ((flags & PC_RATED) && tau_in <= p.tau_max) {  // coldensh_in <= MAX_COLDENSH (raytracing.cu:315)
        const double dtau = tau_out - tau_in;
        const bool thick = fabs(dtau) > ASORA_TAU_PHOTO_LIMIT;
        if (GREY) {
            // photoion_rates_test_gpu (rates.cu:48-64): strength * S_STAR_REF / Vfact * (exp(-tau_in) - exp(-tau_out)), thin
            // cells (tau_out - tau_in) * exp(-tau_in); the division by nHI is deferred like that of the table rates
            const double e_in = exp(-tau_in);
            const double absorbed = thick ? (e_in - exp(-tau_out)) : dtau * e_in;
            deposit_rate<false>(p.phi_ion, nullptr, 0.0, pos, (skn * ASORA_S_STAR_REF) * absorbed);
            return tau_out;
        }
        // Beyond the end of the table both lookups of a thick cell are clamped to the same entry (rates.cu:78-79), so the
        // reference adds exactly T - T = 0 there: nothing to look up or deposit.  In an optically thick box that is most
        // of a large-radius sweep (tau reaches 10^4 within 44 cells of a source in the reference's benchmark field).
        if (thick && tau_in >= p.tau_hi) return tau_out;
        double t_in, t_out, h_in = 0.0, h_out = 0.0;
        photo_lookup2<REP, TEX, HEAT>(thick, tau_in, tau_out, p, log2_tab, t_in, t_out, h_in, h_out);
        // rates.cu:28-38: thick cells absorb T(tau_in) - T(tau_out), thin cells dtau * T_thin(tau_out)
        const double absorbed = thick ? (t_in - t_out) : dtau * t_out;
        deposit_rate<DET>(p.phi_ion, p.det_lo, p.det_scale, pos, skn * absorbed);
        if (HEAT) {  // photorates.f90:118,124 + raytracing.f90:530,537, same prefactor and the same deferred / nHI
            // Thin cells: the heating table is read at tau_out, the convention of the ASORA ionisation rate above
            // (rates.cu:36), not at tau_in as in photorates.f90:124; the two differ by O(dtau) <= 1e-7 relative.
            const double heated = thick ? (h_in - h_out) : dtau * h_out;
            deposit_rate<DET>(p.phi_heat, p.det_lo_heat, p.det_scale_heat, pos, skn * heated);
        }
    }
    return tau_out;
}

template <int REP, bool TEX, bool HEAT, bool DET = false, bool GREY = false>
__device__ __forceinline__ double finish_cell(double tau_in, double path_cells, double inv_np, unsigned flags,
                                              double ntau_p, double sk, size_t pos, const SweepParams& p,
                                              const double2* __restrict__ log2_tab)
{
    return finish_cell_pre<REP, TEX, HEAT, DET, GREY>(tau_in, path_cells, sk * inv_np, flags, ntau_p, pos, p, log2_tab);
}
