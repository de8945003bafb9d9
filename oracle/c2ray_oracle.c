/*
 * c2ray_oracle.c -- CPU restatement of pyc2ray's ray-tracing + chemistry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in pyc2ray_b200/ may import, link or call this file.
 * Allowed users: tests/, __graft_entry__.smoke(), and the cpu_baseline / --impl reference
 * legs of bench.py.  The product path is the CUDA library (pyc2ray_b200/csrc) and fails
 * loudly when that library is missing; it never falls back to this code.
 *
 * What is restated (paths relative to the reference checkout):
 *   src/c2ray/raytracing.f90   do_all_sources :52-119, do_source :127-249, evolve2D :258-340,
 *                              evolve0D :347-567, cinterp :576-815
 *   src/c2ray/photorates.f90   photoion_rates :62-149 (+ nested photo_lookuptable :130-147)
 *   src/c2ray/chemistry.f90    global_pass :13-48, evolve0D_global :53-110,
 *                              do_chemistry :117-204, doric :221-316
 *   src/asora/raytracing.cu    do_all_sources_gpu :79-148, evolve0D_gpu :155-339,
 *                              cinterp_gpu :345-535, linthrd2cart :39-59
 *   src/asora/rates.cu         photoion_rates_gpu :16-41, photo_lookuptable :70-83
 *
 * The Fortran sources cannot be compiled in this image (no Fortran compiler), so this C file
 * is the CPU oracle ("port").  Two flavours of the same per-cell arithmetic are offered, chosen
 * by `flavour`:
 *   ORACLE_FORTRAN (0)  constants and traversal of src/c2ray (single-precision sqrt(2), sqrt(3),
 *                       1.0e-7, 2e30 literals widened to double; T_thin(tau_in); cube traversal
 *                       plane by plane with the optional sub-box / photon-loss exit; 1-indexed
 *                       source positions; Fortran-ordered grids)
 *   ORACLE_ASORA   (1)  constants and traversal of src/asora (double literals 1.73205080757 /
 *                       1.41421356237; T_thin(tau_out); octahedral q-shells clipped to the
 *                       +-N/2 cube; 0-indexed sources; C-ordered grids, flat index i*N*N+j*N+k)
 *
 * Compile with -ffp-contract=off so every a*b+c rounds twice, exactly as gfortran does on
 * baseline x86-64.  The one place where nvcc's default FMA contraction changes a *decision*
 * in the ASORA flavour (the dist2 <= R^2 sphere test) can be switched to the fused form with
 * ORACLE_OPT_FMA_DIST2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_FORTRAN 0
#define ORACLE_ASORA 1

/* option bits */
#define ORACLE_OPT_NORMFLUX_BUG 1 /* raytracing.f90:500,503 uses normflux(NumSrc) for every source */
#define ORACLE_OPT_USE_SUBBOX 2   /* -DUSE_SUBBOX (src/c2ray/Makefile:3) */
#define ORACLE_OPT_FMA_DIST2 4    /* nvcc contracts xs*xs+ys*ys+zs*zs (raytracing.cu:305) into FMAs */
#define ORACLE_OPT_GREY_NOTABLES 8 /* -DGREY_NOTABLES: analytic grey-opacity rates (rates.cu:48-64, photorates.f90:13-57) */

typedef struct {
    double sqrt3, sqrt2;    /* raytracing.f90:608-609 vs raytracing.cu:435,439 */
    double tau_photo_limit; /* photorates.f90:69 vs rates.cu:7 */
    double max_coldensh;    /* raytracing.f90:368 vs raytracing.cu:15 */
    int thin_uses_tau_out;  /* rates.cu:37 (tau_out) vs photorates.f90:121 (tau_in) */
    int fma_dist2;
    int grey_notables;      /* raytracing.cu:317-318, raytracing.f90:499-501 */
} consts_t;

static consts_t make_consts(int flavour, int opts)
{
    consts_t c;
    if (flavour == ORACLE_FORTRAN) {
        c.sqrt3 = (double)sqrtf(3.0f);
        c.sqrt2 = (double)sqrtf(2.0f);
        c.tau_photo_limit = (double)1.0e-7f;
        c.max_coldensh = (double)2e30f;
        c.thin_uses_tau_out = 0;
        c.fma_dist2 = 0;
    } else {
        c.sqrt3 = 1.73205080757;
        c.sqrt2 = 1.41421356237;
        c.tau_photo_limit = 1.0e-7;
        c.max_coldensh = 2e30;
        c.thin_uses_tau_out = 1;
        c.fma_dist2 = (opts & ORACLE_OPT_FMA_DIST2) ? 1 : 0;
    }
    c.grey_notables = (opts & ORACLE_OPT_GREY_NOTABLES) ? 1 : 0;
    return c;
}

/* Fortran modulo / raytracing.cu:23-24 */
static inline int modulo(int a, int b) { return (a % b + b) % b; }
/* Fortran sign(1,x) / raytracing.cu:27 (sign of zero is +1) */
static inline int sign1(int x) { return x >= 0 ? 1 : -1; }

/* grid indexing: `fo` = Fortran order (i fastest), else C order (k fastest); i,j,k 0-based */
static inline size_t gidx(int i, int j, int k, int N, int fo)
{
    return fo ? ((size_t)k * N + j) * N + i : ((size_t)i * N + j) * N + k;
}

/* photorates.f90:130-147 == rates.cu:70-83.  `ntab` is the number of valid table entries
 * (reference bug N6: the caller may pass NumTau == ntab, which makes i1 == ntab reachable;
 * the reference then reads one element past the table, we clamp to the last valid entry). */
static double photo_lookuptable(const double *table, double tau, double minlogtau, double dlogtau,
                                int NumTau, int ntab)
{
    double logtau = log10(fmax(1.0e-20, tau));
    double real_i = fmin((double)(float)NumTau, fmax(0.0, 1.0 + (logtau - minlogtau) / dlogtau));
    int i0 = (int)real_i;
    int i1 = (NumTau < i0 + 1) ? NumTau : i0 + 1;
    double residual = real_i - (double)i0;
    if (i0 > ntab - 1) i0 = ntab - 1;
    if (i1 > ntab - 1) i1 = ntab - 1;
    return table[i0] + residual * (table[i1] - table[i0]);
}

/* photorates.f90:62-127 == rates.cu:16-41 (photo-ionisation part; heating: photoheat_rate below) */
static double photoion_rates(const consts_t *c, double normflux, double coldens_in, double coldens_out,
                             double Vfact, double sig, const double *thin, const double *thick,
                             double minlogtau, double dlogtau, int NumTau, int ntab, double *phi_out)
{
    double tau_in = coldens_in * sig;
    double tau_out = coldens_out * sig;
    double prefact = normflux / Vfact;
    double phi_photo_in, cell;
    if (c->grey_notables) {
        /* photoion_rates_test_gpu (rates.cu:48-64) == photoion_rates_test (photorates.f90:13-57): no tables, the
         * source strength in units of S_STAR_REF = 1e48 (rates.cu:8); the thin branch uses exp(-tau_in) in both */
        prefact = normflux * 1e48 / Vfact;
        phi_photo_in = prefact * exp(-tau_in);
        if (fabs(tau_out - tau_in) > c->tau_photo_limit) {
            double phi_photo_out = prefact * exp(-tau_out);
            cell = phi_photo_in - phi_photo_out;
            if (phi_out) *phi_out = phi_photo_out;
        } else {
            cell = prefact * (tau_out - tau_in) * exp(-tau_in);
            if (phi_out) *phi_out = phi_photo_in - cell;
        }
        return cell;
    }
    phi_photo_in = prefact * photo_lookuptable(thick, tau_in, minlogtau, dlogtau, NumTau, ntab);
    if (fabs(tau_out - tau_in) > c->tau_photo_limit) {
        double phi_photo_out = prefact * photo_lookuptable(thick, tau_out, minlogtau, dlogtau, NumTau, ntab);
        cell = phi_photo_in - phi_photo_out;
        if (phi_out) *phi_out = phi_photo_out;
    } else {
        double targ = c->thin_uses_tau_out ? tau_out : tau_in;
        cell = prefact * (tau_out - tau_in) * photo_lookuptable(thin, targ, minlogtau, dlogtau, NumTau, ntab);
        if (phi_out) *phi_out = phi_photo_in - cell;
    }
    return cell;
}

/* photorates.f90:118,124: photo-heating rate of a cell from the heating tables, with the same thick/thin
 * split and table argument as the ionisation rate above (the ASORA flavour, which has no heating in the
 * reference -- TODO at c2ray_base.py:424-426 --, keeps its own thin-cell argument tau_out, rates.cu:37). */
static double photoheat_rate(const consts_t *c, double normflux, double coldens_in, double coldens_out, double Vfact,
                             double sig, const double *heat_thin, const double *heat_thick, double minlogtau,
                             double dlogtau, int NumTau, int ntab)
{
    double tau_in = coldens_in * sig;
    double tau_out = coldens_out * sig;
    double prefact = normflux / Vfact;
    if (fabs(tau_out - tau_in) > c->tau_photo_limit)
        return prefact * (photo_lookuptable(heat_thick, tau_in, minlogtau, dlogtau, NumTau, ntab) -
                          photo_lookuptable(heat_thick, tau_out, minlogtau, dlogtau, NumTau, ntab));
    {
        double targ = c->thin_uses_tau_out ? tau_out : tau_in;
        return prefact * (tau_out - tau_in) * photo_lookuptable(heat_thin, targ, minlogtau, dlogtau, NumTau, ntab);
    }
}

/* raytracing.f90:807-813 == raytracing.cu:33 */
static inline double weightf(double cd, double sig) { return 1.0 / fmax(0.6, cd * sig); }

/* raytracing.f90:576-815 == raytracing.cu:345-535.  (i,j,k),(i0,j0,k0) are un-wrapped mesh
 * coordinates in whatever indexing the caller uses; `base` is that indexing's origin (1 for the
 * Fortran flavour, 0 for ASORA) so the periodic wrap matches modulo(x-1,m)+1 / modulo_gpu. */
static void cinterp(const consts_t *c, int i, int j, int k, int i0, int j0, int k0, double *cdensi,
                    double *path, const double *coldensh_out, double sig, int N, int fo, int base)
{
    int idel = i - i0, jdel = j - j0, kdel = k - k0;
    int idela = abs(idel), jdela = abs(jdel), kdela = abs(kdel);
    int sgni = sign1(idel), sgnj = sign1(jdel), sgnk = sign1(kdel);
    int im = i - sgni, jm = j - sgnj, km = k - sgnk;
    double di = (double)idel, dj = (double)jdel, dk = (double)kdel;
    double alam, xc, yc, zc, dx, dy, dz, s1, s2, s3, s4, c1, c2, c3, c4, w1, w2, w3, w4;
    int ip = modulo(i - base, N), imp = modulo(im - base, N);
    int jp = modulo(j - base, N), jmp = modulo(jm - base, N);
    int kp = modulo(k - base, N), kmp = modulo(km - base, N);

    if (kdela >= jdela && kdela >= idela) {
        alam = ((double)(km - k0) + sgnk * 0.5) / dk;
        xc = alam * di + (double)i0;
        yc = alam * dj + (double)j0;
        dx = 2.0 * fabs(xc - ((double)im + 0.5 * sgni));
        dy = 2.0 * fabs(yc - ((double)jm + 0.5 * sgnj));
        s1 = (1. - dx) * (1. - dy);
        s2 = (1. - dy) * dx;
        s3 = (1. - dx) * dy;
        s4 = dx * dy;
        c1 = coldensh_out[gidx(imp, jmp, kmp, N, fo)];
        c2 = coldensh_out[gidx(ip, jmp, kmp, N, fo)];
        c3 = coldensh_out[gidx(imp, jp, kmp, N, fo)];
        c4 = coldensh_out[gidx(ip, jp, kmp, N, fo)];
        w1 = s1 * weightf(c1, sig);
        w2 = s2 * weightf(c2, sig);
        w3 = s3 * weightf(c3, sig);
        w4 = s4 * weightf(c4, sig);
        *cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
        if (kdela == 1 && (idela == 1 || jdela == 1)) {
            if (idela == 1 && jdela == 1)
                *cdensi = c->sqrt3 * *cdensi;
            else
                *cdensi = c->sqrt2 * *cdensi;
        }
        *path = sqrt((di * di + dj * dj) / (dk * dk) + 1.0);
    } else if (jdela >= idela && jdela >= kdela) {
        alam = ((double)(jm - j0) + sgnj * 0.5) / dj;
        zc = alam * dk + (double)k0;
        xc = alam * di + (double)i0;
        dz = 2.0 * fabs(zc - ((double)km + 0.5 * sgnk));
        dx = 2.0 * fabs(xc - ((double)im + 0.5 * sgni));
        s1 = (1. - dx) * (1. - dz);
        s2 = (1. - dz) * dx;
        s3 = (1. - dx) * dz;
        s4 = dx * dz;
        c1 = coldensh_out[gidx(imp, jmp, kmp, N, fo)];
        c2 = coldensh_out[gidx(ip, jmp, kmp, N, fo)];
        c3 = coldensh_out[gidx(imp, jmp, kp, N, fo)];
        c4 = coldensh_out[gidx(ip, jmp, kp, N, fo)];
        w1 = s1 * weightf(c1, sig);
        w2 = s2 * weightf(c2, sig);
        w3 = s3 * weightf(c3, sig);
        w4 = s4 * weightf(c4, sig);
        *cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
        if (jdela == 1 && (idela == 1 || kdela == 1)) {
            if (idela == 1 && kdela == 1)
                *cdensi = c->sqrt3 * *cdensi;
            else
                *cdensi = c->sqrt2 * *cdensi;
        }
        *path = sqrt((di * di + dk * dk) / (dj * dj) + 1.0);
    } else {
        alam = ((double)(im - i0) + sgni * 0.5) / di;
        zc = alam * dk + (double)k0;
        yc = alam * dj + (double)j0;
        dz = 2.0 * fabs(zc - ((double)km + 0.5 * sgnk));
        dy = 2.0 * fabs(yc - ((double)jm + 0.5 * sgnj));
        s1 = (1. - dz) * (1. - dy);
        s2 = (1. - dz) * dy;
        s3 = (1. - dy) * dz;
        s4 = dy * dz;
        c1 = coldensh_out[gidx(imp, jmp, kmp, N, fo)];
        c2 = coldensh_out[gidx(imp, jp, kmp, N, fo)];
        c3 = coldensh_out[gidx(imp, jmp, kp, N, fo)];
        c4 = coldensh_out[gidx(imp, jp, kp, N, fo)];
        w1 = s1 * weightf(c1, sig);
        w2 = s2 * weightf(c2, sig);
        w3 = s3 * weightf(c3, sig);
        w4 = s4 * weightf(c4, sig);
        *cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
        if (idela == 1 && (jdela == 1 || kdela == 1)) {
            if (jdela == 1 && kdela == 1)
                *cdensi = c->sqrt3 * *cdensi;
            else
                *cdensi = c->sqrt2 * *cdensi;
        }
        *path = sqrt(1.0 + (dj * dj + dk * dk) / (di * di));
    }
}

/* One (source, cell) update: raytracing.f90:347-567 == raytracing.cu:241-330.
 * Returns 1 if the cell was processed.  `guard` reproduces raytracing.f90:426. */
typedef struct {
    const consts_t *c;
    int N, fo, base, guard;
    double sig, dr, Rmax;
    const double *ndens, *xh_av, *thin, *thick;
    const double *heat_thin, *heat_thick; /* NULL: no heating */
    size_t heat_off;                      /* phi_heat of a cell lives heat_off doubles after its phi_ion */
    double minlogtau, dlogtau;
    int NumTau, ntab;
} cellctx_t;

static int evolve0D(const cellctx_t *x, int i, int j, int k, int i0, int j0, int k0, double flux,
                    double *coldensh_out, double *phi_ion, double *phi_out_ret)
{
    const int N = x->N;
    int p0 = modulo(i - x->base, N), p1 = modulo(j - x->base, N), p2 = modulo(k - x->base, N);
    size_t pos = gidx(p0, p1, p2, N, x->fo);
    double xh_av_p = x->xh_av[pos];
    double nHI_p = x->ndens[pos] * (1.0 - xh_av_p);
    double coldensh_in, path, vol_ph, dist2 = 0.0;
    int stop = 0;
    const double dr = x->dr;

    if (x->guard && coldensh_out[pos] != 0.0) return 0;

    if (i == i0 && j == j0 && k == k0) {
        coldensh_in = 0.0;
        path = 0.5 * dr;
        vol_ph = dr * dr * dr;
    } else {
        double xs, ys, zs;
        cinterp(x->c, i, j, k, i0, j0, k0, &coldensh_in, &path, coldensh_out, x->sig, N, x->fo, x->base);
        path = path * dr;
        xs = dr * (double)(i - i0);
        ys = dr * (double)(j - j0);
        zs = dr * (double)(k - k0);
        if (x->c->fma_dist2)
            dist2 = fma(zs, zs, fma(ys, ys, xs * xs));
        else
            dist2 = xs * xs + ys * ys + zs * zs;
        vol_ph = dist2 * path * (4.0 * 3.14159265358979323846264338);
        if (dist2 / (dr * dr) > x->Rmax * x->Rmax) stop = 1;
        if (coldensh_in > x->c->max_coldensh) stop = 1;
    }
    {
        double cdho = coldensh_in + nHI_p * path;
        double phi = 0.0, pout = 0.0;
        coldensh_out[pos] = cdho;
        if (!stop) {
            phi = photoion_rates(x->c, flux, coldensh_in, cdho, vol_ph, x->sig, x->thin, x->thick,
                                 x->minlogtau, x->dlogtau, x->NumTau, x->ntab, &pout);
            phi = phi / nHI_p;
            /* raytracing.cu:315-329 only touches phi_ion for rated cells; the Fortran adds 0/nHI
             * (raytracing.f90:519,531,536), which is the same unless nHI_p == 0 */
            phi_ion[pos] += phi;
            if (x->heat_thick) { /* raytracing.f90:530,537 */
                double heat = photoheat_rate(x->c, flux, coldensh_in, cdho, vol_ph, x->sig, x->heat_thin,
                                             x->heat_thick, x->minlogtau, x->dlogtau, x->NumTau, x->ntab);
                phi_ion[pos + x->heat_off] += heat / nHI_p;
            }
        } else if (x->c->thin_uses_tau_out == 0) {
            phi_ion[pos] += 0.0 / nHI_p;
        }
        if (phi_out_ret) *phi_out_ret = pout;
    }
    return 1;
}

/* raytracing.cu:39-59 */
static void linthrd2cart(int s, int q, int *i, int *j)
{
    if (s == 0) {
        *i = q;
        *j = 0;
    } else {
        int b = (s - 1) / (2 * q);
        int a = (s - 1) % (2 * q);
        if (a + 2 * b > 2 * q) {
            a = a + 1;
            b = b - 1 - q;
        }
        *i = a + b - q;
        *j = b;
    }
}

/* ASORA traversal for one source: raytracing.cu:188-337.  Returns the number of visited cells. */
static long asora_source(const cellctx_t *x, int q_max, int i0, int j0, int k0, double flux,
                         double *slab, double *phi_ion)
{
    const int N = x->N;
    const int last_r = N / 2 - 1 + modulo(N, 2), last_l = -N / 2; /* raytracing.cu:122-123 */
    long visited = 0;
    for (int q = 0; q <= q_max; q++) {
        int s_end = (q == 0) ? 1 : 4 * q * q + 2;
        int s_end_top = 2 * q * (q + 1) + 1;
        for (int s = 0; s < s_end; s++) {
            int i, j, k, sgn;
            if (s < s_end_top) {
                sgn = 1;
                linthrd2cart(s, q, &i, &j);
            } else {
                sgn = -1;
                linthrd2cart(s - s_end_top, q - 1, &i, &j);
            }
            k = sgn * q - sgn * (abs(i) + abs(j));
            if (i >= last_l && i <= last_r && j >= last_l && j <= last_r && k >= last_l && k <= last_r)
                visited += evolve0D(x, i + i0, j + j0, k + k0, i0, j0, k0, flux, slab, phi_ion, NULL);
        }
    }
    return visited;
}

/* Fortran traversal for one source: raytracing.f90:127-249 (do_source) + :258-340 (evolve2D).
 * srcpos 1-indexed.  Returns visited cells; accumulates sum_nbox / photon_loss. */
static long fortran_source(const cellctx_t *x, int opts, const int *sp, double flux_rate, double flux_ns,
                           int max_subbox, int subboxsize, float loss_fraction, double *slab,
                           double *phi_ion, int *sum_nbox, double *photon_loss)
{
    const int N = x->N;
    const double S_star = 1e48; /* photorates.f90:7 */
    int lastpos_r[3], lastpos_l[3], last_r[3], last_l[3];
    long visited = 0;
    for (int d = 0; d < 3; d++) {
        int a = N / 2 - 1 + N % 2, b = N / 2;
        lastpos_r[d] = sp[d] + (max_subbox < a ? max_subbox : a);
        lastpos_l[d] = sp[d] - (max_subbox < b ? max_subbox : b);
    }
    memset(slab, 0, sizeof(double) * (size_t)N * N * N); /* raytracing.f90:181 */

#define PLANE(kk)                                                                                    \
    do {                                                                                             \
        for (int pass = 0; pass < 2; pass++) {                                                       \
            int jb = pass == 0 ? sp[1] : sp[1] - 1, je = pass == 0 ? last_r[1] : last_l[1];           \
            int jst = pass == 0 ? 1 : -1;                                                            \
            for (int j = jb; pass == 0 ? j <= je : j >= je; j += jst) {                              \
                for (int i = sp[0]; i <= last_r[0]; i++) CELL(i, j, kk);                             \
                for (int i = sp[0] - 1; i >= last_l[0]; i--) CELL(i, j, kk);                         \
            }                                                                                        \
        }                                                                                            \
    } while (0)
#define CELL(ii, jj, kk)                                                                             \
    do {                                                                                             \
        double pout = 0.0;                                                                           \
        int did = evolve0D(x, ii, jj, kk, sp[0], sp[1], sp[2], flux_rate, slab, phi_ion, &pout);     \
        visited += did;                                                                              \
        if (did && (opts & ORACLE_OPT_USE_SUBBOX) &&                                                 \
            ((ii) == last_l[0] || (jj) == last_l[1] || (kk) == last_l[2] || (ii) == last_r[0] ||     \
             (jj) == last_r[1] || (kk) == last_r[2]))                                                \
            photon_loss_src += pout * (x->dr * x->dr * x->dr); /* raytracing.f90:541-543 */          \
    } while (0)

    if (opts & ORACLE_OPT_USE_SUBBOX) {
        int nbox = 0;
        double photon_loss_src = flux_ns * S_star;
        for (int d = 0; d < 3; d++) last_r[d] = last_l[d] = sp[d];
        /* raytracing.f90:193-221; loss_fraction is default real (single precision) */
        while (photon_loss_src > (double)loss_fraction * flux_ns * S_star && last_r[2] < lastpos_r[2] &&
               last_l[2] > lastpos_l[2]) {
            photon_loss_src = 0.0;
            nbox++;
            for (int d = 0; d < 3; d++) {
                int r = sp[d] + subboxsize * nbox, l = sp[d] - subboxsize * nbox;
                last_r[d] = r < lastpos_r[d] ? r : lastpos_r[d];
                last_l[d] = l > lastpos_l[d] ? l : lastpos_l[d];
            }
            for (int k = sp[2]; k <= last_r[2]; k++) PLANE(k);
            for (int k = sp[2] - 1; k >= last_l[2]; k--) PLANE(k);
        }
        *sum_nbox += nbox;
        *photon_loss += photon_loss_src;
    } else {
        double photon_loss_src = 0.0;
        for (int d = 0; d < 3; d++) {
            last_r[d] = lastpos_r[d];
            last_l[d] = lastpos_l[d];
        }
        for (int k = sp[2]; k <= last_r[2]; k++) PLANE(k);
        for (int k = sp[2] - 1; k >= last_l[2]; k--) PLANE(k);
        (void)photon_loss_src;
    }
#undef CELL
#undef PLANE
    return visited;
}

/* ---------------------------------------------------------------------------------------------
 * Public entry: ray-trace all sources.
 *   flavour ORACLE_ASORA  : srcpos = int32[3*NumSrc] interleaved xyz, 0-indexed (sourceutils.py:30);
 *                           grids C-ordered; region = octahedron(q_max) & cube; phi zeroed first
 *                           (raytracing.cu:113).
 *   flavour ORACLE_FORTRAN: srcpos = int32[3*NumSrc] = Fortran srcpos(3,NumSrc), 1-indexed;
 *                           grids Fortran-ordered; phi zeroed first (raytracing.f90:95).
 * coldensh_out (N^3) receives the last processed source's outgoing column density
 * (single thread) -- used by the column-density parity tests with NumSrc == 1.
 * Returns the number of (source, cell) updates performed; stats[0]=sum_nbox, stats[1]=photon_loss.
 * nthreads > 1 parallelises over sources with per-thread scratch (not part of the reference,
 * which is serial: raytracing.f90:177); the per-thread partial phi grids are summed in thread order.
 * ------------------------------------------------------------------------------------------- */
static long do_all_sources_impl(int flavour, int opts, const double *srcflux, const int32_t *srcpos, int NumSrc,
                                int N, double R, double sig, double dr, const double *ndens, const double *xh_av,
                                double *phi_out, double *coldensh_out, const double *thin, const double *thick,
                                const double *heat_thin, const double *heat_thick, double *phi_heat_out,
                                int ntab, double minlogtau, double dlogtau, int NumTau, int max_subbox,
                                int subboxsize, float loss_fraction, int nthreads, double *stats)
{
    const consts_t c = make_consts(flavour, opts);
    const size_t n3 = (size_t)N * N * N;
    const int heating = (heat_thin && heat_thick && phi_heat_out);
    const size_t nacc = heating ? 2 * n3 : n3; /* rates and, behind them, heating rates */
    double *phi_ion = heating ? (double *)calloc(nacc, sizeof(double)) : phi_out;
    cellctx_t x;
    long total = 0;
    int sum_nbox = 0;
    double photon_loss = 0.0;
    x.c = &c;
    x.N = N;
    x.fo = (flavour == ORACLE_FORTRAN);
    x.base = (flavour == ORACLE_FORTRAN) ? 1 : 0;
    x.guard = (flavour == ORACLE_FORTRAN);
    x.sig = sig;
    x.dr = dr;
    x.Rmax = R;
    x.ndens = ndens;
    x.xh_av = xh_av;
    x.thin = thin;
    x.thick = thick;
    x.heat_thin = heating ? heat_thin : NULL;
    x.heat_thick = heating ? heat_thick : NULL;
    x.heat_off = n3;
    x.minlogtau = minlogtau;
    x.dlogtau = dlogtau;
    x.NumTau = NumTau;
    x.ntab = ntab;
    /* raytracing.cu:101 */
    const int q_max = (int)ceil(1.73205080757 * fmin(R, 1.73205080757 * N / 2.0));

    memset(phi_ion, 0, sizeof(double) * nacc);
    if (nthreads < 1) nthreads = 1;
#ifndef _OPENMP
    nthreads = 1;
#endif
    if (nthreads == 1) {
        double *slab = coldensh_out;
        memset(slab, 0, sizeof(double) * n3);
        for (int ns = 0; ns < NumSrc; ns++) {
            const int32_t *sp = srcpos + 3 * ns;
            if (flavour == ORACLE_ASORA) {
                total += asora_source(&x, q_max, sp[0], sp[1], sp[2], srcflux[ns], slab, phi_ion);
            } else {
                int spi[3] = {sp[0], sp[1], sp[2]};
                double frate = (opts & ORACLE_OPT_NORMFLUX_BUG) ? srcflux[NumSrc - 1] : srcflux[ns];
                total += fortran_source(&x, opts, spi, frate, srcflux[ns], max_subbox, subboxsize,
                                        loss_fraction, slab, phi_ion, &sum_nbox, &photon_loss);
            }
        }
    } else {
#ifdef _OPENMP
        double **phis = (double **)calloc(nthreads, sizeof(double *));
#pragma omp parallel num_threads(nthreads) reduction(+ : total, sum_nbox, photon_loss)
        {
            int t = omp_get_thread_num();
            double *slab = (double *)calloc(n3, sizeof(double));
            double *phi = (double *)calloc(nacc, sizeof(double));
            phis[t] = phi;
#pragma omp for schedule(dynamic, 1)
            for (int ns = 0; ns < NumSrc; ns++) {
                const int32_t *sp = srcpos + 3 * ns;
                if (flavour == ORACLE_ASORA) {
                    total += asora_source(&x, q_max, sp[0], sp[1], sp[2], srcflux[ns], slab, phi);
                } else {
                    int spi[3] = {sp[0], sp[1], sp[2]};
                    double frate = (opts & ORACLE_OPT_NORMFLUX_BUG) ? srcflux[NumSrc - 1] : srcflux[ns];
                    total += fortran_source(&x, opts, spi, frate, srcflux[ns], max_subbox, subboxsize,
                                            loss_fraction, slab, phi, &sum_nbox, &photon_loss);
                }
            }
            free(slab);
        }
        for (int t = 0; t < nthreads; t++) {
            if (!phis[t]) continue;
            for (size_t p = 0; p < nacc; p++) phi_ion[p] += phis[t][p];
            free(phis[t]);
        }
        free(phis);
#endif
    }
    if (stats) {
        stats[0] = (double)sum_nbox;
        stats[1] = photon_loss;
    }
    if (heating) {
        memcpy(phi_out, phi_ion, sizeof(double) * n3);
        memcpy(phi_heat_out, phi_ion + n3, sizeof(double) * n3);
        free(phi_ion);
    }
    return total;
}

long oracle_do_all_sources(int flavour, int opts, const double *srcflux, const int32_t *srcpos, int NumSrc,
                           int N, double R, double sig, double dr, const double *ndens, const double *xh_av,
                           double *phi_ion, double *coldensh_out, const double *thin, const double *thick,
                           int ntab, double minlogtau, double dlogtau, int NumTau, int max_subbox,
                           int subboxsize, float loss_fraction, int nthreads, double *stats)
{
    return do_all_sources_impl(flavour, opts, srcflux, srcpos, NumSrc, N, R, sig, dr, ndens, xh_av, phi_ion,
                               coldensh_out, thin, thick, NULL, NULL, NULL, ntab, minlogtau, dlogtau, NumTau,
                               max_subbox, subboxsize, loss_fraction, nthreads, stats);
}

/* The same with photo-heating rates (raytracing.f90:52-110 takes the heating tables and phi_heat): phi_heat
 * receives sum over sources of heat / nHI (raytracing.f90:530,537). */
long oracle_do_all_sources_heat(int flavour, int opts, const double *srcflux, const int32_t *srcpos, int NumSrc,
                                int N, double R, double sig, double dr, const double *ndens, const double *xh_av,
                                double *phi_ion, double *phi_heat, double *coldensh_out, const double *thin,
                                const double *thick, const double *heat_thin, const double *heat_thick, int ntab,
                                double minlogtau, double dlogtau, int NumTau, int max_subbox, int subboxsize,
                                float loss_fraction, int nthreads, double *stats)
{
    return do_all_sources_impl(flavour, opts, srcflux, srcpos, NumSrc, N, R, sig, dr, ndens, xh_av, phi_ion,
                               coldensh_out, thin, thick, heat_thin, heat_thick, phi_heat, ntab, minlogtau, dlogtau,
                               NumTau, max_subbox, subboxsize, loss_fraction, nthreads, stats);
}

/* ---------------------------------------------------------------------------------------------
 * Chemistry: chemistry.f90.  The thresholds are single-precision literals stored in real64
 * (chemistry.f90:9-10, :299).
 * ------------------------------------------------------------------------------------------- */
static const double epsilon_x = 1e-14; /* chemistry.f90:8 */

/* chemistry.f90:221-316 */
static void doric(double xh_old, double dt, double temp_p, double rhe, double phi_p, double bh00,
                  double albpow, double colh0, double temph0, double clumping, double *xh, double *xh_av)
{
    double brech0 = clumping * bh00 * pow(temp_p / 1e4, albpow);
    double sqrtt0 = sqrt(temp_p);
    double acolh0 = colh0 * sqrtt0 * exp(-temph0 / temp_p);
    double aphoth0 = phi_p;
    double aih0 = aphoth0 + rhe * acolh0;
    double delth = aih0 + rhe * brech0;
    double eqxh = aih0 / delth;
    double deltht = delth * dt;
    double ee = exp(-deltht);
    double avg_factor;
    double x = (xh_old - eqxh) * ee + eqxh;
    if (x < epsilon_x) x = epsilon_x;
    if (deltht < (double)1.0e-8f)
        avg_factor = 1.0;
    else
        avg_factor = (1.0 - ee) / deltht;
    double xa = eqxh + (xh_old - eqxh) * avg_factor;
    if (xa < epsilon_x) xa = epsilon_x;
    *xh = x;
    *xh_av = xa;
}

/* chemistry.f90:117-204 (isothermal: the temperature criterion is always met) */
static void do_chemistry(double dt, double ndens_p, double temperature_start, double xh_p, double *xh_av_p,
                         double *xh_intermed_p, double phi_ion_p, double bh00, double albpow, double colh0,
                         double temph0, double abu_c, int *nit_out)
{
    const double minimum_fractional_change = (double)1.0e-3f;
    const double minimum_fraction_of_atoms = (double)1.0e-8f;
    double temperature_end = temperature_start, temperature_previous_iteration;
    int nit = 0;
    for (;;) {
        nit++;
        temperature_previous_iteration = temperature_end;
        double xh_av_p_old = *xh_av_p;
        double de = ndens_p * (*xh_av_p + abu_c);
        doric(xh_p, dt, temperature_end, de, phi_ion_p, bh00, albpow, colh0, temph0, 1.0, xh_intermed_p, xh_av_p);
        if ((fabs((*xh_av_p - xh_av_p_old) / (1.0 - *xh_av_p)) < minimum_fractional_change ||
             (1.0 - *xh_av_p < minimum_fraction_of_atoms)) &&
            (fabs((temperature_end - temperature_previous_iteration) / temperature_end) <
             minimum_fractional_change))
            break;
        if (nit > 400) break;
    }
    if (nit_out) *nit_out = nit;
}

/* chemistry.f90:13-48 + :53-110 on flat arrays of n cells (the update is cell-local, so the
 * memory order of the grids is irrelevant as long as all arrays share it).  xh_av and
 * xh_intermed are updated in place, xh_intermed stored first then xh_av (chemistry.f90:107-108),
 * which matters when the caller aliases them (pyc2ray/chemistry.py:85,91).  Returns conv_flag. */
int oracle_global_pass(double dt, const double *ndens, const double *temp, const double *xh, double *xh_av,
                       double *xh_intermed, const double *phi_ion, double bh00, double albpow, double colh0,
                       double temph0, double abu_c, long n, long *nit_total)
{
    const double minimum_fractional_change = (double)1.0e-3f;
    const double minimum_fraction_of_atoms = (double)1.0e-8f;
    int conv_flag = 0;
    long nits = 0;
    for (long p = 0; p < n; p++) {
        double xh_p = xh[p];
        double xh_av_p = xh_av[p];
        double xh_intermed_p = xh_intermed[p];
        double yh_av_p = 1.0 - xh_av_p;
        int nit;
        do_chemistry(dt, ndens[p], temp[p], xh_p, &xh_av_p, &xh_intermed_p, phi_ion[p], bh00, albpow, colh0,
                     temph0, abu_c, &nit);
        nits += nit;
        double xh_av_p_old = xh_av[p];
        if (fabs(xh_av_p - xh_av_p_old) > minimum_fractional_change &&
            fabs((xh_av_p - xh_av_p_old) / yh_av_p) > minimum_fractional_change &&
            yh_av_p > minimum_fraction_of_atoms)
            conv_flag++;
        xh_intermed[p] = xh_intermed_p;
        xh_av[p] = xh_av_p;
    }
    if (nit_total) *nit_total = nits;
    return conv_flag;
}

/* Number of cells the ASORA sweep visits for one source (octahedron(q_max) & cube):
 * raytracing.cu:101,122-123,241.  Used for the "source-cell updates" metric. */
long oracle_cells_per_source(int N, double R)
{
    const int q_max = (int)ceil(1.73205080757 * fmin(R, 1.73205080757 * N / 2.0));
    const int last_r = N / 2 - 1 + modulo(N, 2), last_l = -N / 2;
    long cnt = 0;
    for (int i = last_l; i <= last_r; i++)
        for (int j = last_l; j <= last_r; j++) {
            int rem = q_max - abs(i) - abs(j);
            if (rem < 0) continue;
            int lo = -rem < last_l ? last_l : -rem, hi = rem > last_r ? last_r : rem;
            cnt += hi - lo + 1;
        }
    return cnt;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
