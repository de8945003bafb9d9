// chemistry.cu -- per-cell implicit hydrogen ionisation solve on the GPU.
//
// Replaces the serial Fortran triple loop global_pass -> evolve0D_global -> do_chemistry -> doric
// (src/c2ray/chemistry.f90:13-316).  One thread per cell, fully coalesced fp64 streams:
// 5 reads + 2 writes = 56 B per cell per pass, HBM-bound.  The convergence counter and the two sums
// the evolve loop needs (pyc2ray/evolve.py:210-217) are fused in and reduced deterministically:
// fixed-shape tree inside a block, per-block partials, then one block adds the partials in order.
#include "asora_common.cuh"

#define CHEM_BLOCK 256

// chemistry.f90:8-10 -- the two thresholds are single-precision literals stored in real64
#define CHEM_EPSILON 1e-14
#define CHEM_MIN_FRAC_CHANGE ((double)1.0e-3f)
#define CHEM_MIN_FRAC_ATOMS ((double)1.0e-8f)

// The two temperature-only factors of doric (chemistry.f90:257-262): recombination and collisional ionisation
// coefficients.  They cost a pow, a sqrt and an exp per cell, more than the rest of a converged cell's update, and the
// run is isothermal: the convergence loop of one time step calls global_pass ~30 times with the same temperatures.
// PRE = true reads them from a grid filled once per (temperature grid, constants) by temperature_factors_kernel.
__device__ __forceinline__ double2 temperature_factors(double temp_p, double bh00, double albpow, double colh0, double temph0)
{
    const double brech0 = 1.0 * bh00 * pow(temp_p / 1e4, albpow);                 // clumping = 1 (chemistry.f90:168)
    const double acolh0 = colh0 * sqrt(temp_p) * exp(-temph0 / temp_p);
    return make_double2(brech0, acolh0);
}

__global__ void __launch_bounds__(CHEM_BLOCK)
temperature_factors_kernel(const double* __restrict__ temp, double2* __restrict__ factors, double bh00, double albpow,
                           double colh0, double temph0, int64_t ncell)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ncell; p += (int64_t)gridDim.x * blockDim.x)
        factors[p] = temperature_factors(temp[p], bh00, albpow, colh0, temph0);
}

template <bool PRE>
__global__ void __launch_bounds__(CHEM_BLOCK)
global_pass_kernel(double dt, const double* __restrict__ ndens, const double* __restrict__ temp,
                   const double2* __restrict__ factors,
                   const double* xh, double* xh_av, double* xh_intermed, const double* __restrict__ phi_ion,
                   double bh00, double albpow, double colh0, double temph0, double abu_c, int64_t ncell,
                   int store_av_first, double* __restrict__ partials, int* __restrict__ iparts)
{
    int my_flag = 0;
    double my_s1 = 0.0, my_s0 = 0.0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < ncell;
         p += (int64_t)gridDim.x * blockDim.x) {
        // chemistry.f90:81-91
        const double ndens_p = ndens[p];
        const double phi_p = phi_ion[p];
        const double xh_p = xh[p];
        const double xh_av_old = xh_av[p];
        const double yh_av_p = 1.0 - xh_av_old;
        double xh_av_p = xh_av_old;
        double xh_int_p = 0.0;

        // doric, temperature-only factors (chemistry.f90:257-262); isothermal, so loop-invariant
        const double2 tf = PRE ? factors[p] : temperature_factors(temp[p], bh00, albpow, colh0, temph0);
        const double brech0 = tf.x, acolh0 = tf.y;

        // do_chemistry fixed point on the time-averaged electron density (chemistry.f90:143-203)
        int nit = 0;
        for (;;) {
            nit++;
            const double prev = xh_av_p;
            const double de = ndens_p * (xh_av_p + abu_c);  // chemistry.f90:162
            // doric (chemistry.f90:279-311)
            const double aih0 = phi_p + de * acolh0;
            const double delth = aih0 + de * brech0;
            // fast_div: MUFU.RCP64H seed + Newton + residual correction (<= 1 ulp), asora_common.cuh.  delth > 0
            // and 1 - x > 0 are ordinary normal numbers here; the library division's slow path is never needed.
            const double eqxh = fast_div(aih0, delth);
            const double deltht = delth * dt;
            const double ee = exp(-deltht);
            double x = (xh_p - eqxh) * ee + eqxh;
            if (x < CHEM_EPSILON) x = CHEM_EPSILON;
            const double avg_factor = (deltht < (double)1.0e-8f) ? 1.0 : fast_div(1.0 - ee, deltht);
            double xa = eqxh + (xh_p - eqxh) * avg_factor;
            if (xa < CHEM_EPSILON) xa = CHEM_EPSILON;
            xh_int_p = x;
            xh_av_p = xa;
            // chemistry.f90:182-189 (the temperature criterion is identically true: isothermal)
            const double ya = 1.0 - xh_av_p;
            if (ya < CHEM_MIN_FRAC_ATOMS || fabs(fast_div(xh_av_p - prev, ya)) < CHEM_MIN_FRAC_CHANGE) break;
            if (nit > 400) break;  // chemistry.f90:192
        }
        // chemistry.f90:96-104
        if (fabs(xh_av_p - xh_av_old) > CHEM_MIN_FRAC_CHANGE &&
            fabs((xh_av_p - xh_av_old) / yh_av_p) > CHEM_MIN_FRAC_CHANGE && yh_av_p > CHEM_MIN_FRAC_ATOMS)
            my_flag++;
        // chemistry.f90:107-108.  When the caller aliases xh_av and xh_intermed the Fortran result is
        // compiler-dependent (dummy arguments may not alias); the reference's own known answer
        // (tutorials/chemistry_solver.ipynb cell 5) corresponds to xh_intermed being stored last.
        if (store_av_first) {
            xh_av[p] = xh_av_p;
            xh_intermed[p] = xh_int_p;
        } else {
            xh_intermed[p] = xh_int_p;
            xh_av[p] = xh_av_p;
        }
        my_s1 += xh_int_p;
        my_s0 += 1.0 - xh_int_p;
    }
    // deterministic block reduction
    __shared__ double sh1[CHEM_BLOCK], sh0[CHEM_BLOCK];
    __shared__ int shf[CHEM_BLOCK];
    sh1[threadIdx.x] = my_s1;
    sh0[threadIdx.x] = my_s0;
    shf[threadIdx.x] = my_flag;
    __syncthreads();
    for (int s = CHEM_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sh1[threadIdx.x] += sh1[threadIdx.x + s];
            sh0[threadIdx.x] += sh0[threadIdx.x + s];
            shf[threadIdx.x] += shf[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x + 0] = sh1[0];
        partials[2 * blockIdx.x + 1] = sh0[0];
        iparts[blockIdx.x] = shf[0];
    }
}

// One block adds the per-block partials in index order -> run-to-run identical sums.
__global__ void __launch_bounds__(CHEM_BLOCK)
global_pass_finish_kernel(const double* __restrict__ partials, const int* __restrict__ iparts, int nparts,
                          double* __restrict__ out_sums, int* __restrict__ out_flag)
{
    __shared__ double sh1[CHEM_BLOCK], sh0[CHEM_BLOCK];
    __shared__ int shf[CHEM_BLOCK];
    double a = 0.0, b = 0.0;
    int f = 0;
    for (int i = threadIdx.x; i < nparts; i += CHEM_BLOCK) {
        a += partials[2 * i];
        b += partials[2 * i + 1];
        f += iparts[i];
    }
    sh1[threadIdx.x] = a;
    sh0[threadIdx.x] = b;
    shf[threadIdx.x] = f;
    __syncthreads();
    for (int s = CHEM_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sh1[threadIdx.x] += sh1[threadIdx.x + s];
            sh0[threadIdx.x] += sh0[threadIdx.x + s];
            shf[threadIdx.x] += shf[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out_sums[0] = sh1[0];
        out_sums[1] = sh0[0];
        *out_flag = shf[0];
    }
}

int chemistry_partial_blocks(int64_t ncell)
{
    // enough CTAs to fill 148 SMs x 8 resident blocks; grid-stride beyond that
    int64_t want = (ncell + CHEM_BLOCK - 1) / CHEM_BLOCK;
    const int64_t cap = 148 * 8 * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

// partials: 2*nblocks doubles followed by 2 result doubles; iparts: nblocks ints followed by 1 result.
cudaError_t launch_temperature_factors(const double* temp, double2* factors, double bh00, double albpow, double colh0,
                                       double temph0, int64_t ncell, cudaStream_t stream)
{
    temperature_factors_kernel<<<chemistry_partial_blocks(ncell), CHEM_BLOCK, 0, stream>>>(temp, factors, bh00, albpow,
                                                                                        colh0, temph0, ncell);
    return cudaGetLastError();
}

// `factors` may be null: the temperature factors are then evaluated per cell inside the pass.
cudaError_t launch_global_pass(double dt, const double* ndens, const double* temp, const double2* factors,
                               const double* xh, double* xh_av, double* xh_intermed, const double* phi_ion, double bh00,
                               double albpow, double colh0, double temph0, double abu_c, int64_t ncell,
                               int store_av_first, double* d_partials, int* d_iparts, int nblocks_max,
                               int* conv_flag, double* sum1, double* sum0, cudaStream_t stream)
{
    int nb = chemistry_partial_blocks(ncell);
    if (nb > nblocks_max) nb = nblocks_max;
    if (factors)
        global_pass_kernel<true><<<nb, CHEM_BLOCK, 0, stream>>>(dt, ndens, temp, factors, xh, xh_av, xh_intermed, phi_ion,
                                                                bh00, albpow, colh0, temph0, abu_c, ncell, store_av_first,
                                                                d_partials, d_iparts);
    else
        global_pass_kernel<false><<<nb, CHEM_BLOCK, 0, stream>>>(dt, ndens, temp, factors, xh, xh_av, xh_intermed, phi_ion,
                                                                 bh00, albpow, colh0, temph0, abu_c, ncell, store_av_first,
                                                                 d_partials, d_iparts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    global_pass_finish_kernel<<<1, CHEM_BLOCK, 0, stream>>>(d_partials, d_iparts, nb, d_partials + 2 * nblocks_max,
                                                            d_iparts + nblocks_max);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    double sums[2];
    int flag;
    e = cudaMemcpyAsync(sums, d_partials + 2 * nblocks_max, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(&flag, d_iparts + nblocks_max, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    if (conv_flag) *conv_flag = flag;
    if (sum1) *sum1 = sums[0];
    if (sum0) *sum0 = sums[1];
    return cudaSuccess;
}
