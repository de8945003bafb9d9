// sweep_cluster.cu -- the large-radius / full-box sweep: one thread-block cluster per wedge of a source, level buffers in
// distributed shared memory.
//
// Replaces evolve0D_gpu + cinterp_gpu (src/asora/raytracing.cu:155-535) for radii whose levels do not fit the shared
// memory of one CTA (q_max up to 385 at 256^3, 98 300 cells in the largest level).  The grid-cooperative variant 2
// (sweep_kernels.cu) keeps the column densities of such sweeps in N^3 scratch grids: every update then costs four L2
// gathers, a store and a RED through a 2 GB working set, one grid-wide barrier per level (profiles/r02a: L2 hit rate
// 28 %, 139 B of DRAM traffic and 528 instructions per update, 24 G updates/s).  Here nothing per-source leaves the SMs:
//
//   wedges    A level (cube shell) splits into 6 faces x 4 sign quadrants = 24 wedges {dominant axis = +-m, minor offsets
//             (+-a, +-b), 0 <= a, b <= m}.  Every upstream cell of a wedge cell lies in the same wedge one level down (the
//             dominant offset steps to m-1, a minor offset steps towards 0 or stays), so the 24 wedges of a source are 24
//             independent sweeps whose levels are (m+1) x (m+1) arrays: upstream corners c1..c4 (raytracing.cu:416-419)
//             are simply (a-1,b-1), (a,b-1), (a-1,b), (a,b) of the previous array.  Cells on a wedge's rim (a zero
//             offset, or a minor offset tying the dominant one) belong to several wedges: they are evaluated by each
//             (with the wedge's own axis as the dominant one, which gives the same interpolation: the bilinear weight of
//             an unstepped tying axis is exactly 0) and rated by the one the reference's tie order picks
//             (raytracing.cu:394,446,491) with non-negative signs on the zero offsets.
//   cluster   One wedge = one cluster of C CTAs.  The rows a of a level are dealt to the CTAs in blocks of four
//             (block-cyclic: balanced at every level size); a CTA keeps its rows of the previous and the current level in
//             its own shared memory and reads the rows it does not own -- one in eight -- from its neighbours'
//             shared memory (mapa + ld.shared::cluster).  One cluster barrier per level (barrier.cluster, hardware)
//             instead of a spin on an L2 counter.
//   geometry  on the fly from (m, a, b): two corrected reciprocal products for the fractions, one rsqrt-seeded square
//             root for the path, no plan (a full-box plan would be 0.5 GB at 256^3).
//
// 24 clusters per source run concurrently and independently, so one source already spreads over 24 C CTAs (192 for
// C = 8: every SM of a B200), and several sources overlap through ordinary CTA scheduling.
#include "asora_common.cuh"
#include "sweep_device.cuh"

#include <cooperative_groups.h>

namespace {

__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared-memory address `addr` (this CTA's window) as seen in CTA `rank` of the cluster
__device__ __forceinline__ unsigned map_to_rank(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ double load_cluster(unsigned addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// sqrt(x) for a normal positive x: MUFU.RSQ64H seed, two Newton steps on 1/sqrt(x), one Heron correction (<= 1 ulp).
__device__ __forceinline__ double wedge_sqrt(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double h = 0.5 * x;
    r = fma(r, fma(-h * r, r, 0.5), r);
    r = fma(r, fma(-h * r, r, 0.5), r);
    const double s = x * r;
    return fma(fma(-s, s, x), 0.5 * r, s);
}

// Rows of a level are dealt to the C = 2^LOGC CTAs of a cluster in blocks of four.
template <int LOGC>
__device__ __forceinline__ unsigned row_owner(int a) { return ((unsigned)a >> 2) & ((1u << LOGC) - 1u); }
template <int LOGC>
__device__ __forceinline__ int row_local(int a) { return ((a >> (2 + LOGC)) << 2) | (a & 3); }

// One cluster of 2^LOGC CTAs per (source, wedge); grid = sources x 24 x C.
//   pitch      doubles per row of a level buffer (>= nlevels)
//   rows_cap   rows per CTA and level buffer
template <int LOGC, int BLOCK, bool HEAT, bool DET>
__global__ void __launch_bounds__(BLOCK)
sweep_wedge_kernel(SweepParams p, int nlevels, int pitch, int rows_cap)
{
    constexpr int C = 1 << LOGC;
    extern __shared__ double2 sh_raw[];
    double2* log2_tab = sh_raw;                                   // 256 entries
    double* buf = reinterpret_cast<double*>(sh_raw + 256);        // [2][rows_cap][pitch]
    const int side = 2 * nlevels - 1;                             // offsets -(nlevels-1) .. nlevels-1
    unsigned* wrap_tab = reinterpret_cast<unsigned*>(buf + (size_t)2 * rows_cap * pitch);  // [3][side]

    const int N = p.N;
    const unsigned rank = cluster_rank();
    const int cluster_id = blockIdx.x >> LOGC;
    const int wedge = cluster_id % 24, ns = p.src_begin + cluster_id / 24;
    const int face = wedge >> 3;                 // 0: x dominant, 1: y, 2: z
    const int og = wedge & 7;                    // sign bits: 4 = di negative, 2 = dj, 1 = dk
    // the reference's minor axes A, B of each dominant axis (raytracing.cu:397-403,449-455,494-500)
    const int axisA = face == 0 ? 1 : 0, axisB = face == 2 ? 1 : 2;
    const int sgn_dom = ((og >> (2 - face)) & 1) ? -1 : 1;
    const int sgnA = ((og >> (2 - axisA)) & 1) ? -1 : 1, sgnB = ((og >> (2 - axisB)) & 1) ? -1 : 1;

    for (int t = threadIdx.x; t < 256; t += BLOCK) log2_tab[t] = __ldg(p.log2_tab + t);
    for (int t = threadIdx.x; t < 2 * rows_cap * pitch; t += BLOCK) buf[t] = 0.0;
    const int c0[3] = {p.src_pos[3 * ns + 0], p.src_pos[3 * ns + 1], p.src_pos[3 * ns + 2]};
    const double sk = p.src_flux[ns] * p.kpref;
    for (int t = threadIdx.x; t < 3 * side; t += BLOCK) {
        const int axis = t / side, d = t - axis * side - (nlevels - 1);
        int w = (c0[axis] + d) % N;
        if (w < 0) w += N;
        const unsigned stride = axis == 0 ? (unsigned)N * N : (axis == 1 ? (unsigned)N : 1u);
        wrap_tab[t] = (unsigned)w * stride;
    }
    const unsigned* wrap_dom = wrap_tab + face * side + (nlevels - 1);
    const unsigned* wrapA = wrap_tab + axisA * side + (nlevels - 1);
    const unsigned* wrapB = wrap_tab + axisB * side + (nlevels - 1);
    // integer squared distances this far from R^2 decide the sphere test without evaluating the reference's expression
    const double R2_lo = p.R2 * (1.0 - 1e-12), R2_hi = p.R2 * (1.0 + 1e-12);
    const unsigned buf_addr = (unsigned)__cvta_generic_to_shared(buf);
    cluster_barrier();  // all buffers of the cluster are zeroed before anybody reads a neighbour's

    for (int m = 0; m < nlevels; m++) {
        const int dom = sgn_dom * m;
        const bool level_in_box = dom >= p.last_l && dom <= p.last_r;
        // minor offsets of this level: 0 .. amax (the level's square clipped by the octahedron and by the cube)
        const int amax = min(m, min(p.q_max - m, max(-p.last_l, p.last_r)));
        const int ncols = amax + 1;
        // my rows of this level: blocks g = rank, rank + C, ... of four rows, up to the block that holds row amax
        const int G = amax >> 2;
        int nloc = 0;
        if (level_in_box && amax >= 0 && G >= (int)rank) {
            const int nblk = (G - (int)rank) / C + 1;
            nloc = 4 * nblk - ((((G - (int)rank) % C) == 0) ? 3 - (amax & 3) : 0);
        }
        const int total = nloc * ncols;
        double* cur = buf + (size_t)(m & 1) * rows_cap * pitch;
        const unsigned prev_addr = buf_addr + (unsigned)(((m & 1) ^ 1) * rows_cap * pitch * 8);
        const double md = (double)m, inv_m = m > 0 ? 1.0 / md : 0.0;
        const float inv_cols = 1.0f / (float)ncols;
        const unsigned pos_dom = wrap_dom[dom];
        for (int t = threadIdx.x; t < total; t += BLOCK) {
            const int lr = (int)(((float)t + 0.5f) * inv_cols);   // t / ncols (t < 2^20: exact in fp32)
            const int b = t - lr * ncols;
            const int a = ((((lr >> 2) << LOGC) + (int)rank) << 2) | (lr & 3);
            const int da = sgnA * a, db = sgnB * b;
            // octahedron & cube (raytracing.cu:101,122-123,241)
            if (m + a + b > p.q_max || da < p.last_l || da > p.last_r || db < p.last_l || db > p.last_r) continue;
            const int n = m * m + a * a + b * b;
            const double dn = u2d((unsigned)n);  // (int -> fp64 on the fp64 pipe instead of the conversion unit)
            unsigned flags = 0;
            bool inside;
            if (dn <= R2_lo) {
                inside = true;
            } else if (dn >= R2_hi) {
                inside = false;
            } else {  // on the sphere's surface to rounding: the reference's own form decides (raytracing.cu:302-305,315)
                int d3[3];
                d3[face] = dom;
                d3[axisA] = da;
                d3[axisB] = db;
                const double xs = p.dr * (double)d3[0], ys = p.dr * (double)d3[1], zs = p.dr * (double)d3[2];
                const double dist2 = __fma_rn(zs, zs, __fma_rn(ys, ys, __dmul_rn(xs, xs)));
                inside = dist2 / (p.dr * p.dr) <= p.R2;
            }
            if (!inside && p.sphere_only && m > 0) continue;
            // the wedge that rates the cell: the reference's dominant-axis tie order (z, then y, then x), and the
            // non-negative side of every zero offset
            bool owner = face == 2 ? true : (face == 1 ? b < m : (a < m && b < m));
            if (m == 0) owner = face == 2;
            if ((m == 0 && sgn_dom < 0) || (a == 0 && sgnA < 0) || (b == 0 && sgnB < 0)) owner = false;
            if (m == 0) flags = PC_SOURCE | (owner ? PC_RATED : 0u);
            else if (inside && owner) flags = PC_RATED;
            const unsigned pos = pos_dom + wrapA[da] + wrapB[db];  // N <= 1600: fits 32 bits
            const double ntau = __ldg(p.nhi + pos);
            double cin = 0.0, path = 0.5, inv_np = ASORA_FOURPI;
            if (m > 0) {
                // fractions a/m, b/m correctly rounded (reciprocal + one correction step); path = sqrt(n)/m
                // (raytracing.cu:444); 1/(n path) by reciprocal
                const double ad = u2d((unsigned)a), bd = u2d((unsigned)b);
                const double qa = ad * inv_m, qb = bd * inv_m;
                const double wA = fma(fma(-md, qa, ad), inv_m, qa), wB = fma(fma(-md, qb, bd), inv_m, qb);
                const double sn = wedge_sqrt(dn), qp = sn * inv_m;
                path = fma(fma(-md, qp, sn), inv_m, qp);
                inv_np = fast_rcp(dn * path);
                if (m == 1 && (a == 1 || b == 1)) flags |= (a == 1 && b == 1) ? PC_DIAG3 : PC_DIAG2;
                // upstream rows a-1 and a, columns b-1 and b of the previous level; indices of zero-weight corners are
                // clamped into the array: whatever finite value sits there is multiplied by an exact 0 (the buffers
                // start zeroed and only ever receive optical depths)
                const int am = max(a - 1, 0), au = min(a, m - 1), bm = max(b - 1, 0), bu = min(b, m - 1);
                const unsigned row_m = map_to_rank(prev_addr + (unsigned)(row_local<LOGC>(am) * pitch * 8), row_owner<LOGC>(am));
                const unsigned row_u = map_to_rank(prev_addr + (unsigned)(row_local<LOGC>(au) * pitch * 8), row_owner<LOGC>(au));
                const double c1 = load_cluster(row_m + bm * 8), c2 = load_cluster(row_u + bm * 8);
                const double c3 = load_cluster(row_m + bu * 8), c4 = load_cluster(row_u + bu * 8);
                cin = interp_coldens<false, true>(c1, c2, c3, c4, wA, wB, flags);
            }
            const double cdho = finish_cell<1, false, HEAT, DET>(cin, path, inv_np, flags, ntau, sk, pos, p, log2_tab);
            cur[lr * pitch + b] = cdho;
            if (p.coldens_out) p.coldens_out[pos] = cdho;  // debug path (all wedges that hold the cell store the same value)
        }
        cluster_barrier();
    }
}

template <int LOGC, int BLOCK>
cudaError_t launch_wedge(const SweepParams& p, int nlevels, int pitch, int rows_cap, size_t smem, cudaStream_t stream)
{
    void* kernel;
    if (p.det_lo)
        kernel = p.phi_heat ? (void*)sweep_wedge_kernel<LOGC, BLOCK, true, true> : (void*)sweep_wedge_kernel<LOGC, BLOCK, false, true>;
    else
        kernel = p.phi_heat ? (void*)sweep_wedge_kernel<LOGC, BLOCK, true, false> : (void*)sweep_wedge_kernel<LOGC, BLOCK, false, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.src_count * 24u << LOGC);
    cfg.blockDim = dim3(BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1u << LOGC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SweepParams pc = p;
    void* args[] = {(void*)&pc, (void*)&nlevels, (void*)&pitch, (void*)&rows_cap};
    return cudaLaunchKernelExC(&cfg, kernel, args);
}

}  // namespace

// Levels of the large-radius sweep, as in the grid-cooperative variant (raytracing.cu:198 bounded by the cube).
int sweep_cluster_levels(const SweepParams& p)
{
    int nlevels = std::min(p.q_max, std::max(-p.last_l, p.last_r)) + 1;
    // sphere-only: no cell beyond Chebyshev distance floor(R) can be inside the sphere
    if (p.sphere_only && std::sqrt(p.R2) + 2.0 < (double)nlevels) nlevels = (int)std::sqrt(p.R2) + 2;
    return nlevels;
}

// Shared memory per CTA for clusters of 2^logc CTAs.
size_t sweep_cluster_smem_bytes(int nlevels, int logc)
{
    const int C = 1 << logc;
    const int pitch = nlevels + 1;
    const int blocks = (nlevels + 3) / 4;              // row blocks of the largest level
    const int rows_cap = ((blocks + C - 1) / C) * 4;
    return (size_t)256 * sizeof(double2) + (size_t)2 * rows_cap * pitch * sizeof(double) +
           (size_t)3 * (2 * nlevels - 1) * sizeof(unsigned);
}

cudaError_t launch_sweep_cluster(const SweepParams& p, int logc, int block, cudaStream_t stream, int* launches, int* levels)
{
    if (p.src_count <= 0) return cudaSuccess;
    const int nlevels = sweep_cluster_levels(p);
    if (levels) *levels = nlevels;
    const int C = 1 << logc;
    const int pitch = nlevels + 1;
    const int rows_cap = ((((nlevels + 3) / 4) + C - 1) / C) * 4;
    const size_t smem = sweep_cluster_smem_bytes(nlevels, logc);
    if (launches) *launches += 1;
#define ASORA_WEDGE(LC, BL) if (logc == LC && block == BL) return launch_wedge<LC, BL>(p, nlevels, pitch, rows_cap, smem, stream);
    ASORA_WEDGE(3, 256)
    ASORA_WEDGE(3, 384)
    ASORA_WEDGE(3, 512)
    ASORA_WEDGE(2, 256)
    ASORA_WEDGE(2, 512)
    ASORA_WEDGE(1, 256)
    ASORA_WEDGE(1, 512)
    ASORA_WEDGE(0, 256)
    ASORA_WEDGE(0, 512)
#undef ASORA_WEDGE
    return cudaErrorInvalidValue;
}
