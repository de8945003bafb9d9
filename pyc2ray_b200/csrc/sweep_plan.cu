// sweep_plan.cu -- host-side construction of the source-independent sweep plan.
//
// The geometry of the short-characteristics interpolation depends only on the offset of a cell from
// its source (src/asora/raytracing.cu:370-386,397-408,444), so it is tabulated once per (N, R, dr)
// and shared by every source: the sweep kernel then does no index arithmetic beyond the periodic
// wrap of one cell.  Cells are grouped in Chebyshev levels (see asora_common.cuh).  Inside a level the
// cells that receive a rate (inside the R sphere) come first, so that whole warps skip the rate
// arithmetic for the octahedron's corners; within each group the order is lexicographic in
// (di,dj,dk), so that consecutive threads touch consecutive k (the contiguous axis of the grids,
// raytracing.cu:30) on four of the six faces of a level, and the upstream slots of consecutive
// cells are consecutive shared-memory words.
#include "asora_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

// raytracing.cu:101  (SQRT3 is the 12-digit literal of raytracing.cu:14)
int asora_qmax(int N, double R)
{
    return (int)std::ceil(ASORA_SQRT3 * std::min(R, ASORA_SQRT3 * N / 2.0));
}

static inline void clip_bounds(int N, int& last_l, int& last_r)
{
    last_r = N / 2 - 1 + (N % 2);  // raytracing.cu:122
    last_l = -N / 2;               // raytracing.cu:123
}

// |octahedron(q_max) & cube|  (raytracing.cu:202,241)
int64_t asora_count_cells(int N, double R)
{
    const int q = asora_qmax(N, R);
    int ll, lr;
    clip_bounds(N, ll, lr);
    int64_t cnt = 0;
    for (int i = ll; i <= lr; i++)
        for (int j = ll; j <= lr; j++) {
            int rem = q - std::abs(i) - std::abs(j);
            if (rem < 0) continue;
            int lo = std::max(-rem, ll), hi = std::min(rem, lr);
            cnt += hi - lo + 1;
        }
    return cnt;
}

static inline int sign1(int x) { return x >= 0 ? 1 : -1; }  // raytracing.cu:27

// Sphere test exactly as the reference kernel evaluates it (raytracing.cu:302-305,315), nvcc contracting
// xs*xs+ys*ys+zs*zs into DMUL,DFMA,DFMA (confirmed against the reference kernel on B200:
// tests/test_gpu_vs_reference_kernel.py, case r_int5).
static inline bool cell_rated(int i, int j, int k, double dr, double R2)
{
    const double xs = dr * (double)i, ys = dr * (double)j, zs = dr * (double)k;
    const double dist2 = std::fma(zs, zs, std::fma(ys, ys, xs * xs));
    return dist2 / (dr * dr) <= R2;
}

// cells of octahedron(q_max) & cube that receive a rate
int64_t asora_count_rated_cells(int N, double R, double dr)
{
    const int Q = asora_qmax(N, R);
    int ll, lr;
    clip_bounds(N, ll, lr);
    const int lo = std::max(ll, -Q), hi = std::min(lr, Q);
    int64_t cnt = 0;
    for (int i = lo; i <= hi; i++)
        for (int j = lo; j <= hi; j++)
            for (int k = lo; k <= hi; k++)
                if (std::abs(i) + std::abs(j) + std::abs(k) <= Q && cell_rated(i, j, k, dr, R * R)) cnt++;
    return cnt;
}

// Upstream corners whose bilinear weight is exactly zero (SURVEY note N3) are pointed at a reserved slot behind the last
// cell of the largest level, which the kernels keep at 0: the weights are formed as products over the other corners'
// max(0.6, tau) (interp_weighted), so whatever a zero-weight corner holds cancels only mathematically, not bit for bit --
// a fixed 0 makes a cell's optical depth independent of how the sweep is split and enumerated (deterministic mode).
// The source cell "interpolates" that slot with weight 1.  Returns the buffer length per level: largest level + 1.
static const int kZeroSlot = 0xffff;
static int resolve_zero_slot(std::vector<PlanCell>& cells, int maxc)
{
    for (PlanCell& c : cells)
        for (int t = 0; t < 4; t++)
            if (c.nb[t] == kZeroSlot) c.nb[t] = (uint16_t)maxc;
    return maxc + 1;
}

// Device copy of plan.cells as two 16-byte streams + the 4-byte offsets stream, and the level bounds.
static bool upload_plan(SweepPlan& plan, std::string& err)
{
    const int64_t total = (int64_t)plan.cells.size();
    std::vector<int4> soa(2 * (size_t)total);
    for (int64_t e = 0; e < total; e++) {
        const int4* src = reinterpret_cast<const int4*>(&plan.cells[e]);
        soa[e] = src[1];
        soa[total + e] = src[2];
    }
    std::vector<unsigned> dwords((size_t)total);
    for (int64_t e = 0; e < total; e++) dwords[e] = (unsigned)soa[total + e].z;
    cudaError_t e = cudaMalloc(&plan.d_cells, sizeof(int4) * soa.size());
    if (e == cudaSuccess) e = cudaMalloc(&plan.d_dwords, sizeof(unsigned) * dwords.size());
    if (e == cudaSuccess)
        e = cudaMemcpy(plan.d_dwords, dwords.data(), sizeof(unsigned) * dwords.size(), cudaMemcpyHostToDevice);
    std::vector<int> bounds(plan.level_start);  // level bounds, then (octant plans) the class boundaries
    bounds.insert(bounds.end(), plan.level_mid.begin(), plan.level_mid.end());
    if (e == cudaSuccess) e = cudaMalloc(&plan.d_level_start, sizeof(int) * bounds.size());
    if (e == cudaSuccess)
        e = cudaMemcpy(plan.d_cells, soa.data(), sizeof(int4) * soa.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(plan.d_level_start, bounds.data(), sizeof(int) * bounds.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        err = std::string("sweep plan upload: ") + cudaGetErrorString(e);
        free_sweep_plan(plan);
        return false;
    }
    plan.valid = true;
    return true;
}

// Cells of the largest level of the largest octant (the all-negative one on even meshes) of an eight-part plan,
// counted without building it: decides whether a large radius fits the shared-memory variant before a plan of
// up to N^3 entries is made.
int sweep_plan_octant_level_cells(int N, double R, double dr, bool sphere_only)
{
    const int Q = asora_qmax(N, R);
    int ll, lr;
    clip_bounds(N, ll, lr);
    const int lo = std::max(ll, -Q);
    const double R2 = R * R;
    std::vector<int> count((size_t)(-lo) + 1, 0);
    for (int i = lo; i <= 0; i++)
        for (int j = lo; j <= 0; j++) {
            const int kmin = std::max(lo, -(Q + i + j));  // |i| + |j| + |k| <= Q
            for (int k = kmin; k <= 0; k++) {
                if (sphere_only && !cell_rated(i, j, k, dr, R2)) continue;
                count[(size_t)std::max(-i, std::max(-j, -k))]++;
            }
        }
    int maxc = 0;
    for (int c : count) maxc = std::max(maxc, c);
    return maxc + 1;  // + the zero slot
}

// One sweep may be split into `parts` in {1,2,4,8} independent pieces by the signs of the offsets on z,
// (y,z) or (x,y,z): every non-zero interpolation weight points one step towards the source, so a cell's
// upstream cells have offsets of the same sign or zero, and a zero offset only ever pairs with the weight of
// the unstepped corner.  A part therefore consists of its open half-space / quadrant / octant plus the
// bounding planes (offset 0 on a constrained axis); the planes are recomputed by every part that touches
// them but receive their rate from the all-positive side only ("owned" cells keep PC_RATED).  Parts of one
// source run as separate CTAs with 1/parts of the shared memory each.
bool build_sweep_plan(SweepPlan& plan, int N, double R, double dr, bool sphere_only, int parts, std::string& err, bool upload)
{
    free_sweep_plan(plan);
    if (!(parts == 1 || parts == 2 || parts == 4 || parts == 8)) {
        err = "sweep plan: parts must be 1, 2, 4 or 8";
        return false;
    }
    const int Q = asora_qmax(N, R);
    int ll, lr;
    clip_bounds(N, ll, lr);
    const int lo = std::max(ll, -Q), hi = std::min(lr, Q);
    const int side = hi - lo + 1;
    if (side > 256) {  // positions inside the swept cube are stored as bytes
        err = "sweep plan: radius too large for the shared-memory variant";
        return false;
    }
    int nlevels = std::max(-lo, hi) + 1;
    const double R2 = R * R;
    auto rated = [&](int i, int j, int k) { return cell_rated(i, j, k, dr, R2); };
    auto level_of = [](int i, int j, int k) { return std::max(std::abs(i), std::max(std::abs(j), std::abs(k))); };
    // membership of the sweep: inside the octahedron, and inside the sphere when only rated cells are swept
    auto member = [&](int i, int j, int k) {
        if (std::abs(i) + std::abs(j) + std::abs(k) > Q) return false;
        return !sphere_only || rated(i, j, k);
    };
    {   // trailing levels may be empty in sphere-only mode
        std::vector<char> used(nlevels, 0);
        for (int i = lo; i <= hi; i++)
            for (int j = lo; j <= hi; j++)
                for (int k = lo; k <= hi; k++)
                    if (member(i, j, k)) used[level_of(i, j, k)] = 1;
        while (nlevels > 1 && !used[nlevels - 1]) nlevels--;
    }
    // constrained axes: bit 0 of the part index = sign on z, bit 1 = y, bit 2 = x (0: d >= 0, 1: d <= 0)
    const int nbits = parts == 1 ? 0 : (parts == 2 ? 1 : (parts == 4 ? 2 : 3));
    auto in_part = [&](int part, int i, int j, int k, bool& owned) {
        const int d[3] = {k, j, i};
        owned = true;
        for (int b = 0; b < nbits; b++) {
            const bool neg = (part >> b) & 1;
            if (neg ? d[b] > 0 : d[b] < 0) return false;
            if (neg && d[b] == 0) owned = false;  // bounding plane: rated by the positive side only
        }
        return true;
    };

    plan.level_start.assign((size_t)parts * (nlevels + 1), 0);
    plan.cells.clear();
    int maxc = 0;
    std::vector<int32_t> slot((size_t)side * side * side);
    auto sidx = [&](int i, int j, int k) { return ((size_t)(i - lo) * side + (j - lo)) * side + (k - lo); };

    for (int part = 0; part < parts; part++) {
        std::fill(slot.begin(), slot.end(), -1);
        // bounding box of the part (the loops below visit nothing else)
        int blo[3] = {lo, lo, lo}, bhi[3] = {hi, hi, hi};  // x, y, z
        for (int b = 0; b < nbits; b++) {
            if ((part >> b) & 1) bhi[2 - b] = 0;
            else blo[2 - b] = 0;
        }
        // pass 1: level sizes and slots (rank inside the level; rate-receiving cells first, then lexicographic)
        std::vector<int> count(nlevels, 0), nfirst(nlevels, 0), fill_a(nlevels, 0), fill_b(nlevels, 0);
        for (int i = blo[0]; i <= bhi[0]; i++)
            for (int j = blo[1]; j <= bhi[1]; j++)
                for (int k = blo[2]; k <= bhi[2]; k++) {
                    bool owned;
                    if (!in_part(part, i, j, k, owned) || !member(i, j, k)) continue;
                    const int m = level_of(i, j, k);
                    count[m]++;
                    if (owned && rated(i, j, k)) nfirst[m]++;
                }
        for (int i = blo[0]; i <= bhi[0]; i++)
            for (int j = blo[1]; j <= bhi[1]; j++)
                for (int k = blo[2]; k <= bhi[2]; k++) {
                    bool owned;
                    if (!in_part(part, i, j, k, owned) || !member(i, j, k)) continue;
                    const int m = level_of(i, j, k);
                    slot[sidx(i, j, k)] = (owned && rated(i, j, k)) ? fill_a[m]++ : nfirst[m] + fill_b[m]++;
                }
        int* ls = plan.level_start.data() + (size_t)part * (nlevels + 1);
        ls[0] = (int)plan.cells.size();
        for (int m = 0; m < nlevels; m++) {
            ls[m + 1] = ls[m] + count[m];
            maxc = std::max(maxc, count[m]);
        }
        if (maxc > 65533) {
            err = "sweep plan: level too large for 16-bit slots";
            return false;
        }
        plan.cells.resize((size_t)ls[nlevels]);
        // pass 2: geometry
        for (int i = blo[0]; i <= bhi[0]; i++)
            for (int j = blo[1]; j <= bhi[1]; j++)
                for (int k = blo[2]; k <= bhi[2]; k++) {
                    bool owned;
                    if (!in_part(part, i, j, k, owned) || !member(i, j, k)) continue;
                    const int ia = std::abs(i), ja = std::abs(j), ka = std::abs(k);
                    const int m = level_of(i, j, k);
                    PlanCell pc;
                    pc.d[0] = (uint8_t)(i - lo);
                    pc.d[1] = (uint8_t)(j - lo);
                    pc.d[2] = (uint8_t)(k - lo);
                    pc.ab = 0;
                    pc.flags = 0;
                    pc.nb[0] = pc.nb[1] = pc.nb[2] = pc.nb[3] = kZeroSlot;
                    if (m == 0) {
                        // source cell: no incoming column, path dr/2, volume dr^3 (raytracing.cu:285-294).
                        // wA = wB = 0 makes it "interpolate" the zero slot with weight 1.
                        pc.flags = PC_SOURCE | (owned ? PC_RATED : 0u);
                        pc.wA = pc.wB = 0.0;
                        pc.path = 0.5;
                        pc.inv_np = ASORA_FOURPI;
                    } else {
                        const int si = sign1(i), sj = sign1(j), sk = sign1(k);
                        const int im = i - si, jm = j - sj, km = k - sk;
                        int a, b, c;                     // |minor A|, |minor B|, |dominant|
                        int n1[3], n2[3], n3[3], n4[3];  // upstream cells c1..c4
                        // dominant-axis selection with the reference's tie order (raytracing.cu:394,446,491)
                        if (ka >= ja && ka >= ia) {
                            a = ia; b = ja; c = ka;  // A = x, B = y (raytracing.cu:416-419)
                            n1[0] = im; n1[1] = jm; n1[2] = km;
                            n2[0] = i;  n2[1] = jm; n2[2] = km;
                            n3[0] = im; n3[1] = j;  n3[2] = km;
                            n4[0] = i;  n4[1] = j;  n4[2] = km;
                        } else if (ja >= ia && ja >= ka) {
                            a = ia; b = ka; c = ja;  // A = x, B = z (raytracing.cu:464-467)
                            n1[0] = im; n1[1] = jm; n1[2] = km;
                            n2[0] = i;  n2[1] = jm; n2[2] = km;
                            n3[0] = im; n3[1] = jm; n3[2] = k;
                            n4[0] = i;  n4[1] = jm; n4[2] = k;
                        } else {
                            a = ja; b = ka; c = ia;  // A = y, B = z (raytracing.cu:509-512)
                            n1[0] = im; n1[1] = jm; n1[2] = km;
                            n2[0] = im; n2[1] = j;  n2[2] = km;
                            n3[0] = im; n3[1] = jm; n3[2] = k;
                            n4[0] = im; n4[1] = j;  n4[2] = k;
                        }
                        // With dx = 1 - a/c (raytracing.cu:397-403 in source-relative coordinates) the
                        // bilinear weights are s1 = wA*wB, s2 = wB*(1-wA), s3 = wA*(1-wB), s4 = (1-wA)*(1-wB).
                        pc.ab = (uint32_t)a | ((uint32_t)b << 8);
                        pc.wA = (double)a / (double)c;
                        pc.wB = (double)b / (double)c;
                        const double da = a, db = b, dc = c;
                        pc.path = std::sqrt((da * da + db * db) / (dc * dc) + 1.0);  // raytracing.cu:444
                        const int n = ia * ia + ja * ja + ka * ka;
                        pc.inv_np = 1.0 / ((double)n * pc.path);
                        if (c == 1 && (a == 1 || b == 1)) pc.flags |= (a == 1 && b == 1) ? PC_DIAG3 : PC_DIAG2;
                        if (owned && rated(i, j, k)) pc.flags |= PC_RATED;
                        if (ka == m && ja < m && ia < m) pc.flags |= PC_ZFACE;
                        // upstream slots; zero-weight corners (which may fall outside the part) -> the zero slot
                        const double s[4] = {pc.wA * pc.wB, pc.wB * (1.0 - pc.wA), pc.wA * (1.0 - pc.wB),
                                             (1.0 - pc.wA) * (1.0 - pc.wB)};
                        int* nn[4] = {n1, n2, n3, n4};
                        for (int t = 0; t < 4; t++) {
                            int sl = kZeroSlot;
                            if (s[t] != 0.0) {
                                const int* q = nn[t];
                                const bool in = q[0] >= lo && q[0] <= hi && q[1] >= lo && q[1] <= hi && q[2] >= lo && q[2] <= hi;
                                const int32_t v = in ? slot[sidx(q[0], q[1], q[2])] : -1;
                                if (v < 0 || level_of(q[0], q[1], q[2]) != m - 1) {
                                    err = "sweep plan: internal error, upstream cell not in the previous level of its part";
                                    return false;
                                }
                                sl = v;
                            }
                            pc.nb[t] = (uint16_t)sl;
                        }
                    }
                    plan.cells[(size_t)ls[m] + slot[sidx(i, j, k)]] = pc;
                }
    }
    plan.N = N;
    plan.R = R;
    plan.dr = dr;
    plan.q_max = Q;
    plan.nlevels = nlevels;
    plan.max_level_cells = resolve_zero_slot(plan.cells, maxc);
    plan.lo = lo;
    plan.side = side;
    plan.sphere_only = sphere_only;
    plan.parts = parts;
    plan.ncells = (int64_t)plan.cells.size();

    plan.octant = false;
    plan.level_mid.clear();
    return upload ? upload_plan(plan, err) : true;
}

// ---------------------------------------------------------------------------------------------------
// Octant plan (sweep_octant.cu)
// ---------------------------------------------------------------------------------------------------
// Everything the plan tabulates is invariant under flipping the sign of any offset: the dominant-axis choice and its
// tie order compare absolute values (raytracing.cu:394,446,491), the fractions, the path and the sphere test
// (raytracing.cu:302-305: squares) depend on |di|, |dj|, |dk| only, and the upstream cells are one step towards the
// source on every axis.  The octant plan therefore lists only the closed positive octant (di, dj, dk >= 0) of
// octahedron(q_max) & cube, and the kernel applies every entry to its up to eight mirror images, which share the
// entry's fetch, decode and interpolation weights.  The eight images keep separate level buffers with identical slot
// numbering, so the upstream slots of an entry are the same numbers in every image.  A cell with a zero offset lies on
// the plane between two images: it is evaluated once (by the image with a clear sign bit on that axis, which also owns
// its rate) and stored into the buffers of all the images it borders; `flags >> 5` is the mask of zero offsets in the
// octant-bit order (bit 2 = x, bit 1 = y, bit 0 = z).  Needs a mirror-symmetric cell set: -lo == hi on every axis
// (always for odd N; for even N while q_max <= N/2 - 1).
// Order inside a level: by zero mask in the order 0, 4, 2, 6, 1, 5, 3, 7 (bit 0 most significant), so that for a thread
// that iterates the images of the low log2(OPT) bits itself the cells without a zero offset on those axes ("class A":
// all OPT images distinct) form a prefix of the level and the plane cells it can de-duplicate ("class B") the rest;
// plan.level_mid holds the three boundaries (OPT = 8, 4, 2).  Inside each zero-mask group rated cells come first
// (whole warps skip the rate arithmetic of the octahedron's corners), then lexicographic order with k fastest.
bool build_octant_plan(SweepPlan& plan, int N, double R, double dr, bool sphere_only, std::string& err, bool upload)
{
    free_sweep_plan(plan);
    const int Q = asora_qmax(N, R);
    int ll, lr;
    clip_bounds(N, ll, lr);
    const int lo = std::max(ll, -Q), hi = std::min(lr, Q);
    if (-lo != hi) {
        err = "octant plan: the swept region is not mirror-symmetric (even mesh, q_max > N/2 - 1)";
        return false;
    }
    if (hi > 126) {
        err = "octant plan: radius too large for byte-sized offsets";
        return false;
    }
    const int side = hi + 1;  // offsets 0..hi
    const double R2 = R * R;
    auto rated = [&](int i, int j, int k) { return cell_rated(i, j, k, dr, R2); };
    auto level_of = [](int i, int j, int k) { return std::max(i, std::max(j, k)); };
    auto member = [&](int i, int j, int k) {
        if (i + j + k > Q) return false;
        return !sphere_only || rated(i, j, k);
    };
    auto zmask_of = [](int i, int j, int k) { return (i == 0 ? 4 : 0) | (j == 0 ? 2 : 0) | (k == 0 ? 1 : 0); };
    int nlevels = hi + 1;
    {
        std::vector<char> used(nlevels, 0);
        for (int i = 0; i <= hi; i++)
            for (int j = 0; j <= hi; j++)
                for (int k = 0; k <= hi; k++)
                    if (member(i, j, k)) used[level_of(i, j, k)] = 1;
        while (nlevels > 1 && !used[nlevels - 1]) nlevels--;
    }
    // group of a cell inside its level: 2 * bit-reversed zmask + (unrated)
    auto group_of = [&](int i, int j, int k) {
        const int z = zmask_of(i, j, k);
        const int zr = ((z & 1) << 2) | (z & 2) | ((z >> 2) & 1);
        return 2 * zr + (rated(i, j, k) ? 0 : 1);
    };
    std::vector<int> gcount((size_t)nlevels * 16, 0);
    for (int i = 0; i <= hi; i++)
        for (int j = 0; j <= hi; j++)
            for (int k = 0; k <= hi; k++)
                if (member(i, j, k)) gcount[(size_t)level_of(i, j, k) * 16 + group_of(i, j, k)]++;
    plan.level_start.assign((size_t)nlevels + 1, 0);
    std::vector<int> gstart((size_t)nlevels * 16, 0);
    int maxc = 0;
    for (int m = 0; m < nlevels; m++) {
        int c = 0;
        for (int g = 0; g < 16; g++) {
            gstart[(size_t)m * 16 + g] = c;
            c += gcount[(size_t)m * 16 + g];
        }
        plan.level_start[m + 1] = plan.level_start[m] + c;
        maxc = std::max(maxc, c);
    }
    // class boundaries: groups 0-1 have zmask 0; 0-3 have no zero on y, z; 0-7 have no zero on z
    plan.level_mid.assign((size_t)3 * nlevels, 0);
    for (int m = 0; m < nlevels; m++) {
        plan.level_mid[(size_t)0 * nlevels + m] = plan.level_start[m] + gstart[(size_t)m * 16 + 2];   // OPT = 8
        plan.level_mid[(size_t)1 * nlevels + m] = plan.level_start[m] + gstart[(size_t)m * 16 + 4];   // OPT = 4
        plan.level_mid[(size_t)2 * nlevels + m] = plan.level_start[m] + gstart[(size_t)m * 16 + 8];   // OPT = 2
    }
    if (maxc > 65533) {
        err = "octant plan: level too large for 16-bit slots";
        return false;
    }
    std::vector<int32_t> slot((size_t)side * side * side, -1);
    auto sidx = [&](int i, int j, int k) { return ((size_t)i * side + j) * side + k; };
    {
        std::vector<int> fill((size_t)nlevels * 16, 0);
        for (int i = 0; i <= hi; i++)
            for (int j = 0; j <= hi; j++)
                for (int k = 0; k <= hi; k++)
                    if (member(i, j, k)) {
                        const size_t g = (size_t)level_of(i, j, k) * 16 + group_of(i, j, k);
                        slot[sidx(i, j, k)] = gstart[g] + fill[g]++;
                    }
    }
    plan.cells.assign((size_t)plan.level_start[nlevels], PlanCell());
    for (int i = 0; i <= hi; i++)
        for (int j = 0; j <= hi; j++)
            for (int k = 0; k <= hi; k++) {
                if (!member(i, j, k)) continue;
                const int m = level_of(i, j, k);
                PlanCell pc;
                pc.d[0] = (uint8_t)i;
                pc.d[1] = (uint8_t)j;
                pc.d[2] = (uint8_t)k;
                pc.ab = 0;
                pc.flags = (uint8_t)(zmask_of(i, j, k) << 5);
                pc.nb[0] = pc.nb[1] = pc.nb[2] = pc.nb[3] = kZeroSlot;
                if (m == 0) {  // source cell (raytracing.cu:285-294), see build_sweep_plan
                    pc.flags |= PC_SOURCE | PC_RATED;
                    pc.wA = pc.wB = 0.0;
                    pc.path = 0.5;
                    pc.inv_np = ASORA_FOURPI;
                } else {
                    // one step towards the source; a zero offset steps to -1, outside the octant, with weight 0
                    const int im = i - 1, jm = j - 1, km = k - 1;
                    int a, b, c;
                    int n1[3], n2[3], n3[3], n4[3];
                    if (k >= j && k >= i) {  // raytracing.cu:394
                        a = i; b = j; c = k;
                        n1[0] = im; n1[1] = jm; n1[2] = km;
                        n2[0] = i;  n2[1] = jm; n2[2] = km;
                        n3[0] = im; n3[1] = j;  n3[2] = km;
                        n4[0] = i;  n4[1] = j;  n4[2] = km;
                    } else if (j >= i && j >= k) {  // raytracing.cu:446
                        a = i; b = k; c = j;
                        n1[0] = im; n1[1] = jm; n1[2] = km;
                        n2[0] = i;  n2[1] = jm; n2[2] = km;
                        n3[0] = im; n3[1] = jm; n3[2] = k;
                        n4[0] = i;  n4[1] = jm; n4[2] = k;
                    } else {  // raytracing.cu:491
                        a = j; b = k; c = i;
                        n1[0] = im; n1[1] = jm; n1[2] = km;
                        n2[0] = im; n2[1] = j;  n2[2] = km;
                        n3[0] = im; n3[1] = jm; n3[2] = k;
                        n4[0] = im; n4[1] = j;  n4[2] = k;
                    }
                    pc.ab = (uint32_t)a | ((uint32_t)b << 8);
                    pc.wA = (double)a / (double)c;
                    pc.wB = (double)b / (double)c;
                    const double da = a, db = b, dc = c;
                    pc.path = std::sqrt((da * da + db * db) / (dc * dc) + 1.0);  // raytracing.cu:444
                    const int n = i * i + j * j + k * k;
                    pc.inv_np = 1.0 / ((double)n * pc.path);
                    if (c == 1 && (a == 1 || b == 1)) pc.flags |= (a == 1 && b == 1) ? PC_DIAG3 : PC_DIAG2;
                    if (rated(i, j, k)) pc.flags |= PC_RATED;
                    if (k == m && j < m && i < m) pc.flags |= PC_ZFACE;
                    const double s[4] = {pc.wA * pc.wB, pc.wB * (1.0 - pc.wA), pc.wA * (1.0 - pc.wB),
                                         (1.0 - pc.wA) * (1.0 - pc.wB)};
                    int* nn[4] = {n1, n2, n3, n4};
                    for (int t = 0; t < 4; t++) {
                        int sl = kZeroSlot;
                        if (s[t] != 0.0) {
                            const int* q = nn[t];
                            const bool in = q[0] >= 0 && q[1] >= 0 && q[2] >= 0 && q[0] <= hi && q[1] <= hi && q[2] <= hi;
                            const int32_t v = in ? slot[sidx(q[0], q[1], q[2])] : -1;
                            if (v < 0 || level_of(q[0], q[1], q[2]) != m - 1) {
                                err = "octant plan: internal error, upstream cell not in the previous level";
                                return false;
                            }
                            sl = v;
                        }
                        pc.nb[t] = (uint16_t)sl;
                    }
                }
                plan.cells[(size_t)plan.level_start[m] + slot[sidx(i, j, k)]] = pc;
            }
    plan.N = N;
    plan.R = R;
    plan.dr = dr;
    plan.q_max = Q;
    plan.nlevels = nlevels;
    plan.max_level_cells = resolve_zero_slot(plan.cells, maxc);
    plan.lo = -hi;
    plan.side = 2 * hi + 1;
    plan.sphere_only = sphere_only;
    plan.parts = 1;
    plan.octant = true;
    plan.ncells = (int64_t)plan.cells.size();
    return upload ? upload_plan(plan, err) : true;
}

void free_sweep_plan(SweepPlan& plan)
{
    if (plan.d_cells) cudaFree(plan.d_cells);
    if (plan.d_level_start) cudaFree(plan.d_level_start);
    if (plan.d_dwords) cudaFree(plan.d_dwords);
    plan.d_dwords = nullptr;
    plan.d_cells = nullptr;
    plan.d_level_start = nullptr;
    plan.cells.clear();
    plan.level_start.clear();
    plan.level_mid.clear();
    plan.valid = false;
}
