// K2 -- neighbour search on the sorted particles.
//
// Replaces Particles::gridNNS (/root/reference/demonstrator/src/Particles.cpp:324-365),
// Domain::getNeighborCells (Domain.cpp:83-118), Particles::createGhostParticles (:2113-2191) and the
// brute-force Particles::ghostNNS (:2237-2260).
//
// One thread per particle walks the 3^D stencil in the reference's order (x outer, y, z inner);
// a stencil cell is a contiguous range of the sorted arrays, so candidates are read with unit
// stride and lanes of the same cell broadcast.  The cutoff test is the reference's expression,
// evaluated without FMA contraction (dist_sqr_exact), so the neighbour SETS are bit-exact.
//
// Periodic images are not materialised as ghost particles: a stencil cell that wraps around the box
// yields candidates whose image position is computed on the fly with the reference's formulas
// (image_coord) and whose existence test is the reference's threshold (image_exists).  Because
// cellSize >= h (Domain.cpp:10-22) every ghost within h of a particle lives in a wrapped stencil
// cell, so the result equals the brute-force search over all ghosts.  Ghost entries are appended
// after the regular ones and ordered by parent original index (== ascending ghost index).
#include "mlh_internal.cuh"

namespace {

template <int D>
struct StencilCell {
    int cell; // local cell id or -1
    int code; // image code (2 bits per dim), 0 = not wrapped
};

// neighbour cell (cx+ox, cy+oy, cz+oz) of the local grid with periodic wrap information
template <int D, bool PER>
__device__ __forceinline__ StencilCell<D> stencil_cell(const Grid &g, const int *ci, const int *off) {
    StencilCell<D> r;
    r.code = 0;
    int n[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < D; ++k) {
        int v = ci[k] + off[k];
        if (g.sliced && k == g.slab_dim) {
            // local layers: index is always inside the local grid for owned cells; the wrap is
            // decided by the GLOBAL layer
            int gl = g.layer0 + v;
            if (v < 0 || v >= g.lcells[k]) { r.cell = -1; return r; }
            if (gl < 0) {
                if (!PER) { r.cell = -1; return r; }
                r.code |= 2 << (2 * k);
            } else if (gl >= g.cells[k]) {
                if (!PER) { r.cell = -1; return r; }
                r.code |= 1 << (2 * k);
            }
        } else {
            if (v < 0) {
                if (!PER) { r.cell = -1; return r; }
                v = g.cells[k] - 1;
                r.code |= 2 << (2 * k); // parents near the high side, image below min
            } else if (v >= g.cells[k]) {
                if (!PER) { r.cell = -1; return r; }
                v = 0;
                r.code |= 1 << (2 * k); // parents near the low side, image above max
            }
        }
        n[k] = v;
    }
    r.cell = n[0] + g.lcells[0] * (n[1] + g.lcells[1] * n[2]);
    return r;
}

template <int D, bool PER>
__global__ void __launch_bounds__(128) k_neighbours(const Params p) {
    int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.own_end) return;
    const Grid &g = p.grid;
    double xi[3];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = p.d.x[k][i];
    int c = p.d.cell[i];
    int ci[3];
    ci[0] = c % g.lcells[0];
    ci[1] = (c / g.lcells[0]) % g.lcells[1];
    ci[2] = D == 3 ? c / (g.lcells[0] * g.lcells[1]) : 0;

    int cnt = 0;
    bool overflow = false;
    const int idi = p.d.id[i];
    constexpr int NSTENCIL = D == 3 ? 27 : 9;
    // ---- pass A: regular neighbours, Particles.cpp:335-359.  Per stencil cell the thread also records where that
    // cell's group starts in its list (grp) and WHICH particles of the cell it lists (bit k = k-th particle of the
    // cell): with these a partner finds its own slot in this list without a search (k_face_index, k4_flux.cu). ----
    int off[3] = {0, 0, 0};
    int sci = -1; // stencil cell index in the reference's order
    for (off[0] = -1; off[0] <= 1; ++off[0])
        for (off[1] = -1; off[1] <= 1; ++off[1])
            for (off[2] = (D == 3 ? -1 : 0); off[2] <= (D == 3 ? 1 : 0); ++off[2]) {
                ++sci;
                unsigned long long mask = 0ull;
                unsigned short g0 = (unsigned short)(cnt < p.max_ni ? cnt : p.max_ni);
                StencilCell<D> sc = stencil_cell<D, PER>(g, ci, off);
                if (sc.cell >= 0 && sc.code == 0) {
                    int s = p.d.cell_start[sc.cell], e = p.d.cell_start[sc.cell + 1];
                    if (e - s > 8 * p.max_ni + 64) { // collapsed cell (NaN state), see k1
                        overflow = true;
                        e = s;
                    }
                    if (e - s > 64) g0 |= 0x8000u; // more particles than mask bits: the partner searches this group
                    // four candidates per trip: their loads and cutoff tests are independent (the loop is latency-bound),
                    // hits are then recorded in ascending j as the reference does.  (Requesting the NEXT four before testing
                    // these cost 40 more registers and half the occupancy: 0.210 -> 0.233 ms at 61^3, r02e -- dropped.)
                    for (int j0 = s; j0 < e; j0 += 4) {
                        bool hit[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = j0 + u;
                            double d[3];
                            const int jl = j < e ? j : e - 1; // clamped load address for the tail
#pragma unroll
                            for (int k = 0; k < D; ++k) d[k] = __dsub_rn(p.d.x[k][jl], xi[k]);
                            hit[u] = (j < e) && (j != i) && (dist_sqr_exact<D>(d) < p.hSqr);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!hit[u]) continue;
                            const int j = j0 + u;
                            // bits 26..30: stencil cell, bit 31: j has the lower original index -- consumed (and cleared)
                            // by the ownership pass below, which then needs neither id[j] nor a search for the group
                            if (cnt < p.max_ni)
                                p.d.nnl[(size_t)cnt * p.ncap + i] = j | (sci << MLH_NNL_IDX_BITS) | (p.d.id[j] < idi ? (int)0x80000000 : 0);
                            else
                                overflow = true;
                            ++cnt;
                            mask |= 1ull << ((j - s) & 63);
                        }
                    }
                }
                p.d.grp[(size_t)sci * p.ncap + i] = g0;
                p.d.nbm[(size_t)sci * p.ncap + i] = mask;
            }
    int nreg = cnt < p.max_ni ? cnt : p.max_ni;
    p.d.noi[i] = nreg;
    int ng = 0;
    if (PER) {
        // ---- pass B: periodic images, Particles.cpp:2113-2191 + :2237-2260 ----
        for (off[0] = -1; off[0] <= 1; ++off[0])
            for (off[1] = -1; off[1] <= 1; ++off[1])
                for (off[2] = (D == 3 ? -1 : 0); off[2] <= (D == 3 ? 1 : 0); ++off[2]) {
                    StencilCell<D> sc = stencil_cell<D, PER>(g, ci, off);
                    if (sc.cell < 0 || sc.code == 0) continue;
                    int rcode = reverse_code(sc.code);
                    int s = p.d.cell_start[sc.cell], e = p.d.cell_start[sc.cell + 1];
                    for (int j = s; j < e; ++j) {
                        double d[3];
                        bool exists = true;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            int ck = (sc.code >> (2 * k)) & 3;
                            double xj = p.d.x[k][j];
                            exists = exists && image_exists(xj, ck, g.bmin[k], g.bmax[k], p.h);
                            d[k] = __dsub_rn(image_coord(xj, ck, g.bmin[k], g.bmax[k]), xi[k]);
                        }
                        bool hit = exists && (dist_sqr_exact<D>(d) < p.hSqr);
                        if (!hit && p.symmetric_seam) {
                            // the pair as particle j sees it (image of i with the opposite code)
                            bool ex2 = true;
                            double d2[3];
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                int ck = (rcode >> (2 * k)) & 3;
                                ex2 = ex2 && image_exists(xi[k], ck, g.bmin[k], g.bmax[k], p.h);
                                d2[k] = __dsub_rn(image_coord(xi[k], ck, g.bmin[k], g.bmax[k]), p.d.x[k][j]);
                            }
                            hit = ex2 && (dist_sqr_exact<D>(d2) < p.hSqr);
                        }
                        if (hit) {
                            if (cnt < p.max_ni)
                                p.d.nnl[(size_t)cnt * p.ncap + i] = j | (sc.code << MLH_NNL_IDX_BITS);
                            else
                                overflow = true;
                            ++cnt;
                        }
                    }
                }
        int ntot = cnt < p.max_ni ? cnt : p.max_ni;
        ng = ntot - nreg;
        // order ghost entries by parent original index (ghostNNS scans ghosts in creation order,
        // which is ascending parent index, Particles.cpp:2116,2241)
        for (int a = nreg + 1; a < ntot; ++a) {
            int ea = p.d.nnl[(size_t)a * p.ncap + i];
            int ka = p.d.id[ea & MLH_NNL_IDX_MASK];
            int b = a - 1;
            while (b >= nreg) {
                int eb = p.d.nnl[(size_t)b * p.ncap + i];
                if (p.d.id[eb & MLH_NNL_IDX_MASK] <= ka) break;
                p.d.nnl[(size_t)(b + 1) * p.ncap + i] = eb;
                --b;
            }
            p.d.nnl[(size_t)(b + 1) * p.ncap + i] = ea;
        }
    }
    p.d.noig[i] = ng;
    if (overflow) atomicOr(p.d.flags, MLH_F_MAX_INTERACTIONS);
    // ---- face ownership: the flux pass (k4_flux.cu) evaluates every pair once, from its owner, and both endpoints
    // gather +-F.  Owner = the endpoint with the lower ORIGINAL index (the one that solves the face in the reference,
    // Particles.cpp:1841,1889); always this particle when the partner has no list on this rank (halo of another slab)
    // or does not list the pair (one-sided periodic pair, quirk Q9).  A regular slot owned by the partner j carries
    // (index of i inside its cell | stencil cell of i as j sees it << 12) for k_face_index. ----
    {
        const int ntot = nreg + ng;
        const int li = i - p.d.cell_start[c];
        const unsigned li_enc = (unsigned)(li < 0xFFF ? li : 0xFFF);
        int nown = 0;
        for (int s0 = 0; s0 < ntot; s0 += 4) {
          int e4[4]; // the entries of four slots are read back together (independent loads)
#pragma unroll
          for (int q = 0; q < 4; ++q) e4[q] = p.d.nnl[(size_t)(s0 + q < ntot ? s0 + q : ntot - 1) * p.ncap + i];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int s = s0 + q;
            if (s >= ntot) break;
            const size_t at = (size_t)s * p.ncap + i;
            const int e = e4[q];
            const int j = e & MLH_NNL_IDX_MASK;
            const bool regular = !PER || s < nreg;
            const bool canon = regular ? e >= 0 : !(p.d.id[j] < idi);
            if (regular) p.d.nnl[at] = j; // strip the tags of pass A
            bool listed = true; // does j list this pair too?  (the test of pass B from j's side; exact for regular pairs)
            if (PER && s >= nreg && !p.symmetric_seam) {
                const int cview = reverse_code((int)((unsigned)e >> MLH_NNL_IDX_BITS)); // image of i as j sees it
                double dd[3];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int ck = (cview >> (2 * k)) & 3;
                    listed = listed && image_exists(xi[k], ck, g.bmin[k], g.bmax[k], p.h);
                    dd[k] = __dsub_rn(image_coord(xi[k], ck, g.bmin[k], g.bmax[k]), p.d.x[k][j]);
                }
                listed = listed && (dist_sqr_exact<D>(dd) < p.hSqr);
                if (!listed) atomicAdd(&p.d.counters[0], 1u); // one-sided pair (statistics for the parity harness)
            }
            const bool own = canon || !listed || j < p.own_begin || j >= p.own_end;
            unsigned v;
            if (own) {
                v = ((unsigned)nown << 2) | 2u | (canon ? 0u : 1u);
                ++nown;
            } else if (regular) {
                const int cg = (e >> MLH_NNL_IDX_BITS) & 31; // stencil cell of this slot (tag of pass A)
                v = (li_enc | ((unsigned)(NSTENCIL - 1 - cg) << 12)) << 2;
            } else {
                v = MLH_FMAP_GHOST_SEARCH;
            }
            p.d.fmap[at] = v;
          }
        }
        p.d.nown[i] = nown;
    }
    // longest list of this step: bounds the slot loop of the persistent face kernels (k4_flux.cu)
    {
        const unsigned am = __activemask();
        const unsigned len = __reduce_max_sync(am, (unsigned)(nreg + ng));
        if ((threadIdx.x & 31) == (__ffs(am) - 1)) atomicMax(&p.d.counters[3], len);
    }
}

} // namespace

int mlh_launch_neighbours(mlh_ctx *c) {
    Params &p = c->p;
    int n = p.own_end - p.own_begin;
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.counters + 3, 0, sizeof(unsigned), c->stream));
    mlh_prof_begin(c, KID_NEIGHBOURS);
    if (p.D == 2 && p.periodic)
        k_neighbours<2, true><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else if (p.D == 2)
        k_neighbours<2, false><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else if (p.periodic)
        k_neighbours<3, true><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    else
        k_neighbours<3, false><<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_NEIGHBOURS);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
