// K2 -- neighbour search on the sorted particles: one WARP per search-grid cell.
//
// Replaces Particles::gridNNS (/root/reference/demonstrator/src/Particles.cpp:324-365),
// Domain::getNeighborCells (Domain.cpp:83-118), Particles::createGhostParticles (:2113-2191) and the
// brute-force Particles::ghostNNS (:2237-2260).
//
// A stencil cell is a contiguous index range of the sorted arrays.  The warp that owns cell C lays the 3^D
// ranges of its stencil end to end IN THE REFERENCE'S ORDER (x outer, y, z inner; ascending index inside a
// cell) and cuts that candidate sequence into tiles of 32: lane l of tile t holds candidate 32 t + l
// (coordinates, original id) in registers, K2_U tiles at a time.  The particles of C -- staged once in shared
// memory -- are then tested against a tile with ONE instruction stream for 32 candidates: the cutoff
// `pow(dx,2)+pow(dy,2)[+pow(dz,2)] < h*h` without FMA contraction (dist_sqr_exact), so the SETS are bit-exact;
// `__ballot_sync` gives the hit mask, and a hit lands in list slot
//     (hits of this particle so far) + popc(mask & lanes below)
// -- the list order is the reference's by construction (Particles.cpp:335-359), no sorting, no per-thread
// walk over 27 ranges.  [The previous thread-per-particle walk: 22 of 32 lanes busy, 614 warp instructions per
// particle, 0.208 ms at 61^3 -- profiles/r02_k_neighbours_ncu_full.txt.]
//
// The same ballots give the face bookkeeping for free (no group masks, no id gathers):
//   * ownership of the pair = lower ORIGINAL index (Particles.cpp:1841,1889), ids ride with the candidates;
//   * the rank of the slot among the particle's owned slots = running count + popc(owned mask & lanes below);
//   * r = how many particles of C BELOW this one list the same candidate j = a per-lane hit counter, because
//     the particles of C are visited in ascending order.  j's list is ordered stencil cell by stencil cell,
//     so this particle sits in j's list at grp[mirrored cell][j] + r: k_face_index (k4_flux.cu) finds the
//     partner's slot with one gather instead of a search;
//   * the hits of a particle in the tiles at hand are compacted into a shared-memory queue (queue position = list
//     slot), then written out one lane per hit; where the stencil cell changes between consecutive hits, that lane
//     also records grp[c][i], the first slot of stencil cell c in the list of i.
//
// Periodic images are not materialised as ghost particles: a stencil cell that wraps around the box yields
// candidates whose image position is computed on the fly with the reference's formulas (image_coord) and whose
// existence test is the reference's threshold (image_exists).  Because cellSize >= h (Domain.cpp:10-22) every
// ghost within h of a particle lives in a wrapped stencil cell, so the result equals the brute-force search
// over all ghosts.  Only the cells on the rim of the box have wrapped stencil cells; their few image entries
// are appended per particle (one lane each) after the regular ones, ordered by parent original index
// (== ascending ghost index of ghostNNS).
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

// offsets of stencil cell number k in the reference's order (Domain.cpp:83-118: x outer, y, z inner)
template <int D>
__device__ __forceinline__ void stencil_offset(int k, int *off) {
    if (D == 3) {
        off[0] = k / 9 - 1;
        off[1] = (k / 3) % 3 - 1;
        off[2] = k % 3 - 1;
    } else {
        off[0] = k / 3 - 1;
        off[1] = k % 3 - 1;
        off[2] = 0;
    }
}

// ---- periodic images of one particle (one lane): Particles.cpp:2113-2191 + :2237-2260, appended after the `nreg`
// regular entries; then the ownership of those slots.  Returns the number of image entries; *nown is advanced. ----
template <int D>
__device__ int ghost_entries(const Params &p, int i, int c, int nreg, unsigned *nown_io, bool *overflow) {
    constexpr bool PER = true;
    const Grid &g = *p.d.grid;
    double xi[3];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = p.d.x[k][i];
    int ci[3];
    ci[0] = c % g.lcells[0];
    ci[1] = (c / g.lcells[0]) % g.lcells[1];
    ci[2] = D == 3 ? c / (g.lcells[0] * g.lcells[1]) : 0;
    const int idi = p.d.id[i];
    constexpr int NS = D == 3 ? 27 : 9;
    int cnt = nreg;
    for (int sk = 0; sk < NS; ++sk) {
        int off[3];
        stencil_offset<D>(sk, off);
        StencilCell<D> sc = stencil_cell<D, PER>(g, ci, off);
        if (sc.cell < 0 || sc.code == 0) continue;
        const int rcode = reverse_code(sc.code);
        const int s = p.d.cell_start[sc.cell], e = p.d.cell_start[sc.cell + 1];
        for (int j = s; j < e; ++j) {
            double d[3];
            bool exists = true;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const int ck = (sc.code >> (2 * k)) & 3;
                const double xj = p.d.x[k][j];
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
                    const int ck = (rcode >> (2 * k)) & 3;
                    ex2 = ex2 && image_exists(xi[k], ck, g.bmin[k], g.bmax[k], p.h);
                    d2[k] = __dsub_rn(image_coord(xi[k], ck, g.bmin[k], g.bmax[k]), p.d.x[k][j]);
                }
                hit = ex2 && (dist_sqr_exact<D>(d2) < p.hSqr);
            }
            if (hit) {
                if (cnt < p.max_ni)
                    p.d.nnl[(size_t)cnt * p.ncap + i] = j | (sc.code << MLH_NNL_IDX_BITS);
                else
                    *overflow = true;
                ++cnt;
            }
        }
    }
    const int ntot = cnt < p.max_ni ? cnt : p.max_ni;
    // order the image entries by parent original index (ghostNNS scans the ghosts in creation order, which is
    // ascending parent index, Particles.cpp:2116,2241)
    for (int a = nreg + 1; a < ntot; ++a) {
        const int ea = p.d.nnl[(size_t)a * p.ncap + i];
        const int ka = p.d.id[ea & MLH_NNL_IDX_MASK];
        int b = a - 1;
        while (b >= nreg) {
            const int eb = p.d.nnl[(size_t)b * p.ncap + i];
            if (p.d.id[eb & MLH_NNL_IDX_MASK] <= ka) break;
            p.d.nnl[(size_t)(b + 1) * p.ncap + i] = eb;
            --b;
        }
        p.d.nnl[(size_t)(b + 1) * p.ncap + i] = ea;
    }
    // ownership of the image slots: lower original index, or the only side that lists the pair (quirk Q9), or the
    // partner lives on another rank
    unsigned nown = *nown_io;
    for (int s = nreg; s < ntot; ++s) {
        const size_t at = (size_t)s * p.ncap + i;
        const int e = p.d.nnl[at];
        const int j = e & MLH_NNL_IDX_MASK;
        const bool canon = !(p.d.id[j] < idi);
        bool listed = true; // does j list this pair too?  (the test above from j's side)
        if (!p.symmetric_seam) {
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
        const bool offrank = j < p.own_begin || j >= p.own_end;
        unsigned word = MLH_FMAP_SKIP;
        if (canon || !listed || offrank) {
            word = MLH_K2_OWNED | (canon ? 0u : 1u) | (nown << MLH_K2_RANK_SHIFT) | MLH_K2_GHOST |
                   ((!listed || offrank) ? MLH_K2_NOPARTNER : 0u);
            ++nown;
        }
        p.d.fmap[at] = word;
    }
    *nown_io = nown;
    return ntot - nreg;
}

constexpr int K2_WARPS = 4;  // warps (= cells in flight) per block
constexpr int K2_U = 5;      // candidate tiles held in registers at a time (3D: ~9 tiles per cell, 2D: 4-5)
constexpr int K2_MAXCH = 32; // tiles per cell: 1024 candidates; a denser stencil overflows every list anyway
constexpr int K2_MAXP = 32;  // particles of the cell per pass

// per-particle progress word carried by lane l for particle l of the pass
__device__ __forceinline__ unsigned k2_pack(unsigned cnt, unsigned nown, unsigned gk) { return cnt | (nown << 11) | (gk << 22); }

template <int D, bool PER>
__global__ void __launch_bounds__(32 * K2_WARPS) k_neighbours_cell(const Params p) {
    constexpr int NS = D == 3 ? 27 : 9;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned KEEP = (31u << MLH_K2_R_SHIFT) | MLH_K2_R_OVER | (31u << MLH_K2_SC_SHIFT) | MLH_K2_NOPARTNER;
    __shared__ double s_x[K2_WARPS][D][K2_MAXP];     // the cell's own particles
    __shared__ int s_qj[K2_WARPS][32 * K2_U];        // hits of the current particle in the current tiles: sorted index ..
    __shared__ unsigned s_qi[K2_WARPS][32 * K2_U];   // .. and pair info (MLH_K2_* bits, bit 0 = the partner has the lower id)
    __shared__ int s_off[K2_WARPS][NS + 1];          // first candidate number of each stencil cell
    __shared__ int s_start[K2_WARPS][NS];            // first sorted index of each stencil cell
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const Grid &g = *p.d.grid;
    const int nwarps = gridDim.x * K2_WARPS;
    const unsigned max_ni = (unsigned)p.max_ni, ncap = (unsigned)p.ncap;
    // cells that hold owned particles: all of them, or (slab decomposition) the layers between the two halo layers
    int cell_begin = 0, cell_end = g.ncells;
    if (g.sliced) {
        const int layer_cells = g.ncells / g.lcells[g.slab_dim];
        cell_begin = layer_cells;
        cell_end = g.ncells - layer_cells;
    }
    bool overflow = false;
    unsigned maxlen = 0;
    for (int c = cell_begin + blockIdx.x * K2_WARPS + w; c < cell_end; c += nwarps) {
        const int s_c = p.d.cell_start[c], n_c = p.d.cell_start[c + 1] - s_c;
        if (n_c <= 0) continue; // (warp-uniform)
        // ---- the candidate sequence: regular stencil cells end to end in the reference's order ----
        int cs = 0, cn = 0;
        if (lane < NS) {
            int ci[3], off[3];
            ci[0] = c % g.lcells[0];
            ci[1] = (c / g.lcells[0]) % g.lcells[1];
            ci[2] = D == 3 ? c / (g.lcells[0] * g.lcells[1]) : 0;
            stencil_offset<D>(lane, off);
            const StencilCell<D> sc = stencil_cell<D, PER>(g, ci, off);
            if (sc.cell >= 0 && sc.code == 0) {
                cs = p.d.cell_start[sc.cell];
                cn = p.d.cell_start[sc.cell + 1] - cs;
            }
        }
        int inc = cn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        int total = __shfl_sync(FULL, inc, 31);
        __syncwarp(); // the previous cell's readers of s_off / s_start are done
        if (lane < NS) {
            s_off[w][lane] = inc - cn;
            s_start[w][lane] = cs;
        }
        if (lane == 0) s_off[w][NS] = total;
        int nch = (total + 31) >> 5;
        if (nch > K2_MAXCH) { // more candidates than any list can hold (e.g. a NaN state collapsed into one cell)
            overflow = true;
            nch = K2_MAXCH;
            total = 32 * K2_MAXCH;
        }
        __syncwarp();
        for (int b0 = 0; b0 < n_c; b0 += K2_MAXP) {
            const int nb = min(K2_MAXP, n_c - b0), i_base = s_c + b0;
            __syncwarp(); // the previous pass has finished with s_x
            int idi_l = 0;
            if (lane < nb) {
#pragma unroll
                for (int k = 0; k < D; ++k) s_x[w][k][lane] = p.d.x[k][i_base + lane];
                idi_l = p.d.id[i_base + lane];
            }
            // lane l: hits so far | owned slots so far << 11 | next stencil cell whose group start is unwritten << 22
            unsigned cw_l = 0u;
            // r is only counted over the first 32 particles of a cell; beyond that k_face_index searches
            const unsigned rinc = b0 == 0 ? (1u << MLH_K2_R_SHIFT) : 0u;
            __syncwarp();
            for (int c0 = 0; c0 < nch; c0 += K2_U) {
                // ---- K2_U tiles of candidates into registers ----
                double xj[K2_U][D];
                int jj[K2_U], idj[K2_U];
                unsigned info[K2_U]; // mirrored stencil cell | NOPARTNER (other rank's halo) | r (or R_OVER)
#pragma unroll
                for (int u = 0; u < K2_U; ++u) {
                    const int q = (c0 + u) * 32 + lane;
                    jj[u] = -1;
                    idj[u] = 0;
                    info[u] = 0u;
#pragma unroll
                    for (int k = 0; k < D; ++k) xj[u][k] = 1e300; // never within h of anything
                    if (c0 + u < nch && q < total) {
                        int k = 0; // stencil cell that holds candidate q: largest k with s_off[k] <= q
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1)
                            if (k + step < NS && s_off[w][k + step] <= q) k += step;
                        const int j = s_start[w][k] + (q - s_off[w][k]);
                        jj[u] = j;
                        idj[u] = p.d.id[j];
#pragma unroll
                        for (int k2 = 0; k2 < D; ++k2) xj[u][k2] = p.d.x[k2][j];
                        const bool offr = j < p.own_begin || j >= p.own_end;
                        info[u] = ((unsigned)(NS - 1 - k) << MLH_K2_SC_SHIFT) | (offr ? MLH_K2_NOPARTNER : 0u) |
                                  (b0 == 0 ? 0u : MLH_K2_R_OVER);
                    }
                }
                // ---- every particle of this pass against those tiles, ascending (so the per-lane hit counters r count
                // the particles of the cell BELOW the current one that list the lane's candidate) ----
                for (int l = 0; l < nb; ++l) {
                    const unsigned i = (unsigned)(i_base + l);
                    double xi[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) xi[k] = s_x[w][k][l];
                    const int idi = __shfl_sync(FULL, idi_l, l);
                    const unsigned cw = __shfl_sync(FULL, cw_l, l);
                    unsigned cnt = cw & 0x7ffu, nown = (cw >> 11) & 0x7ffu, gk = cw >> 22;
                    unsigned qn = 0u; // hits of this particle in these tiles
#pragma unroll
                    for (int u = 0; u < K2_U; ++u) {
                        if (c0 + u >= nch) break; // (warp-uniform)
                        double d[3];
#pragma unroll
                        for (int k = 0; k < D; ++k) d[k] = __dsub_rn(xj[u][k], xi[k]);
                        const bool hit = (dist_sqr_exact<D>(d) < p.hSqr) & (jj[u] != (int)i); // Particles.cpp:341-347
                        const unsigned m = __ballot_sync(FULL, hit);
                        const unsigned pos = qn + __popc(m & lt);
                        if (hit) { // list order = candidate order: the queue position IS the list slot (minus cnt)
                            s_qj[w][pos] = jj[u];
                            s_qi[w][pos] = info[u] | (idi < idj[u] ? 0u : 1u);
                            info[u] += rinc;
                        }
                        qn += __popc(m);
                    }
                    __syncwarp();
                    // ---- emit: one lane per hit ----
                    for (unsigned base = 0; base < qn; base += 32) {
                        const unsigned e = base + lane, slot = cnt + e;
                        const bool valid = e < qn;
                        const int j = valid ? s_qj[w][e] : 0;
                        const unsigned inf = valid ? s_qi[w][e] : 0u;
                        const bool fits = valid & (slot < max_ni);
                        // owner = lower ORIGINAL index (Particles.cpp:1841,1889), or the only side with a list entry
                        const bool own = fits & (!(inf & 1u) | ((inf & MLH_K2_NOPARTNER) != 0u));
                        const unsigned mo = __ballot_sync(FULL, own);
                        if (fits) {
                            const unsigned at = slot * ncap + i;
                            p.d.nnl[at] = j;
                            p.d.fmap[at] = own ? (MLH_K2_OWNED | (inf & 1u) | ((nown + __popc(mo & lt)) << MLH_K2_RANK_SHIFT) | (inf & KEEP))
                                               : MLH_FMAP_SKIP; // the partner owns the pair and will fill this slot
                        }
                        if (valid & !fits) overflow = true;
                        // group starts: the stencil cells (pk, kk] begin at this slot
                        const unsigned kk = valid ? (unsigned)(NS - 1) - ((inf >> MLH_K2_SC_SHIFT) & 31u) : (unsigned)NS;
                        unsigned pk = __shfl_up_sync(FULL, kk, 1);
                        if (lane == 0) pk = gk - 1u; // (gk = 0: wraps to ~0, pk + 1 = 0)
                        if (valid) {
                            const unsigned short gv = (unsigned short)(slot < max_ni ? slot : max_ni);
                            for (unsigned k = pk + 1u; k <= kk; ++k) p.d.grp[k * ncap + i] = gv;
                        }
                        nown += __popc(mo);
                        const unsigned nvalid = min(32u, qn - base);
                        gk = __shfl_sync(FULL, kk, nvalid - 1u) + 1u;
                    }
                    cnt += qn;
                    if (lane == l) cw_l = k2_pack(cnt, nown, gk);
                    __syncwarp(); // the queue is reused by the next particle
                }
            }
            // ---- per particle: remaining group starts, list lengths, periodic images, owned-slot count ----
            if (lane < nb) {
                const int i = i_base + lane;
                const unsigned cnt_l = cw_l & 0x7ffu;
                unsigned nown = (cw_l >> 11) & 0x7ffu;
                if (cnt_l > max_ni) overflow = true;
                const int nreg = (int)(cnt_l < max_ni ? cnt_l : max_ni);
                for (unsigned k = cw_l >> 22; k < (unsigned)NS; ++k) p.d.grp[k * ncap + (unsigned)i] = (unsigned short)nreg;
                p.d.noi[i] = nreg;
                int ng = 0;
                if (PER) ng = ghost_entries<D>(p, i, c, nreg, &nown, &overflow);
                p.d.noig[i] = ng;
                p.d.nown[i] = (int)nown;
                const unsigned len = (unsigned)(nreg + ng);
                maxlen = len > maxlen ? len : maxlen;
            }
        }
    }
    if (overflow) atomicOr(p.d.flags, MLH_F_MAX_INTERACTIONS);
    // longest list of this step (diagnostics)
    maxlen = __reduce_max_sync(FULL, maxlen);
    if (lane == 0 && maxlen) atomicMax(&p.d.counters[3], maxlen);
}

} // namespace

int mlh_launch_neighbours(mlh_ctx *c) {
    Params &p = c->p;
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.counters + 3, 0, sizeof(unsigned), c->stream));
    if (p.own_end <= p.own_begin) return MLH_OK;
    // one warp per cell, handed out grid-stride; the cell count is the host's when it knows the grid, else the capacity
    // of the cell arrays (device-built grid: the surplus warps find no cell)
    const int ncell = c->grid_host_current ? p.grid.ncells : c->max_cells;
    int blocks = mlh_blocks(ncell, K2_WARPS);
    const int cap = c->num_sms * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    mlh_prof_begin(c, KID_NEIGHBOURS);
    if (p.D == 2 && p.periodic)
        k_neighbours_cell<2, true><<<blocks, 32 * K2_WARPS, 0, c->stream>>>(p);
    else if (p.D == 2)
        k_neighbours_cell<2, false><<<blocks, 32 * K2_WARPS, 0, c->stream>>>(p);
    else if (p.periodic)
        k_neighbours_cell<3, true><<<blocks, 32 * K2_WARPS, 0, c->stream>>>(p);
    else
        k_neighbours_cell<3, false><<<blocks, 32 * K2_WARPS, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_NEIGHBOURS);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
