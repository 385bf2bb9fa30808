// K2 -- neighbour search on the sorted particles: one WARP per search-grid cell.
//
// Replaces Particles::gridNNS (/root/reference/demonstrator/src/Particles.cpp:324-365),
// Domain::getNeighborCells (Domain.cpp:83-118), Particles::createGhostParticles (:2113-2191) and the
// brute-force Particles::ghostNNS (:2237-2260).
//
// A stencil cell is a contiguous index range of the sorted arrays.  The warp that owns cell C lays the 3^D
// ranges of its stencil end to end IN THE REFERENCE'S ORDER (x outer, y, z inner; ascending index inside a
// cell) and cuts that candidate sequence into tiles of 32.  Two phases per cell:
//   tests     lane l of tile t holds candidate 32 t + l in registers (K2_U tiles at a time); the particles of C --
//             staged once in shared memory -- are tested against a tile with ONE instruction stream for 32
//             candidates: the cutoff `pow(dx,2)+pow(dy,2)[+pow(dz,2)] < h*h` without FMA contraction
//             (dist_sqr_exact), so the SETS are bit-exact.  `__ballot_sync` turns the 32 outcomes into one hit mask
//             per (particle, tile), parked in shared memory; each lane also collects, per candidate it holds, the
//             bit mask of the particles that hit it.  12 instructions per 32 tests, nothing data-dependent.
//             [The previous thread-per-particle walk spent ~11 instructions per candidate at 22 of 32 lanes:
//             614 warp instructions per particle, 0.208 ms at 61^3, profiles/r02_k_neighbours_ncu_full.txt.]
//   emission  lane l walks the masks of particle l in candidate order = the reference's list order
//             (Particles.cpp:335-359): the n-th set bit is list slot n.  All lanes emit their n-th entry in the same
//             iteration, so slot n of consecutive particles leaves as one contiguous run (the lists are slot-major).
// The masks also give the face bookkeeping without gathers or searches:
//   * ownership of the pair = lower ORIGINAL index (Particles.cpp:1841,1889); ids sit in the candidate table;
//   * r = how many particles of C BELOW this one list the same candidate j = popcount of the candidate's
//     particle mask below this lane.  j's list is ordered stencil cell by stencil cell, so this particle sits in j's
//     list at grp[mirrored cell][j] + r: k_face_index (k4_flux.cu) finds the partner's slot with one gather;
//   * grp[c][i] (first slot of stencil cell c in the list of i) falls out where the stencil cell of consecutive
//     entries changes.
// [First version of the warp-per-cell search, profiles/r2c_k_neighbours_cell_v1_*: it emitted the hits of one particle
// with one lane per hit right after every group of tiles -- 697 warp instructions per particle, 4-byte stores scattered
// over as many rows as there were hits; 0.261 ms at 61^3 and 1.43 ms at KH 1M against 0.206 / 0.52 ms before.]
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
__device__ int ghost_entries(const Params &p, const Grid &g, int i, int c, int nreg, unsigned *nown_io, bool *overflow) {
    constexpr bool PER = true;
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
constexpr int K2_U = 4;      // candidate tiles held in registers at a time
constexpr int K2_MAXCH = 16; // tiles per pass over the candidates: 512 (3D: ~280 per cell, 2D: ~140); denser cells take more passes
constexpr int K2_MAXP = 32;  // particles of the cell per batch

template <int D, bool PER>
__global__ void __launch_bounds__(32 * K2_WARPS) k_neighbours_cell(const Params p) {
    constexpr int NS = D == 3 ? 27 : 9;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ Grid s_grid;                                   // the search grid (device-built in non-periodic runs)
    __shared__ double s_x[K2_WARPS][D][K2_MAXP];              // the cell's own particles
    __shared__ unsigned s_j[K2_WARPS][32 * K2_MAXCH];         // candidate: sorted index | mirrored stencil cell << 26 | other rank's halo << 31
    __shared__ int s_id[K2_WARPS][32 * K2_MAXCH];             // candidate: original id
    __shared__ unsigned s_col[K2_WARPS][32 * K2_MAXCH];       // candidate: which particles of the batch list it (bit l)
    __shared__ unsigned s_m[K2_WARPS][K2_MAXP][K2_MAXCH + 1]; // hit mask of (particle, tile); +1: conflict-free columns
    __shared__ int s_off[K2_WARPS][NS + 1];                   // first candidate number of each stencil cell
    __shared__ int s_start[K2_WARPS][NS];                     // first sorted index of each stencil cell
    if (threadIdx.x == 0) s_grid = *p.d.grid;
    __syncthreads();
    const Grid &g = s_grid;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned ltl = (1u << lane) - 1u;
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
        const int total = __shfl_sync(FULL, inc, 31);
        __syncwarp(); // the previous cell's readers of the tables are done
        if (lane < NS) {
            s_off[w][lane] = inc - cn;
            s_start[w][lane] = cs;
        }
        if (lane == 0) s_off[w][NS] = total;
        __syncwarp();
        for (int b0 = 0; b0 < n_c; b0 += K2_MAXP) {
            const int nb = min(K2_MAXP, n_c - b0), i_base = s_c + b0;
            __syncwarp(); // the previous batch has finished with s_x / s_grp
            int idi = 0;
            if (lane < nb) {
#pragma unroll
                for (int k = 0; k < D; ++k) s_x[w][k][lane] = p.d.x[k][i_base + lane];
                idi = p.d.id[i_base + lane];
            }
            // lane l = particle l of the batch: its list so far
            unsigned cnt = 0u, nown = 0u; // entries, owned entries
            int gk = 0;                   // next stencil cell whose group start is unwritten (warp-uniform)
            __syncwarp();
            for (int q0 = 0; q0 < total; q0 += 32 * K2_MAXCH) { // passes over the candidates (one in all but pathological cells)
                const int npass = min(32 * K2_MAXCH, total - q0), ntile = (npass + 31) >> 5;
                // ---- candidate tables of this pass ----
                for (int t = 0; t < ntile; ++t) {
                    const int ql = t * 32 + lane, q = q0 + ql;
                    unsigned enc = 0xFFFFFFFFu;
                    int id = 0;
                    if (ql < npass) {
                        int k = 0; // stencil cell that holds candidate q: largest k with s_off[k] <= q
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1)
                            if (k + step < NS && s_off[w][k + step] <= q) k += step;
                        const int j = s_start[w][k] + (q - s_off[w][k]);
                        const bool offr = j < p.own_begin || j >= p.own_end;
                        enc = (unsigned)j | ((unsigned)(NS - 1 - k) << MLH_NNL_IDX_BITS) | (offr ? 0x80000000u : 0u);
                        id = p.d.id[j];
                    }
                    s_j[w][ql] = enc;
                    s_id[w][ql] = id;
                }
                __syncwarp();
                // ---- tests: every particle of the batch against the tiles, 32 candidates per instruction; only the hit
                // masks are kept (per particle and tile, and transposed per candidate) ----
                for (int c0 = 0; c0 < ntile; c0 += K2_U) {
                    double xj[K2_U][D];
                    int jj[K2_U];
                    unsigned col[K2_U];
#pragma unroll
                    for (int u = 0; u < K2_U; ++u) {
                        col[u] = 0u;
                        jj[u] = -1;
#pragma unroll
                        for (int k = 0; k < D; ++k) xj[u][k] = 1e300; // never within h of anything
                        const int ql = (c0 + u) * 32 + lane;
                        if (c0 + u < ntile && ql < npass) {
                            jj[u] = (int)(s_j[w][ql] & MLH_NNL_IDX_MASK);
#pragma unroll
                            for (int k = 0; k < D; ++k) xj[u][k] = p.d.x[k][jj[u]];
                        }
                    }
                    for (int l = 0; l < nb; ++l) {
                        const int i = i_base + l;
                        double xi[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) xi[k] = s_x[w][k][l];
#pragma unroll
                        for (int u = 0; u < K2_U; ++u) {
                            if (c0 + u >= ntile) break; // (warp-uniform)
                            double d[3];
#pragma unroll
                            for (int k = 0; k < D; ++k) d[k] = __dsub_rn(xj[u][k], xi[k]);
                            const bool hit = (dist_sqr_exact<D>(d) < p.hSqr) & (jj[u] != i); // Particles.cpp:341-347
                            const unsigned m = __ballot_sync(FULL, hit);
                            if (lane == 0) s_m[w][l][c0 + u] = m;
                            col[u] |= hit ? (1u << l) : 0u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < K2_U; ++u)
                        if (c0 + u < ntile) s_col[w][(c0 + u) * 32 + lane] = col[u];
                }
                __syncwarp();
                // ---- emission: lane l walks the hit masks of particle l; every iteration emits the next hit of every
                // particle, i.e. list slot `cnt` of consecutive particles -> one coalesced run per array ----
                const unsigned cnt_before = cnt;
                if (lane < nb) {
                    const unsigned i = (unsigned)(i_base + lane);
                    int t = 0;
                    unsigned m = s_m[w][lane][0];
                    for (;;) {
                        while (m == 0u && ++t < ntile) m = s_m[w][lane][t];
                        if (t >= ntile) break;
                        const int ql = t * 32 + (__ffs(m) - 1);
                        m &= m - 1u;
                        const unsigned enc = s_j[w][ql];
                        const unsigned slot = cnt++;
                        if (slot < max_ni) {
                            const bool canon = idi < s_id[w][ql];
                            // owner = lower ORIGINAL index (Particles.cpp:1841,1889), or the only side with a list entry
                            // (enc bit 31 = the partner is another rank's halo particle)
                            unsigned word = MLH_FMAP_SKIP; // the partner owns the pair and will fill this slot
                            if (canon | ((int)enc < 0)) {
                                // enc bits 26..31 (mirrored stencil cell, halo flag) are the word's bits 18..23
                                static_assert(MLH_K2_SC_SHIFT == MLH_NNL_IDX_BITS - 8 && MLH_K2_NOPARTNER == (1u << 23), "field layout");
                                word = (MLH_K2_OWNED | (nown << MLH_K2_RANK_SHIFT)) + (canon ? 0u : 1u) + ((enc >> 8) & 0xFC0000u) +
                                       (b0 == 0 ? ((unsigned)__popc(s_col[w][ql] & ltl) << MLH_K2_R_SHIFT) : MLH_K2_R_OVER);
                                ++nown;
                            }
                            const unsigned at = slot * ncap + i;
                            p.d.nnl[at] = (int)(enc & MLH_NNL_IDX_MASK);
                            p.d.fmap[at] = word;
                        }
                    }
                }
                // ---- group starts (grp[k][i] = entries of i that lie in stencil cells before k): for every stencil cell whose
                // first candidate falls into this pass, the hits of particle l before that candidate = (hits in the pass's
                // earlier tiles) + popc(mask of its tile & lanes below it).  k and the tile are warp-uniform: no divergence,
                // and a row of grp leaves as one run over the lanes ----
                {
                    unsigned pref = cnt_before; // entries of particle l before this pass
                    int tcur = 0;
                    while (gk < NS && s_off[w][gk] < q0 + npass) {
                        const int ql = s_off[w][gk] - q0, ch = ql >> 5;
                        for (; tcur < ch; ++tcur)
                            if (lane < nb) pref += __popc(s_m[w][lane][tcur]);
                        if (lane < nb) {
                            const unsigned gv = pref + __popc(s_m[w][lane][ch] & ((1u << (ql & 31)) - 1u));
                            p.d.grp[(unsigned)gk * ncap + (unsigned)(i_base + lane)] = (unsigned short)(gv < max_ni ? gv : max_ni);
                        }
                        ++gk;
                    }
                }
                __syncwarp(); // tables are rebuilt by the next pass
            }
            // ---- per particle: remaining group starts, list lengths, periodic images, owned-slot count ----
            if (lane < nb) {
                const int i = i_base + lane;
                if (cnt > max_ni) overflow = true;
                const int nreg = (int)(cnt < max_ni ? cnt : max_ni);
                for (int k = gk; k < NS; ++k) p.d.grp[(unsigned)k * ncap + (unsigned)i] = (unsigned short)nreg; // empty trailing cells
                p.d.noi[i] = nreg;
                int ng = 0;
                if (PER) ng = ghost_entries<D>(p, g, i, c, nreg, &nown, &overflow);
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
