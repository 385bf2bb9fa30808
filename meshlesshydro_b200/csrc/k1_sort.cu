// K1 -- search-grid cell keys, counting sort and reorder of the SoA state.
//
// Replaces Particles::assignParticlesAndCells (/root/reference/demonstrator/src/Particles.cpp:270-322)
// and the per-cell std::vector<int> of Domain::Cell (Domain.h:17-41): instead of pushing particle
// indices into per-cell vectors, the particles themselves are reordered so that a cell is a
// contiguous index range [cell_start[c], cell_start[c+1]) of every SoA array.  Inside a cell the
// order is ascending ORIGINAL index, i.e. exactly the push_back order of the reference
// (Particles.cpp:319), which keeps the neighbour-list order -- and with it every per-particle
// sum order -- identical to the reference's.
//
// HBM-bound: reads pos(D)+state(2D+2 incl. pos)+id, writes the same (14/19 doubles per particle).
#include "mlh_internal.cuh"

namespace {

// cell key: floor((x-min)/cellSize) with the `== cells -> -1` clamp, Particles.cpp:279-302 (bit-exact:
// IEEE subtract, divide, floor -- no contraction possible).
template <int D>
__global__ void __launch_bounds__(256) k_cell_key(const Params p) {
    __shared__ Grid s_grid; // the grid lives in device memory (device-built in non-periodic runs): one copy per block
    if (threadIdx.x == 0) s_grid = *p.d.grid;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.ncur) return;
    const Grid &g = s_grid;
    int idx[3] = {0, 0, 0};
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double x = p.d.cx[k][i];
        double f = floor(__ddiv_rn(__dsub_rn(x, g.bmin[k]), g.cell_size[k]));
        int fi = (f >= 2147483647.0 || f <= -2147483648.0 || f != f) ? -1 : (int)f;
        if (fi == g.cells[k]) fi -= 1;
        if (fi < 0 || fi >= g.cells[k]) {
            bad = true; // the reference would index out of bounds here
            fi = fi < 0 ? 0 : g.cells[k] - 1;
        }
        if (g.sliced && k == g.slab_dim) {
            // local layer of this rank's grid (owned layers + one halo layer each side)
            int ll = fi - g.layer0;
            if (p.periodic) {
                if (ll < 0) ll += g.cells[k];
                if (ll >= g.cells[k]) ll -= g.cells[k];
            }
            if (ll < 0 || ll >= g.lcells[k]) {
                atomicOr(p.d.flags, MLH_F_MIGRATION);
                ll = ll < 0 ? 0 : g.lcells[k] - 1;
            }
            fi = ll;
        }
        idx[k] = fi;
    }
    if (bad) atomicOr(p.d.flags, MLH_F_OUT_OF_GRID);
    int key = idx[0] + g.lcells[0] * (idx[1] + g.lcells[1] * idx[2]);
    p.d.ckey[i] = key;
    p.d.crank[i] = atomicAdd(&p.d.cell_count[key], 1);
}

// ---- exclusive scan of cell_count -> cell_start (three small kernels, 1024 items per block) ----
constexpr int SCAN_T = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_sums[SCAN_T / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_T / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_T / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    int base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[SCAN_T / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const int *in, int *out, int *tile_sums, int n) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int *tile_sums, int ntiles) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += SCAN_T) {
        int i = base + threadIdx.x;
        int v = i < ntiles ? tile_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry; // grand total
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int *out, const int *tile_sums, int n, int total_slot, int total_value) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int add = tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[total_slot] = total_value >= 0 ? total_value : tile_sums[gridDim.x];
}

__global__ void __launch_bounds__(256) k_scatter_perm(const Params p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.ncur) return;
    p.d.perm[p.d.cell_start[p.d.ckey[i]] + p.d.crank[i]] = i;
}

// Reorder into cell order.  perm (k_scatter_perm) lists the members of a cell in arrival order of the atomics; the
// canonical order inside a cell is ascending ORIGINAL index (Cell::prtcls push_back order, Particles.cpp:319).  Cells
// hold O(10) particles because cellSize >= h (Domain.cpp:10-22), so every particle simply counts the members of its
// cell with a smaller id -- independent loads, one thread per particle -- and writes itself to cell_start + rank.
// [A separate one-thread-per-cell insertion sort took 37 us at 61^3 for 22k threads of dependent loads: profiles/r01o.]
template <int D>
__global__ void __launch_bounds__(256) k_gather_sorted(const Params p) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= p.ncur) return;
    const int src = p.d.perm[a];
    const int id = p.d.cid[src];
    const int key = p.d.ckey[src];
    const int s = p.d.cell_start[key], e = p.d.cell_start[key + 1];
    int dst = a;
    if (e - s > 8 * p.max_ni + 64) {
        // far more particles in one cell (edge >= h) than any neighbour list can hold: the search would overflow
        // MAX_NUM_INTERACTIONS anyway (reference: exit(1)); typical cause is a NaN state collapsing into one cell.
        // Raise the flag instead of spending O(n^2) here (arrival order is kept).
        atomicOr(p.d.flags, MLH_F_MAX_INTERACTIONS);
    } else {
        int rank = 0;
        for (int b = s; b < e; ++b) rank += p.d.cid[p.d.perm[b]] < id ? 1 : 0;
        dst = s + rank;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        p.d.x[k][dst] = p.d.cx[k][src];
        p.d.v[k][dst] = p.d.cv[k][src];
    }
    p.d.m[dst] = p.d.cm[src];
    p.d.u[dst] = p.d.cu[src];
    p.d.id[dst] = id;
    p.d.cell[dst] = key;
}

} // namespace

int mlh_exclusive_scan(mlh_ctx *c, const int *in, int *out, int *tmp, int n) {
    cudaStream_t st = c->stream;
    const int ntiles = mlh_blocks(n, SCAN_TILE);
    k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(in, out, tmp, n);
    k_scan_sums<<<1, SCAN_T, 0, st>>>(tmp, ntiles);       // afterwards tmp[ntiles] = grand total
    k_scan_add<<<ntiles, SCAN_T, 0, st>>>(out, tmp, n, n, -1);
    c->launches += 3;
    return MLH_OK;
}

int mlh_launch_sort(mlh_ctx *c) {
    Params &p = c->p;
    // the scan covers the whole allocated cell array: counts beyond the grid's last cell are zero, so every entry from
    // cell_start[ncells] on equals n -- the kernels never need the host to know ncells (device-built grid)
    const int n = p.ncur, nc = c->max_cells - 1;
    cudaStream_t st = c->stream;
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.cell_count, 0, sizeof(int) * (size_t)(nc + 1), st));
    mlh_prof_begin(c, KID_KEY);
    if (p.D == 2)
        k_cell_key<2><<<mlh_blocks(n, 256), 256, 0, st>>>(p);
    else
        k_cell_key<3><<<mlh_blocks(n, 256), 256, 0, st>>>(p);
    mlh_prof_end(c, KID_KEY);
    int ntiles = mlh_blocks(nc, SCAN_TILE);
    mlh_prof_begin(c, KID_SCAN);
    k_scan_tiles<<<ntiles, SCAN_T, 0, st>>>(p.d.cell_count, p.d.cell_start, p.d.scan_tmp, nc);
    k_scan_sums<<<1, SCAN_T, 0, st>>>(p.d.scan_tmp, ntiles);
    k_scan_add<<<ntiles, SCAN_T, 0, st>>>(p.d.cell_start, p.d.scan_tmp, nc, nc, n);
    mlh_prof_end(c, KID_SCAN);
    c->launches += 2; // prof_end counts one launch per bracket; scan has three
    mlh_prof_begin(c, KID_SCATTER);
    k_scatter_perm<<<mlh_blocks(n, 256), 256, 0, st>>>(p);
    mlh_prof_end(c, KID_SCATTER);
    mlh_prof_begin(c, KID_GATHER);
    if (p.D == 2)
        k_gather_sorted<2><<<mlh_blocks(n, 256), 256, 0, st>>>(p);
    else
        k_gather_sorted<3><<<mlh_blocks(n, 256), 256, 0, st>>>(p);
    mlh_prof_end(c, KID_GATHER);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    p.n = n;
    return MLH_OK;
}
