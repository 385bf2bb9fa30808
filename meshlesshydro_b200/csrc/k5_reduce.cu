// K0 / K5 helpers -- bounding box (non-periodic grid rebuild), conservation sums, un-permutation.
//
//   Particles::getDomainLimits   /root/reference/demonstrator/src/Particles.cpp:228-267  (quirks Q2, Q8)
//   Particles::sumVolume/sumMass/sumEnergy/sumMomentumX/Y/Z   :2830-2886
// All HBM-bound streaming reductions (warp shuffle -> block -> one atomic per block).
#include "mlh_internal.cuh"
#include <cfloat>

namespace {

__global__ void k_bbox_init(const Params p) {
    int k = threadIdx.x;
    if (k < 3) {
        unsigned long long *keys = (unsigned long long *)p.d.bbox;
        keys[k] = dbl_key(DBL_MAX);     // running minimum starts at numeric_limits::max()  (:230)
        keys[3 + k] = dbl_key(DBL_MIN); // running maximum starts at numeric_limits::min()  (:231, quirk Q2)
        p.d.bbox[6 + k] = -DBL_MAX;     // coordinate of original particle 0 (set by whoever owns it)
        p.d.bbox[9 + k] = -DBL_MAX;     // coordinate of original particle 1 (multi-GPU form of quirk Q8, mlh_capi.cu)
    }
}

// min over all particles, max over ORIGINAL index >= 1: x[0] always lowers the running minimum first and is
// therefore never tested against the maximum (`if (x<min) .. else if (x>max)`, :240-244, quirk Q8)
template <int D>
__global__ void __launch_bounds__(256) k_bbox(const Params p) {
    double mn[D], mx[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        mn[k] = DBL_MAX;
        mx[k] = DBL_MIN;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.ncur; i += gridDim.x * blockDim.x) {
        const int id = p.d.cid[i];
        const bool first = id == 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double x = p.d.cx[k][i];
            mn[k] = x < mn[k] ? x : mn[k];
            if (!first) mx[k] = x > mx[k] ? x : mx[k];
            else p.d.bbox[6 + k] = x;
            if (id == 1) p.d.bbox[9 + k] = x;
        }
    }
    mlh_bbox_block_reduce<D>(p, mn, mx);
}

// Exact Q8 semantics for the one case the parallel reduction cannot see: x[0] is the strict maximum
// along an axis.  Then the sequential loop of the reference decides which later elements were ever
// compared against the running maximum; replay it in ORIGINAL order (one thread per axis, cold path).
__global__ void k_inverse_ids(const Params p, int *inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.ncur) inv[p.d.cid[i]] = i;
}
template <int D>
__global__ void k_bbox_replay(const Params p, const int *inv) {
    int k = threadIdx.x;
    if (k >= D) return;
    unsigned long long *keys = (unsigned long long *)p.d.bbox;
    if (!(p.d.bbox[6 + k] > key_dbl(keys[3 + k]))) return;
    double mn = DBL_MAX, mx = DBL_MIN;
    for (int id = 0; id < p.ncur; ++id) {
        double x = p.d.cx[k][inv[id]];
        if (x < mn)
            mn = x;
        else if (x > mx)
            mx = x;
    }
    keys[k] = dbl_key(mn);
    keys[3 + k] = dbl_key(mx);
}

// Domain::createGrid (Domain.cpp:9-54) on the device, fed by the bounding box the update kernel of the previous step
// reduced (getDomainLimits, Particles.cpp:228-267): non-periodic runs rebuild the search grid every step
// (MeshlessScheme.cpp:41-51) and no longer stop the stream for it.  One block: if original particle 0 is the strict
// maximum along an axis (quirk Q8) the block first builds the id -> index table and that axis is replayed in original
// order, exactly as the reference's sequential loop.  The arithmetic is the host's (make_grid, mlh_capi.cu): IEEE
// subtract, divide, floor.  A grid that does not fit the allocated cell arrays raises MLH_F_OUT_OF_GRID and leaves the
// previous grid in place (the host grows the arrays from its one-step-late mirror long before that).
template <int D>
__global__ void __launch_bounds__(1024) k_make_grid(const Params p, int max_cells) {
    __shared__ int need;
    unsigned long long *keys = (unsigned long long *)p.d.bbox;
    if (threadIdx.x == 0) {
        // the reduction (min over all, max over original index >= 1) differs from the sequential loop only if the largest
        // coordinate among the indices >= 1 belongs to particle 1 AND lies below particle 0 (see mlh_build_grid)
        int nd = 0;
        for (int k = 0; k < D; ++k) {
            const double m1 = key_dbl(keys[3 + k]);
            nd |= (p.d.bbox[6 + k] > m1 && p.d.bbox[9 + k] >= m1) ? 1 : 0;
        }
        need = nd;
    }
    __syncthreads();
    if (need) {
        int *inv = p.d.perm; // free before the sort
        for (int i = threadIdx.x; i < p.ncur; i += blockDim.x) inv[p.d.cid[i]] = i;
        __syncthreads();
        if (threadIdx.x < D) {
            const int k = threadIdx.x;
            const double m1 = key_dbl(keys[3 + k]);
            if (p.d.bbox[6 + k] > m1 && p.d.bbox[9 + k] >= m1) {
                double mn = DBL_MAX, mx = DBL_MIN;
                for (int id = 0; id < p.ncur; ++id) {
                    const double x = p.d.cx[k][inv[id]];
                    if (x < mn)
                        mn = x;
                    else if (x > mx)
                        mx = x;
                }
                keys[k] = dbl_key(mn);
                keys[3 + k] = dbl_key(mx);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Grid g = *p.d.grid;
        long long nc = 1;
        bool ok = true;
        for (int k = 0; k < D; ++k) {
            const double lo = key_dbl(keys[k]), hi = key_dbl(keys[3 + k]);
            const double cells = floor(__ddiv_rn(__dsub_rn(hi, lo), p.h));
            if (!(cells >= 1.) || cells > 2.0e9) {
                ok = false;
                break;
            }
            g.bmin[k] = lo;
            g.bmax[k] = hi;
            g.cells[k] = g.lcells[k] = (int)cells;
            g.cell_size[k] = __ddiv_rn(__dsub_rn(hi, lo), (double)g.cells[k]);
            nc *= g.cells[k];
            if (nc + 1 > (long long)max_cells) {
                ok = false;
                break;
            }
        }
        if (ok) {
            g.ncells = (int)nc;
            *p.d.grid = g;
        } else {
            atomicOr(p.d.flags, MLH_F_OUT_OF_GRID);
        }
    }
}

// sums over the CUR set (state) -- V uses omega of the SRT set (same order)
template <int D>
__global__ void __launch_bounds__(256) k_sums(const Params p, int n, const double *m, const double *u,
                                              const double *v0, const double *v1, const double *v2, const double *omega) {
    double s[6] = {0., 0., 0., 0., 0., 0.};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double mi = m[i], a = v0[i], b = v1[i], c = D == 3 ? v2[i] : 0.;
        if (omega) s[0] += 1. / omega[i];
        s[1] += mi;
        s[2] += D == 3 ? mi * (u[i] + .5 * (a * a + b * b + c * c)) : mi * (u[i] + .5 * (a * a + b * b));
        s[3] += mi * a;
        s[4] += mi * b;
        if (D == 3) s[5] += mi * c;
    }
    __shared__ double sh[6][8];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = s[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
        atomicAdd(&p.d.sums[threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(256) k_unpermute_f64(const double *src, const int *ids, double *dst, int n, int comps, int stride, int id_base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[(size_t)(ids[i] - id_base) * stride + comps] = src[i];
}
__global__ void __launch_bounds__(256) k_unpermute_i32(const int *src, const int *ids, int *dst, int n, int id_base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[ids[i] - id_base] = src[i];
}

} // namespace

int mlh_launch_bbox(mlh_ctx *c) {
    Params &p = c->p;
    mlh_prof_begin(c, KID_BBOX);
    k_bbox_init<<<1, 32, 0, c->stream>>>(p);
    int blocks = mlh_blocks(p.ncur, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (p.D == 2)
        k_bbox<2><<<blocks, 256, 0, c->stream>>>(p);
    else
        k_bbox<3><<<blocks, 256, 0, c->stream>>>(p);
    mlh_prof_end(c, KID_BBOX);
    c->launches += 1;
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}

// single GPU only (ids are 0..N-1 and all resident); p.d.perm is free before the sort and serves as scratch
int mlh_launch_bbox_q8_replay(mlh_ctx *c) {
    Params &p = c->p;
    k_inverse_ids<<<mlh_blocks(p.ncur, 256), 256, 0, c->stream>>>(p, p.d.perm);
    if (p.D == 2)
        k_bbox_replay<2><<<1, 32, 0, c->stream>>>(p, p.d.perm);
    else
        k_bbox_replay<3><<<1, 32, 0, c->stream>>>(p, p.d.perm);
    c->launches += 2;
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}

int mlh_launch_make_grid(mlh_ctx *c) {
    Params &p = c->p;
    if (p.D == 2)
        k_make_grid<2><<<1, 1024, 0, c->stream>>>(p, c->max_cells);
    else
        k_make_grid<3><<<1, 1024, 0, c->stream>>>(p, c->max_cells);
    c->launches += 1;
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}

// sums over the current state; which set holds it depends on the phase (see mlh_capi.cu)
int mlh_launch_sums(mlh_ctx *c) {
    Params &p = c->p;
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.sums, 0, 6 * sizeof(double), c->stream));
    const bool in_srt = c->phase >= 1;
    int n = in_srt ? (p.own_end - p.own_begin) : p.ncur;
    int off = in_srt ? p.own_begin : 0;
    const double *m = in_srt ? p.d.m + off : p.d.cm;
    const double *u = in_srt ? p.d.u + off : p.d.cu;
    const double *v0 = in_srt ? p.d.v[0] + off : p.d.cv[0];
    const double *v1 = in_srt ? p.d.v[1] + off : p.d.cv[1];
    const double *v2 = p.D == 3 ? (in_srt ? p.d.v[2] + off : p.d.cv[2]) : nullptr;
    // omega is defined once density has run in this or a previous step; it lives in SRT order, which
    // is also the order of the CUR set produced from it
    const double *omega = p.d.omega + p.own_begin;
    int blocks = mlh_blocks(n, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    mlh_prof_begin(c, KID_SUMS);
    if (p.D == 2)
        k_sums<2><<<blocks, 256, 0, c->stream>>>(p, n, m, u, v0, v1, v2, omega);
    else
        k_sums<3><<<blocks, 256, 0, c->stream>>>(p, n, m, u, v0, v1, v2, omega);
    mlh_prof_end(c, KID_SUMS);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}

int mlh_launch_unpermute_f64(mlh_ctx *c, const double *src, const int *ids, double *dst, int n, int comps, int stride) {
    mlh_prof_begin(c, KID_UNPERMUTE);
    k_unpermute_f64<<<mlh_blocks(n, 256), 256, 0, c->stream>>>(src, ids, dst, n, comps, stride, 0);
    mlh_prof_end(c, KID_UNPERMUTE);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
int mlh_launch_unpermute_i32(mlh_ctx *c, const int *src, const int *ids, int *dst, int n) {
    mlh_prof_begin(c, KID_UNPERMUTE);
    k_unpermute_i32<<<mlh_blocks(n, 256), 256, 0, c->stream>>>(src, ids, dst, n, 0);
    mlh_prof_end(c, KID_UNPERMUTE);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    return MLH_OK;
}
