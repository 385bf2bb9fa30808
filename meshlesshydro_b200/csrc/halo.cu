// Multi-GPU slab decomposition: NCCL halo exchange, particle migration, dt / bbox / sum reductions.
//
// The reference is serial; the only "exchange" it has is the copy of parent state into the periodic
// ghost particles at fixed points of the step (/root/reference/demonstrator/src/MeshlessScheme.cpp:60,109,122,129
// -> Particles::createGhostParticles :2113-2191, updateGhostState :2193-2206, updateGhostGradients
// :2208-2222).  The slab decomposition reuses exactly those points:
//
//   exchange 1 (start of step, = createGhostParticles): raw state (pos, v, m, u, id) of the two boundary cell
//              layers goes to the slab neighbours, together with the particles that drifted across the slab
//              boundary (migration).  The search grid has cell edge >= h (Domain.cpp:10-22), so ONE halo
//              layer per side is enough for every later pass.
//   exchange 2 (after K3, = updateGhostState): the packed record pk1 (x, v, rho, P, c_s, omega) of the boundary layers.
//   exchange 3 (after K3b, = updateGhostGradients): the packed record pk2 (B^-1, limited gradients); dt: allreduce-min.
//
// Slabs are whole cell layers along the slowest-varying cell axis (y in 2D, z in 3D: cell id = iX + iY*cellsX
// + iZ*cellsX*cellsY, Particles.cpp:298-302).  After the cell sort the local particle arrays are ordered
// [lower halo layer | owned layers | upper halo layer] and, inside a layer, by (cell, original id) -- the SAME
// order on the sending and on the receiving rank.  Exchanges 2 and 3 are therefore ONE contiguous-range
// ncclSend/ncclRecv per side of the packed per-particle records, no packing kernel, no index lists.  A face cut by a slab boundary is
// evaluated by both ranks from bit-identical inputs in the canonical orientation (quirk Q4, global original
// ids), so no flux is exchanged and conservation holds to round-off.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the host is
// Python, else the system one): single-GPU users need no NCCL at all.
#include "mlh_internal.cuh"

#include <cfloat>
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void *handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};
NcclApi g_nccl = {};

bool load_nccl(char *err, size_t errlen) {
    if (g_nccl.handle) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(err, errlen, "cannot load libnccl.so.2: %s", dlerror());
        return false;
    }
#define SYM(field, name)                                                  \
    *(void **)(&g_nccl.field) = dlsym(h, name);                           \
    if (!g_nccl.field) {                                                  \
        snprintf(err, errlen, "libnccl.so.2 lacks %s", name);             \
        return false;                                                     \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return true;
}

#define MLH_NCCL_CHECK(c, expr)                                                                          \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            snprintf((c)->err, sizeof((c)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,           \
                     g_nccl.GetErrorString(_r));                                                         \
            return MLH_E_COMM;                                                                           \
        }                                                                                                \
    } while (0)

// global cell layer of a coordinate along the slab axis: the reference's cell formula (Particles.cpp:279-295)
__device__ __forceinline__ int global_layer(const Grid &g, double x) {
    const int k = g.slab_dim;
    double f = floor(__ddiv_rn(__dsub_rn(x, g.bmin[k]), g.cell_size[k]));
    int fi = (f >= 2147483647.0 || f <= -2147483648.0 || f != f) ? -1 : (int)f;
    if (fi == g.cells[k]) fi -= 1;
    return fi;
}

// Exchange 1, sender side: classify every owned particle by its global layer and append it to the
// send buffer of the lower and/or upper slab neighbour (boundary-layer copy or migrant).
//   buf layout: field f of direction d at buf[(d * nf + f) * cap + slot], fields = x[D], v[D], m, u; ids apart
template <int D>
__global__ void __launch_bounds__(256) k_halo_pack(const Params p, int lo, int hi, int periodic, int has_dn, int has_up,
                                                   double *buf, int *idbuf, int cap, int *counts) {
    __shared__ Grid s_grid;
    if (threadIdx.x == 0) s_grid = *p.d.grid;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.ncur) return;
    const Grid &g = s_grid;
    const int L = g.cells[g.slab_dim];
    const int l = global_layer(g, p.d.cx[g.slab_dim][i]);
    int below = lo - 1, above = hi; // layers just outside the slab
    if (periodic) {
        if (below < 0) below += L;
        if (above >= L) above -= L;
    }
    const bool owned = l >= lo && l < hi;
    const bool to_dn = has_dn && ((owned && l == lo) || (!owned && l == below));
    const bool to_up = has_up && ((owned && l == hi - 1) || (!owned && l == above));
    if (!owned && l != below && l != above) atomicOr(p.d.flags, MLH_F_MIGRATION);
    constexpr int NF = 2 * D + 2;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (!(d == 0 ? to_dn : to_up)) continue;
        const int slot = atomicAdd(&counts[d], 1);
        if (slot >= cap) {
            atomicOr(p.d.flags, MLH_F_HALO_OVERFLOW);
            continue;
        }
        double *b = buf + (size_t)d * NF * cap + slot;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            b[(size_t)k * cap] = p.d.cx[k][i];
            b[(size_t)(D + k) * cap] = p.d.cv[k][i];
        }
        b[(size_t)(2 * D) * cap] = p.d.cm[i];
        b[(size_t)(2 * D + 1) * cap] = p.d.cu[i];
        idbuf[(size_t)d * cap + slot] = p.d.cid[i];
    }
}

int halo_alloc(mlh_ctx *c) {
    if (c->halo_buf) return MLH_OK;
    const Params &p = c->p;
    const int NF = 2 * p.D + 2;
    long cap = (c->capacity - c->n_owned) / 2;
    if (cap < 1024) cap = 1024;
    c->halo_cap = (int)cap;
    MLH_CUDA_CHECK(c, cudaMalloc(&c->halo_buf, sizeof(double) * (size_t)2 * NF * cap));
    MLH_CUDA_CHECK(c, cudaMalloc(&c->halo_ids, sizeof(int) * (size_t)2 * cap));
    MLH_CUDA_CHECK(c, cudaMalloc(&c->halo_counts, sizeof(int) * 8));
    MLH_CUDA_CHECK(c, cudaMallocHost(&c->h_counts, sizeof(int) * 8));
    return MLH_OK;
}

} // namespace

extern "C" int mlh_comm_unique_id(char *id128) {
    if (!id128) return MLH_E_INVALID;
    char err[256];
    if (!load_nccl(err, sizeof(err))) return MLH_E_COMM;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return MLH_E_COMM;
    static_assert(sizeof(id) == MLH_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(id128, &id, sizeof(id));
    return MLH_OK;
}

extern "C" int mlh_comm_init(mlh_ctx *c, const char *id128) {
    if (!c || !id128) return MLH_E_INVALID;
    if (c->cfg.nranks <= 1) {
        snprintf(c->err, sizeof(c->err), "mlh_comm_init: context was created with nranks <= 1");
        return MLH_E_INVALID;
    }
    if (!load_nccl(c->err, sizeof(c->err))) return MLH_E_COMM;
    cudaSetDevice(c->cfg.device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    MLH_NCCL_CHECK(c, g_nccl.CommInitRank(&comm, c->cfg.nranks, id, c->cfg.rank));
    c->nccl_comm = comm;
    return MLH_OK;
}

void mlh_comm_destroy(mlh_ctx *c) {
    if (c->nccl_comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
    if (c->halo_buf) cudaFree(c->halo_buf);
    if (c->halo_ids) cudaFree(c->halo_ids);
    if (c->halo_counts) cudaFree(c->halo_counts);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    c->halo_buf = nullptr;
    c->halo_ids = nullptr;
    c->halo_counts = nullptr;
    c->h_counts = nullptr;
}

// slab neighbours of this rank (-1: none)
static void slab_neighbours(const mlh_ctx *c, int *dn, int *up) {
    const int R = c->cfg.nranks, r = c->cfg.rank;
    *dn = r - 1;
    *up = r + 1;
    if (c->p.periodic) {
        if (*dn < 0) *dn = R - 1;
        if (*up >= R) *up = 0;
    } else {
        if (*up >= R) *up = -1;
    }
}

// allreduce of the bounding box (non-periodic runs rebuild the grid from it every step,
// MeshlessScheme.cpp:41-51).  d.bbox = [min3 | max3 over id != 0 | x of id 0 or -DBL_MAX]; in place.
int mlh_comm_bbox(mlh_ctx *c) {
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    if (!comm) {
        snprintf(c->err, sizeof(c->err), "multi-GPU context without communicator: call mlh_comm_init");
        return MLH_E_STATE;
    }
    double *b = c->p.d.bbox;
    MLH_NCCL_CHECK(c, g_nccl.GroupStart());
    // [0..5] are order-preserving keys (dbl_key): min / max of the keys == key of the min / max
    MLH_NCCL_CHECK(c, g_nccl.AllReduce(b, b, 3, ncclUint64, ncclMin, comm, c->stream));
    MLH_NCCL_CHECK(c, g_nccl.AllReduce(b + 3, b + 3, 3, ncclUint64, ncclMax, comm, c->stream));
    MLH_NCCL_CHECK(c, g_nccl.AllReduce(b + 6, b + 6, 3, ncclDouble, ncclMax, comm, c->stream));
    MLH_NCCL_CHECK(c, g_nccl.GroupEnd());
    c->launches += 1;
    return MLH_OK;
}

// Exchange 1.  On entry the CUR set holds the ncur particles this rank owned at the end of the last
// step (or uploaded); on exit it additionally holds the neighbours' boundary layers and the particles
// that migrated in; migrants that left stay as halo copies.  The grid (make_grid) must be current.
int mlh_halo_exchange_particles(mlh_ctx *c) {
    Params &p = c->p;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    if (!comm) {
        snprintf(c->err, sizeof(c->err), "multi-GPU context without communicator: call mlh_comm_init");
        return MLH_E_STATE;
    }
    int rc = halo_alloc(c);
    if (rc != MLH_OK) return rc;
    const int D = p.D, NF = 2 * D + 2, cap = c->halo_cap;
    int dn, up;
    slab_neighbours(c, &dn, &up);
    cudaStream_t st = c->stream;
    int *cnt = c->halo_counts; // [0] to dn, [1] to up, [2] from dn, [3] from up
    MLH_CUDA_CHECK(c, cudaMemsetAsync(cnt, 0, sizeof(int) * 8, st));
    mlh_prof_begin(c, KID_HALO);
    if (D == 2)
        k_halo_pack<2><<<mlh_blocks(p.ncur, 256), 256, 0, st>>>(p, c->layer_lo, c->layer_hi, p.periodic, dn >= 0, up >= 0,
                                                             c->halo_buf, c->halo_ids, cap, cnt);
    else
        k_halo_pack<3><<<mlh_blocks(p.ncur, 256), 256, 0, st>>>(p, c->layer_lo, c->layer_hi, p.periodic, dn >= 0, up >= 0,
                                                             c->halo_buf, c->halo_ids, cap, cnt);
    mlh_prof_end(c, KID_HALO);
    MLH_CUDA_CHECK(c, cudaGetLastError());
    // counts: tell each neighbour how many particles follow.  Between one pair of ranks NCCL matches
    // sends and receives in issue order; both sides issue (dn-send, up-recv, up-send, dn-recv), so with
    // two ranks in a periodic box (dn == up) the peer's "to dn" message lands in our "from up" slot.
    MLH_NCCL_CHECK(c, g_nccl.GroupStart());
    if (dn >= 0) MLH_NCCL_CHECK(c, g_nccl.Send(cnt + 0, 1, ncclInt32, dn, comm, st));
    if (up >= 0) MLH_NCCL_CHECK(c, g_nccl.Recv(cnt + 3, 1, ncclInt32, up, comm, st));
    if (up >= 0) MLH_NCCL_CHECK(c, g_nccl.Send(cnt + 1, 1, ncclInt32, up, comm, st));
    if (dn >= 0) MLH_NCCL_CHECK(c, g_nccl.Recv(cnt + 2, 1, ncclInt32, dn, comm, st));
    MLH_NCCL_CHECK(c, g_nccl.GroupEnd());
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_counts, cnt, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(st));
    const int s_dn = c->h_counts[0], s_up = c->h_counts[1], r_dn = c->h_counts[2], r_up = c->h_counts[3];
    if (s_dn > cap || s_up > cap) {
        snprintf(c->err, sizeof(c->err), "halo send buffer overflow (%d / %d particles, capacity %d): raise mlh_config.capacity",
                 s_dn, s_up, cap);
        return MLH_E_DEVICE_FLAG;
    }
    const long total = (long)p.ncur + r_dn + r_up;
    if (total > c->capacity) {
        snprintf(c->err, sizeof(c->err), "rank %d: %ld particles incl. halo exceed capacity %ld: raise mlh_config.capacity",
                 c->cfg.rank, total, c->capacity);
        return MLH_E_DEVICE_FLAG;
    }
    const size_t o_dn = (size_t)p.ncur, o_up = (size_t)p.ncur + r_dn;
    double *dst[8];
    for (int k = 0; k < D; ++k) {
        dst[k] = p.d.cx[k];
        dst[D + k] = p.d.cv[k];
    }
    dst[2 * D] = p.d.cm;
    dst[2 * D + 1] = p.d.cu;
    MLH_NCCL_CHECK(c, g_nccl.GroupStart());
    for (int f = 0; f < NF; ++f) {
        const double *b_dn = c->halo_buf + (size_t)(0 * NF + f) * cap, *b_up = c->halo_buf + (size_t)(1 * NF + f) * cap;
        if (dn >= 0 && s_dn) MLH_NCCL_CHECK(c, g_nccl.Send(b_dn, s_dn, ncclDouble, dn, comm, st));
        if (up >= 0 && r_up) MLH_NCCL_CHECK(c, g_nccl.Recv(dst[f] + o_up, r_up, ncclDouble, up, comm, st));
        if (up >= 0 && s_up) MLH_NCCL_CHECK(c, g_nccl.Send(b_up, s_up, ncclDouble, up, comm, st));
        if (dn >= 0 && r_dn) MLH_NCCL_CHECK(c, g_nccl.Recv(dst[f] + o_dn, r_dn, ncclDouble, dn, comm, st));
    }
    if (dn >= 0 && s_dn) MLH_NCCL_CHECK(c, g_nccl.Send(c->halo_ids, s_dn, ncclInt32, dn, comm, st));
    if (up >= 0 && r_up) MLH_NCCL_CHECK(c, g_nccl.Recv(p.d.cid + o_up, r_up, ncclInt32, up, comm, st));
    if (up >= 0 && s_up) MLH_NCCL_CHECK(c, g_nccl.Send(c->halo_ids + cap, s_up, ncclInt32, up, comm, st));
    if (dn >= 0 && r_dn) MLH_NCCL_CHECK(c, g_nccl.Recv(p.d.cid + o_dn, r_dn, ncclInt32, dn, comm, st));
    MLH_NCCL_CHECK(c, g_nccl.GroupEnd());
    c->launches += 2; // counts + payload (each one fused NCCL kernel)
    p.ncur = (int)total;
    return MLH_OK;
}

// After the cell sort: where the halo layers and the boundary layers sit in the sorted arrays.
int mlh_halo_read_layout(mlh_ctx *c) {
    Params &p = c->p;
    const Grid &g = p.grid;
    const int layer_cells = g.ncells / g.lcells[g.slab_dim];
    const int Lloc = g.lcells[g.slab_dim]; // owned layers + 2
    const int idx[4] = {layer_cells, 2 * layer_cells, (Lloc - 2) * layer_cells, (Lloc - 1) * layer_cells};
    for (int k = 0; k < 4; ++k)
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_counts + 4 + k, p.d.cell_start + idx[k], sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    p.own_begin = c->h_counts[4];
    c->lo_layer_end = c->h_counts[5];   // owned bottom layer = [own_begin, lo_layer_end)
    c->hi_layer_begin = c->h_counts[6]; // owned top layer    = [hi_layer_begin, own_end)
    p.own_end = c->h_counts[7];
    c->n_owned = p.own_end - p.own_begin;
    return MLH_OK;
}

// Exchanges 2 and 3: refresh the halo ranges of the given SoA arrays from the neighbours' boundary
// layers (contiguous ranges, identical order on both sides -- see the header comment).
int mlh_halo_refresh(mlh_ctx *c, double *const *arrays, int narrays, int width) {
    Params &p = c->p;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    int dn, up;
    slab_neighbours(c, &dn, &up);
    cudaStream_t st = c->stream;
    const size_t w = (size_t)width; // doubles per particle in these arrays
    const size_t n_lo_halo = p.own_begin * w, n_hi_halo = (size_t)(p.n - p.own_end) * w;
    const size_t n_lo_layer = (size_t)(c->lo_layer_end - p.own_begin) * w, n_hi_layer = (size_t)(p.own_end - c->hi_layer_begin) * w;
    mlh_prof_begin(c, KID_HALO);
    MLH_NCCL_CHECK(c, g_nccl.GroupStart());
    for (int a = 0; a < narrays; ++a) {
        double *arr = arrays[a];
        if (dn >= 0 && n_lo_layer) MLH_NCCL_CHECK(c, g_nccl.Send(arr + p.own_begin * w, n_lo_layer, ncclDouble, dn, comm, st));
        if (up >= 0 && n_hi_halo) MLH_NCCL_CHECK(c, g_nccl.Recv(arr + p.own_end * w, n_hi_halo, ncclDouble, up, comm, st));
        if (up >= 0 && n_hi_layer) MLH_NCCL_CHECK(c, g_nccl.Send(arr + c->hi_layer_begin * w, n_hi_layer, ncclDouble, up, comm, st));
        if (dn >= 0 && n_lo_halo) MLH_NCCL_CHECK(c, g_nccl.Recv(arr, n_lo_halo, ncclDouble, dn, comm, st));
    }
    MLH_NCCL_CHECK(c, g_nccl.GroupEnd());
    mlh_prof_end(c, KID_HALO);
    return MLH_OK;
}

// dt = min over ranks (compGlobalTimestep is a global minimum, Particles.cpp:1446-1485); positive doubles
// order like their bit patterns, so the accumulator is reduced as a double
int mlh_comm_min_dt(mlh_ctx *c) {
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    MLH_NCCL_CHECK(c, g_nccl.AllReduce(c->p.d.dt_bits, c->p.d.dt_bits, 1, ncclDouble, ncclMin, comm, c->stream));
    c->launches += 1;
    return MLH_OK;
}

int mlh_comm_sum(mlh_ctx *c, double *dev, int n) {
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    MLH_NCCL_CHECK(c, g_nccl.AllReduce(dev, dev, n, ncclDouble, ncclSum, comm, c->stream));
    c->launches += 1;
    return MLH_OK;
}
