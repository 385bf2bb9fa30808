// C-ABI layer of the B200 MFV hot path (include/mlh_gpu.h): context, HBM pool, phase sequencing
// (MeshlessScheme::run order, /root/reference/demonstrator/src/MeshlessScheme.cpp:39-253), transfers
// and the parity/measurement hooks.  No physics here -- that is in k1..k5.
#include "mlh_internal.cuh"

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

static char g_create_err[512] = "";

static const char *kKernelNames[KID_COUNT] = {
    "k0_bbox",        "k1_cell_key",   "k1_scan",        "k1_scatter_perm", "k1_sort_within_cells",
    "k1_gather",      "k2_neighbours", "k3_density_matrix", "k3b_gradient_limit", "k2b_face_index", "k4_select_dt",
    "k4a_face_states", "k4b1_face_setup", "k4b_face_riemann", "k4b3_face_finish", "k4c_flux_sum_update", "k5_sums", "k5_unpermute", "halo_exchange"};
static_assert(sizeof(kKernelNames) / sizeof(kKernelNames[0]) == KID_COUNT, "one name per KernelId");

// ------------------------------------------------------------------------------------------------
// profiling brackets (CUDA events on the context's stream)
// ------------------------------------------------------------------------------------------------
static void prof_flush(mlh_ctx *c) {
    if (c->ev_used == 0) return;
    cudaEventSynchronize(c->ev[2 * (c->ev_used - 1) + 1]);
    for (int k = 0; k < c->ev_used; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[2 * k], c->ev[2 * k + 1]);
        c->prof_ms[c->ev_kernel[k]] += ms;
    }
    c->ev_used = 0;
}
void mlh_prof_begin(mlh_ctx *c, int kid) {
    if (!c->profiling) return;
    if (c->ev_used == 64) prof_flush(c);
    c->ev_kernel[c->ev_used] = kid;
    cudaEventRecord(c->ev[2 * c->ev_used], c->stream);
}
void mlh_prof_end(mlh_ctx *c, int kid) {
    c->launches++;
    c->prof_launches[kid]++;
    if (!c->profiling) return;
    cudaEventRecord(c->ev[2 * c->ev_used + 1], c->stream);
    c->ev_used++;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
    char *base;
    size_t off;
    template <typename T> T *take(size_t count) {
        T *ptr = base ? (T *)(base + off) : nullptr;
        off = align_up(off + count * sizeof(T), 256);
        return ptr;
    }
};

static void carve_pool(mlh_ctx *c, Carver &cv) {
    Params &p = c->p;
    const size_t n = (size_t)p.ncap;
    const int D = p.D;
    DevPtrs &d = p.d;
    for (int k = 0; k < D; ++k) {
        d.x[k] = cv.take<double>(n);
        d.v[k] = cv.take<double>(n);
        d.cx[k] = cv.take<double>(n);
        d.cv[k] = cv.take<double>(n);
    }
    d.m = cv.take<double>(n); d.u = cv.take<double>(n);
    d.cm = cv.take<double>(n); d.cu = cv.take<double>(n);
    d.rho = cv.take<double>(n); d.P = cv.take<double>(n); d.omega = cv.take<double>(n); d.cs = cv.take<double>(n);
    for (int k = 0; k < D * D; ++k) d.B[k] = cv.take<double>(n);
    for (int f = 0; f < 5; ++f) {
        if (f == 3 && D == 2) continue;
        for (int a = 0; a < D; ++a) d.g[f * 3 + a] = cv.take<double>(n);
    }
    d.pk1 = cv.take<double>(n * (size_t)MLH_PK1(D));
    d.pk2 = cv.take<double>(n * (size_t)MLH_PK2(D));
    d.id = cv.take<int>(n); d.cid = cv.take<int>(n); d.cell = cv.take<int>(n);
    d.noi = cv.take<int>(n); d.noig = cv.take<int>(n);
    d.ckey = cv.take<int>(n); d.crank = cv.take<int>(n); d.perm = cv.take<int>(n);
    d.nnl = cv.take<int>(n * (size_t)p.max_ni);
    d.fmap = cv.take<unsigned>(n * (size_t)p.max_ni);
    d.grp = cv.take<unsigned short>(n * (size_t)27);
    d.nown = cv.take<int>(n);
    d.face_start = cv.take<int>(n + 1);
    d.face_scan_tmp = cv.take<int>(n / 1024 + 4);
    d.fa = cv.take<int>((size_t)p.fcap);
    d.fe = cv.take<int>((size_t)p.fcap);
    d.F = cv.take<double>((size_t)p.fcap * MLH_FREC(D));
    if (p.debug_capture) {
        for (int f = 0; f < 5; ++f) {
            if (f == 3 && D == 2) continue;
            for (int a = 0; a < D; ++a) d.gpre[f * 3 + a] = cv.take<double>(n);
        }
        for (int k = 0; k < 2 + D; ++k) d.flux[k] = cv.take<double>(n);
        d.dbg_face = cv.take<double>((size_t)p.fcap * (4 * D + 4));
    }
    d.dt_bits = cv.take<unsigned long long>(1);
    d.dt_used = cv.take<double>(1);
    d.bbox = cv.take<double>(12);
    d.grid = cv.take<Grid>(1);
    d.sums = cv.take<double>(6);
    d.flags = cv.take<unsigned>(1);
    d.counters = cv.take<unsigned>(4);
}

// Domain::createGrid, Domain.cpp:9-54 (same double arithmetic on the host)
static int make_grid(mlh_ctx *c, const double *bmin, const double *bmax) {
    Params &p = c->p;
    Grid &g = p.grid;
    long long nc = 1;
    for (int k = 0; k < 3; ++k) {
        g.cells[k] = 1;
        g.lcells[k] = 1;
        g.bmin[k] = g.bmax[k] = g.cell_size[k] = 0.;
    }
    for (int k = 0; k < p.D; ++k) {
        g.bmin[k] = bmin[k];
        g.bmax[k] = bmax[k];
        double cells = floor((bmax[k] - bmin[k]) / p.h);
        if (!(cells >= 1.) || cells > 2.0e9) {
            snprintf(c->err, sizeof(c->err), "search grid has %g cells along axis %d (box [%g,%g], kernelSize %g)",
                     cells, k, bmin[k], bmax[k], p.h);
            return MLH_E_INVALID;
        }
        g.cells[k] = (int)cells;
        g.cell_size[k] = (bmax[k] - bmin[k]) / (double)g.cells[k];
        g.lcells[k] = g.cells[k];
    }
    g.slab_dim = p.D - 1;
    g.sliced = 0;
    g.layer0 = 0;
    if (c->cfg.nranks > 1) {
        g.sliced = 1;
        int lo, hi;
        mlh_slab_range(g.cells[g.slab_dim], c->cfg.nranks, c->cfg.rank, &lo, &hi);
        c->layer_lo = lo;
        c->layer_hi = hi;
        c->n_layers_global = g.cells[g.slab_dim];
        g.layer0 = lo - 1;
        g.lcells[g.slab_dim] = hi - lo + 2;
    }
    for (int k = 0; k < p.D; ++k) {
        nc *= g.lcells[k];
        if (p.periodic && g.cells[k] < 3) {
            snprintf(c->err, sizeof(c->err), "periodic box needs >= 3 search cells per axis (axis %d has %d)", k, g.cells[k]);
            return MLH_E_INVALID;
        }
    }
    if (nc + 1 > 2147483647LL) {
        snprintf(c->err, sizeof(c->err), "search grid too large (%lld cells)", nc);
        return MLH_E_INVALID;
    }
    g.ncells = (int)nc;
    return MLH_OK;
}

// host grid -> device copy read by K1 / K2 / the halo kernels (pinned staging: the caller has just synchronised the
// stream or is outside the time loop, so the staging buffer is not in flight)
static int upload_grid(mlh_ctx *c) {
    if (!c->pool) return MLH_OK; // done by mlh_upload once the pool exists
    c->h_grid[0] = c->p.grid;
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->p.d.grid, &c->h_grid[0], sizeof(Grid), cudaMemcpyHostToDevice, c->stream));
    c->grid_host_current = true;
    c->grid_mirror_pending = false;
    return MLH_OK;
}

// cell arrays for at least `need` entries (cell_count / cell_start / scan_tmp); synchronises when it has to grow
static int reserve_cells(mlh_ctx *c, long long need) {
    Params &p = c->p;
    if (need <= c->max_cells) return MLH_OK;
    if (p.d.cell_count) {
        cudaStreamSynchronize(c->stream);
        cudaFree(p.d.cell_count);
        cudaFree(p.d.cell_start);
        cudaFree(p.d.scan_tmp);
    }
    c->max_cells = (int)std::min<long long>(2147483647LL, (long long)(need * 1.25) + 1024);
    MLH_CUDA_CHECK(c, cudaMalloc(&p.d.cell_count, sizeof(int) * (size_t)c->max_cells));
    MLH_CUDA_CHECK(c, cudaMalloc(&p.d.cell_start, sizeof(int) * (size_t)c->max_cells));
    MLH_CUDA_CHECK(c, cudaMalloc(&p.d.scan_tmp, sizeof(int) * (size_t)(c->max_cells / 1024 + 2)));
    return MLH_OK;
}

static int check_ctx(mlh_ctx *c) {
    if (!c) return MLH_E_INVALID;
    cudaSetDevice(c->cfg.device);
    return MLH_OK;
}

__global__ void k_iota(int *dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = i;
}

// DFMA throughput probe: 8 independent chains per thread, `iters` x 8 DFMA each
__global__ void __launch_bounds__(256) k_dfma_peak(double *sink, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
#pragma unroll 4
    for (int k = 0; k < iters; ++k) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    double t = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (t == 123.456) sink[0] = t; // never true; keeps the chains alive
}

// ------------------------------------------------------------------------------------------------
// public API
// ------------------------------------------------------------------------------------------------
extern "C" {

int mlh_host_alloc(unsigned long bytes, void **ptr) {
    if (!ptr || bytes == 0) return MLH_E_INVALID;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaMallocHost(%lu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        *ptr = nullptr;
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MLH_E_NO_DEVICE : MLH_E_CUDA;
    }
    return MLH_OK;
}
int mlh_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? MLH_OK : MLH_E_CUDA; }

int mlh_measure_fp64_peak(int device, double *tflops) {
    if (!tflops) return MLH_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MLH_E_NO_DEVICE; }
    if (device < 0 || device >= ndev) return MLH_E_INVALID;
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double *sink = nullptr;
    if (cudaMalloc(&sink, 64) != cudaSuccess) return MLH_E_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    double best = 0.;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, 256>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? MLH_OK : MLH_E_CUDA;
}

int mlh_abi_version(void) { return MLH_ABI_VERSION; }

void mlh_default_config(mlh_config *cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->dim = 2;            // parameter.h:9
    cfg->periodic = 1;       // :12
    cfg->max_interactions = 128;
    cfg->slope_limiting = 1; // :28
    cfg->pairwise_limiter = 1; // :34
    cfg->meshless_finite_mass = 0;
    cfg->move_particles = 1; // :45
    cfg->abs_mode = MLH_ABS_FABS;
    cfg->q13_mode = MLH_Q13_ZERO_Z;
    cfg->q3_mode = MLH_Q3_REFERENCE;
    cfg->first_order_quad_point = 1; // :55
    cfg->cfl = .2;           // :18
    cfg->beta = 4.;          // :31
    cfg->psi1 = .5;
    cfg->psi2 = .25;
    cfg->kernel_size = .025; // demonstrator/config.info:30
    cfg->gamma = 1.6666666666666667;
    cfg->box[0] = 0.; cfg->box[1] = 0.; cfg->box[2] = 1.; cfg->box[3] = 1.;
    cfg->nranks = 1;
}

const char *mlh_last_error(const mlh_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int mlh_create(const mlh_config *cfg, mlh_ctx **out) {
    if (!cfg || !out) return MLH_E_INVALID;
    *out = nullptr;
    if (cfg->dim != 2 && cfg->dim != 3) {
        snprintf(g_create_err, sizeof(g_create_err), "dim must be 2 or 3 (got %d)", cfg->dim);
        return MLH_E_INVALID;
    }
    if (cfg->max_interactions > 1023) { // the slot -> face map packs a per-particle rank into 10 bits
        snprintf(g_create_err, sizeof(g_create_err), "max_interactions must be <= 1023 (got %d)", cfg->max_interactions);
        return MLH_E_INVALID;
    }
    if (!(cfg->kernel_size > 0.) || !(cfg->gamma > 1.)) {
        snprintf(g_create_err, sizeof(g_create_err), "kernel_size must be > 0 and gamma > 1");
        return MLH_E_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        snprintf(g_create_err, sizeof(g_create_err),
                 "no CUDA device (%s): the MFV path has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return MLH_E_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        snprintf(g_create_err, sizeof(g_create_err), "device %d out of range (%d devices)", cfg->device, ndev);
        return MLH_E_INVALID;
    }
    mlh_ctx *c = new mlh_ctx();
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    if (c->cfg.nranks < 1) c->cfg.nranks = 1;
    if (c->cfg.max_interactions <= 0) c->cfg.max_interactions = 128;
    cudaSetDevice(cfg->device);
    {
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, cfg->device);
        c->num_sms = prop.multiProcessorCount; // 148 on B200: persistent kernels launch a multiple of it
    }
    Params &p = c->p;
    p.D = cfg->dim;
    p.periodic = cfg->periodic ? 1 : 0;
    p.max_ni = c->cfg.max_interactions;
    p.slope_limiting = cfg->slope_limiting;
    p.pairwise = cfg->pairwise_limiter;
    p.mfm = cfg->meshless_finite_mass;
    p.move_particles = cfg->move_particles;
    p.abs_mode = cfg->abs_mode;
    p.q13_mode = cfg->q13_mode;
    p.q3_mode = cfg->q3_mode;
    p.symmetric_seam = cfg->symmetric_seam;
    p.debug_capture = cfg->debug_capture;
    p.quad_h4 = cfg->first_order_quad_point ? 0 : 1;
    p.h4 = cfg->kernel_size / 4.; // `kernelSize/4.` of Particles.cpp:1358,1515
    p.h = cfg->kernel_size;
    p.hSqr = cfg->kernel_size * cfg->kernel_size; // Particles.cpp:334
    p.gamma = cfg->gamma;
    p.cfl = cfg->cfl;
    p.beta = cfg->beta;
    p.psi1 = cfg->psi1;
    p.psi2 = cfg->psi2;
    {   // Kernel::cubicSpline constants, Particles.cpp:10-15
        double h2 = p.h / 2.;
        p.h2 = h2;
        p.inv_h2 = 1. / h2; // correctly rounded on the host
        p.sigma = (p.D == 2) ? 10. / (7. * M_PI * h2 * h2) : 1. / (M_PI * h2 * h2 * h2);
        p.sigma4 = p.sigma / 4.;
    }
    {   // RiemannSolver(gamma)
        const double g = cfg->gamma;
        p.rs.gamma = g;
        p.rs.gp1d2g = 0.5 * (g + 1.) / g;
        p.rs.gm1d2g = 0.5 * (g - 1.) / g;
        p.rs.gm1dgp1 = (g - 1.) / (g + 1.);
        p.rs.tdgp1 = 2. / (g + 1.);
        p.rs.tdgm1 = 2. / (g - 1.);
        p.rs.gm1d2 = 0.5 * (g - 1.);
        p.rs.tgdgm1 = 2. * g / (g - 1.);
        p.rs.ginv = 1. / g;
        p.rs.sqrt_tdgp1 = sqrt(p.rs.tdgp1);
        p.rs.root_n = 0;
        for (int n = 5; n <= 7; n += 2)
            if (fabs(p.rs.gm1d2g * n - 1.) < 4e-16) p.rs.root_n = n;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        snprintf(g_create_err, sizeof(g_create_err), "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return MLH_E_CUDA;
    }
    for (int k = 0; k < 128; ++k) cudaEventCreate(&c->ev[k]);
    cudaEventCreate(&c->timer[0]);
    cudaEventCreate(&c->timer[1]);
    cudaMallocHost(&c->h_small, 64 * sizeof(double));
    cudaMallocHost(&c->h_grid, 2 * sizeof(Grid)); // [0] host -> device staging, [1] asynchronous mirror of the device grid
    cudaEventCreateWithFlags(&c->ev_grid, cudaEventDisableTiming);
    cudaMallocHost(&c->h_flags, 8 * sizeof(unsigned));
    if (p.periodic) {
        double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
        for (int k = 0; k < p.D; ++k) {
            bmin[k] = cfg->box[k];
            bmax[k] = cfg->box[p.D + k];
        }
        int rc = make_grid(c, bmin, bmax); // MeshlessScheme.cpp:17
        if (rc == MLH_OK) rc = reserve_cells(c, (long long)c->p.grid.ncells + 2);
        if (rc != MLH_OK) {
            snprintf(g_create_err, sizeof(g_create_err), "%s", c->err);
            mlh_destroy(c);
            return rc;
        }
    }
    *out = c;
    return MLH_OK;
}

int mlh_destroy(mlh_ctx *c) {
    if (!c) return MLH_E_INVALID;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    mlh_comm_destroy(c);
    if (c->pool) cudaFree(c->pool);
    if (c->dl_scratch) cudaFree(c->dl_scratch);
    if (c->rf_buf) cudaFree(c->rf_buf);
    if (c->stage) cudaFree(c->stage);
    if (c->p.d.cell_count) {
        cudaFree(c->p.d.cell_count);
        cudaFree(c->p.d.cell_start);
        cudaFree(c->p.d.scan_tmp);
    }
    for (int k = 0; k < 128; ++k) cudaEventDestroy(c->ev[k]);
    cudaEventDestroy(c->timer[0]);
    cudaEventDestroy(c->timer[1]);
    cudaFreeHost(c->h_small);
    cudaFreeHost(c->h_flags);
    cudaFreeHost(c->h_grid);
    cudaEventDestroy(c->ev_grid);
    cudaStreamDestroy(c->stream);
    delete c;
    return MLH_OK;
}

unsigned mlh_error_flags(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK || !c->pool) return 0;
    cudaMemcpyAsync(c->h_flags, c->p.d.flags, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    return c->h_flags[0];
}

int mlh_upload(mlh_ctx *c, long N, const double *x, const double *y, const double *z, const double *vx,
               const double *vy, const double *vz, const double *m, const double *u, const int *global_ids) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    Params &p = c->p;
    if (N <= 0 || N > MLH_NNL_IDX_MASK || !x || !y || !vx || !vy || !m || !u || (p.D == 3 && (!z || !vz))) {
        snprintf(c->err, sizeof(c->err), "mlh_upload: bad arguments (N=%ld, need 0 < N < 2^26, z/vz required in 3D)", N);
        return MLH_E_INVALID;
    }
    if (c->cfg.nranks == 1 && global_ids) {
        // the download / parity paths scatter by id into N-element buffers: ids must be a permutation of 0..N-1
        std::vector<unsigned char> seen((size_t)N, 0);
        for (long i = 0; i < N; ++i) {
            const int id = global_ids[i];
            if (id < 0 || id >= N || seen[(size_t)id]) {
                snprintf(c->err, sizeof(c->err), "mlh_upload: global_ids must be a permutation of 0..N-1 on a single GPU (entry %ld = %d)", i, id);
                return MLH_E_INVALID;
            }
            seen[(size_t)id] = 1;
        }
    }
    long cap = c->cfg.capacity > 0 ? c->cfg.capacity : (c->cfg.nranks > 1 ? N + N / 2 + 4096 : N);
    if (cap < N) cap = N;
    cap = (long)align_up((size_t)cap, 32);
    if ((unsigned long long)cap * (unsigned long long)p.max_ni >= (1ull << 32)) { // slot arrays are indexed with 32 bits
        snprintf(c->err, sizeof(c->err), "mlh_upload: capacity %ld x max_interactions %d exceeds 2^32 list slots: lower max_interactions", cap, p.max_ni);
        return MLH_E_INVALID;
    }
    if (!c->pool || cap > c->capacity) {
        if (c->pool) {
            cudaStreamSynchronize(c->stream);
            cudaFree(c->pool);
            c->pool = nullptr;
            if (c->dl_scratch) cudaFree(c->dl_scratch);
            c->dl_scratch = nullptr;
        }
        p.ncap = (int)cap;
        c->capacity = cap;
        {   // every pair of the lists is one face; pairs inside this rank appear in two lists
            long long fc = (long long)cap * p.max_ni / 2 + 1024;
            if (fc > (1LL << 30) - 1) fc = (1LL << 30) - 1;
            p.fcap = (int)fc;
        }
        Carver sizing{nullptr, 0};
        carve_pool(c, sizing);
        c->pool_bytes = sizing.off;
        cudaError_t e = cudaMalloc(&c->pool, c->pool_bytes);
        if (e != cudaSuccess) {
            snprintf(c->err, sizeof(c->err), "cudaMalloc of %.2f GB for %ld particles failed: %s", c->pool_bytes / 1e9, cap,
                     cudaGetErrorString(e));
            cudaGetLastError();
            c->pool = nullptr;
            return MLH_E_CUDA;
        }
        Carver cv{(char *)c->pool, 0};
        carve_pool(c, cv);
        MLH_CUDA_CHECK(c, cudaMemsetAsync(c->pool, 0, c->pool_bytes, c->stream));
    }
    const size_t nb = sizeof(double) * (size_t)N;
    cudaStream_t st = c->stream;
    const double *xs[3] = {x, y, z}, *vs[3] = {vx, vy, vz};
    for (int k = 0; k < p.D; ++k) {
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(p.d.cx[k], xs[k], nb, cudaMemcpyHostToDevice, st));
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(p.d.cv[k], vs[k], nb, cudaMemcpyHostToDevice, st));
    }
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(p.d.cm, m, nb, cudaMemcpyHostToDevice, st));
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(p.d.cu, u, nb, cudaMemcpyHostToDevice, st));
    if (global_ids) {
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(p.d.cid, global_ids, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    } else {
        k_iota<<<mlh_blocks(N, 256), 256, 0, st>>>(p.d.cid, (int)N);
        c->launches++;
    }
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.flags, 0, sizeof(unsigned), st));
    MLH_CUDA_CHECK(c, cudaMemsetAsync(p.d.counters, 0, 4 * sizeof(unsigned), st));
    c->grid_host_current = false;
    if (p.periodic) { // fixed grid (MeshlessScheme.cpp:17): the device copy follows the pool
        int rcg = upload_grid(c);
        if (rcg != MLH_OK) return rcg;
    }
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(st));
    p.ncur = (int)N;
    p.n = 0;
    p.own_begin = 0;
    p.own_end = 0;
    c->n_owned = N;
    c->have_state = true;
    c->bbox_valid = false;
    c->phase = 0;
    return MLH_OK;
}

int mlh_build_grid(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (!c->have_state || c->phase != 0) {
        snprintf(c->err, sizeof(c->err), "mlh_build_grid: needs an uploaded/advanced state (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    Params &p = c->p;
    const bool multi = c->cfg.nranks > 1;
    if (!p.periodic) { // MeshlessScheme.cpp:41-51: grid rebuilt from the particle bounding box every step
        // the update kernel of the previous step has already reduced the box of its new positions (k_flux_sum_update);
        // a freshly uploaded state needs the stand-alone pass
        if (!multi && c->bbox_valid && c->max_cells > 0 && p.grid.ncells > 0) {
            // ---- single GPU, inside the time loop: Domain::createGrid runs on the device, no host round trip ----
            c->bbox_valid = false;
            if (c->grid_mirror_pending && cudaEventQuery(c->ev_grid) == cudaSuccess) {
                // the grid of an earlier step has arrived: keep the cell arrays well ahead of it
                p.grid = c->h_grid[1];
                c->grid_mirror_pending = false;
                if ((long long)p.grid.ncells + 2 > (long long)(0.8 * c->max_cells)) {
                    int rcr = reserve_cells(c, 2LL * p.grid.ncells + 2);
                    if (rcr != MLH_OK) return rcr;
                }
            } else {
                cudaGetLastError(); // cudaErrorNotReady is not an error
            }
            int rc = mlh_launch_make_grid(c);
            if (rc != MLH_OK) return rc;
            c->grid_host_current = false;
            if (!c->grid_mirror_pending) {
                MLH_CUDA_CHECK(c, cudaMemcpyAsync(&c->h_grid[1], p.d.grid, sizeof(Grid), cudaMemcpyDeviceToHost, c->stream));
                MLH_CUDA_CHECK(c, cudaEventRecord(c->ev_grid, c->stream));
                c->grid_mirror_pending = true;
            }
        } else {
            int rc = c->bbox_valid ? MLH_OK : mlh_launch_bbox(c);
            if (rc != MLH_OK) return rc;
            c->bbox_valid = false;
            if (multi && (rc = mlh_comm_bbox(c)) != MLH_OK) return rc;
            MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_small, p.d.bbox, 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
            for (int k = 0; k < 6; ++k) { // [0..5] travel as order-preserving keys
                unsigned long long key;
                memcpy(&key, &c->h_small[k], sizeof(key));
                c->h_small[k] = key_dbl(key);
            }
            // quirk Q8: the sequential `if (x<min) .. else if (x>max)` loop never tests a particle that lowers the running
            // minimum against the maximum.  The reduction above (min over all, max over original index >= 1) equals it
            // unless the largest coordinate among the indices >= 1 is itself such a record low -- which, given that it is
            // the largest, can only be original particle 1 sitting below particle 0.
            bool q8 = false;
            for (int k = 0; k < p.D; ++k) q8 = q8 || (c->h_small[6 + k] > c->h_small[3 + k] && c->h_small[9 + k] >= c->h_small[3 + k]);
            if (q8) {
                if (multi) {
                    snprintf(c->err, sizeof(c->err), "getDomainLimits quirk Q8: original particles 0 and 1 are the two largest along an axis -- "
                                                     "the sequential replay is not supported with nranks > 1");
                    return MLH_E_INVALID;
                }
                if ((rc = mlh_launch_bbox_q8_replay(c)) != MLH_OK) return rc;
                MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_small, p.d.bbox, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
                for (int k = 0; k < 6; ++k) {
                    unsigned long long key;
                    memcpy(&key, &c->h_small[k], sizeof(key));
                    c->h_small[k] = key_dbl(key);
                }
            }
            rc = make_grid(c, c->h_small, c->h_small + 3);
            if (rc == MLH_OK) rc = reserve_cells(c, (long long)p.grid.ncells + 2);
            if (rc == MLH_OK) rc = upload_grid(c);
            if (rc != MLH_OK) return rc;
        }
    }
    if (multi) { // exchange 1: boundary layers + migrants (createGhostParticles point of the step)
        int rc = mlh_halo_exchange_particles(c);
        if (rc != MLH_OK) return rc;
    }
    int rc = mlh_launch_sort(c);
    if (rc != MLH_OK) return rc;
    p.own_begin = 0;
    p.own_end = p.n;
    if (multi && (rc = mlh_halo_read_layout(c)) != MLH_OK) return rc;
    c->phase = 1;
    return MLH_OK;
}

int mlh_neighbours(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (c->phase != 1) {
        snprintf(c->err, sizeof(c->err), "mlh_neighbours: call mlh_build_grid first (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    int rc = mlh_launch_neighbours(c);
    if (rc == MLH_OK) rc = mlh_launch_face_index(c);
    if (rc == MLH_OK) c->phase = 2;
    return rc;
}

int mlh_density_matrix(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (c->phase != 2) {
        snprintf(c->err, sizeof(c->err), "mlh_density_matrix: call mlh_neighbours first (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    int rc = mlh_launch_density(c);
    if (rc == MLH_OK && c->cfg.nranks > 1) { // exchange 2 (updateGhostState point, MeshlessScheme.cpp:109)
        Params &p = c->p;
        double *arr[1] = {p.d.pk1}; // x, v, rho, P, cs, omega of the boundary layers, one contiguous range per side
        rc = mlh_halo_refresh(c, arr, 1, MLH_PK1(p.D));
    }
    if (rc == MLH_OK) c->phase = 3;
    return rc;
}

int mlh_gradients_limit(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (c->phase != 3) {
        snprintf(c->err, sizeof(c->err), "mlh_gradients_limit: call mlh_density_matrix first (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    int rc = mlh_launch_gradient(c);
    if (rc == MLH_OK && c->cfg.nranks > 1) { // exchange 3 (updateGhostGradients point, :129) + global dt
        Params &p = c->p;
        double *arr[1] = {p.d.pk2}; // Binv + limited gradients of the boundary layers
        rc = mlh_halo_refresh(c, arr, 1, MLH_PK2(p.D));
        if (rc == MLH_OK) rc = mlh_comm_min_dt(c);
    }
    if (rc == MLH_OK) c->phase = 4;
    return rc;
}

int mlh_timestep(mlh_ctx *c, double *dt_cfl) {
    if (check_ctx(c) != MLH_OK || !dt_cfl) return MLH_E_INVALID;
    if (c->phase != 4) {
        snprintf(c->err, sizeof(c->err), "mlh_timestep: call mlh_gradients_limit first (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_small, c->p.d.dt_bits, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    *dt_cfl = c->h_small[0];
    return MLH_OK;
}

int mlh_flux_update(mlh_ctx *c, double dt) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (c->phase != 4) {
        snprintf(c->err, sizeof(c->err), "mlh_flux_update: call mlh_gradients_limit first (phase %d)", c->phase);
        return MLH_E_STATE;
    }
    if (!(dt >= 0.)) {
        snprintf(c->err, sizeof(c->err), "mlh_flux_update: dt must be >= 0");
        return MLH_E_INVALID;
    }
    int rc = mlh_launch_flux(c, dt, -1.);
    if (rc == MLH_OK) c->phase = 0;
    return rc;
}

int mlh_prepare(mlh_ctx *c, double *dt_cfl) {
    int rc;
    if ((rc = mlh_build_grid(c)) != MLH_OK) return rc;
    if ((rc = mlh_neighbours(c)) != MLH_OK) return rc;
    if ((rc = mlh_density_matrix(c)) != MLH_OK) return rc;
    if ((rc = mlh_gradients_limit(c)) != MLH_OK) return rc;
    if (dt_cfl) return mlh_timestep(c, dt_cfl);
    return MLH_OK;
}

int mlh_advance(mlh_ctx *c, double dt) { return mlh_flux_update(c, dt); }

int mlh_step(mlh_ctx *c, double dt_fixed, double dt_max, double *dt_used) {
    int rc = mlh_prepare(c, nullptr);
    if (rc != MLH_OK) return rc;
    rc = mlh_launch_flux(c, dt_fixed, dt_max);
    if (rc != MLH_OK) return rc;
    c->phase = 0;
    if (dt_used) {
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_small, c->p.d.dt_used, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        *dt_used = c->h_small[0];
    }
    return MLH_OK;
}

long mlh_num_particles(mlh_ctx *c) { return c ? c->n_owned : -1; }

int mlh_grid_info(mlh_ctx *c, int *cells3, double *cell_size3, double *bounds6) {
    if (!c) return MLH_E_INVALID;
    if (!c->grid_host_current && c->pool && c->p.grid.ncells > 0) { // the device built the grid of this step: fetch it
        cudaSetDevice(c->cfg.device);
        if (cudaMemcpyAsync(&c->h_grid[0], c->p.d.grid, sizeof(Grid), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess)
            return MLH_E_CUDA;
        c->p.grid = c->h_grid[0];
        c->grid_host_current = true;
    }
    const Grid &g = c->p.grid;
    for (int k = 0; k < 3; ++k) {
        if (cells3) cells3[k] = g.cells[k];
        if (cell_size3) cell_size3[k] = g.cell_size[k];
    }
    if (bounds6)
        for (int k = 0; k < c->p.D; ++k) {
            bounds6[k] = g.bmin[k];
            bounds6[c->p.D + k] = g.bmax[k];
        }
    return MLH_OK;
}

// which arrays hold the current state
struct StateView {
    const double *x[3], *v[3], *m, *u;
    const int *ids;
    int n;
};
static StateView current_state(mlh_ctx *c) {
    const Params &p = c->p;
    StateView s;
    if (c->phase == 0) {
        for (int k = 0; k < 3; ++k) { s.x[k] = p.d.cx[k]; s.v[k] = p.d.cv[k]; }
        s.m = p.d.cm; s.u = p.d.cu; s.ids = p.d.cid; s.n = p.ncur;
    } else {
        const int o = p.own_begin;
        for (int k = 0; k < 3; ++k) { s.x[k] = p.d.x[k] ? p.d.x[k] + o : nullptr; s.v[k] = p.d.v[k] ? p.d.v[k] + o : nullptr; }
        s.m = p.d.m + o; s.u = p.d.u + o; s.ids = p.d.id + o; s.n = p.own_end - p.own_begin;
    }
    return s;
}

// device array (n values, device order) -> host array; single GPU: scattered to original index
static int fetch_f64(mlh_ctx *c, const double *src, const int *ids, int n, double *host, int comp, int stride, double *tmp) {
    if (c->cfg.nranks > 1) {
        // device order; the caller gets the ids alongside (contiguous per component only)
        if (stride != 1) return MLH_E_INVALID;
        MLH_CUDA_CHECK(c, cudaMemcpyAsync(host, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        return MLH_OK;
    }
    int rc = mlh_launch_unpermute_f64(c, src, ids, tmp, n, comp, stride);
    return rc;
}

int mlh_download_state(mlh_ctx *c, double *x, double *y, double *z, double *vx, double *vy, double *vz, double *m,
                       double *u, int *ids_out) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    if (!c->have_state) {
        snprintf(c->err, sizeof(c->err), "mlh_download_state: no state uploaded");
        return MLH_E_STATE;
    }
    StateView s = current_state(c);
    const int n = s.n;
    // un-permutation scratch: (2D+2) arrays of ncap doubles, allocated once
    if (!c->dl_scratch) MLH_CUDA_CHECK(c, cudaMalloc(&c->dl_scratch, sizeof(double) * (size_t)c->p.ncap * 8));
    double *hx[3] = {x, y, z}, *hv[3] = {vx, vy, vz};
    struct Item { const double *src; double *dst; } items[8];
    int ni = 0;
    for (int k = 0; k < c->p.D; ++k) {
        items[ni++] = {s.x[k], hx[k]};
        items[ni++] = {s.v[k], hv[k]};
    }
    items[ni++] = {s.m, m};
    items[ni++] = {s.u, u};
    int rc = MLH_OK;
    for (int q = 0; q < ni && rc == MLH_OK; ++q) {
        if (!items[q].dst) continue;
        if (c->cfg.nranks > 1) { // device order, ids alongside
            MLH_CUDA_CHECK(c, cudaMemcpyAsync(items[q].dst, items[q].src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        } else {
            double *tmp = c->dl_scratch + (size_t)q * c->p.ncap;
            rc = mlh_launch_unpermute_f64(c, items[q].src, s.ids, tmp, n, 0, 1);
            if (rc == MLH_OK)
                MLH_CUDA_CHECK(c, cudaMemcpyAsync(items[q].dst, tmp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (rc == MLH_OK && ids_out) {
        if (c->cfg.nranks > 1) {
            MLH_CUDA_CHECK(c, cudaMemcpyAsync(ids_out, s.ids, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        } else {
            for (int i = 0; i < n; ++i) ids_out[i] = i;
        }
    }
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return rc;
}

int mlh_download_diag(mlh_ctx *c, double *rho, double *P, double *rhoGrad, int *noi) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    const Params &p = c->p;
    if (p.n == 0) {
        snprintf(c->err, sizeof(c->err), "mlh_download_diag: call mlh_prepare first");
        return MLH_E_STATE;
    }
    const int o = p.own_begin, n = p.own_end - p.own_begin, D = p.D;
    const int *ids = p.d.id + o;
    // un-permutation scratch shared with mlh_download_state (8 x ncap doubles, allocated once): dump2file calls this
    // at every snapshot (MeshlessScheme.cpp:165-195)
    if (!c->dl_scratch) MLH_CUDA_CHECK(c, cudaMalloc(&c->dl_scratch, sizeof(double) * (size_t)c->p.ncap * 8));
    double *tmp = c->dl_scratch;
    int rc = MLH_OK;
    auto one = [&](const double *src, double *dst) {
        if (!dst || rc != MLH_OK) return;
        if (c->cfg.nranks > 1) {
            cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        } else {
            rc = mlh_launch_unpermute_f64(c, src, ids, tmp, n, 0, 1);
            cudaMemcpyAsync(dst, tmp, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        }
        cudaStreamSynchronize(c->stream);
    };
    one(p.d.rho + o, rho);
    one(p.d.P + o, P);
    if (rhoGrad && rc == MLH_OK) {
        if (c->cfg.nranks > 1) { // device order like the state (ids from mlh_download_state), interleaved on the host
            std::vector<double> comp((size_t)n);
            for (int a = 0; a < D; ++a) {
                cudaMemcpyAsync(comp.data(), p.d.g[0 * 3 + a] + o, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
                cudaStreamSynchronize(c->stream);
                for (int i = 0; i < n; ++i) rhoGrad[(size_t)i * D + a] = comp[(size_t)i];
            }
        } else {
            for (int a = 0; a < D && rc == MLH_OK; ++a) rc = mlh_launch_unpermute_f64(c, p.d.g[0 * 3 + a] + o, ids, tmp, n, a, D);
            cudaMemcpyAsync(rhoGrad, tmp, sizeof(double) * (size_t)n * D, cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
        }
    }
    if (noi && rc == MLH_OK) {
        int *itmp = (int *)tmp;
        if (c->cfg.nranks > 1) {
            cudaMemcpyAsync(noi, p.d.noi + o, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        } else {
            rc = mlh_launch_unpermute_i32(c, p.d.noi + o, ids, itmp, n);
            cudaMemcpyAsync(noi, itmp, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        }
        cudaStreamSynchronize(c->stream);
    }
    return rc;
}

int mlh_sums(mlh_ctx *c, double *out6) {
    if (check_ctx(c) != MLH_OK || !out6) return MLH_E_INVALID;
    if (!c->have_state) return MLH_E_STATE;
    int rc = mlh_launch_sums(c);
    if (rc != MLH_OK) return rc;
    if (c->cfg.nranks > 1 && (rc = mlh_comm_sum(c, c->p.d.sums, 6)) != MLH_OK) return rc;
    MLH_CUDA_CHECK(c, cudaMemcpyAsync(c->h_small, c->p.d.sums, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 6; ++k) out6[k] = c->h_small[k];
    return MLH_OK;
}

void *mlh_stream(mlh_ctx *c) { return c ? (void *)c->stream : nullptr; }

int mlh_synchronize(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    MLH_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return MLH_OK;
}

int mlh_profile_enable(mlh_ctx *c, int on) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    prof_flush(c);
    c->profiling = on != 0;
    if (on) {
        for (int k = 0; k < KID_COUNT; ++k) {
            c->prof_ms[k] = 0.;
            c->prof_launches[k] = 0;
        }
    }
    return MLH_OK;
}

int mlh_profile_read(mlh_ctx *c, int max_entries, const char **names, double *total_ms, long *launches) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    prof_flush(c);
    int n = KID_COUNT < max_entries ? KID_COUNT : max_entries;
    for (int k = 0; k < n; ++k) {
        if (names) names[k] = kKernelNames[k];
        if (total_ms) total_ms[k] = c->prof_ms[k];
        if (launches) launches[k] = c->prof_launches[k];
    }
    return n;
}

long mlh_launch_count(mlh_ctx *c) { return c ? c->launches : -1; }

int mlh_timer_start(mlh_ctx *c) {
    if (check_ctx(c) != MLH_OK) return MLH_E_INVALID;
    MLH_CUDA_CHECK(c, cudaEventRecord(c->timer[0], c->stream));
    return MLH_OK;
}
int mlh_timer_stop(mlh_ctx *c, double *ms) {
    if (check_ctx(c) != MLH_OK || !ms) return MLH_E_INVALID;
    MLH_CUDA_CHECK(c, cudaEventRecord(c->timer[1], c->stream));
    MLH_CUDA_CHECK(c, cudaEventSynchronize(c->timer[1]));
    float f = 0.f;
    MLH_CUDA_CHECK(c, cudaEventElapsedTime(&f, c->timer[0], c->timer[1]));
    *ms = f;
    return MLH_OK;
}

int mlh_slab_range(int n_layers, int nranks, int rank, int *lo, int *hi) {
    if (n_layers <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || !lo || !hi) return MLH_E_INVALID;
    // equal cell-layer counts, remainder to the lowest ranks
    int base = n_layers / nranks, rem = n_layers % nranks;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
    return MLH_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// parity harness
// ------------------------------------------------------------------------------------------------
namespace {
// row `id` of out = original ids (or codes) of the list entries of the particle with that original id
__global__ void k_export_lists(const Params p, int which, int *out, int n_rows) {
    int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.own_end) return;
    int row = p.d.id[i];
    if (row < 0 || row >= n_rows) return;
    int nreg = p.d.noi[i], ng = p.d.noig[i];
    for (int s = 0; s < p.max_ni; ++s) {
        int val = -1;
        if (which == 0) {
            if (s < nreg) val = p.d.id[p.d.nnl[(size_t)s * p.ncap + i] & MLH_NNL_IDX_MASK];
        } else if (s < ng) {
            int e = p.d.nnl[(size_t)(nreg + s) * p.ncap + i];
            val = which == 1 ? p.d.id[e & MLH_NNL_IDX_MASK] : (int)((unsigned)e >> MLH_NNL_IDX_BITS);
        }
        out[(size_t)row * p.max_ni + s] = val;
    }
}
// per face: ORIGINAL ids of the canonical endpoint a (lower id: the one that solves the face in the reference,
// Particles.cpp:1841,1889) and of its partner b, and the periodic-image code of b as a lists it
__global__ void k_export_face_pairs(const Params p, int nfaces, int *out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    const int fav = p.d.fa[f], e = p.d.fe[f];
    const int i = fav & 0x7FFFFFFF, j = e & MLH_NNL_IDX_MASK;
    const int code = (int)((unsigned)e >> MLH_NNL_IDX_BITS);
    const bool canon = fav >= 0;
    out[3 * f + 0] = p.d.id[canon ? i : j];
    out[3 * f + 1] = p.d.id[canon ? j : i];
    out[3 * f + 2] = canon ? code : reverse_code(code);
}
__global__ void k_export_face_flux(const Params p, int nfaces, double *out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    for (int nu = 0; nu < p.D + 2; ++nu) out[(size_t)f * (p.D + 2) + nu] = p.d.F[(size_t)f * MLH_FREC(p.D) + nu];
}
// Particles::checkFluxSymmetry (Particles.cpp:2888-2976) for the unique-face layout: the reference compares Fij + Fji of
// every slot pair against FLUX_SYM_TOL; here both endpoints read ONE stored flux, so the check is structural -- every
// slot must map to a face whose record names this particle (as owner or as the owner's list entry) and the two
// endpoints must add it with opposite signs.  out[0] slots with a face, [1] violations, [2] faces used from both
// sides, [3] faces only their owner uses (partner on another rank, or one-sided seam pair, quirk Q9: where the
// reference prints "fluxes are NOT symmetric").
__global__ void k_check_flux_symmetry(const Params p, unsigned long long *out) {
    const int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned slots = 0, bad = 0, partner = 0, owned = 0;
    if (i < p.own_end) {
        const int ntot = p.d.noi[i] + p.d.noig[i];
        for (int s = 0; s < ntot; ++s) {
            const size_t at = (size_t)s * p.ncap + i;
            const unsigned v = p.d.fmap[at];
            if (v == MLH_FMAP_SKIP) continue;
            const int f = (int)(v >> 2);
            if (f >= p.fcap) continue;
            ++slots;
            const int fav = p.d.fa[f], e = p.d.fe[f];
            const int owner = fav & 0x7FFFFFFF, j = p.d.nnl[at] & MLH_NNL_IDX_MASK;
            if (v & 2u) { // this particle owns the face
                ++owned;
                if (owner != i || (e & MLH_NNL_IDX_MASK) != j || ((unsigned)fav >> 31) != (v & 1u)) ++bad;
            } else {      // the partner owns it: it must name this particle, and add +F where this one adds -F
                ++partner;
                if (owner != j || (e & MLH_NNL_IDX_MASK) != i || !(v & 1u) || ((unsigned)fav >> 31) != 0u) ++bad;
            }
        }
    }
    atomicAdd(out + 0, (unsigned long long)slots);
    atomicAdd(out + 1, (unsigned long long)bad);
    atomicAdd(out + 2, (unsigned long long)partner);
    atomicAdd(out + 3, (unsigned long long)(owned));
}
__global__ void k_iota_sorted_index(const Params p, int *out) {
    int i = p.own_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.own_end) return;
    out[p.d.id[i]] = i;
}
} // namespace

extern "C" long mlh_debug_fetch(mlh_ctx *c, const char *field, void *dst, long dst_elems) {
    if (check_ctx(c) != MLH_OK || !field) return MLH_E_INVALID;
    const Params &p = c->p;
    const int D = p.D;
    const std::string f(field);
    if (c->cfg.nranks > 1) {
        // sharded runs: only per-rank statistics (device order); the per-particle harness is single-GPU
        if (f == "noi" || f == "noiGhosts") {
            const int n = p.own_end - p.own_begin;
            if (p.n == 0) return MLH_E_STATE;
            if (!dst) return n;
            if (dst_elems < n) return MLH_E_INVALID;
            cudaMemcpyAsync(dst, (f == "noi" ? p.d.noi : p.d.noig) + p.own_begin, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
            return n;
        }
        if (f != "num_faces" && f != "counters" && f != "one_sided_pairs" && f != "flux_symmetry") {
            snprintf(c->err, sizeof(c->err), "mlh_debug_fetch(%s) is single-GPU only", field);
            return MLH_E_INVALID;
        }
    }
    StateView sv = current_state(c);
    // ---- state ----
    const double *src = nullptr;
    const int *ids = sv.ids;
    int n = sv.n, comps = 1;
    const double *multi[15] = {nullptr};
    if (f == "x") src = sv.x[0];
    else if (f == "y") src = sv.x[1];
    else if (f == "z" && D == 3) src = sv.x[2];
    else if (f == "vx") src = sv.v[0];
    else if (f == "vy") src = sv.v[1];
    else if (f == "vz" && D == 3) src = sv.v[2];
    else if (f == "m") src = sv.m;
    else if (f == "u") src = sv.u;
    if (!src) {
        // derived quantities live in SRT order
        if (p.n == 0) {
            snprintf(c->err, sizeof(c->err), "mlh_debug_fetch(%s): nothing computed yet", field);
            return MLH_E_STATE;
        }
        ids = p.d.id + p.own_begin;
        n = p.own_end - p.own_begin;
        const int o = p.own_begin;
        if (f == "rho") src = p.d.rho + o;
        else if (f == "P") src = p.d.P + o;
        else if (f == "omega") src = p.d.omega + o;
        else if (f == "cs") src = p.d.cs + o;
        else if (f == "Binv") { comps = D * D; for (int k = 0; k < comps; ++k) multi[k] = p.d.B[k] + o; }
        else if (f == "rhoGrad") { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.g[0 * 3 + a] + o; }
        else if (f == "vxGrad") { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.g[1 * 3 + a] + o; }
        else if (f == "vyGrad") { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.g[2 * 3 + a] + o; }
        else if (f == "vzGrad" && D == 3) { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.g[3 * 3 + a] + o; }
        else if (f == "PGrad") { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.g[4 * 3 + a] + o; }
        else if (f == "vF" && p.debug_capture) { comps = D; for (int a = 0; a < D; ++a) multi[a] = p.d.flux[2 + a] + o; }
        else if (f == "mF" && p.debug_capture) src = p.d.flux[0] + o;
        else if (f == "eF" && p.debug_capture) src = p.d.flux[1] + o;
    }
    if (src || multi[0]) {
        long count = (long)n * comps;
        if (!dst) return count;
        if (dst_elems < count) return MLH_E_INVALID;
        double *tmp = nullptr;
        if (cudaMalloc(&tmp, sizeof(double) * (size_t)count) != cudaSuccess) return MLH_E_CUDA;
        int rc = MLH_OK;
        if (src) rc = mlh_launch_unpermute_f64(c, src, ids, tmp, n, 0, 1);
        else
            for (int k = 0; k < comps && rc == MLH_OK; ++k) rc = mlh_launch_unpermute_f64(c, multi[k], ids, tmp, n, k, comps);
        cudaMemcpyAsync(dst, tmp, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        return rc == MLH_OK ? count : rc;
    }
    if (p.n == 0) return MLH_E_STATE;
    const int o = p.own_begin;
    n = p.own_end - p.own_begin;
    ids = p.d.id + o;
    if (f == "gradPre" && p.debug_capture) {
        long count = (long)(D + 2) * n * D;
        if (!dst) return count;
        if (dst_elems < count) return MLH_E_INVALID;
        double *tmp = nullptr;
        if (cudaMalloc(&tmp, sizeof(double) * (size_t)count) != cudaSuccess) return MLH_E_CUDA;
        int blk = 0, rc = MLH_OK;
        for (int fl = 0; fl < 5; ++fl) {
            if (fl == 3 && D == 2) continue;
            for (int a = 0; a < D && rc == MLH_OK; ++a)
                rc = mlh_launch_unpermute_f64(c, p.d.gpre[fl * 3 + a] + o, ids, tmp + (size_t)blk * n * D, n, a, D);
            ++blk;
        }
        cudaMemcpyAsync(dst, tmp, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        return rc == MLH_OK ? count : rc;
    }
    // ---- integers ----
    const int *isrc = nullptr;
    if (f == "cell") isrc = p.d.cell + o;
    else if (f == "noi") isrc = p.d.noi + o;
    else if (f == "noiGhosts") isrc = p.d.noig + o;
    if (isrc || f == "sorted_index") {
        if (!dst) return n;
        if (dst_elems < n) return MLH_E_INVALID;
        int *tmp = nullptr;
        if (cudaMalloc(&tmp, sizeof(int) * (size_t)n) != cudaSuccess) return MLH_E_CUDA;
        int rc = MLH_OK;
        if (isrc) rc = mlh_launch_unpermute_i32(c, isrc, ids, tmp, n);
        else k_iota_sorted_index<<<mlh_blocks(n, 256), 256, 0, c->stream>>>(p, tmp);
        cudaMemcpyAsync(dst, tmp, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        return rc == MLH_OK ? n : rc;
    }
    int which = f == "nnl" ? 0 : (f == "nnlGhosts" ? 1 : (f == "nnlGhostCodes" ? 2 : -1));
    if (which >= 0) {
        long count = (long)n * p.max_ni;
        if (!dst) return count;
        if (dst_elems < count) return MLH_E_INVALID;
        int *tmp = nullptr;
        if (cudaMalloc(&tmp, sizeof(int) * (size_t)count) != cudaSuccess) return MLH_E_CUDA;
        k_export_lists<<<mlh_blocks(n, 128), 128, 0, c->stream>>>(p, which, tmp, n);
        cudaMemcpyAsync(dst, tmp, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        return count;
    }
    if (f == "face_pairs" || f == "face_rec" || f == "face_F") { // per-face intermediates (a16/a17/a19): single GPU
        int nf = 0;
        cudaMemcpyAsync(&nf, p.d.face_start + p.own_end, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        if (nf > p.fcap) nf = p.fcap;
        const int per = f == "face_pairs" ? 3 : (f == "face_rec" ? 4 * D + 4 : D + 2);
        const long count = (long)nf * per;
        if (!dst) return count;
        if (dst_elems < count) return MLH_E_INVALID;
        if (f == "face_rec") {
            if (!p.debug_capture || c->phase != 0) {
                snprintf(c->err, sizeof(c->err), "mlh_debug_fetch(face_rec): needs debug_capture and a completed flux pass");
                return MLH_E_STATE;
            }
            cudaMemcpyAsync(dst, p.d.dbg_face, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, c->stream);
            cudaStreamSynchronize(c->stream);
            return count;
        }
        void *tmp = nullptr;
        const size_t bytes = (size_t)count * (f == "face_pairs" ? sizeof(int) : sizeof(double));
        if (count == 0) return 0;
        if (cudaMalloc(&tmp, bytes) != cudaSuccess) return MLH_E_CUDA;
        if (f == "face_pairs")
            k_export_face_pairs<<<mlh_blocks(nf, 256), 256, 0, c->stream>>>(p, nf, (int *)tmp);
        else
            k_export_face_flux<<<mlh_blocks(nf, 256), 256, 0, c->stream>>>(p, nf, (double *)tmp);
        cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        return count;
    }
    if (f == "flux_symmetry") { // Particles::checkFluxSymmetry, structural form (see k_check_flux_symmetry)
        if (!dst) return 4;
        if (dst_elems < 4) return MLH_E_INVALID;
        unsigned long long *tmp = nullptr;
        if (cudaMalloc(&tmp, 4 * sizeof(unsigned long long)) != cudaSuccess) return MLH_E_CUDA;
        cudaMemsetAsync(tmp, 0, 4 * sizeof(unsigned long long), c->stream);
        k_check_flux_symmetry<<<mlh_blocks(n > 0 ? n : 1, 128), 128, 0, c->stream>>>(p, tmp);
        unsigned long long h[4];
        cudaMemcpyAsync(h, tmp, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(tmp);
        int *o4 = (int *)dst;
        o4[0] = (int)std::min<unsigned long long>(h[0], 2147483647ull);
        o4[1] = (int)std::min<unsigned long long>(h[1], 2147483647ull);
        o4[2] = (int)std::min<unsigned long long>(h[2], 2147483647ull);
        o4[3] = (int)std::min<unsigned long long>(h[3] - h[2], 2147483647ull); // owned faces without a partner slot
        return 4;
    }
    if (f == "num_faces") { // faces of this step's list (k_face_index)
        if (!dst) return 1;
        if (dst_elems < 1) return MLH_E_INVALID;
        cudaMemcpyAsync(dst, p.d.face_start + p.own_end, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        return 1;
    }
    if (f == "one_sided_pairs" || f == "counters") {
        if (!dst) return 4;
        if (dst_elems < 4) return MLH_E_INVALID;
        cudaMemcpyAsync(dst, p.d.counters, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
        return 4;
    }
    snprintf(c->err, sizeof(c->err), "mlh_debug_fetch: unknown field '%s'", field);
    return MLH_E_INVALID;
}
