// Multi-GPU slab decomposition: NCCL halo exchange (placeholder until the exchange kernels land).
#include "mlh_internal.cuh"

extern "C" int mlh_comm_unique_id(char *id128) {
    (void)id128;
    return MLH_E_COMM;
}
extern "C" int mlh_comm_init(mlh_ctx *c, const char *id128) {
    (void)id128;
    if (c) snprintf(c->err, sizeof(c->err), "multi-GPU exchange not built into this library yet");
    return MLH_E_COMM;
}
