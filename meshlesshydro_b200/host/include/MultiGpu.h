// Multi-GPU launcher of the host layer: one process per GPU on ONE node (north_star: "slab-decomposed across the GPUs
// of one 8xB200 box"), no MPI.  The reference is a serial program (no MPI_/omp call sites); this
// is the piece a C++ user needs instead of torchrun + meshlesshydro_b200/multigpu.py:
//
//   * the parent reads the config and the initial distribution into host arrays that live in an anonymous
//     MAP_SHARED mapping, then forks ranks 1..R-1 (no CUDA call has happened yet); the parent is rank 0;
//   * every rank creates its own context (device = rank), rank 0 makes the NCCL unique id and hands it over through
//     the shared mapping, all call mlh_comm_init;
//   * every rank uploads the particles of its slab (same cell-layer rule as the device: Particles.cpp:279-302 on the
//     slowest axis, mlh_slab_range) and runs the SAME MeshlessScheme::run() loop -- dt and the conservation sums are
//     global values, so all ranks take the same branches;
//   * at a snapshot every rank downloads its owned particles and scatters them by original id into the shared host
//     arrays; after a process-shared barrier rank 0 writes the one HDF5 file the reference writes.
#ifndef DEMONSTRATOR_MULTIGPU_H
#define DEMONSTRATOR_MULTIGPU_H

#include <cstddef>

namespace mgpu {
void plan(int nranks);          // before any Particles is constructed; nranks <= 1 keeps everything single-process
int planned();                  // ranks asked for (>= 1)
void *sharedAlloc(size_t bytes); // zeroed MAP_SHARED | MAP_ANONYMOUS memory (valid in all ranks if called before launch)
void sharedFree(void *p, size_t bytes);
void launch();                  // fork ranks 1..R-1; afterwards rank() / nranks() are final
int rank();
int nranks();
bool active();                  // launched with more than one rank
void barrier();                 // all ranks
char *ncclId();                 // 128 bytes in the shared mapping
void fail();                    // mark the run as failed (other ranks abort at their next barrier)
int finish(int rc);             // rank 0: wait for the other ranks, return the worst exit code; others: exit
} // namespace mgpu

#endif // DEMONSTRATOR_MULTIGPU_H
