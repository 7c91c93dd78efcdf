// kmcb200_mpi_compat.h -- the handful of MPI names the reference's superstep (src/kmc_main.cpp:161-171,184-240,328-545)
// and its solver entry points use, for builds WITHOUT <mpi.h> (this image has no MPI; the reference runs one MPI rank per
// GPU, src/kmc_main.cpp:72-101).  A host that has a real MPI defines KMCB200_HAVE_MPI before including gpu_solvers_b200.hpp
// and none of this is seen.
//
// Model: one process per GPU.  MPI_Comm is a pointer to a small rank/size record; MPI_COMM_WORLD is the process's world
// record, filled by kmcb200_world_init() (rank / size from the launcher's environment: RANK / WORLD_SIZE as set by
// torchrun, or KMCB200_RANK / KMCB200_WORLD_SIZE).  The data-moving collectives of the superstep are NOT emulated here:
// in the B200 build the potentials are exchanged inside sum_and_gather_potential (NVLink peer memory), so the two
// MPI_Gatherv calls of the superstep (device pointers, root 0) are synchronisation points only.
#pragma once
#ifndef KMCB200_HAVE_MPI

#include <chrono>
#include <cstdlib>

struct kmcb200_mpi_comm {
    int rank = 0, size = 1;
};
typedef kmcb200_mpi_comm *MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Request;

#define MPI_COMM_NULL ((MPI_Comm) nullptr)
#define MPI_IN_PLACE ((void *)1)
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_SUCCESS 0

inline kmcb200_mpi_comm *kmcb200_world() {
    static kmcb200_mpi_comm world;
    return &world;
}
#define MPI_COMM_WORLD (kmcb200_world())

// rank / size of this process from the launcher's environment (torchrun: RANK / WORLD_SIZE)
inline void kmcb200_world_init() {
    const char *r = std::getenv("KMCB200_RANK") ? std::getenv("KMCB200_RANK") : std::getenv("RANK");
    const char *s = std::getenv("KMCB200_WORLD_SIZE") ? std::getenv("KMCB200_WORLD_SIZE") : std::getenv("WORLD_SIZE");
    kmcb200_world()->rank = r ? std::atoi(r) : 0;
    kmcb200_world()->size = s ? std::atoi(s) : 1;
}
inline int MPI_Init(int *, char ***) { kmcb200_world_init(); return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm c, int *rank) { *rank = c ? c->rank : 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm c, int *size) { *size = c ? c->size : 1; return MPI_SUCCESS; }
inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// host-side barrier of the processes of `c`: provided by the library's bootstrap when size > 1 (gpu_solvers_b200.hpp)
int kmcb200_host_barrier(MPI_Comm c);
inline int MPI_Barrier(MPI_Comm c) { return (c && c->size > 1) ? kmcb200_host_barrier(c) : MPI_SUCCESS; }
// root-0 gathers of DEVICE buffers in the superstep: synchronisation only (see the header comment).  (The root's
// "MPI_IN_PLACE, NULL, NULL" form binds NULL to the int parameters, as it does with a real mpi.h.)
inline int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm c) {
    return MPI_Barrier(c);
}
#endif  // KMCB200_HAVE_MPI
