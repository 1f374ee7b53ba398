/* oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A minimal in-process stand-in for the 13 MPI entry points that the reference's
 * DDStore method-0 path touches, so that /root/reference/include/ddstore.hpp and
 * /root/reference/src/ddstore.cxx can be compiled VERBATIM (no MPI in this image).
 *
 * Model: one "rank" == one thread of a single process. A communicator is a pointer to
 * {group, rank}; collectives are pthread-barrier based; an RMA window is the table of
 * the ranks' base pointers, and MPI_Get is a memcpy out of the target's buffer -- i.e.
 * what an intra-node MPI does for a shared-memory window, minus lock/flush cost
 * (an OPTIMISTIC stand-in).
 *
 * Call sites served (reference file:line):
 *   MPI_Comm_size/rank   src/ddstore.cxx:22-23,29-30,37-38
 *   MPI_Alloc_mem        include/ddstore.hpp:44,115
 *   MPI_Win_create       include/ddstore.hpp:56-61,127-132
 *   MPI_Allgather        include/ddstore.hpp:76,147
 *   MPI_Allreduce        include/ddstore.hpp:80,151
 *   MPI_Win_lock/Get/unlock  include/ddstore.hpp:222-237
 *   MPI_Win_fence        src/ddstore.cxx:59,73
 *   MPI_Finalized/Win_free   src/ddstore.cxx:82,89
 */
#ifndef DDS_ORACLE_MPI_SHIM_H
#define DDS_ORACLE_MPI_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct shim_group;
struct shim_comm { struct shim_group *group; int rank; };
struct shim_win;

typedef struct shim_comm *MPI_Comm;
typedef struct shim_win *MPI_Win;
typedef long MPI_Aint;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Info;

#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_MAX 100
#define MPI_LOCK_SHARED 2

MPI_Comm shim_comm_self(void);
#define MPI_COMM_SELF (shim_comm_self())

/* shim-only: build / destroy a world of `size` thread-ranks */
struct shim_group *shim_group_create(int size);
MPI_Comm shim_group_comm(struct shim_group *g, int rank);
void shim_group_destroy(struct shim_group *g);

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Free_mem(void *base);
int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win);
int MPI_Win_free(MPI_Win *win);
int MPI_Win_fence(int assert_, MPI_Win win);
int MPI_Win_lock(int lock_type, int rank, int assert_, MPI_Win win);
int MPI_Win_unlock(int rank, MPI_Win win);
int MPI_Get(void *origin_addr, int origin_count, MPI_Datatype origin_datatype, int target_rank,
            MPI_Aint target_disp, int target_count, MPI_Datatype target_datatype, MPI_Win win);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Finalized(int *flag);

#ifdef __cplusplus
}
#endif
#endif
