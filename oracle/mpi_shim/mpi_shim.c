/* oracle/mpi_shim/mpi_shim.c -- TEST INFRASTRUCTURE, not product code.
 * Thread-rank implementation of the MPI subset declared in mpi.h (see that header for
 * the reference call sites it serves). One process, `size` threads, shared address space:
 * an RMA window is the table of every rank's base pointer; MPI_Get is a memcpy. */
#define _GNU_SOURCE
#include "mpi.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct shim_group {
    int size;
    pthread_barrier_t bar;
    struct shim_comm *comms;  /* [size] */
    const void **slots;       /* [size] scratch for collectives */
};

struct shim_win {
    int size;
    char **base;       /* [size] */
    MPI_Aint *bytes;   /* [size] */
    int *disp_unit;    /* [size] */
    struct shim_group *group;
    int freed_count;
    pthread_mutex_t mu;
};

static struct shim_group *self_group(void) {
    static struct shim_group *g = NULL;
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_mutex_lock(&mu);
    if (!g) g = shim_group_create(1);
    pthread_mutex_unlock(&mu);
    return g;
}

MPI_Comm shim_comm_self(void) { return &self_group()->comms[0]; }

struct shim_group *shim_group_create(int size) {
    struct shim_group *g = (struct shim_group *)calloc(1, sizeof(*g));
    g->size = size;
    pthread_barrier_init(&g->bar, NULL, (unsigned)size);
    g->comms = (struct shim_comm *)calloc((size_t)size, sizeof(struct shim_comm));
    g->slots = (const void **)calloc((size_t)size, sizeof(void *));
    for (int r = 0; r < size; r++) {
        g->comms[r].group = g;
        g->comms[r].rank = r;
    }
    return g;
}

MPI_Comm shim_group_comm(struct shim_group *g, int rank) { return &g->comms[rank]; }

void shim_group_destroy(struct shim_group *g) {
    if (!g) return;
    pthread_barrier_destroy(&g->bar);
    free(g->comms);
    free((void *)g->slots);
    free(g);
}

static void group_barrier(struct shim_group *g) {
    if (g->size > 1) pthread_barrier_wait(&g->bar);
}

int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm->group->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm->rank; return MPI_SUCCESS; }

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr) {
    (void)info;
    void *p = NULL;
    size_t n = size > 0 ? (size_t)size : 64;
    if (posix_memalign(&p, 64, n) != 0) return 1;
    *(void **)baseptr = p;
    return MPI_SUCCESS;
}

int MPI_Free_mem(void *base) { free(base); return MPI_SUCCESS; }

int MPI_Win_create(void *base, MPI_Aint size, int disp_unit, MPI_Info info, MPI_Comm comm, MPI_Win *win) {
    (void)info;
    struct shim_group *g = comm->group;
    struct shim_win *w = NULL;
    if (comm->rank == 0) {
        w = (struct shim_win *)calloc(1, sizeof(*w));
        w->size = g->size;
        w->base = (char **)calloc((size_t)g->size, sizeof(char *));
        w->bytes = (MPI_Aint *)calloc((size_t)g->size, sizeof(MPI_Aint));
        w->disp_unit = (int *)calloc((size_t)g->size, sizeof(int));
        w->group = g;
        pthread_mutex_init(&w->mu, NULL);
        g->slots[0] = w;
    }
    group_barrier(g);
    w = (struct shim_win *)g->slots[0];
    w->base[comm->rank] = (char *)base;
    w->bytes[comm->rank] = size;
    w->disp_unit[comm->rank] = disp_unit;
    group_barrier(g);
    *win = w;
    return MPI_SUCCESS;
}

int MPI_Win_free(MPI_Win *win) {
    /* Called by every rank (possibly sequentially from one thread at teardown): the last
     * caller releases the table. Not collective-blocking on purpose. */
    struct shim_win *w = *win;
    if (!w) return MPI_SUCCESS;
    pthread_mutex_lock(&w->mu);
    int last = (++w->freed_count == w->size);
    pthread_mutex_unlock(&w->mu);
    if (last) {
        free(w->base);
        free(w->bytes);
        free(w->disp_unit);
        pthread_mutex_destroy(&w->mu);
        free(w);
    }
    *win = NULL;
    return MPI_SUCCESS;
}

int MPI_Win_fence(int assert_, MPI_Win win) {
    (void)assert_;
    group_barrier(win->group);
    return MPI_SUCCESS;
}

int MPI_Win_lock(int lock_type, int rank, int assert_, MPI_Win win) {
    (void)lock_type; (void)rank; (void)assert_; (void)win;
    return MPI_SUCCESS; /* shared lock on read-only data: nothing to do (optimistic) */
}

int MPI_Win_unlock(int rank, MPI_Win win) {
    (void)rank; (void)win;
    return MPI_SUCCESS;
}

static size_t type_bytes(MPI_Datatype t) { return (size_t)t; /* codes are the byte widths */ }

int MPI_Get(void *origin_addr, int origin_count, MPI_Datatype origin_datatype, int target_rank,
            MPI_Aint target_disp, int target_count, MPI_Datatype target_datatype, MPI_Win win) {
    (void)target_count; (void)target_datatype;
    const char *src = win->base[target_rank] + (size_t)target_disp * (size_t)win->disp_unit[target_rank];
    memcpy(origin_addr, src, (size_t)origin_count * type_bytes(origin_datatype));
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm) {
    (void)recvcount; (void)recvtype;
    struct shim_group *g = comm->group;
    size_t n = (size_t)sendcount * type_bytes(sendtype);
    g->slots[comm->rank] = sendbuf;
    group_barrier(g);
    for (int r = 0; r < g->size; r++) memcpy((char *)recvbuf + (size_t)r * n, g->slots[r], n);
    group_barrier(g);
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op, MPI_Comm comm) {
    struct shim_group *g = comm->group;
    if (datatype != MPI_INT || op != MPI_MAX) {
        fprintf(stderr, "mpi_shim: only MPI_Allreduce(MPI_INT, MPI_MAX) is implemented\n");
        abort();
    }
    g->slots[comm->rank] = sendbuf;
    group_barrier(g);
    for (int i = 0; i < count; i++) {
        int m = ((const int *)g->slots[0])[i];
        for (int r = 1; r < g->size; r++) {
            int v = ((const int *)g->slots[r])[i];
            if (v > m) m = v;
        }
        ((int *)recvbuf)[i] = m;
    }
    group_barrier(g);
    return MPI_SUCCESS;
}

int MPI_Finalized(int *flag) { *flag = 0; return MPI_SUCCESS; }
