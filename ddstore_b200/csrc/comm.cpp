// ddstore_b200/csrc/comm.cpp -- communicators for the store's two collectives (bootstrap all-gather,
// fence barrier). Replaces the MPI_Comm the reference is constructed with
// (/root/reference/include/ddstore.hpp:29-31, MPI_Allgather/Allreduce at :76,:80, MPI_Win_fence at
// src/ddstore.cxx:59,73). Three kinds: self, POSIX-shm (ranks = processes or threads of one box, the
// NVSwitch domain), and user callbacks (mpi4py / torch.distributed adapters live in Python).
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ddstore_b200.h"
#include "internal.h"

namespace {

constexpr uint32_t kMagic = 0xDD5B200u;
constexpr size_t kSlotBytes = 4096;

struct ShmHeader {
    std::atomic<uint32_t> magic;
    uint32_t size;
    std::atomic<uint32_t> attached;
    std::atomic<uint32_t> bar_count;
    std::atomic<uint32_t> bar_sense;
    uint32_t pad[11];
};
static_assert(sizeof(ShmHeader) == 64, "header is one cache line");

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double comm_timeout_s() {
    if (const char *e = getenv("DDS_COMM_TIMEOUT_S")) return atof(e) > 0 ? atof(e) : 120.0;
    return 120.0;
}

void backoff(unsigned spins) {
    if (spins < 64)
        sched_yield();
    else {
        timespec ts = {0, 50000};
        nanosleep(&ts, nullptr);
    }
}

} // namespace

struct dds_comm {
    int kind; // 0 self, 1 shm, 2 callbacks
    int rank, size;
    // shm
    ShmHeader *hdr = nullptr;
    unsigned char *slots = nullptr;
    size_t map_bytes = 0;
    uint32_t sense = 0;
    std::string shm_name;
    // callbacks
    dds_allgather_fn ag = nullptr;
    dds_barrier_fn bar = nullptr;
    void *ctx = nullptr;
};

static int shm_barrier(dds_comm *c) {
    if (c->size == 1) return DDS_OK;
    c->sense ^= 1u;
    ShmHeader *h = c->hdr;
    if (h->bar_count.fetch_add(1, std::memory_order_acq_rel) == (uint32_t)c->size - 1) {
        h->bar_count.store(0, std::memory_order_relaxed);
        h->bar_sense.store(c->sense, std::memory_order_release);
        return DDS_OK;
    }
    const double t0 = now_s(), limit = comm_timeout_s();
    unsigned spins = 0;
    while (h->bar_sense.load(std::memory_order_acquire) != c->sense) {
        backoff(spins++);
        if ((spins & 1023u) == 0 && now_s() - t0 > limit)
            return dds_internal::fail(DDS_ERR_COMM, "shm communicator: barrier timed out (a rank died or never arrived)");
    }
    return DDS_OK;
}

extern "C" {

dds_comm_t *dds_comm_self(void) {
    dds_comm *c = new dds_comm;
    c->kind = 0;
    c->rank = 0;
    c->size = 1;
    return c;
}

dds_comm_t *dds_comm_shm(const char *key, int rank, int size) {
    if (!key || size < 1 || rank < 0 || rank >= size) {
        dds_internal::fail(DDS_ERR_ARG, "dds_comm_shm: bad key/rank/size");
        return nullptr;
    }
    std::string name = "/dds_b200_";
    for (const char *p = key; *p; p++) name += (isalnum((unsigned char)*p) || *p == '_' || *p == '-') ? *p : '_';
    const size_t bytes = sizeof(ShmHeader) + (size_t)size * kSlotBytes;
    const double t0 = now_s(), limit = comm_timeout_s();

    int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    bool creator = fd >= 0;
    if (creator) {
        if (ftruncate(fd, (off_t)bytes) != 0) {
            close(fd);
            shm_unlink(name.c_str());
            dds_internal::fail(DDS_ERR_COMM, "dds_comm_shm: ftruncate failed");
            return nullptr;
        }
    } else {
        unsigned spins = 0;
        while (true) {
            fd = shm_open(name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break;
                close(fd);
                fd = -1;
            }
            backoff(spins++);
            if (now_s() - t0 > limit) {
                dds_internal::fail(DDS_ERR_COMM, "dds_comm_shm: timed out waiting for the segment to appear");
                return nullptr;
            }
        }
    }
    void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) {
        if (creator) shm_unlink(name.c_str());
        dds_internal::fail(DDS_ERR_COMM, "dds_comm_shm: mmap failed");
        return nullptr;
    }
    dds_comm *c = new dds_comm;
    c->kind = 1;
    c->rank = rank;
    c->size = size;
    c->hdr = (ShmHeader *)m;
    c->slots = (unsigned char *)m + sizeof(ShmHeader);
    c->map_bytes = bytes;
    c->shm_name = name;
    if (creator) {
        c->hdr->size = (uint32_t)size;
        c->hdr->attached.store(0);
        c->hdr->bar_count.store(0);
        c->hdr->bar_sense.store(0);
        c->hdr->magic.store(kMagic, std::memory_order_release);
    } else {
        unsigned spins = 0;
        while (c->hdr->magic.load(std::memory_order_acquire) != kMagic) {
            backoff(spins++);
            if (now_s() - t0 > limit) {
                munmap(m, bytes);
                delete c;
                dds_internal::fail(DDS_ERR_COMM, "dds_comm_shm: segment never initialised");
                return nullptr;
            }
        }
    }
    if (c->hdr->size != (uint32_t)size || c->hdr->attached.fetch_add(1) >= (uint32_t)size) {
        munmap(m, bytes);
        delete c;
        dds_internal::fail(DDS_ERR_COMM,
                           "dds_comm_shm: stale or mismatched segment for this key (remove /dev/shm/dds_b200_<key> "
                           "or use a unique key per job)");
        return nullptr;
    }
    // everyone has the segment mapped after this barrier; the name can go away so no stale file survives
    if (shm_barrier(c) != DDS_OK) {
        if (creator) shm_unlink(name.c_str()); // a failed rendezvous must not leave a half-used segment behind
        munmap(m, bytes);
        delete c;
        return nullptr;
    }
    if (creator) shm_unlink(name.c_str());
    return c;
}

dds_comm_t *dds_comm_callbacks(int rank, int size, dds_allgather_fn allgather, dds_barrier_fn barrier, void *ctx) {
    if (size < 1 || rank < 0 || rank >= size || !allgather || !barrier) {
        dds_internal::fail(DDS_ERR_ARG, "dds_comm_callbacks: bad arguments");
        return nullptr;
    }
    dds_comm *c = new dds_comm;
    c->kind = 2;
    c->rank = rank;
    c->size = size;
    c->ag = allgather;
    c->bar = barrier;
    c->ctx = ctx;
    return c;
}

int dds_comm_rank(const dds_comm_t *c) { return c ? c->rank : -1; }
int dds_comm_size(const dds_comm_t *c) { return c ? c->size : -1; }

int dds_comm_barrier(dds_comm_t *c) {
    if (!c) return dds_internal::fail(DDS_ERR_ARG, "null communicator");
    switch (c->kind) {
    case 0: return DDS_OK;
    case 1: return shm_barrier(c);
    default:
        if (c->bar(c->ctx) != 0) return dds_internal::fail(DDS_ERR_COMM, "communicator barrier callback failed");
        return DDS_OK;
    }
}

int dds_comm_allgather(dds_comm_t *c, const void *send, void *recv, size_t n) {
    if (!c) return dds_internal::fail(DDS_ERR_ARG, "null communicator");
    if (c->kind == 0) {
        memcpy(recv, send, n);
        return DDS_OK;
    }
    if (c->kind == 2) {
        if (c->ag(c->ctx, send, recv, n) != 0) return dds_internal::fail(DDS_ERR_COMM, "communicator allgather callback failed");
        return DDS_OK;
    }
    // shm: slot-sized pieces
    for (size_t off = 0; off < n || (n == 0 && off == 0); off += kSlotBytes) {
        size_t piece = n - off < kSlotBytes ? n - off : kSlotBytes;
        memcpy(c->slots + (size_t)c->rank * kSlotBytes, (const char *)send + off, piece);
        if (int rc = shm_barrier(c)) return rc;
        for (int r = 0; r < c->size; r++)
            memcpy((char *)recv + (size_t)r * n + off, c->slots + (size_t)r * kSlotBytes, piece);
        if (int rc = shm_barrier(c)) return rc;
        if (n == 0) break;
    }
    return DDS_OK;
}

void dds_comm_free(dds_comm_t *c) {
    if (!c) return;
    if (c->kind == 1 && c->hdr) munmap((void *)c->hdr, c->map_bytes);
    delete c;
}

} // extern "C"
