// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A flat C entry-point layer over the UNMODIFIED reference class `DDStore`
// (/root/reference/include/ddstore.hpp:26-258, /root/reference/src/ddstore.cxx), compiled
// from where the sources lie (never copied into this repo) against oracle/mpi_shim.
// Built by oracle/Makefile into oracle/_ref/libddstore_ref.so. Used by
//   * tests/            -- to pin oracle/ddstore_oracle.c against the real reference
//   * tests/golden/make_golden.py -- to generate the committed golden vectors
//   * bench.py          -- as cpu_baseline kind "reference" and the --impl reference arm
// One rank == one thread (see mpi_shim/mpi.h). Collective calls (add/init/epoch fences)
// fan out over `size` threads here; get() is non-collective and is called directly.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ddstore.hpp" /* the reference's own header, via -I/root/reference/include */

extern "C" {
/* method-1 (libfabric) entry points are referenced from the header templates but are out of
 * scope; they must never run. */
void init_fabric(struct fabric_state *) { fprintf(stderr, "ref_driver: libfabric path is out of scope\n"); abort(); }
int handshake(struct fabric_state *, MPI_Comm) { abort(); }
int read_from_remote(struct fabric_state *, int, uint64_t) { abort(); }
}

namespace {

enum { DT_INT32 = 0, DT_INT64 = 1, DT_UINT8 = 2, DT_FLOAT32 = 3, DT_FLOAT64 = 4, DT_BOOL = 5 };

struct World {
    int size;
    shim_group *group;
    std::vector<DDStore *> ranks;
    std::set<std::string> names;
    std::string last_error;
};

/* dtype -> the C type src/pyddstore.pyx:69-80 instantiates the templates with */
template <typename F>
void with_type(int dtype, F &&f) {
    switch (dtype) {
    case DT_INT32: f((int *)nullptr); break;
    case DT_INT64: f((long *)nullptr); break;
    case DT_UINT8: f((char *)nullptr); break;
    case DT_FLOAT32: f((float *)nullptr); break;
    case DT_FLOAT64: f((double *)nullptr); break;
    case DT_BOOL: f((char *)nullptr); break;
    default: throw std::invalid_argument("ref_driver: bad dtype code");
    }
}

int fanout(World *w, const std::function<void(int)> &fn) {
    std::vector<std::string> errs((size_t)w->size);
    std::vector<std::thread> th;
    for (int r = 0; r < w->size; r++)
        th.emplace_back([&, r] {
            try {
                fn(r);
            } catch (const std::exception &e) {
                errs[(size_t)r] = e.what();
                if (errs[(size_t)r].empty()) errs[(size_t)r] = "exception";
            }
        });
    for (auto &t : th) t.join();
    int rc = 0;
    for (int r = 0; r < w->size; r++)
        if (!errs[(size_t)r].empty()) {
            if (!rc) w->last_error = errs[(size_t)r];
            rc = 1;
        }
    return rc;
}

} // namespace

extern "C" {

void *ref_world_create(int size) {
    World *w = new World;
    w->size = size;
    w->group = shim_group_create(size);
    for (int r = 0; r < size; r++) w->ranks.push_back(new DDStore(0, shim_group_comm(w->group, r)));
    return w;
}

void ref_world_destroy(void *h) {
    World *w = (World *)h;
    for (auto *d : w->ranks) delete d; /* ~DDStore -> free() -> MPI_Win_free per rank */
    shim_group_destroy(w->group);
    delete w;
}

const char *ref_last_error(void *h) { return ((World *)h)->last_error.c_str(); }

/* int sortedsearch(std::vector<long>&, long)  -- src/ddstore.cxx:5-17, called verbatim */
int ref_sortedsearch(const long *vec, int n, long num) {
    std::vector<long> v(vec, vec + n);
    return sortedsearch(v, num);
}

/* collective DDStore::add<T> (include/ddstore.hpp:39-108) on every rank; bufs[r] has nrows[r] x disp[r] */
int ref_add(void *h, const char *name, int dtype, const void *const *bufs, const long *nrows, const int *disp) {
    World *w = (World *)h;
    std::string nm(name);
    int rc = fanout(w, [&](int r) {
        with_type(dtype, [&](auto *tp) {
            typedef typename std::remove_pointer<decltype(tp)>::type T;
            w->ranks[(size_t)r]->add<T>(nm, (T *)bufs[r], nrows[r], disp[r]);
        });
    });
    w->names.insert(nm);
    return rc;
}

/* collective DDStore::init (include/ddstore.hpp:110-179) */
int ref_init(void *h, const char *name, const long *nrows, const int *disp, int itemsize) {
    World *w = (World *)h;
    std::string nm(name);
    int rc = fanout(w, [&](int r) { w->ranks[(size_t)r]->init(nm, nrows[r], disp[r], itemsize); });
    w->names.insert(nm);
    return rc;
}

/* local DDStore::update<T> (include/ddstore.hpp:181-195) */
int ref_update(void *h, int rank, const char *name, int dtype, const void *buf, long nrows, long offset) {
    World *w = (World *)h;
    std::string nm(name);
    if (!w->names.count(nm)) { w->last_error = "ref_driver: unknown variable (UB in the reference)"; return 2; }
    try {
        with_type(dtype, [&](auto *tp) {
            typedef typename std::remove_pointer<decltype(tp)>::type T;
            w->ranks[(size_t)rank]->update<T>(nm, (T *)buf, nrows, offset);
        });
    } catch (const std::exception &e) { w->last_error = e.what(); return 1; }
    return 0;
}

/* DDStore::get<T> (include/ddstore.hpp:197-248), one request, as rank `rank` */
int ref_get(void *h, int rank, const char *name, int dtype, long start, long count, void *buf) {
    World *w = (World *)h;
    std::string nm(name);
    if (!w->names.count(nm)) { w->last_error = "ref_driver: unknown variable (UB in the reference)"; return 2; }
    try {
        with_type(dtype, [&](auto *tp) {
            typedef typename std::remove_pointer<decltype(tp)>::type T;
            w->ranks[(size_t)rank]->get<T>(nm, start, count, (T *)buf);
        });
    } catch (const std::exception &e) { w->last_error = e.what(); return 1; }
    return 0;
}

/* The loader pattern (examples/vae/distdataset.py:79-89): n blocking get() calls in a row, each
 * into the next free bytes of `out` (row_bytes = disp*itemsize of the variable). Returns elapsed
 * nanoseconds, or -1 on the first failing request (its index in *bad). */
long long ref_get_loop(void *h, int rank, const char *name, int dtype, const long *starts, const long *counts,
                       long n, long row_bytes, char *out, long *bad) {
    World *w = (World *)h;
    std::string nm(name);
    if (!w->names.count(nm)) { w->last_error = "ref_driver: unknown variable"; return -2; }
    DDStore *d = w->ranks[(size_t)rank];
    auto t0 = std::chrono::steady_clock::now();
    long i = 0;
    try {
        with_type(dtype, [&](auto *tp) {
            typedef typename std::remove_pointer<decltype(tp)>::type T;
            char *p = out;
            for (i = 0; i < n; i++) {
                d->get<T>(nm, starts[i], counts[i], (T *)p);
                p += counts[i] * row_bytes;
            }
        });
    } catch (const std::exception &e) {
        w->last_error = e.what();
        if (bad) *bad = i;
        return -1;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
}

/* All ranks run ref_get_loop concurrently (one thread per rank == "all the host threads it can
 * use"); starts/counts/outs are per-rank arrays. Returns the slowest rank's nanoseconds. */
long long ref_get_loop_all(void *h, const char *name, int dtype, const long *const *starts,
                           const long *const *counts, long n, long row_bytes, char *const *outs) {
    World *w = (World *)h;
    std::vector<long long> ns((size_t)w->size, 0);
    std::atomic<int> ready(0);
    std::vector<std::thread> th;
    for (int r = 0; r < w->size; r++)
        th.emplace_back([&, r] {
            ready.fetch_add(1);
            while (ready.load() < w->size) { /* start line */ }
            ns[(size_t)r] = ref_get_loop(h, r, name, dtype, starts[r], counts[r], n, row_bytes, outs[r], nullptr);
        });
    for (auto &t : th) t.join();
    long long worst = 0;
    for (auto v : ns) {
        if (v < 0) return v;
        if (v > worst) worst = v;
    }
    return worst;
}

/* collective epoch_begin / epoch_end (src/ddstore.cxx:51-77) */
int ref_epoch_begin(void *h) {
    World *w = (World *)h;
    return fanout(w, [&](int r) { w->ranks[(size_t)r]->epoch_begin(); });
}
int ref_epoch_end(void *h) {
    World *w = (World *)h;
    return fanout(w, [&](int r) { w->ranks[(size_t)r]->epoch_end(); });
}

/* DDStore::query (src/ddstore.cxx:46-49): copies out itemsize, disp and lenlist[size] */
int ref_query(void *h, int rank, const char *name, int *itemsize, int *disp, long *lenlist) {
    World *w = (World *)h;
    std::string nm(name);
    if (!w->names.count(nm)) { w->last_error = "ref_driver: unknown variable"; return 2; }
    VarInfo_t vi;
    w->ranks[(size_t)rank]->query(nm, vi);
    *itemsize = vi.itemsize;
    *disp = vi.disp;
    for (size_t i = 0; i < vi.lenlist.size(); i++) lenlist[i] = vi.lenlist[i];
    return 0;
}

} // extern "C"
