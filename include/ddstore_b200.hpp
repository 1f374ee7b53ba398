// include/ddstore_b200.hpp -- `class DDStore` with the reference's public shape
// (/root/reference/include/ddstore.hpp:26-258), implemented as a thin header-only wrapper over the C-ABI in
// ddstore_b200.h. A C++ caller of the reference switches by including this header, linking
// libddstore_b200.so, and passing a dds_comm_t* where it used to pass an MPI_Comm.
//
// Same member names, argument order and meaning; the same exception types and texts:
//   std::invalid_argument("Invalid data type" | "Invalid start on target" | "Invalid count on target" |
//                         "Invalid disp")                      ddstore.hpp:82,153,190,203,211,214
//   std::logic_error("Fence already activated" | "Fence is not activated")   ddstore.cxx:58,72
// Everything the reference leaves undefined throws with the C-ABI's message instead: std::out_of_range for an
// unknown variable (the reference default-inserts one, UB), std::runtime_error otherwise (no device, CUDA, comm).
#ifndef DDSTORE_B200_HPP
#define DDSTORE_B200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "ddstore_b200.h"

struct VarInfo { // ddstore.hpp:10-22 (window/base/fabric_state replaced by what exists here)
    std::string name;
    int itemsize;
    int disp;
    std::vector<long> lenlist;
    bool active;
    bool fence_active;
    void *base; // device pointer of the local shard
};
typedef struct VarInfo VarInfo_t;

inline int sortedsearch(std::vector<long> &vec, long num) { // src/ddstore.cxx:5-17
    std::vector<int64_t> v(vec.begin(), vec.end());
    return dds_sortedsearch(v.data(), (int)v.size(), (int64_t)num);
}

class DDStore {
  public:
    // DDStore() -- MPI_COMM_SELF, ddstore.cxx:19-24
    DDStore() : own_comm_(dds_comm_self()), comm_(own_comm_), store_(nullptr) { open(0, -1); }
    // DDStore(MPI_Comm comm), ddstore.cxx:26-31
    explicit DDStore(dds_comm_t *comm, int device = -1) : own_comm_(nullptr), comm_(comm), store_(nullptr) {
        open(0, device);
    }
    // DDStore(int method, MPI_Comm comm), ddstore.cxx:33-39
    DDStore(int method, dds_comm_t *comm, int device = -1) : own_comm_(nullptr), comm_(comm), store_(nullptr) {
        open(method, device);
    }
    ~DDStore() { // ddstore.cxx:41-44 (local teardown; the collective one is free())
        if (store_) dds_destroy(store_);
        if (own_comm_) dds_comm_free(own_comm_);
    }
    DDStore(const DDStore &) = delete;
    DDStore &operator=(const DDStore &) = delete;

    void query(std::string name, VarInfo_t &varinfo) { // ddstore.cxx:46-49
        dds_varinfo_t vi;
        check(dds_query(store_, name.c_str(), &vi));
        varinfo.name = name;
        varinfo.itemsize = vi.itemsize;
        varinfo.disp = vi.disp;
        varinfo.lenlist.assign(vi.lenlist, vi.lenlist + vi.nranks);
        varinfo.active = true;
        varinfo.fence_active = vi.fence_active != 0;
        varinfo.base = vi.local_base;
    }
    void epoch_begin() { check(dds_epoch_begin(store_)); } // ddstore.cxx:51-63
    void epoch_end() { check(dds_epoch_end(store_)); }     // ddstore.cxx:65-77
    void free() { check(dds_free(store_)); }               // ddstore.cxx:79-96

    template <typename T>
    void add(std::string name, T *buffer, long nrows, int disp) { // ddstore.hpp:39-108
        check(dds_add(store_, name.c_str(), buffer, nrows, disp, (int)sizeof(T), 0));
    }
    void init(std::string name, long nrows, int disp, int itemsize) { // ddstore.hpp:110-179
        check(dds_init(store_, name.c_str(), nrows, disp, itemsize));
    }
    template <typename T>
    void update(std::string name, T *buffer, long nrows, long offset = 0) { // ddstore.hpp:181-195
        check(dds_update(store_, name.c_str(), buffer, nrows, offset, (int)sizeof(T), 0));
    }
    template <typename T>
    void get(std::string name, long start, long count, T *buffer) { // ddstore.hpp:197-248
        check(dds_get(store_, name.c_str(), start, count, (int)sizeof(T), buffer, 0));
    }

    // ---- beyond the reference: the batched get() (one kernel launch for the whole batch) -----------------
    // Packs request i = (starts[i], counts[i]) at byte offset sum_{j<i} counts[j]*disp*sizeof(T) of dst.
    // counts == nullptr: every request fetches fixed_count rows. Returns the packed bytes.
    template <typename T>
    long get_batch(std::string name, const long *starts, const long *counts, long fixed_count, long nreq, T *dst,
                   long dst_capacity_bytes, long *dst_offsets = nullptr, bool on_device = false,
                   void *cuda_stream = nullptr) {
        int64_t total = 0, bad = -1;
        unsigned flags = on_device ? (DDS_IDX_ON_DEVICE | DDS_DST_ON_DEVICE) : 0u;
        check(dds_get_batch(store_, name.c_str(), (const int64_t *)starts, (const int64_t *)counts, fixed_count, nreq,
                            (int)sizeof(T), dst, dst_capacity_bytes, (int64_t *)dst_offsets, flags, cuda_stream, &total,
                            &bad));
        return (long)total;
    }
    // device-pointer variants of add/get for callers that already hold the data in HBM
    template <typename T>
    void add_device(std::string name, const T *dev_buffer, long nrows, int disp) {
        check(dds_add(store_, name.c_str(), dev_buffer, nrows, disp, (int)sizeof(T), 1));
    }
    template <typename T>
    void get_device(std::string name, long start, long count, T *dev_buffer) {
        check(dds_get(store_, name.c_str(), start, count, (int)sizeof(T), dev_buffer, 1));
    }

    int rank() const { return dds_rank(store_); }
    int size() const { return dds_size(store_); }
    dds_store_t *handle() { return store_; }

  private:
    void open(int method, int device) {
        store_ = dds_create(comm_, device, method);
        if (!store_) throw std::runtime_error(dds_last_error());
    }
    static void check(int rc) {
        switch (rc) {
        case DDS_OK: return;
        case DDS_ERR_DTYPE:
        case DDS_ERR_START:
        case DDS_ERR_COUNT:
        case DDS_ERR_DISP: throw std::invalid_argument(dds_last_error());
        case DDS_ERR_FENCE_ACTIVE:
        case DDS_ERR_FENCE_INACTIVE: throw std::logic_error(dds_last_error());
        case DDS_ERR_UNKNOWN_VAR: throw std::out_of_range(dds_last_error());
        default: throw std::runtime_error(dds_last_error());
        }
    }
    dds_comm_t *own_comm_;
    dds_comm_t *comm_;
    dds_store_t *store_;
};

#endif
