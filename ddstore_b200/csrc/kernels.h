/* ddstore_b200/csrc/kernels.h -- the thin C-ABI between the host C++ store (store.cpp) and the
 * CUDA side (kernels.cu). Plain pointers, sizes and a cudaStream_t passed as void*. Nothing here
 * is public; the public boundary is include/ddstore_b200.h. */
#ifndef DDSK_KERNELS_H
#define DDSK_KERNELS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDSK_MAX_RANKS 64
#define DDSK_MAX_MULTI 4 /* variables per multi-array launch */

/* Device-visible description of one variable: what the reference keeps in VarInfo
 * (/root/reference/include/ddstore.hpp:10-22) minus the MPI window, plus the peer-mapped shard base
 * of every owner (the "window"). Passed BY VALUE as a kernel parameter (constant bank). */
typedef struct ddsk_var {
    const void *bases[DDSK_MAX_RANKS]; /* shard base of rank r as mapped into THIS process (IPC / peer) */
    int64_t lenlist[DDSK_MAX_RANKS];   /* inclusive cumulative row counts, ddstore.hpp:84-89 */
    int64_t row_bytes;                 /* disp * itemsize, the window's disp_unit, ddstore.hpp:58 */
    int32_t nranks;
    int32_t pad_;
} ddsk_var_t;

/* status word written by the kernels: 0xFFFF... = ok, else (first_bad_request << 8) | code */
#define DDSK_STATUS_OK 0xFFFFFFFFFFFFFFFFull
#define DDSK_CODE_START 2
#define DDSK_CODE_COUNT 3
#define DDSK_CODE_CAPACITY 12
#define DDSK_CODE_WATCHDOG 14
#define DDSK_CODE_SAMPLE 15

/* scratch a store owns for the batched path (all device memory) */
typedef struct ddsk_scratch {
    unsigned long long *status; /* 1 word */
    unsigned int *counters;     /* 4 words: [0] segment ticket, [1] finished warps, [2] plan ticket,
                                   [3] finished plan tiles -- all self-resetting */
    uint64_t *req_src;          /* [cap_req]   planned source address per request (0 = skip) */
    int64_t *req_dst;           /* [cap_req+1] exclusive scan of request bytes */
    int64_t *tile_sums;         /* [cap_req/128 + 2] tile sums (separate plan kernels) / look-back words (fused plan) */
    int64_t cap_req;
    unsigned long long *host_mirror; /* device alias of 2 pinned host words: status, packed total (written by the
                                        last warp of every gather launch); NULL = not used */
    unsigned int epoch;         /* host-side launch counter tagging the look-back words (22 bits, 0 = never) */
    /* slots used by overlapped variable-count launches: their counters only ever grow; these are the values they
     * will have once every launch queued on the slot so far has retired */
    unsigned int ticket_base, tiles_base, finish_target;
} ddsk_scratch_t;

/* `flags` of both launchers: bit 0 = reset the status word first, bit 1 = have the kernel's last warp mirror status +
 * total into scr->host_mirror (synchronous calls; costs ~2 us at the kernel's end, so async queues skip it);
 * bit 2 = independent batch (static segment striding instead of the ticket counters; for the variable entry `scr` must
 * then be a scratch slot of its own, used with monotonic counters and planned in-kernel), bit 4 = its predecessor in
 * the queue was one too (skip griddepcontrol.wait: the two overlap).
 * Fixed-count batch: every request fetches `count` rows; offsets are i*count*row_bytes.
 * One launch: validate + owner lookup + gather + pack. */
int ddsk_gather_fixed(const ddsk_var_t *var, const int64_t *starts_dev, int64_t count, int64_t nreq, void *dst_dev,
                      int64_t dst_capacity, int64_t *offsets_dev_or_null, const ddsk_scratch_t *scr, int flags,
                      void *stream);

/* Where the (start row, row count) of request i comes from (all device pointers): explicit arrays, or -- when
 * sample_ids is set -- the per-sample table of the variable: start = table_start[sample_ids[i]], etc. */
typedef struct ddsk_index {
    const int64_t *starts, *counts;
    const int64_t *sample_ids;
    const int64_t *table_start, *table_count;
    int64_t nsamples;
} ddsk_index_t;

/* Variable-count batch: plan (lookup + validate + exclusive scan) then gather + pack. */
int ddsk_gather_var(const ddsk_var_t *var, const ddsk_index_t *index, int64_t nreq, void *dst_dev,
                    int64_t dst_capacity, int64_t *offsets_dev_or_null, ddsk_scratch_t *scr, int flags,
                    void *stream);

/* Multi-array batch: the rows of the SAME nreq sample ids in nvars (<= DDSK_MAX_MULTI) variables, one launch. vars_dev =
 * device array of the variables' windows; table_*[v] = sample index of variable v; dst[v]/cap[v]/offsets[v] per variable
 * (offsets[v] nullable, nreq+1 entries). On return of the stream, totals are req_dst-derived (see store.cpp). */
typedef struct ddsk_multi {
    int nvars;
    const ddsk_var_t *vars_dev;
    const int64_t *table_start[DDSK_MAX_MULTI], *table_count[DDSK_MAX_MULTI];
    int64_t nsamples[DDSK_MAX_MULTI];
    void *dst[DDSK_MAX_MULTI];
    int64_t cap[DDSK_MAX_MULTI];
    int64_t *offsets[DDSK_MAX_MULTI];
} ddsk_multi_t;
int ddsk_gather_multi(const ddsk_multi_t *m, const int64_t *sample_ids_dev, int64_t nreq, ddsk_scratch_t *scr, int flags,
                      void *stream);

/* Synthetic payload: element (global_row g, col c) = low itemsize bytes of
 * splitmix64(seed ^ (g*disp + c)). Bench / test helper, fills a local shard in place. */
int ddsk_synth_fill(void *base_dev, int64_t first_global_row, int64_t nrows, int64_t disp, int itemsize, uint64_t seed,
                    void *stream);

/* launch geometry actually used (for bench reporting / DESIGN.md) */
void ddsk_gather_geometry(int *ctas, int *warps_per_cta, int *stages, int *chunk_bytes, int *smem_bytes);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long ddsk_launch_count(void);

const char *ddsk_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif
