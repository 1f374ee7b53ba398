// ddstore_b200/csrc/kernels.cu -- the get() hot path as hand-written sm_100a CUDA.
//
// What the reference does per sample (include/ddstore.hpp:197-238 + src/ddstore.cxx:5-17):
//   owner = sortedsearch(lenlist, start); offset = lenlist[owner-1] (or 0); two range checks;
//   MPI_Get of count*disp*itemsize bytes from the owner's window at row (start-offset).
// What this file does per BATCH, in one persistent kernel (dds_gather_kernel):
//   the same lookup + checks for every request, then a gather of all payloads from the owners'
//   HBM shards (local, or peer-mapped over NVLink/NVSwitch -- CUDA VMM blocks shared by file
//   descriptor, see vmm.cpp) packed back to back into one contiguous device buffer.
//
// Kernel design (bandwidth-bound byte mover, no tensor cores):
//   * The packed destination byte range [0, T) is cut into segments that warps claim dynamically
//     (one atomic per segment; statically strided for DDS_OVERLAP launches), so load balance is by
//     BYTES, not by request count (lengths differ 100x in the variable-length configs) and not by
//     owner (remote rows are slower than local).
//   * Every warp is an autonomous pipeline with a private ring of S shared-memory stages. A stage
//     carries a GROUP of up to 32 pieces, one per lane (consecutive small requests, or one <= CH-byte
//     piece of a large one); every lane issues the 1-D TMA bulk load of its own piece
//     (cp.async.bulk global->shared, mbarrier complete_tx), S-1 stages ahead, on the
//     16-byte-aligned superset of the source range, so arbitrary element alignment (4-byte
//     floats, single bytes) is legal for TMA.
//   * Drain, per piece:
//       - staged bytes, destination, size all 16-byte aligned -> the piece's own lane issues one
//         TMA bulk store shared->global, all lanes at once;
//       - same 16-byte phase, ragged ends -> lane 0 bulk-stores the body, byte stores for head/tail;
//       - different phase -> all lanes read two aligned 16-byte vectors from shared memory,
//         funnel-shift (or word-select) them into place and issue aligned 128-bit stores.
//   * Request offsets in the packed buffer are an exclusive prefix sum of request sizes: arithmetic
//     in the fixed-count entry; for variable counts a warp-shuffle scan with decoupled look-back
//     INSIDE the gather launch (<= 8192 requests) or in two small plan kernels before it. The
//     (start, count) of a request may come from a device-resident per-sample index (sample ids in).
//   * Launches carry the programmatic-dependent-launch attribute; independent batches
//     (DDS_OVERLAP) skip the grid wait and overlap head-to-tail.
//
// Nothing here calls a library kernel; everything is launched from the ddsk_* functions at the end.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "kernels.h"

namespace {

thread_local char g_cuda_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            snprintf(g_cuda_err, sizeof(g_cuda_err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                     __FILE__, __LINE__);                                                                \
            return (int)e__ ? (int)e__ : -1;                                                             \
        }                                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------------------
// PTX helpers (sm_100a): mbarrier, 1-D TMA bulk copies, shared-memory vector access
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void stg128(void *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// owner lookup + range checks, exactly the reference's arithmetic
// ------------------------------------------------------------------------------------------------
// src/ddstore.cxx:5-17: first i>=1 with vec[i-1] <= num < vec[i]; else 0 (also when out of range).
__device__ __forceinline__ int dev_sortedsearch(const ddsk_var_t &v, int64_t num) {
    int rtn = 0;
    for (int i = 1; i < v.nranks; i++) {
        if (v.lenlist[i - 1] <= num && num < v.lenlist[i]) {
            rtn = i;
            break;
        }
    }
    return rtn;
}

// include/ddstore.hpp:205-214. Returns 0 or DDSK_CODE_*; *src = address of the first payload byte.
__device__ __forceinline__ int dev_locate(const ddsk_var_t &v, int64_t start, int64_t count, uint64_t *src) {
    int t = dev_sortedsearch(v, start);
    int64_t off = t > 0 ? v.lenlist[t - 1] : 0;
    if (start < off) return DDSK_CODE_START;
    if (count < 0 || start + count > v.lenlist[t]) return DDSK_CODE_COUNT; /* count<0 is UB in the reference */
    *src = (uint64_t)v.bases[t] + (uint64_t)(start - off) * (uint64_t)v.row_bytes; /* ddstore.hpp:229-236 */
    return 0;
}

__device__ __forceinline__ void report(unsigned long long *status, int64_t req, int code) {
    atomicMin(status, ((unsigned long long)req << 8) | (unsigned long long)code);
}

// ------------------------------------------------------------------------------------------------
// gather kernel
// ------------------------------------------------------------------------------------------------
// where request i's (start row, row count) comes from: explicit arrays, or a per-sample table indexed by ids[i]
struct PlanSrc {
    const int64_t *starts, *counts;       // explicit (ids == nullptr)
    const int64_t *ids;                   // sample ids (device-resident sample index)
    const int64_t *tab_start, *tab_count; // [nsamples] row_start / row_count of every sample of this variable
    int64_t nsamples;
    // multi-array batches (config 4: node_feat + edge_index of the same samples in ONE launch): request i belongs to
    // variable i / per_var and to sample ids[i % per_var]; every variable has its own window and sample index
    int nvars;                                   // 0/1: single variable
    int64_t per_var;                             // requests per variable (= number of sample ids)
    const ddsk_var_t *mvars;                     // [nvars] windows, device memory
    const int64_t *mtab_start[DDSK_MAX_MULTI], *mtab_count[DDSK_MAX_MULTI];
    int64_t mnsamples[DDSK_MAX_MULTI];
};

// Lookup + checks of K requests per thread -> (source address or 0, byte size). Written as three unrolled passes
// (ids, then table rows, then arithmetic) so that the K independent -- and for the sample index, dependent
// two-level -- global loads of a thread are all in flight together instead of one DRAM latency after another.
template <int K>
__device__ __forceinline__ void plan_many(const ddsk_var_t &var, const PlanSrc &p, const int64_t (&idx)[K], int64_t nreq,
                                          unsigned long long *status, uint64_t (&src)[K], int64_t (&nbytes)[K]) {
    int64_t start[K], count[K];
    bool live[K], badid[K];
    const ddsk_var_t *vp[K];
    if (p.nvars > 1) {
        int64_t id[K];
        int v[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            live[k] = idx[k] < nreq;
            v[k] = live[k] ? (int)(idx[k] / p.per_var) : 0;
            id[k] = live[k] ? p.ids[idx[k] - (int64_t)v[k] * p.per_var] : 0;
            vp[k] = &p.mvars[v[k]];
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            badid[k] = live[k] && (id[k] < 0 || id[k] >= p.mnsamples[v[k]]);
            const bool ok = live[k] && !badid[k];
            start[k] = ok ? p.mtab_start[v[k]][id[k]] : 0;
            count[k] = ok ? p.mtab_count[v[k]][id[k]] : 0;
        }
    } else if (p.ids) {
        int64_t id[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            live[k] = idx[k] < nreq;
            id[k] = live[k] ? p.ids[idx[k]] : 0;
            vp[k] = &var;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            badid[k] = live[k] && (id[k] < 0 || id[k] >= p.nsamples);
            const bool ok = live[k] && !badid[k];
            start[k] = ok ? p.tab_start[id[k]] : 0;
            count[k] = ok ? p.tab_count[id[k]] : 0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; k++) {
            live[k] = idx[k] < nreq;
            badid[k] = false;
            start[k] = live[k] ? p.starts[idx[k]] : 0;
            count[k] = live[k] ? p.counts[idx[k]] : 0;
            vp[k] = &var;
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        src[k] = 0;
        nbytes[k] = 0;
        if (!live[k]) continue;
        if (badid[k]) {
            report(status, idx[k], DDSK_CODE_SAMPLE);
            continue;
        }
        uint64_t s = 0;
        const int code = dev_locate(*vp[k], start[k], count[k], &s);
        if (code) {
            report(status, idx[k], code);
            continue;
        }
        src[k] = s;
        nbytes[k] = count[k] * vp[k]->row_bytes;
    }
}

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

struct GatherArgs {
    ddsk_var_t var;
    const int64_t *starts;   // FIXED: start row per request
    int64_t count;           // FIXED: rows per request
    uint64_t *req_src; // VAR: planned source address (0 = skip)
    int64_t *req_dst;  // VAR: [nreq+1] exclusive scan; req_dst[nreq] = total bytes
    PlanSrc plan;      // VAR with fused_plan: where (start, count) of request i comes from
    unsigned long long *tile_state; // VAR with fused_plan: one look-back word per 128-request tile
    unsigned int epoch;             // tags tile_state words of THIS launch (no per-launch memset)
    int fused_plan;
    int64_t nreq;
    char *dst;
    int64_t dst_cap;
    int64_t *offsets_out; // FIXED: optional [nreq+1]
    unsigned long long *status;
    unsigned int *counters;
    // multi-array batches (plan.nvars > 1): the packed result of variable v goes to mdst[v]
    char *mdst[DDSK_MAX_MULTI];
    int64_t mcap[DDSK_MAX_MULTI];
    int64_t *moffsets[DDSK_MAX_MULTI]; // optional per-variable [per_var + 1] byte offsets
    int overlap;   // declared independent of its neighbours in the queue: segments are strided statically instead of
                   // ticketed; a variable-count launch then works in its own scratch slot with monotonic counters
    int skip_wait; // ... and the launch before it was one too: do not wait for it to finish
    unsigned int ticket_base, tiles_base, finish_target; // VAR + overlap: values of the slot's monotonic counters
                                                         // [2] (plan tickets), [3] (tiles done), [1] (warps finished)
                                                         // after every earlier user of the slot
    unsigned long long *host_mirror; // zero-copy pinned host words: [0] status, [1] packed total (written at kernel end)
};

// One pipeline stage carries a GROUP of up to 32 pieces (one per lane): consecutive requests of the walk, or one
// <= CH-byte piece of a large request. Small requests therefore still put ~CH bytes in flight per stage.
struct Piece {
    uint64_t src;  // first payload byte (0: nothing to copy)
    int64_t dpos;  // byte position in the packed buffer
    uint32_t n;    // payload bytes (0: lane idle)
    uint32_t off;  // byte offset of this piece's aligned superset inside the stage
};

template <bool FIXED, int CH>
struct ChunkWalker {
    // warp-uniform state
    int64_t seg_pos = 0, seg_end = 0, T = 0, seg_bytes = 0, nseg = 0, nb = 0;
    int64_t gwarp = 0, nwarps = 1; // this warp's global index / warps in the grid (first segment = gwarp)
    bool first_claim = true, static_claims = false;
    int64_t cur_seg = 0;
    int64_t r = 0, win_base = -64;
    // per-lane window of 32 request descriptors
    uint64_t w_src = 0;
    int64_t w_dst = 0, w_n = 0;

    __device__ __forceinline__ void load_window(const GatherArgs &a, int lane) {
        win_base = r;
        int64_t idx = r + lane;
        w_src = 0;
        w_dst = 0;
        w_n = 0;
        if (idx < a.nreq) {
            if (FIXED) {
                uint64_t s = 0;
                int code = dev_locate(a.var, a.starts[idx], a.count, &s);
                if (code) report(a.status, idx, code); // the reference's two checks; every request with bytes to
                                                       // fetch passes through some warp's window at least once
                w_src = code ? 0 : s; // invalid request: keep its slot in the packed layout, copy nothing
                w_dst = idx * nb;
                w_n = nb;
            } else {
                w_src = a.req_src[idx];
                w_dst = a.req_dst[idx];
                w_n = a.req_dst[idx + 1] - w_dst;
            }
        }
    }

    // largest r in [0, nreq) with req_dst[r] <= pos, 32-ary search across the lanes
    __device__ __forceinline__ int64_t locate_var(const GatherArgs &a, int64_t pos, int lane) {
        int64_t lo = 0, hi = a.nreq;
        while (hi - lo > 1) {
            int64_t step = (hi - lo + 31) / 32;
            int64_t idx = lo + (int64_t)(lane + 1) * step;
            bool le = (idx < hi) && (a.req_dst[idx] <= pos);
            int k = __popc(__ballot_sync(0xffffffffu, le));
            lo = lo + (int64_t)k * step;
            hi = min(hi, lo + step);
        }
        return lo;
    }

    // Next group of the walk. Returns the bytes to expect in the stage (0: no more work); `pc` is this lane's piece.
    template <int STAGE>
    __device__ __forceinline__ uint32_t next_group(const GatherArgs &a, int lane, Piece &pc) {
        while (true) {
            if (seg_pos >= seg_end) {
                // the first segment of warp g is segment g (no ticket: spares ~1800 same-address atomics at the
                // start of every launch); later ones come from the ticket counter, offset by the warp count
                int64_t seg;
                if (first_claim) {
                    first_claim = false;
                    seg = gwarp;
                } else if (static_claims) {
                    seg = cur_seg + nwarps; // overlapped launches share no mutable state: plain striding
                } else {
                    unsigned int t = 0;
                    if (lane == 0) t = atomicAdd(&a.counters[0], 1u);
                    seg = nwarps + (int64_t)__shfl_sync(0xffffffffu, t, 0);
                }
                cur_seg = seg;
                if (seg >= nseg) return 0;
                seg_pos = seg * seg_bytes;
                seg_end = min(T, seg_pos + seg_bytes);
                r = FIXED ? seg_pos / nb : locate_var(a, seg_pos, lane);
            }
            if (r >= a.nreq) { // defensive: cannot happen while seg_pos < T
                seg_pos = seg_end;
                continue;
            }
            if (r < win_base || r >= win_base + 32) load_window(a, lane);
            {   // fast path: the current request alone (nearly) fills a stage, or is cut by CH / the segment end
                const int wl = (int)(r - win_base);
                const uint64_t s0 = __shfl_sync(0xffffffffu, w_src, wl);
                const int64_t d0 = __shfl_sync(0xffffffffu, w_dst, wl);
                const int64_t e0 = d0 + __shfl_sync(0xffffffffu, w_n, wl);
                const int64_t p0 = max(d0, seg_pos);
                const int64_t len = min(min(e0, seg_end) - p0, (int64_t)CH);
                if (len >= CH / 2 || p0 + len < e0) {
                    seg_pos = p0 + len;
                    if (seg_pos >= e0) r++;
                    if (s0 == 0) continue; // rejected request (FIXED): its slot stays untouched
                    const uint64_t src = s0 + (uint64_t)(p0 - d0);
                    pc.src = lane == 0 ? src : 0;
                    pc.dpos = p0;
                    pc.n = lane == 0 ? (uint32_t)len : 0u;
                    pc.off = 0;
                    return ((uint32_t)(src & 15u) + (uint32_t)len + 15u) & ~15u;
                }
            }
            // group path: lane j looks at request r + j (as long as the 32-entry window covers it)
            const int srcl = (int)(r - win_base) + lane;
            const uint64_t q_src = __shfl_sync(0xffffffffu, w_src, srcl & 31);
            const int64_t q_d0 = __shfl_sync(0xffffffffu, w_dst, srcl & 31);
            const int64_t q_n = __shfl_sync(0xffffffffu, w_n, srcl & 31);
            const int64_t q_end = q_d0 + q_n;
            const bool valid = srcl < 32 && r + lane < a.nreq && q_d0 < seg_end;
            const int64_t p0 = max(q_d0, seg_pos);
            int64_t len = min(q_end, seg_end) - p0;
            len = max(len, (int64_t)0);
            len = min(len, (int64_t)CH);
            const bool complete = p0 + len >= q_end; // this piece finishes its request
            const bool copy = valid && len > 0 && q_src != 0;
            const uint64_t src = q_src + (uint64_t)(p0 - q_d0);
            const uint32_t sz = copy ? (((uint32_t)(src & 15u) + (uint32_t)len + 15u) & ~15u) : 0u;
            uint32_t end = sz; // inclusive scan of the padded sizes -> stage offsets
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, end, d);
                if (lane >= d) end += o;
            }
            const unsigned ok = __ballot_sync(0xffffffffu, valid && end <= (uint32_t)STAGE);
            const unsigned part = __ballot_sync(0xffffffffu, valid && !complete);
            int m = ok == 0xffffffffu ? 32 : __ffs(~ok) - 1;   // leading lanes that are valid and fit
            if (part) m = min(m, __ffs(part));                  // ... up to and including the first partial piece
            // lane 0 is always valid and fits (STAGE >= CH + 30), so m >= 1
            const bool active = lane < m;
            const int64_t new_pos = __shfl_sync(0xffffffffu, p0 + len, m - 1);
            const uint32_t total = __shfl_sync(0xffffffffu, end, m - 1);
            r += __popc(__ballot_sync(0xffffffffu, active && complete));
            seg_pos = max(seg_pos, new_pos);
            if (total == 0) continue; // only empty / rejected requests in this run
            pc.src = (active && copy) ? src : 0;
            pc.dpos = p0;
            pc.n = (active && copy) ? (uint32_t)len : 0u;
            pc.off = end - sz;
            return total;
        }
    }
};

// Re-phase loop: output vector j = staged bytes [q16 + 16j + 4*WS + bs, +16). Specialised on the word shift WS
// (and on whether a sub-word byte shift is needed at all) so the loop body is branch-free: two aligned 128-bit
// shared loads, at most four funnel shifts, one aligned 128-bit global store.
template <int WS, bool BYTES>
__device__ __forceinline__ void rephase_loop(uint32_t sbase, char *dv, uint32_t nv, uint32_t bs8, int lane) {
#pragma unroll 4
    for (uint32_t j = (uint32_t)lane; j < nv; j += 32) {
        const uint4 lo = lds128(sbase + (j << 4));
        const uint4 hi = lds128(sbase + (j << 4) + 16);
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint4 out;
        if (BYTES) {
            out.x = __funnelshift_r(w[WS + 0], w[WS + 1], bs8);
            out.y = __funnelshift_r(w[WS + 1], w[WS + 2], bs8);
            out.z = __funnelshift_r(w[WS + 2], w[WS + 3], bs8);
            out.w = __funnelshift_r(w[WS + 3], w[WS + 4], bs8);
        } else { // 4-byte-aligned shift (float32 / int32 / int64 rows): pure word selection
            out.x = w[WS + 0];
            out.y = w[WS + 1];
            out.z = w[WS + 2];
            out.w = w[WS + 3];
        }
        stg128(dv + ((size_t)j << 4), out);
    }
}

// Drain one staged piece: payload byte k lives at shared address sb + a + k and goes to d[k].
template <int CH>
__device__ __forceinline__ void drain_chunk(uint32_t sb, uint32_t a, char *d, uint32_t n, int lane) {
    uint32_t head = (16u - (uint32_t)((uint64_t)d & 15u)) & 15u;
    if (head > n) head = n;
    uint32_t nv = (n - head) >> 4;
    uint32_t tail = n - head - (nv << 4);
    uint32_t s = a + head; // shared offset of the first body byte, 0..30
    uint32_t sh = s & 15u;
    if (nv) {
        if (sh == 0) {
            // source and destination share the 16-byte phase: one bulk store moves the whole body
            if (lane == 0) {
                fence_proxy_async();
                tma_store_1d(d + head, sb + s, nv << 4);
            }
        } else {
            const uint32_t sbase = sb + (s & ~15u);
            const uint32_t bs8 = (sh & 3u) * 8u;
            char *dv = d + head;
            switch ((sh >> 2) * 2u + (bs8 ? 1u : 0u)) { // warp-uniform
            case 0: rephase_loop<0, false>(sbase, dv, nv, bs8, lane); break; // unreachable (sh == 0), kept for the table
            case 1: rephase_loop<0, true>(sbase, dv, nv, bs8, lane); break;
            case 2: rephase_loop<1, false>(sbase, dv, nv, bs8, lane); break;
            case 3: rephase_loop<1, true>(sbase, dv, nv, bs8, lane); break;
            case 4: rephase_loop<2, false>(sbase, dv, nv, bs8, lane); break;
            case 5: rephase_loop<2, true>(sbase, dv, nv, bs8, lane); break;
            case 6: rephase_loop<3, false>(sbase, dv, nv, bs8, lane); break;
            default: rephase_loop<3, true>(sbase, dv, nv, bs8, lane); break;
            }
        }
    }
    if ((uint32_t)lane < head) d[lane] = (char)lds8(sb + a + (uint32_t)lane);
    if ((uint32_t)lane < tail) {
        uint32_t k = head + (nv << 4) + (uint32_t)lane;
        d[k] = (char)lds8(sb + a + k);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused plan (variable counts): the lookup + checks + exclusive scan of request sizes run INSIDE the gather
// launch. Warps take 128-request tiles by ticket; a tile's offset comes from a decoupled look-back over the
// tiles before it (each lane polls one predecessor). Tiles are ticketed in running order, so a tile only ever
// waits for tiles held by warps that are already running: no co-residency assumption, no deadlock.
// ------------------------------------------------------------------------------------------------
constexpr int TILE_ITEMS = 4;
constexpr int TILE_REQ = 32 * TILE_ITEMS;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// tile_state word: [63:42] epoch (22 bits) | [41:40] flag (1 = aggregate, 2 = inclusive prefix) | [39:0] bytes
__device__ __forceinline__ unsigned long long tile_pack(unsigned int epoch, unsigned int flag, int64_t v) {
    return ((unsigned long long)(epoch & 0x3FFFFFu) << 42) | ((unsigned long long)flag << 40) |
           ((unsigned long long)v & 0xFFFFFFFFFFull);
}

__device__ __forceinline__ void plan_in_kernel(const GatherArgs &a, int lane) {
    const int64_t ntiles = (a.nreq + TILE_REQ - 1) / TILE_REQ;
    while (true) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(&a.counters[2], 1u) - a.ticket_base; // base = 0 for self-resetting counters
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((int64_t)tile >= ntiles) break; // (every warp makes exactly one failing claim: the host counts on that)
        // lane owns TILE_ITEMS consecutive requests
        int64_t idx[TILE_ITEMS], nb[TILE_ITEMS];
        uint64_t sv[TILE_ITEMS];
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) idx[k] = (int64_t)tile * TILE_REQ + lane * TILE_ITEMS + k;
        plan_many<TILE_ITEMS>(a.var, a.plan, idx, a.nreq, a.status, sv, nb);
        int64_t mine = 0;
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) mine += nb[k];
        const int64_t incl = warp_incl_scan(mine, lane);
        const int64_t agg = __shfl_sync(0xffffffffu, incl, 31);
        // publish the aggregate, then look back
        if (lane == 0) st_release_u64(&a.tile_state[tile], tile_pack(a.epoch, tile == 0 ? 2u : 1u, agg));
        int64_t excl = 0;
        if (tile > 0) {
            int64_t base = (int64_t)tile - 1;
            while (true) {
                const int64_t t = base - lane; // lane 0 polls the nearest predecessor
                unsigned long long wv = tile_pack(a.epoch, 2u, 0);
                if (t >= 0) {
                    const uint64_t t0 = globaltimer_ns();
                    do {
                        wv = ld_acquire_u64(&a.tile_state[t]);
                        if (globaltimer_ns() - t0 > 4000000000ull) { // never expected; do not hang the box
                            report(a.status, a.nreq, DDSK_CODE_WATCHDOG);
                            __trap();
                        }
                    } while ((unsigned int)(wv >> 42) != (a.epoch & 0x3FFFFFu) || ((wv >> 40) & 3u) == 0);
                }
                const unsigned int flag = (unsigned int)((wv >> 40) & 3u);
                const int64_t val = (int64_t)(wv & 0xFFFFFFFFFFull);
                const unsigned inc = __ballot_sync(0xffffffffu, flag == 2u);
                const int stop = inc ? __ffs(inc) - 1 : 31; // nearest tile that already knows its inclusive prefix
                int64_t c = lane <= stop ? val : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                excl += c;
                if (inc) break;
                base -= 32;
            }
            if (lane == 0) st_release_u64(&a.tile_state[tile], tile_pack(a.epoch, 2u, excl + agg));
        }
        int64_t run = excl + incl - mine;
#pragma unroll
        for (int k = 0; k < TILE_ITEMS; k++) {
            if (idx[k] < a.nreq) {
                a.req_src[idx[k]] = sv[k];
                a.req_dst[idx[k]] = run;
                if (a.offsets_out) a.offsets_out[idx[k]] = run;
                run += nb[k];
            }
        }
        if ((int64_t)tile == ntiles - 1 && lane == 31) {
            a.req_dst[a.nreq] = excl + agg;
            if (a.offsets_out) a.offsets_out[a.nreq] = excl + agg;
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            atomicAdd(&a.counters[3], 1u);
        }
    }
    // every tile has an owner that is running; wait until all of them have written their part of the plan
    if (lane == 0) {
        const uint64_t t0 = globaltimer_ns();
        while ((int64_t)(unsigned int)(ld_acquire_u32(&a.counters[3]) - a.tiles_base) < ntiles) {
            __nanosleep(200);
            if (globaltimer_ns() - t0 > 4000000000ull) {
                report(a.status, a.nreq, DDSK_CODE_WATCHDOG);
                __trap();
            }
        }
    }
    __syncwarp();
}

template <bool FIXED, int NW, int S, int CH>
__global__ void __launch_bounds__(NW * 32, 1) dds_gather_kernel(const __grid_constant__ GatherArgs a) {
    constexpr int STAGE = CH + 32; // room for the aligned superset of a misaligned CH-byte range
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[NW][S];
    __shared__ __align__(16) struct PieceDesc {
        int64_t dpos;
        uint32_t n, pack; // pack = stage offset | (source misalignment << 16)
    } desc[NW][S][32];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * NW + warp;
    const int64_t nwarps = (int64_t)gridDim.x * NW;

    // Programmatic dependent launch: let the NEXT kernel of the stream start launching right away (its CTAs
    // take over each SM as ours retire), and do not touch global memory before the PREVIOUS kernel (which may
    // have produced our indices / plan, and resets the ticket counters) has completed and flushed.
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(smem_u32(&full_bar[warp][s]), 1);
        fence_mbar_init();
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // An overlapped batch shares nothing with the launches before it (own destination, indices already in place,
    // no ticket counters), so its CTAs start moving bytes as soon as an SM frees up: the tail of batch k and the
    // head of batch k+1 overlap, whatever else is running on the GPU.
    // (The first batch of such a run still waits, so a ticketed launch before the run has retired for good before
    // any ticketed launch after the run can start.)
    if (!a.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!FIXED && a.overlap) {
        // the scratch slot of an overlapped variable-count batch is reused every few launches: its previous user must
        // have retired completely (all of ITS CTAs have started long ago, so this cannot deadlock)
        if (lane == 0) {
            const uint64_t t0 = globaltimer_ns();
            while ((int)(ld_acquire_u32(&a.counters[1]) - a.finish_target) < 0) {
                __nanosleep(100);
                if (globaltimer_ns() - t0 > 4000000000ull) {
                    report(a.status, a.nreq, DDSK_CODE_WATCHDOG);
                    __trap();
                }
            }
        }
    }
    __syncwarp();

    if (!FIXED && a.fused_plan) plan_in_kernel(a, lane);

    // ---- total bytes, segment geometry -------------------------------------------------------
    ChunkWalker<FIXED, CH> w;
    w.gwarp = gwarp;
    w.nwarps = nwarps;
    w.static_claims = a.overlap != 0;
    w.nb = FIXED ? a.count * a.var.row_bytes : 0;
    w.T = FIXED ? w.nb * a.nreq : *(volatile const int64_t *)&a.req_dst[a.nreq];
    bool over = w.T > a.dst_cap;
    // multi-array batch: the walk runs over the concatenation of the variables' packed results; vbase[v] is where
    // variable v starts in that virtual space (unused slots are +inf so dst_of() never selects them)
    const bool multi = !FIXED && a.plan.nvars > 1;
    int64_t vbase[DDSK_MAX_MULTI + 1];
#pragma unroll
    for (int v = 0; v <= DDSK_MAX_MULTI; v++) vbase[v] = INT64_MAX;
    if (multi) {
        over = false;
#pragma unroll
        for (int v = 0; v < DDSK_MAX_MULTI; v++)
            if (v < a.plan.nvars) vbase[v] = *(volatile const int64_t *)&a.req_dst[(int64_t)v * a.plan.per_var];
#pragma unroll
        for (int v = 0; v < DDSK_MAX_MULTI; v++)
            if (v < a.plan.nvars) {
                const int64_t endv = v + 1 < a.plan.nvars ? vbase[v + 1] : w.T;
                over |= endv - vbase[v] > a.mcap[v];
            }
    }
    auto dst_of = [&](int64_t dpos) -> char * {
        if (!multi) return a.dst + dpos;
        int64_t b = vbase[0]; // static indices only: the tables stay in registers / the constant bank
        char *d = a.mdst[0];
#pragma unroll
        for (int k = 1; k < DDSK_MAX_MULTI; k++)
            if (dpos >= vbase[k]) {
                b = vbase[k];
                d = a.mdst[k];
            }
        return d + (dpos - b);
    };
    {
        // FIXED: a claim is one atomic + a division, so small segments (8 per warp) keep the tail short.
        // VAR: every claim also costs a 32-ary search over req_dst (2-3 dependent L2 round trips), so segments are
        // at least 4 chunks (measured: config 3 at B=4096 47.6 -> 41.8 us).
        int64_t target = w.T / (nwarps * 8);
        target = max((int64_t)(FIXED ? CH : 4 * CH), min(target, (int64_t)1 << 20));
        if (FIXED && w.nb > 0 && w.nb <= target)
            w.seg_bytes = (target / w.nb) * w.nb; // whole requests per segment
        else
            w.seg_bytes = (target / CH) * CH;
        w.nseg = w.T > 0 ? (w.T + w.seg_bytes - 1) / w.seg_bytes : 0;
    }
    if (over) {
        if (gwarp == 0 && lane == 0) report(a.status, a.nreq, DDSK_CODE_CAPACITY);
        w.nseg = 0;
    }

    // ---- FIXED with nothing to walk (count == 0, or the batch does not fit): run the reference's two checks here,
    //      so an invalid request is still the error that gets reported
    if (FIXED && (w.nb == 0 || over)) {
        for (int64_t i = gwarp * 32 + lane; i < a.nreq; i += nwarps * 32) {
            uint64_t s;
            int code = dev_locate(a.var, a.starts[i], a.count, &s);
            if (code) report(a.status, i, code);
        }
    }

    // ---- per-warp pipeline -------------------------------------------------------------------
    const uint32_t ring = smem_u32(smem_dyn) + (uint32_t)warp * (uint32_t)(S * STAGE);

    uint32_t issued = 0, consumed = 0;
    bool more = w.nseg > 0;
    while (true) {
        // issue up to S-1 groups ahead
        while (more && issued - consumed < (uint32_t)(S - 1)) {
            Piece pc;
            const uint32_t total = w.template next_group<STAGE>(a, lane, pc);
            if (total == 0) {
                more = false;
                break;
            }
            const uint32_t st = issued % S;
            const uint32_t bar = smem_u32(&full_bar[warp][st]);
            // the stage's previous tenant was drained >= 2 drains ago; the bulk stores a lane issued for it (if any)
            // are at most that lane's second most recent bulk group
            bulk_wait_read<1>();
            if (lane == 0) mbar_expect_tx(bar, total);
            __syncwarp();
            const uint32_t al = (uint32_t)(pc.src & 15u);
            desc[warp][st][lane].dpos = pc.dpos;
            desc[warp][st][lane].n = pc.n;
            desc[warp][st][lane].pack = pc.off | (al << 16);
            if (pc.n) // every lane issues its own piece's TMA load; all complete on the stage's mbarrier
                tma_load_1d(ring + st * STAGE + pc.off, (const void *)(pc.src - al), (al + pc.n + 15u) & ~15u, bar);
            issued++;
        }
        if (consumed == issued) break;
        // drain the oldest group
        const uint32_t st = consumed % S;
        const uint32_t parity = (consumed / S) & 1u;
        const uint32_t bar = smem_u32(&full_bar[warp][st]);
        if (!mbar_try_wait(bar, parity)) {
            const uint64_t t0 = globaltimer_ns();
            while (!mbar_try_wait(bar, parity)) {
                if (globaltimer_ns() - t0 > 4000000000ull) { // 4 s: a lost TMA completion must not hang the box
                    report(a.status, a.nreq, DDSK_CODE_WATCHDOG);
                    __trap();
                }
            }
        }
        __syncwarp();
        const int64_t my_dpos = desc[warp][st][lane].dpos;
        const uint32_t my_n = desc[warp][st][lane].n;
        const uint32_t my_pack = desc[warp][st][lane].pack;
        // Pieces whose staged bytes, destination and size are all 16-byte aligned (every piece of an aligned
        // fixed-stride batch) are stored by their own lane, all lanes at once: one TMA bulk store each, no loop.
        char *const my_dst = dst_of(my_dpos);
        const bool direct = my_n != 0 && (((uint32_t)(uint64_t)my_dst | my_n | (my_pack >> 16)) & 15u) == 0;
        if (direct) tma_store_1d(my_dst, ring + st * STAGE + (my_pack & 0xffffu), my_n);
        // the rest (re-phase, or <16-byte heads/tails) is drained cooperatively, piece by piece
        unsigned todo = __ballot_sync(0xffffffffu, my_n != 0 && !direct);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t dpos = __shfl_sync(0xffffffffu, my_dpos, j);
            const uint32_t n = __shfl_sync(0xffffffffu, my_n, j);
            const uint32_t pk = __shfl_sync(0xffffffffu, my_pack, j);
            drain_chunk<CH>(ring + st * STAGE + (pk & 0xffffu), pk >> 16, dst_of(dpos), n, lane);
        }
        bulk_commit(); // every lane: one (possibly empty) bulk group per drained stage
        __syncwarp();  // all lanes are done reading the stage before it is refilled
        consumed++;
    }
    bulk_wait_read<0>(); // every lane: its stages have been read out; the global writes complete with the grid
    __syncwarp();
    if (multi) { // per-variable byte offsets = plan offsets rebased to the variable's start
        for (int64_t i = gwarp * 32 + lane; i < a.nreq + a.plan.nvars; i += nwarps * 32) {
            // entry (v, j) for j in [0, per_var]: i enumerates nvars * (per_var + 1) slots
            const int v = (int)(i / (a.plan.per_var + 1));
            const int64_t j = i - (int64_t)v * (a.plan.per_var + 1);
            if (v < a.plan.nvars && a.moffsets[v]) {
                const int64_t basev = a.req_dst[(int64_t)v * a.plan.per_var]; // == vbase[v]; read from memory so that
                a.moffsets[v][j] = a.req_dst[(int64_t)v * a.plan.per_var + j] - basev; // vbase[] is never indexed dynamically
            }
        }
    }
    if (FIXED && a.offsets_out) { // arithmetic offsets, written off the critical path
        for (int64_t i = gwarp * 32 + lane; i <= a.nreq; i += nwarps * 32) a.offsets_out[i] = i * w.nb;
    }

    if (!FIXED && a.overlap && lane == 0) { // this warp is done with the slot's plan arrays
        __threadfence();
        atomicAdd(&a.counters[1], 1u);
    }
    // ---- self-resetting ticket counters (not used by overlapped launches) ---------------------
    if (lane == 0 && !a.overlap) {
        __threadfence();
        unsigned int done = atomicAdd(&a.counters[1], 1u);
        if (done == (unsigned int)(nwarps - 1)) {
            a.counters[0] = 0;
            a.counters[1] = 0;
            a.counters[2] = 0;
            a.counters[3] = 0;
            __threadfence();
            if (a.host_mirror) { // the last warp publishes status + total straight into pinned host memory: the host
                                 // reads them after the stream sync, no D2H copy in the call
                a.host_mirror[0] = *(volatile unsigned long long *)a.status;
                a.host_mirror[1] = (unsigned long long)w.T;
                __threadfence_system();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// plan kernels (variable counts): lookup + checks + exclusive scan of request bytes
// ------------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 256;
constexpr int PLAN_ITEMS = 4;
constexpr int PLAN_TILE = PLAN_THREADS * PLAN_ITEMS;
constexpr int PLAN1_THREADS = 1024; // single-CTA plan for small batches
constexpr int PLAN1_ITEMS = 8;
constexpr int PLAN1_MAX = PLAN1_THREADS * PLAN1_ITEMS;

// block-wide exclusive scan of one value per thread (NT threads); returns exclusive prefix, *total = block sum
template <int NT>
__device__ __forceinline__ int64_t block_excl_scan(int64_t v, int64_t *total) {
    __shared__ int64_t warp_tot[NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t inc = warp_incl_scan(v, lane);
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int64_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++) {
        int64_t t = warp_tot[k];
        if (k < wid) base += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

// small batches (<= 8192 requests): the whole plan in ONE CTA, one launch
__global__ void __launch_bounds__(PLAN1_THREADS) dds_plan_single_kernel(const __grid_constant__ ddsk_var_t var,
                                                                        const __grid_constant__ PlanSrc p, int64_t nreq,
                                                                        uint64_t *__restrict__ req_src,
                                                                        int64_t *__restrict__ req_dst,
                                                                        int64_t *__restrict__ offsets_out,
                                                                        unsigned long long *status) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // blocked: thread t owns the ipt (<= 8) consecutive requests [t*ipt, (t+1)*ipt), so every thread has work
    const int ipt = (int)((nreq + PLAN1_THREADS - 1) / PLAN1_THREADS);
    const int64_t base = (int64_t)threadIdx.x * ipt;
    int64_t idx[PLAN1_ITEMS], nb[PLAN1_ITEMS];
    uint64_t sv[PLAN1_ITEMS];
#pragma unroll
    for (int k = 0; k < PLAN1_ITEMS; k++) idx[k] = k < ipt ? base + k : nreq;
    plan_many<PLAN1_ITEMS>(var, p, idx, nreq, status, sv, nb);
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < PLAN1_ITEMS; k++) {
        if (idx[k] < nreq) {
            req_src[idx[k]] = sv[k];
            mine += nb[k];
        }
    }
    int64_t tot;
    int64_t run = block_excl_scan<PLAN1_THREADS>(mine, &tot);
#pragma unroll
    for (int k = 0; k < PLAN1_ITEMS; k++) {
        if (idx[k] < nreq) {
            req_dst[idx[k]] = run;
            if (offsets_out) offsets_out[idx[k]] = run;
            run += nb[k];
        }
    }
    if (threadIdx.x == 0) {
        req_dst[nreq] = tot;
        if (offsets_out) offsets_out[nreq] = tot;
    }
}

// pass 1 (large batches): per request source address + byte size (size parked in req_dst), per tile byte sum
__global__ void __launch_bounds__(PLAN_THREADS) dds_plan_lookup_kernel(const __grid_constant__ ddsk_var_t var,
                                                                       const __grid_constant__ PlanSrc p, int64_t nreq,
                                                                       uint64_t *__restrict__ req_src,
                                                                       int64_t *__restrict__ req_dst,
                                                                       int64_t *__restrict__ tile_sums,
                                                                       unsigned long long *status) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int64_t base = (int64_t)blockIdx.x * PLAN_TILE;
    int64_t idx[PLAN_ITEMS], nb[PLAN_ITEMS];
    uint64_t sv[PLAN_ITEMS];
#pragma unroll
    for (int k = 0; k < PLAN_ITEMS; k++) idx[k] = base + (int64_t)k * PLAN_THREADS + threadIdx.x; // striped
    plan_many<PLAN_ITEMS>(var, p, idx, nreq, status, sv, nb);
    int64_t mine = 0;
#pragma unroll
    for (int k = 0; k < PLAN_ITEMS; k++) {
        if (idx[k] < nreq) {
            req_src[idx[k]] = sv[k];
            req_dst[idx[k]] = nb[k];
            mine += nb[k];
        }
    }
    int64_t tot;
    block_excl_scan<PLAN_THREADS>(mine, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// pass 2: exclusive scan. Each CTA sums the tiles before it, then scans its own tile in place.
__global__ void __launch_bounds__(PLAN_THREADS) dds_plan_scan_kernel(int64_t nreq, int64_t *__restrict__ req_dst,
                                                                     const int64_t *__restrict__ tile_sums,
                                                                     int64_t *__restrict__ offsets_out) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int64_t part = 0;
    for (int64_t t = threadIdx.x; t < (int64_t)blockIdx.x; t += PLAN_THREADS) part += tile_sums[t];
    int64_t tile_base;
    block_excl_scan<PLAN_THREADS>(part, &tile_base);
    // layout inside a tile is striped (item k of thread t = k*THREADS + t): scan stripe by stripe
    const int64_t base = (int64_t)blockIdx.x * PLAN_TILE;
    int64_t running = tile_base;
#pragma unroll
    for (int k = 0; k < PLAN_ITEMS; k++) {
        int64_t i = base + (int64_t)k * PLAN_THREADS + threadIdx.x;
        int64_t v = i < nreq ? req_dst[i] : 0;
        int64_t tot;
        int64_t ex = block_excl_scan<PLAN_THREADS>(v, &tot);
        if (i < nreq) {
            req_dst[i] = running + ex;
            if (offsets_out) offsets_out[i] = running + ex;
        }
        running += tot;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        req_dst[nreq] = running;
        if (offsets_out) offsets_out[nreq] = running;
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic payload generator (bench / test helper)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <typename T>
__global__ void dds_synth_kernel(T *__restrict__ base, uint64_t first_elem, uint64_t nelem, uint64_t seed) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nelem; i += (uint64_t)gridDim.x * blockDim.x)
        base[i] = (T)splitmix64(seed ^ (first_elem + i));
}

// ------------------------------------------------------------------------------------------------
// launch geometry
// ------------------------------------------------------------------------------------------------
struct Geometry {
    int nw, stages, ch;
};
constexpr Geometry kGeoms[] = {{8, 4, 4096}, {8, 6, 4096}, {16, 3, 4096}, {4, 4, 8192}, {12, 4, 4096}, {4, 6, 4096}};
constexpr int kNumGeoms = (int)(sizeof(kGeoms) / sizeof(kGeoms[0]));

// Measured on B200 (profiles/r1_configs.md): 12 warps x 4 stages is as fast as 8 x 4 on 4 KiB+ rows and clearly
// faster on the instruction-heavier variable / re-phase path; rows under 2 KiB want even more warps (16 x 3).
constexpr int kGeomLarge = 4, kGeomSmall = 2, kGeomVar = 4;

int g_geom_fixed_env = -1; // DDS_GATHER_GEOM      (tuning: force one variant for the fixed-count entry)
int g_geom_var_env = -1;   // DDS_GATHER_GEOM_VAR  (... for the variable-count entry; defaults to the former)
bool g_geom_init = false;
int g_sms = 0;
int g_ctas_per_sm = 1;
int g_pdl = 1;
int g_fused_plan = 1; // DDS_FUSED_PLAN: 1 = auto (fused for <= 8192 requests), 0 = never, 2 = always (A/B switch)

int pick_geometry() {
    if (g_geom_init) return 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    if (const char *e = getenv("DDS_GATHER_GEOM")) g_geom_fixed_env = atoi(e);
    if (g_geom_fixed_env >= kNumGeoms) g_geom_fixed_env = -1;
    g_geom_var_env = g_geom_fixed_env;
    if (const char *e = getenv("DDS_GATHER_GEOM_VAR")) g_geom_var_env = atoi(e);
    if (g_geom_var_env >= kNumGeoms) g_geom_var_env = -1;
    if (const char *e = getenv("DDS_GATHER_CTAS_PER_SM")) g_ctas_per_sm = atoi(e) > 0 ? atoi(e) : 1;
    if (const char *e = getenv("DDS_PDL")) g_pdl = atoi(e) != 0;
    if (const char *e = getenv("DDS_FUSED_PLAN")) g_fused_plan = atoi(e);
    g_geom_init = true;
    return 0;
}

int geometry_for(bool fixed, int64_t request_bytes) {
    if (fixed) {
        if (g_geom_fixed_env >= 0) return g_geom_fixed_env;
        return request_bytes < 2048 ? kGeomSmall : kGeomLarge;
    }
    return g_geom_var_env >= 0 ? g_geom_var_env : kGeomVar;
}

template <bool FIXED, int NW, int S, int CH>
int launch_gather_t(const GatherArgs &args, cudaStream_t stream) {
    constexpr int smem = NW * S * (CH + 32);
    static std::atomic<unsigned long long> configured{0}; // bit d: attribute set on device d (it is per device)
    auto kern = dds_gather_kernel<FIXED, NW, S, CH>;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 64 || !(configured.load() & (1ull << dev))) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 64) configured.fetch_or(1ull << dev);
    }
    int per_sm = g_ctas_per_sm;
    while (per_sm > 1 && per_sm * (smem + 2048) > 227 * 1024) per_sm--;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(g_sms * per_sm));
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args));
    g_launches++;
    return 0;
}

// launch with the programmatic-dependent-launch attribute (the kernels call griddepcontrol.wait themselves)
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
    g_launches++;
    return 0;
}

template <bool FIXED>
int launch_gather(const GatherArgs &args, cudaStream_t stream) {
    if (int rc = pick_geometry()) return rc;
    switch (geometry_for(FIXED, FIXED ? args.count * args.var.row_bytes : 0)) {
    case 1: return launch_gather_t<FIXED, 8, 6, 4096>(args, stream);
    case 2: return launch_gather_t<FIXED, 16, 3, 4096>(args, stream);
    case 3: return launch_gather_t<FIXED, 4, 4, 8192>(args, stream);
    case 4: return launch_gather_t<FIXED, 12, 4, 4096>(args, stream);
    case 5: return launch_gather_t<FIXED, 4, 6, 4096>(args, stream);
    default: return launch_gather_t<FIXED, 8, 4, 4096>(args, stream);
    }
}

} // namespace

// ------------------------------------------------------------------------------------------------
// the thin C-ABI the host C++ calls
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *ddsk_last_cuda_error(void) { return g_cuda_err; }
unsigned long long ddsk_launch_count(void) { return g_launches.load(); }

void ddsk_gather_geometry(int *ctas, int *warps_per_cta, int *stages, int *chunk_bytes, int *smem_bytes) {
    if (pick_geometry()) {
        *ctas = *warps_per_cta = *stages = *chunk_bytes = *smem_bytes = 0;
        return;
    }
    const Geometry &g = kGeoms[geometry_for(true, 4096)];
    int smem = g.nw * g.stages * (g.ch + 32);
    int per_sm = g_ctas_per_sm;
    while (per_sm > 1 && per_sm * (smem + 2048) > 227 * 1024) per_sm--;
    *ctas = g_sms * per_sm;
    *warps_per_cta = g.nw;
    *stages = g.stages;
    *chunk_bytes = g.ch;
    *smem_bytes = smem;
}

int ddsk_gather_fixed(const ddsk_var_t *var, const int64_t *starts_dev, int64_t count, int64_t nreq, void *dst_dev,
                      int64_t dst_capacity, int64_t *offsets_dev_or_null, const ddsk_scratch_t *scr, int reset_status,
                      void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (reset_status & 1) CUDA_TRY(cudaMemsetAsync(scr->status, 0xFF, sizeof(unsigned long long), st));
    if (nreq <= 0) return 0;
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    a.var = *var;
    a.starts = starts_dev;
    a.count = count;
    a.nreq = nreq;
    a.dst = (char *)dst_dev;
    a.dst_cap = dst_capacity;
    a.offsets_out = offsets_dev_or_null;
    a.status = scr->status;
    a.overlap = (reset_status & 4) ? 1 : 0;    // bit 2: independent batch -> static segment striding
    a.skip_wait = (reset_status & 16) ? 1 : 0; // bit 4: ... whose predecessor was one too -> no grid wait
    a.counters = scr->counters;
    a.host_mirror = reset_status & 2 ? scr->host_mirror : nullptr; // bit 1 of the flags word: mirror wanted
    return launch_gather<true>(a, st);
}

int ddsk_gather_var(const ddsk_var_t *var, const ddsk_index_t *index, int64_t nreq, void *dst_dev, int64_t dst_capacity,
                    int64_t *offsets_dev_or_null, ddsk_scratch_t *scr, int reset_status, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (reset_status & 1) CUDA_TRY(cudaMemsetAsync(scr->status, 0xFF, sizeof(unsigned long long), st));
    if (nreq <= 0) return 0;
    if (nreq > scr->cap_req) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "ddsk_gather_var: scratch too small (%lld > %lld)", (long long)nreq,
                 (long long)scr->cap_req);
        return -2;
    }
    if (int rc = pick_geometry()) return rc;
    PlanSrc p;
    memset(&p, 0, sizeof(p));
    p.starts = index->starts;
    p.counts = index->counts;
    p.ids = index->sample_ids;
    p.tab_start = index->table_start;
    p.tab_count = index->table_count;
    p.nsamples = index->nsamples;
    // measured (profiles/r1_configs.md): the in-kernel plan wins below ~8K requests (one launch instead of two or
    // three: B=4096 54 -> 48 us), separate plan kernels are ~2 % faster above (they overlap the previous gather's tail)
    const bool ovl = (reset_status & 4) != 0; // independent batch in its own scratch slot: always planned in-kernel
    const bool fused = ovl || (g_fused_plan == 1 ? nreq <= 8192 : g_fused_plan == 2);
    if (!fused) {
        // one CTA is enough for explicit (start, count) arrays (coalesced loads); the sample-index lookups are random
        // two-level gathers and want more SMs' worth of memory parallelism (measured: 1 CTA costs +25 us at B=4096)
        if (nreq <= (p.ids ? (int64_t)PLAN_TILE : (int64_t)PLAN1_MAX)) {
            if (int rc = launch_pdl(dds_plan_single_kernel, dim3(1), dim3(PLAN1_THREADS), st, *var, p, nreq, scr->req_src,
                                    scr->req_dst, offsets_dev_or_null, scr->status))
                return rc;
        } else {
            const int tiles = (int)((nreq + PLAN_TILE - 1) / PLAN_TILE);
            if (int rc = launch_pdl(dds_plan_lookup_kernel, dim3(tiles), dim3(PLAN_THREADS), st, *var, p, nreq,
                                    scr->req_src, scr->req_dst, scr->tile_sums, scr->status))
                return rc;
            if (int rc = launch_pdl(dds_plan_scan_kernel, dim3(tiles), dim3(PLAN_THREADS), st, nreq, scr->req_dst,
                                    (const int64_t *)scr->tile_sums, offsets_dev_or_null))
                return rc;
        }
    }
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    a.var = *var;
    a.req_src = scr->req_src;
    a.req_dst = scr->req_dst;
    a.nreq = nreq;
    a.dst = (char *)dst_dev;
    a.dst_cap = dst_capacity;
    a.status = scr->status;
    a.counters = scr->counters;
    a.host_mirror = reset_status & 2 ? scr->host_mirror : nullptr;
    if (fused) {
        a.plan = p;
        a.tile_state = (unsigned long long *)scr->tile_sums;
        scr->epoch = (scr->epoch + 1) & 0x3FFFFFu;
        if (scr->epoch == 0) { // 22-bit tag wrapped: clear the words so a stale tag can never look current
            CUDA_TRY(cudaMemsetAsync(scr->tile_sums, 0, (size_t)(scr->cap_req / 128 + 2) * 8, st));
            scr->epoch = 1;
        }
        a.epoch = scr->epoch;
        a.fused_plan = 1;
        a.offsets_out = offsets_dev_or_null;
    }
    if (ovl) {
        a.overlap = 1;
        a.skip_wait = (reset_status & 16) ? 1 : 0;
        a.ticket_base = scr->ticket_base;
        a.tiles_base = scr->tiles_base;
        a.finish_target = scr->finish_target;
        const Geometry &g = kGeoms[geometry_for(false, 0)];
        int per_sm = g_ctas_per_sm;
        while (per_sm > 1 && per_sm * (g.nw * g.stages * (g.ch + 32) + 2048) > 227 * 1024) per_sm--;
        const unsigned int nwarps = (unsigned int)(g_sms * per_sm * g.nw);
        const unsigned int ntiles = (unsigned int)((nreq + TILE_REQ - 1) / TILE_REQ);
        scr->ticket_base += ntiles + nwarps; // every warp makes exactly one failing ticket claim
        scr->tiles_base += ntiles;
        scr->finish_target += nwarps;
    }
    return launch_gather<false>(a, st);
}

int ddsk_gather_multi(const ddsk_multi_t *m, const int64_t *sample_ids_dev, int64_t nreq, ddsk_scratch_t *scr, int flags,
                      void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (flags & 1) CUDA_TRY(cudaMemsetAsync(scr->status, 0xFF, sizeof(unsigned long long), st));
    if (nreq <= 0 || m->nvars <= 0) return 0;
    if (m->nvars > DDSK_MAX_MULTI) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "ddsk_gather_multi: more than %d variables", DDSK_MAX_MULTI);
        return -2;
    }
    const int64_t total_req = nreq * m->nvars;
    if (total_req > scr->cap_req) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "ddsk_gather_multi: scratch too small (%lld > %lld)", (long long)total_req,
                 (long long)scr->cap_req);
        return -2;
    }
    if (int rc = pick_geometry()) return rc;
    PlanSrc p;
    memset(&p, 0, sizeof(p));
    p.ids = sample_ids_dev;
    p.nvars = m->nvars;
    p.per_var = nreq;
    p.mvars = m->vars_dev;
    for (int v = 0; v < m->nvars; v++) {
        p.mtab_start[v] = m->table_start[v];
        p.mtab_count[v] = m->table_count[v];
        p.mnsamples[v] = m->nsamples[v];
    }
    ddsk_var_t dummy;
    memset(&dummy, 0, sizeof(dummy));
    const bool fused = g_fused_plan == 1 ? total_req <= 8192 : g_fused_plan == 2;
    if (!fused) {
        const int tiles = (int)((total_req + PLAN_TILE - 1) / PLAN_TILE);
        if (int rc = launch_pdl(dds_plan_lookup_kernel, dim3(tiles), dim3(PLAN_THREADS), st, dummy, p, total_req,
                                scr->req_src, scr->req_dst, scr->tile_sums, scr->status))
            return rc;
        if (int rc = launch_pdl(dds_plan_scan_kernel, dim3(tiles), dim3(PLAN_THREADS), st, total_req, scr->req_dst,
                                (const int64_t *)scr->tile_sums, (int64_t *)nullptr))
            return rc;
    }
    GatherArgs a;
    memset(&a, 0, sizeof(a));
    a.req_src = scr->req_src;
    a.req_dst = scr->req_dst;
    a.nreq = total_req;
    a.dst = nullptr;
    a.dst_cap = INT64_MAX;
    a.status = scr->status;
    a.counters = scr->counters;
    a.host_mirror = flags & 2 ? scr->host_mirror : nullptr;
    a.plan = p; // the gather needs nvars / per_var even when the plan ran in its own kernels
    for (int v = 0; v < m->nvars; v++) {
        a.mdst[v] = (char *)m->dst[v];
        a.mcap[v] = m->cap[v];
        a.moffsets[v] = m->offsets[v];
    }
    if (fused) {
        a.tile_state = (unsigned long long *)scr->tile_sums;
        scr->epoch = (scr->epoch + 1) & 0x3FFFFFu;
        if (scr->epoch == 0) {
            CUDA_TRY(cudaMemsetAsync(scr->tile_sums, 0, (size_t)(scr->cap_req / 128 + 2) * 8, st));
            scr->epoch = 1;
        }
        a.epoch = scr->epoch;
        a.fused_plan = 1;
    }
    return launch_gather<false>(a, st);
}

int ddsk_synth_fill(void *base_dev, int64_t first_global_row, int64_t nrows, int64_t disp, int itemsize, uint64_t seed,
                    void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t nelem = (uint64_t)nrows * (uint64_t)disp;
    uint64_t first = (uint64_t)first_global_row * (uint64_t)disp;
    if (nelem == 0) return 0;
    int blocks = (int)((nelem + 255) / 256 < 148 * 16 ? (nelem + 255) / 256 : 148 * 16);
    switch (itemsize) {
    case 1: dds_synth_kernel<uint8_t><<<blocks, 256, 0, st>>>((uint8_t *)base_dev, first, nelem, seed); break;
    case 2: dds_synth_kernel<uint16_t><<<blocks, 256, 0, st>>>((uint16_t *)base_dev, first, nelem, seed); break;
    case 4: dds_synth_kernel<uint32_t><<<blocks, 256, 0, st>>>((uint32_t *)base_dev, first, nelem, seed); break;
    case 8: dds_synth_kernel<uint64_t><<<blocks, 256, 0, st>>>((uint64_t *)base_dev, first, nelem, seed); break;
    default:
        snprintf(g_cuda_err, sizeof(g_cuda_err), "ddsk_synth_fill: unsupported itemsize %d", itemsize);
        return -2;
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

} // extern "C"
