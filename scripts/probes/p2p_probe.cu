// scripts/probes/p2p_probe.cu -- measurement probe (not product code): peer-read bandwidth over NVLink for the
// access patterns the gather kernel can use. GPU0 (and optionally GPU1 at the same time) reads random 4 KiB rows
// of the OTHER GPU's buffer and writes them to a local packed buffer.
//   mode 0: LDG.128 (ld.global.nc.v4) -> STG.128, warp per row, UNROLL independent loads
//   mode 1: TMA bulk load global->shared (mbarrier) -> TMA bulk store shared->global, per-warp ring (as in kernels.cu)
//   mode 2: cp.async 16 B (LDGSTS) global->shared -> TMA bulk store
// Round 2 adds the OWNER-PUSH pattern: a GPU bulk-loads random LOCAL rows and bulk-STORES them into the peer's packed
// buffer (posted writes: the reverse direction carries acks instead of read requests), one way and both ways at once,
// a half-pull / half-push mix, and cudaMemcpyPeerAsync in both directions at once as the copy-engine reference.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_probe p2p_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <thread>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_ldg(const uint4 *__restrict__ src, uint4 *__restrict__ dst, const int *__restrict__ rows, int nrows, int row_vec) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < nrows; r += nw) {
        const uint4 *s = src + (size_t)rows[r] * row_vec;
        uint4 *d = dst + (size_t)r * row_vec;
        for (int j = lane; j < row_vec; j += 32 * 8) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int jj = j + u * 32;
                if (jj < row_vec)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + jj));
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int jj = j + u * 32;
                if (jj < row_vec) d[jj] = v[u];
            }
        }
    }
}

// owner-push with plain stores: LDG.128 from LOCAL rows, STG.128 into the destination (the peer's packed buffer)
__global__ void k_push_stg(const uint4 *__restrict__ src, uint4 *dst, const int *__restrict__ rows, int nrows, int row_vec) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < nrows; r += nw) {
        const uint4 *s = src + (size_t)rows[r] * row_vec;
        uint4 *d = dst + (size_t)r * row_vec;
        for (int j = lane; j < row_vec; j += 32 * 8) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int jj = j + u * 32;
                if (jj < row_vec) v[u] = s[jj];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int jj = j + u * 32;
                if (jj < row_vec) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(d + jj), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
            }
        }
    }
}

template <int S>
__global__ void k_tma(const char *__restrict__ src, char *__restrict__ dst, const int *__restrict__ rows, int nrows, int row_bytes, int use_cpasync) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bars[32][S];
    int warp_in = threadIdx.x >> 5, lane = threadIdx.x & 31, nwb = blockDim.x >> 5;
    int warp = blockIdx.x * nwb + warp_in, nw = gridDim.x * nwb;
    uint32_t ring = smem_u32(sm) + warp_in * S * row_bytes;
    if (lane == 0) {
        for (int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp_in][s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    int issued = 0, consumed = 0;
    int next = warp;
    while (true) {
        while (next < nrows && issued - consumed < S - 1) {
            int st = issued % S;
            const char *s = src + (size_t)rows[next] * row_bytes;
            uint32_t bar = smem_u32(&bars[warp_in][st]);
            if (!use_cpasync) {
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + st * row_bytes), "l"(s), "r"(row_bytes), "r"(bar) : "memory");
                }
            } else {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                for (int o = lane * 16; o < row_bytes; o += 512)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + st * row_bytes + o), "l"(s + o) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
            }
            issued++;
            next += nw;
        }
        if (consumed == issued) break;
        int st = consumed % S;
        uint32_t par = (consumed / S) & 1, bar = smem_u32(&bars[warp_in][st]);
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
        __syncwarp();
        if (lane == 0) {
            int r = warp + consumed * nw;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)r * row_bytes), "r"(ring + st * row_bytes), "r"(row_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
        consumed++;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Side { int dev; char *buf; char *out; char *out2; int *rows; cudaStream_t st; cudaEvent_t e0, e1; };

int main(int argc, char **argv) {
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t buf_bytes = 8ull << 30; const int row_bytes = 4096; const int nrows = 65536;
    const int total_rows = buf_bytes / row_bytes;
    Side sd[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); cudaDeviceEnablePeerAccess(1 - d, 0);
        sd[d].dev = d; CK(cudaMalloc(&sd[d].buf, buf_bytes)); CK(cudaMemset(sd[d].buf, d + 1, buf_bytes));
        CK(cudaMalloc(&sd[d].out, (size_t)nrows * row_bytes)); CK(cudaMalloc(&sd[d].out2, (size_t)nrows * row_bytes));
        CK(cudaMalloc(&sd[d].rows, nrows * 4));
        std::vector<int> h(nrows); srand(7 + d); for (auto &x : h) x = (int)(((uint64_t)rand() * 1315423911ull) % total_rows);
        CK(cudaMemcpy(sd[d].rows, h.data(), nrows * 4, cudaMemcpyHostToDevice));
        CK(cudaStreamCreate(&sd[d].st)); CK(cudaEventCreate(&sd[d].e0)); CK(cudaEventCreate(&sd[d].e1));
    }
    // remote: 0 = local rows -> local out, 1 = PULL (peer rows -> local out), 2 = PUSH (local rows -> the peer's out)
    auto run = [&](int d, int mode, int remote, int nwarps, int S, int iters) -> float {
        CK(cudaSetDevice(d));
        const char *src = remote == 1 ? sd[1 - d].buf : sd[d].buf;
        char *outp = remote == 2 ? sd[1 - d].out2 : sd[d].out;
        int smem = nwarps * S * row_bytes;
        auto launch = [&]() {
            if (mode == 3) k_push_stg<<<148 * 2, 512, 0, sd[d].st>>>((const uint4 *)src, (uint4 *)outp, sd[d].rows, nrows, row_bytes / 16);
            else if (mode == 0) k_ldg<<<148 * 2, 512, 0, sd[d].st>>>((const uint4 *)src, (uint4 *)outp, sd[d].rows, nrows, row_bytes / 16);
            else if (S == 4) { cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_tma<4><<<148, nwarps * 32, smem, sd[d].st>>>(src, outp, sd[d].rows, nrows, row_bytes, mode == 2); }
            else { cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_tma<2><<<148, nwarps * 32, smem, sd[d].st>>>(src, outp, sd[d].rows, nrows, row_bytes, mode == 2); }
        };
        for (int i = 0; i < 3; i++) launch();
        CK(cudaEventRecord(sd[d].e0, sd[d].st));
        for (int i = 0; i < iters; i++) launch();
        CK(cudaEventRecord(sd[d].e1, sd[d].st));
        CK(cudaEventSynchronize(sd[d].e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, sd[d].e0, sd[d].e1));
        return (float)((double)nrows * row_bytes * iters / (ms * 1e-3) / 1e9);
    };
    const char *mn[] = {"LDG.128", "TMA bulk", "cp.async16"};
    const char *what = argc > 1 ? argv[1] : "pull";
    if (!strcmp(what, "push_tma") || !strcmp(what, "push_stg")) {
        const int pm = !strcmp(what, "push_tma") ? 1 : 3, nw = 12, S = 4;
        float pull_uni = run(0, 1, 1, nw, S, 20), push_uni = run(0, pm, 2, nw, S, 20);
        printf("%s: PULL one way %7.1f | PUSH one way %7.1f GB/s\n", what, pull_uni, push_uni); fflush(stdout);
        float b[2], c[2];
        { std::thread t0([&] { b[0] = run(0, pm, 2, nw, S, 20); }), t1([&] { b[1] = run(1, pm, 2, nw, S, 20); }); t0.join(); t1.join(); }
        printf("%s: PUSH both ways %7.1f + %7.1f GB/s\n", what, b[0], b[1]); fflush(stdout);
        // GPU0 pulls from GPU1 while GPU1 pushes to GPU0: all payload flows 1 -> 0 (requests + acks flow 0 -> 1)
        { std::thread t0([&] { c[0] = run(0, 1, 1, nw, S, 20); }), t1([&] { c[1] = run(1, pm, 2, nw, S, 20); }); t0.join(); t1.join(); }
        printf("%s: pull + push in the SAME direction %7.1f + %7.1f GB/s\n", what, c[0], c[1]);
        return 0;
    }
    if (!strcmp(what, "copy")) {
        float r[2];
        auto cp = [&](int d) {
            CK(cudaSetDevice(d));
            CK(cudaEventRecord(sd[d].e0, sd[d].st));
            for (int i = 0; i < 10; i++) CK(cudaMemcpyPeerAsync(sd[d].out, d, sd[1 - d].buf, 1 - d, (size_t)nrows * row_bytes, sd[d].st));
            CK(cudaEventRecord(sd[d].e1, sd[d].st)); CK(cudaEventSynchronize(sd[d].e1));
            float ms; CK(cudaEventElapsedTime(&ms, sd[d].e0, sd[d].e1));
            r[d] = (float)((double)nrows * row_bytes * 10 / (ms * 1e-3) / 1e9);
        };
        cp(0);
        printf("cudaMemcpyPeerAsync 256 MiB contiguous, one direction: %.1f GB/s\n", r[0]);
        std::thread t0([&] { cp(0); }), t1([&] { cp(1); }); t0.join(); t1.join();
        printf("cudaMemcpyPeerAsync 256 MiB contiguous, both directions at once: %.1f + %.1f GB/s\n", r[0], r[1]);
        return 0;
    }
    for (int mode = 0; mode < 3; mode++)
        for (int nw : {8, 12}) for (int S : {4, 2}) {
            if (mode == 0 && (nw != 8 || S != 4)) continue;
            float loc = run(0, mode, 0, nw, S, 20);
            float uni = run(0, mode, 1, nw, S, 20);
            float bi[2];
            std::thread t0([&] { bi[0] = run(0, mode, 1, nw, S, 20); }), t1([&] { bi[1] = run(1, mode, 1, nw, S, 20); });
            t0.join(); t1.join();
            printf("%-10s warps/SM %2d stages %d : local %7.1f GB/s | remote uni %7.1f GB/s | remote bidir %7.1f + %7.1f GB/s\n", mn[mode], nw, S, loc, uni, bi[0], bi[1]);
        }
    // cudaMemcpyPeer reference
    CK(cudaSetDevice(0));
    CK(cudaEventRecord(sd[0].e0, sd[0].st));
    for (int i = 0; i < 10; i++) CK(cudaMemcpyPeerAsync(sd[0].out, 0, sd[1].buf, 1, (size_t)nrows * row_bytes, sd[0].st));
    CK(cudaEventRecord(sd[0].e1, sd[0].st)); CK(cudaEventSynchronize(sd[0].e1));
    float ms; CK(cudaEventElapsedTime(&ms, sd[0].e0, sd[0].e1));
    printf("cudaMemcpyPeerAsync 256 MiB contiguous: %.1f GB/s\n", (double)nrows * row_bytes * 10 / (ms * 1e-3) / 1e9);
    return 0;
}
