// scripts/probes/gather4_probe.cu -- measurement probe (not product code): for rows of 512 B, does the sm_100 TMA
// tile::gather4 form (4 rows of a 2-D tensor map per instruction) beat the 1-D bulk copies the store's gather kernel uses
// (one cp.async.bulk per row)? Random rows of a local 4 GiB buffer are packed into a contiguous output, per-warp ring of
// S stages x 4 KiB (8 rows per stage), one 4 KiB bulk store per stage in both variants; only the loads differ:
//   mode 0: 8 lanes issue one 512 B cp.async.bulk each            (what dds_gather_kernel's packed groups do)
//   mode 1: lane 0 issues two cp.async.bulk.tensor.2d.tile::gather4 (4 rows each)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_probe gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int ROW = 512, RPS = 8, STAGE = ROW * RPS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int S, int MODE>
__global__ void k_rows(const char *__restrict__ src, const __grid_constant__ CUtensorMap tmap, char *__restrict__ dst,
                       const int *__restrict__ rows, int nrows) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bars[16][S];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nwb = blockDim.x >> 5;
    const int gw = blockIdx.x * nwb + w, nw = gridDim.x * nwb;
    const uint32_t ring = smem_u32(sm) + w * S * STAGE;
    if (lane == 0) {
        for (int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[w][s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int ngroups = nrows / RPS;
    int issued = 0, consumed = 0, next = gw;
    while (true) {
        while (next < ngroups && issued - consumed < S - 1) {
            const int st = issued % S;
            const uint32_t bar = smem_u32(&bars[w][st]);
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE) : "memory");
            }
            __syncwarp();
            if (MODE == 0) {
                if (lane < RPS) {
                    const char *s = src + (size_t)rows[next * RPS + lane] * ROW;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + st * STAGE + lane * ROW), "l"(s), "r"(ROW), "r"(bar) : "memory");
                }
            } else {
                const int r = lane < RPS ? rows[next * RPS + lane] : 0;
                const int r0 = __shfl_sync(~0u, r, 0), r1 = __shfl_sync(~0u, r, 1), r2 = __shfl_sync(~0u, r, 2), r3 = __shfl_sync(~0u, r, 3);
                const int r4 = __shfl_sync(~0u, r, 4), r5 = __shfl_sync(~0u, r, 5), r6 = __shfl_sync(~0u, r, 6), r7 = __shfl_sync(~0u, r, 7);
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(ring + st * STAGE), "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(ring + st * STAGE + 4 * ROW), "l"(&tmap), "r"(0), "r"(r4), "r"(r5), "r"(r6), "r"(r7), "r"(bar) : "memory");
                }
            }
            issued++;
            next += nw;
        }
        if (consumed == issued) break;
        const int st = consumed % S;
        const uint32_t par = (consumed / S) & 1, bar = smem_u32(&bars[w][st]);
        uint32_t ok = 0;
        long spins = 0;
        while (!ok) {
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
            if (++spins > 200000000L) { if (lane == 0 && gw == 0) printf("timeout\n"); __trap(); }
        }
        __syncwarp();
        if (lane == 0) {
            const int g = gw + consumed * nw;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)g * STAGE), "r"(ring + st * STAGE), "r"(STAGE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
        consumed++;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
    const size_t buf_bytes = 4ull << 30;
    const int total_rows = (int)(buf_bytes / ROW), nrows = 262144;
    const int boxh = argc > 1 ? atoi(argv[1]) : 1; // box height of the tensor map (1 or 4): which one gather4 wants
    char *buf, *out, *out2;
    int *rows;
    CK(cudaMalloc(&buf, buf_bytes)); CK(cudaMalloc(&out, (size_t)nrows * ROW)); CK(cudaMalloc(&out2, (size_t)nrows * ROW));
    CK(cudaMalloc(&rows, nrows * 4));
    std::vector<uint32_t> h(buf_bytes / 4);
    for (size_t i = 0; i < h.size(); i += 1024) h[i] = (uint32_t)(i * 2654435761u);
    CK(cudaMemcpy(buf, h.data(), buf_bytes, cudaMemcpyHostToDevice));
    std::vector<int> hr(nrows); srand(5); for (auto &x : hr) x = (int)(((uint64_t)rand() * 1315423911ull) % total_rows);
    CK(cudaMemcpy(rows, hr.data(), nrows * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t gdim[2] = {ROW / 4, (cuuint64_t)total_rows}, gstr[1] = {ROW};
    cuuint32_t box[2] = {ROW / 4, (cuuint32_t)boxh}, estr[2] = {1, 1};
    CUresult cr = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 1; }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](int mode, int nwarps, char *o) -> float {
        constexpr int S = 4;
        const int smem = nwarps * S * STAGE;
        auto k0 = k_rows<S, 0>; auto k1 = k_rows<S, 1>;
        CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        auto launch = [&]() { if (mode == 0) k0<<<148, nwarps * 32, smem>>>(buf, tm, o, rows, nrows); else k1<<<148, nwarps * 32, smem>>>(buf, tm, o, rows, nrows); };
        for (int i = 0; i < 3; i++) launch();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; i++) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        return (float)((double)nrows * ROW * 20 / (ms * 1e-3) / 1e9);
    };
    for (int nw : {8, 12}) {
        float a = run(0, nw, out), b = run(1, nw, out2);
        std::vector<char> x((size_t)nrows * ROW), y((size_t)nrows * ROW);
        CK(cudaMemcpy(x.data(), out, x.size(), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(y.data(), out2, y.size(), cudaMemcpyDeviceToHost));
        printf("512 B rows, B=%d, %2d warps/SM x 4 stages x 4 KiB, box height %d: bulk-per-row %7.1f GB/s | tile::gather4 %7.1f GB/s | outputs %s\n",
               nrows, nw, boxh, a, b, memcmp(x.data(), y.data(), x.size()) == 0 ? "identical" : "DIFFER");
    }
    return 0;
}
