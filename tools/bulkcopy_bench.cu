// Dev microbenchmark (not part of the product): per-SM and chip-wide throughput of 1-D TMA bulk copies
// (cp.async.bulk global -> shared, mbarrier complete_tx) as a function of the copy size and of the number of copies
// in flight, next to 16-byte cp.async (LDGSTS.128) issued by one warp.  Decides how the pencil sweeps stage their rows.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulkcopy_bench bulkcopy_bench.cu && ./bulkcopy_bench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbarExpect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbarTry(uint32_t b, uint32_t ph)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(b) : "memory");
}

// mode 0: stages of `ncopy` bulk copies of `size` bytes each (issued by lanes 0..ncopy-1 of warp 0), `nstage` stages in flight
template <int MODE>
__global__ void k(const char* __restrict__ src, size_t perCta, int size, int ncopy, int nstage, int iters, unsigned long long* out)
{
    extern __shared__ __align__(128) char sm[];
    __shared__ unsigned long long bar[16];
    const int lane = threadIdx.x & 31;
    const char* base = src + (size_t)blockIdx.x * perCta;
    const int stageBytes = size * ncopy;
    if (threadIdx.x < nstage) mbarInit(s32(&bar[threadIdx.x]), MODE == 0 ? 1 : 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned long long t0 = clock64();
    if (threadIdx.x < 32) {
        if (MODE == 0) {
            size_t off = 0;
            // prologue
            for (int s = 0; s < nstage; ++s) {
                if (lane == 0) mbarExpect(s32(&bar[s]), stageBytes);
                __syncwarp();
                if (lane < ncopy) bulk(s32(sm + s * stageBytes + lane * size), base + off + (size_t)lane * size, size, s32(&bar[s]));
                off += stageBytes;
                if (off + stageBytes > perCta) off = 0;
            }
            int st = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                while (!mbarTry(s32(&bar[st]), ph)) {}
                // consume: one 8-byte load per lane so the data is really there
                volatile double* p = (volatile double*)(sm + st * stageBytes);
                double v = p[lane];
                if (v == 123.456) out[1] = 1;
                __syncwarp();
                if (it + nstage < iters) {
                    if (lane == 0) mbarExpect(s32(&bar[st]), stageBytes);
                    __syncwarp();
                    if (lane < ncopy) bulk(s32(sm + st * stageBytes + lane * size), base + off + (size_t)lane * size, size, s32(&bar[st]));
                    off += stageBytes;
                    if (off + stageBytes > perCta) off = 0;
                }
                if (++st == nstage) { st = 0; ph ^= 1; }
            }
        } else {
            // 16-byte cp.async per lane: stageBytes / 512 instructions per stage, completion through cp.async.mbarrier.arrive
            size_t off = 0;
            auto issue = [&](int s) {
                for (int q = 0; q < stageBytes; q += 512)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(sm + s * stageBytes + q + lane * 16)), "l"(base + off + q + lane * 16));
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(&bar[s])) : "memory");
                off += stageBytes;
                if (off + stageBytes > perCta) off = 0;
            };
            for (int s = 0; s < nstage; ++s) issue(s);
            int st = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                while (!mbarTry(s32(&bar[st]), ph)) {}
                volatile double* p = (volatile double*)(sm + st * stageBytes);
                double v = p[lane];
                if (v == 123.456) out[1] = 1;
                __syncwarp();
                if (it + nstage < iters) issue(st);
                if (++st == nstage) { st = 0; ph ^= 1; }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main()
{
    const size_t perCta = 8u << 20;
    const int nCta = 148;
    char* src;
    cudaMalloc(&src, perCta * nCta);
    cudaMemset(src, 0, perCta * nCta);
    unsigned long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("mode grid size ncopy nstage  bytes_in_flight  GB/s_total  GB/s_per_SM  ns_per_stage\n");
    for (int mode = 0; mode < 2; ++mode)
        for (int grid : {1, 148})
            for (int size : {1024, 2048, 4096, 8192, 16384})
                for (int ncopy : {1, 4, 12})
                    for (int nstage : {2, 4, 8}) {
                        const int stageBytes = size * ncopy;
                        if ((size_t)stageBytes * nstage > 190 * 1024) continue;
                        if (mode == 1 && (ncopy != 1)) continue;
                        const int iters = 2000;
                        for (int rep = 0; rep < 2; ++rep) {
                            cudaEventRecord(e0);
                            if (mode == 0) k<0><<<grid, 64, (size_t)stageBytes * nstage>>>(src, perCta, size, ncopy, nstage, iters, out);
                            else k<1><<<grid, 64, (size_t)stageBytes * nstage>>>(src, perCta, size, ncopy, nstage, iters, out);
                            cudaEventRecord(e1);
                            cudaEventSynchronize(e1);
                        }
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        cudaError_t e = cudaGetLastError();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        const double bytes = (double)stageBytes * iters * grid;
                        printf("%d %4d %6d %3d %2d  %8d  %9.1f  %7.2f  %8.1f\n", mode, grid, size, ncopy, nstage, stageBytes * nstage, bytes / ms * 1e-6,
                               bytes / ms * 1e-6 / grid, ms * 1e6 / iters);
                    }
    return 0;
}
