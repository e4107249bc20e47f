// Dev microbenchmark: streaming handoff of 256-byte rows through L2 between CTA pairs, consumer polling with an
// 8-deep prefetch ring that is armed BEFORE the producer starts (the pencil pipeline's start-up case).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SENT 0x7FF4DEADBEEF5A5AULL
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int MODE> __device__ __forceinline__ unsigned long long ldp(const unsigned long long* p)
{
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// pair p: block 2p produces rows of buf[p], block 2p+1 consumes.  rows: n, gapCycles between rows, startDelay ns
template <int MODE>
__global__ void k_stream(unsigned long long* buf, int n, int gap, int startDelayNs, unsigned long long* tProd, unsigned long long* tCons, int rowsPerPair)
{
    const int pair = blockIdx.x >> 1, lane = threadIdx.x;
    unsigned long long* b = buf + (size_t)pair * rowsPerPair * 32;
    if ((blockIdx.x & 1) == 0) {
        unsigned long long t0 = gtime();
        while (gtime() - t0 < (unsigned long long)startDelayNs) {}
        for (int i = 0; i < n; ++i) {
            long long c = clock64();
            while (clock64() - c < gap) {}
            if (lane == 0) tProd[pair * n + i] = gtime();
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(b + i * 32 + lane), "l"((unsigned long long)(i + 1)));
        }
    } else {
        unsigned long long r[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) r[d] = ldp<MODE>(b + d * 32 + lane);
        for (int i0 = 0; i0 < n; i0 += 8) {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                while (r[d] == SENT) r[d] = ldp<MODE>(b + (i0 + d) * 32 + lane);
                if (lane == 0) tCons[pair * n + i0 + d] = gtime();
                r[d] = ldp<MODE>(b + (i0 + d + 8) * 32 + lane);
            }
        }
    }
}
__global__ void k_arm(unsigned long long* buf, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = SENT; }
template <int MODE> void run(int pairs, int gap, int delayNs)
{
    const int n = 160, rpp = 4096;
    unsigned long long *buf, *tp, *tc;
    cudaMalloc(&buf, (size_t)pairs * rpp * 32 * 8); cudaMalloc(&tp, pairs * n * 8); cudaMalloc(&tc, pairs * n * 8);
    unsigned long long* hp = new unsigned long long[pairs * n]; unsigned long long* hc = new unsigned long long[pairs * n];
    for (int rep = 0; rep < 2; ++rep) {
        k_arm<<<256, 256>>>(buf, (size_t)pairs * rpp * 32);
        cudaDeviceSynchronize();
        k_stream<MODE><<<2 * pairs, 32>>>(buf, n, gap, delayNs, tp, tc, rpp);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(hp, tp, pairs * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc, tc, pairs * n * 8, cudaMemcpyDeviceToHost);
    double first = 0, mid = 0, mx = 0;
    for (int p = 0; p < pairs; ++p) {
        for (int i = 0; i < n; ++i) { double d = (double)hc[p * n + i] - (double)hp[p * n + i]; if (d > mx) mx = d; }
        first += (double)hc[p * n] - (double)hp[p * n];
        mid += (double)hc[p * n + 80] - (double)hp[p * n + 80];
    }
    printf("mode %d pairs %3d gap %4d delay %5d ns: handoff lag first row %.0f ns, row 80 %.0f ns, max %.0f ns  (%s)\n", MODE, pairs, gap, delayNs,
           first / pairs, mid / pairs, mx, cudaGetErrorString(cudaGetLastError()));
    cudaFree(buf); cudaFree(tp); cudaFree(tc); delete[] hp; delete[] hc;
}
int main()
{
    for (int pairs : {1, 32, 64}) {
        run<0>(pairs, 400, 3000);
        run<1>(pairs, 400, 3000);
        run<2>(pairs, 400, 3000);
    }
    run<0>(64, 400, 0);
    run<0>(64, 100, 3000);
    return 0;
}
