// Dev microbenchmarks (B200): latencies that bound the pencil pipeline.  nvcc -arch=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_dp_chain(double* out, double a, double b, int n, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        x = x * a;      // DMUL
        x = x - b;      // DADD
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_chain(double* out, int n, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = __shfl_up_sync(0xffffffffu, x, 1);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fp32_chain(float* out, float a, float b, int n, long long* cyc)
{
    float x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x = x * a; x = x - b; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ping-pong through shared memory between warp 0 and warp 1 of one CTA
__global__ void k_smem_pingpong(int n, long long* cyc)
{
    __shared__ volatile int flag[2];
    if (threadIdx.x == 0) { flag[0] = 0; flag[1] = 0; }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    if (lane == 0) {
        for (int i = 1; i <= n; ++i) {
            if (w == 0) { flag[0] = i; while (flag[1] != i) {} }
            else { while (flag[0] != i) {} flag[1] = i; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// ping-pong through global memory (L2) between CTA 0 and CTA 1
__global__ void k_gmem_pingpong(volatile int* flag, int n, long long* cyc)
{
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int i = 1; i <= n; ++i) {
            if (blockIdx.x == 0) { flag[0] = i; while (flag[32] != i) {} }
            else { while (flag[0] != i) {} flag[32] = i; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
// one-way streaming handoff through L2: producer writes a row of 32 doubles every `gap` cycles, consumer polls them
__global__ void k_gmem_stream(volatile double* buf, int n, int gap, long long* cyc)
{
    const int lane = threadIdx.x;
    if (blockIdx.x == 0) {
        for (int i = 0; i < n; ++i) {
            long long t = clock64();
            while (clock64() - t < gap) {}
            buf[i * 32 + lane] = 1.0 + i;
        }
    } else {
        long long t0 = clock64();
        long long polls = 0;
        for (int i = 0; i < n; ++i) {
            double v = buf[i * 32 + lane];
            while (v == 0.0) { v = buf[i * 32 + lane]; ++polls; }
        }
        long long t1 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = polls; }
    }
}
int main()
{
    double* d; float* f; long long* c; int* flag; double* buf;
    cudaMalloc(&d, 1024 * 8); cudaMalloc(&f, 1024 * 4); cudaMalloc(&c, 64); cudaMalloc(&flag, 4096); cudaMalloc(&buf, 1 << 22);
    cudaMemset(d, 0, 8192); cudaMemset(f, 0, 4096); cudaMemset(flag, 0, 4096);
    long long h[2];
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_dp_chain<<<1, 32>>>(d, 1.0000001, 1e-9, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("fp64 dependent DMUL+DADD pair: %.1f cycles\n", (double)h[0] / n);
        k_fp32_chain<<<1, 32>>>(f, 1.0000001f, 1e-9f, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("fp32 dependent FMUL+FADD pair: %.1f cycles\n", (double)h[0] / n);
        k_shfl_chain<<<1, 32>>>(d, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("dependent 64-bit shfl: %.1f cycles\n", (double)h[0] / n);
        k_smem_pingpong<<<1, 64>>>(n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("smem ping-pong round trip (2 handoffs): %.1f cycles\n", (double)h[0] / n);
        cudaMemset(flag, 0, 4096);
        k_gmem_pingpong<<<2, 32>>>(flag, n, c); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        printf("gmem ping-pong round trip (2 handoffs): %.1f cycles\n", (double)h[0] / n);
        for (int gap : {100, 400, 1000}) {
            cudaMemset(buf, 0, 1 << 22);
            k_gmem_stream<<<2, 32>>>(buf, n, gap, c); cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost);
            printf("gmem stream gap %d: consumer %.1f cycles/row, %.2f polls/row\n", gap, (double)h[0] / n, (double)h[1] / n);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
