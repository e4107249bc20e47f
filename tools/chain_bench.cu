// Dev microbenchmark (not part of the product): cycles per row-step of ONE warp running the DIC forward-substitution
// chain (w = c0 - c3*vz - c2*vy - c1*vx; vy by shuffle) with the pieces of the real loop added one at a time, and
// with different orderings / hand-over protocols of the z channel.  The channel ring feeds itself (the slot read at
// step t was written 16 steps earlier), so a number is the throughput limit of a consumer stage whose producer is ahead.
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o chain_bench chain_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double ldsV(uint32_t p) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(p)); return v; }
__device__ __forceinline__ void lds2V(uint32_t p, double& a, double& b) { asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(p)); }
__device__ __forceinline__ void lds2Vq(uint32_t p, double& a, long long& b) { asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=d"(a), "=l"(b) : "r"(p)); }
__device__ __forceinline__ void stsV(uint32_t p, double v) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(p), "d"(v)); }
__device__ __forceinline__ void sts2Vq(uint32_t p, double a, long long b) { asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(p), "d"(a), "l"(b)); }
__device__ __forceinline__ void stg(double* p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ void mbarInit(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbarArrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool mbarTry(uint32_t b, uint32_t ph)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
    return ok;
}
__device__ __forceinline__ bool sent(double v) { return __double2hiint(v) == 0x7ff4dead; }
__device__ __forceinline__ double sentV() { return __longlong_as_double(0x7FF4DEADBEEF5A5ALL); }

// PROTO 0: no z channel.  1: sentinel slots, order [check, re-arm, prefetch next | inputs(t) | math | stores]
// 2: sentinel slots, software pipelined: [inputs(t+1), prefetch z(t+1), re-arm(t) | math(t) | stores(t) | check(t+1)]
// 3: {value, seq} 16-byte slots, no re-arm, software pipelined like 2
template <int PROTO, int Z, bool YRING>
__global__ void k(double* out, unsigned long long* cyc, int steps)
{
    __shared__ __align__(128) double ring[8 * 5 * 32 * Z];     // 8 rows x 5 doubles x 32 lanes x Z planes
    __shared__ __align__(16) double chan[32 * 32 * 2];
    __shared__ double yring[64 * 2];
    __shared__ unsigned long long bar[2];
    if (threadIdx.x == 0) { mbarInit(s32(&bar[0]), 1); mbarInit(s32(&bar[1]), 1); }
    const int lane = threadIdx.x & 31;
    for (int x = threadIdx.x; x < 8 * 5 * 32 * Z; x += blockDim.x) ring[x] = 1e-3 * (x % 7);
    for (int x = threadIdx.x; x < 32 * 32 * 2; x += blockDim.x) chan[x] = PROTO == 3 ? ((x & 1) ? __longlong_as_double((long long)(x / 64)) : 0.5) : 0.5;
    for (int x = threadIdx.x; x < 128; x += blockDim.x) yring[x] = 0.25;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    double prev[Z];
#pragma unroll
    for (int z = 0; z < Z; ++z) prev[z] = 1.0 + z;
    const uint32_t rs = s32(ring) + lane * 16, rs8 = s32(ring) + lane * 8;
    const uint32_t cs = s32(chan) + lane * (PROTO == 3 ? 16 : 8), CW = PROTO == 3 ? 512 : 256;
    const uint32_t ys = s32(yring);
    double* op = out + lane + (size_t)blockIdx.x * 1024 * 32 * Z;
    auto loadIn = [&](int r, double (&c)[Z][5], double (&yv)[Z]) {
#pragma unroll
        for (int z = 0; z < Z; ++z) {
            lds2V(rs + (z * 5 * 8 + r * 2) * 256, c[z][0], c[z][1]);
            lds2V(rs + (z * 5 * 8 + 16 + r * 2) * 256, c[z][2], c[z][3]);
            c[z][4] = ldsV(rs8 + (z * 5 * 8 + 32 + r) * 256);
            if (YRING) yv[z] = ldsV(ys + z * 512 + r * 8);
        }
    };
    double cA[Z][5], cB[Z][5], yA[Z], yB[Z];
#pragma unroll
    for (int z = 0; z < Z; ++z) { yA[z] = yB[z] = 0; }
    double vzN = 0.25;
    long long sqN = 0;
    if (PROTO == 1 || PROTO == 2 || PROTO >= 4) vzN = ldsV(cs);
    if (PROTO == 3) lds2Vq(cs, vzN, sqN);
    if (PROTO >= 2 || PROTO == 0) loadIn(0, cA, yA);
    unsigned long long t0 = clock64();
    for (int t0s = 0; t0s < steps; t0s += 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int t = t0s + r;
            double (&c)[Z][5] = (r & 1) ? cB : cA;
            double (&cn)[Z][5] = (r & 1) ? cA : cB;
            double (&yv)[Z] = (r & 1) ? yB : yA;
            double (&yn)[Z] = (r & 1) ? yA : yB;
            double vz0 = 0.25;
            if (PROTO == 1) {
                double v = vzN;
                if (sent(v)) { do { v = ldsV(cs + (t & 31) * CW); } while (sent(v)); }
                stsV(cs + (t & 31) * CW, sentV());
                vzN = ldsV(cs + ((t + 1) & 31) * CW);
                vz0 = v;
                loadIn(r, c, yv);
            } else if (PROTO == 2 || PROTO == 4 || PROTO == 5 || PROTO == 6 || PROTO == 7 || PROTO == 8) {
                vz0 = vzN;                                   // checked at the end of the previous step
                loadIn((r + 1) & 7, cn, yn);
                vzN = ldsV(cs + ((t + 1) & 31) * CW);
                stsV(cs + (t & 31) * CW, sentV());
            } else if (PROTO == 3) {
                vz0 = vzN;
                loadIn((r + 1) & 7, cn, yn);
                lds2Vq(cs + ((t + 1) & 31) * CW, vzN, sqN);
            } else {
                if ((PROTO == 9 || PROTO == 10) && (r & 3) == 0) {
                    const int b = (t >> 2) & 1;
                    const uint32_t ph = (uint32_t)((t >> 3) & 1);
                    if (lane == 0) mbarArrive(s32(&bar[b]));
                    while (!mbarTry(s32(&bar[b]), ph)) {}
                }
                if (PROTO == 10 && (r & 3) == 0) {
                    double chk = ldsV(cs + ((t + 7) & 31) * CW);
                    if (!sent(chk) && chk == 123.0) { do { chk = ldsV(cs + ((t + 7) & 31) * CW); } while (chk == 123.0); }
                }
                loadIn((r + 1) & 7, cn, yn);
            }
            if (PROTO == 8) {
                double vyS[Z], p0[Z], p3[Z], p1[Z];
#pragma unroll
                for (int z = Z - 1; z >= 0; --z) {
                    vyS[z] = __shfl_up_sync(0xffffffffu, prev[z], 1);
                    p0[z] = c[z][0] * c[z][4];
                    p3[z] = c[z][3] * (z > 0 ? prev[z - 1] : vz0);
                    p1[z] = c[z][1] * prev[z];
                    p0[z] -= p3[z];
                }
                if (sent(vzN)) { do { vzN = ldsV(cs + ((t + 1) & 31) * CW); } while (sent(vzN)); }
#pragma unroll
                for (int z = Z - 1; z >= 0; --z) {
                    double vy = vyS[z];
                    if (YRING) vy = lane == 0 ? yv[z] : vy;
                    double w = p0[z];
                    w -= c[z][2] * vy;
                    w -= p1[z];
                    stg(op + (size_t)((t & 1023) * 32) + z * 32 * 1024, w);
                    prev[z] = w;
                }
                stsV(cs + ((t + 16) & 31) * CW, prev[Z - 1] * 0.0 + 0.5);
                continue;
            }
#pragma unroll
            for (int z = Z - 1; z >= 0; --z) {
                const double vx = prev[z];
                double vy = __shfl_up_sync(0xffffffffu, prev[z], 1);
                if (YRING) vy = lane == 0 ? yv[z] : vy;
                const double vz = z > 0 ? prev[z - 1] : vz0;
                double w = c[z][0] * c[z][4];
                w -= c[z][3] * vz;
                w -= c[z][2] * vy;
                w -= c[z][1] * vx;
                stg(op + (size_t)((t & 1023) * 32) + z * 32 * 1024, w);
                prev[z] = w;
            }
            // hand-over of the last plane (to a slot 16 steps ahead of the reader) -- value forced to 0.5 so the run stays finite
            const double hv = prev[Z - 1] * 0.0 + 0.5;
            if (PROTO == 1 || PROTO == 2 || PROTO == 4 || PROTO == 5 || PROTO == 10) stsV(cs + ((t + 16) & 31) * CW, hv);
            if (PROTO == 6) stsV(cs + ((t + 16) & 31) * CW, 0.5);
            if (PROTO == 7) {
                stsV(cs + ((t + 16) & 31) * CW, hv);
                if (__builtin_expect(sent(vzN), 0)) { do { vzN = ldsV(cs + ((t + 1) & 31) * CW); } while (sent(vzN)); }
            }
            if (PROTO == 5) {
                if (__any_sync(0xffffffffu, sent(vzN))) { do { vzN = ldsV(cs + ((t + 1) & 31) * CW); } while (__any_sync(0xffffffffu, sent(vzN))); }
            }
            if (PROTO == 6) {
                if (sent(vzN)) { do { vzN = ldsV(cs + ((t + 1) & 31) * CW); } while (sent(vzN)); }
            }
            if (PROTO == 3) sts2Vq(cs + ((t + 16) & 31) * CW, hv, (long long)(t + 16));
            if (PROTO == 2) {
                if (sent(vzN)) { do { vzN = ldsV(cs + ((t + 1) & 31) * CW); } while (sent(vzN)); }
            }
            if (PROTO == 3) {
                if (sqN != (long long)(t + 1)) { do { lds2Vq(cs + ((t + 1) & 31) * CW, vzN, sqN); } while (sqN != (long long)(t + 1)); }
            }
        }
    }
    unsigned long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int z = 0; z < Z; ++z) s += prev[z];
    if (lane == 0) { cyc[blockIdx.x] = t1 - t0; out[blockIdx.x] = s; }
}

template <int PROTO, int Z, bool YRING>
void run(double* out, unsigned long long* cyc)
{
    const int steps = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k<PROTO, Z, YRING><<<1, 64>>>(out, cyc, steps);
        cudaDeviceSynchronize();
    }
    cudaError_t e = cudaGetLastError();
    printf("proto %d Z %d yring %d: %7.1f cycles/step  %s\n", PROTO, Z, (int)YRING, (double)cyc[0] / steps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    double* out;
    unsigned long long* cyc;
    cudaMalloc(&out, (size_t)4 * 1024 * 32 * 8 * 4);
    cudaMallocManaged(&cyc, 1024 * 8);
    run<0, 1, false>(out, cyc); run<1, 1, false>(out, cyc); run<2, 1, false>(out, cyc); run<3, 1, false>(out, cyc);
    run<0, 1, true>(out, cyc);  run<1, 1, true>(out, cyc);  run<2, 1, true>(out, cyc);  run<3, 1, true>(out, cyc);
    run<0, 2, false>(out, cyc); run<1, 2, false>(out, cyc); run<2, 2, false>(out, cyc); run<3, 2, false>(out, cyc);
    run<0, 2, true>(out, cyc);  run<1, 2, true>(out, cyc);  run<2, 2, true>(out, cyc);  run<3, 2, true>(out, cyc);
    run<4, 1, true>(out, cyc);  run<5, 1, true>(out, cyc);  run<6, 1, true>(out, cyc);
    run<9, 1, true>(out, cyc); run<10, 1, true>(out, cyc); run<9, 2, true>(out, cyc); run<10, 2, true>(out, cyc);
    run<7, 1, true>(out, cyc); run<7, 2, true>(out, cyc); run<8, 1, true>(out, cyc); run<8, 2, true>(out, cyc);
    run<4, 2, true>(out, cyc);  run<5, 2, true>(out, cyc);  run<6, 2, true>(out, cyc);

    return 0;
}
