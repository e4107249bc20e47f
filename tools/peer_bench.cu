// tools/peer_bench.cu -- development microbenchmark: latency of a 1-double all-reduce between GPUs done by kernels over
// peer memory.  One process, N devices with peer access; each device runs ONE warp doing K all-reduces back to back, so
// time / K is the pure protocol latency.  Variants:
//   0  values, fence.sys, sequence flag (two-phase; fv_peer.cuh's first protocol)
//   1  LL: each 8-byte store carries 4 bytes of data and a 4-byte sequence number -- no fence
//   2  only fence.sys in a loop (no remote traffic)      3  remote store + fence.sys in a loop
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/peer_bench peer_bench.cu ; run: bin/peer_bench [ndev] [K]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define MAXR 8
struct Mail {
    double v[2][MAXR];
    unsigned long long seq[2][MAXR];
    unsigned long long ll[2][MAXR][2];
};
struct Tab { Mail* box[MAXR]; int rank, n; };

__global__ void k_two_phase(Tab t, int K, double* out)
{
    if (threadIdx.x) return;
    double x = t.rank + 1.0;
    Mail* mine = t.box[t.rank];
    for (int it = 1; it <= K; ++it) {
        const int par = it & 1;
        for (int r = 0; r < t.n; ++r) *(volatile double*)&t.box[r]->v[par][t.rank] = x;
        __threadfence_system();
        for (int r = 0; r < t.n; ++r) *(volatile unsigned long long*)&t.box[r]->seq[par][t.rank] = (unsigned long long)it;
        for (int r = 0; r < t.n; ++r) while (*(volatile unsigned long long*)&mine->seq[par][r] != (unsigned long long)it) {}
        __threadfence_system();
        double a = 0;
        for (int r = 0; r < t.n; ++r) a += *(volatile double*)&mine->v[par][r];
        x = a * 0.5;
    }
    *out = x;
}
__global__ void k_ll(Tab t, int K, double* out)
{
    const int lane = threadIdx.x;
    double x = t.rank + 1.0;
    Mail* mine = t.box[t.rank];
    Mail* peer = lane < t.n ? t.box[lane] : nullptr;
    for (int it = 1; it <= K; ++it) {
        const int par = it & 1;
        const unsigned long long b = (unsigned long long)__double_as_longlong(x);
        const unsigned long long tag = (unsigned long long)(unsigned)it << 32;
        if (peer) {
            *(volatile unsigned long long*)&peer->ll[par][t.rank][0] = (b & 0xffffffffull) | tag;
            *(volatile unsigned long long*)&peer->ll[par][t.rank][1] = (b >> 32) | tag;
        }
        double mineV = 0.0;
        if (lane < t.n) {
            unsigned long long w0, w1;
            do { w0 = *(volatile unsigned long long*)&mine->ll[par][lane][0]; } while ((w0 >> 32) != (unsigned)it);
            do { w1 = *(volatile unsigned long long*)&mine->ll[par][lane][1]; } while ((w1 >> 32) != (unsigned)it);
            mineV = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        double a = 0;
        for (int r = 0; r < t.n; ++r) a += __shfl_sync(0xffffffffu, mineV, r);
        x = a * 0.5;
    }
    if (lane == 0) *out = x;
}
__global__ void k_fence(Tab t, int K, double* out, int remote)
{
    if (threadIdx.x) return;
    Mail* other = t.box[(t.rank + 1) % t.n];
    for (int it = 1; it <= K; ++it) {
        if (remote) *(volatile double*)&other->v[0][t.rank] = (double)it;
        __threadfence_system();
    }
    *out = 1.0;
}

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 2, K = argc > 2 ? atoi(argv[2]) : 2000;
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (n > nd) n = nd;
    if (n > MAXR) n = MAXR;
    std::vector<Mail*> box(n);
    std::vector<double*> out(n);
    std::vector<cudaStream_t> st(n);
    std::vector<cudaEvent_t> e0(n), e1(n);
    for (int d = 0; d < n; ++d) {
        CK(cudaSetDevice(d));
        for (int p = 0; p < n; ++p) if (p != d) { int can = 0; cudaDeviceCanAccessPeer(&can, d, p); if (!can) { printf("no peer access %d->%d\n", d, p); return 1; } cudaDeviceEnablePeerAccess(p, 0); }
        cudaGetLastError();
        CK(cudaMalloc(&box[d], sizeof(Mail)));
        CK(cudaMalloc(&out[d], 8));
        CK(cudaStreamCreate(&st[d]));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
    }
    const char* names[] = {"two-phase (store, fence.sys, flag)", "LL (data+tag per 8-byte store, no fence)", "fence.sys only", "remote store + fence.sys"};
    for (int variant = 0; variant < 4; ++variant) {
        for (int rep = 0; rep < 3; ++rep) {
            for (int d = 0; d < n; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(box[d], 0, sizeof(Mail))); CK(cudaDeviceSynchronize()); }
            for (int d = 0; d < n; ++d) {
                CK(cudaSetDevice(d));
                Tab t;
                for (int p = 0; p < n; ++p) t.box[p] = box[p];
                t.rank = d; t.n = n;
                CK(cudaEventRecord(e0[d], st[d]));
                if (variant == 0) k_two_phase<<<1, 32, 0, st[d]>>>(t, K, out[d]);
                else if (variant == 1) k_ll<<<1, 32, 0, st[d]>>>(t, K, out[d]);
                else k_fence<<<1, 32, 0, st[d]>>>(t, K, out[d], variant == 3);
                CK(cudaEventRecord(e1[d], st[d]));
            }
            float worst = 0;
            double res = 0;
            for (int d = 0; d < n; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaDeviceSynchronize());
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
                if (ms > worst) worst = ms;
                CK(cudaMemcpy(&res, out[d], 8, cudaMemcpyDeviceToHost));
            }
            if (rep == 2) printf("%d GPUs  %-44s %8.3f us per op   (result %.6g)\n", n, names[variant], worst * 1000.0 / K, res);
        }
    }
    return 0;
}
