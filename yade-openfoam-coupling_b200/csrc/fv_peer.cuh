// fv_peer.cuh -- the collectives of the decomposed PCG iteration, done by the iteration's OWN kernels over NVLink peer
// memory (csrc/fv_dist.cu maps the peers' buffers with CUDA IPC at fy_dist_init).
//
// Why not NCCL here: a PCG iteration on a slab of the box is ~0.2 ms of kernels and needs three global sums of ONE
// double and one halo exchange of a few hundred KB.  As separate NCCL launches each costs 20-30 us of launch + protocol
// latency (measured: profiles/r2h_*), four times per iteration -- as much as the sweeps the decomposition shortens.
// Fused, the kernel that produces a partial sum finishes the sum itself:
//   all-reduce   one warp of the last block of the producing kernel stores its rank's partial sums into EVERY rank's
//                mailbox (slot = source rank, lane = destination rank), polls its own mailbox -- local memory -- until
//                all ranks' words have arrived, adds the values in RANK ORDER (every rank gets bit-identical totals) and
//                runs the finishing step (alpha, beta, convergence test) that the single-domain kernel runs in its
//                last block.  The words are self-validating: every 8-byte store carries 4 bytes of the value and the
//                4-byte sequence number of the collective (8-byte stores are single transactions on NVLink), so no
//                fence and no separate flag is needed -- one one-way NVLink latency per all-reduce
//                (tools/peer_bench.cu measures it).  Two mailboxes alternate by sequence parity: a rank can be at most
//                one collective ahead of a peer that still reads.
//   halo         the kernel that writes the search direction pA stores its boundary rows ALSO into the neighbours'
//                ghost rows (every rank keeps the global layout, so the address offset is the same on every rank); its
//                last block then raises a sequence flag in each neighbour's mailbox, and the blocks of the next kernel
//                (Amul) wait for the flags of their rank's neighbours before they read.
// No kernel ever waits for something a peer produces LATER in its stream than what the peer is waiting for itself, so
// the waits cannot form a cycle; every wait has a poll limit that raises the engine's error flag instead of hanging.
#pragma once
#include <cuda_runtime.h>

#define FY_PEER_MAXR 16
#define FY_PEER_SPIN_LIMIT (1u << 27)

struct PeerMail {
    unsigned long long word[2][FY_PEER_MAXR][4];   // [sequence parity][source rank][value 0 lo, hi, value 1 lo, hi]: {data32 | seq32 << 32}
    unsigned long long halo[4];                    // sequence of the last halo received FROM the neighbour zlo zhi ylo yhi
};
// one per rank, in that rank's device memory; the kernels get a pointer to it
struct PeerDev {
    PeerMail* box[FY_PEER_MAXR];        // every rank's mailbox as mapped into this process (own: the local pointer)
    double* pa[4];                      // the neighbours' search-direction vectors: zlo zhi ylo yhi (null: no neighbour)
    int nbr[4];                         // their ranks (-1: none)
    int rank, nranks;
    unsigned long long redSeq, haloSeq; // collectives done so far (the ranks run in lockstep: same counts everywhere)
    unsigned int dirCount, pad;         // blocks of the halo-writing kernel that have finished
    int* error;
};

__device__ __forceinline__ void peerSt(unsigned long long* p, unsigned long long v) { *(volatile unsigned long long*)p = v; }
__device__ __forceinline__ unsigned long long peerLd(const unsigned long long* p) { return *(const volatile unsigned long long*)p; }

// SUM over the ranks of tot[0..NV), in place.  WARP-collective: all 32 lanes of one warp of the rank call it (the values
// of lane 0 count); every lane returns the totals.
template <int NV>
__device__ __forceinline__ void peerAllReduce(PeerDev* pd, double (&tot)[NV])
{
    static_assert(NV <= 2, "mailbox slots hold two values");
    const int lane = threadIdx.x & 31;
    const int me = pd->rank, n = pd->nranks;
    const unsigned long long s = pd->redSeq + 1;
    PeerMail* const mine = pd->box[me];
    PeerMail* const dst = lane < n ? pd->box[lane] : nullptr;
    const int par = (int)(s & 1);
    const unsigned long long tag = (s & 0xffffffffull) << 32;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(__shfl_sync(0xffffffffu, tot[q], 0));
        if (dst) {
            peerSt(&dst->word[par][me][2 * q], (b & 0xffffffffull) | tag);
            peerSt(&dst->word[par][me][2 * q + 1], (b >> 32) | tag);
        }
    }
    double got[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        got[q] = 0.0;
        if (lane < n) {
            unsigned long long w[2];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                unsigned int spin = 0;
                do {
                    w[hf] = peerLd(&mine->word[par][lane][2 * q + hf]);
                    if (++spin > FY_PEER_SPIN_LIMIT) { atomicExch(pd->error, 1); break; }
                } while ((w[hf] >> 32) != (s & 0xffffffffull));
            }
            got[q] = __longlong_as_double((long long)((w[0] & 0xffffffffull) | (w[1] << 32)));
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double a = 0.0;
        for (int r = 0; r < n; ++r) a += __shfl_sync(0xffffffffu, got[q], r);      // rank order: the same sum on every rank
        tot[q] = a;
    }
    if (lane == 0) pd->redSeq = s;
}

// the halo-writing kernel, after its last store: every thread of every block calls it (wrote: this thread stored into a
// neighbour's memory)
__device__ __forceinline__ void peerHaloPublish(PeerDev* pd, bool wrote)
{
    if (wrote) __threadfence_system();                     // the remote stores are performed before the block reports
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    const unsigned int t = atomicAdd(&pd->dirCount, 1u);
    if (t != gridDim.x * gridDim.y * gridDim.z - 1) return;
    __threadfence_system();
    pd->dirCount = 0u;
    const unsigned long long s = pd->haloSeq + 1;
    for (int d = 0; d < 4; ++d)
        if (pd->nbr[d] >= 0) peerSt(&pd->box[pd->nbr[d]]->halo[d ^ 1], s);      // I am the neighbour's neighbour on the other side
    pd->haloSeq = s;
}

// the halo-reading kernel, before its first load: every thread of every block calls it
__device__ __forceinline__ void peerHaloWait(PeerDev* pd)
{
    if (threadIdx.x < 4) {
        const int d = threadIdx.x;
        if (pd->nbr[d] >= 0) {
            const unsigned long long s = pd->haloSeq;
            const unsigned long long* f = &pd->box[pd->rank]->halo[d];
            unsigned int spin = 0;
            while (peerLd(f) < s)
                if (++spin > FY_PEER_SPIN_LIMIT) { atomicExch(pd->error, 1); break; }
        }
        __threadfence();
    }
    __syncthreads();
}
