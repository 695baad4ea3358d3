// SyncBN statistics exchange over peer memory (NVLink / NVSwitch P2P), the one exchange step on the training path
// that sits ON the critical chain: every BatchNorm needs the {sum, sumsq} of the whole global minibatch before the
// next layer can normalise (net/batchnormlayer.py:154-159), 61 times forward and 61 times backward per step, 2*C
// doubles each.  An NCCL all-reduce costs 10-20 us of launch / proxy latency per call at this size; here the
// exchange is ONE single-CTA kernel in stream order (capturable into the step's CUDA graph):
//
//   every rank r:  writes its 2*C local sums into slot r of EVERY rank's exchange buffer (st.global over NVLink),
//                  fences system-wide, then writes the sync point's sequence number into flag r of every rank;
//                  waits until all `world` flags of its OWN buffer carry that sequence number (ld.acquire.sys);
//                  adds the world slots in rank order (the same order everywhere: replicas stay bit-identical) and
//                  overwrites its local sums with the global ones.
//
// Sequence numbers grow monotonically (a device-side counter advanced by the kernel itself, so graph replays need no
// new arguments) and every sync point has its own slot region: nothing is ever reset and a rank that runs ahead cannot
// overwrite data another rank still has to read (it cannot pass a sync point before everyone has arrived at it).
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int XT = 256;

// layout of a rank's exchange buffer for one sync point: [world][n] doubles (slots) then [world] u64 (flags)
__global__ void __launch_bounds__(XT)
k_stats_exchange(double *__restrict__ stats, int n, unsigned char *const *__restrict__ peers, int64_t region_off, int rank,
                 int world, unsigned long long *__restrict__ seq_counter, unsigned int *__restrict__ err) {
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    pdl_trigger();          // programmatic dependent launch: the launch latency hides under the producer's tail ...
    pdl_wait();             // ... and nothing is read before the producer of `stats` has completed
    if (tid == 0) s_seq = *seq_counter + 1ull;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const size_t flags_off = (size_t)region_off + (size_t)world * n * sizeof(double);
    // 1. my sums -> slot `rank` of every rank (my own buffer included)
    for (int i = tid; i < n * world; i += XT) {
        const int p = i / n, c = i - p * n;
        double *dst = reinterpret_cast<double *>(peers[p] + region_off) + (size_t)rank * n + c;
        asm volatile("st.global.relaxed.sys.f64 [%0], %1;" ::"l"(dst), "d"(stats[c]) : "memory");
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag `rank` of every rank := seq
    if (tid < world) {
        unsigned long long *f = reinterpret_cast<unsigned long long *>(peers[tid] + flags_off) + rank;
        asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    }
    // 3. wait for everybody's flag in MY buffer
    if (tid < world) {
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(peers[rank] + flags_off) + tid;
        unsigned long long v = 0;
        unsigned int polls = 0;
        do {
            asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= seq) break;
            __nanosleep(100);
        } while (++polls < (1u << 25));              // seconds: only if a rank died
        if (v < seq) atomicExch(err, 0xDEADu);
    }
    __syncthreads();
    // 4. global sums, same order on every rank
    const double *mine = reinterpret_cast<const double *>(peers[rank] + region_off);
    for (int c = tid; c < n; c += XT) {
        double acc = 0.0;
        for (int p = 0; p < world; ++p) {
            double v;
            asm volatile("ld.global.relaxed.sys.f64 %0, [%1];" : "=d"(v) : "l"(mine + (size_t)p * n + c) : "memory");
            acc += v;
        }
        stats[c] = acc;
    }
    if (tid == 0) *seq_counter = seq;
}

}  // namespace

extern "C" int dpp_peer_alloc(int64_t bytes, void **ptr_out, unsigned char *ipc_handle_out64) {
    DPP_CHECK_ARG(bytes > 0 && ptr_out && ipc_handle_out64);
    void *p = nullptr;
    DPP_CUDA(cudaMalloc(&p, (size_t)bytes));
    DPP_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    DPP_CUDA(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(h) == 64, "ipc handle size");
    memcpy(ipc_handle_out64, &h, 64);
    *ptr_out = p;
    return DPP_OK;
}

extern "C" int dpp_peer_open(const unsigned char *ipc_handle64, void **ptr_out) {
    DPP_CHECK_ARG(ipc_handle64 && ptr_out);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    DPP_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return DPP_OK;
}

extern "C" int dpp_peer_close(void *ptr) {
    if (ptr) DPP_CUDA(cudaIpcCloseMemHandle(ptr));
    return DPP_OK;
}

extern "C" int dpp_peer_free(void *ptr) {
    if (ptr) DPP_CUDA(cudaFree(ptr));
    return DPP_OK;
}

extern "C" int dpp_stats_exchange(double *stats, int n, void *const *peers_dev, int64_t region_off, int rank, int world,
                                  unsigned long long *seq_counter, unsigned int *err_flag, void *stream) {
    DPP_CHECK_ARG(stats && n > 0 && peers_dev && region_off >= 0 && rank >= 0 && rank < world && world <= XT && seq_counter &&
                  err_flag);
    DPP_CUDA(launch_pdl(1, k_stats_exchange, dim3(1), dim3(XT), 0, S(stream), stats, n,
                        reinterpret_cast<unsigned char *const *>(peers_dev), region_off, rank, world, seq_counter, err_flag));
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
