// comm.cu -- the multi-GPU exchange of a commitment group, fused into ONE kernel over NVLink / NVSwitch peer memory.
//
// Multi-GPU MSM shards the scalars and the key by row: every rank ends its Pippenger pipeline with one
// partial sum per commitment (128-byte XYZZ point) and the commitment is the sum of the ranks' partials.  Point
// addition is not an NCCL reduction operator, so the library path was all-gather (NCCL) + a combine kernel: two more
// launches and NCCL's protocol latency on the critical path of EVERY commitment group, four times per fold step.
// Here the last kernel of the pipeline does the exchange itself:
//
//   1. adds the two halves of the weighted bucket sum (X + Y) -> this rank's partial (4-warp cooperative addition, coop.cuh),
//   2. stores it into the mailbox of every peer (plain 16-byte stores to peer memory mapped through CUDA IPC: NVLink
//      P2P writes), fences system-wide, then raises a per-(rank, commitment) sequence flag in each peer's mailbox,
//   3. spins on the flags in its OWN mailbox (local HBM polls) until every rank's partial has arrived,
//   4. adds the partials by a log2(world)-step butterfly over the block's logical lanes and normalises to affine.
//
// No host round trip, no NCCL launch; every rank ends with the same affine commitments.  Mailboxes are double-buffered
// by the parity of the call sequence number: a rank can be at most one call ahead of a peer (it needs the peer's flag
// of call s to finish call s), so call s + 1 never overwrites a slot a peer may still be reading.  A spin that lasts
// longer than SB_COMM_TIMEOUT_NS raises the status word instead of hanging the GPU (a rank died): the next call fails.
#include <string.h>

#include <new>

#include "common.cuh"
#include "curve.cuh"
#include "coop.cuh"

namespace sb {

constexpr int COMM_MAX_WORLD = 16;
constexpr unsigned long long COMM_TIMEOUT_NS = 4000000000ull;   // 4 s

struct CommView {
    char* peers[COMM_MAX_WORLD];   // every rank's mailbox as seen from this device (peers[rank] = the local one)
    int rank, world;
    uint32_t max_batch;
};

SB_D size_t mailbox_slot_off(const CommView& c, uint32_t parity, uint32_t src, uint32_t b) {
    return (((size_t)parity * c.world + src) * c.max_batch + b) * 128;
}
SB_D size_t mailbox_flag_off(const CommView& c, uint32_t parity, uint32_t src, uint32_t b) {
    const size_t slots = (size_t)2 * c.world * c.max_batch * 128;
    return slots + (((size_t)parity * c.world + src) * c.max_batch + b) * 4;
}
static size_t mailbox_bytes(int world, size_t max_batch) { return (size_t)2 * world * max_batch * (128 + 4) + 256; }

SB_D unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
SB_D void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
SB_D uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SB_D uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// grid = batch blocks of 128 threads = 32 logical lanes of the 4-warp cooperative group law (coop.cuh).
// in: pairs ? xy[b][2] (the two halves of the weighted bucket sum) : partial[b].
template <class F>
__global__ void __launch_bounds__(COOP_THREADS)
k_exchange_combine(CommView c, const unsigned int* __restrict__ seq_counter, const XYZZ<F>* __restrict__ in, int pairs, Affine<F>* __restrict__ out_xy,
                   unsigned int* __restrict__ status) {
    __shared__ CoopBuf sh;
    __shared__ XYZZ<F> mine;
    __shared__ int failed;
    // The call sequence number lives in DEVICE memory (k_seq_bump raises it behind this kernel, in stream order), not in a
    // kernel argument: a captured pipeline (CUDA graph, msm.cu) replays with identical arguments.  Every rank issues the
    // same exchanges in the same order on a communicator, so the counters agree.
    const uint32_t seq = *reinterpret_cast<const volatile unsigned int*>(seq_counter) + 1u;
    const uint32_t b = blockIdx.x, parity = seq & 1u;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) failed = 0;
    // 1. this rank's partial: X + Y (every logical lane computes the same sum; one addition of latency)
    XYZZ<F> acc = in[pairs ? 2 * b : b];
    if (pairs) {
        const XYZZ<F> y = in[2 * b + 1];
        coop4_add(acc, y, sh);
    }
    if (tid == 0) mine = acc;
    __syncthreads();
    // 2. the partial goes to every rank's mailbox: 8 x 16-byte stores per peer
    if (tid < c.world * 8) {
        const int peer = tid >> 3, part = tid & 7;
        const uint4 v = reinterpret_cast<const uint4*>(&mine)[part];
        uint4* dst = reinterpret_cast<uint4*>(c.peers[peer] + mailbox_slot_off(c, parity, (uint32_t)c.rank, b)) + part;
        *dst = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < c.world) st_release_sys(reinterpret_cast<uint32_t*>(c.peers[tid] + mailbox_flag_off(c, parity, (uint32_t)c.rank, b)), seq);
    // 3. wait for every rank's partial in the local mailbox
    if (tid < c.world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(c.peers[c.rank] + mailbox_flag_off(c, parity, (uint32_t)tid, b));
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) != seq) {
            if (global_ns() - t0 > COMM_TIMEOUT_NS) {
                failed = 1;
                atomicExch(status, 1u);
                break;
            }
        }
    }
    __syncthreads();
    if (failed) return;
    // 4. logical lane r takes rank r's partial; a log2(world)-step butterfly adds them (same order on every rank)
    XYZZ<F> q = XYZZ<F>::identity();
    if (lane < c.world) {
        const uint4* src = reinterpret_cast<const uint4*>(c.peers[c.rank] + mailbox_slot_off(c, parity, (uint32_t)lane, b));
#pragma unroll
        for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(&q)[k] = ld_volatile_v4(src + k);
    }
    int span = 1;
    while (span < c.world) span <<= 1;
#pragma unroll 1
    for (int d = span >> 1; d >= 1; d >>= 1) {
        XYZZ<F> t = coop_shfl_xor(q, d);
        coop4_add(q, t, sh);
    }
    if (tid == 0) {
        Affine<F> a = xyzz_to_affine<false>(q);
        uint4* d = reinterpret_cast<uint4*>(out_xy + b);
        const uint4* s = reinterpret_cast<const uint4*>(&a);
#pragma unroll
        for (int k = 0; k < 4; k++) d[k] = s[k];
    }
}

__global__ void k_seq_bump(unsigned int* seq_counter) { *seq_counter += 1u; }

}  // namespace sb

struct sb_comm {
    int rank, world;
    size_t max_batch;
    char* local;
    char* peers[sb::COMM_MAX_WORLD];
    bool opened[sb::COMM_MAX_WORLD];
    unsigned int* d_status;   // [0] status word, [1] call sequence counter
    bool connected;
};

using namespace sb;

namespace sb {
void msm_graphs_drop(const void* ck, cudaStream_t st, const void* comm, bool all);   // msm.cu
// used by msm.cu's sharded commit: run the exchange on `in` (pairs: xy halves in the MSM workspace)
int comm_exchange_enqueue(sb_comm* c, int curve, const void* d_in, int pairs, size_t batch, void* d_out_xy, cudaStream_t st) {
    if (!c || !c->connected) {
        set_error("sb_comm: communicator not connected");
        return SB_ERR_ARG;
    }
    if (batch > c->max_batch) {
        set_error("sb_comm: batch %zu exceeds the communicator's max_batch %zu", batch, c->max_batch);
        return SB_ERR_ARG;
    }
    if (!batch) return SB_OK;
    CommView v;
    for (int r = 0; r < COMM_MAX_WORLD; r++) v.peers[r] = r < c->world ? c->peers[r] : nullptr;
    v.rank = c->rank;
    v.world = c->world;
    v.max_batch = (uint32_t)c->max_batch;
    unsigned int* seq_counter = c->d_status + 1;
    if (curve == CURVE_BN256)
        k_exchange_combine<Fq><<<(unsigned)batch, COOP_THREADS, 0, st>>>(v, seq_counter, (const XYZZ<Fq>*)d_in, pairs, (Affine<Fq>*)d_out_xy, c->d_status);
    else if (curve == CURVE_GRUMPKIN)
        k_exchange_combine<Fr><<<(unsigned)batch, COOP_THREADS, 0, st>>>(v, seq_counter, (const XYZZ<Fr>*)d_in, pairs, (Affine<Fr>*)d_out_xy, c->d_status);
    else {
        set_error("sb_comm: unknown curve %d", curve);
        return SB_ERR_ARG;
    }
    SB_KERNEL_CHECK();
    k_seq_bump<<<1, 1, 0, st>>>(seq_counter);
    SB_KERNEL_CHECK();
    return SB_OK;
}
}  // namespace sb

extern "C" {

int sb_comm_create(int rank, int world, size_t max_batch, sb_comm_t* out, unsigned char ipc_handle_out[64]) {
    if (!out || !ipc_handle_out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || !max_batch || max_batch > 4096) {
        set_error("sb_comm_create: bad argument (world <= %d, max_batch <= 4096)", COMM_MAX_WORLD);
        return SB_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    SB_TRY(ensure_runtime());
    sb_comm* c = new (std::nothrow) sb_comm();
    if (!c) return SB_ERR_OOM;
    memset(c, 0, sizeof(*c));
    c->rank = rank;
    c->world = world;
    c->max_batch = max_batch;
    const size_t bytes = mailbox_bytes(world, max_batch);
    cudaError_t e = cudaMalloc((void**)&c->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_status, 8);
    if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, 8);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("sb_comm_create: %s", cudaGetErrorString(e));
        if (c->local) cudaFree(c->local);
        if (c->d_status) cudaFree(c->d_status);
        delete c;
        return SB_ERR_CUDA;
    }
    memcpy(ipc_handle_out, &h, 64);
    c->peers[rank] = c->local;
    *out = c;
    return SB_OK;
}

int sb_comm_connect(sb_comm_t c, const unsigned char* all_handles) {
    if (!c || !all_handles) {
        set_error("sb_comm_connect: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank || c->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("sb_comm_connect: cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink is required)", r, cudaGetErrorString(e));
            return SB_ERR_NCCL;
        }
        c->peers[r] = (char*)p;
        c->opened[r] = true;
    }
    c->connected = true;
    return SB_OK;
}

void sb_comm_destroy(sb_comm_t c) {
    if (!c) return;
    cudaDeviceSynchronize();
    {
        RtLock lk(runtime().mu);
        msm_graphs_drop(nullptr, nullptr, c, false);
    }
    for (int r = 0; r < c->world; r++)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers[r]);
    if (c->local) cudaFree(c->local);
    if (c->d_status) cudaFree(c->d_status);
    delete c;
}

/* 0 while every exchange completed; 1 after a rank failed to show up within the timeout (synchronises the stream first). */
int sb_comm_status(sb_comm_t c, void* stream) {
    if (!c) return -1;
    unsigned int h = 0;
    if (stream) cudaStreamSynchronize((cudaStream_t)stream);
    if (cudaMemcpy(&h, c->d_status, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)h;
}

int sb_comm_allsum_points_device(sb_comm_t c, int curve, const void* d_partials_xyzz, size_t batch, void* d_out_xy, void* stream) {
    if (!d_partials_xyzz || !d_out_xy) {
        set_error("sb_comm_allsum_points_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return comm_exchange_enqueue(c, curve, d_partials_xyzz, 0, batch, d_out_xy, stream ? (cudaStream_t)stream : rt.stream);
}

}  // extern "C"
