// runtime.cu -- device selection, error strings, scratch memory, and the on-device arithmetic self test.
#include <string.h>

#include <atomic>
#include <map>
#include <memory>
#include <vector>

#include "common.cuh"
#include "curve.cuh"
#include "coop.cuh"

namespace sb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }   // a replayed graph of n kernels (msm.cu)
uint64_t launch_count_now() { return g_launches.load(std::memory_order_relaxed); }
void msm_graphs_drop(const void* ck, cudaStream_t st, const void* comm, bool all);        // msm.cu

struct ProfileRec { cudaEvent_t e0, e1; uint64_t units; int tag; };
static std::atomic<bool> g_profile{false};
static std::mutex g_prof_mu;   // the records are touched from any thread that enqueues (ProfScope) and by sb_profile_collect
static std::vector<ProfileRec> g_prof;
static std::vector<ProfileRec> g_prof_pool;
bool profile_enabled() { return g_profile.load(std::memory_order_relaxed); }
int profile_begin(cudaStream_t st, int tag, uint64_t units) {
    if (!profile_enabled()) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfileRec rec;
    if (!g_prof_pool.empty()) {
        rec = g_prof_pool.back();
        g_prof_pool.pop_back();
    } else {
        cudaEventCreate(&rec.e0);
        cudaEventCreate(&rec.e1);
    }
    rec.units = units;
    rec.tag = tag;
    cudaEventRecord(rec.e0, st);
    g_prof.push_back(rec);
    return (int)g_prof.size() - 1;
}
void profile_end(cudaStream_t st, int handle) {
    if (handle < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (handle >= (int)g_prof.size()) return;
    cudaEventRecord(g_prof[handle].e1, st);
}

Runtime& runtime() {
    static Runtime rt;
    return rt;
}

static int init_locked(Runtime& rt, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s): libsirius_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        return SB_ERR_CUDA;
    }
    if (device >= 0) {
        SB_CUDA_TRY(cudaSetDevice(device));
    }
    int cur = 0;
    SB_CUDA_TRY(cudaGetDevice(&cur));
    if (rt.ready.load() && rt.device == cur) return SB_OK;
    cudaDeviceProp prop;
    SB_CUDA_TRY(cudaGetDeviceProperties(&prop, cur));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", cur, prop.major, prop.minor);
        return SB_ERR_CUDA;
    }
    if (rt.stream) {   // re-initialisation on another device: the old device's scratch goes with its stream
        cudaStreamSynchronize(rt.stream);
        ws_release_all();
        cudaStreamDestroy(rt.stream);
    }
    SB_CUDA_TRY(cudaStreamCreateWithFlags(&rt.stream, cudaStreamNonBlocking));
    rt.device = cur;
    rt.sm_count = prop.multiProcessorCount;
    for (size_t i = 1; i < rt.devs.size(); i++)
        if (rt.devs[i].stream) {
            cudaSetDevice(rt.devs[i].device);
            cudaStreamDestroy(rt.devs[i].stream);
        }
    cudaSetDevice(cur);
    rt.devs.assign(1, Runtime::Dev{cur, rt.stream, rt.sm_count});
    rt.ready.store(true, std::memory_order_release);
    return SB_OK;
}

int ensure_runtime() {
    Runtime& rt = runtime();
    if (rt.ready.load(std::memory_order_acquire)) {
        // calls may come from any host thread (cargo test): bind the thread to the library's device
        cudaError_t e = cudaSetDevice(rt.device);
        if (e != cudaSuccess) {
            set_error("cudaSetDevice(%d): %s", rt.device, cudaGetErrorString(e));
            return SB_ERR_CUDA;
        }
        return SB_OK;
    }
    RtLock lk(rt.mu);
    return init_locked(rt, -1);
}

int Scratch::reserve(size_t bytes) {
    if (bytes <= cap) return SB_OK;
    Runtime& rt = runtime();
    if (ptr) {
        (void)rt;
        cudaDeviceSynchronize();  // earlier work on ANY stream may still read the old buffer
        cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&ptr, want);
    }
    if (e != cudaSuccess) {
        ptr = nullptr;
        set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return SB_ERR_OOM;
    }
    cap = want;
    return SB_OK;
}

void Scratch::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

// per-stream scratch slots (callers hold runtime().mu)
struct StreamSlots { Scratch slot[WS_NUM_SLOTS]; };
static std::map<cudaStream_t, std::unique_ptr<StreamSlots>> g_stream_ws;
Scratch& ws_slot(cudaStream_t st, int slot) {
    auto it = g_stream_ws.find(st);
    if (it == g_stream_ws.end()) it = g_stream_ws.emplace(st, std::unique_ptr<StreamSlots>(new StreamSlots())).first;
    return it->second->slot[slot];
}
void ws_release_stream(cudaStream_t st) {
    auto it = g_stream_ws.find(st);
    if (it == g_stream_ws.end()) return;
    for (int i = 0; i < WS_NUM_SLOTS; i++) it->second->slot[i].release();
    g_stream_ws.erase(it);
}
void ws_release_all() {
    for (auto& kv : g_stream_ws)
        for (int i = 0; i < WS_NUM_SLOTS; i++) kv.second->slot[i].release();
    g_stream_ws.clear();
}

// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void k_selftest_field(const F* a, const F* b, size_t n, F* o_ptx, F* o_port, F* o_add, F* o_sub, F* o_inv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i];
    o_ptx[i] = mul(x, y);
    o_port[i] = mul_portable(x, y);
    o_add[i] = add(x, y);
    o_sub[i] = sub(x, y);
    o_inv[i] = (i & 1) ? inv_safegcd(x) : (x.is_zero() ? x : ((i & 2) ? inv_binary(x) : inv(x)));  // all three inversion routines
}

// lazy-domain operations (field.cuh): PTX forms on operands anywhere in [0, 2p); raw results out
template <class F>
__global__ void k_selftest_lazy(const F* a, const F* b, size_t n, F* o_mul, F* o_sub, F* o_dbl, F* o_canon) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i];
    o_mul[i] = mul_lazy(x, y);
    o_sub[i] = (i & 1) ? sub_lazy(x, y) : add_lazy(x, neg_lazy(y));   // both routes to x - y
    o_dbl[i] = (i & 1) ? dbl_lazy(x) : add_lazy(x, x);
    F c = canon(x);
    if (is_zero_lazy(x)) c.v[7] |= 0x80000000u;  // flag bit (values are < 2^255)
    o_canon[i] = c;
}

// the same chain of group operations through the 4-warp cooperative forms (coop.cuh) and the single-lane forms; it
// visits every exceptional case: generic add, P + P through add, doubling, P + (-P), identity on either side
template <class F, bool COOP>
SB_D void selftest_chain(const Affine<F>& pa, const Affine<F>& pb, CoopBuf* sh, Affine<F>& out_t, Affine<F>& out_u, Affine<F>& out_v) {
    auto addp = [&](XYZZ<F>& x, const XYZZ<F>& y) {
        if constexpr (COOP) coop4_add(x, y, *sh);
        else xyzz_add<true>(x, y);
    };
    auto dblp = [&](XYZZ<F>& x) {
        if constexpr (COOP) coop4_double(x, *sh);
        else x = xyzz_double<true>(x);
    };
    XYZZ<F> a = XYZZ<F>::from_affine(pa), b = XYZZ<F>::from_affine(pb);
    XYZZ<F> t = a;
    addp(t, b);                 // generic (or an exceptional case when the inputs say so)
    XYZZ<F> t2 = t;
    addp(t, t2);                // P + P through the addition
    dblp(t);                    // doubling
    addp(t, b);                 // generic with non-trivial ZZ
    XYZZ<F> u = t;
    XYZZ<F> nt = t;
    nt.y = neg(nt.y);
    addp(u, nt);                // P + (-P) -> identity
    XYZZ<F> v = t;
    addp(v, u);                 // + identity
    addp(u, t);                 // identity + P
    dblp(u);
    out_t = xyzz_to_affine<true>(t);
    out_u = xyzz_to_affine<true>(u);
    out_v = xyzz_to_affine<true>(v);
}
template <class F>
__global__ void __launch_bounds__(COOP_THREADS) k_selftest_coop(const Affine<F>* pa, const Affine<F>* pb, size_t n, Affine<F>* out_coop, Affine<F>* out_plain) {
    __shared__ CoopBuf sh;
    const size_t i = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
    Affine<F> za, zb;
    za.x = F::zero(); za.y = F::zero(); zb = za;
    const Affine<F> a = i < n ? pa[i] : za, b = i < n ? pb[i] : zb;   // lanes past the end run the chain on identities (block-uniform trip counts)
    Affine<F> t, u, v;
    selftest_chain<F, true>(a, b, &sh, t, u, v);
    if (i < n && threadIdx.x < 32) { out_coop[3 * i] = t; out_coop[3 * i + 1] = u; out_coop[3 * i + 2] = v; }
    if (threadIdx.x < 32) {
        selftest_chain<F, false>(a, b, nullptr, t, u, v);
        if (i < n) { out_plain[3 * i] = t; out_plain[3 * i + 1] = u; out_plain[3 * i + 2] = v; }
    }
}

}  // namespace sb

using namespace sb;

extern "C" {

/* Group-law self test: for every pair (a[i], b[i]) of affine points the chain ((a+b)+(a+b)) doubled, + b, then P + (-P), + identity,
 * identity + P, through the 4-warp cooperative operations (out_coop) and the single-lane ones (out_plain); 3 affine points per pair. */
int sb_selftest_coop(int curve, const uint64_t* a_xy, const uint64_t* b_xy, size_t n, uint64_t* out_coop_xy, uint64_t* out_plain_xy) {
    if (!a_xy || !b_xy || !out_coop_xy || !out_plain_xy || !n) {
        set_error("sb_selftest_coop: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    char* d = nullptr;
    SB_CUDA_TRY(cudaMalloc(&d, n * 64 * 8));
    int rc = SB_OK;
    cudaError_t e = cudaMemcpyAsync(d, a_xy, n * 64, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n * 64, b_xy, n * 64, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((n + 31) / 32);
        if (curve == CURVE_BN256)
            k_selftest_coop<Fq><<<blocks, COOP_THREADS, 0, rt.stream>>>((const Affine<Fq>*)d, (const Affine<Fq>*)(d + n * 64), n, (Affine<Fq>*)(d + n * 128), (Affine<Fq>*)(d + n * 320));
        else
            k_selftest_coop<Fr><<<blocks, COOP_THREADS, 0, rt.stream>>>((const Affine<Fr>*)d, (const Affine<Fr>*)(d + n * 64), n, (Affine<Fr>*)(d + n * 128), (Affine<Fr>*)(d + n * 320));
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_coop_xy, d + n * 128, n * 192, cudaMemcpyDeviceToHost, rt.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_plain_xy, d + n * 320, n * 192, cudaMemcpyDeviceToHost, rt.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rt.stream);
    if (e != cudaSuccess) {
        set_error("sb_selftest_coop: %s", cudaGetErrorString(e));
        rc = SB_ERR_CUDA;
    }
    cudaFree(d);
    return rc;
}

const char* sb_last_error(void) { return g_err; }
int sb_version(void) { return 1; }

uint64_t sb_launch_count(void) { return g_launches.load(); }

void sb_profile_enable(int on) { g_profile.store(on != 0); }

/* Per-tag sums of the instrumented launch groups since the last call (synchronises the recorded events).
 * Arrays of SB_PROF_NUM_TAGS entries each; any may be NULL. */
int sb_profile_collect(double* total_ms, uint64_t* total_units, uint64_t* launches) {
    for (int t = 0; t < PROF_NUM_TAGS; t++) {
        if (total_ms) total_ms[t] = 0;
        if (total_units) total_units[t] = 0;
        if (launches) launches[t] = 0;
    }
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& rec : g_prof) {
        cudaError_t e = cudaEventSynchronize(rec.e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, rec.e0, rec.e1);
        if (e != cudaSuccess) {
            set_error("sb_profile_collect: %s", cudaGetErrorString(e));
            return SB_ERR_CUDA;
        }
        if (total_ms) total_ms[rec.tag] += ms;
        if (total_units) total_units[rec.tag] += rec.units;
        if (launches) launches[rec.tag] += 1;
        g_prof_pool.push_back(rec);
    }
    g_prof.clear();
    return SB_OK;
}

int sb_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

int sb_init(int device) {
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return init_locked(rt, device);
}

/* One process, several GPUs (`sb_init(devs, n)`): devices[0] becomes the primary device (every _device entry point and
 * every non-commit kernel runs there); keys registered from host memory afterwards are split block-cyclically over all n devices and
 * sb_msm / sb_msm_batch shard each commit over them, exchanging the 128-byte partial sums by peer copies.  n = 1 is sb_init. */
int sb_init_devices(const int* devices, int n) {
    if (!devices || n < 1 || n > 16) {
        set_error("sb_init_devices: bad argument");
        return SB_ERR_ARG;
    }
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    rt.ready.store(false);   // force re-initialisation of the primary even when the device is unchanged
    SB_TRY(init_locked(rt, devices[0]));
    for (int i = 1; i < n; i++) {
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) {
                set_error("sb_init_devices: device %d listed twice", devices[i]);
                return SB_ERR_ARG;
            }
        cudaDeviceProp prop;
        SB_CUDA_TRY(cudaGetDeviceProperties(&prop, devices[i]));
        if (prop.major < 10) {
            set_error("device %d is sm_%d%d; this library is built for sm_100a only", devices[i], prop.major, prop.minor);
            return SB_ERR_CUDA;
        }
        SB_CUDA_TRY(cudaSetDevice(devices[i]));
        Runtime::Dev d{devices[i], nullptr, prop.multiProcessorCount};
        SB_CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        rt.devs.push_back(d);
    }
    // peer access in both directions (partial sums travel by cudaMemcpyPeerAsync; NVLink when the GPUs share an NVSwitch)
    for (size_t i = 0; i < rt.devs.size(); i++) {
        SB_CUDA_TRY(cudaSetDevice(rt.devs[i].device));
        for (size_t j = 0; j < rt.devs.size(); j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, rt.devs[i].device, rt.devs[j].device);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(rt.devs[j].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", rt.devs[i].device, rt.devs[j].device, cudaGetErrorString(e));
                    return SB_ERR_CUDA;
                }
                cudaGetLastError();
            }
        }
    }
    SB_CUDA_TRY(cudaSetDevice(rt.device));
    return SB_OK;
}

int sb_num_devices(void) { return (int)runtime().devs.size(); }

void sb_shutdown(void) {
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    if (rt.stream) {
        cudaStreamSynchronize(rt.stream);
        cudaDeviceSynchronize();
        msm_graphs_drop(nullptr, nullptr, nullptr, true);
        ws_release_all();
        cudaStreamDestroy(rt.stream);
        rt.stream = nullptr;
    }
    rt.ready.store(false);
}

/* Frees the per-stream device scratch the library keeps for `stream` (call before destroying a stream that was
 * passed to _device entry points; the stream must be idle). */
void sb_stream_release(void* stream) {
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    if (stream) cudaStreamSynchronize((cudaStream_t)stream);
    msm_graphs_drop(nullptr, (cudaStream_t)stream, nullptr, stream == nullptr);
    ws_release_stream((cudaStream_t)stream);
}

int sb_selftest_field(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_mul_ptx, uint64_t* out_mul_portable,
                      uint64_t* out_add, uint64_t* out_sub, uint64_t* out_inv) {
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    char* d = nullptr;
    size_t bytes = n * 32;
    SB_CUDA_TRY(cudaMalloc(&d, bytes * 7));
    int rc = SB_OK;
    do {
        if (cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, rt.stream) != cudaSuccess ||
            cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyHostToDevice, rt.stream) != cudaSuccess) {
            set_error("selftest H2D failed");
            rc = SB_ERR_CUDA;
            break;
        }
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (field == FIELD_FR)
            k_selftest_field<Fr><<<blocks, 128, 0, rt.stream>>>((const Fr*)d, (const Fr*)(d + bytes), n, (Fr*)(d + 2 * bytes), (Fr*)(d + 3 * bytes),
                                                               (Fr*)(d + 4 * bytes), (Fr*)(d + 5 * bytes), (Fr*)(d + 6 * bytes));
        else
            k_selftest_field<Fq><<<blocks, 128, 0, rt.stream>>>((const Fq*)d, (const Fq*)(d + bytes), n, (Fq*)(d + 2 * bytes), (Fq*)(d + 3 * bytes),
                                                               (Fq*)(d + 4 * bytes), (Fq*)(d + 5 * bytes), (Fq*)(d + 6 * bytes));
        cudaError_t e = cudaGetLastError();
        uint64_t* outs[5] = {out_mul_ptx, out_mul_portable, out_add, out_sub, out_inv};
        for (int k = 0; k < 5 && e == cudaSuccess; k++) e = cudaMemcpyAsync(outs[k], d + (2 + k) * bytes, bytes, cudaMemcpyDeviceToHost, rt.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt.stream);
        if (e != cudaSuccess) {
            set_error("selftest failed: %s", cudaGetErrorString(e));
            rc = SB_ERR_CUDA;
        }
    } while (0);
    cudaFree(d);
    return rc;
}

int sb_selftest_lazy(int field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out_mul, uint64_t* out_sub, uint64_t* out_dbl,
                     uint64_t* out_canon) {
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    char* d = nullptr;
    size_t bytes = n * 32;
    SB_CUDA_TRY(cudaMalloc(&d, bytes * 6));
    int rc = SB_OK;
    do {
        if (cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, rt.stream) != cudaSuccess ||
            cudaMemcpyAsync(d + bytes, b, bytes, cudaMemcpyHostToDevice, rt.stream) != cudaSuccess) {
            set_error("selftest H2D failed");
            rc = SB_ERR_CUDA;
            break;
        }
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (field == FIELD_FR)
            k_selftest_lazy<Fr><<<blocks, 128, 0, rt.stream>>>((const Fr*)d, (const Fr*)(d + bytes), n, (Fr*)(d + 2 * bytes), (Fr*)(d + 3 * bytes),
                                                              (Fr*)(d + 4 * bytes), (Fr*)(d + 5 * bytes));
        else
            k_selftest_lazy<Fq><<<blocks, 128, 0, rt.stream>>>((const Fq*)d, (const Fq*)(d + bytes), n, (Fq*)(d + 2 * bytes), (Fq*)(d + 3 * bytes),
                                                              (Fq*)(d + 4 * bytes), (Fq*)(d + 5 * bytes));
        cudaError_t e = cudaGetLastError();
        uint64_t* outs[4] = {out_mul, out_sub, out_dbl, out_canon};
        for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaMemcpyAsync(outs[k], d + (2 + k) * bytes, bytes, cudaMemcpyDeviceToHost, rt.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt.stream);
        if (e != cudaSuccess) {
            set_error("selftest failed: %s", cudaGetErrorString(e));
            rc = SB_ERR_CUDA;
        }
    } while (0);
    cudaFree(d);
    return rc;
}

}  // extern "C"
