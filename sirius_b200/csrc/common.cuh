// common.cuh -- error plumbing, the library stream and a grow-only device workspace.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/sirius_b200.h"

namespace sb {

void set_error(const char* fmt, ...);
void count_launch();

// Optional CUDA-event instrumentation of the dominant kernel (MSM bucket accumulation), used by bench.py to
// report the roofline of that kernel from inside the timed region.
enum ProfTag : int {
    PROF_DECOMPOSE = 0, PROF_SORT = 1, PROF_ACCUMULATE = 2, PROF_FIXUP = 3, PROF_REDUCE = 4, PROF_FINALIZE = 5,
    PROF_CROSS_TERMS = 6, PROF_FOLD = 7, PROF_NTT = 8, PROF_PG = 9, PROF_NUM_TAGS = 10
};
bool profile_enabled();
int profile_begin(cudaStream_t st, int tag, uint64_t units);
void profile_end(cudaStream_t st, int handle);
struct ProfScope {  // CUDA events around the launches issued while the scope is alive (only when profiling is on)
    cudaStream_t st;
    int h;
    ProfScope(cudaStream_t s, int tag, uint64_t units = 0) : st(s), h(profile_begin(s, tag, units)) {}
    ~ProfScope() { profile_end(st, h); }
};

#define SB_CUDA_TRY(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            sb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SB_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

#define SB_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc != SB_OK) return _rc; \
    } while (0)

// every kernel launch of this library goes through here: counts launches (bench.py's gpu_launches)
#define SB_KERNEL_CHECK()                   \
    do {                                    \
        sb::count_launch();                 \
        SB_CUDA_TRY(cudaGetLastError());    \
    } while (0)

// Library-wide state for the device this process drives (one process per GPU).
using RtMutex = std::recursive_mutex;
using RtLock = std::lock_guard<std::recursive_mutex>;
struct Runtime {
    RtMutex mu;             // serialises the HOST side of library calls (commit is called sequentially by the
                            // reference, SURVEY 3.3, but cargo test runs tests on many threads).  Recursive: a
                            // host-memory entry point holds it across stage -> enqueue -> download -> sync and
                            // calls the matching _device entry point inside.
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;   // the library's own stream: host-memory entry points run here
    std::atomic<bool> ready{false};
    // sb_init_devices: the devices a single process drives (devs[0] = the primary above).  Host-memory commits against a
    // key registered while several devices are active are sharded over them inside the library (msm.cu).
    struct Dev { int device; cudaStream_t stream; int sm_count; };
    std::vector<Dev> devs;
};
Runtime& runtime();
int ensure_runtime();

// Grow-only scratch buffer; contents are dead between calls.
struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

// Device scratch is kept PER CUDA STREAM: two threads driving two streams never share a buffer, and calls on
// one stream are ordered by the stream itself, so every _device entry point is re-entrant across streams
// (SURVEY 8b "Threading").  The caller holds runtime().mu while it looks a slot up and enqueues.
enum WsSlot : int {
    WS_MSM = 0, WS_EXPR_ARGS, WS_EXPR_STAGE, WS_FOLD_CONSTS, WS_FOLD_STAGE, WS_NTT_TMP, WS_NTT_STAGE, WS_PG, WS_LINCOMB_ARGS,
    WS_PG_STAGE, WS_BI_STAGE, WS_LK, WS_LK_STAGE, WS_INV_SHIFT, WS_MSM_HOST, WS_NUM_SLOTS
};
Scratch& ws_slot(cudaStream_t st, int slot);
void ws_release_stream(cudaStream_t st);   // frees every slot of `st` (all streams when st == nullptr-all flag is used by sb_shutdown)
void ws_release_all();

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sb
