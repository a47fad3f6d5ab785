// common.cuh -- error plumbing, the library stream and a grow-only device workspace.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

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
struct Runtime {
    std::mutex mu;          // serialises library calls that share the workspace (commit is called sequentially
                            // by the reference, SURVEY 3.3, but cargo test runs tests on many threads)
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool ready = false;
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

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sb
