// msm.cu -- Pedersen commitment = multi-scalar multiplication on one B200.
//
// Replaces the body of `CommitmentKey::commit` (reference src/commitment.rs:81-90), i.e. halo2's CPU
// `best_multiexp(v, &ck[..v.len()]).to_affine()`, by a device pipeline:
//
//   register (once per CommitmentKey):  table[w][i] = 2^(c*w) * ck[i]            k_precompute
//   commit:  scalars --from-Montgomery, signed c-bit digits--> (bucket, table index, sign)      k_decompose
//            counting sort of the n*W digit entries by bucket                    k_scan_* / k_scatter
//            bucket sums: fixed-size chunks of the sorted entries, one chunk per thread, mixed XYZZ adds
//                         (skew-proof: a bucket of any size is split across as many threads as it needs)
//                                                                                k_accumulate, k_fixup
//            sum_b (b+1) * B_b: row/column sums, octal digit sums, quad-lane tail  k_rowcol_sums / k_digit_sums / k_weighted_finish / k_reduce_final
//            XYZZ -> affine                                                      k_finalize
//
// Because every window's base multiple is precomputed, all W windows share ONE set of 2^(c-1) buckets and
// there is no per-window doubling chain at the end.  All arithmetic is exact, so any grouping of the
// additions gives the reference's bit pattern.
//
// Roofline: the work is ~W mixed additions per scalar (8M+2S 254-bit Montgomery products each) -- integer-pipe
// bound; algorithmic HBM traffic is 96 B/point.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include <new>

#include "common.cuh"
#include "curve.cuh"
#include "affine.cuh"
#include "quad.cuh"
#include "coop.cuh"

namespace sb {

constexpr int LS_MIN_LOG = 4;   // sorted entries per accumulate chunk (one thread each): 2^4 .. 2^8, chosen per call
constexpr int LS_MAX_LOG = 8;
constexpr int FIX_SEQ = 6;      // buckets split in <= fix_seq pieces (>= FIX_SEQ, chosen per plan) are summed serially by their own lane ..
constexpr int FIX_TREE = 1024;  // .. up to FIX_TREE pieces by a cooperative block (k_fixup_tree: lane-strided runs + a 5-step lane tree) ..
                                // .. and beyond that (all-equal scalars) by the 256-thread k_fixup_heavy
constexpr int HEAVY_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // items per thread in the scan kernels
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

// ------------------------------------------------------------------------------------------------
// 128-bit global loads/stores of field-sized objects
// ------------------------------------------------------------------------------------------------
template <class T>
SB_D T load_vec(const T* p) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiples only");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
    return r;
}
template <class T>
SB_D T load_vec_nc(const T* p) {  // read-only path for data not written by this kernel
    static_assert(sizeof(T) % 16 == 0, "16-byte multiples only");
    T r;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = __ldg(s + i);
    return r;
}
template <class T>
SB_D void store_vec(T* p, const T& v) {
    static_assert(sizeof(T) % 16 == 0, "16-byte multiples only");
    uint4* d = reinterpret_cast<uint4*>(p);
    const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); i++) d[i] = s[i];
}

// ------------------------------------------------------------------------------------------------
// warp helpers on XYZZ points
// ------------------------------------------------------------------------------------------------
template <class F>
SB_D XYZZ<F> shfl_xor_point(const XYZZ<F>& v, int mask) {
    XYZZ<F> r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
    return r;
}
// ------------------------------------------------------------------------------------------------
// register-time: table of window multiples
// ------------------------------------------------------------------------------------------------
template <class F>
__global__ void k_precompute(const Affine<F>* __restrict__ bases, size_t n, int c, int W, Affine<F>* __restrict__ table) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = load_vec_nc(bases + i);
    store_vec(table + i, p);
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    for (int w = 1; w < W; w++) {
        for (int j = 0; j < c; j++) xyzz_double_call(acc);
        Affine<F> q = xyzz_to_affine<false>(acc);
        store_vec(table + (size_t)w * n + i, q);
    }
}

// Synthetic key: out[i] = [first + i + 1] * G (the benches' stand-in for CommitmentKey::setup, whose
// hash_to_curve lives in the un-vendored halo2curves).  One thread per point: double-and-add on
// the index, then normalise.
template <class F>
__global__ void k_index_multiples(Affine<F> g, uint64_t first, size_t n, Affine<F>* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = first + i + 1;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (int bit = 63 - __clzll((long long)s); bit >= 0; bit--) {
        xyzz_double_call(acc);
        if ((s >> bit) & 1) {
            XYZZ<F> gg = XYZZ<F>::from_affine(g);
            xyzz_add_call(acc, gg);
        }
    }
    store_vec(out + i, xyzz_to_affine<false>(acc));
}

// is_on_curve for every point of a key file (reference src/commitment.rs:148-157: `p.is_on_curve()` over the
// loaded cache; halo2curves accepts the identity (0,0)).  bad_count += number of points off y^2 = x^3 + b.
template <class F>
__global__ void k_on_curve(const Affine<F>* __restrict__ pts, size_t n, int b_small, int b_negative, unsigned long long* bad_count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = load_vec_nc(pts + i);
    if (p.is_identity()) return;
    F bb = F::zero();
    bb.v[0] = (uint32_t)b_small;
    bb = to_mont(bb);
    if (b_negative) bb = neg(bb);
    // coordinates must be canonical (< p) for the comparison to mean anything: reduce-once keeps valid inputs unchanged
    F lhs = sqr(p.y);
    F rhs = add(mul(sqr(p.x), p.x), bb);
    bool canonical = true;
    {
        uint32_t t[8];
        for (int k = 0; k < 8; k++) t[k] = p.x.v[k];
        reduce_once_portable<typename F::Params>(t);
        for (int k = 0; k < 8; k++) canonical &= (t[k] == p.x.v[k]);
        for (int k = 0; k < 8; k++) t[k] = p.y.v[k];
        reduce_once_portable<typename F::Params>(t);
        for (int k = 0; k < 8; k++) canonical &= (t[k] == p.y.v[k]);
    }
    if (!canonical || lhs != rhs) atomicAdd(bad_count, 1ull);
}

// ------------------------------------------------------------------------------------------------
// commit-time: digits, counting sort
// ------------------------------------------------------------------------------------------------
// dig[w*n + i] = (0 (digit 0) or ((|d|) | sign<<31), rank in its bucket) for the signed base-2^c digit d of scalar i, window w.
// Batched: `total` = batch * n scalars (batch b = i / n shares the same key, its buckets are [b*K, (b+1)*K)),
// the scalar vectors being `stride` elements apart.
template <class S>
__global__ void k_decompose(const S* __restrict__ scalars, uint32_t n, uint32_t total, size_t stride, uint32_t K, int c, int W,
                            uint2* __restrict__ dig, uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t batch = i / n;
    const uint32_t bucket_base = batch * K;
    S s = from_mont(load_vec_nc(scalars + (size_t)batch * stride + (i - batch * n)));  // canonical integer (`to_repr()`)
    uint32_t limb[9];
#pragma unroll
    for (int k = 0; k < 8; k++) limb[k] = s.v[k];
    limb[8] = 0;
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1;
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        int off = w * c;
        int li = off >> 5, sh = off & 31;
        uint32_t raw = 0;
        if (li < 8) {
            uint64_t two = (uint64_t)limb[li] | ((uint64_t)limb[li + 1] << 32);
            raw = (uint32_t)(two >> sh) & mask;
        }
        raw += carry;
        uint32_t out = 0;
        if (raw > half) {
            uint32_t mag = (1u << c) - raw;  // digit = raw - 2^c  (negative, |d| < 2^(c-1)); raw == 2^c -> 0
            carry = 1;
            if (mag) out = mag | 0x80000000u;
        } else {
            carry = 0;
            out = raw;
        }
        // the counter's old value is this entry's rank inside its bucket: the scatter pass needs no second atomic
        uint32_t rank = 0;
        if (out) rank = atomicAdd(&counts[bucket_base + (out & 0x7fffffffu) - 1], 1u);
        dig[(size_t)w * total + i] = make_uint2(out, rank);
    }
}

// The three scan kernels below build, in one pass each, the bucket offsets of every reduction round (affine.cuh):
// blockIdx.y = r scans cnt_r(b) = ceil(counts[b] / 2^r); row r of `offsets` (KB + 1 entries) and of `tile_sums`
// (SCAN_MAX_TILES entries).  r = 0 is the plain counting-sort scan.
constexpr uint32_t SCAN_MAX_TILES = 8192;
SB_D uint32_t round_count(uint32_t c, uint32_t r) { return (c + ((1u << r) - 1u)) >> r; }

__global__ void k_scan_tile_sums(const uint32_t* __restrict__ counts, uint32_t K, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const uint32_t r = blockIdx.y;
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t idx = base + k;
        s += (idx < K) ? round_count(counts[idx], r) : 0u;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int wi = 0; wi < SCAN_THREADS / 32; wi++) t += warp_tot[wi];
        tile_sums[r * SCAN_MAX_TILES + blockIdx.x] = t;
    }
}

// exclusive scan of up to 8192 tile sums in one block of 1024 threads (8 items per thread); one block per round
__global__ void k_scan_tiles(uint32_t* tile_sums_all, uint32_t num_tiles) {
    __shared__ uint32_t sh[1024];
    uint32_t* tile_sums = tile_sums_all + blockIdx.x * SCAN_MAX_TILES;
    uint32_t v[8];
    uint32_t base = threadIdx.x * 8;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = (base + k < num_tiles) ? tile_sums[base + k] : 0u;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < num_tiles) tile_sums[base + k] = excl;
        excl += v[k];
    }
}

// offsets[r][i] = exclusive prefix of cnt_r; offsets[r][K] = total
__global__ void k_scan_apply(const uint32_t* __restrict__ counts, uint32_t K, const uint32_t* __restrict__ tile_excl_all,
                             uint32_t* __restrict__ offsets_all) {
    __shared__ uint32_t sh[SCAN_THREADS];
    const uint32_t r = blockIdx.y;
    const uint32_t* tile_excl = tile_excl_all + r * SCAN_MAX_TILES;
    uint32_t* offsets = offsets_all + (size_t)r * ((size_t)K + 1);
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t idx = base + k;
        v[k] = (idx < K) ? round_count(counts[idx], r) : 0u;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < SCAN_THREADS; d <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = tile_excl[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t idx = base + k;
        if (idx < K) offsets[idx] = run;
        run += v[k];
        if (idx == K - 1) offsets[K] = run;
    }
}

// entry (table index | sign<<31) placed at offsets[bucket] + its rank (from k_decompose); the bucket of a sorted position is
// recovered from `offsets` (k_chunk_heads + a walk in k_accumulate), so no key array is written
__global__ void k_scatter(const uint2* __restrict__ dig, uint32_t n, uint32_t total, uint32_t K, uint32_t n_ck, int W,
                          const uint32_t* __restrict__ offsets, uint32_t* __restrict__ eidx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t batch = i / n;
    const uint32_t pt = i - batch * n;
    for (int w = 0; w < W; w++) {
        const uint2 dr = dig[(size_t)w * total + i];
        const uint32_t d = dr.x;
        if (d) {
            uint32_t b = batch * K + (d & 0x7fffffffu) - 1;
            eidx[offsets[b] + dr.y] = ((uint32_t)w * n_ck + pt) | (d & 0x80000000u);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// commit-time, large inputs: two-level partition sort (no per-entry global atomics)
// ------------------------------------------------------------------------------------------------
// The counting sort above pays one global atomic per digit entry (n*W of them: ~0.7 ms of a 16 ms step).  For
// 512 <= KB <= 2^20 buckets the entries are instead partitioned twice through shared-memory histograms:
//   level 1  partition = bucket >> 8 (P <= 4096 partitions).  k_digits_parts writes the digits and counts entries per
//            partition (shared-memory counters, P global atomics per block of 1024 scalars); k_scan_small turns the
//            counts into partition offsets; k_partition re-reads the digits and moves (bucket, table index | sign)
//            pairs into their partition, reserving one run per (block, partition) with a single global atomic.
//   level 2  inside a partition only 256 buckets occur: k_bucket_hist counts them in shared memory (-> counts[],
//            scanned by the usual k_scan_* into offsets[]); k_place reserves one run per (tile, bucket) and writes the
//            sorted entries eidx[].
// Order inside a bucket differs from the atomic path; the sums are exact, so the commitment does not.
constexpr int PART_LO_BITS = 8;
constexpr uint32_t PART_BUCKETS = 1u << PART_LO_BITS;   // buckets per partition
constexpr uint32_t PART_MAX = 4096;                     // partitions (shared-memory counters of k_partition: 2 * 4 B each)
constexpr int PART_SPT = 4;                             // scalars per thread in the level-1 kernels
constexpr int PART_EPT = 8;                             // entries per thread and tile in k_place

// digit w of a canonical scalar held in limb[0..8] (limb[8] = 0), with the carry of the signed recoding
SB_D uint32_t signed_digit(const uint32_t* limb, int w, int c, uint32_t& carry) {
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    const int off = w * c;
    const int li = off >> 5, sh = off & 31;
    uint32_t raw = 0;
    if (li < 8) {
        uint64_t two = (uint64_t)limb[li] | ((uint64_t)limb[li + 1] << 32);
        raw = (uint32_t)(two >> sh) & mask;
    }
    raw += carry;
    if (raw > half) {
        const uint32_t mag = (1u << c) - raw;  // digit = raw - 2^c (negative); raw == 2^c -> 0
        carry = 1;
        return mag ? (mag | 0x80000000u) : 0u;
    }
    carry = 0;
    return raw;
}

template <class S>
__global__ void __launch_bounds__(256)
k_digits_parts(const S* __restrict__ scalars, uint32_t n, uint32_t total, size_t stride, uint32_t K, int c, int W, uint32_t P,
               uint32_t* __restrict__ dig, uint32_t* __restrict__ part_count) {
    extern __shared__ uint32_t sh_cnt[];
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < PART_SPT; s++) {
        const uint32_t i = (blockIdx.x * PART_SPT + s) * blockDim.x + threadIdx.x;
        if (i >= total) break;
        const uint32_t batch = i / n;
        const uint32_t bucket_base = batch * K;
        S sc = from_mont(load_vec_nc(scalars + (size_t)batch * stride + (i - batch * n)));
        uint32_t limb[9];
#pragma unroll
        for (int k = 0; k < 8; k++) limb[k] = sc.v[k];
        limb[8] = 0;
        uint32_t carry = 0;
        for (int w = 0; w < W; w++) {
            const uint32_t out = signed_digit(limb, w, c, carry);
            dig[(size_t)w * total + i] = out;
            if (out) atomicAdd(&sh_cnt[(bucket_base + (out & 0x7fffffffu) - 1) >> PART_LO_BITS], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
        if (sh_cnt[i]) atomicAdd(&part_count[i], sh_cnt[i]);
}

// exclusive scan of up to 8192 counters in one block of 1024 threads: off[i] = sum_{j<i} cnt[j], off[n] = total
__global__ void k_scan_small(const uint32_t* __restrict__ cnt, uint32_t n, uint32_t* __restrict__ off) {
    __shared__ uint32_t sh[1024];
    uint32_t v[8];
    const uint32_t base = threadIdx.x * 8;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = (base + k < n) ? cnt[base + k] : 0u;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (base + k < n) off[base + k] = excl;
        excl += v[k];
        if (base + k + 1 == n) off[n] = excl;
    }
}

// exclusive prefix of cnt[0..m) into off[0..m) by the 256 threads of a block (tmp: 256 words of shared memory)
SB_D void block_excl_scan(const uint32_t* cnt, uint32_t* off, uint32_t m, uint32_t* tmp) {
    const uint32_t ch = (m + 255u) / 256u;
    const uint32_t lo = threadIdx.x * ch, hi = (lo + ch < m) ? lo + ch : m;
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += cnt[i];
    tmp[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const uint32_t t = (threadIdx.x >= (unsigned)d) ? tmp[threadIdx.x - d] : 0u;
        __syncthreads();
        tmp[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = tmp[threadIdx.x] - s;
    for (uint32_t i = lo; i < hi; i++) {
        off[i] = run;
        run += cnt[i];
    }
    __syncthreads();
}

// One block = 256 scalars = 256 * W entries: (bucket, table index | sign) pairs are ranked by level-1 partition in
// shared memory and leave as one contiguous run per partition (coalesced 8-byte stores; a run is reserved in the
// partition with a single global atomic).  Shared memory: cnt[P] | loc[P] | base[P] | tmp[256] | stage[256 * W] pairs.
__global__ void __launch_bounds__(256)
k_partition(const uint32_t* __restrict__ dig, uint32_t n, uint32_t total, uint32_t K, uint32_t n_ck, int W, uint32_t P,
            const uint32_t* __restrict__ part_off, uint32_t* __restrict__ part_cursor, uint2* __restrict__ out1) {
    extern __shared__ uint32_t sh[];
    uint32_t* cnt = sh;               // entries of this block per partition, later the running local rank
    uint32_t* loc = sh + P;           // start of the partition's run inside the staging area
    uint32_t* base = sh + 2 * P;      // start of the run inside the partition (global)
    uint32_t* tmp = sh + 3 * P;
    uint2* stage = reinterpret_cast<uint2*>(sh + 3 * P + 256 + ((3 * P) & 1u));   // 8-byte aligned
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < total;
    const uint32_t batch = live ? i / n : 0;
    const uint32_t pt = i - batch * n;
    const uint32_t bucket_base = batch * K;
    if (live)
        for (int w = 0; w < W; w++) {
            const uint32_t d = dig[(size_t)w * total + i];
            if (d) atomicAdd(&cnt[(bucket_base + (d & 0x7fffffffu) - 1) >> PART_LO_BITS], 1u);
        }
    __syncthreads();
    block_excl_scan(cnt, loc, P, tmp);
    for (uint32_t q = threadIdx.x; q < P; q += blockDim.x) {
        const uint32_t m = cnt[q];
        if (m) base[q] = part_off[q] + atomicAdd(&part_cursor[q], m);
        cnt[q] = 0;
    }
    __syncthreads();
    if (live)
        for (int w = 0; w < W; w++) {
            const uint32_t d = dig[(size_t)w * total + i];
            if (d) {
                const uint32_t b = bucket_base + (d & 0x7fffffffu) - 1;
                const uint32_t hi = b >> PART_LO_BITS;
                const uint32_t r = atomicAdd(&cnt[hi], 1u);
                stage[loc[hi] + r] = make_uint2(b, ((uint32_t)w * n_ck + pt) | (d & 0x80000000u));
            }
        }
    __syncthreads();
    // the block's entry count = end of the last partition's run
    const uint32_t m_total = loc[P - 1] + cnt[P - 1];
    for (uint32_t idx = threadIdx.x; idx < m_total; idx += blockDim.x) {
        const uint2 e = stage[idx];
        const uint32_t hi = e.x >> PART_LO_BITS;
        out1[base[hi] + (idx - loc[hi])] = e;
    }
}

// grid (P, S): block (hi, s) strides over partition hi and counts its 256 buckets
__global__ void __launch_bounds__(256)
k_bucket_hist(const uint2* __restrict__ out1, const uint32_t* __restrict__ part_off, uint32_t KB, uint32_t* __restrict__ counts) {
    __shared__ uint32_t h[PART_BUCKETS];
    const uint32_t hi = blockIdx.x;
    const uint32_t start = part_off[hi], end = part_off[hi + 1];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t pos = (uint64_t)start + blockIdx.y * 256u + threadIdx.x; pos < end; pos += (uint64_t)gridDim.y * 256u)
        atomicAdd(&h[out1[pos].x & (PART_BUCKETS - 1)], 1u);
    __syncthreads();
    const uint32_t b = hi * PART_BUCKETS + threadIdx.x;
    if (b < KB && h[threadIdx.x]) atomicAdd(&counts[b], h[threadIdx.x]);
}

// grid (P, S): tiles of 256 * PART_EPT entries of one partition are ranked by bucket in shared memory and leave as one
// contiguous run per bucket (a run is reserved in eidx[] with a single global atomic per (tile, bucket))
__global__ void __launch_bounds__(256)
k_place(const uint2* __restrict__ out1, const uint32_t* __restrict__ part_off, const uint32_t* __restrict__ offsets,
        uint32_t* __restrict__ cursor, uint32_t KB, uint32_t* __restrict__ eidx) {
    __shared__ uint32_t cnt[PART_BUCKETS];
    __shared__ uint32_t loc[PART_BUCKETS];
    __shared__ uint32_t base[PART_BUCKETS];
    __shared__ uint32_t tmp[256];
    __shared__ uint2 stage[256 * PART_EPT];
    const uint32_t hi = blockIdx.x;
    const uint32_t start = part_off[hi], end = part_off[hi + 1];
    const uint32_t TILE = 256u * PART_EPT;
    for (uint64_t tile = (uint64_t)start + (uint64_t)blockIdx.y * TILE; tile < end; tile += (uint64_t)gridDim.y * TILE) {
        cnt[threadIdx.x] = 0;
        __syncthreads();
        uint2 e[PART_EPT];
#pragma unroll
        for (int j = 0; j < PART_EPT; j++) {
            const uint64_t pos = tile + (uint64_t)j * 256u + threadIdx.x;
            e[j] = make_uint2(0xffffffffu, 0u);
            if (pos < end) {
                e[j] = out1[pos];
                atomicAdd(&cnt[e[j].x & (PART_BUCKETS - 1)], 1u);
            }
        }
        __syncthreads();
        block_excl_scan(cnt, loc, PART_BUCKETS, tmp);
        {
            const uint32_t b = hi * PART_BUCKETS + threadIdx.x;
            const uint32_t m = cnt[threadIdx.x];
            if (m && b < KB) base[threadIdx.x] = offsets[b] + atomicAdd(&cursor[b], m);
            cnt[threadIdx.x] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PART_EPT; j++) {
            if (e[j].x != 0xffffffffu) {
                const uint32_t lo = e[j].x & (PART_BUCKETS - 1);
                const uint32_t r = atomicAdd(&cnt[lo], 1u);
                stage[loc[lo] + r] = e[j];
            }
        }
        __syncthreads();
        const uint32_t m_total = loc[PART_BUCKETS - 1] + cnt[PART_BUCKETS - 1];
        for (uint32_t idx = threadIdx.x; idx < m_total; idx += 256u) {
            const uint2 v = stage[idx];
            const uint32_t lo = v.x & (PART_BUCKETS - 1);
            eidx[base[lo] + (idx - loc[lo])] = v.y;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// bucket sums
// ------------------------------------------------------------------------------------------------
// chunk_head[t] = bucket that contains sorted position t*LS (LS = chunk length, any value; one thread per bucket, writes one entry per chunk start
// inside its range: #chunks writes in total instead of one key per entry)
__global__ void k_chunk_heads(const uint32_t* __restrict__ offsets, uint32_t KB, uint32_t LS, uint32_t* __restrict__ chunk_head) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= KB) return;
    const uint32_t o = offsets[b], o2 = offsets[b + 1];
    if (o2 == o) return;
    for (uint32_t t = (o + LS - 1) / LS; (uint64_t)t * LS < o2; t++) chunk_head[t] = b;
}

// One reduction round of the batched-affine bucket sums (affine.cuh): thread g produces outputs [g*B, (g+1)*B) of
// A_{r+1}; the block shares one inversion through a product tree in shared memory.  Threads below the active
// count of a tree level are contiguous, so a level costs ceil(active / 32) warp-products, ~4 per thread in all.
template <class F, bool INDEXED, int B>
__global__ void __launch_bounds__(PR_THREADS, 2)
k_pair_round(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ eidx, const uint32_t* __restrict__ off_in,
             const uint32_t* __restrict__ off_out, uint32_t KB, Affine<F>* __restrict__ dst) {
    __shared__ F node[PR_NODES];
    __shared__ F ninv[PR_NODES];
    const uint32_t total_out = off_out[KB];
    if ((uint64_t)blockIdx.x * (PR_THREADS * B) >= total_out) return;  // whole block idle (the grid is an upper bound)
    const int tid = threadIdx.x;
    const uint32_t g = blockIdx.x * PR_THREADS + tid;
    uint32_t pos[B];
    uint8_t kind[B];
    F cp[B];
    const PairSrc<F, INDEXED> src{pts, eidx};
    node[tid] = pair_forward<F, INDEXED, B>(src, off_in, off_out, KB, g, pos, kind, cp);
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < PR_LEVELS; l++) {
        if (tid < (PR_THREADS >> (l + 1))) pr_tree_up(node, l, tid);
        __syncthreads();
    }
    if (tid == 32 * (int)(blockIdx.x & 7u)) ninv[pr_level_off(PR_LEVELS)] = inv_safegcd(node[pr_level_off(PR_LEVELS)]);
    __syncthreads();
#pragma unroll 1
    for (int l = PR_LEVELS - 1; l >= 0; l--) {
        if (tid < (PR_THREADS >> l)) pr_tree_down(node, ninv, l, tid);
        __syncthreads();
    }
    pair_backward<F, INDEXED, B>(src, pos, kind, cp, ninv[tid], dst + (size_t)g * B);
}

// Thread t owns sorted entries [t*LS, (t+1)*LS).  A bucket lying entirely inside the chunk is
// written to buckets[]; a piece of a bucket that continues into a neighbouring chunk goes to PH[t] (piece
// starts at the chunk start) or PT[t] (piece ends at the chunk end) and is finished by k_fixup.
// DIRECT = false: entry -> window table (eidx); DIRECT = true: the entries ARE points (output of the affine rounds).
template <class F, int MINB, bool DIRECT>
__global__ void __launch_bounds__(128, MINB)
k_accumulate(const Affine<F>* __restrict__ table, const uint32_t* __restrict__ chunk_head, const uint32_t* __restrict__ eidx,
             const uint32_t* __restrict__ offsets, uint32_t KB, uint32_t LS, XYZZ<F>* __restrict__ buckets,
             XYZZ<F>* __restrict__ PH, XYZZ<F>* __restrict__ PT) {
    const uint32_t M = offsets[KB];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t start64 = (uint64_t)t * LS;
    if (start64 >= M) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (M - start < LS) ? M : start + LS;

    uint32_t cur = chunk_head[t];               // offsets[cur] <= start < offsets[cur + 1]
    uint32_t cur_end = offsets[cur + 1];
    uint32_t seg_start = start;
    XYZZ<F> acc = XYZZ<F>::identity();
    uint32_t e = DIRECT ? start : eidx[start];
    Affine<F> base = load_vec_nc(table + (e & 0x7fffffffu));
#pragma unroll 1
    for (uint32_t pos = start; pos < end; pos++) {
        // prefetch the next entry's base while this one is being added
        uint32_t e_next = 0;
        Affine<F> base_next;
        if (pos + 1 < end) {
            e_next = DIRECT ? pos + 1 : eidx[pos + 1];
            base_next = load_vec_nc(table + (e_next & 0x7fffffffu));
        }
        xyzz_madd_lazy(acc, base, (e >> 31) != 0);   // coordinates stay in [0, 2p) inside the run (field.cuh, lazy domain)
        if (pos + 1 == end || pos + 1 == cur_end) {
            const uint32_t seg_end = pos + 1;
            const uint32_t o = offsets[cur];
            acc = canon_point(acc);
            if (o == seg_start && cur_end == seg_end) store_vec(buckets + cur, acc);
            else if (seg_start == start) store_vec(PH + t, acc);
            else store_vec(PT + t, acc);
            acc = XYZZ<F>::identity();
            seg_start = seg_end;
            if (seg_end < end) {                 // next non-empty bucket
                do {
                    cur++;
                    cur_end = offsets[cur + 1];
                } while (cur_end <= seg_end);
            }
        }
        e = e_next;
        base = base_next;
    }
}

// One thread per bucket: empty buckets are set to the identity; a bucket split over 2..FIX_SEQ chunks is summed
// from its pieces here; a bucket split over more chunks (skewed scalars: many equal digits) is queued for
// k_fixup_heavy.  (Throughput-bound for batched commits: plain lanes, not quad-lane groups -- measured.)
template <class F>
__global__ void __launch_bounds__(128)
k_fixup(const uint32_t* __restrict__ offsets, uint32_t KB, uint32_t LS, XYZZ<F>* __restrict__ buckets,
        const XYZZ<F>* __restrict__ PH, const XYZZ<F>* __restrict__ PT, uint32_t* __restrict__ heavy_count,
        uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ tree_list, uint32_t fix_seq) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= KB) return;
    const uint32_t o = offsets[b], o2 = offsets[b + 1];
    if (o2 == o) {
        store_vec(buckets + b, XYZZ<F>::identity());
        return;
    }
    const uint32_t t0 = o / LS;
    const uint32_t np = (o2 - 1) / LS - t0 + 1;
    if (np == 1) return;  // written by k_accumulate
    if (np > fix_seq) {   // heavy_count[0]: buckets for k_fixup_heavy, heavy_count[1]: buckets for k_fixup_tree
        if (np > (uint32_t)FIX_TREE) heavy_list[atomicAdd(heavy_count, 1u)] = b;
        else tree_list[atomicAdd(heavy_count + 1, 1u)] = b;
        return;
    }
    XYZZ<F> acc = (o - t0 * LS) ? load_vec(PT + t0) : load_vec(PH + t0);
#pragma unroll 1
    for (uint32_t p = 1; p < np; p++) {
        XYZZ<F> q = load_vec(PH + t0 + p);
        xyzz_add_call(acc, q);
    }
    store_vec(buckets + b, acc);
}

// One block per queued bucket: threads stride over its pieces, then a shared-memory tree.
template <class F>
__global__ void __launch_bounds__(HEAVY_THREADS)
k_fixup_heavy(const uint32_t* __restrict__ offsets, uint32_t LS, XYZZ<F>* __restrict__ buckets, const XYZZ<F>* __restrict__ PH,
              const XYZZ<F>* __restrict__ PT, const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list) {
    __shared__ XYZZ<F> sh[HEAVY_THREADS];
    const uint32_t count = *heavy_count;
    for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
        const uint32_t b = heavy_list[h];
        const uint32_t o = offsets[b], o2 = offsets[b + 1];
        const uint32_t t0 = o / LS;
        const uint32_t np = (o2 - 1) / LS - t0 + 1;
        const bool head_is_tail = (o - t0 * LS) != 0;   // the bucket starts inside chunk t0: its first piece is PT[t0]
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t p = threadIdx.x; p < np; p += HEAVY_THREADS) {
            XYZZ<F> q = (p == 0 && head_is_tail) ? load_vec(PT + t0) : load_vec(PH + t0 + p);
            xyzz_add_call(acc, q);
        }
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int d = HEAVY_THREADS / 2; d >= 1; d >>= 1) {
            if ((int)threadIdx.x < d) {
                XYZZ<F> x = sh[threadIdx.x];
                XYZZ<F> y = sh[threadIdx.x + d];
                xyzz_add_call(x, y);
                sh[threadIdx.x] = x;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) store_vec(buckets + b, sh[0]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// sum_b (b+1) B_b : row/column sums, then octal digit sums of the two short vectors
// ------------------------------------------------------------------------------------------------
// With b = r * C + q (C = 2^lc columns, R = K / C rows):
//     sum_b (b+1) B_b = C * sum_r r * Row_r  +  sum_q (q+1) * Col_q,      Row_r = sum_q B_{r,q},  Col_q = sum_r B_{r,q}.
// Stage 1 (k_rowcol_sums) does the 2K plain additions with ordinary lanes (it is throughput-bound for batched
// commits): one warp per row / per column, a short serial run per lane and a 5-step shuffle tree.
// Stage 2 (k_digit_sums, k_weighted_finish) reduces a short vector E_0..E_{n-1} with weights (j + w0) by octal digit sums:
//     sum_j j * E_j = sum_i 8^i sum_{v=1..7} v * D[i][v],   D[i][v] = sum_{j: digit_i(j) = v} E_j   (plain sums, one warp each)
// and finishes with quad-lane additions (quad.cuh): 7-term weighted sums, Horner over the digit positions.
// Stage 3 (k_reduce_final) adds the two parts and normalises to affine.
template <class F>
SB_D XYZZ<F> warp_sum_call(XYZZ<F> v) {  // all lanes end with the warp total (scalar additions)
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        XYZZ<F> t = shfl_xor_point(v, d);
        xyzz_add_call(v, t);
    }
    return v;
}

// vec[batch][0..R) = row sums, vec[batch][R..R+C) = column sums
template <class F>
__global__ void __launch_bounds__(128)
k_rowcol_sums(const XYZZ<F>* __restrict__ buckets_all, int log_k, int lc, XYZZ<F>* __restrict__ vec_all) {
    const uint32_t K = 1u << log_k, C = 1u << lc, R = K >> lc;
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= R + C) return;
    const XYZZ<F>* buckets = buckets_all + (size_t)blockIdx.y * K;
    XYZZ<F> acc = XYZZ<F>::identity();
    if (w < R) {  // row w: C consecutive buckets
        const uint32_t per = (C + 31) >> 5;
#pragma unroll 1
        for (uint32_t t = 0; t < per; t++) {
            const uint32_t q = (uint32_t)lane * per + t;
            if (q < C) {
                XYZZ<F> x = load_vec(buckets + (size_t)w * C + q);
                xyzz_add_call(acc, x);
            }
        }
    } else {      // column w - R: R buckets, stride C
        const uint32_t col = w - R;
        const uint32_t per = (R + 31) >> 5;
#pragma unroll 1
        for (uint32_t t = 0; t < per; t++) {
            const uint32_t r = (uint32_t)lane * per + t;
            if (r < R) {
                XYZZ<F> x = load_vec(buckets + (size_t)r * C + col);
                xyzz_add_call(acc, x);
            }
        }
    }
    acc = warp_sum_call(acc);
    if (lane == 0) store_vec(vec_all + (size_t)blockIdx.y * (R + C) + w, acc);
}

// Octal digit sums, one warp (= one block, so every warp gets a scheduler of its own: the sums are a latency chain)
// per (part, digit position, digit value): D[batch][part][pos][v] = sum of the entries E_j whose octal digit `pos`
// of j equals v.  part 0 = the row sums (n = R entries), part 1 = the column sums (n = C entries).
template <class F>
__global__ void __launch_bounds__(32)
k_digit_sums(const XYZZ<F>* __restrict__ vec_all, int log_k, int lc, int max_pos, XYZZ<F>* __restrict__ D_all) {
    const uint32_t K = 1u << log_k, C = 1u << lc, R = K >> lc;
    const int which = blockIdx.x / (max_pos * 8);
    const int rem = blockIdx.x - which * (max_pos * 8);
    const int pos = rem >> 3, v = rem & 7;
    const uint32_t n = which == 0 ? R : C;
    const int log_n = which == 0 ? (log_k - lc) : lc;
    const XYZZ<F>* E = vec_all + (size_t)blockIdx.y * (R + C) + (which == 0 ? 0 : R);
    const int lane = threadIdx.x;
    const int npos = (log_n + 2) / 3 > 0 ? (log_n + 2) / 3 : 1;
    XYZZ<F> acc = XYZZ<F>::identity();
    if (pos < npos) {
        const int bits = (log_n - 3 * pos) < 3 ? (log_n - 3 * pos) : 3;
        if (v < (1 << bits)) {
            const uint32_t cnt = n >> bits;
            const uint32_t low_mask = (1u << (3 * pos)) - 1;
#pragma unroll 1
            for (uint32_t j = lane; j < cnt; j += 32) {
                const uint32_t idx = ((j >> (3 * pos)) << (3 * pos + bits)) | ((uint32_t)v << (3 * pos)) | (j & low_mask);
                XYZZ<F> x = load_vec(E + idx);
                xyzz_add_call(acc, x);
            }
        }
    }
    acc = warp_sum_call(acc);
    if (lane == 0) store_vec(D_all + (((size_t)blockIdx.y * 2 + which) * max_pos + pos) * 8 + v, acc);
}

// grid (2, batch): blockIdx.x = 0 -> X = C * sum_r r * Row_r (weights from 0), 1 -> Y = sum_q (q+1) * Col_q, from
// the digit sums:  sum_j j * E_j = sum_pos 8^pos * sum_{v=1..7} v * D[pos][v].
// Warp `pos` (< npos) forms the 7-term weighted sum of its position by suffix sums over the 8 four-lane groups
// (quad-lane additions); one more warp sums D[0][*] (= sum of all entries, the "+1" of the column weights); then
// lane 0 of warp 0 runs the Horner chain with single-lane doublings (8.9k cycles each against 10.6k for the
// quad-lane form, profiles/r1_microbench3.txt).
template <class F>
__global__ void __launch_bounds__(32 * 9)
k_weighted_finish(const XYZZ<F>* __restrict__ D_all, int log_k, int lc, int max_pos, XYZZ<F>* __restrict__ xy_all) {
    __shared__ XYZZ<F> Xp[8];
    __shared__ XYZZ<F> S0;
    const int which = blockIdx.x;
    const int log_n = which == 0 ? (log_k - lc) : lc;
    const int npos = (log_n + 2) / 3 > 0 ? (log_n + 2) / 3 : 1;
    const XYZZ<F>* D = D_all + (((size_t)blockIdx.y * 2 + which) * max_pos) * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2;
    if (warp < npos) {  // position `warp`: sum_v v * D[v]
        XYZZ<F> suf = load_vec(D + warp * 8 + g);
#pragma unroll 1
        for (int dist = 1; dist < 8; dist <<= 1) {
            XYZZ<F> t = group_shfl_down(suf, dist);
            if (g + dist < 8) quad_add(suf, t);
        }
        XYZZ<F> term = (g >= 1) ? suf : XYZZ<F>::identity();
        term = warp_group_sum(term);
        if (lane == 0) Xp[warp] = term;
    } else if (warp == max_pos) {  // the extra warp: total of all entries
        XYZZ<F> d = load_vec(D + g);
        d = warp_group_sum(d);
        if (lane == 0) S0 = d;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        XYZZ<F> horner = XYZZ<F>::identity();
        for (int i = npos - 1; i >= 0; i--) {
            xyzz_double_call(horner);
            xyzz_double_call(horner);
            xyzz_double_call(horner);
            XYZZ<F> q = Xp[i];
            xyzz_add_call(horner, q);
        }
        if (which == 1) {  // weights q + 1
            XYZZ<F> s0 = S0;
            xyzz_add_call(horner, s0);
        } else {           // the row part carries the factor C = 2^lc
            for (int k = 0; k < lc; k++) xyzz_double_call(horner);
        }
        store_vec(xy_all + (size_t)blockIdx.y * 2 + which, horner);
    }
}

// ------------------------------------------------------------------------------------------------
// the same tail on the 4-warp cooperative group law (coop.cuh): a block of 128 threads = 32 logical lanes, each
// addition ~4 dependent products instead of 14.  These kernels replace k_fixup / k_rowcol_sums / k_digit_sums /
// k_weighted_finish (kept above for A/B measurements, sb_msm_tune key 3).
// ------------------------------------------------------------------------------------------------
// fix-up: logical lane = bucket.  Pieces of a bucket are summed serially, but every addition is cooperative; lanes whose
// bucket has fewer pieces idle (identity operand) until the block's longest bucket is done.
template <class F>
__global__ void __launch_bounds__(COOP_THREADS)
k_fixup_coop(const uint32_t* __restrict__ offsets, uint32_t KB, uint32_t LS, XYZZ<F>* __restrict__ buckets,
             const XYZZ<F>* __restrict__ PH, const XYZZ<F>* __restrict__ PT, uint32_t* __restrict__ heavy_count,
             uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ tree_list, uint32_t fix_seq) {
    __shared__ CoopBuf sh;
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 32u + (uint32_t)lane;
    uint32_t np = 0, t0 = 0;
    bool head_tail = false;
    if (b < KB) {
        const uint32_t o = offsets[b], o2 = offsets[b + 1];
        if (o2 == o) {
            if (role == 0) store_vec(buckets + b, XYZZ<F>::identity());
        } else {
            t0 = o / LS;
            np = (o2 - 1) / LS - t0 + 1;
            head_tail = (o - t0 * LS) != 0;
            if (np == 1) np = 0;   // written by k_accumulate
            else if (np > fix_seq) {
                if (role == 0) {
                    if (np > (uint32_t)FIX_TREE) heavy_list[atomicAdd(heavy_count, 1u)] = b;
                    else tree_list[atomicAdd(heavy_count + 1, 1u)] = b;
                }
                np = 0;
            }
        }
    }
    const uint32_t np_max = __reduce_max_sync(0xffffffffu, np);   // identical in the 4 warps (replicated data)
    XYZZ<F> acc = XYZZ<F>::identity();
    if (np) acc = head_tail ? load_vec(PT + t0) : load_vec(PH + t0);
    XYZZ<F> q = XYZZ<F>::identity();
    if (1 < np) q = load_vec(PH + t0 + 1);
#pragma unroll 1
    for (uint32_t p = 1; p < np_max; p++) {
        XYZZ<F> q_next = XYZZ<F>::identity();   // the next piece is in flight while this one is added
        if (p + 1 < np) q_next = load_vec(PH + t0 + p + 1);
        coop4_add(acc, q, sh);
        q = q_next;
    }
    if (np && role == 0) store_vec(buckets + b, acc);
}

// Buckets split over FIX_SEQ < np <= FIX_TREE chunks (uniform scalars: the few buckets of the short top window, which
// take n / 2^(top bits) entries each; witness-like scalars: the small-value buckets of window 0).  A cooperative block
// takes TWO queued buckets, 16 logical lanes each: lane l adds pieces l, l + 16, .. (block-uniform trip count), then a
// 4-step xor tree -- 4 + np / 16 dependent additions where the lane-per-bucket loop of k_fixup would need np - 1.
// (A cooperative addition costs an SM ~2.7 us whatever the number of busy lanes -- a lone block already keeps the four
// integer pipes issuing -- so the kernel's time is blocks / 148 x additions x 2.7 us: two buckets per block halve it.)
template <class F>
__global__ void __launch_bounds__(COOP_THREADS)
k_fixup_tree(const uint32_t* __restrict__ offsets, uint32_t LS, XYZZ<F>* __restrict__ buckets, const XYZZ<F>* __restrict__ PH,
             const XYZZ<F>* __restrict__ PT, const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ tree_list) {
    __shared__ CoopBuf sh;
    const uint32_t lane = threadIdx.x & 31u, gl = lane & 15u, half = lane >> 4;
    const uint32_t count = heavy_count[1];
    for (uint32_t h0 = blockIdx.x * 2u; h0 < count; h0 += gridDim.x * 2u) {
        const uint32_t h = h0 + half;
        uint32_t b = 0, t0 = 0, np = 0;
        bool head_is_tail = false;
        if (h < count) {
            b = tree_list[h];
            const uint32_t o = offsets[b], o2 = offsets[b + 1];
            t0 = o / LS;
            np = (o2 - 1) / LS - t0 + 1;
            head_is_tail = (o - t0 * LS) != 0;   // the bucket starts inside chunk t0: its first piece is PT[t0]
        }
        const uint32_t np_max = __reduce_max_sync(0xffffffffu, np);   // identical in the 4 warps (replicated data)
        XYZZ<F> acc = XYZZ<F>::identity();
        if (gl < np) acc = (gl == 0 && head_is_tail) ? load_vec(PT + t0) : load_vec(PH + t0 + gl);
        XYZZ<F> q = XYZZ<F>::identity();
        if (gl + 16u < np) q = load_vec(PH + t0 + gl + 16u);
#pragma unroll 1
        for (uint32_t p = 16u; p < np_max; p += 16u) {
            XYZZ<F> q_next = XYZZ<F>::identity();   // the next piece is in flight while this one is added
            if (p + 16u + gl < np) q_next = load_vec(PH + t0 + p + 16u + gl);
            coop4_add(acc, q, sh);
            q = q_next;
        }
#pragma unroll 1
        for (int d = 8; d >= 1; d >>= 1) {
            XYZZ<F> t = coop_shfl_xor(acc, d);
            coop4_add(acc, t, sh);
        }
        if (threadIdx.x < 32u && gl == 0u && h < count) store_vec(buckets + b, acc);
    }
}

// row / column sums.  A row (or column) of cnt buckets is summed by a GROUP of G logical lanes (G = 32, 16, 8 or 4, chosen
// on the host): lane gl of the group adds entries gl, gl + G, .. serially, then a log2(G)-step xor tree.  A block holds
// 32 / G rows, grid (ceil((R + C) / (32 / G)), batch).  One row per block (G = 32) is the shortest chain, but every
// cooperative addition costs the same 14 warp-products whether 1 or 32 of its lanes do useful work (the xor tree is an
// all-reduce): with the thousands of short rows of a batched commitment the kernel is bound by that throughput, and
// narrower groups do the same sums with 2-3x fewer block-additions.
template <class F>
__global__ void __launch_bounds__(COOP_THREADS)
k_rowcol_coop(const XYZZ<F>* __restrict__ buckets_all, int log_k, int lc, int log_g, XYZZ<F>* __restrict__ vec_all) {
    __shared__ CoopBuf sh;
    const uint32_t K = 1u << log_k, C = 1u << lc, R = K >> lc;
    const uint32_t G = 1u << log_g, rows_per_block = 32u >> log_g;
    const uint32_t lane = threadIdx.x & 31u, gl = lane & (G - 1u);
    const uint32_t w = blockIdx.x * rows_per_block + (lane >> log_g);
    const XYZZ<F>* buckets = buckets_all + (size_t)blockIdx.y * K;
    const bool live = w < R + C, row = w < R;
    const uint32_t cnt = live ? (row ? C : R) : 0u;
    const uint32_t per = ((R > C ? R : C) + G - 1u) >> log_g;   // block-uniform trip count
    auto fetch = [&](uint32_t t) {
        const uint32_t j = t * G + gl;
        if (j >= cnt) return XYZZ<F>::identity();
        return row ? load_vec(buckets + (size_t)w * C + j) : load_vec(buckets + (size_t)j * C + (w - R));
    };
    XYZZ<F> acc = fetch(0);
    XYZZ<F> x = per > 1u ? fetch(1) : XYZZ<F>::identity();
#pragma unroll 1
    for (uint32_t t = 1; t < per; t++) {
        XYZZ<F> x_next = XYZZ<F>::identity();   // the next entry is in flight while this one is added
        if (t + 1u < per) x_next = fetch(t + 1u);
        coop4_add(acc, x, sh);
        x = x_next;
    }
#pragma unroll 1
    for (uint32_t d = G >> 1; d >= 1u; d >>= 1) {
        XYZZ<F> t = coop_shfl_xor(acc, (int)d);
        coop4_add(acc, t, sh);
    }
    if (threadIdx.x < 32u && gl == 0u && live) store_vec(vec_all + (size_t)blockIdx.y * (R + C) + w, acc);
}

// grid (2, batch): blockIdx.x = 0 -> X = C * sum_r r * Row_r, 1 -> Y = sum_q (q + 1) * Col_q.
// Logical lane l holds the m = n / 32 consecutive entries E[l*m .. l*m+m): a running sum gives S_l = sum E and
// T_l = sum_i (i + 1) E[l*m + i]; then sum_j (j + w0) E_j = m * sum_{l >= 1} suffix_l(S) + sum_l (T_l - (1 - w0) S_l):
// one suffix scan and one sum over the lanes (shuffles), log2(m) doublings -- no scalar multiplication, no digit sums.
template <class F>
__global__ void __launch_bounds__(COOP_THREADS)
k_weighted_coop(const XYZZ<F>* __restrict__ vec_all, int log_k, int lc, XYZZ<F>* __restrict__ xy_all) {
    __shared__ CoopBuf sh;
    const uint32_t K = 1u << log_k, C = 1u << lc, R = K >> lc;
    const int which = blockIdx.x;
    const uint32_t n = which == 0 ? R : C;
    const XYZZ<F>* E = vec_all + (size_t)blockIdx.y * (R + C) + (which == 0 ? 0 : R);
    const int lane = threadIdx.x & 31;
    const uint32_t m = n >= 32u ? (n >> 5) : 1u;
    int log_m = 0;
    while ((1u << log_m) < m) log_m++;
    XYZZ<F> S = XYZZ<F>::identity(), T = XYZZ<F>::identity();
#pragma unroll 1
    for (int i = (int)m - 1; i >= 0; i--) {
        const uint32_t j = (uint32_t)lane * m + (uint32_t)i;
        XYZZ<F> x = XYZZ<F>::identity();
        if (j < n) x = load_vec(E + j);
        if (i == (int)m - 1) {
            S = x;
            T = x;
        } else {
            coop4_add(S, x, sh);
            coop4_add(T, S, sh);
        }
    }
    XYZZ<F> V = T;
    if (which == 0) {   // weights start at 0: T - S
        XYZZ<F> ns = S;
        ns.y = neg(ns.y);
        coop4_add(V, ns, sh);
    }
    XYZZ<F> suf = S;    // inclusive suffix sums over the lanes
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        XYZZ<F> t = coop_shfl_down(suf, d);
        coop4_add(suf, t, sh);
    }
    XYZZ<F> A = lane >= 1 ? suf : XYZZ<F>::identity();
    for (int k = 0; k < log_m; k++) coop4_double(A, sh);
    coop4_add(V, A, sh);
    V = coop_lane_sum(V, sh);
    if (which == 0)
        for (int k = 0; k < lc; k++) coop4_double(V, sh);   // the row part carries the factor C = 2^lc
    if (threadIdx.x == 0) store_vec(xy_all + (size_t)blockIdx.y * 2 + which, V);
}

// root = C*X + Y (both already weighted) ; out = affine(root)
template <class F>
__global__ void k_reduce_final(const XYZZ<F>* __restrict__ xy_all, uint32_t batch, Affine<F>* out_xy, XYZZ<F>* out_xyzz) {
    const uint32_t b = blockIdx.x;
    if (b >= batch) return;
    XYZZ<F> acc = load_vec(xy_all + (size_t)b * 2);
    XYZZ<F> y = load_vec(xy_all + (size_t)b * 2 + 1);
    quad_add(acc, y);
    if (threadIdx.x == 0) {
        if (out_xyzz) store_vec(out_xyzz + b, acc);
        if (out_xy) store_vec(out_xy + b, xyzz_to_affine<false>(acc));
    }
}

template <class F>
__global__ void k_identity_out(uint32_t batch, Affine<F>* out_xy, XYZZ<F>* out_xyzz) {
    const uint32_t b = blockIdx.x;
    if (threadIdx.x != 0 || b >= batch) return;
    if (out_xyzz) store_vec(out_xyzz + b, XYZZ<F>::identity());
    if (out_xy) {
        Affine<F> z;
        z.x = F::zero();
        z.y = F::zero();
        store_vec(out_xy + b, z);
    }
}

// out[b] = affine( sum_{g < count} parts[g * stride + b] ) : one 4-lane group per commitment (multi-GPU gather)
template <class F>
__global__ void k_combine(const XYZZ<F>* __restrict__ parts, int count, size_t stride, Affine<F>* out_xy) {
    const uint32_t b = blockIdx.x;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (int i = 0; i < count; i++) {
        XYZZ<F> q = load_vec(parts + (size_t)i * stride + b);
        quad_add(acc, q);
    }
    if (threadIdx.x == 0) store_vec(out_xy + b, xyzz_to_affine<false>(acc));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
}  // namespace sb

struct sb_ck_table {
    int c, W;
    uint32_t K;
    void* table;  // Affine[W][n]
};
struct sb_ck {
    int curve;
    size_t n;
    int c, W;      // primary table (also tables[0])
    uint32_t K;
    void* table;
    std::vector<sb_ck_table> tables;  // every registered window width; the commit picks the cheapest per call
    // single process, several GPUs (sb_init_devices): the key is split block-cyclically (SHARD_BLOCK points per block) over
    // the devices; shards[g] is an ordinary key on device g holding blocks g, g + G, g + 2G, ...  (empty = not sharded)
    std::vector<sb_ck*> shards;
    int dev_index = 0;   // index into Runtime::devs of the device holding `tables`
};

namespace sb {


// Window width per key size, measured on B200 (profiles/r1_window_tuning.txt).  Only widths at which the window
// count W = floor(254/c)+1 drops are worth having (c = 13, 15, 16, 17, 19, 20): a wider window with the same W only
// doubles the buckets.
static int default_window_bits(size_t n) {
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    if (lg <= 12) return 10;
    if (lg <= 14) return 11;
    if (lg <= 16) return 13;
    if (lg <= 17) return 15;
    if (lg <= 19) return 16;
    if (lg <= 22) return 17;
    if (lg <= 23) return 19;
    return 20;
}

struct MsmPlan {
    int c, W;            // window width / count of the table chosen for this call
    const void* table;
    size_t n, total, nW, chunks;
    uint32_t batch, K, KB, tiles;
    uint32_t chunk_len;  // sorted entries per thread of k_accumulate
    uint32_t fix_seq;    // buckets split over more chunks than this leave the lane-per-bucket fix-up for the cooperative tree
    uint32_t parts;      // level-1 partitions of the two-level sort (0 = per-entry atomic counting sort)
    int rounds;          // batched-affine reduction rounds before the XYZZ chunk kernel (0 = none), affine.cuh
    int pair_b;          // outputs per thread in k_pair_round (8 or 16)
    size_t m_final;      // upper bound of the entries left after the rounds
    size_t off_dig, off_counts, off_offsets, off_tiles, off_ekey, off_eidx, off_buckets, off_ph, off_pt,
        off_heavy, off_nodes_a, off_nodes_b, off_digits, off_out_xy, off_out_xyzz, off_scalars, off_round_a, off_round_b, off_parts, off_cursor, off_out1, total_bytes;
};

// Tuning knobs (sb_msm_tune): number of batched-affine rounds (-1 = automatic) and outputs per thread.
static int g_affine_rounds = []() {
    const char* e = getenv("SB_MSM_AFFINE_ROUNDS");
    return e ? atoi(e) : -1;
}();
static int g_sort_mode = []() {   // 0 = automatic, 1 = always the per-entry atomic counting sort, 2 = partition sort when legal
    const char* e = getenv("SB_MSM_SORT");
    return e ? atoi(e) : 0;
}();
static int g_pair_b = []() {
    const char* e = getenv("SB_MSM_PAIR_B");
    return e ? atoi(e) : 16;
}();
static int g_tail_mode = []() {   // 1 = 4-warp cooperative tail kernels (coop.cuh), 0 = the round-1 single-lane / quad-lane tail
    const char* e = getenv("SB_MSM_TAIL");
    return e ? atoi(e) : 1;
}();
constexpr int MAX_AFFINE_ROUNDS = 8;

// Automatic choice (-1): no affine rounds -- on B200 a round is memory-latency bound and loses to the XYZZ chunk kernel at
// every measured shape (profiles/r1_affine_sweep_bn256.txt); the rounds run only when asked for.
static int auto_affine_rounds() { return 0; }

static int make_plan(const sb_ck* ck, size_t n, size_t batch, bool stage_scalars, MsmPlan& p) {
    p = MsmPlan{};
    p.n = n;
    p.batch = (uint32_t)batch;
    p.total = n * batch;
    {   // cheapest registered window width: ~W mixed additions per scalar + ~12 addition-equivalents per bucket
        // (fix-up, row/column sums and their latency), fitted to profiles/r1_window_tuning.txt
        const sb_ck_table* best = nullptr;
        double best_cost = 0;
        for (const sb_ck_table& t : ck->tables) {
            double cost = (double)p.total * t.W + 12.0 * (double)batch * t.K;
            if (!best || cost < best_cost) { best = &t; best_cost = cost; }
        }
        p.c = best->c;
        p.W = best->W;
        p.K = best->K;
        p.table = best->table;
    }
    p.nW = p.total * (size_t)p.W;
    const uint64_t kb = (uint64_t)p.K * batch;
    if (p.nW >= (1ull << 31) || kb > (1ull << 25)) {
        set_error("sb_msm: batch too large (n*W*batch = %zu, buckets = %llu)", p.nW, (unsigned long long)kb);
        return SB_ERR_ARG;
    }
    p.KB = (uint32_t)kb;
    if ((kb + SCAN_TILE - 1) / SCAN_TILE > SCAN_MAX_TILES) {
        set_error("sb_msm: too many buckets (%llu)", (unsigned long long)kb);
        return SB_ERR_ARG;
    }
    {   // two-level partition sort for big inputs (the atomic path stays for tiny / huge bucket counts)
        const uint64_t P = (kb + PART_BUCKETS - 1) >> PART_LO_BITS;
        const bool legal = kb >= 2 * PART_BUCKETS && P <= PART_MAX;
        // automatic: only up to 256 partitions (<= 65 536 buckets).  With more, a block's 256*W entries spread over so
        // many partitions that its runs shrink to a few entries (scattered stores again): at 2^23 scalars, c = 19,
        // 1024 partitions the commit takes 34.4 ms against 23.0 ms with the counting sort (profiles/r1_large_msm_check.txt)
        const bool want = g_sort_mode == 2 || (g_sort_mode == 0 && p.nW >= ((size_t)1 << 18) && P <= 256);
        const size_t part_smem = ((size_t)3 * P + 256 + 2) * 4 + (size_t)256 * p.W * 8;   // k_partition's staging area
        p.parts = (legal && want && part_smem <= 200 * 1024) ? (uint32_t)P : 0u;
    }
    p.rounds = g_affine_rounds >= 0 ? std::min(g_affine_rounds, MAX_AFFINE_ROUNDS) : auto_affine_rounds();
    p.pair_b = g_pair_b == 8 ? 8 : 16;
    // entries left for the chunk kernel: every round halves each bucket, rounding up
    p.m_final = p.nW;
    for (int r = 0; r < p.rounds; r++) p.m_final = p.m_final / 2 + p.KB;
    // chunk length (one thread per chunk): as long as possible -- fewer bucket pieces for k_fixup -- while keeping
    // >= 4 resident warps per scheduler (148 SMs x 4 SMSPs x 4 warps x 32 lanes = 75 776 threads) in k_accumulate
    // and no longer than ~1/4 of the average bucket (measured: chunks spanning several buckets run ~30 % slower)
    {
        // Small commits (the row shards of a multi-GPU step, KB <= 32768 buckets) are bound by the LATENCY of the per-commitment
        // tail, not by throughput: two resident blocks per SM already keep the integer pipe ~full (a lone warp reaches 70 % of the
        // mixed-addition peak), and chunks twice as long halve the pieces the fix-up has to add serially -- measured on one rank's
        // share of the 8-GPU step: 3.74 -> 3.56 ms (fix-up 0.50 -> 0.34 ms).  Large commits keep four blocks per SM and chunks
        // of at most a quarter bucket (longer chunks cost the accumulate kernel what the fix-up gains, profiles/r2_chunk_policy.txt).
        static const int wave_blocks_env = []() {
            const char* e = getenv("SB_MSM_WAVE_BLOCKS");
            const int v = e ? atoi(e) : 0;
            return v >= 1 && v <= 4 ? v : 0;
        }();
        static const int bucket_div_env = []() {
            const char* e = getenv("SB_MSM_BUCKET_DIV");
            const int v = e ? atoi(e) : 0;
            return v >= 1 && v <= 8 ? v : 0;
        }();
        const bool small = p.KB <= 32768u;
        const int wave_blocks = wave_blocks_env ? wave_blocks_env : (small ? 2 : 4);
        const int bucket_div = bucket_div_env ? bucket_div_env : (small ? 2 : 4);
        const size_t wave_threads = (size_t)(runtime().sm_count > 0 ? runtime().sm_count : 148) * wave_blocks * 128;  // one full wave of k_accumulate
        const size_t m = p.rounds ? (p.nW >> p.rounds) : p.nW;
        const size_t per_bucket = m / (p.KB ? p.KB : 1);
        int l = LS_MIN_LOG;
        while (l < LS_MAX_LOG && (m >> (l + 1)) >= wave_threads && ((size_t)bucket_div << l) < per_bucket) l++;
        size_t len = (size_t)1 << l;
        // shorten the chunks so that the threads fill a whole number of waves (every thread does the same work: a
        // partly filled last wave costs a full wave's latency on the SMs it touches)
        if (m >= wave_threads * 16) {
            const size_t waves = (m + len * wave_threads - 1) / (len * wave_threads);
            len = (m + waves * wave_threads - 1) / (waves * wave_threads);
            if (len < 16) len = 16;
        }
        p.chunk_len = (uint32_t)len;
        // The tree kernel is for the FEW long buckets (a cooperative addition costs 2.6x the issue slots of a plain one): the
        // threshold sits well above the pieces an average bucket has, or a tenth of all buckets ends up there (k = 20, batched
        // cross-term commits: 240 entries per bucket in chunks of 64 -- fix-up 1.8 -> 2.7 ms per step with a fixed threshold of 6)
        const size_t typical = per_bucket / len + 1;
        p.fix_seq = (uint32_t)std::min<size_t>(32, std::max<size_t>(FIX_SEQ, 2 * typical + 2));
        // With very many buckets the lane-per-bucket kernel is throughput-bound for longer than its longest serial run lasts
        // (k = 20, 2^19 buckets: 0.44 ms against 27 pieces x 6.8 us), so the long runs are free there and the tree only adds its
        // own time (measured: +0.47 ms per 12.6 M-point commit, profiles/r2g_k20_launches.csv): keep it for the small commits
        if (p.KB > 65536u) p.fix_seq = 32;
    }
    p.chunks = (p.m_final + p.chunk_len - 1) / p.chunk_len;
    p.tiles = (p.KB + SCAN_TILE - 1) / SCAN_TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    p.off_dig = take(p.nW * (p.parts ? 4 : 8));   // digits (partition sort) or (digit, rank) pairs
    p.off_parts = take(p.parts ? ((size_t)p.parts * 3 + 2) * 4 : 0);   // partition counts | cursors | offsets (+1)
    p.off_cursor = take(p.parts ? (size_t)p.KB * 4 : 0);
    p.off_out1 = take(p.parts ? p.nW * 8 : 0);
    p.off_counts = take(((size_t)p.KB + 2) * 4);  // +2: the heavy / tree bucket counters live behind the counts (one memset)
    p.off_offsets = take(((size_t)p.KB + 1) * 4 * (size_t)(p.rounds + 1));  // one row per round
    p.off_tiles = take((size_t)SCAN_MAX_TILES * 4 * (size_t)(p.rounds + 1));
    p.off_ekey = take((p.chunks + 1) * 4);  // chunk heads
    p.off_eidx = take(p.nW * 4);
    p.off_buckets = take((size_t)p.KB * 128);
    p.off_ph = take(p.chunks * 128);
    p.off_pt = take(p.chunks * 128);
    p.off_heavy = take((p.chunks / 2 + 4) * 4 * 2);   // heavy list | tree list (a listed bucket spans > FIX_SEQ - 2 >= 2 whole chunks)
    {
        const int log_k = p.c - 1, lc = (log_k + 1) / 2;
        p.off_nodes_a = take((((size_t)1 << lc) + ((size_t)p.K >> lc)) * batch * 128);  // row + column sums
        p.off_nodes_b = take((size_t)batch * 2 * 128);                                   // X, Y
        p.off_digits = take((size_t)batch * 2 * 8 * 8 * 128);                            // octal digit sums
    }
    p.off_out_xy = take(64 * batch);
    p.off_out_xyzz = take(128 * batch);
    p.off_scalars = take(stage_scalars ? p.total * 32 : 0);
    // round outputs ping-pong: A_1, A_3, .. in `a` (<= nW/2 + KB points), A_2, A_4, .. in `b` (<= nW/4 + KB)
    p.off_round_a = take(p.rounds >= 1 ? (p.nW / 2 + p.KB + 1) * 64 : 0);
    p.off_round_b = take(p.rounds >= 2 ? (p.nW / 4 + p.KB + 1) * 64 : 0);
    p.total_bytes = off;
    return SB_OK;
}

int comm_exchange_enqueue(::sb_comm* c, int curve, const void* d_in, int pairs, size_t batch, void* d_out_xy, cudaStream_t st);   // comm.cu

static size_t g_part_smem_set[64] = {0};   // per device: largest dynamic shared-memory size k_partition has been opted into

template <class F, class S>
static int msm_enqueue(const sb_ck* ck, const MsmPlan& p, char* ws, const void* d_scalars, size_t stride, void* d_out_xy,
                       void* d_out_xyzz, cudaStream_t st, ::sb_comm* comm = nullptr) {
    const uint32_t n = (uint32_t)p.n, total = (uint32_t)p.total;
    auto* dig = (uint2*)(ws + p.off_dig);
    auto* counts = (uint32_t*)(ws + p.off_counts);
    auto* offsets = (uint32_t*)(ws + p.off_offsets);
    auto* tiles = (uint32_t*)(ws + p.off_tiles);
    auto* chunk_head = (uint32_t*)(ws + p.off_ekey);
    auto* eidx = (uint32_t*)(ws + p.off_eidx);
    auto* buckets = (XYZZ<F>*)(ws + p.off_buckets);
    auto* PH = (XYZZ<F>*)(ws + p.off_ph);
    auto* PT = (XYZZ<F>*)(ws + p.off_pt);
    auto* heavy_list = (uint32_t*)(ws + p.off_heavy);
    uint32_t* tree_list = heavy_list + (p.chunks / 2 + 4);
    const uint32_t K = p.K, KB = p.KB;
    uint32_t* heavy_count = counts + KB;

    const uint32_t P = p.parts;
    auto* part_count = (uint32_t*)(ws + p.off_parts);   // [P] counts, [P] cursors, [P + 1] offsets
    uint32_t* part_cursor = part_count + P;
    uint32_t* part_off = part_count + 2 * (size_t)P;
    auto* bucket_cursor = (uint32_t*)(ws + p.off_cursor);
    auto* out1 = (uint2*)(ws + p.off_out1);
    const unsigned part_blocks = (unsigned)((p.total + 256 * PART_SPT - 1) / (256 * PART_SPT));
    {
        ProfScope ps(st, PROF_DECOMPOSE, p.total);
        SB_CUDA_TRY(cudaMemsetAsync(counts, 0, ((size_t)KB + 2) * 4, st));
        if (P) {
            SB_CUDA_TRY(cudaMemsetAsync(part_count, 0, (size_t)P * 2 * 4, st));
            SB_CUDA_TRY(cudaMemsetAsync(bucket_cursor, 0, (size_t)KB * 4, st));
            k_digits_parts<S><<<part_blocks, 256, P * 4, st>>>((const S*)d_scalars, n, total, stride, K, p.c, p.W, P, (uint32_t*)dig, part_count);
        } else {
            k_decompose<S><<<(total + 255) / 256, 256, 0, st>>>((const S*)d_scalars, n, total, stride, K, p.c, p.W, dig, counts);
        }
        SB_KERNEL_CHECK();
    }
    std::unique_ptr<ProfScope> sort_scope(new ProfScope(st, PROF_SORT, p.nW));
    const int R = p.rounds;
    unsigned part_split = 1;
    if (P) {   // level 1 partition, then the per-bucket counts of level 2
        k_scan_small<<<1, 1024, 0, st>>>(part_count, P, part_off);
        SB_KERNEL_CHECK();
        const size_t part_smem = ((size_t)3 * P + 256 + 2) * 4 + (size_t)256 * p.W * 8;
        int cur_dev = 0;
        cudaGetDevice(&cur_dev);
        size_t& smem_set = g_part_smem_set[cur_dev & 63];
        if (part_smem > smem_set) {   // ONE tracker per device for the one (non-template) kernel: the attribute only ever grows
            SB_CUDA_TRY(cudaFuncSetAttribute(k_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
            smem_set = part_smem;
        }
        k_partition<<<(total + 255) / 256, 256, part_smem, st>>>((const uint32_t*)dig, n, total, K, (uint32_t)ck->n, p.W, P, part_off, part_cursor, out1);
        SB_KERNEL_CHECK();
        part_split = 2048u / P;
        if (part_split < 1) part_split = 1;
        if (part_split > 64) part_split = 64;
        k_bucket_hist<<<dim3(P, part_split), 256, 0, st>>>(out1, part_off, KB, counts);
        SB_KERNEL_CHECK();
    }
    {
        dim3 gs(p.tiles, R + 1);
        k_scan_tile_sums<<<gs, SCAN_THREADS, 0, st>>>(counts, KB, tiles);
        SB_KERNEL_CHECK();
        k_scan_tiles<<<R + 1, 1024, 0, st>>>(tiles, p.tiles);
        SB_KERNEL_CHECK();
        k_scan_apply<<<gs, SCAN_THREADS, 0, st>>>(counts, KB, tiles, offsets);
        SB_KERNEL_CHECK();
    }
    if (!P) {
        k_scatter<<<(total + 255) / 256, 256, 0, st>>>(dig, n, total, K, (uint32_t)ck->n, p.W, offsets, eidx);
        SB_KERNEL_CHECK();
    } else {
        k_place<<<dim3(P, part_split), 256, 0, st>>>(out1, part_off, offsets, bucket_cursor, KB, eidx);
        SB_KERNEL_CHECK();
    }
    const uint32_t* off_final = offsets + (size_t)R * ((size_t)KB + 1);   // offsets of the entries the chunk kernel sees
    k_chunk_heads<<<(KB + 255) / 256, 256, 0, st>>>(off_final, KB, p.chunk_len, chunk_head);
    sort_scope.reset();
    SB_KERNEL_CHECK();
    const Affine<F>* acc_src = (const Affine<F>*)p.table;
    {
        ProfScope ps(st, PROF_ACCUMULATE, p.nW);  // units: mixed additions (upper bound: zero digits are skipped)
        // batched-affine rounds: A_0 = table[eidx] -> A_1 (buffer a) -> A_2 (buffer b) -> A_3 (a) ..
        auto* buf_a = (Affine<F>*)(ws + p.off_round_a);
        auto* buf_b = (Affine<F>*)(ws + p.off_round_b);
        size_t bound = p.nW;
        for (int r = 0; r < R; r++) {
            bound = bound / 2 + KB;                      // upper bound of this round's outputs
            const uint32_t* off_in = offsets + (size_t)r * ((size_t)KB + 1);
            const uint32_t* off_out = offsets + (size_t)(r + 1) * ((size_t)KB + 1);
            Affine<F>* dst = (r & 1) ? buf_b : buf_a;
            const size_t per_block = (size_t)PR_THREADS * p.pair_b;
            const unsigned blocks = (unsigned)((bound + per_block - 1) / per_block);
            if (r == 0) {
                if (p.pair_b == 8) k_pair_round<F, true, 8><<<blocks, PR_THREADS, 0, st>>>((const Affine<F>*)p.table, eidx, off_in, off_out, KB, dst);
                else k_pair_round<F, true, 16><<<blocks, PR_THREADS, 0, st>>>((const Affine<F>*)p.table, eidx, off_in, off_out, KB, dst);
            } else {
                if (p.pair_b == 8) k_pair_round<F, false, 8><<<blocks, PR_THREADS, 0, st>>>(acc_src, nullptr, off_in, off_out, KB, dst);
                else k_pair_round<F, false, 16><<<blocks, PR_THREADS, 0, st>>>(acc_src, nullptr, off_in, off_out, KB, dst);
            }
            SB_KERNEL_CHECK();
            acc_src = dst;
        }
        size_t blocks = (p.chunks + 127) / 128;
        static const int minb = []() {
            const char* e = getenv("SB_ACC_MINB");
            return e ? atoi(e) : 4;
        }();
        if (R > 0)
            k_accumulate<F, 4, true><<<(unsigned)blocks, 128, 0, st>>>(acc_src, chunk_head, nullptr, off_final, KB, p.chunk_len, buckets, PH, PT);
        else if (minb == 5)
            k_accumulate<F, 5, false><<<(unsigned)blocks, 128, 0, st>>>((const Affine<F>*)p.table, chunk_head, eidx, offsets, KB, p.chunk_len, buckets, PH, PT);
        else if (minb == 6)
            k_accumulate<F, 6, false><<<(unsigned)blocks, 128, 0, st>>>((const Affine<F>*)p.table, chunk_head, eidx, offsets, KB, p.chunk_len, buckets, PH, PT);
        else if (minb == 3)
            k_accumulate<F, 3, false><<<(unsigned)blocks, 128, 0, st>>>((const Affine<F>*)p.table, chunk_head, eidx, offsets, KB, p.chunk_len, buckets, PH, PT);
        else
            k_accumulate<F, 4, false><<<(unsigned)blocks, 128, 0, st>>>((const Affine<F>*)p.table, chunk_head, eidx, offsets, KB, p.chunk_len, buckets, PH, PT);
        SB_KERNEL_CHECK();
    }
    {
        ProfScope ps(st, PROF_FIXUP, KB);
        // one plain lane per bucket for the short runs (<= FIX_SEQ pieces: throughput-bound, and a cooperative addition
        // costs an SM 2.6x the issue slots of a plain one), the cooperative tree for the long ones.  SB_MSM_TAIL=2 keeps
        // the all-cooperative fix-up of small commits for A/B runs.
        if (g_tail_mode == 2 && KB <= 32768u) k_fixup_coop<F><<<(KB + 31) / 32, COOP_THREADS, 0, st>>>(off_final, KB, p.chunk_len, buckets, PH, PT, heavy_count, heavy_list, tree_list, p.fix_seq);
        else k_fixup<F><<<(KB + 127) / 128, 128, 0, st>>>(off_final, KB, p.chunk_len, buckets, PH, PT, heavy_count, heavy_list, tree_list, p.fix_seq);
        SB_KERNEL_CHECK();
        k_fixup_tree<F><<<592, COOP_THREADS, 0, st>>>(off_final, p.chunk_len, buckets, PH, PT, heavy_count, tree_list);
        SB_KERNEL_CHECK();
        k_fixup_heavy<F><<<296, HEAVY_THREADS, 0, st>>>(off_final, p.chunk_len, buckets, PH, PT, heavy_count, heavy_list);
        SB_KERNEL_CHECK();
    }
    {
        const int log_k = p.c - 1;
        const int lc = (log_k + 1) / 2;
        const uint32_t C = 1u << lc, R = K >> lc;
        auto* vec = (XYZZ<F>*)(ws + p.off_nodes_a);   // [batch][R + C]
        auto* xy = (XYZZ<F>*)(ws + p.off_nodes_b);    // [batch][2]
        {
            ProfScope ps(st, PROF_REDUCE, KB);
            if (g_tail_mode) {
                // group width, from a model fitted to the launches of profiles/r2_shard8_launches_tree.csv and r2f_ncu_full_summary.txt:
                // a dependent cooperative addition of this kernel takes ~4.5 us (it waits for a 128-byte bucket load), an SM sustains
                // ~0.4 of them per us with 3 resident blocks, so time ~ ops x max(4.5, blocks per SM / 0.4).  Narrower groups pay
                // when there are thousands of SHORT rows (batched cross-term commits of a multi-GPU shard: 768 rows of 64: 75 -> 55 us);
                // a narrower group must win by 10 % to be taken (one row per block is the form measured longest)
                int log_g = 5;
                double best_t = 0;
                const double sms = (double)(runtime().sm_count > 0 ? runtime().sm_count : 148);
                for (int lg = 5; lg >= 3; lg--) {
                    const uint32_t G = 1u << lg, cmax = R > C ? R : C;
                    const double ops = (double)((cmax + G - 1) / G - 1) + lg;
                    const double blocks = (double)((R + C + (32u >> lg) - 1) / (32u >> lg)) * p.batch;
                    const double t = ops * std::max(4.5, blocks / sms / 0.4);
                    if (lg == 5 || t < 0.9 * best_t) { best_t = t; log_g = lg; }
                }
                k_rowcol_coop<F><<<dim3((R + C + (32u >> log_g) - 1) / (32u >> log_g), p.batch), COOP_THREADS, 0, st>>>(buckets, log_k, lc, log_g, vec);
                SB_KERNEL_CHECK();
                k_weighted_coop<F><<<dim3(2, p.batch), COOP_THREADS, 0, st>>>(vec, log_k, lc, xy);
                SB_KERNEL_CHECK();
            } else {
            dim3 g1((R + C + 3) / 4, p.batch);
            k_rowcol_sums<F><<<g1, 128, 0, st>>>(buckets, log_k, lc, vec);
            SB_KERNEL_CHECK();
            const int max_log = (log_k - lc) > lc ? (log_k - lc) : lc;
            const int max_pos = (max_log + 2) / 3 > 0 ? (max_log + 2) / 3 : 1;   // <= 8: log_k <= 24
            auto* dsum = (XYZZ<F>*)(ws + p.off_digits);   // [batch][2][max_pos][8]
            dim3 gd(2 * max_pos * 8, p.batch);
            k_digit_sums<F><<<gd, 32, 0, st>>>(vec, log_k, lc, max_pos, dsum);
            SB_KERNEL_CHECK();
            dim3 g2(2, p.batch);
            k_weighted_finish<F><<<g2, 32 * (max_pos + 1), 0, st>>>(dsum, log_k, lc, max_pos, xy);
            SB_KERNEL_CHECK();
            }
        }
        {
            ProfScope pf(st, PROF_FINALIZE, p.batch);
            if (comm) {   // multi-GPU: X + Y, the exchange over peer memory, the sum of the ranks' partials and the normalisation in ONE kernel
                SB_TRY(comm_exchange_enqueue(comm, ck->curve, xy, 1, p.batch, d_out_xy, st));
            } else {
                k_reduce_final<F><<<p.batch, 4, 0, st>>>(xy, p.batch, (Affine<F>*)d_out_xy, (XYZZ<F>*)d_out_xyzz);
                SB_KERNEL_CHECK();
            }
        }
    }
    return SB_OK;
}

static int msm_dispatch(const sb_ck* ck, const MsmPlan& p, char* ws, const void* d_scalars, size_t stride, void* d_out_xy,
                        void* d_out_xyzz, cudaStream_t st, ::sb_comm* comm = nullptr) {
    if (p.batch == 0) return SB_OK;
    if (p.n == 0 && comm) {   // this rank owns no rows of the vector: its partial is the identity, the peers still wait for it
        void* xy = ws + p.off_nodes_b;
        if (ck->curve == CURVE_BN256) k_identity_out<Fq><<<p.batch * 2, 32, 0, st>>>(p.batch * 2, nullptr, (XYZZ<Fq>*)xy);
        else k_identity_out<Fr><<<p.batch * 2, 32, 0, st>>>(p.batch * 2, nullptr, (XYZZ<Fr>*)xy);
        SB_KERNEL_CHECK();
        return comm_exchange_enqueue(comm, ck->curve, xy, 1, p.batch, d_out_xy, st);
    }
    if (p.n == 0) {
        if (ck->curve == CURVE_BN256) k_identity_out<Fq><<<p.batch, 32, 0, st>>>(p.batch, (Affine<Fq>*)d_out_xy, (XYZZ<Fq>*)d_out_xyzz);
        else k_identity_out<Fr><<<p.batch, 32, 0, st>>>(p.batch, (Affine<Fr>*)d_out_xy, (XYZZ<Fr>*)d_out_xyzz);
        SB_KERNEL_CHECK();
        return SB_OK;
    }
    if (ck->curve == CURVE_BN256) return msm_enqueue<Fq, Fr>(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
    return msm_enqueue<Fr, Fq>(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
}

// ------------------------------------------------------------------------------------------------
// CUDA-graph cache of whole commitment pipelines.
// A prover calls the same commitment again and again: same key, same device buffers, same stream (the device-resident
// session of sirius_b200/device.py commits W_in and T every fold step).  The pipeline is ~20 launches and memsets whose
// arguments depend only on those inputs and on the stream's scratch address, so the second call with an identical
// signature is captured into a graph and every later one is ONE cudaGraphLaunch: the host spends ~10 us instead of
// ~100 us per pipeline (it sits on the critical path of the short phases of a multi-GPU shard), and the kernels of the
// sort chain follow each other without launch gaps.  Eager launches stay for first calls, for profiling runs (event
// records between the kernels) and with SB_MSM_GRAPH=0.  Entries die with their key, stream or communicator.
// ------------------------------------------------------------------------------------------------
struct GraphKey {
    const sb_ck* ck;
    const void* table;
    size_t n, stride, batch, ws_bytes;
    const void* in;
    void *out_xy, *out_xyzz;
    cudaStream_t st;
    ::sb_comm* comm;
    char* ws;
    uint64_t epoch;
    bool operator==(const GraphKey& o) const {
        return ck == o.ck && table == o.table && n == o.n && stride == o.stride && batch == o.batch && ws_bytes == o.ws_bytes && in == o.in &&
               out_xy == o.out_xy && out_xyzz == o.out_xyzz && st == o.st && comm == o.comm && ws == o.ws && epoch == o.epoch;
    }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;   // nullptr: seen once, not captured yet
    bool refused;           // capture failed once: stay eager
    uint64_t launches, last_use;
};
static std::vector<GraphEntry> g_graphs;   // under runtime().mu
static uint64_t g_graph_epoch = 1, g_graph_clock = 0;
static int g_graph_on = []() {
    const char* e = getenv("SB_MSM_GRAPH");
    return (!e || atoi(e) != 0) ? 1 : 0;
}();
constexpr size_t GRAPH_CACHE_MAX = 64;
uint64_t launch_count_now();   // runtime.cu
void count_launches(uint64_t n);

void msm_graphs_drop(const void* ck, cudaStream_t st, const void* comm, bool all) {   // caller holds runtime().mu
    for (size_t i = 0; i < g_graphs.size();) {
        const GraphKey& k = g_graphs[i].key;
        if (all || (ck && k.ck == ck) || (st && k.st == st) || (comm && k.comm == comm)) {
            if (g_graphs[i].exec) cudaGraphExecDestroy(g_graphs[i].exec);
            g_graphs[i] = g_graphs.back();
            g_graphs.pop_back();
        } else {
            i++;
        }
    }
}

static int msm_dispatch_cached(const sb_ck* ck, const MsmPlan& p, char* ws, const void* d_scalars, size_t stride, void* d_out_xy, void* d_out_xyzz,
                               cudaStream_t st, ::sb_comm* comm = nullptr) {
    if (!g_graph_on || profile_enabled() || p.batch == 0 || p.n == 0) return msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
    const GraphKey key{ck, p.table, p.n, stride, (size_t)p.batch, p.total_bytes, d_scalars, d_out_xy, d_out_xyzz, st, comm, ws, g_graph_epoch};
    GraphEntry* e = nullptr;
    for (GraphEntry& g : g_graphs)
        if (g.key == key) { e = &g; break; }
    if (!e) {   // first sight: run eagerly (grows the scratch, sets kernel attributes), remember the signature
        if (g_graphs.size() >= GRAPH_CACHE_MAX) {
            size_t old = 0;
            for (size_t i = 1; i < g_graphs.size(); i++)
                if (g_graphs[i].last_use < g_graphs[old].last_use) old = i;
            if (g_graphs[old].exec) cudaGraphExecDestroy(g_graphs[old].exec);
            g_graphs[old] = g_graphs.back();
            g_graphs.pop_back();
        }
        g_graphs.push_back(GraphEntry{key, nullptr, false, 0, ++g_graph_clock});
        return msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
    }
    e->last_use = ++g_graph_clock;
    if (e->refused) return msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
    if (!e->exec) {   // second call: capture
        const uint64_t l0 = launch_count_now();
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            e->refused = true;
            return msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
        }
        const int rc = msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        cudaGraphExec_t exec = nullptr;
        if (rc != SB_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            e->refused = true;
            if (rc != SB_OK) return rc;   // an argument error of the pipeline itself
            return msm_dispatch(ck, p, ws, d_scalars, stride, d_out_xy, d_out_xyzz, st, comm);
        }
        cudaGraphDestroy(graph);
        e->exec = exec;
        e->launches = launch_count_now() - l0;   // counted once during the capture: this call launches the graph once
        SB_CUDA_TRY(cudaGraphLaunch(exec, st));
        return SB_OK;
    }
    SB_CUDA_TRY(cudaGraphLaunch(e->exec, st));
    count_launches(e->launches);
    return SB_OK;
}

static int ck_add_table(sb_ck* ck, const void* d_bases, int c, cudaStream_t st) {
    if (c < 2 || c > 24) {
        set_error("sb_ck: window_bits %d out of range [2,24]", c);
        return SB_ERR_ARG;
    }
    for (const sb_ck_table& t : ck->tables)
        if (t.c == c) return SB_OK;
    const int W = 254 / c + 1;
    const size_t n = ck->n;
    if ((uint64_t)n * (uint64_t)W >= (1ull << 31)) {
        set_error("sb_ck: n*W = %zu*%d does not fit the 31-bit table index", n, W);
        return SB_ERR_ARG;
    }
    sb_ck_table t;
    t.c = c;
    t.W = W;
    t.K = 1u << (c - 1);
    t.table = nullptr;
    if (n) {
        cudaError_t e = cudaMalloc(&t.table, n * (size_t)W * 64);
        if (e != cudaSuccess) {
            set_error("sb_ck: cudaMalloc(%zu) failed: %s", n * (size_t)W * 64, cudaGetErrorString(e));
            return SB_ERR_OOM;
        }
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (ck->curve == CURVE_BN256) k_precompute<Fq><<<blocks, 128, 0, st>>>((const Affine<Fq>*)d_bases, n, c, W, (Affine<Fq>*)t.table);
        else k_precompute<Fr><<<blocks, 128, 0, st>>>((const Affine<Fr>*)d_bases, n, c, W, (Affine<Fr>*)t.table);
        cudaError_t e2 = cudaGetLastError();
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
        if (e2 != cudaSuccess) {
            set_error("sb_ck: precompute failed: %s", cudaGetErrorString(e2));
            cudaFree(t.table);
            return SB_ERR_CUDA;
        }
    }
    ck->tables.push_back(t);
    return SB_OK;
}

static int ck_build(int curve, const void* d_bases, size_t n, int window_bits, cudaStream_t st, sb_ck_t* out) {
    if (curve != CURVE_BN256 && curve != CURVE_GRUMPKIN) {
        set_error("sb_ck_register: unknown curve %d", curve);
        return SB_ERR_ARG;
    }
    int c = window_bits > 0 ? window_bits : default_window_bits(n);
    sb_ck* ck = new (std::nothrow) sb_ck();
    if (!ck) return SB_ERR_OOM;
    ck->curve = curve;
    ck->n = n;
    int rc = ck_add_table(ck, d_bases, c, st);
    if (rc != SB_OK) {
        delete ck;
        return rc;
    }
    ck->c = ck->tables[0].c;
    ck->W = ck->tables[0].W;
    ck->K = ck->tables[0].K;
    ck->table = ck->tables[0].table;
    *out = ck;
    return SB_OK;
}

// ---- single process, several GPUs ------------------------------------------------------------------------------------
constexpr size_t SHARD_BLOCK = 4096;   // points per block of the block-cyclic split: any prefix of the key is balanced to within a block

// number of the first n global indices that land on device g of G
static size_t shard_count(size_t n, int g, int G) {
    const size_t cycle = SHARD_BLOCK * (size_t)G;
    const size_t q = n / cycle, rem = n - q * cycle;
    const size_t lo = (size_t)g * SHARD_BLOCK;
    const size_t extra = rem > lo ? std::min(rem - lo, SHARD_BLOCK) : 0;
    return q * SHARD_BLOCK + extra;
}

// device g's elements (elem_bytes each) of the first n of a host array -> contiguous device memory, on `st`
static cudaError_t shard_upload(void* d_dst, const void* h_src, size_t n, size_t elem_bytes, int g, int G, cudaStream_t st) {
    const size_t cycle = SHARD_BLOCK * (size_t)G;
    const size_t q = n / cycle, rem = n - q * cycle;
    const size_t lo = (size_t)g * SHARD_BLOCK;
    const char* src = (const char*)h_src + lo * elem_bytes;
    cudaError_t e = cudaSuccess;
    if (q) e = cudaMemcpy2DAsync(d_dst, SHARD_BLOCK * elem_bytes, src, cycle * elem_bytes, SHARD_BLOCK * elem_bytes, q, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && rem > lo) {
        const size_t extra = std::min(rem - lo, SHARD_BLOCK);
        e = cudaMemcpyAsync((char*)d_dst + q * SHARD_BLOCK * elem_bytes, src + q * cycle * elem_bytes, extra * elem_bytes, cudaMemcpyHostToDevice, st);
    }
    return e;
}

static int ck_register_sharded(int curve, const uint64_t* bases_xy, size_t n, int window_bits, sb_ck_t* out) {
    Runtime& rt = runtime();
    const int G = (int)rt.devs.size();
    sb_ck* parent = new (std::nothrow) sb_ck();
    if (!parent) return SB_ERR_OOM;
    parent->curve = curve;
    parent->n = n;
    int rc = SB_OK;
    for (int g = 0; g < G && rc == SB_OK; g++) {
        const size_t ng = shard_count(n, g, G);
        cudaSetDevice(rt.devs[g].device);
        void* d_bases = nullptr;
        cudaError_t e = cudaMalloc(&d_bases, ng * 64 + 64);
        if (e == cudaSuccess) e = shard_upload(d_bases, bases_xy, n, 64, g, G, rt.devs[g].stream);
        if (e != cudaSuccess) {
            set_error("sb_ck_register (device %d): %s", rt.devs[g].device, cudaGetErrorString(e));
            rc = SB_ERR_CUDA;
        }
        sb_ck* sh = nullptr;
        if (rc == SB_OK) rc = ck_build(curve, d_bases, ng, window_bits, rt.devs[g].stream, &sh);
        if (d_bases) cudaFree(d_bases);
        if (rc == SB_OK) {
            sh->dev_index = g;
            parent->shards.push_back(sh);
        }
    }
    cudaSetDevice(rt.device);
    if (rc != SB_OK) {
        sb_ck_release(parent);
        return rc;
    }
    parent->c = parent->shards[0]->c;
    parent->W = parent->shards[0]->W;
    parent->K = parent->shards[0]->K;
    parent->table = nullptr;
    *out = parent;
    return SB_OK;
}

static Scratch g_gather;   // device 0: [G][batch] XYZZ partials + batch affine results (sharded host commits; under rt.mu)

// sb_msm_batch on a sharded key: every device commits its block-cyclic share of the scalars (uploaded straight from the
// caller's arrays by one host thread per device), the 128-byte partial sums travel to the primary device by peer copies,
// one combine kernel adds them.  Bit-identical to the single-device result (exact arithmetic).
static int msm_batch_sharded(sb_ck* ck, const uint64_t* const* scalars_mont, size_t n, size_t batch, uint64_t* out_xy) {
    Runtime& rt = runtime();
    const int G = (int)ck->shards.size();
    if (G != (int)rt.devs.size()) {
        set_error("sb_msm: the key was registered for %d devices, the runtime now drives %zu", G, rt.devs.size());
        return SB_ERR_ARG;
    }
    std::vector<MsmPlan> plans(G);
    std::vector<char*> wss(G);
    std::vector<size_t> counts(G);
    for (int g = 0; g < G; g++) {   // workspaces are reserved by the calling thread (the per-stream slot map is not thread-safe)
        counts[g] = shard_count(n, g, G);
        SB_TRY(make_plan(ck->shards[g], counts[g], batch, true, plans[g]));
        cudaSetDevice(rt.devs[g].device);
        Scratch& ws = ws_slot(rt.devs[g].stream, WS_MSM);
        int rc = ws.reserve(plans[g].total_bytes);
        if (rc != SB_OK) {
            cudaSetDevice(rt.device);
            return rc;
        }
        wss[g] = (char*)ws.ptr;
    }
    cudaSetDevice(rt.device);
    SB_TRY(g_gather.reserve((size_t)G * batch * 128 + batch * 64 + 256));
    char* gather = (char*)g_gather.ptr;
    std::vector<cudaEvent_t> done(G);
    std::vector<int> rcs(G, SB_OK);
    std::vector<std::string> errs(G);
    auto worker = [&](int g) {
        const Runtime::Dev& dv = rt.devs[g];
        cudaSetDevice(dv.device);
        const MsmPlan& p = plans[g];
        char* ws = wss[g];
        cudaError_t e = cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming);
        for (size_t b = 0; b < batch && e == cudaSuccess; b++)
            e = shard_upload(ws + p.off_scalars + b * counts[g] * 32, scalars_mont[b], n, 32, g, G, dv.stream);
        int rc = SB_OK;
        if (e == cudaSuccess) rc = msm_dispatch(ck->shards[g], p, ws, ws + p.off_scalars, counts[g], nullptr, ws + p.off_out_xyzz, dv.stream);
        if (rc != SB_OK) errs[g] = sb_last_error();
        if (e == cudaSuccess && rc == SB_OK)
            e = cudaMemcpyPeerAsync(gather + (size_t)g * batch * 128, rt.device, ws + p.off_out_xyzz, dv.device, batch * 128, dv.stream);
        if (e == cudaSuccess && rc == SB_OK) e = cudaEventRecord(done[g], dv.stream);
        if (e != cudaSuccess) {
            errs[g] = std::string("device ") + std::to_string(dv.device) + ": " + cudaGetErrorString(e);
            rc = SB_ERR_CUDA;
        }
        rcs[g] = rc;
    };
    {
        std::vector<std::thread> threads;
        for (int g = 1; g < G; g++) threads.emplace_back(worker, g);
        worker(0);
        for (auto& t : threads) t.join();
    }
    cudaSetDevice(rt.device);
    int rc = SB_OK;
    for (int g = 0; g < G; g++)
        if (rcs[g] != SB_OK && rc == SB_OK) {
            set_error("sb_msm (sharded): %s", errs[g].c_str());
            rc = rcs[g];
        }
    if (rc == SB_OK) {
        for (int g = 0; g < G; g++) SB_CUDA_TRY(cudaStreamWaitEvent(rt.stream, done[g], 0));
        char* d_out = gather + (size_t)G * batch * 128;
        rc = sb_msm_combine_batch_device(ck->curve, gather, G, batch, batch, d_out, rt.stream);
        if (rc == SB_OK) {
            SB_CUDA_TRY(cudaMemcpyAsync(out_xy, d_out, 64 * batch, cudaMemcpyDeviceToHost, rt.stream));
            SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
        }
    }
    for (int g = 0; g < G; g++) {   // the partial-sum buffers are reused by the next call: every device must be done with them
        cudaSetDevice(rt.devs[g].device);
        if (rcs[g] == SB_OK) cudaStreamSynchronize(rt.devs[g].stream);
        if (done[g]) cudaEventDestroy(done[g]);
    }
    cudaSetDevice(rt.device);
    return rc;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sb_ck_register(int curve, const uint64_t* bases_xy, size_t n, int window_bits, sb_ck_t* out) {
    if (!out || (!bases_xy && n)) {
        set_error("sb_ck_register: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    if (rt.devs.size() > 1 && (curve == CURVE_BN256 || curve == CURVE_GRUMPKIN)) return ck_register_sharded(curve, bases_xy, n, window_bits, out);
    void* d_bases = nullptr;
    if (n) {
        SB_CUDA_TRY(cudaMalloc(&d_bases, n * 64));
        cudaError_t e = cudaMemcpyAsync(d_bases, bases_xy, n * 64, cudaMemcpyHostToDevice, rt.stream);
        if (e != cudaSuccess) {
            cudaFree(d_bases);
            set_error("sb_ck_register: H2D failed: %s", cudaGetErrorString(e));
            return SB_ERR_CUDA;
        }
    }
    int rc = ck_build(curve, d_bases, n, window_bits, rt.stream, out);
    if (d_bases) cudaFree(d_bases);
    return rc;
}

int sb_ck_register_device(int curve, const void* d_bases_xy, size_t n, int window_bits, void* stream, sb_ck_t* out) {
    if (!out || (!d_bases_xy && n)) {
        set_error("sb_ck_register_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return ck_build(curve, d_bases_xy, n, window_bits, stream ? (cudaStream_t)stream : rt.stream, out);
}

void sb_ck_release(sb_ck_t ck) {
    if (!ck) return;
    {
        RtLock lk(runtime().mu);
        msm_graphs_drop(ck, nullptr, nullptr, false);
    }
    for (sb_ck* sh : ck->shards) sb_ck_release(sh);
    if (!ck->tables.empty()) {
        Runtime& rt = runtime();
        const bool other = ck->dev_index > 0 && ck->dev_index < (int)rt.devs.size();
        if (other) cudaSetDevice(rt.devs[ck->dev_index].device);
        for (sb_ck_table& t : ck->tables)
            if (t.table) cudaFree(t.table);
        if (other) cudaSetDevice(rt.device);
    }
    delete ck;
}

/* A further window width for the same key: table 0 holds the generators themselves (window 0 of any table), so
 * additional tables are derived on the device.  Commits then use whichever registered width is cheapest. */
int sb_ck_add_window(sb_ck_t ck, int window_bits, void* stream) {
    if (ck && !ck->shards.empty()) {   // sharded key: the width is added on every device (its own stream)
        SB_TRY(ensure_runtime());
        Runtime& rt = runtime();
        RtLock lk(rt.mu);
        int rc = SB_OK;
        for (size_t g = 0; g < ck->shards.size() && rc == SB_OK; g++) {
            cudaSetDevice(rt.devs[g].device);
            rc = ck_add_table(ck->shards[g], ck->shards[g]->tables[0].table, window_bits, rt.devs[g].stream);
        }
        cudaSetDevice(rt.device);
        return rc;
    }
    if (!ck || ck->tables.empty()) {
        set_error("sb_ck_add_window: bad key");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    return ck_add_table(ck, ck->tables[0].table, window_bits, stream ? (cudaStream_t)stream : rt.stream);
}

size_t sb_ck_len(sb_ck_t ck) { return ck ? ck->n : 0; }

/* Tuning: key 0 = batched-affine rounds per commit (-1 automatic, 0 off, up to 8), key 1 = outputs per thread of a
 * round (8 or 16), key 2 = sort (0 automatic, 1 per-entry atomic counting sort, 2 two-level partition sort whenever the
 * bucket count allows).  Results are bit-identical for every setting. */
int sb_msm_tune(int key, int value) {
    RtLock lk(runtime().mu);   // the knobs are read by make_plan / msm_enqueue under the same lock
    g_graph_epoch++;           // captured pipelines were planned with the old setting
    if (key == 0 && value >= -1 && value <= MAX_AFFINE_ROUNDS) g_affine_rounds = value;
    else if (key == 1 && (value == 8 || value == 16)) g_pair_b = value;
    else if (key == 2 && value >= 0 && value <= 2) g_sort_mode = value;
    else if (key == 3 && value >= 0 && value <= 2) g_tail_mode = value;
    else if (key == 4 && (value == 0 || value == 1)) g_graph_on = value;
    else {
        set_error("sb_msm_tune: bad key/value %d/%d", key, value);
        return SB_ERR_ARG;
    }
    return SB_OK;
}
int sb_ck_window_bits(sb_ck_t ck) { return ck ? ck->c : 0; }

static int check_len(sb_ck_t ck, size_t n) {
    if (n > ck->n) {
        set_error("Can't commit too long input: input len: %zu, but limit is %zu", n, ck->n);
        return SB_ERR_TOO_LONG;
    }
    return SB_OK;
}

int sb_msm_batch_device(sb_ck_t ck, const void* d_scalars_mont, size_t n, size_t stride, size_t batch, void* d_out_xy,
                        void* d_out_xyzz, void* stream) {
    if (!ck || (!d_scalars_mont && n && batch) || (!d_out_xy && !d_out_xyzz) || stride < n) {
        set_error("sb_msm_batch_device: bad argument");
        return SB_ERR_ARG;
    }
    if (!ck->shards.empty()) {
        set_error("sb_msm_batch_device: this key is sharded over %zu devices (sb_init_devices); use the host entry points sb_msm / sb_msm_batch", ck->shards.size());
        return SB_ERR_ARG;
    }
    SB_TRY(check_len(ck, n));
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    MsmPlan p;
    SB_TRY(make_plan(ck, n, batch, false, p));
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    Scratch& ws = ws_slot(st, WS_MSM);
    SB_TRY(ws.reserve(p.total_bytes));
    return msm_dispatch_cached(ck, p, (char*)ws.ptr, d_scalars_mont, stride, d_out_xy, d_out_xyzz, st);
}

/* Row-sharded commitment group: this rank's scalars against this rank's slice of the key; the partial sums are
 * exchanged over peer memory inside the pipeline's last kernel (comm.cu) and every rank receives the affine totals. */
int sb_msm_batch_sharded_device(sb_ck_t ck, sb_comm_t comm, const void* d_scalars_mont, size_t n, size_t stride, size_t batch, void* d_out_xy,
                                void* stream) {
    if (!ck || !comm || (!d_scalars_mont && n && batch) || !d_out_xy || stride < n || !ck->shards.empty()) {
        set_error("sb_msm_batch_sharded_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(check_len(ck, n));
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    MsmPlan p;
    SB_TRY(make_plan(ck, n, batch, false, p));
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    Scratch& ws = ws_slot(st, WS_MSM);
    SB_TRY(ws.reserve(p.total_bytes));
    return msm_dispatch_cached(ck, p, (char*)ws.ptr, d_scalars_mont, stride, d_out_xy, nullptr, st, comm);
}

int sb_msm_device(sb_ck_t ck, const void* d_scalars_mont, size_t n, void* d_out_xy, void* d_out_xyzz, void* stream) {
    return sb_msm_batch_device(ck, d_scalars_mont, n, n, 1, d_out_xy, d_out_xyzz, stream);
}

int sb_msm_batch(sb_ck_t ck, const uint64_t* const* scalars_mont, size_t n, size_t batch, uint64_t* out_xy) {
    if (!ck || (!scalars_mont && batch) || !out_xy) {
        set_error("sb_msm_batch: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(check_len(ck, n));
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    RtLock lk(rt.mu);
    if (batch == 0) return SB_OK;
    for (size_t b = 0; b < batch && n; b++)
        if (!scalars_mont[b]) {
            set_error("sb_msm_batch: null scalar vector %zu", b);
            return SB_ERR_ARG;
        }
    if (!ck->shards.empty()) return msm_batch_sharded(ck, scalars_mont, n, batch, out_xy);
    MsmPlan p;
    SB_TRY(make_plan(ck, n, batch, true, p));
    Scratch& wss = ws_slot(rt.stream, WS_MSM);
    SB_TRY(wss.reserve(p.total_bytes));
    char* ws = (char*)wss.ptr;
    for (size_t b = 0; b < batch && n; b++) {
        if (!scalars_mont[b]) {
            set_error("sb_msm_batch: null scalar vector %zu", b);
            return SB_ERR_ARG;
        }
        SB_CUDA_TRY(cudaMemcpyAsync(ws + p.off_scalars + b * n * 32, scalars_mont[b], n * 32, cudaMemcpyHostToDevice, rt.stream));
    }
    SB_TRY(msm_dispatch(ck, p, ws, ws + p.off_scalars, n, ws + p.off_out_xy, nullptr, rt.stream));
    SB_CUDA_TRY(cudaMemcpyAsync(out_xy, ws + p.off_out_xy, 64 * batch, cudaMemcpyDeviceToHost, rt.stream));
    SB_CUDA_TRY(cudaStreamSynchronize(rt.stream));
    return SB_OK;
}

int sb_msm(sb_ck_t ck, const uint64_t* scalars_mont, size_t n, uint64_t out_xy[8]) {
    if (!ck || (!scalars_mont && n) || !out_xy) {
        set_error("sb_msm: null argument");
        return SB_ERR_ARG;
    }
    const uint64_t* one[1] = {scalars_mont};
    return sb_msm_batch(ck, one, n, 1, out_xy);
}

int sb_points_on_curve_device(int curve, const void* d_points_xy, size_t n, void* d_bad_count_u64, void* stream) {
    if ((!d_points_xy && n) || !d_bad_count_u64) {
        set_error("sb_points_on_curve_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    SB_CUDA_TRY(cudaMemsetAsync(d_bad_count_u64, 0, 8, st));
    if (!n) return SB_OK;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (curve == CURVE_BN256) k_on_curve<Fq><<<blocks, 256, 0, st>>>((const Affine<Fq>*)d_points_xy, n, 3, 0, (unsigned long long*)d_bad_count_u64);
    else if (curve == CURVE_GRUMPKIN) k_on_curve<Fr><<<blocks, 256, 0, st>>>((const Affine<Fr>*)d_points_xy, n, 17, 1, (unsigned long long*)d_bad_count_u64);
    else {
        set_error("sb_points_on_curve_device: unknown curve %d", curve);
        return SB_ERR_ARG;
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}

int sb_points_on_curve(int curve, const uint64_t* points_xy, size_t n, uint64_t* bad_count) {
    if ((!points_xy && n) || !bad_count) {
        set_error("sb_points_on_curve: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    void* d = nullptr;
    SB_CUDA_TRY(cudaMalloc(&d, n * 64 + 64));
    int rc = SB_OK;
    cudaError_t e = cudaMemcpyAsync((char*)d + 64, points_xy, n * 64, cudaMemcpyHostToDevice, rt.stream);
    if (e == cudaSuccess) rc = sb_points_on_curve_device(curve, (char*)d + 64, n, d, rt.stream);
    if (e == cudaSuccess && rc == SB_OK) e = cudaMemcpyAsync(bad_count, d, 8, cudaMemcpyDeviceToHost, rt.stream);
    if (e == cudaSuccess && rc == SB_OK) e = cudaStreamSynchronize(rt.stream);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("sb_points_on_curve: %s", cudaGetErrorString(e));
        return SB_ERR_CUDA;
    }
    return rc;
}

int sb_index_multiples_device(int curve, const uint64_t gen_xy[8], uint64_t first, size_t n, void* d_out_xy, void* stream) {
    if (!gen_xy || (!d_out_xy && n)) {
        set_error("sb_index_multiples_device: null argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (!n) return SB_OK;
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (curve == CURVE_BN256) {
        Affine<Fq> g;
        memcpy(&g, gen_xy, 64);
        k_index_multiples<Fq><<<blocks, 128, 0, st>>>(g, first, n, (Affine<Fq>*)d_out_xy);
    } else if (curve == CURVE_GRUMPKIN) {
        Affine<Fr> g;
        memcpy(&g, gen_xy, 64);
        k_index_multiples<Fr><<<blocks, 128, 0, st>>>(g, first, n, (Affine<Fr>*)d_out_xy);
    } else {
        set_error("sb_index_multiples_device: unknown curve %d", curve);
        return SB_ERR_ARG;
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}

int sb_msm_combine_device(int curve, const void* d_partials_xyzz, int count, void* d_out_xy, void* stream) {
    return sb_msm_combine_batch_device(curve, d_partials_xyzz, count, 1, 1, d_out_xy, stream);
}

int sb_msm_combine_batch_device(int curve, const void* d_partials_xyzz, int count, size_t batch, size_t stride, void* d_out_xy, void* stream) {
    if (!d_partials_xyzz || !d_out_xy || count < 0 || stride < batch) {
        set_error("sb_msm_combine_batch_device: bad argument");
        return SB_ERR_ARG;
    }
    SB_TRY(ensure_runtime());
    Runtime& rt = runtime();
    cudaStream_t st = stream ? (cudaStream_t)stream : rt.stream;
    if (!batch) return SB_OK;
    if (curve == CURVE_BN256) k_combine<Fq><<<(unsigned)batch, 4, 0, st>>>((const XYZZ<Fq>*)d_partials_xyzz, count, stride, (Affine<Fq>*)d_out_xy);
    else if (curve == CURVE_GRUMPKIN) k_combine<Fr><<<(unsigned)batch, 4, 0, st>>>((const XYZZ<Fr>*)d_partials_xyzz, count, stride, (Affine<Fr>*)d_out_xy);
    else {
        set_error("sb_msm_combine_device: unknown curve %d", curve);
        return SB_ERR_ARG;
    }
    SB_KERNEL_CHECK();
    return SB_OK;
}

}  // extern "C"
