// Host emulation of the batched-affine reduction rounds (sirius_b200/csrc/affine.cuh, kernel k_pair_round in
// msm.cu): the same per-thread phase functions and product-tree steps the kernel calls, executed thread by thread
// with the barriers replaced by loop boundaries.  Checked (tests/test_host_affine.py) against plain XYZZ bucket sums
// and against the oracle, on machines without a GPU.
#include <cstddef>
#include <cstring>
#include <vector>
#include "../../sirius_b200/csrc/affine.cuh"

using namespace sb;

template <class F, bool INDEXED, int B>
static void emulate_round(const Affine<F>* pts, const uint32_t* eidx, const uint32_t* off_in, const uint32_t* off_out, uint32_t KB,
                          Affine<F>* dst, size_t dst_cap, long* stats) {
    const uint32_t total_out = off_out[KB];
    const size_t per_block = (size_t)PR_THREADS * B;
    const size_t blocks = (total_out + per_block - 1) / per_block + 1;  // one idle block, as the kernel's grid is an upper bound
    std::vector<uint32_t> pos((size_t)PR_THREADS * B);
    std::vector<uint8_t> kind((size_t)PR_THREADS * B);
    std::vector<F> cp((size_t)PR_THREADS * B);
    std::vector<F> node(PR_NODES), ninv(PR_NODES);
    const PairSrc<F, INDEXED> src{pts, eidx};
    for (size_t blk = 0; blk < blocks; blk++) {
        if ((uint64_t)blk * per_block >= total_out) continue;
        for (int tid = 0; tid < PR_THREADS; tid++) {
            const uint32_t g = (uint32_t)(blk * PR_THREADS + tid);
            node[tid] = pair_forward<F, INDEXED, B>(src, off_in, off_out, KB, g, &pos[(size_t)tid * B], &kind[(size_t)tid * B], &cp[(size_t)tid * B]);
        }
        for (int l = 0; l < PR_LEVELS; l++)
            for (int tid = 0; tid < (PR_THREADS >> (l + 1)); tid++) pr_tree_up(node.data(), l, tid);
        ninv[pr_level_off(PR_LEVELS)] = inv_safegcd(node[pr_level_off(PR_LEVELS)]);
        for (int l = PR_LEVELS - 1; l >= 0; l--)
            for (int tid = 0; tid < (PR_THREADS >> l); tid++) pr_tree_down(node.data(), ninv.data(), l, tid);
        for (int tid = 0; tid < PR_THREADS; tid++) {
            const size_t g = blk * PR_THREADS + tid;
            for (int j = 0; j < B; j++) {
                const uint8_t k = kind[(size_t)tid * B + j] & 7;
                stats[k]++;
                if (k != PR_NONE && g * B + j >= dst_cap) stats[7]++;  // would write out of bounds
            }
            pair_backward<F, INDEXED, B>(src, &pos[(size_t)tid * B], &kind[(size_t)tid * B], &cp[(size_t)tid * B], ninv[tid], dst + g * B);
        }
    }
}

// Reduces the sorted entries by `rounds` affine rounds, then sums what is left per bucket with XYZZ mixed additions
// (what k_accumulate + k_fixup do) and normalises.  out_affine[b] = bucket sum; ref_affine[b] = the same sum by
// XYZZ additions alone.  Returns the number of out-of-bounds writes the round buffers would have seen (must be 0).
template <class F>
static long run_case(const Affine<F>* table, const uint32_t* eidx, const uint32_t* off0, uint32_t KB, int rounds, int B,
                     Affine<F>* out_affine, Affine<F>* ref_affine, long* stats) {
    const size_t M = off0[KB];
    std::vector<std::vector<uint32_t>> off(rounds + 1, std::vector<uint32_t>(KB + 1));
    for (uint32_t b = 0; b <= KB; b++) off[0][b] = off0[b];
    for (int r = 1; r <= rounds; r++) {
        uint32_t run = 0;
        for (uint32_t b = 0; b < KB; b++) {
            off[r][b] = run;
            const uint32_t c = off0[b + 1] - off0[b];
            run += (c + ((1u << r) - 1u)) >> r;
        }
        off[r][KB] = run;
    }
    // the kernel's buffers: a <= M/2 + KB, b <= M/4 + KB
    const size_t cap_a = M / 2 + KB + 1, cap_b = M / 4 + KB + 1;
    std::vector<Affine<F>> buf_a(cap_a + PR_THREADS * 16), buf_b(cap_b + PR_THREADS * 16);
    const Affine<F>* cur = table;
    for (int r = 0; r < rounds; r++) {
        Affine<F>* dst = (r & 1) ? buf_b.data() : buf_a.data();
        const size_t cap = (r & 1) ? cap_b : cap_a;
        if (r == 0) {
            if (B == 8) emulate_round<F, true, 8>(table, eidx, off[0].data(), off[1].data(), KB, dst, cap, stats);
            else emulate_round<F, true, 16>(table, eidx, off[0].data(), off[1].data(), KB, dst, cap, stats);
        } else {
            if (B == 8) emulate_round<F, false, 8>(cur, nullptr, off[r].data(), off[r + 1].data(), KB, dst, cap, stats);
            else emulate_round<F, false, 16>(cur, nullptr, off[r].data(), off[r + 1].data(), KB, dst, cap, stats);
        }
        cur = dst;
    }
    for (uint32_t b = 0; b < KB; b++) {
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t p = off[rounds][b]; p < off[rounds][b + 1]; p++) {
            if (rounds == 0) {
                const uint32_t e = eidx[p];
                xyzz_madd(acc, table[e & 0x7fffffffu], (e >> 31) != 0);
            } else xyzz_madd(acc, cur[p], false);
        }
        out_affine[b] = xyzz_to_affine(acc);
        XYZZ<F> ref = XYZZ<F>::identity();
        for (uint32_t p = off0[b]; p < off0[b + 1]; p++) {
            const uint32_t e = eidx[p];
            xyzz_madd(ref, table[e & 0x7fffffffu], (e >> 31) != 0);
        }
        ref_affine[b] = xyzz_to_affine(ref);
    }
    return stats[7];
}

extern "C" long haf_bucket_sums(int curve, const uint64_t* table, const uint32_t* eidx, const uint32_t* off0, uint32_t KB, int rounds,
                                int B, uint64_t* out_affine, uint64_t* ref_affine, long* stats /* [8] */) {
    for (int i = 0; i < 8; i++) stats[i] = 0;
    if (curve == CURVE_BN256)
        return run_case<Fq>((const Affine<Fq>*)table, eidx, off0, KB, rounds, B, (Affine<Fq>*)out_affine, (Affine<Fq>*)ref_affine, stats);
    return run_case<Fr>((const Affine<Fr>*)table, eidx, off0, KB, rounds, B, (Affine<Fr>*)out_affine, (Affine<Fr>*)ref_affine, stats);
}
