// gx_walk.cuh -- K3: start-cell selection and traceback walk on the GPU.
//
// Replaces /root/reference/src/alignment/algo.rs:287-441 (retrace):
//   start cell  algo.rs:306-323  global (m,n); local = last maximum in row-major order
//   score       algo.rs:331
//   walk        algo.rs:339-422  state-less choice S > I > D (here: the stored 2-bit code),
//               off-by-one match label is_match(i,j) (algo.rs:354, sequence.rs:102-115),
//               open/extend labels from last_choice (algo.rs:373-379,388-394),
//               checked_sub index update (algo.rs:412-417), stop at (0,0) (algo.rs:419-421).
// One warp per pair.  All 32 lanes replay the same scalar walk (uniform loads, no divergence);
// the lanes are used to (a) prefetch the code lines a diagonal run will touch next and
// (b) buffer 32 op bytes so that the ops list is written with coalesced 32-byte stores.
#pragma once
#include "gx_common.cuh"

namespace gx {

template <int K>
__device__ __forceinline__ const uint32_t *code_word_addr(const uint8_t *codes, const PairDesc *pd, uint32_t i, uint32_t j,
                                                          uint32_t &shift) {
    constexpr int W = 32 * K;
    constexpr int SPC = 64 / K;
    const uint32_t jj = j - 1, ii = i - 1;
    const uint32_t s = jj / W;
    const uint32_t l = (jj % W) / K;
    const uint32_t k = jj % K;
    const uint32_t p = ii >> PANEL_H_LOG2;
    const uint32_t r = ii & (PANEL_H - 1);
    const uint32_t t = r + l;
    const uint32_t chunk = t / SPC;
    const uint32_t u = t % SPC;
    const uint32_t bitpos = u * 2 * K + 2 * k;
    shift = bitpos & 31u;
    const uint64_t off = pd->codes_off + (uint64_t)(p * pd->S + s) * pd->tile_code_bytes + ((uint64_t)chunk * 32u + l) * 16u +
                         (bitpos >> 5) * 4u;
    return reinterpret_cast<const uint32_t *>(codes + off);
}

template <int K>
__global__ void __launch_bounds__(32) gx_walk_kernel(const WalkParams P) {
    const uint32_t q = blockIdx.x;
    if (q >= P.n_pairs) return;
    const int lane = threadIdx.x;
    const PairDesc *pd = P.pairs + q;
    const uint32_t m = pd->m, n = pd->n;
    const uint8_t *s1 = P.blob + pd->s1_off;
    const uint8_t *s2 = P.blob + pd->s2_off;
    const bool local = P.is_local != 0;

    // ---- start cell and score
    long long score;
    uint32_t i = m, j = n;
    if (local) {
        int bv = 0, bi = (int)m, bj = (int)n;  // boundary cells have V = 0; the last cell of the table wins a tie at 0
        if (m > 0 && n > 0) {
            bv = -1;
            bi = 0;
            bj = 0;
            const uint32_t ntile = pd->S * pd->P;
            for (uint32_t x = lane; x < ntile; x += 32) {
                const int4 tb = P.tile_best[pd->tile_base + x];
                const bool take = (tb.x > bv) || (tb.x == bv && (tb.y > bi || (tb.y == bi && tb.z > bj)));
                if (take) {
                    bv = tb.x;
                    bi = tb.y;
                    bj = tb.z;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                const bool take = (ov > bv) || (ov == bv && (oi > bi || (oi == bi && oj > bj)));
                bv = take ? ov : bv;
                bi = take ? oi : bi;
                bj = take ? oj : bj;
            }
        }
        score = bv;
        if (P.have_best) {
            i = (uint32_t)bi;
            j = (uint32_t)bj;
        }
    } else {
        if (m == 0 && n == 0) score = 0;
        else if (m == 0) score = (long long)P.h + (long long)n * P.g;   // algo.rs:213-220
        else if (n == 0) score = (long long)P.h + (long long)m * P.g;   // algo.rs:204-211
        else score = (long long)P.top[pd->top_off + (n - 1)].x - P.hg;
    }

    DevResult res;
    res.score = score;
    res.start_i = i;
    res.start_j = j;
    res.end_i = i;
    res.end_j = j;
    res.n_ops = 0;
    res.matches = res.mismatches = res.gap_extensions = res.opening_gaps = 0;
    res.lcs_at_first_max = 0;
    res.fill_ms = res.walk_ms = 0.0;

    if (P.traceback) {
        uint8_t *ops = P.ops + pd->ops_off;
        uint32_t nops = 0, n_match = 0, n_mis = 0, n_ext = 0, n_open = 0;
        uint32_t last = 0;  // AlignmentChoice::Match, algo.rs:338
        uint32_t myop = 0;
        for (;;) {
            uint32_t c;
            if (i == 0 && j == 0) c = 0;               // origin: sub_score == max == 0
            else if (j == 0) c = local ? 3u : 2u;      // column 0: only delete_score is finite (global); local stops
            else if (i == 0) c = local ? 3u : 1u;      // row 0
            else {
                uint32_t sh;
                const uint32_t *wp = code_word_addr<K>(P.codes, pd, i, j, sh);
                c = (__ldg(wp) >> sh) & 3u;
                if ((nops & 7u) == 0u) {
                    // prefetch along the diagonal: lane l touches the line of cell (i-8(l+1), j-8(l+1))
                    const uint32_t d = 8u * (uint32_t)(lane + 1);
                    if (i > d && j > d) {
                        uint32_t sh2;
                        const uint32_t *pp = code_word_addr<K>(P.codes, pd, i - d, j - d, sh2);
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
                    }
                }
            }
            if (c == 3u) break;
            uint32_t op;
            bool i_none = false, j_none = false;
            uint32_t ni = i, nj = j;
            if (c == 0u) {
                const int a = (i < m) ? (int)__ldg(s1 + i) : -1;   // Option<u8>: None == None is a match
                const int b = (j < n) ? (int)__ldg(s2 + j) : -1;
                if (a == b) { op = 0u; n_match++; }
                else { op = 1u; n_mis++; }
                last = op;
                if (i == 0) i_none = true; else ni = i - 1;
                if (j == 0) j_none = true; else nj = j - 1;
            } else if (c == 1u) {
                if (last == 2u) { op = 2u; n_ext++; }
                else { op = 4u; n_open++; }
                last = 2u;
                if (j == 0) j_none = true; else nj = j - 1;
            } else {
                if (last == 3u) { op = 3u; n_ext++; }
                else { op = 5u; n_open++; }
                last = 3u;
                if (i == 0) i_none = true; else ni = i - 1;
            }
            if ((nops & 31u) == (uint32_t)lane) myop = op;
            res.end_i = i;
            res.end_j = j;
            nops++;
            if ((nops & 31u) == 0u) ops[nops - 32u + lane] = (uint8_t)myop;
            if (i_none && j_none) break;
            i = i_none ? 0u : ni;
            j = j_none ? 0u : nj;
            if (i == 0 && j == 0) break;
        }
        if ((uint32_t)lane < (nops & 31u)) ops[(nops & ~31u) + lane] = (uint8_t)myop;
        res.n_ops = nops;
        res.matches = n_match;
        res.mismatches = n_mis;
        res.gap_extensions = n_ext;
        res.opening_gaps = n_open;
    }
    if (lane == 0) P.results[q] = res;
}

}  // namespace gx
