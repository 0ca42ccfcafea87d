// gx_lcs.cuh -- second return value of alignment_table (GX_FLAG_LCS_AT_MAX).
//
// Reference: /root/reference/src/alignment/algo.rs:112-121 (max_matches), :250-255 (the three match-count lanes),
// :258-262 (max_cell: FIRST interior cell, row-major, strict `<`, that attains the table-wide maximum of
// score_max(0,0,0,is_local)), :279-281 (return alignment_table[max_cell].max_matches()).
// The match-count lanes are the plain LCS-length DP  L[i][j] = max(L[i][j-1], L[i-1][j], L[i-1][j-1] + eq)
// (SURVEY.md 3.2), independent of the scores, so the value is  LCS(s1[0..i*), s2[0..j*))  for max_cell = (i*, j*).
//
// (i*, j*) comes from a score-only pass of the fill kernel with TRACK == 3 (first maximum).  The LCS length is
// computed with the bit-vector recurrence  U = V & M[c];  V = (V + U) | (V & ~U)  (one bit per column of s2, all ones
// at the start, LCS = number of zero bits), one warp per pair: lane l owns 32 consecutive 32-bit words, rows flow
// through the lanes as a systolic skew and the addition's carry crosses lanes with one shuffle per step.  Tables wider
// than 32768 columns are processed in column blocks; the carry out of a block is kept per row in a bit array.
#pragma once
#include "gx_common.cuh"

namespace gx {

constexpr int LCS_WPL = 32;                          // 32-bit words per lane
constexpr int LCS_BLOCK_WORDS = 32 * LCS_WPL;        // words per column block
constexpr int LCS_BLOCK_COLS = 32 * LCS_BLOCK_WORDS; // 32768 columns

struct LcsParams {
    const uint8_t *blob;
    const PairDesc *pairs;
    uint32_t n_pairs;
    const int4 *tile_first;      // per tile: (V, i, j) of the first maximum (fill pass with TRACK == 3)
    uint32_t *masks;             // [pair][256][LCS_BLOCK_WORDS] match masks of the current column block
    uint32_t *carry;             // [pair][2][carry_words] carry bits between column blocks (ping-pong), one bit per row
    uint32_t carry_words;
    DevResult *results;
};

__global__ void __launch_bounds__(32) gx_lcs_kernel(const LcsParams P) {
    const uint32_t q = blockIdx.x;
    if (q >= P.n_pairs) return;
    const int lane = threadIdx.x;
    const PairDesc *pd = P.pairs + q;
    const uint8_t *s1 = P.blob + pd->s1_off;
    const uint8_t *s2 = P.blob + pd->s2_off;

    // ---- max_cell: larger value, then smaller i, then smaller j; (0,0) when the table has no interior cell
    int bv = INT32_MIN, bi = 0, bj = 0;
    const uint32_t ntile = pd->S * pd->P;
    for (uint32_t x = lane; x < ntile; x += 32) {
        const int4 tb = P.tile_first[pd->tile_base + x];
        const bool take = (tb.z <= (int)pd->n) && ((tb.x > bv) || (tb.x == bv && tb.x != INT32_MIN && (tb.y < bi || (tb.y == bi && tb.z < bj))));
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
        const bool take = (ov > bv) || (ov == bv && ov != INT32_MIN && (oi < bi || (oi == bi && oj < bj)));
        bv = take ? ov : bv;
        bi = take ? oi : bi;
        bj = take ? oj : bj;
    }
    const uint32_t rows = (bv == INT32_MIN) ? 0u : (uint32_t)bi;   // i*
    const uint32_t cols = (bv == INT32_MIN) ? 0u : (uint32_t)bj;   // j*

    uint32_t *masks = P.masks + (size_t)q * 256 * LCS_BLOCK_WORDS;
    uint32_t *carry_a = P.carry + (size_t)q * 2 * P.carry_words;
    uint32_t *carry_b = carry_a + P.carry_words;
    unsigned long long zeros = 0;

    for (uint32_t c0 = 0; c0 < cols; c0 += LCS_BLOCK_COLS) {
        const uint32_t bc = min((uint32_t)LCS_BLOCK_COLS, cols - c0);   // columns of this block
        const uint32_t nw = (bc + 31) / 32;
        const bool first_block = (c0 == 0), last_block = (c0 + bc == cols);
        // ---- match masks of the block: masks[sym][w] bit b <=> s2[c0 + 32 w + b] == sym   (lane owns word w: no atomics)
        __syncwarp();
        for (uint32_t x = lane; x < 256u * LCS_BLOCK_WORDS / 4; x += 32) reinterpret_cast<uint4 *>(masks)[x] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        for (uint32_t w = lane; w < nw; w += 32) {
            const uint32_t nb = min(32u, bc - 32u * w);
            for (uint32_t b = 0; b < nb; ++b) {
                const uint32_t sym = s2[c0 + 32u * w + b];
                masks[sym * LCS_BLOCK_WORDS + w] |= 1u << b;
            }
        }
        __threadfence_block();
        __syncwarp();

        // ---- systolic pass over the rows: lane l works on row t - l at step t
        uint32_t V[LCS_WPL], M[LCS_WPL], Mn[LCS_WPL];   // Mn: masks of the next row, in flight while this row computes
#pragma unroll
        for (int w = 0; w < LCS_WPL; ++w) {
            V[w] = 0xffffffffu;
            M[w] = Mn[w] = 0u;
        }
        const uint32_t *mbase = masks + lane * LCS_WPL;
        auto load_masks = [&](int r) {
            if (r >= 0 && (uint32_t)r < rows) {
                const uint4 *mp = reinterpret_cast<const uint4 *>(mbase + (uint32_t)s1[r] * LCS_BLOCK_WORDS);
#pragma unroll
                for (int v4 = 0; v4 < LCS_WPL / 4; ++v4) {
                    const uint4 v = mp[v4];
                    Mn[4 * v4 + 0] = v.x;
                    Mn[4 * v4 + 1] = v.y;
                    Mn[4 * v4 + 2] = v.z;
                    Mn[4 * v4 + 3] = v.w;
                }
            }
        };
        uint32_t cout_prev = 0;      // carry out of this lane's last word at the previous step
        uint32_t cbits = 0;          // lane 31: carry-out bits of 32 consecutive rows
        const int steps = (int)rows + 31;
        load_masks(0 - lane);
        for (int t = 0; t < steps; ++t) {
            const int r = t - lane;
#pragma unroll
            for (int w = 0; w < LCS_WPL; ++w) M[w] = Mn[w];
            load_masks(r + 1);
            const bool active = (r >= 0) && ((uint32_t)r < rows);
            uint32_t cin = __shfl_up_sync(0xffffffffu, cout_prev, 1);
            if (lane == 0) cin = (first_block || !active) ? 0u : ((carry_a[(uint32_t)r >> 5] >> ((uint32_t)r & 31u)) & 1u);
            uint32_t carry = cin;
            if (active) {
#pragma unroll
                for (int w = 0; w < LCS_WPL; ++w) {
                    const uint32_t U = V[w] & M[w];
                    const unsigned long long sum = (unsigned long long)V[w] + U + carry;
                    carry = (uint32_t)(sum >> 32);
                    V[w] = (uint32_t)sum | (V[w] & ~U);
                }
            }
            cout_prev = active ? carry : 0u;
            if (lane == 31 && !last_block && active) {
                cbits |= carry << ((uint32_t)r & 31u);
                if ((((uint32_t)r & 31u) == 31u) || (uint32_t)r + 1 == rows) {
                    carry_b[(uint32_t)r >> 5] = cbits;
                    cbits = 0;
                }
            }
        }
        // zero bits inside the block's real columns
        uint32_t z = 0;
#pragma unroll
        for (int w = 0; w < LCS_WPL; ++w) {
            const uint32_t gw = (uint32_t)lane * LCS_WPL + w;          // word index inside the block
            uint32_t valid = 0u;
            if (gw < nw) valid = (32u * (gw + 1) <= bc) ? 0xffffffffu : ((1u << (bc - 32u * gw)) - 1u);
            z += (uint32_t)__popc(~V[w] & valid);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) z += __shfl_xor_sync(0xffffffffu, z, off);
        zeros += z;
        // the next block reads what this one wrote
        __threadfence_block();
        __syncwarp();
        uint32_t *tmp = carry_a;
        carry_a = carry_b;
        carry_b = tmp;
    }
    if (lane == 0) P.results[q].lcs_at_first_max = zeros;
}

// ------------------------------------------------------------------------------------------------
// Small-table visualiser support: the three score planes print_scores_table prints
// (/root/reference/src/alignment/display.rs:190-220), i.e. insert_score / delete_score / sub_score of every cell exactly
// as alignment_table stores them (algo.rs:195-248), int64, boundary "minus infinity" = i64::MIN + |g+h| (algo.rs:166).
// The reference only prints tables with m < 200 and n < 2000, so one CTA sweeping the anti-diagonals is plenty.
struct PlanesParams {
    const uint8_t *s1, *s2;
    uint32_t m, n;
    long long a, b, g, h;
    int is_local;
    long long *pi, *pd, *ps;     // row-major (m+1) x (n+1)
};

__device__ __forceinline__ long long planes_mx(long long I, long long S, long long D, long long x, long long y, long long z, bool local) {
    // ComputeScore::score_max (algo.rs:98-107): max(I+x, S+y, D+z, local ? 0 : i64::MIN), wrapping adds
    const long long vi = (long long)((unsigned long long)I + (unsigned long long)x);
    const long long vs = (long long)((unsigned long long)S + (unsigned long long)y);
    const long long vd = (long long)((unsigned long long)D + (unsigned long long)z);
    long long v = vi > vs ? vi : vs;
    v = v > vd ? v : vd;
    const long long f = local ? 0ll : (long long)0x8000000000000000ull;
    return v > f ? v : f;
}

__global__ void __launch_bounds__(256) gx_planes_kernel(const PlanesParams P) {
    const uint32_t C = P.n + 1;
    const long long gh = P.g + P.h;
    const long long neg_inf = (long long)(0x8000000000000000ull + (unsigned long long)(gh < 0 ? -gh : gh));
    const bool local = P.is_local != 0;
    for (uint32_t d = 0; d <= P.m + P.n; ++d) {
        for (uint32_t i = threadIdx.x; i <= P.m && i <= d; i += blockDim.x) {
            const uint32_t j = d - i;
            if (j > P.n) continue;
            long long ci, cd, cs;
            if (i == 0 && j == 0) {
                ci = cd = cs = 0;                                   // algo.rs:195-202
            } else if (j == 0) {
                ci = neg_inf; cd = P.h + (long long)i * P.g; cs = neg_inf;   // algo.rs:204-211
            } else if (i == 0) {
                ci = P.h + (long long)j * P.g; cd = neg_inf; cs = neg_inf;   // algo.rs:213-220
            } else {
                const size_t top = (size_t)i * C + (j - 1), left = (size_t)(i - 1) * C + j, tl = (size_t)(i - 1) * C + (j - 1);
                const bool eq = P.s1[i - 1] == P.s2[j - 1];         // is_match(i-1, j-1), algo.rs:227
                ci = planes_mx(P.pi[top], P.ps[top], P.pd[top], P.g, gh, gh, local);        // algo.rs:229-235
                cd = planes_mx(P.pi[left], P.ps[left], P.pd[left], gh, gh, P.g, local);     // algo.rs:236-242
                cs = (eq ? P.a : P.b) + planes_mx(P.pi[tl], P.ps[tl], P.pd[tl], 0, 0, 0, local);   // algo.rs:244-248
            }
            const size_t at = (size_t)i * C + j;
            P.pi[at] = ci;
            P.pd[at] = cd;
            P.ps[at] = cs;
        }
        __syncthreads();
    }
}

}  // namespace gx
