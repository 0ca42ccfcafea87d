// gx_walk.cuh -- K3: start-cell selection and traceback walk on the GPU.
//
// Replaces /root/reference/src/alignment/algo.rs:287-441 (retrace):
//   start cell  algo.rs:306-323  global (m,n); local = last maximum in row-major order
//   score       algo.rs:331
//   walk        algo.rs:339-422  state-less choice S > I > D (here: the stored 2-bit code),
//               off-by-one match label is_match(i,j) (algo.rs:354, sequence.rs:102-115),
//               open/extend labels from last_choice (algo.rs:373-379,388-394),
//               checked_sub index update (algo.rs:412-417), stop at (0,0) (algo.rs:419-421).
// One warp per pair, two levels of parallelism inside the inherently sequential walk:
// (a) a 256-row window of code chunks spanning the whole strip is fetched into shared memory with one round
//     of asynchronous 16-byte copies (cp.async) -- one HBM latency per window instead of one per code line;
// (b) run following: the path consists of runs of equal codes, so lane x inspects the cell x moves ahead in
//     the current direction, a ballot finds the run length, and up to 32 ops (labels, counters, coalesced
//     byte stores) are emitted per iteration.
#pragma once
#include "gx_common.cuh"
#include "gx_fill.cuh"

namespace gx {

// dynamic shared memory of one walk warp: two window buffers (codes + label characters).
// A window is 256 rows of one strip = 256/R (+1) row blocks + 31 steps of lane skew, 64/(R*K) steps per code chunk.
__host__ __device__ constexpr uint32_t walk_chunks(int K, int R) { return (uint32_t)((256 / R + 1 + 31 + 64 / (R * K) - 1) / (64 / (R * K)) + 1); }
__host__ __device__ constexpr uint32_t walk_buf_bytes(int K, int R) {
    return (uint32_t)(walk_chunks(K, R) * 32 * 16 + (256 + 32) + (32 * K + 32) + 15) & ~15u;
}
__host__ __device__ constexpr uint32_t walk_smem_bytes(int K, int R) { return 2 * walk_buf_bytes(K, R); }

template <int K, int R>
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
                // a winner right of the table can only come from a tile whose maximum is 0 (see gx_fill.cuh): the boundary
                // cell (m, n) this reduction starts from covers it
                const bool take = (tb.z <= (int)n) && ((tb.x > bv) || (tb.x == bv && (tb.y > bi || (tb.y == bi && tb.z > bj))));
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
            // Table maximum 0: every cell ties, and the last one in row-major order is (m, n) (Iterator::max_by,
            // algo.rs:311-322).  The tiles cannot be trusted to say so: a padded column right of the table also holds 0
            // and may have won its tile, which the filter above then dropped together with the tile's real zeros.
            if (bv <= 0) {
                bv = 0;
                bi = (int)m;
                bj = (int)n;
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
        using G = Geo<K, R>;
        constexpr int SPC = G::SPC;
        constexpr int WR = 256;                               // rows per window; the window spans the whole strip width
        constexpr int NCH = (int)walk_chunks(K, R);           // code chunks per fill-lane in a window
        // Two window buffers in dynamic shared memory: the walk reads buffer `cb`; the other one receives the window the
        // walk will probably need next (prefetched with cp.async while the walk runs).  Per buffer: code chunks
        // [chunk][fill-lane] + the label characters of the window: is_match(i,j) reads s1[i] and s2[j] (0-based: the
        // characters AFTER the cell's own, algo.rs:354), i.e. s1[i0w+1 ..] for the rows and s2[j0w+1 ..] for the columns.
        extern __shared__ __align__(16) uint8_t walk_smem[];
        constexpr uint32_t BUF_BYTES = walk_buf_bytes(K, R);
        uint32_t cb = 0;
        const uint4 *win = reinterpret_cast<const uint4 *>(walk_smem);
        const uint8_t *s1w = walk_smem + NCH * 32 * 16;
        const uint8_t *s2w = s1w + (WR + 32);
        uint32_t s1w0 = 0, s2w0 = 0;                          // sequence index of s1w[0] / s2w[0]
        // prefetched window (in buffer cb ^ 1): tile (np, ns), local rows [nr0, nr1]; np = 0xffffffff: none
        uint32_t np = 0xffffffffu, ns = 0, nr0 = 0, nr1 = 0;
        uint8_t *ops = P.ops + pd->ops_off;
        uint32_t nops = 0, n_match = 0, n_mis = 0, n_ext = 0, n_open = 0;
        uint32_t last = 0;  // AlignmentChoice::Match, algo.rs:338
        // current window: tile (wp, ws), local rows [wr0, wr1], first chunk wc0
        uint32_t wp = 0xffffffffu, ws = 0, wr0 = 0, wr1 = 0, wc0 = 0;

        // code of cell (ci, cj): 0 S / 1 I / 2 D / 3 stop; 7 = not in the current window (a run must end before it)
        // Straight-line (no branches): the lookup is on the loop-carried chain of the walk.
        const uint32_t bnd_row0 = local ? 3u : 1u, bnd_col0 = local ? 3u : 2u;
        auto cell_code = [&](uint32_t ci, uint32_t cj) __attribute__((always_inline)) -> uint32_t {
            const uint32_t jj = cj - 1u, ii = ci - 1u;            // wrap to 0xffffffff on the boundaries: never "in window"
            const uint32_t l = (jj % G::W) / K, k = jj % K;
            const uint32_t r = ii & (PANEL_H - 1);
            const bool inwin = ((ii >> PANEL_H_LOG2) == wp) & ((jj / G::W) == ws) & ((r - wr0) <= (wr1 - wr0));
            const uint32_t t = r / R + l;                         // the step at which fill-lane l worked on this row's block
            const uint32_t bitpos = (((t % SPC) * R + r % R) * K + k) * 2;
            const uint32_t widx = inwin ? (((t / SPC - wc0) * 32 + l) * 4 + (bitpos >> 5)) : 0u;
            const uint32_t word = reinterpret_cast<const uint32_t *>(win)[widx];
            uint32_t code = inwin ? ((word >> (bitpos & 31u)) & 3u) : 7u;
            code = (cj == 0u) ? bnd_col0 : code;                  // column 0: only delete_score is finite; local stops
            code = (ci == 0u) ? ((cj == 0u) ? 0u : bnd_row0) : code;   // row 0; origin: sub_score == max == 0 (algo.rs:195-202)
            return code;
        };

        unsigned long long dbg_iters = 0, dbg_reloads = 0;
        long long dbg_reload_cyc = 0;
        const long long dbg_t0 = clock64();
        if (i == 0 && j == 0) {
            // only possible for m == n == 0: one Match at (0,0) (None == None), then both checked_sub fail
            if (lane == 0) ops[0] = 0;
            nops = 1;
            n_match = 1;
        } else {
            uint32_t c0 = cell_code(i, j);
            uint32_t end_i = i, end_j = j;
            for (;;) {
                dbg_iters++;
                if (c0 == 7u) {
                    dbg_reloads++;
                    const long long dbg_r0 = P.debug ? clock64() : 0;
                    const uint32_t jj = j - 1, ii = i - 1;
                    const uint32_t tp = ii >> PANEL_H_LOG2, ts = jj / G::W, tr = ii & (PANEL_H - 1);
                    // copies of `nch` chunks of tile (p, s) starting at chunk c0w into buffer b (asynchronous)
                    auto issue_codes = [&](uint32_t b, uint32_t p, uint32_t s_, uint32_t r0, uint32_t r1) __attribute__((always_inline)) {
                        const uint32_t c0w = (r0 / R) / SPC;
                        const uint32_t nch = (r1 / R + 31) / SPC - c0w + 1;   // <= NCH
                        const uint4 *tile = reinterpret_cast<const uint4 *>(P.codes + pd->codes_off +
                                                                             (uint64_t)(p * pd->S + s_) * pd->tile_code_bytes) +
                                            (size_t)c0w * 32 + lane;
                        uint32_t dst = smem_u32(walk_smem + b * BUF_BYTES) + (uint32_t)lane * 16u;
                        for (uint32_t q2 = 0; q2 < nch; ++q2) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(tile + (size_t)q2 * 32) : "memory");
                            dst += 32 * 16;
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    };
                    __syncwarp();
                    const bool hit = (np == tp) && (ns == ts) && (tr >= nr0) && (tr <= nr1);
                    if (hit) {
                        cb ^= 1u;               // the prefetched window becomes the current one
                        wp = np;
                        ws = ns;
                        wr0 = nr0;
                        wr1 = nr1;
                    } else {
                        // drain whatever is still landing in the other buffer, then load the window that ends at this row:
                        // rows [r-255, r] x all 32 fill-lanes of the tile
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        wp = tp;
                        ws = ts;
                        wr1 = tr;
                        wr0 = (wr1 >= WR - 1) ? wr1 - (WR - 1) : 0;
                        issue_codes(cb, wp, ws, wr0, wr1);
                    }
                    np = 0xffffffffu;
                    wc0 = (wr0 / R) / SPC;
                    uint8_t *base = walk_smem + cb * BUF_BYTES;
                    win = reinterpret_cast<const uint4 *>(base);
                    uint8_t *s1wm = base + NCH * 32 * 16, *s2wm = s1wm + (WR + 32);
                    s1w = s1wm;
                    s2w = s2wm;
                    // label characters: rows wr0..wr1 of panel wp -> s1 indices (i-1)+1, columns of strip ws -> s2 indices (j-1)+1
                    s1w0 = (wp << PANEL_H_LOG2) + wr0 + 1u;
                    s2w0 = ws * G::W + 1u;
                    // (all loads first, then the stores: one memory latency for the whole window, not one per byte)
                    {
                        uint8_t a1[WR / 32], a2[K];
#pragma unroll
                        for (int q2 = 0; q2 < WR / 32; ++q2) {
                            const uint32_t k = (uint32_t)lane + 32u * q2;
                            a1[q2] = (k <= wr1 - wr0 && s1w0 + k < m) ? __ldg(s1 + s1w0 + k) : (uint8_t)0;
                        }
#pragma unroll
                        for (int q2 = 0; q2 < K; ++q2) {
                            const uint32_t k = (uint32_t)lane + 32u * q2;
                            a2[q2] = (s2w0 + k < n) ? __ldg(s2 + s2w0 + k) : (uint8_t)0;
                        }
#pragma unroll
                        for (int q2 = 0; q2 < WR / 32; ++q2) s1wm[lane + 32 * q2] = a1[q2];
#pragma unroll
                        for (int q2 = 0; q2 < K; ++q2) s2wm[lane + 32 * q2] = a2[q2];
                    }
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                    c0 = cell_code(i, j);
                    // predict the next window and start fetching it into the other buffer: a diagonal path leaves this window
                    // on the left after (columns to the strip's left edge) rows, at the top after (rows above the entry) rows
                    {
                        const uint32_t cin = jj % G::W + 1u, rows_up = tr - wr0 + 1u;
                        const bool left_ok = ws > 0, top_ok = (wr0 > 0) || (wp > 0);
                        const bool go_left = left_ok && (cin <= rows_up || !top_ok);
                        if (go_left) {
                            np = wp;
                            ns = ws - 1u;
                            nr0 = wr0;
                            nr1 = wr1;
                        } else if (top_ok) {
                            ns = ws;
                            if (wr0 > 0) {
                                np = wp;
                                nr1 = wr0 - 1u;
                                nr0 = (nr1 >= WR - 1) ? nr1 - (WR - 1) : 0;
                            } else {
                                np = wp - 1u;
                                nr1 = PANEL_H - 1;
                                nr0 = PANEL_H - WR;
                            }
                        }
                        if (np != 0xffffffffu) issue_codes(cb ^ 1u, np, ns, nr0, nr1);
                    }
                    if (P.debug) dbg_reload_cyc += clock64() - dbg_r0;
                }
                if (c0 == 3u) break;                      // local alignment ends on a boundary cell (algo.rs:401-405)
                // every lane looks x steps ahead in the direction of c0; the run ends at the first different code
                const uint32_t x = (uint32_t)lane;
                const bool diag = (c0 == 0u);
                const uint32_t di = (c0 != 1u) ? 1u : 0u, dj = (c0 != 2u) ? 1u : 0u;
                const bool reach = (x * di <= i) & (x * dj <= j);
                const uint32_t ci = i - (reach ? x * di : 0u), cj = j - (reach ? x * dj : 0u);
                uint32_t cx = cell_code(ci, cj);
                cx = (!reach || (x > 0 && ci == 0 && cj == 0)) ? 7u : cx;   // (0,0) is never emitted after a move (algo.rs:419-421)
                const uint32_t same = __ballot_sync(0xffffffffu, cx == c0);
                const uint32_t run = (same == 0xffffffffu) ? 32u : (uint32_t)(__ffs((int)~same) - 1);   // >= 1
                // the cell the walk reaches next is the one lane `run` just looked at: its code starts the next iteration
                const uint32_t c_next = __shfl_sync(0xffffffffu, cx, (int)(run & 31u));
                const bool mine = x < run;
                // labels.  Diagonal: is_match(i, j), Option<u8> equality with None == None (sequence.rs:113-114); run cells lie in
                // the window, so their characters are in s1w / s2w.  Gaps: open/extend from last_choice (algo.rs:373-379, 388-394).
                const bool lab = diag & mine;
                const int a = (lab && ci < m) ? (int)s1w[ci - s1w0] : -1;
                const int b = (lab && cj < n) ? (int)s2w[cj - s2w0] : -1;
                const uint32_t mmask = __ballot_sync(0xffffffffu, lab && a == b);
                const uint32_t nm = (uint32_t)__popc(mmask);
                const uint32_t ext = (c0 == 1u) ? 2u : 3u;          // Insert / Delete
                const bool opens = !diag && (last != ext);
                const uint32_t gop = (x == 0 && opens) ? ext + 2u : ext;   // OpenInsert = 4, OpenDelete = 5
                const uint32_t op = diag ? ((a == b) ? 0u : 1u) : gop;
                if (mine) ops[nops + x] = (uint8_t)op;
                n_match += nm;
                n_mis += diag ? run - nm : 0u;
                n_open += opens ? 1u : 0u;
                n_ext += diag ? 0u : (opens ? run - 1u : run);
                last = diag ? 0u : ext;
                nops += run;
                // last emitted cell, then the checked_sub move (algo.rs:412-417); run cells are all inside the table
                end_i = i - (run - 1u) * di;
                end_j = j - (run - 1u) * dj;
                const bool i_none = di && (i < run), j_none = dj && (j < run);
                if (i_none && j_none) break;
                i = i_none ? 0u : i - run * di;
                j = j_none ? 0u : j - run * dj;
                if (i == 0 && j == 0) break;
                // lane `run` looked at exactly (i, j) unless the run used all 32 lanes, the move was clamped, or that
                // cell was outside the window (7): then look it up (and reload the window at the top of the loop).
                // A warp-uniform branch: one lone warp runs the walk, so instructions issued are what an iteration
                // costs, and the second lookup is a fifth of them.
                if (run < 32u && !i_none && !j_none && c_next != 7u) c0 = c_next;
                else c0 = cell_code(i, j);
            }
            res.end_i = end_i;
            res.end_j = end_j;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (P.debug) {
            res.lcs_at_first_max = dbg_iters | (dbg_reloads << 32);
            res.fill_ms = (double)(clock64() - dbg_t0);
            res.walk_ms = (double)dbg_reload_cyc;
        }
        res.n_ops = nops;
        res.matches = n_match;
        res.mismatches = n_mis;
        res.gap_extensions = n_ext;
        res.opening_gaps = n_open;
    }
    if (lane == 0) P.results[q] = res;
}

}  // namespace gx
