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

// dynamic shared memory of one walk CTA: a 1 KB control block (descriptor ring between the two warps, tail word,
// debug counters) and two code-window buffers.
// A window is 256 rows of one strip = 256/R (+1) row blocks + 31 steps of lane skew, 64/(R*K) steps per code chunk.
__host__ __device__ constexpr uint32_t walk_chunks(int K, int R) { return (uint32_t)((256 / R + 1 + 31 + 64 / (R * K) - 1) / (64 / (R * K)) + 1); }
__host__ __device__ constexpr uint32_t walk_buf_bytes(int K, int R) { return walk_chunks(K, R) * 32 * 16; }
constexpr uint32_t WALK_CTRL_BYTES = 1024;
constexpr uint32_t WALK_RING = 32;          // run descriptors in flight between the path warp and the emit warp
__host__ __device__ constexpr uint32_t walk_smem_bytes(int K, int R) { return WALK_CTRL_BYTES + 2 * walk_buf_bytes(K, R); }

__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_volatile_uint4(uint4 *p, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_volatile_uint4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}

// One CTA of TWO warps per pair (one warp when no traceback is wanted).  The walk is m+n dependent steps, so what it costs
// is the instructions ONE warp has to issue per step of the dependent chain (a lone warp issues ~0.4 instructions per
// clock).  The chain is therefore kept to the path finding alone:
//   path warp (warp 0)  follows the stored direction codes: code windows in shared memory (cp.async, double-buffered,
//                       next window prefetched), run following (lane x inspects the cell x moves ahead, a ballot gives
//                       the run length), and pushes one descriptor (i, j, direction, run length) per run into a ring;
//   emit warp (warp 1)  pops descriptors and does everything that is not on the chain: the off-by-one match labels
//                       is_match(i, j) (characters straight from global memory / L1), open/extend labels, the four
//                       counters, the coalesced op stores, the end cell.
template <int K, int R>
__global__ void __launch_bounds__(64) gx_walk_kernel(const WalkParams P) {
    const uint32_t q = blockIdx.x;
    if (q >= P.n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
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

    if (!P.traceback) {
        if (threadIdx.x == 0) P.results[q] = res;
        return;
    }
    using G = Geo<K, R>;
    constexpr int SPC = G::SPC;
    constexpr int WR = 256;                               // rows per window; the window spans the whole strip width
    constexpr int NCH = (int)walk_chunks(K, R);           // code chunks per fill-lane in a window
    extern __shared__ __align__(16) uint8_t walk_smem[];
    // descriptor ring, LL style: {i, j, code | run << 8, sequence number}.  One lane writes a descriptor with ONE 16-byte
    // shared-memory store and the emit warp polls the slot until it carries the sequence number it expects: no head word,
    // no fence on the path warp's dependent chain.  `tail` (descriptors consumed) is only read when the ring looks full.
    uint4 *ring = reinterpret_cast<uint4 *>(walk_smem);
    uint32_t *tail = reinterpret_cast<uint32_t *>(walk_smem + WALK_RING * 16);
    unsigned long long *dbg = reinterpret_cast<unsigned long long *>(walk_smem + WALK_RING * 16 + 16);
    uint8_t *bufs = walk_smem + WALK_CTRL_BYTES;
    if (threadIdx.x < WALK_RING) ring[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);     // sequence numbers start at 1
    if (threadIdx.x == 0) sts_volatile_u32(tail, 0u);
    __syncthreads();

    if (wid == 0) {
        // ================================================================ path warp
        // Two window buffers: the walk reads buffer `cb`; the other one receives the window the walk will probably need
        // next (prefetched with cp.async while the walk runs).  Per buffer: code chunks [chunk][fill-lane].
        constexpr uint32_t BUF_BYTES = walk_buf_bytes(K, R);
        uint32_t cb = 0;
        const uint4 *win = reinterpret_cast<const uint4 *>(bufs);
        // prefetched window (in buffer cb ^ 1): tile (np, ns), local rows [nr0, nr1]; np = 0xffffffff: none
        uint32_t np = 0xffffffffu, ns = 0, nr0 = 0, nr1 = 0;
        // current window: tile (wp, ws), local rows [wr0, wr1], first chunk wc0
        uint32_t wp = 0xffffffffu, ws = 0, wr0 = 0, wr1 = 0, wc0 = 0;
        uint32_t pushed = 0, tail_seen = 0;
        auto push = [&](uint32_t pi, uint32_t pj, uint32_t meta) __attribute__((always_inline)) {
            if (pushed - tail_seen >= WALK_RING) {           // looks full: refresh the consumer's count (rare: the emit warp is faster)
                do {
                    tail_seen = lds_volatile_u32(tail);
                } while (pushed - tail_seen >= WALK_RING);
            }
            pushed++;
            // (the lap number rides in the top bits of the first half as well: a reader that caught the two 8-byte halves of the
            // slot from different laps -- should 16-byte shared accesses ever be split -- does not accept it)
            if (lane == 0) sts_volatile_uint4(ring + (pushed - 1u) % WALK_RING, make_uint4(pi | (((pushed >> 5) & 7u) << 29), pj, meta, pushed));
        };

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
            GX_CHECK(P.check, widx < walk_buf_bytes(K, R) / 4, 21);
            const uint32_t word = reinterpret_cast<const uint32_t *>(win)[widx];
            uint32_t code = inwin ? ((word >> (bitpos & 31u)) & 3u) : 7u;
            code = (cj == 0u) ? bnd_col0 : code;                  // column 0: only delete_score is finite; local stops
            code = (ci == 0u) ? ((cj == 0u) ? 0u : bnd_row0) : code;   // row 0; origin: sub_score == max == 0 (algo.rs:195-202)
            return code;
        };

        unsigned long long dbg_iters = 0, dbg_reloads = 0;
        long long dbg_reload_cyc = 0;
        bool left_band = false;
        const long long dbg_t0 = clock64();
        if (!(i == 0 && j == 0)) {
            uint32_t c0 = cell_code(i, j);
            for (;;) {
                dbg_iters++;
                if (c0 == 7u) {
                    dbg_reloads++;
                    const long long dbg_r0 = P.debug ? clock64() : 0;
                    const uint32_t jj = j - 1, ii = i - 1;
                    const uint32_t tp = ii >> PANEL_H_LOG2, ts = jj / G::W, tr = ii & (PANEL_H - 1);
                    // code band: a tile away from the table's diagonal holds no codes.  The path has left the band: give
                    // up -- the host repeats the execute with codes everywhere (exactness never depends on the band).
                    if (P.tile_codes && P.tile_codes[pd->tile_base + tp * pd->S + ts] == 0) {
                        left_band = true;
                        break;
                    }
                    // copies of `nch` chunks of tile (p, s) starting at chunk c0w into buffer b (asynchronous)
                    auto issue_codes = [&](uint32_t b, uint32_t p, uint32_t s_, uint32_t r0, uint32_t r1) __attribute__((always_inline)) {
                        const uint32_t c0w = (r0 / R) / SPC;
                        const uint32_t nch = (r1 / R + 31) / SPC - c0w + 1;   // <= NCH
                        GX_CHECK(P.check, nch <= (uint32_t)NCH && (uint64_t)(c0w + nch) * 512 <= pd->tile_code_bytes &&
                                              pd->codes_off + (uint64_t)(p * pd->S + s_ + 1) * pd->tile_code_bytes <= P.code_bytes, 23);
                        const uint4 *tile = reinterpret_cast<const uint4 *>(P.codes + pd->codes_off +
                                                                             (uint64_t)(p * pd->S + s_) * pd->tile_code_bytes) +
                                            (size_t)c0w * 32 + lane;
                        uint32_t dst = smem_u32(bufs + b * BUF_BYTES) + (uint32_t)lane * 16u;
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
                    win = reinterpret_cast<const uint4 *>(bufs + cb * BUF_BYTES);
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
                        if (np != 0xffffffffu && P.tile_codes && P.tile_codes[pd->tile_base + np * pd->S + ns] == 0)
                            np = 0xffffffffu;   // nothing to prefetch there
                        if (np != 0xffffffffu) issue_codes(cb ^ 1u, np, ns, nr0, nr1);
                    }
                    if (P.debug) dbg_reload_cyc += clock64() - dbg_r0;
                }
                if (c0 == 3u) break;                      // local alignment ends on a boundary cell (algo.rs:401-405)
                // every lane looks x steps ahead in the direction of c0; the run ends at the first different code
                const uint32_t x = (uint32_t)lane;
                const uint32_t di = (c0 != 1u) ? 1u : 0u, dj = (c0 != 2u) ? 1u : 0u;
                const bool reach = (x * di <= i) & (x * dj <= j);
                const uint32_t ci = i - (reach ? x * di : 0u), cj = j - (reach ? x * dj : 0u);
                uint32_t cx = cell_code(ci, cj);
                cx = (!reach || (x > 0 && ci == 0 && cj == 0)) ? 7u : cx;   // (0,0) is never emitted after a move (algo.rs:419-421)
                const uint32_t same = __ballot_sync(0xffffffffu, cx == c0);
                const uint32_t run = (same == 0xffffffffu) ? 32u : (uint32_t)(__ffs((int)~same) - 1);   // >= 1
                // the cell the walk reaches next is the one lane `run` just looked at: its code starts the next iteration
                const uint32_t c_next = __shfl_sync(0xffffffffu, cx, (int)(run & 31u));
                push(i, j, c0 | (run << 8));              // labels, counters and op stores happen in the emit warp
                // the checked_sub move (algo.rs:412-417); run cells are all inside the table
                const bool i_none = di && (i < run), j_none = dj && (j < run);
                if (i_none && j_none) break;
                i = i_none ? 0u : i - run * di;
                j = j_none ? 0u : j - run * dj;
                if (i == 0 && j == 0) break;
                // lane `run` looked at exactly (i, j) unless the run used all 32 lanes, the move was clamped, or that
                // cell was outside the window (7): then look it up (and reload the window at the top of the loop).
                // A warp-uniform branch: instructions issued are what an iteration costs.
                if (run < 32u && !i_none && !j_none && c_next != 7u) c0 = c_next;
                else c0 = cell_code(i, j);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (left_band && lane == 0) atomicAdd(P.left_band, 1u);
        if (lane == 0) {
            dbg[0] = dbg_iters | (dbg_reloads << 32);
            dbg[1] = (unsigned long long)(clock64() - dbg_t0);
            dbg[2] = (unsigned long long)dbg_reload_cyc;
        }
        __syncwarp();
        __threadfence_block();                            // debug counters before the end marker
        push(0u, 0u, 0xffffffffu);                        // end of path
    } else {
        // ================================================================ emit warp
        uint8_t *ops = P.ops + pd->ops_off;
        uint32_t nops = 0, n_match = 0, n_mis = 0, n_ext = 0, n_open = 0;
        uint32_t last = 0;  // AlignmentChoice::Match, algo.rs:338
        uint32_t end_i = i, end_j = j;
        if (i == 0 && j == 0) {
            // only possible for m == n == 0: one Match at (0,0) (None == None), then both checked_sub fail
            if (lane == 0) ops[0] = 0;
            nops = 1;
            n_match = 1;
        }
        uint32_t popped = 0;
        for (;;) {
            uint4 d;
            do {
                d = lds_volatile_uint4(ring + popped % WALK_RING);
            } while (d.w != popped + 1u || (d.x >> 29) != (((popped + 1u) >> 5) & 7u));
            d.x &= 0x1fffffffu;                           // row indices stay below 2^29 (gx_check_scores)
            __syncwarp();                                 // every lane has its copy: the slot may be reused
            popped++;
            if (lane == 0) sts_volatile_u32(tail, popped);
            if (d.z == 0xffffffffu) break;
            const uint32_t ri = d.x, rj = d.y, c0 = d.z & 3u, run = d.z >> 8;
            const uint32_t x = (uint32_t)lane;
            const bool diag = (c0 == 0u);
            const uint32_t di = (c0 != 1u) ? 1u : 0u, dj = (c0 != 2u) ? 1u : 0u;
            const bool mine = x < run;
            const uint32_t ci = ri - (mine ? x * di : 0u), cj = rj - (mine ? x * dj : 0u);
            // labels.  Diagonal: is_match(i, j), Option<u8> equality with None == None (sequence.rs:113-114): the characters
            // AFTER the cell's own (0-based s1[i], s2[j]; algo.rs:354).  Gaps: open/extend from last_choice (algo.rs:373-379, 388-394).
            const bool lab = diag & mine;
            const int a = (lab && ci < m) ? (int)__ldg(s1 + ci) : -1;
            const int b = (lab && cj < n) ? (int)__ldg(s2 + cj) : -1;
            const uint32_t mmask = __ballot_sync(0xffffffffu, lab && a == b);
            const uint32_t nm = (uint32_t)__popc(mmask);
            const uint32_t ext = (c0 == 1u) ? 2u : 3u;          // Insert / Delete
            const bool opens = !diag && (last != ext);
            const uint32_t gop = (x == 0 && opens) ? ext + 2u : ext;   // OpenInsert = 4, OpenDelete = 5
            const uint32_t op = diag ? ((a == b) ? 0u : 1u) : gop;
            GX_CHECK(P.check, !mine || (nops + x <= m + n && pd->ops_off + nops + x < P.ops_bytes), 22);
            if (mine) ops[nops + x] = (uint8_t)op;
            n_match += nm;
            n_mis += diag ? run - nm : 0u;
            n_open += opens ? 1u : 0u;
            n_ext += diag ? 0u : (opens ? run - 1u : run);
            last = diag ? 0u : ext;
            nops += run;
            end_i = ri - (run - 1u) * di;                 // last emitted cell
            end_j = rj - (run - 1u) * dj;
        }
        res.end_i = end_i;
        res.end_j = end_j;
        res.n_ops = nops;
        res.matches = n_match;
        res.mismatches = n_mis;
        res.gap_extensions = n_ext;
        res.opening_gaps = n_open;
        if (P.debug) {
            res.lcs_at_first_max = dbg[0];
            res.fill_ms = (double)dbg[1];
            res.walk_ms = (double)dbg[2];
        }
        if (lane == 0) P.results[q] = res;
    }
}

}  // namespace gx
