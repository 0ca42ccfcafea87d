// gx_fill.cuh -- K1/K2: anti-diagonal wavefront fill of the affine-gap S/D/I tables.
//
// Replaces the CPU fill of /root/reference/src/alignment/algo.rs:191-268 (alignment_table).
// Per cell the reference keeps six 8-byte lanes (algo.rs:25-35); here the state is E = V + (h+g)
// and D per column in registers, I runs along the row inside the thread, and the
// only thing written per cell is a 2-bit direction code (0 S, 1 I, 2 D -- the order retrace()
// tests them in, algo.rs:351-400).
//
// Decomposition
//   tile  = (panel p, strip s) of one pair: PANEL_H rows x (32*K) columns, owned by ONE WARP.
//   lane  = K consecutive columns, register-blocked; the 32 lanes run a systolic skew:
//           at step t lane l updates row t-l, so lane l needs from lane l-1 exactly the
//           (E,I) that lane l-1 produced one step earlier -> two __shfl_up per step.
//   strip -> strip hand-off (right boundary column, 8 B per row) goes through an L2-resident
//           buffer with an LL-style protocol: one 64-bit store carries E, I and a parity bit of
//           the current execute, the consumer polls the data itself, no flag, no fence.
//   panel -> panel hand-off (bottom row of a tile, (E,D) per column) goes through `top` with a
//           release/acquire counter per strip (once per 4096 rows).
//   tiles are taken from a global ticket counter in an order in which every dependency has a
//   smaller ticket, so a waiting warp only ever waits for a warp that is already resident.
//   s1 (the row sequence) is staged per tile into shared memory by a 1-D TMA bulk copy
//   (cp.async.bulk + mbarrier); s2 characters of the lane's K columns live in registers.
//
// Integer work per cell (global, score only): 2x VIADDMNMX (I, D), ISETP+SEL (match/mismatch),
// IADD (S), VIMNMX3 (V), IADD (E) = 7 -- the figure bench.py's roofline uses.
#pragma once
#include "gx_common.cuh"

namespace gx {

constexpr unsigned FULL = 0xffffffffu;
// A dependency wait that lasts longer than this many polls (>= ~1 s) is a bug or a lost device: raise the
// abort word instead of hanging the GPU; every waiter also leaves as soon as it sees the word set.
constexpr uint32_t SPIN_LIMIT = 1u << 23;

__device__ __forceinline__ bool spin_check(uint32_t &spins, uint32_t *abort_word) {
    if (++spins > SPIN_LIMIT) atomicExch(abort_word, 1u);
    if ((spins & 63u) == 0u || spins > SPIN_LIMIT) {
        if (*reinterpret_cast<volatile uint32_t *>(abort_word) != 0u) return true;
    }
    return false;
}

template <int K, bool LOCAL, bool CODES, int TRACK, bool MASKED, bool PAD>
__device__ __forceinline__ void run_block(int (&eu)[K], int (&du)[K], const int (&c2)[K], int &elast, int &ilast, int &vd,
                                          int &best, int &best_r, const int g, const int hg, const int ap, const int bp,
                                          const uint8_t *s1base /* s1 row 0 of this tile */, const uint2 *inring,
                                          uint2 *outring, uint4 *code_dst, const int t0, const int rows, const int lane,
                                          const int kvalid) {
    constexpr int SPC = 64 / K;       // steps per 16-byte code chunk
    constexpr int GROUPS = 32 / SPC;
    constexpr int KB = Log2<K>::value;
#pragma unroll 1
    for (int grp = 0; grp < GROUPS; ++grp) {
        uint32_t cw[4] = {0u, 0u, 0u, 0u};
        const uint2 *inr = inring + grp * SPC;
        const int tg = t0 + grp * SPC;
        const uint8_t *s1p = s1base + (tg - lane);
#pragma unroll
        for (int uu = 0; uu < SPC; ++uu) {
            const int r = tg + uu - lane;
            const uint2 bnd = inr[uu];
            int el = __shfl_up_sync(FULL, elast, 1);
            int il = __shfl_up_sync(FULL, ilast, 1);
            if (lane == 0) {
                el = (int)bnd.x;
                il = (int)bnd.y;
            }
            bool active = true;
            int c1;
            if (MASKED) {
                active = (r >= 0) && (r < rows);
                const int rc = min(max(r, 0), rows - 1);
                c1 = s1base[rc];
            } else {
                c1 = s1p[uu];
            }
            int e = el, irun = il, ed = vd;
            int rowbest = -1;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int In = LOCAL ? __viaddmax_s32_relu(irun, g, e) : __viaddmax_s32(irun, g, e);
                const int Dn = LOCAL ? __viaddmax_s32_relu(du[k], g, eu[k]) : __viaddmax_s32(du[k], g, eu[k]);
                const int Sn = ed + ((c1 == c2[k]) ? ap : bp);
                const int Vn = LOCAL ? __vimax3_s32_relu(In, Dn, Sn) : __vimax3_s32(In, Dn, Sn);
                if (CODES) {
                    const uint32_t code = (Sn == Vn) ? 0u : ((In == Vn) ? 1u : 2u);
                    constexpr int dummy = 0;
                    (void)dummy;
                    const int bitpos = uu * 2 * K + 2 * k;
                    cw[bitpos >> 5] |= code << (bitpos & 31);
                }
                ed = eu[k];
                const int En = Vn + hg;
                if (MASKED) {
                    eu[k] = active ? En : eu[k];
                    du[k] = active ? Dn : du[k];
                } else {
                    eu[k] = En;
                    du[k] = Dn;
                }
                e = En;
                irun = In;
                if (TRACK == 2) {
                    int key = (Vn << KB) | k;
                    if (PAD) key = (k < kvalid) ? key : -1;
                    rowbest = max(rowbest, key);
                } else if (TRACK == 1) {
                    int key = Vn;
                    if (PAD) key = (k < kvalid) ? key : -1;
                    rowbest = max(rowbest, key);
                }
            }
            if (MASKED) vd = active ? el : vd;
            else vd = el;
            elast = e;
            ilast = irun;
            if (TRACK == 2) {
                // last maximum in row-major order wins (Iterator::max_by, algo.rs:311-322): a later row
                // replaces an equal value; inside the row the key's low bits prefer the larger column.
                const bool upd = active && ((rowbest | (K - 1)) >= best);
                best = upd ? rowbest : best;
                best_r = upd ? r : best_r;
            } else if (TRACK == 1) {
                best = active ? max(best, rowbest) : best;
            }
            if (lane == 31 && active) outring[r & 31] = make_uint2((uint32_t)e, (uint32_t)irun);
        }
        if (CODES) st_cs_uint4(code_dst + grp * 32, make_uint4(cw[0], cw[1], cw[2], cw[3]));
    }
}

// steps of one tile with `rows` rows: rows + 31, rounded up to whole 32-step blocks
__host__ __device__ __forceinline__ uint32_t tile_blocks(uint32_t rows) { return (rows + 31 + 31) / 32; }

template <int K, bool LOCAL, bool CODES, int TRACK>
__global__ void __launch_bounds__(CTA_THREADS, (K <= 8) ? 3 : 2) gx_fill_kernel(const FillParams P) {
    constexpr int W = 32 * K;
    constexpr int SPC = 64 / K;
    constexpr int GROUPS = 32 / SPC;
    constexpr int KB = Log2<K>::value;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    uint8_t *wsm = smem + wib * WARP_SMEM;
    uint8_t *s1buf = wsm;
    uint2 *inring = reinterpret_cast<uint2 *>(wsm + WARP_SMEM_S1);
    uint2 *outring = inring + 32;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(outring + 32);

    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t phase = 0;
    const int g = P.g, hg = P.hg, ap = P.ap, bp = P.bp, h = P.h;
    const uint32_t parity = P.parity;
    uint32_t *abort_word = P.ticket + 1;
    bool dead = false;

    for (;;) {
        uint32_t tk = 0;
        if (lane == 0) tk = atomicAdd(P.ticket, 1u);
        tk = __shfl_sync(FULL, tk, 0);
        if (tk >= P.n_tiles) break;
        const TileDesc td = P.tiles[tk];
        const PairDesc *pd = P.pairs + td.pair;
        const int m = (int)pd->m, n = (int)pd->n, S = (int)pd->S;
        const int p = (int)td.p, s = (int)td.s;
        const int i0 = p << PANEL_H_LOG2;
        const int rows = min(PANEL_H, m - i0);
        const int jl = s * W + lane * K;  // columns jl+1 .. jl+K (1-based) belong to this lane
        const int kvalid = min(max(n - jl, 0), K);
        const bool has_pad = (s + 1) * W > n;

        // ---- stage s1[i0 .. i0+rows) with a TMA bulk copy (16-byte aligned window around it)
        const uint8_t *s1g = P.blob + pd->s1_off + i0;
        const uint32_t delta = (uint32_t)(reinterpret_cast<uintptr_t>(s1g) & 15u);
        __syncwarp();
        if (lane == 0) {
            fence_proxy_async_smem();  // earlier generic-proxy reads of s1buf are done (syncwarp above)
            const uint32_t bytes = (delta + (uint32_t)rows + 15u) & ~15u;
            mbar_expect_tx(mbar, bytes);
            tma_bulk_g2s(s1buf, s1g - delta, bytes, mbar);
        }

        // ---- s2 characters of this lane's columns; padding columns never match (256)
        int c2[K];
        {
            const uint8_t *s2g = P.blob + pd->s2_off + jl;
#pragma unroll
            for (int k = 0; k < K; ++k) c2[k] = (k < kvalid) ? (int)__ldg(s2g + k) : 256;
        }

        // ---- top boundary of the tile: (E,D) of row i0 for this lane's columns
        int eu[K], du[K];
        if (p == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                eu[k] = (LOCAL ? 0 : h + (jl + k + 1) * g) + hg;  // algo.rs:213-220 (row 0), V = insert_score
                du[k] = NEG32;
            }
        } else {
            // the previous panel of this strip must be complete (its `top` rows written)
            const uint32_t *pr = P.progress + pd->progress_off + s;
            uint32_t spins = 0;
            while (ld_acquire_u32(pr) < (uint32_t)p) {
                if (spin_check(spins, abort_word)) {
                    dead = true;
                    break;
                }
                __nanosleep(200);
            }
            if (dead) break;
            const int2 *tp = P.top + pd->top_off + jl;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int2 v = make_int2(NEG32, NEG32);
                if (k < kvalid) v = ld_cg_int2(tp + k);
                eu[k] = v.x;
                du[k] = v.y;
            }
        }
        const unsigned long long *cb_in = (s > 0) ? P.colbuf + pd->colbuf_off + (uint64_t)(s - 1) * m + i0 : nullptr;
        unsigned long long *cb_out = (s < S - 1) ? P.colbuf + pd->colbuf_off + (uint64_t)s * m + i0 : nullptr;

        // ---- diagonal seed: E of (row i0, column jl)
        int vd = __shfl_up_sync(FULL, eu[K - 1], 1);
        if (lane == 0) {
            if (s == 0) {
                vd = ((p == 0) ? 0 : (LOCAL ? 0 : h + i0 * g)) + hg;      // algo.rs:195-211 (column 0)
            } else if (p == 0) {
                vd = (LOCAL ? 0 : h + jl * g) + hg;                       // row 0
            } else {
                unsigned long long v;
                uint32_t spins = 0;
                while ((((uint32_t)((v = ld_relaxed_u64(cb_in - 1)) >> 32)) & 1u) != parity) {
                    if (spin_check(spins, abort_word)) {
                        dead = true;
                        break;
                    }
                    __nanosleep(100);
                }
                vd = (int)(uint32_t)v;
            }
        }
        dead = __any_sync(FULL, dead);
        if (dead) break;

        int elast = 0, ilast = 0;
        int best = -1, best_r = 0;
        const uint32_t nblk = tile_blocks((uint32_t)rows);
        uint4 *code_base = nullptr;
        if (CODES)
            code_base = reinterpret_cast<uint4 *>(P.codes + pd->codes_off + (uint64_t)(p * S + s) * pd->tile_code_bytes) + lane;

        // ---- left boundary prefetch (LL protocol): entry of local row 32*b + lane
        unsigned long long nxt = 0;
        auto issue = [&](uint32_t b) {
            const int r = (int)(32u * b) + lane;
            if (s > 0 && r < rows) nxt = ld_relaxed_u64(cb_in + r);
        };
        issue(0);

        // wait for the s1 segment
        {
            uint32_t tries = 0;
            while (!mbar_try_wait(mbar, phase)) {
                if (++tries > (1u << 24)) {
                    atomicExch(abort_word, 1u);
                    dead = true;
                    break;
                }
            }
            dead = __any_sync(FULL, dead);
            if (dead) break;
        }
        phase ^= 1u;
        const uint8_t *s1base = s1buf + delta;

        for (uint32_t b = 0; b < nblk; ++b) {
            // settle the in-ring of this block
            uint2 cur;
            const int rb = (int)(32u * b) + lane;
            if (s == 0) {
                cur.x = (uint32_t)((LOCAL ? 0 : h + (i0 + rb + 1) * g) + hg);  // algo.rs:204-211: V = delete_score
                cur.y = (uint32_t)NEG32;
            } else {
                const bool need = rb < rows;
                uint32_t spins = 0;
                for (;;) {
                    const bool ok = !need || ((((uint32_t)(nxt >> 32)) & 1u) == parity);
                    if (__all_sync(FULL, ok)) break;
                    if (__any_sync(FULL, spin_check(spins, abort_word))) {
                        dead = true;
                        break;
                    }
                    if (!ok) {
                        __nanosleep(64);
                        nxt = ld_relaxed_u64(cb_in + rb);
                    }
                }
                if (dead) break;
                cur.x = (uint32_t)nxt;
                cur.y = (uint32_t)(((int)(uint32_t)(nxt >> 32)) >> 1);
            }
            if (b + 1 < nblk) issue(b + 1);
            inring[lane] = cur;
            __syncwarp();

            const int t0 = (int)(32u * b);
            const bool full = (b >= 1) && (t0 + 31 <= rows - 1);
            uint4 *cdst = CODES ? code_base + (size_t)b * GROUPS * 32 : nullptr;
            if (full) {
                if ((TRACK != 0) && has_pad)
                    run_block<K, LOCAL, CODES, TRACK, false, true>(eu, du, c2, elast, ilast, vd, best, best_r, g, hg, ap, bp, s1base,
                                                                   inring, outring, cdst, t0, rows, lane, kvalid);
                else
                    run_block<K, LOCAL, CODES, TRACK, false, false>(eu, du, c2, elast, ilast, vd, best, best_r, g, hg, ap, bp,
                                                                    s1base, inring, outring, cdst, t0, rows, lane, kvalid);
            } else {
                run_block<K, LOCAL, CODES, TRACK, true, (TRACK != 0)>(eu, du, c2, elast, ilast, vd, best, best_r, g, hg, ap, bp, s1base,
                                                                      inring, outring, cdst, t0, rows, lane, kvalid);
            }

            // flush the right boundary rows lane 31 finished in this block: rows 32(b-1)+1 .. 32b
            __syncwarp();
            if (cb_out != nullptr) {
                const int ro = (lane == 0) ? t0 : t0 - 32 + lane;
                if (ro >= 0 && ro < rows) {
                    const uint2 v = outring[lane];
                    const unsigned long long packed =
                        (unsigned long long)v.x | ((unsigned long long)((v.y << 1) | parity) << 32);
                    st_relaxed_u64(cb_out + ro, packed);
                }
            }
            __syncwarp();
        }

        if (dead) break;
        // ---- bottom row -> top buffer (next panel of this strip, and the global score)
        {
            int2 *tp = P.top + pd->top_off + jl;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (k < kvalid) st_cg_int2(tp + k, make_int2(eu[k], du[k]));
            __syncwarp();
            if (lane == 0) st_release_u32(P.progress + pd->progress_off + s, (uint32_t)(p + 1));
        }

        // ---- tile-level local maximum (value, i, j): larger value, then larger i, then larger j
        if (TRACK != 0) {
            int bv, bi, bj;
            if (TRACK == 2) {
                bv = (best < 0) ? -1 : (best >> KB);
                bi = i0 + best_r + 1;
                bj = jl + (best & (K - 1)) + 1;
            } else {
                bv = best;
                bi = 0;
                bj = 0;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int ov = __shfl_xor_sync(FULL, bv, off);
                const int oi = __shfl_xor_sync(FULL, bi, off);
                const int oj = __shfl_xor_sync(FULL, bj, off);
                const bool take = (ov > bv) || (ov == bv && (oi > bi || (oi == bi && oj > bj)));
                bv = take ? ov : bv;
                bi = take ? oi : bi;
                bj = take ? oj : bj;
            }
            if (lane == 0) P.tile_best[pd->tile_base + p * S + s] = make_int4(bv, bi, bj, 0);
        }
    }
}

}  // namespace gx
