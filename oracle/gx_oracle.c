/*
 * gx_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the affine-gap NW/SW hot path of nlaha/genomics-rs
 * (src/alignment/algo.rs + src/sequence.rs:102-115).  It exists so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
 * check and time-compare the CUDA path.  Nothing under genomics_rs_b200/ may
 * import, link or call it; the product path has no CPU fallback.
 *
 * Parity pinning: the reference is Rust-nightly and cannot be compiled in this image
 * (no cargo/rustc), so the oracle is pinned against the reference's own golden
 * vectors, tests/test_alignment.rs:23-53, 55-90, 92-139 (all global mode).  Local
 * mode has NO reference test: for local mode this oracle is "parity unpinned" -- it
 * follows algo.rs line by line (citations below) but no reference-produced vector
 * exists to confirm it.
 *
 * Variants
 *   gxo_align_faithful  the reference's algorithm as written: (m+1)x(n+1) table of
 *                       48-byte cells, column-major, i-outer/j-inner, 4x score_max +
 *                       3x max_matches per cell, first-max tracking, stateless retrace.
 *   gxo_align_linear    same results from O(n) rolling int64 rows + 2-bit direction
 *                       codes; used where 48 B/cell cannot be allocated.
 *   gxo_score_linear    score (and local last-argmax) only, O(n) memory.
 *   gxo_score_batch     gxo_score_linear over a blob of pairs (pthread parallel-for).
 *   gxo_nw_score_blocked multi-threaded global score for very long pairs.
 *   gxo_nw_band         one column band of a global table with an explicit left boundary
 *                       column (checker of the multi-GPU column-band decomposition).
 *
 * Reference type/function map
 *   cell_t           <- AlignmentCell            algo.rs:25-35
 *   score_max        <- ComputeScore::score_max  algo.rs:98-107
 *   max_matches      <- ComputeScore::max_matches algo.rs:112-121
 *   is_match         <- SequenceOperations::is_match (reverse=false) sequence.rs:102-115
 *   fill loop        <- alignment_table          algo.rs:151-282
 *   walk             <- retrace                  algo.rs:287-441
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <pthread.h>

/* AlignmentChoice discriminants, algo.rs:124-133 (#[repr(u8)]) */
enum { GXO_MATCH = 0, GXO_MISMATCH = 1, GXO_INSERT = 2, GXO_DELETE = 3, GXO_OPEN_INSERT = 4, GXO_OPEN_DELETE = 5 };

typedef struct {
    int64_t score;
    uint64_t start_i, start_j, end_i, end_j;
    uint64_t n_ops, matches, mismatches, gap_extensions, opening_gaps;
    uint64_t lcs_at_first_max;       /* 2nd return value of alignment_table, algo.rs:279-281 */
    uint64_t first_max_i, first_max_j; /* max_cell, algo.rs:157-158,258-262 */
    double fill_ms, walk_ms;
    int32_t status;                  /* 0 ok; 1 ops capacity too small; 2 alloc failure; 3 impossible state (algo.rs:407-408) */
    int32_t pad;
} gxo_result;

typedef struct { /* AlignmentCell, #[repr(C)], 48 bytes */
    int64_t ins, del, sub;
    uint64_t mi, md, ms;
} cell_t;

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* release-mode i64 '+' wraps (algo.rs is built with opt-level 3, overflow checks off) */
static inline int64_t wadd(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t max2(int64_t a, int64_t b) { return a > b ? a : b; }

/* algo.rs:98-107 */
static inline int64_t score_max(const cell_t *c, int64_t im, int64_t sm, int64_t dm, int is_local) {
    int64_t f = is_local ? 0 : INT64_MIN;
    return max2(max2(wadd(c->ins, im), wadd(c->sub, sm)), max2(wadd(c->del, dm), f));
}
/* algo.rs:112-121 */
static inline uint64_t max_matches(const cell_t *c) {
    uint64_t v = c->mi > c->ms ? c->mi : c->ms;
    return v > c->md ? v : c->md;
}
/* sequence.rs:102-115 with reverse_sequences=false: Option<u8> equality, None==None is true */
static inline int is_match(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n, uint64_t i, uint64_t j) {
    int a = i < m ? (int)s1[i] : -1;
    int b = j < n ? (int)s2[j] : -1;
    return a == b;
}

/* FNV-1a-64 over the walk, per op the 8 LE bytes of code*1000003 + i*10007 + j */
static inline uint64_t hash_step(uint64_t hsh, uint64_t code, uint64_t i, uint64_t j) {
    uint64_t v = code * 1000003ull + i * 10007ull + j;
    for (int b = 0; b < 8; b++) {
        hsh ^= (v >> (8 * b)) & 0xff;
        hsh *= 1099511628211ull;
    }
    return hsh;
}
#define FNV_OFFSET 1469598103934665603ull

uint64_t gxo_hash_ops(const uint8_t *ops, uint64_t n_ops, uint64_t start_i, uint64_t start_j) {
    /* replays (choice,i,j) from the start cell with the checked_sub rules of algo.rs:412-417 */
    uint64_t h = FNV_OFFSET, i = start_i, j = start_j;
    for (uint64_t k = 0; k < n_ops; k++) {
        uint8_t c = ops[k];
        h = hash_step(h, c, i, j);
        if (c == GXO_MATCH || c == GXO_MISMATCH) { i = i ? i - 1 : 0; j = j ? j - 1 : 0; }
        else if (c == GXO_INSERT || c == GXO_OPEN_INSERT) { j = j ? j - 1 : 0; }
        else { i = i ? i - 1 : 0; }
    }
    return h;
}

/* ------------------------------------------------------------------------------------------ */
/* Faithful variant                                                                           */
/* ------------------------------------------------------------------------------------------ */
int gxo_align_faithful(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n,
                       int64_t a, int64_t b, int64_t g, int64_t h, int is_local,
                       gxo_result *out, uint8_t *ops, uint32_t *ops_i, uint32_t *ops_j, uint64_t ops_cap) {
    memset(out, 0, sizeof(*out));
    const uint64_t R = m + 1, C = n + 1;
    /* Array2::zeros((m+1,n+1).f()), algo.rs:172: zero-filled, column-major */
    cell_t *T = (cell_t *)calloc(R * C, sizeof(cell_t));
    if (!T) { out->status = 2; return 2; }
#define AT(i, j) T[(uint64_t)(i) + (uint64_t)(j) * R]
    int64_t gh = wadd(g, h);
    int64_t neg_inf = wadd(INT64_MIN, gh < 0 ? -gh : gh); /* algo.rs:166 */
    int64_t maximum_score = INT64_MIN;                  /* algo.rs:157 */
    uint64_t max_i = 0, max_j = 0;                      /* algo.rs:158 */

    double t0 = now_ms();
    for (uint64_t i = 0; i < R; i++) {       /* algo.rs:191 */
        for (uint64_t j = 0; j < C; j++) {   /* algo.rs:192 */
            cell_t c;
            if (i == 0 && j == 0) {          /* algo.rs:195-202 */
                memset(&c, 0, sizeof c);
            } else if (j == 0) {             /* algo.rs:204-211 */
                c.ins = neg_inf; c.del = wadd(h, (int64_t)i * g); c.sub = neg_inf; c.mi = c.md = c.ms = 0;
            } else if (i == 0) {             /* algo.rs:213-220 */
                c.ins = wadd(h, (int64_t)j * g); c.del = neg_inf; c.sub = neg_inf; c.mi = c.md = c.ms = 0;
            } else {                         /* algo.rs:221-265 */
                const cell_t *top_left = &AT(i - 1, j - 1);
                const cell_t *left = &AT(i - 1, j); /* the reference's name for T[i-1][j] */
                const cell_t *top = &AT(i, j - 1);  /* the reference's name for T[i][j-1] */
                int eq = is_match(s1, m, s2, n, i - 1, j - 1);
                c.ins = score_max(top, g, gh, gh, is_local);
                c.del = score_max(left, gh, gh, g, is_local);
                c.sub = wadd(eq ? a : b, score_max(top_left, 0, 0, 0, is_local));
                c.mi = max_matches(top);
                c.md = max_matches(left);
                c.ms = max_matches(top_left) + (eq ? 1 : 0);
                int64_t mcs = score_max(&c, 0, 0, 0, is_local);
                if (maximum_score < mcs) { max_i = i; max_j = j; maximum_score = mcs; }
            }
            AT(i, j) = c;
        }
    }
    out->fill_ms = now_ms() - t0;
    out->first_max_i = max_i; out->first_max_j = max_j;
    out->lcs_at_first_max = max_matches(&AT(max_i, max_j)); /* algo.rs:279 */

    /* ---- retrace, algo.rs:287-441 ---- */
    t0 = now_ms();
    uint64_t i = m, j = n; /* algo.rs:308 */
    if (is_local) {
        /* indexed_iter() is logical row-major; Iterator::max_by keeps the LAST maximum. algo.rs:311-322 */
        int64_t best = INT64_MIN; int have = 0;
        for (uint64_t ii = 0; ii < R; ii++)
            for (uint64_t jj = 0; jj < C; jj++) {
                int64_t v = score_max(&AT(ii, jj), 0, 0, 0, 1);
                if (!have || v >= best) { best = v; i = ii; j = jj; have = 1; }
            }
    }
    out->start_i = i; out->start_j = j;
    out->score = score_max(&AT(i, j), 0, 0, 0, is_local); /* algo.rs:331 */
    out->end_i = i; out->end_j = j;
    int last = GXO_MATCH; /* algo.rs:338 */
    uint64_t k = 0;
    int status = 0;
    for (;;) {
        const cell_t *c = &AT(i, j);
        int64_t mx = score_max(c, 0, 0, 0, is_local);
        int code;
        int i_none = 0, j_none = 0;
        uint64_t ni = i, nj = j;
        if (mx == c->sub) {                          /* algo.rs:353-369 */
            if (is_match(s1, m, s2, n, i, j)) { code = GXO_MATCH; out->matches++; }
            else { code = GXO_MISMATCH; out->mismatches++; }
            last = code;
            if (i == 0) i_none = 1; else ni = i - 1;
            if (j == 0) j_none = 1; else nj = j - 1;
        } else if (mx == c->ins) {                   /* algo.rs:371-385 */
            if (last == GXO_INSERT) { out->gap_extensions++; code = GXO_INSERT; }
            else { out->opening_gaps++; code = GXO_OPEN_INSERT; }
            last = GXO_INSERT;
            if (j == 0) j_none = 1; else nj = j - 1;
        } else if (mx == c->del) {                   /* algo.rs:387-400 */
            if (last == GXO_DELETE) { out->gap_extensions++; code = GXO_DELETE; }
            else { out->opening_gaps++; code = GXO_OPEN_DELETE; }
            last = GXO_DELETE;
            if (i == 0) i_none = 1; else ni = i - 1;
        } else {                                     /* algo.rs:401-409 */
            if (!(is_local && mx == 0)) status = 3;
            break;
        }
        if (k >= ops_cap && ops) { status = 1; break; }
        if (ops) ops[k] = (uint8_t)code;
        if (ops_i) ops_i[k] = (uint32_t)i;
        if (ops_j) ops_j[k] = (uint32_t)j;
        out->end_i = i; out->end_j = j;
        k++;
        /* algo.rs:412-417 */
        if (i_none && j_none) break;
        i = i_none ? 0 : ni;
        j = j_none ? 0 : nj;
        if (i == 0 && j == 0) break;                 /* algo.rs:419-421 */
    }
    out->n_ops = k;
    out->walk_ms = now_ms() - t0;
    out->status = status;
    free(T);
#undef AT
    return status;
}

/* ------------------------------------------------------------------------------------------ */
/* Linear-memory variant (rolling int64 rows + 2-bit codes), SURVEY 3.4                        */
/* valid when h <= 0, g < 0, h+g < 0 (checked by caller); bit-identical results then.          */
/* ------------------------------------------------------------------------------------------ */
#define NEG64 (INT64_MIN / 4)

static int compact_walk(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n, int is_local,
                        const uint8_t *codes, uint64_t code_stride /* bytes per row */,
                        uint64_t si, uint64_t sj, gxo_result *out,
                        uint8_t *ops, uint32_t *ops_i, uint32_t *ops_j, uint64_t ops_cap) {
    uint64_t i = si, j = sj, k = 0;
    int last = GXO_MATCH, status = 0;
    out->end_i = i; out->end_j = j;
    for (;;) {
        int c;
        if (i == 0 && j == 0) c = 0;
        else if (j == 0) c = is_local ? 3 : 2;
        else if (i == 0) c = is_local ? 3 : 1;
        else c = (codes[(i - 1) * code_stride + ((j - 1) >> 2)] >> (2 * ((j - 1) & 3))) & 3;
        if (c == 3) break;
        int code;
        int i_none = 0, j_none = 0;
        uint64_t ni = i, nj = j;
        if (c == 0) {
            if (is_match(s1, m, s2, n, i, j)) { code = GXO_MATCH; out->matches++; }
            else { code = GXO_MISMATCH; out->mismatches++; }
            last = code;
            if (i == 0) i_none = 1; else ni = i - 1;
            if (j == 0) j_none = 1; else nj = j - 1;
        } else if (c == 1) {
            if (last == GXO_INSERT) { out->gap_extensions++; code = GXO_INSERT; }
            else { out->opening_gaps++; code = GXO_OPEN_INSERT; }
            last = GXO_INSERT;
            if (j == 0) j_none = 1; else nj = j - 1;
        } else {
            if (last == GXO_DELETE) { out->gap_extensions++; code = GXO_DELETE; }
            else { out->opening_gaps++; code = GXO_OPEN_DELETE; }
            last = GXO_DELETE;
            if (i == 0) i_none = 1; else ni = i - 1;
        }
        if (k >= ops_cap && ops) { status = 1; break; }
        if (ops) ops[k] = (uint8_t)code;
        if (ops_i) ops_i[k] = (uint32_t)i;
        if (ops_j) ops_j[k] = (uint32_t)j;
        out->end_i = i; out->end_j = j;
        k++;
        if (i_none && j_none) break;
        i = i_none ? 0 : ni;
        j = j_none ? 0 : nj;
        if (i == 0 && j == 0) break;
    }
    out->n_ops = k;
    return status;
}

int gxo_align_linear(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n,
                     int64_t a, int64_t b, int64_t g, int64_t h, int is_local, int want_traceback,
                     gxo_result *out, uint8_t *ops, uint32_t *ops_i, uint32_t *ops_j, uint64_t ops_cap) {
    memset(out, 0, sizeof(*out));
    if (!(h <= 0 && g < 0 && h + g < 0)) { out->status = 4; return 4; }
    const int64_t hg = h + g;
    int64_t *V = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    int64_t *D = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    uint64_t *L = (uint64_t *)calloc(n + 1, sizeof(uint64_t)); /* LCS lane, algo.rs:250-255 */
    uint64_t stride = (n + 3) / 4;
    uint8_t *codes = NULL;
    if (want_traceback && m && n) {
        codes = (uint8_t *)calloc(m * stride, 1);
        if (!codes) { free(V); free(D); free(L); out->status = 2; return 2; }
    }
    if (!V || !D || !L) { free(V); free(D); free(L); free(codes); out->status = 2; return 2; }
    double t0 = now_ms();
    V[0] = 0; D[0] = NEG64;
    for (uint64_t j = 1; j <= n; j++) { V[j] = is_local ? 0 : h + (int64_t)j * g; D[j] = NEG64; }
    int64_t best = 0; uint64_t bi = m, bj = n;   /* local: boundaries count with V=0; last cell in row-major order wins ties */
    int64_t fmax = INT64_MIN; uint64_t fi = 0, fj = 0, flcs = 0;
    if (is_local) { bi = 0; bj = n; } /* row 0 is all V=0: its last cell is the running last-argmax */
    for (uint64_t i = 1; i <= m; i++) {
        int64_t vdiag = V[0];
        uint64_t ldiag = L[0];
        V[0] = is_local ? 0 : h + (int64_t)i * g;
        int64_t I = NEG64;
        int64_t vleft = V[0];
        uint64_t lleft = 0;
        uint8_t c1 = s1[i - 1];
        uint8_t *crow = codes ? codes + (i - 1) * stride : NULL;
        if (is_local && 0 >= best) { best = 0; bi = i; bj = 0; }
        for (uint64_t j = 1; j <= n; j++) {
            int eq = (c1 == s2[j - 1]);
            int64_t In = max2(I + g, vleft + hg);
            int64_t Dn = max2(D[j] + g, V[j] + hg);
            int64_t Sn = vdiag + (eq ? a : b);
            if (is_local) { In = max2(In, 0); Dn = max2(Dn, 0); }
            int64_t Vn = max2(max2(In, Dn), Sn);
            if (is_local) Vn = max2(Vn, 0);
            /* LCS lane */
            uint64_t lup = L[j];
            uint64_t ln = ldiag + (eq ? 1 : 0);
            if (lleft > ln) ln = lleft;
            if (lup > ln) ln = lup;
            if (fmax < Vn) { fmax = Vn; fi = i; fj = j; flcs = ln; }
            if (is_local && Vn >= best) { best = Vn; bi = i; bj = j; }
            if (crow) {
                int code = (Sn == Vn) ? 0 : (In == Vn) ? 1 : 2;
                crow[(j - 1) >> 2] |= (uint8_t)(code << (2 * ((j - 1) & 3)));
            }
            vdiag = V[j]; ldiag = lup;
            V[j] = Vn; D[j] = Dn; L[j] = ln;
            I = In; vleft = Vn; lleft = ln;
        }
    }
    out->fill_ms = now_ms() - t0;
    out->first_max_i = fi; out->first_max_j = fj; out->lcs_at_first_max = flcs;
    if (is_local) { out->start_i = bi; out->start_j = bj; out->score = best; }
    else { out->start_i = m; out->start_j = n; out->score = V[n]; if (m == 0) out->score = (n == 0) ? 0 : h + (int64_t)n * g; }
    if (!is_local && n == 0 && m > 0) out->score = h + (int64_t)m * g;
    out->end_i = out->start_i; out->end_j = out->start_j;
    int status = 0;
    if (want_traceback) {
        t0 = now_ms();
        status = compact_walk(s1, m, s2, n, is_local, codes, stride, out->start_i, out->start_j, out, ops, ops_i, ops_j, ops_cap);
        out->walk_ms = now_ms() - t0;
    }
    out->status = status;
    free(V); free(D); free(L); free(codes);
    return status;
}

/* Score only; local also returns the last-argmax start cell.  O(n) memory, int64. */
int gxo_score_linear(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n,
                     int64_t a, int64_t b, int64_t g, int64_t h, int is_local,
                     int64_t *score, uint64_t *start_i, uint64_t *start_j) {
    if (!(h <= 0 && g < 0 && h + g < 0)) return 4;
    const int64_t hg = h + g;
    int64_t *V = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    int64_t *D = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    if (!V || !D) { free(V); free(D); return 2; }
    V[0] = 0; D[0] = NEG64;
    for (uint64_t j = 1; j <= n; j++) { V[j] = is_local ? 0 : h + (int64_t)j * g; D[j] = NEG64; }
    int64_t best = 0; uint64_t bi = 0, bj = n;
    for (uint64_t i = 1; i <= m; i++) {
        int64_t vdiag = V[0];
        V[0] = is_local ? 0 : h + (int64_t)i * g;
        int64_t I = NEG64, vleft = V[0];
        uint8_t c1 = s1[i - 1];
        if (is_local && 0 >= best) { best = 0; bi = i; bj = 0; }
        for (uint64_t j = 1; j <= n; j++) {
            int64_t In = max2(I + g, vleft + hg);
            int64_t Dn = max2(D[j] + g, V[j] + hg);
            int64_t Sn = vdiag + ((c1 == s2[j - 1]) ? a : b);
            int64_t Vn = max2(max2(In, Dn), Sn);
            if (is_local) { In = max2(In, 0); Dn = max2(Dn, 0); Vn = max2(Vn, 0); if (Vn >= best) { best = Vn; bi = i; bj = j; } }
            vdiag = V[j]; V[j] = Vn; D[j] = Dn; I = In; vleft = Vn;
        }
    }
    if (is_local) { *score = best; if (start_i) *start_i = bi; if (start_j) *start_j = bj; }
    else { *score = V[n]; if (n == 0 && m > 0) *score = h + (int64_t)m * g; if (start_i) *start_i = m; if (start_j) *start_j = n; }
    free(V); free(D);
    return 0;
}

/* tiny pthread parallel-for (this image's gcc has no usable libgomp spec) */
typedef void (*pf_body)(int64_t idx, void *ctx);
typedef struct { pf_body body; void *ctx; int64_t lo, hi, chunk; volatile int64_t *next; } pf_arg;
static void *pf_worker(void *p) {
    pf_arg *a = (pf_arg *)p;
    for (;;) {
        int64_t s = __atomic_fetch_add(a->next, a->chunk, __ATOMIC_RELAXED);
        if (s >= a->hi) break;
        int64_t e = s + a->chunk < a->hi ? s + a->chunk : a->hi;
        for (int64_t i = s; i < e; i++) a->body(i, a->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t lo, int64_t hi, int64_t chunk, int n_threads, pf_body body, void *ctx) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    volatile int64_t next = lo;
    pf_arg a = { body, ctx, lo, hi, chunk, &next };
    if (n_threads == 1 || hi - lo <= chunk) { pf_worker(&a); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &a) == 0) started++;
    pf_worker(&a);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

typedef struct {
    const uint8_t *blob; const uint64_t *off1, *len1, *off2, *len2;
    int64_t a, b, g, h; int is_local; int64_t *scores; int rc;
} batch_ctx;
static void batch_body(int64_t p, void *vc) {
    batch_ctx *c = (batch_ctx *)vc;
    int r = gxo_score_linear(c->blob + c->off1[p], c->len1[p], c->blob + c->off2[p], c->len2[p],
                             c->a, c->b, c->g, c->h, c->is_local, &c->scores[p], NULL, NULL);
    if (r) c->rc = r;
}
int gxo_score_batch(const uint8_t *blob, const uint64_t *off1, const uint64_t *len1,
                    const uint64_t *off2, const uint64_t *len2, uint64_t n_pairs,
                    int64_t a, int64_t b, int64_t g, int64_t h, int is_local, int n_threads,
                    int64_t *scores) {
    batch_ctx c = { blob, off1, len1, off2, len2, a, b, g, h, is_local, scores, 0 };
    parallel_for(0, (int64_t)n_pairs, 64, n_threads, batch_body, &c);
    return c.rc;
}

/* Multi-threaded global score for very long pairs: column blocks pipelined over row blocks
 * (block (r,c) needs (r,c-1) and (r-1,c)); threads sweep anti-diagonals of blocks.  Exact int64.
 * Used once to freeze the config-5 constant; not a reference capability (48 TB table there). */
typedef struct {
    const uint8_t *s1, *s2; uint64_t m, n; int64_t a, b, g, hg; uint64_t blk, nbc;
    int64_t *Vr, *Dr, *Vc, *Ic, *corner; uint64_t d;
} blk_ctx;
static void blk_body(int64_t r, void *vc) {
    blk_ctx *x = (blk_ctx *)vc;
    const uint64_t blk = x->blk, nbc = x->nbc, m = x->m, n = x->n;
    const int64_t a = x->a, b = x->b, g = x->g, hg = x->hg;
    int64_t *Vr = x->Vr, *Dr = x->Dr, *Vc = x->Vc, *Ic = x->Ic;
    uint64_t c = x->d - (uint64_t)r;
    uint64_t i0 = (uint64_t)r * blk + 1, i1 = i0 + blk - 1; if (i1 > m) i1 = m;
    uint64_t j0 = c * blk + 1, j1 = j0 + blk - 1; if (j1 > n) j1 = n;
    int64_t vd0 = x->corner[(uint64_t)r * (nbc + 1) + c]; /* V[i0-1][j0-1] */
    for (uint64_t i = i0; i <= i1; i++) {
        int64_t vdiag = vd0;
        int64_t vleft = Vc[i], I = Ic[i];
        vd0 = vleft; /* V[i][j0-1] is the diagonal of the next row's first column */
        uint8_t c1 = x->s1[i - 1];
        for (uint64_t j = j0; j <= j1; j++) {
            int64_t In = max2(I + g, vleft + hg);
            int64_t Dn = max2(Dr[j] + g, Vr[j] + hg);
            int64_t Sn = vdiag + ((c1 == x->s2[j - 1]) ? a : b);
            int64_t Vn = max2(max2(In, Dn), Sn);
            vdiag = Vr[j]; Vr[j] = Vn; Dr[j] = Dn; I = In; vleft = Vn;
        }
        Vc[i] = vleft; Ic[i] = I;
    }
    x->corner[((uint64_t)r + 1) * (nbc + 1) + (c + 1)] = Vr[j1];
}
int gxo_nw_score_blocked(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n,
                         int64_t a, int64_t b, int64_t g, int64_t h, int n_threads, uint64_t blk, int64_t *score) {
    if (!(h <= 0 && g < 0 && h + g < 0)) return 4;
    if (m == 0 || n == 0) { *score = (m == 0 && n == 0) ? 0 : h + (int64_t)(m + n) * g; return 0; }
    const int64_t hg = h + g;
    if (blk == 0) blk = 4096;
    uint64_t nbr = (m + blk - 1) / blk, nbc = (n + blk - 1) / blk;
    /* row state per column (V,D) shared by all row-blocks in a column block; column state (V,I) per row */
    int64_t *Vr = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    int64_t *Dr = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    int64_t *Vc = (int64_t *)malloc((m + 1) * sizeof(int64_t));
    int64_t *Ic = (int64_t *)malloc((m + 1) * sizeof(int64_t));
    int64_t *corner = (int64_t *)malloc((nbr + 1) * (nbc + 1) * sizeof(int64_t)); /* V at block top-left corners */
    if (!Vr || !Dr || !Vc || !Ic || !corner) { free(Vr); free(Dr); free(Vc); free(Ic); free(corner); return 2; }
    Vr[0] = 0;
    for (uint64_t j = 1; j <= n; j++) { Vr[j] = h + (int64_t)j * g; Dr[j] = NEG64; }
    Vc[0] = 0;
    for (uint64_t i = 1; i <= m; i++) { Vc[i] = h + (int64_t)i * g; Ic[i] = NEG64; }
    /* corner[r][c] = V[r*blk][c*blk] */
    for (uint64_t c = 0; c <= nbc; c++) { uint64_t j = c * blk; if (j > n) j = n; corner[c] = Vr[j]; }
    for (uint64_t r = 0; r <= nbr; r++) { uint64_t i = r * blk; if (i > m) i = m; corner[r * (nbc + 1)] = Vc[i]; }
    blk_ctx bc = { s1, s2, m, n, a, b, g, hg, blk, nbc, Vr, Dr, Vc, Ic, corner, 0 };
    for (uint64_t d = 0; d < nbr + nbc - 1; d++) {
        int64_t rlo = (d >= nbc) ? (int64_t)(d - nbc + 1) : 0;
        int64_t rhi = (d < nbr) ? (int64_t)d : (int64_t)nbr - 1;
        bc.d = d;
        parallel_for(rlo, rhi + 1, 1, n_threads, blk_body, &bc);
    }
    *score = Vr[n];
    free(Vr); free(Dr); free(Vc); free(Ic); free(corner);
    return 0;
}

/* One column band of the global table: columns col0+1 .. col0+nb of s2 (s2band = s2 + col0), all m rows.
 * Same recurrences as gxo_score_linear (algo.rs:221-265).  The left boundary is column col0 of the full
 * table: for col0 == 0 the column-0 formulas of algo.rs:195-211 (in_V/in_I ignored, may be NULL), otherwise
 * in_V[i-1], in_I[i-1] = V and insert_score of cell (i, col0), i = 1..m, as produced by the band to the left.
 * Row 0 is algo.rs:213-220 at absolute column col0+j.  out_V/out_I (m entries, may be NULL) receive column
 * col0+nb; *score = V[m][col0+nb].  Exact int64. */
int gxo_nw_band(const uint8_t *s1, uint64_t m, const uint8_t *s2band, uint64_t nb, uint64_t col0,
                int64_t a, int64_t b, int64_t g, int64_t h,
                const int64_t *in_V, const int64_t *in_I, int64_t *out_V, int64_t *out_I, int64_t *score) {
    if (!(h <= 0 && g < 0 && h + g < 0)) return 4;
    if (col0 > 0 && (!in_V || !in_I) && m > 0) return 5;
    const int64_t hg = h + g;
    int64_t *V = (int64_t *)malloc((nb + 1) * sizeof(int64_t));
    int64_t *D = (int64_t *)malloc((nb + 1) * sizeof(int64_t));
    if (!V || !D) { free(V); free(D); return 2; }
    V[0] = col0 ? h + (int64_t)col0 * g : 0;                       /* V[0][col0] */
    D[0] = NEG64;
    for (uint64_t j = 1; j <= nb; j++) { V[j] = h + (int64_t)(col0 + j) * g; D[j] = NEG64; }
    for (uint64_t i = 1; i <= m; i++) {
        int64_t vdiag = V[0];
        int64_t I;
        if (col0 == 0) { V[0] = h + (int64_t)i * g; I = NEG64; }    /* algo.rs:204-211 */
        else { V[0] = in_V[i - 1]; I = in_I[i - 1]; }
        int64_t vleft = V[0];
        uint8_t c1 = s1[i - 1];
        for (uint64_t j = 1; j <= nb; j++) {
            int64_t In = max2(I + g, vleft + hg);
            int64_t Dn = max2(D[j] + g, V[j] + hg);
            int64_t Sn = vdiag + ((c1 == s2band[j - 1]) ? a : b);
            int64_t Vn = max2(max2(In, Dn), Sn);
            vdiag = V[j]; V[j] = Vn; D[j] = Dn; I = In; vleft = Vn;
        }
        if (out_V) out_V[i - 1] = vleft;
        if (out_I) out_I[i - 1] = I;
    }
    if (score) *score = V[nb];
    free(V); free(D);
    return 0;
}

/* The three score planes of the reference's table (what print_scores_table prints, display.rs:190-220), row-major
 * (m+1) x (n+1) int64 each: insert_score, delete_score, sub_score exactly as alignment_table stores them
 * (algo.rs:195-248), boundary "minus infinity" = i64::MIN + |g+h| (algo.rs:166). */
int gxo_planes(const uint8_t *s1, uint64_t m, const uint8_t *s2, uint64_t n,
               int64_t a, int64_t b, int64_t g, int64_t h, int is_local, int64_t *pi, int64_t *pd, int64_t *ps) {
    const uint64_t C = n + 1;
    const int64_t gh = wadd(g, h);
    const int64_t neg_inf = wadd(INT64_MIN, gh < 0 ? -gh : gh);
    for (uint64_t i = 0; i <= m; i++)
        for (uint64_t j = 0; j <= n; j++) {
            cell_t c;
            memset(&c, 0, sizeof c);
            if (i == 0 && j == 0) {
            } else if (j == 0) { c.ins = neg_inf; c.del = wadd(h, (int64_t)i * g); c.sub = neg_inf; }
            else if (i == 0) { c.ins = wadd(h, (int64_t)j * g); c.del = neg_inf; c.sub = neg_inf; }
            else {
                cell_t tl = { pi[(i - 1) * C + j - 1], pd[(i - 1) * C + j - 1], ps[(i - 1) * C + j - 1], 0, 0, 0 };
                cell_t left = { pi[(i - 1) * C + j], pd[(i - 1) * C + j], ps[(i - 1) * C + j], 0, 0, 0 };
                cell_t top = { pi[i * C + j - 1], pd[i * C + j - 1], ps[i * C + j - 1], 0, 0, 0 };
                int eq = is_match(s1, m, s2, n, i - 1, j - 1);
                c.ins = score_max(&top, g, gh, gh, is_local);
                c.del = score_max(&left, gh, gh, g, is_local);
                c.sub = wadd(eq ? a : b, score_max(&tl, 0, 0, 0, is_local));
            }
            pi[i * C + j] = c.ins; pd[i * C + j] = c.del; ps[i * C + j] = c.sub;
        }
    return 0;
}

uint64_t gxo_sizeof_result(void) { return sizeof(gxo_result); }
