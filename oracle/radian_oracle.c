/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, float64, log domain) of the
 * comprna/radian hot path.  It is the checker for the CUDA path, never a product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (radian_b200/csrc) shares no code with it and
 * uses a different formulation (linear domain, power-of-two rescaling, lane-parallel
 * selection), so an agreement between the two is a real cross-check.
 *
 * Parity pinning: the reference has no tests or golden vectors.  This
 * file is pinned against outputs of the *unmodified* reference run in the build
 * container (oracle/make_golden.py -> tests/golden/ npz files) by tests/test_oracle_golden.py.
 *
 * Reference lines followed (paths relative to /root/reference/radian):
 *   decode.py:16-17    log()                 -> ref_log
 *   decode.py:52-64    combine_dists         -> combine
 *   decode.py:67-76    normalise / entropy   -> frame_entropy_*, row_entropy
 *   decode.py:79-96    apply_rna_model       -> gate logic inside search()
 *   decode.py:124-210  beam_search           -> radian_oracle_beam_search
 *   matrix_assembly.py:12-53                 -> radian_oracle_assemble
 * Third-party arithmetic restated here: numpy.logaddexp (npy_logaddexp, numpy 2.3.5 in
 * the container, pinned ~=1.19.5 by requirements.txt:5) and sklearn normalize(norm="l1")
 * incl. _handle_zeros_in_scale (scikit-learn 1.9.0 here, ~=1.1.2 pinned).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_BASES 4
#define BLANK 4

/* decode.py:16-17 */
static inline double ref_log(double x) { return x == 0.0 ? -INFINITY : log(x); }

/* numpy npy_logaddexp (npymath): log(exp(x)+exp(y)) */
static inline double lae(double x, double y)
{
    if (x == y) return x + 0.693147180559945309417232121458176568; /* same-sign inf too */
    double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp; /* NaN */
}

/* ------------------------------------------------------------------ trie of labelings */
typedef struct {
    int32_t parent;
    int32_t child[N_BASES];
    int32_t len;      /* number of symbols */
    uint32_t ctx;     /* last 16 symbols, 2 bits each, newest in the low bits */
    int8_t sym;       /* last symbol, -1 for the root */
    double h_ext;     /* memo of entropy(lm[last L symbols]) (decode.py:86-90); NaN = unset */
} Node;

typedef struct {
    Node *v;
    int64_t n, cap;
} Trie;

static int32_t trie_new(Trie *t, int32_t parent, int sym)
{
    if (t->n == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 1024;
        t->v = (Node *)realloc(t->v, (size_t)t->cap * sizeof(Node));
    }
    Node *nd = &t->v[t->n];
    nd->parent = parent;
    for (int c = 0; c < N_BASES; ++c) nd->child[c] = -1;
    nd->sym = (int8_t)sym;
    nd->h_ext = NAN;
    if (parent < 0) {
        nd->len = 0;
        nd->ctx = 0;
    } else {
        nd->len = t->v[parent].len + 1;
        nd->ctx = (t->v[parent].ctx << 2) | (uint32_t)sym;
    }
    return (int32_t)(t->n++);
}

/* one BeamEntry (decode.py:20-26).  node >= 0: labeling is that trie node.
 * node < 0: labeling = labeling(parent) + (sym,), not materialised unless it survives. */
typedef struct {
    double ptot, pnb, pb;
    int32_t node, parent;
    int8_t sym;
} Entry;

typedef struct {
    Entry *v;
    int n, cap;
} BeamList;

static Entry *bl_push(BeamList *b)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 64;
        b->v = (Entry *)realloc(b->v, (size_t)b->cap * sizeof(Entry));
    }
    Entry *e = &b->v[b->n++];
    e->ptot = e->pnb = e->pb = -INFINITY; /* BeamEntry defaults, decode.py:23-25 */
    e->node = e->parent = -1;
    e->sym = -1;
    return e;
}

/* stable descending sort of indices by ptot == sorted(reverse=True, key=pr_total)
 * (decode.py:35-39).  Insertion sort on an index array: n <= 5*bw. */
static void sort_desc_stable(const Entry *v, int n, int *idx)
{
    for (int i = 0; i < n; ++i) idx[i] = i;
    for (int i = 1; i < n; ++i) {
        int k = idx[i];
        double key = v[k].ptot;
        int j = i - 1;
        while (j >= 0 && v[idx[j]].ptot < key) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = k;
    }
}

/* entropy(np.asarray(row)) for a float64 table row, decode.py:73-76 */
static double row_entropy(const double *r)
{
    double s = 0.0;
    for (int i = 0; i < N_BASES; ++i)
        if (r[i] > 0) s = s + r[i] * log(r[i]);
    return -s;
}

/* s_entropies[t] for a float64 matrix row, decode.py:135-138 with 67-76 */
static double frame_entropy_f64(const double *p)
{
    double S = ((0.0 + p[0]) + p[1]) + p[2] + p[3];
    double q[N_BASES];
    for (int i = 0; i < N_BASES; ++i) q[i] = (S == 0.0) ? p[i] : p[i] / S;
    double s = 0.0;
    for (int i = 0; i < N_BASES; ++i)
        if (q[i] > 0) s = s + q[i] * log(q[i]);
    return -s;
}

/* Same for a float32 matrix under numpy >= 2 promotion (NEP 50), which is what the
 * container's reference run does: float32 sum and division, math.log in float64 rounded
 * to float32 by the weak-scalar multiply, float32 accumulation (SURVEY.md 8c drift note).
 * Returned as the float32 value; the gate compares it with (float)s_threshold. */
static float frame_entropy_f32(const float *p)
{
    float S = ((0.0f + p[0]) + p[1]) + p[2] + p[3];
    float q[N_BASES];
    for (int i = 0; i < N_BASES; ++i) q[i] = (S == 0.0f) ? p[i] : p[i] / S;
    float s = 0.0f;
    for (int i = 0; i < N_BASES; ++i)
        if (q[i] > 0) {
            float lg = (float)log((double)q[i]);
            s = s + q[i] * lg;
        }
    return -s;
}

typedef struct {
    const void *mat;
    int is_f64;
    const double *table;
    int L;
    double s_thr, r_thr;
    uint64_t n_lookup, n_combine;
} Ctx;

static inline double mat_at(const Ctx *c, int64_t t, int k)
{
    return c->is_f64 ? ((const double *)c->mat)[t * 5 + k] : (double)((const float *)c->mat)[t * 5 + k];
}

/* apply_rna_model (decode.py:79-96): fills d[0..3] with the distribution to take logs of.
 * h_memo is the entropy memo slot of the context (entr_cache). */
static void apply_rna_model(Ctx *c, int64_t t, uint32_t ctx_idx, double *h_memo, int s_gate_open,
                            double *d)
{
    const double *r = c->table + (size_t)ctx_idx * 4;
    c->n_lookup++;
    if (isnan(*h_memo)) *h_memo = row_entropy(r);
    if (*h_memo < c->r_thr && s_gate_open) {
        c->n_combine++;
        /* combine_dists, decode.py:52-64 */
        if (c->is_f64) {
            const double *p = (const double *)c->mat + t * 5;
            double S = ((p[0] + p[1]) + p[2]) + p[3];
            for (int i = 0; i < N_BASES; ++i) d[i] = ((r[i] + p[i] / S) / 2) * S;
        } else {
            const float *p = (const float *)c->mat + t * 5;
            float S = ((p[0] + p[1]) + p[2]) + p[3];
            for (int i = 0; i < N_BASES; ++i) {
                float q = p[i] / S;
                d[i] = ((r[i] + (double)q) / 2) * (double)S;
            }
        }
    } else {
        for (int i = 0; i < N_BASES; ++i) d[i] = mat_at(c, t, i);
    }
}

/*
 * beam_search (decode.py:100-212).  Returns 0, or -1 on bad arguments.
 *  out_seq      symbols 0..3 of the best labeling (decode order, not reversed)
 *  out_scores   pr_total of the stable-sorted final candidates, first `topk`
 *  counters     [0] number of lm[context] reads (decode.py:83), [1] combine_dists calls
 */
int radian_oracle_beam_search(const void *mat, int is_f64, int64_t T, int beam_width,
                              const double *table, int L, double s_thr, double r_thr,
                              uint8_t *out_seq, int64_t out_cap, int64_t *out_len,
                              double *out_scores, int topk, int *n_final, uint64_t *counters)
{
    if (beam_width < 1 || T < 0) return -1;
    if (table && (L < 1 || L > 15)) return -1;
    Ctx cx = {mat, is_f64, table, L, s_thr, r_thr, 0, 0};
    const uint32_t ctx_mask = table ? (uint32_t)((1ull << (2 * L)) - 1) : 0;

    Trie tr = {0, 0, 0};
    int32_t root = trie_new(&tr, -1, -1);

    BeamList last = {0, 0, 0}, curr = {0, 0, 0};
    Entry *e0 = bl_push(&last);
    e0->node = root;
    e0->pb = 0.0;   /* log(1), decode.py:131 */
    e0->ptot = 0.0; /* decode.py:132 */

    int *order = NULL;
    int order_cap = 0;
    /* curr.entries[labeling] lookup: slot of a materialised node in curr, valid when stamp matches */
    int32_t *slot = NULL;
    int64_t *stamp = NULL;
    int64_t slot_cap = 0;

    for (int64_t t = 0; t < T; ++t) {
        /* s_entropies[t] and the signal half of the gate (decode.py:93) */
        int s_open = 0;
        if (table) {
            if (is_f64)
                s_open = frame_entropy_f64((const double *)mat + t * 5) > s_thr;
            else
                s_open = frame_entropy_f32((const float *)mat + t * 5) > (float)s_thr;
        }
        double lp_raw[5];
        for (int k = 0; k < 5; ++k) lp_raw[k] = ref_log(mat_at(&cx, t, k));

        if (last.n > order_cap) {
            order_cap = last.n * 2;
            order = (int *)realloc(order, (size_t)order_cap * sizeof(int));
        }
        sort_desc_stable(last.v, last.n, order);
        int nbest = last.n < beam_width ? last.n : beam_width;

        /* materialise surviving pending labelings so that every best beam is a trie node */
        for (int b = 0; b < nbest; ++b) {
            Entry *X = &last.v[order[b]];
            if (X->node < 0) {
                int32_t nn = trie_new(&tr, X->parent, X->sym);
                tr.v[X->parent].child[X->sym] = nn;
                X->node = nn;
            }
        }
        if (tr.n + 8 > slot_cap) {
            int64_t nc = (tr.n + 8) * 2;
            slot = (int32_t *)realloc(slot, (size_t)nc * sizeof(int32_t));
            stamp = (int64_t *)realloc(stamp, (size_t)nc * sizeof(int64_t));
            for (int64_t i = slot_cap; i < nc; ++i) stamp[i] = -1;
            slot_cap = nc;
        }

        curr.n = 0;
        for (int b = 0; b < nbest; ++b) {
            const Entry X = last.v[order[b]]; /* copy: curr pushes never alias, but keep it simple */
            const Node nd = tr.v[X.node];
            const int len = nd.len;
            const int lastc = nd.sym;

            /* ---- COPY BEAM (decode.py:150-175) */
            double pnb = -INFINITY;
            if (len > 0) {
                double lpc;
                if (table && len >= L + 1) {
                    double d[4];
                    uint32_t cidx = (nd.ctx >> 2) & ctx_mask;            /* labeling[-(L+1):-1] */
                    apply_rna_model(&cx, t, cidx, &tr.v[nd.parent].h_ext, s_open, d);
                    lpc = ref_log(d[lastc]);
                } else {
                    lpc = lp_raw[lastc];
                }
                pnb = X.pnb + lpc;
            }
            double pb = X.ptot + lp_raw[BLANK];
            Entry *ce;
            if (stamp[X.node] == t) {
                ce = &curr.v[slot[X.node]];
            } else {
                ce = bl_push(&curr);
                ce->node = X.node;
                stamp[X.node] = t;
                slot[X.node] = curr.n - 1;
            }
            ce->pnb = lae(ce->pnb, pnb);
            ce->pb = lae(ce->pb, pb);
            ce->ptot = lae(ce->ptot, lae(pb, pnb));

            /* ---- EXTEND BEAM (decode.py:177-201) */
            double lpe[4];
            if (table && len >= L) {
                double d[4];
                uint32_t cidx = nd.ctx & ctx_mask;                        /* labeling[-L:] */
                apply_rna_model(&cx, t, cidx, &tr.v[X.node].h_ext, s_open, d);
                for (int c = 0; c < 4; ++c) lpe[c] = ref_log(d[c]);
            } else {
                for (int c = 0; c < 4; ++c) lpe[c] = lp_raw[c];
            }
            for (int c = 0; c < N_BASES; ++c) {
                double base = (len > 0 && lastc == c) ? X.pb : X.ptot; /* decode.py:192-195 */
                double v = base + lpe[c];
                int32_t ch = tr.v[X.node].child[c];
                Entry *ee;
                if (ch >= 0 && stamp[ch] == t) {
                    ee = &curr.v[slot[ch]];
                } else {
                    ee = bl_push(&curr);
                    if (ch >= 0) {
                        ee->node = ch;
                        stamp[ch] = t;
                        slot[ch] = curr.n - 1;
                    } else {
                        ee->parent = X.node;
                        ee->sym = (int8_t)c;
                    }
                }
                ee->pnb = lae(ee->pnb, v);
                ee->ptot = lae(ee->ptot, v);
            }
        }
        BeamList tmp = last;
        last = curr;
        curr = tmp;
    }

    /* decode.py:207-210 */
    if (last.n > order_cap) {
        order_cap = last.n * 2;
        order = (int *)realloc(order, (size_t)order_cap * sizeof(int));
    }
    sort_desc_stable(last.v, last.n, order);
    const Entry *best = &last.v[order[0]];
    int64_t n;
    int32_t walk;
    int pending = -1;
    if (best->node >= 0) {
        n = tr.v[best->node].len;
        walk = best->node;
    } else {
        n = tr.v[best->parent].len + 1;
        walk = best->parent;
        pending = best->sym;
    }
    if (out_len) *out_len = n;
    if (out_seq && n <= out_cap) {
        int64_t i = n;
        if (pending >= 0) out_seq[--i] = (uint8_t)pending;
        while (i > 0) {
            out_seq[--i] = (uint8_t)tr.v[walk].sym;
            walk = tr.v[walk].parent;
        }
    }
    if (n_final) *n_final = last.n;
    if (out_scores)
        for (int k = 0; k < topk && k < last.n; ++k) out_scores[k] = last.v[order[k]].ptot;
    if (counters) {
        counters[0] = cx.n_lookup;
        counters[1] = cx.n_combine;
    }
    int rc = (out_seq && n > out_cap) ? -2 : 0;
    free(tr.v);
    free(last.v);
    free(curr.v);
    free(order);
    free(slot);
    free(stamp);
    return rc;
}

/* Decode reads [r0, r1) of a concatenated batch; thread-safe (no shared state).  Used by the
 * Python wrapper to run reads on several host threads for the CPU baseline. */
int radian_oracle_beam_search_batch(const void *mat, int is_f64, const int64_t *frame_offsets,
                                    int64_t r0, int64_t r1, int beam_width, const double *table, int L,
                                    double s_thr, double r_thr, uint8_t *out_seq,
                                    const int64_t *seq_offsets, int64_t *out_len, double *out_score,
                                    uint64_t *counters)
{
    for (int64_t r = r0; r < r1; ++r) {
        int64_t T = frame_offsets[r + 1] - frame_offsets[r];
        const char *m = (const char *)mat + (size_t)frame_offsets[r] * 5 * (is_f64 ? 8 : 4);
        uint64_t cnt[2];
        int rc = radian_oracle_beam_search(m, is_f64, T, beam_width, table, L, s_thr, r_thr,
                                           out_seq + seq_offsets[r], seq_offsets[r + 1] - seq_offsets[r],
                                           &out_len[r], out_score ? &out_score[r] : NULL, 1, NULL, cnt);
        if (rc) return rc;
        if (counters) {
            counters[2 * r] = cnt[0];
            counters[2 * r + 1] = cnt[1];
        }
    }
    return 0;
}

/*
 * assemble_matrices (matrix_assembly.py:6-53).  chunks: n matrices of chunk_len[k] x 5 float32
 * stored back to back; chunk k starts at global row k*step (create_vstack, :12-34).  For every
 * global row the *first* chunk covering it wins (average_dist discards np.add's result, :52);
 * rows covered by more than one chunk are cast to float64 and L1-normalised by sklearn
 * normalize (:53), others are passed through.  Output is float64 when any row was
 * normalised (np.asarray promotion, :44), else float32; the caller passes which one it
 * expects in out_is_f64 and gets -3 if that is wrong.  Returns T via *out_T.
 */
int radian_oracle_assemble(const float *chunks, const int32_t *chunk_len, int n_chunks, int step,
                           void *out, int out_is_f64, int64_t out_cap_rows, int64_t *out_T)
{
    if (step <= 0) return -1;
    int64_t T = 0;
    for (int k = 0; k < n_chunks; ++k) {
        /* create_vstack appends one row at a time; a chunk starting past the end is an IndexError */
        if (chunk_len[k] > 0 && (int64_t)k * step > T) return -4;
        int64_t end = (int64_t)k * step + chunk_len[k];
        if (chunk_len[k] > 0 && end > T) T = end;
    }
    if (out_T) *out_T = T;
    if (T > out_cap_rows) return -2;
    int64_t *first = (int64_t *)malloc((size_t)(T + 1) * sizeof(int64_t)); /* source row index */
    int32_t *cover = (int32_t *)calloc((size_t)(T + 1), sizeof(int32_t));
    int64_t row0 = 0;
    for (int k = 0; k < n_chunks; ++k) {
        for (int j = 0; j < chunk_len[k]; ++j) {
            int64_t t = (int64_t)k * step + j;
            if (cover[t]++ == 0) first[t] = row0 + j;
        }
        row0 += chunk_len[k];
    }
    int any_f64 = 0;
    for (int64_t t = 0; t < T; ++t) any_f64 |= cover[t] > 1;
    int rc = 0;
    if (any_f64 != (out_is_f64 != 0)) {
        rc = -3;
    } else {
        for (int64_t t = 0; t < T; ++t) {
            const float *src = chunks + first[t] * 5;
            if (!out_is_f64) {
                memcpy((float *)out + t * 5, src, 5 * sizeof(float));
            } else if (cover[t] > 1) {
                double x[5], nrm = 0.0;
                for (int i = 0; i < 5; ++i) {
                    x[i] = (double)src[i];
                    nrm = nrm + fabs(x[i]);
                }
                if (nrm < 10 * 2.220446049250313e-16) nrm = 1.0; /* _handle_zeros_in_scale */
                for (int i = 0; i < 5; ++i) ((double *)out)[t * 5 + i] = x[i] / nrm;
            } else {
                for (int i = 0; i < 5; ++i) ((double *)out)[t * 5 + i] = (double)src[i];
            }
        }
    }
    free(first);
    free(cover);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Chunk-mode stitching: sequence_assembly.py:19-48 (simple_assembly, add_count) followed by
 * np.argmax(consensus, axis=0) + index2base (basecall.py:122-123, sequence_assembly.py:90-97).
 *
 * The alignment of consecutive fragments is difflib.SequenceMatcher(None, prev, cur) of the
 * Python standard library (third-party to the reference; CPython 3.12 Lib/difflib.py), restated
 * here: __chain_b with the autojunk "popular element" rule (len(b) >= 200: elements occurring
 * more than len(b)//100 + 1 times are left out of b2j), find_longest_match (the j2len dynamic
 * programme, first strictly longer match wins, then extension over equal elements on both
 * sides), get_matching_blocks (recursion left and right of every block, sort, collapse of
 * adjacent blocks, sentinel), and max(blocks, key=size) = first block of maximal size.
 * Symbols are 0..3.  Returns 0, or -6 where the reference raises IndexError (a fragment that
 * does not fit the vote buffer after its single 1000-column growth step).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int i, j, k;
} blk_t;

static blk_t longest_match(const uint8_t *a, const uint8_t *b, int alo, int ahi, int blo, int bhi,
                           const uint8_t *popular, int *row0, int *row1)
{
    int besti = alo, bestj = blo, bestsize = 0;
    int *old = row0, *cur = row1; /* indexed j + 1 */
    for (int j = blo; j <= bhi; ++j) old[j] = 0;
    for (int i = alo; i < ahi; ++i) {
        cur[blo] = 0;
        for (int j = blo; j < bhi; ++j) {
            int k = 0;
            if (a[i] == b[j] && !popular[b[j]]) {
                k = old[j] + 1; /* j2len.get(j - 1, 0) + 1 */
                if (k > bestsize) {
                    besti = i - k + 1;
                    bestj = j - k + 1;
                    bestsize = k;
                }
            }
            cur[j + 1] = k;
        }
        int *t = old;
        old = cur;
        cur = t;
    }
    while (besti > alo && bestj > blo && a[besti - 1] == b[bestj - 1]) {
        --besti;
        --bestj;
        ++bestsize;
    }
    while (besti + bestsize < ahi && bestj + bestsize < bhi && a[besti + bestsize] == b[bestj + bestsize]) ++bestsize;
    blk_t r = {besti, bestj, bestsize};
    return r;
}

static int blk_cmp(const void *x, const void *y)
{
    const blk_t *p = (const blk_t *)x, *q = (const blk_t *)y;
    if (p->i != q->i) return p->i < q->i ? -1 : 1;
    if (p->j != q->j) return p->j < q->j ? -1 : 1;
    return (p->k > q->k) - (p->k < q->k);
}

/* displacement block[0] - block[1] of the first largest matching block */
static int stitch_disp(const uint8_t *a, int la, const uint8_t *b, int lb)
{
    uint8_t popular[4] = {0, 0, 0, 0};
    if (lb >= 200) {
        int cnt[4] = {0, 0, 0, 0};
        for (int j = 0; j < lb; ++j) cnt[b[j]]++;
        for (int c = 0; c < 4; ++c) popular[c] = cnt[c] > lb / 100 + 1;
    }
    int nmax = (la < lb ? la : lb) + 2;
    blk_t *blocks = (blk_t *)malloc(sizeof(blk_t) * (size_t)nmax);
    int *queue = (int *)malloc(sizeof(int) * 4 * (size_t)(2 * nmax + 2));
    int *row0 = (int *)malloc(sizeof(int) * (size_t)(lb + 2));
    int *row1 = (int *)malloc(sizeof(int) * (size_t)(lb + 2));
    int nb = 0, qh = 0, qt = 0;
    queue[0] = 0, queue[1] = la, queue[2] = 0, queue[3] = lb;
    qt = 1;
    while (qh < qt) {
        const int alo = queue[4 * qh], ahi = queue[4 * qh + 1], blo = queue[4 * qh + 2], bhi = queue[4 * qh + 3];
        ++qh;
        const blk_t x = longest_match(a, b, alo, ahi, blo, bhi, popular, row0, row1);
        if (x.k) {
            blocks[nb++] = x;
            if (alo < x.i && blo < x.j) {
                queue[4 * qt] = alo, queue[4 * qt + 1] = x.i, queue[4 * qt + 2] = blo, queue[4 * qt + 3] = x.j;
                ++qt;
            }
            if (x.i + x.k < ahi && x.j + x.k < bhi) {
                queue[4 * qt] = x.i + x.k, queue[4 * qt + 1] = ahi, queue[4 * qt + 2] = x.j + x.k, queue[4 * qt + 3] = bhi;
                ++qt;
            }
        }
    }
    qsort(blocks, (size_t)nb, sizeof(blk_t), blk_cmp);
    /* collapse adjacent blocks, append the sentinel, take the first block of maximal size */
    int bi = la, bj = lb, bk = 0, have = 0;
    int i1 = 0, j1 = 0, k1 = 0;
    for (int n = 0; n <= nb; ++n) {
        const int last = (n == nb);
        if (!last && i1 + k1 == blocks[n].i && j1 + k1 == blocks[n].j) {
            k1 += blocks[n].k;
        } else {
            if (k1 && (!have || k1 > bk)) {
                bi = i1, bj = j1, bk = k1;
                have = 1;
            }
            if (!last) i1 = blocks[n].i, j1 = blocks[n].j, k1 = blocks[n].k;
        }
    }
    if (!have) bi = la, bj = lb; /* only the (la, lb, 0) sentinel */
    free(blocks);
    free(queue);
    free(row0);
    free(row1);
    return bi - bj;
}

int radian_oracle_stitch(const uint8_t *sym, const int64_t *frag_off, int n_frags, int32_t *votes, int64_t cap,
                         uint8_t *consensus, int64_t *out_len)
{
    /* votes: 4 x cap int32, zeroed here; cap >= 1000 * (n_frags + 1) is always enough */
    int64_t census_len = 1000, pos = 0, length = 0;
    memset(votes, 0, sizeof(int32_t) * 4 * (size_t)cap);
    *out_len = 0;
    for (int f = 0; f < n_frags; ++f) {
        const uint8_t *cur = sym + frag_off[f];
        const int64_t len = frag_off[f + 1] - frag_off[f];
        int64_t start = 0;
        if (f > 0) {
            const uint8_t *prev = sym + frag_off[f - 1];
            const int64_t disp = stitch_disp(prev, (int)(frag_off[f] - frag_off[f - 1]), cur, (int)len);
            if (disp + pos + len > census_len) census_len += 1000;
            start = pos + disp;
            pos += disp;
        }
        int64_t skip = 0;
        if (start < 0) {
            skip = -start;
            start = 0;
        }
        for (int64_t i = skip; i < len; ++i) {
            const int64_t col = start + (i - skip);
            if (col >= census_len || col >= cap) return -6; /* IndexError in add_count */
            votes[(int64_t)cur[i] * cap + col]++;
        }
        if (f > 0 && pos + len > length) length = pos + len;
    }
    if (length > census_len) length = census_len;
    if (length < 0) length = 0;
    for (int64_t c = 0; c < length; ++c) {
        int best = 0;
        for (int s = 1; s < 4; ++s)
            if (votes[(int64_t)s * cap + c] > votes[(int64_t)best * cap + c]) best = s;
        consensus[c] = (uint8_t)best;
    }
    *out_len = length;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Signal preprocessing: preprocess.py:23-49 (mad_normalise and helpers) and :4-21 (get_windows),
 * call sites basecall.py:78,83.  signal is the raw int16 array of the fast5 file.
 *   median = np.median(signal)                       (float64, mean of the two middle values)
 *   mad    = np.median(|signal - median|)            (float64)
 *   z      = (x - median) / (1.4826 * mad), clipped to +-outlier
 * np.vectorize takes its output dtype from the first element: when the first sample is clipped
 * and outlier_z_score is a Python int (argparse type=int, basecall.py:25) the whole result is
 * int64, every z truncated towards zero.  out holds 8 bytes per sample either way; *is_int64
 * says which.  Returns 0, -7 (empty signal, ValueError preprocess.py:24-25) or -8 (MAD is zero,
 * ValueError preprocess.py:47-48).
 * ------------------------------------------------------------------------------------------ */
static int cmp_dbl(const void *x, const void *y)
{
    const double a = *(const double *)x, b = *(const double *)y;
    return (a > b) - (a < b);
}

static double median_sorted(const double *v, int64_t n)
{
    /* np.median: mean of the two middle elements (np.mean of two float64) */
    return (n & 1) ? v[n / 2] : (v[n / 2 - 1] + v[n / 2]) / 2.0;
}

int radian_oracle_mad_normalise(const int16_t *signal, int64_t n, double outlier, int outlier_is_int, void *out,
                                int *is_int64)
{
    if (n == 0) return -7;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) tmp[i] = (double)signal[i];
    qsort(tmp, (size_t)n, sizeof(double), cmp_dbl);
    const double median = median_sorted(tmp, n);
    for (int64_t i = 0; i < n; ++i) tmp[i] = fabs((double)signal[i] - median);
    qsort(tmp, (size_t)n, sizeof(double), cmp_dbl);
    const double mad = median_sorted(tmp, n);
    free(tmp);
    if (mad == 0.0) return -8;
    const double scale = 1.4826 * mad;
    double *o = (double *)out;
    for (int64_t i = 0; i < n; ++i) {
        double z = ((double)signal[i] - median) / scale;
        if (z > outlier) z = outlier;
        else if (z < -1 * outlier) z = -1 * outlier;
        o[i] = z;
    }
    const double z0 = ((double)signal[0] - median) / scale;
    *is_int64 = outlier_is_int && (z0 > outlier || z0 < -1 * outlier);
    if (*is_int64) {
        int64_t *oi = (int64_t *)out;
        for (int64_t i = 0; i < n; ++i) oi[i] = (int64_t)o[i]; /* C cast = truncation, as numpy's astype */
    }
    return 0;
}

/* get_windows: number of windows and pad_end for a signal of n samples (preprocess.py:9-20) */
int64_t radian_oracle_windows(int64_t n, int window, int step, int *pad_end)
{
    int64_t start = 0, count = 0;
    while (start + window <= n) {
        ++count;
        start += step;
    }
    *pad_end = (int)(window - (n - start));
    return count + 1;
}
