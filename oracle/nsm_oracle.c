/*
 * TEST INFRASTRUCTURE — CPU restatement (plain C) of the reference's pair-scoring path over the
 * packed arrays of napkon_string_matching/gpu/pack.py.  Never linked into, loaded by or called
 * from the product; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
 *
 * It deliberately shares no algorithm with the CUDA kernels: set intersection is a plain
 * two-pointer merge over the sorted ids (no signatures, no filters, no early exits) and the
 * LCS is the textbook two-row dynamic programme (not the bit-parallel recurrence).
 *
 * Reference statements followed (all under /root/reference/napkon_string_matching/):
 *   compare_terms              types/comparable_data.py:248-265   (ora_compare_terms)
 *   intersection_vs_union      compare/score_functions.py:6-13    (ora_jaccard)
 *   fuzzy_match -> QRatio/100  compare/score_functions.py:20-27   (ora_qratio; rapidfuzz 2.1.x
 *                              restated, parity unpinned — see oracle/reference_port.py)
 *   cross product + threshold  types/comparable_data.py:191,223-232,243  (ora_*_allpairs)
 *   category predicate         types/comparable_data.py:464-476   (keep_categories)
 *
 * Build: oracle/build.sh  ->  oracle/_build/libnsm_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint32_t *item_level_off; /* [n_items+1] */
    const uint32_t *level_off;      /* sets: [n_levels+1] token ranges; strings: [n_levels] start of each string */
    const uint32_t *level_len;      /* strings: [n_levels] lengths; sets: NULL */
    const uint32_t *tok;            /* sets: sorted unique ids; strings: unused */
    const uint8_t *chr;             /* strings: alphabet codes; sets: unused */
    uint32_t n_items;
} ora_side_t;

typedef struct {
    uint32_t left;
    uint32_t right;
    double score;
} ora_pair_t;

enum { ORA_FLAG_ZERO_UNION = 1, ORA_FLAG_INDEX_ERROR = 2 };
enum { ORA_JACCARD = 0, ORA_QRATIO = 1 };

static uint32_t merge_count(const uint32_t *a, uint32_t na, const uint32_t *b, uint32_t nb) {
    uint32_t i = 0, j = 0, c = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) i++;
        else if (a[i] > b[j]) j++;
        else { c++; i++; j++; }
    }
    return c;
}

/* len(A & B) / len(A | B); 0/0 is reported as NaN + flag (ZeroDivisionError in the reference) */
static double ora_jaccard(const ora_side_t *L, uint32_t gl, const ora_side_t *R, uint32_t gr,
                          uint32_t *flags) {
    uint32_t la = L->level_off[gl], na = L->level_off[gl + 1] - la;
    uint32_t lb = R->level_off[gr], nb = R->level_off[gr + 1] - lb;
    uint32_t inter = merge_count(L->tok + la, na, R->tok + lb, nb);
    uint32_t uni = na + nb - inter;
    if (uni == 0) { *flags |= ORA_FLAG_ZERO_UNION; return NAN; }
    return (double)inter / (double)uni;
}

static uint32_t lcs_dp(const uint8_t *a, uint32_t na, const uint8_t *b, uint32_t nb) {
    uint32_t *prev = (uint32_t *)calloc(2 * (size_t)(nb + 1), sizeof(uint32_t));
    uint32_t *cur = prev + nb + 1, *tmp, i, j, res;
    for (i = 0; i < na; i++) {
        cur[0] = 0;
        for (j = 1; j <= nb; j++) {
            if (a[i] == b[j - 1]) cur[j] = prev[j - 1] + 1;
            else cur[j] = prev[j] > cur[j - 1] ? prev[j] : cur[j - 1];
        }
        tmp = prev; prev = cur; cur = tmp;
    }
    res = prev[nb];
    free(prev < cur ? prev : cur);
    return res;
}

/* QRatio(a, b) / 100 on processed strings: 0 if either is empty, else
 * ((1.0 - dist/lensum) * 100) / 100 with dist = lensum - 2*LCS */
static double ora_qratio(const ora_side_t *L, uint32_t gl, const ora_side_t *R, uint32_t gr) {
    uint32_t la = L->level_off[gl], na = L->level_len[gl];
    uint32_t lb = R->level_off[gr], nb = R->level_len[gr];
    if (na == 0 || nb == 0) return 0.0;
    uint32_t lcs = lcs_dp(L->chr + la, na, R->chr + lb, nb);
    uint32_t lensum = na + nb, dist = lensum - 2 * lcs;
    double norm_dist = (double)dist / (double)lensum;
    double norm_sim = 1.0 - norm_dist;
    return (norm_sim * 100.0) / 100.0;
}

/* score = sum_{i=1..max(Kl,Kr)} f(left[min(i,Kl-1)], right[min(i,Kr-1)]) * 2^-i ; flat: f(l0,r0) */
static double ora_compare_terms(int func, int flat, const ora_side_t *L, uint32_t li,
                                const ora_side_t *R, uint32_t ri, uint32_t *flags) {
    uint32_t lg0 = L->item_level_off[li], kl = L->item_level_off[li + 1] - lg0;
    uint32_t rg0 = R->item_level_off[ri], kr = R->item_level_off[ri + 1] - rg0;
    uint32_t kmax = kl > kr ? kl : kr, i;
    double score = 0.0, factor = 1.0;
    if (kmax == 0) return 0.0;
    if (kl == 0 || kr == 0) { *flags |= ORA_FLAG_INDEX_ERROR; return NAN; }
    if (flat)
        return func == ORA_JACCARD ? ora_jaccard(L, lg0, R, rg0, flags) : ora_qratio(L, lg0, R, rg0);
    for (i = 1; i <= kmax; i++) {
        uint32_t gl = lg0 + (i < kl - 1 ? i : kl - 1), gr = rg0 + (i < kr - 1 ? i : kr - 1);
        double s = func == ORA_JACCARD ? ora_jaccard(L, gl, R, gr, flags) : ora_qratio(L, gl, R, gr);
        factor /= 2;
        score += s * factor;
    }
    return score;
}

static int keep_categories(int mode, uint64_t ml, uint64_t mr) {
    if (mode == 0) return 1;
    if (mode == 1) return (ml & mr) != 0 || (ml == 0 && mr == 0); /* list / list */
    return (ml & mr) != 0;                                         /* scalar in list, scalar == scalar */
}

/* Row-major all-pairs.  Returns the number of kept pairs (may exceed cap; only the first cap
 * are written, in row-major order). */
int64_t ora_allpairs(int func, int flat, const ora_side_t *L, const ora_side_t *R, uint32_t l_begin,
                     uint32_t l_end, double threshold, const uint64_t *l_cat, const uint64_t *r_cat,
                     int cat_mode, ora_pair_t *out, int64_t cap, uint32_t *flags_out) {
    int64_t n_rows = (int64_t)l_end - (int64_t)l_begin, row;
    int64_t *row_count = (int64_t *)calloc((size_t)(n_rows > 0 ? n_rows : 1) + 1, sizeof(int64_t));
    uint32_t flags = 0;
    int pass;
    int64_t total = 0;
    for (pass = 0; pass < 2; pass++) {
#pragma omp parallel for schedule(dynamic, 4) reduction(| : flags)
        for (row = 0; row < n_rows; row++) {
            uint32_t li = l_begin + (uint32_t)row, ri;
            int64_t pos = row_count[row], n = 0;
            uint32_t f = 0;
            for (ri = 0; ri < R->n_items; ri++) {
                double s;
                if (!keep_categories(cat_mode, l_cat ? l_cat[li] : 0, r_cat ? r_cat[ri] : 0)) continue;
                s = ora_compare_terms(func, flat, L, li, R, ri, &f);
                if (s >= threshold) {
                    if (pass == 1 && pos + n < cap) {
                        out[pos + n].left = li; out[pos + n].right = ri; out[pos + n].score = s;
                    }
                    n++;
                }
            }
            if (pass == 0) row_count[row] = n;
            flags |= f;
        }
        if (pass == 0) { /* exclusive prefix sum */
            int64_t acc = 0;
            for (row = 0; row < n_rows; row++) { int64_t c = row_count[row]; row_count[row] = acc; acc += c; }
            total = acc;
            if (out == NULL || cap == 0) break;
        }
    }
    free(row_count);
    if (flags_out) *flags_out = flags;
    return total;
}

/* One score per listed pair (used to spot-check huge runs). */
void ora_score_pairs(int func, int flat, const ora_side_t *L, const ora_side_t *R, const uint32_t *li,
                     const uint32_t *ri, int64_t n, double *out, uint32_t *flags_out) {
    uint32_t flags = 0;
    int64_t k;
#pragma omp parallel for schedule(static) reduction(| : flags)
    for (k = 0; k < n; k++) {
        uint32_t f = 0;
        out[k] = ora_compare_terms(func, flat, L, li[k], R, ri[k], &f);
        flags |= f;
    }
    if (flags_out) *flags_out = flags;
}

/* ---- CPU-baseline support only (oracle/shims/rapidfuzz, bench.py's reference arm) ------------
 * rapidfuzz computes the Indel distance behind QRatio with a bit-parallel LCS in C++; a shim that
 * ran the dynamic programme above would understate the reference's speed several times over.
 * This is Hyyro's recurrence over 64-bit words on UTF-32 code points.  No parity check uses it
 * (tests/test_oracle.py checks IT against the dynamic programme). */
uint32_t ora_lcs_utf32(const uint32_t *a, uint32_t na, const uint32_t *b, uint32_t nb) {
    enum { STACK_WORDS = 8, STACK_UNIQ = 128 };
    uint32_t words, n_uniq = 0, i, w, lcs = 0;
    uint64_t stack_masks[STACK_UNIQ * STACK_WORDS], stack_s[STACK_WORDS];
    uint32_t stack_cp[STACK_UNIQ];
    int16_t direct[256];
    uint64_t *masks = stack_masks, *S = stack_s;
    uint32_t *cps = stack_cp;
    if (na == 0 || nb == 0) return 0;
    if (na > nb) { const uint32_t *t = a; uint32_t tn = na; a = b; na = nb; b = t; nb = tn; }
    words = (na + 63) / 64;
    if (words > STACK_WORDS || na > STACK_UNIQ) {
        masks = (uint64_t *)calloc((size_t)na * words + words, sizeof(uint64_t));
        cps = (uint32_t *)malloc((size_t)na * sizeof(uint32_t));
        S = masks + (size_t)na * words;
    }
    memset(direct, 0xff, sizeof(direct));
    for (i = 0; i < na; i++) {
        uint32_t c = a[i], k;
        if (c < 256 && direct[c] >= 0) k = (uint32_t)direct[c];
        else {
            for (k = 0; k < n_uniq && cps[k] != c; k++) {}
            if (k == n_uniq) {
                cps[n_uniq++] = c;
                for (w = 0; w < words; w++) masks[(size_t)k * words + w] = 0;
                if (c < 256) direct[c] = (int16_t)k;
            }
        }
        masks[(size_t)k * words + i / 64] |= 1ull << (i % 64);
    }
    for (w = 0; w < words; w++) S[w] = ~0ull;
    for (i = 0; i < nb; i++) {
        uint32_t c = b[i], k;
        uint64_t carry = 0;
        const uint64_t *M;
        if (c < 256) { if (direct[c] < 0) continue; k = (uint32_t)direct[c]; }
        else { for (k = 0; k < n_uniq && cps[k] != c; k++) {} if (k == n_uniq) continue; }
        M = masks + (size_t)k * words;
        for (w = 0; w < words; w++) {
            uint64_t u = S[w] & M[w], sum = S[w] + u, sum2 = sum + carry;
            carry = (sum < u) | (sum2 < sum);
            S[w] = sum2 | (S[w] - u);
        }
    }
    for (w = 0; w < words; w++) lcs += (uint32_t)__builtin_popcountll(~S[w]);
    /* bits beyond the pattern length never change from 1: they do not count */
    if (masks != stack_masks) { free(masks); free(cps); }
    return lcs;
}
