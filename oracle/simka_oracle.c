/*
 * simka_oracle.c -- CPU restatement of the Simka hot path.  TEST INFRASTRUCTURE ONLY
 * (see simka_oracle.h for the contract and the parity pin).
 *
 * Reference paths: S/ = src/, GC/ = thirdparty/gatb-core/gatb-core/src/gatb/.
 * Nothing here is copied from the reference; each function restates the cited behaviour.
 */
#define _GNU_SOURCE
#include "simka_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * k-mer model
 * ---------------------------------------------------------------------------------------- */

/* GC/tools/misc/api/Data.hpp:179 -- nucleotide code = (c>>1)&3  => A=0 C=1 T=2 G=3 (N maps to 3) */
static inline unsigned nt_code(char c) { return ((unsigned char)c >> 1) & 3u; }
/* GC/tools/misc/api/Data.cpp:3 -- valid iff c in ACGTacgt */
static inline int nt_valid(char c) {
    switch (c) {
    case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1;
    default: return 0;
    }
}

/* GC/tools/math/LargeInt1.pri:137-155 -- reverse the 2-bit groups, complement = xor 2 */
uint64_t orc_revcomp(uint64_t x, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((x & 3u) ^ 2u);
        x >>= 2;
    }
    return r;
}

/* GC/kmer/impl/Model.hpp:726-765 (iterate), :858-884 (ModelCanonical first/next) */
static size_t kmers_impl(const char *seq, size_t len, int k, uint64_t *out, uint8_t *valid, int canonical) {
    if (len < (size_t)k) return 0;
    const uint64_t mask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint64_t fwd = 0, rev = 0;
    int bad = -1; /* index (from the right end) of the last bad char inside the window */
    size_t n = 0;
    for (size_t i = 0; i < len; i++) {
        unsigned c = nt_code(seq[i]);
        fwd = ((fwd << 2) | c) & mask;
        rev = (rev >> 2) | ((uint64_t)(c ^ 2u) << (2 * (k - 1)));
        if (!nt_valid(seq[i])) bad = k - 1; else bad--;
        if (i + 1 >= (size_t)k) {
            out[n] = canonical ? (fwd < rev ? fwd : rev) : fwd;
            valid[n] = (bad < 0);
            n++;
        }
    }
    return n;
}
size_t orc_canonical_kmers(const char *seq, size_t len, int k, uint64_t *canon, uint8_t *valid) {
    return kmers_impl(seq, len, k, canon, valid, 1);
}

/* the same for 32 <= k <= 63: Kmer<64>, Type = LargeInt<2> (GC/tools/math/LargeInt2.pri: value[1] is the high word; operator<
 * compares value[1] first, then value[0], i.e. the order of the 128-bit integer); span dispatch KSIZE_LIST "32 64 96 128",
 * CMakeLists.txt:66-71, Integer::apply GC/tools/math/Integer.hpp:76-85 */
typedef unsigned __int128 u128;
static size_t kmers_impl128(const char *seq, size_t len, int k, u128 *out, uint8_t *valid) {
    if (len < (size_t)k) return 0;
    const u128 mask = (((u128)1) << (2 * k)) - 1;  /* k <= 63 */
    u128 fwd = 0, rev = 0;
    int bad = -1;
    size_t n = 0;
    for (size_t i = 0; i < len; i++) {
        unsigned c = nt_code(seq[i]);
        fwd = ((fwd << 2) | c) & mask;
        rev = (rev >> 2) | ((u128)(c ^ 2u) << (2 * (k - 1)));
        if (!nt_valid(seq[i])) bad = k - 1; else bad--;
        if (i + 1 >= (size_t)k) {
            out[n] = fwd < rev ? fwd : rev;
            valid[n] = (bad < 0);
            n++;
        }
    }
    return n;
}
/* test hook: canonical k-mers as (lo, hi) word pairs */
size_t orc_canonical_kmers128(const char *seq, size_t len, int k, uint64_t *lo, uint64_t *hi, uint8_t *valid) {
    if (k < 1 || k > 63 || len < (size_t)k) return 0;
    u128 *tmp = malloc((len - k + 1) * sizeof(u128));
    const size_t n = kmers_impl128(seq, len, k, tmp, valid);
    for (size_t i = 0; i < n; i++) { lo[i] = (uint64_t)tmp[i]; hi[i] = (uint64_t)(tmp[i] >> 64); }
    free(tmp);
    return n;
}
size_t orc_direct_kmers(const char *seq, size_t len, int k, uint64_t *fwd, uint8_t *valid) {
    return kmers_impl(seq, len, k, fwd, valid, 0);
}

/* GC/kmer/impl/Model.hpp:1220-1251 -- ban m-mers containing AA except at the first position */
static int mmer_allowed(uint32_t mmer, int m) {
    uint64_t mmask_m1 = (1ull << ((m - 2) * 2)) - 1;
    uint64_t mask_ma1 = 0x5555555555555555ull & mmask_m1;
    uint64_t a1 = mmer;
    a1 = ~(a1 | (a1 >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 == 0;
}
/* GC/kmer/impl/Model.hpp:1039-1062 (LUT) and :1254-1287 (computeNewMinimizerOriginal) */
uint32_t orc_minimizer(uint64_t kmer_fwd, int k, int m) {
    const uint32_t mmask = (1u << (2 * m)) - 1;
    uint32_t best = mmask; /* default minimizer = 4^m - 1 */
    uint64_t val = kmer_fwd;
    for (int idx = k - m; idx >= 0; idx--) {
        uint32_t mm = (uint32_t)(val & mmask);
        uint32_t rc = (uint32_t)orc_revcomp(mm, m);
        if (rc < mm) mm = rc;
        if (!mmer_allowed(mm, m)) mm = mmask;
        if (mm < best) best = mm;
        val >>= 2;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------
 * run state
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    uint64_t *kmers;      /* low words */
    uint64_t *kmers_hi;   /* high words, k > 31 only (NULL otherwise: the high word is 0) */
    uint32_t *counts;
    size_t n;
    uint64_t nb_reads, nb_distinct, nb_kmers, chord_n2;
} sample_t;

typedef struct {
    /* S/core/SimkaDistance.hpp:86-125 */
    uint64_t *distinct_shared, *braycurtis_num;      /* sym */
    uint64_t *nb_shared;                             /* N*N */
    long double *chord_ninj;                         /* N*N */
    uint64_t *hellinger, *kulczynski, *canberra, *whittaker;
    long double *kl;
    uint64_t nb_distinct_kmers, nb_shared_kmers;
} stats_t;

struct orc_run {
    const char *seqs;
    const uint64_t *read_off, *sample_read_off;
    int n, k, simple, complex_, nthreads;
    int64_t amin, amax;
    sample_t *samples;
    stats_t st;
    int counted, merged;
};

static size_t sym_size(int n) { return (size_t)n * (n + 1) / 2; }
/* S/core/SimkaAlgorithm.hpp:364 */
static inline size_t sym_index(int n, int i, int j) {
    return (size_t)j + (size_t)(n - 1) * i - ((size_t)i * (i - 1) / 2);
}

static void stats_alloc(stats_t *s, int n) {
    size_t nn = (size_t)n * n, ns = sym_size(n);
    memset(s, 0, sizeof *s);
    s->distinct_shared = calloc(ns, 8);
    s->braycurtis_num = calloc(ns, 8);
    s->nb_shared = calloc(nn, 8);
    s->chord_ninj = calloc(nn, sizeof(long double));
    s->hellinger = calloc(nn, 8);
    s->kulczynski = calloc(nn, 8);
    s->canberra = calloc(nn, 8);
    s->whittaker = calloc(nn, 8);
    s->kl = calloc(nn, sizeof(long double));
}
static void stats_free(stats_t *s) {
    free(s->distinct_shared); free(s->braycurtis_num); free(s->nb_shared); free(s->chord_ninj);
    free(s->hellinger); free(s->kulczynski); free(s->canberra); free(s->whittaker); free(s->kl);
    memset(s, 0, sizeof *s);
}
/* S/core/SimkaDistance.cpp:156-213 (operator+=) */
static void stats_add(stats_t *a, const stats_t *b, int n) {
    size_t nn = (size_t)n * n, ns = sym_size(n);
    a->nb_distinct_kmers += b->nb_distinct_kmers;
    a->nb_shared_kmers += b->nb_shared_kmers;
    for (size_t i = 0; i < ns; i++) {
        a->distinct_shared[i] += b->distinct_shared[i];
        a->braycurtis_num[i] += b->braycurtis_num[i];
    }
    for (size_t i = 0; i < nn; i++) {
        a->nb_shared[i] += b->nb_shared[i];
        a->chord_ninj[i] += b->chord_ninj[i];
        a->hellinger[i] += b->hellinger[i];
        a->kulczynski[i] += b->kulczynski[i];
        a->canberra[i] += b->canberra[i];
        a->whittaker[i] += b->whittaker[i];
        a->kl[i] += b->kl[i];
    }
}

orc_run *orc_run_create(const char *seqs, const uint64_t *read_off, const uint64_t *sample_read_off,
                        int nsamples, int k, int64_t amin, int64_t amax, int simple, int complex_,
                        int nthreads) {
    if (k < 1 || k > 63 || nsamples < 1) return NULL; /* spans 32 and 64 (Kmer<32>, Kmer<64>): GC/kmer/impl/Model.hpp:401-406 */
    orc_run *r = calloc(1, sizeof *r);
    if (!r) return NULL;
    r->seqs = seqs; r->read_off = read_off; r->sample_read_off = sample_read_off;
    r->n = nsamples; r->k = k; r->amin = amin; r->amax = amax;
    r->simple = simple; r->complex_ = complex_;
#ifdef _OPENMP
    r->nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    r->nthreads = 1;
#endif
    r->samples = calloc((size_t)nsamples, sizeof(sample_t));
    stats_alloc(&r->st, nsamples);
    return r;
}

void orc_run_destroy(orc_run *r) {
    if (!r) return;
    for (int s = 0; s < r->n; s++) { free(r->samples[s].kmers); free(r->samples[s].kmers_hi); free(r->samples[s].counts); }
    free(r->samples);
    stats_free(&r->st);
    free(r);
}

/* ------------------------------------------------------------------------------------------
 * stage 1: exact multiset count of valid canonical k-mers per sample + abundance filter
 *   GC/kmer/impl/SortingCountAlgorithm.cpp:600-653 (what is computed: exact counts, emitted in
 *   increasing k-mer order per partition), S/minikc/MiniKC.hpp:54-79 (filter + totals),
 *   S/SimkaCount.cpp:303-317 (the four per-sample totals)
 * ---------------------------------------------------------------------------------------- */

static void radix_sort_u64(uint64_t *a, uint64_t *tmp, size_t n, int bits) {
    uint64_t *src = a, *dst = tmp;
    for (int shift = 0; shift < bits; shift += 11) {
        size_t hist[2048] = {0};
        for (size_t i = 0; i < n; i++) hist[(src[i] >> shift) & 2047]++;
        size_t sum = 0; int single = 0;
        for (int d = 0; d < 2048; d++) { size_t c = hist[d]; if (c == n) single = 1; hist[d] = sum; sum += c; }
        if (single) continue;
        for (size_t i = 0; i < n; i++) dst[hist[(src[i] >> shift) & 2047]++] = src[i];
        uint64_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * 8);
}

static void radix_sort_u128(u128 *a, u128 *tmp, size_t n, int bits) {
    u128 *src = a, *dst = tmp;
    for (int shift = 0; shift < bits; shift += 11) {
        size_t hist[2048] = {0};
        for (size_t i = 0; i < n; i++) hist[(size_t)(src[i] >> shift) & 2047]++;
        size_t sum = 0; int single = 0;
        for (int d = 0; d < 2048; d++) { size_t c = hist[d]; if (c == n) single = 1; hist[d] = sum; sum += c; }
        if (single) continue;
        for (size_t i = 0; i < n; i++) dst[hist[(size_t)(src[i] >> shift) & 2047]++] = src[i];
        u128 *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(u128));
}

/* k > 31: same counting on 128-bit k-mers (the DSK path; MiniKC and its quirks are k <= 15 only) */
static void count_sample128(orc_run *r, int s) {
    sample_t *sm = &r->samples[s];
    const int k = r->k;
    uint64_t r0 = r->sample_read_off[s], r1 = r->sample_read_off[s + 1];
    size_t cap = 0, maxlen = 0;
    for (uint64_t q = r0; q < r1; q++) {
        uint64_t len = r->read_off[q + 1] - r->read_off[q];
        if (len >= (uint64_t)k) cap += len - k + 1;
        if (len > maxlen) maxlen = len;
    }
    u128 *keys = malloc((cap ? cap : 1) * sizeof(u128)), *tmp = malloc((cap ? cap : 1) * sizeof(u128));
    u128 *kb = malloc((maxlen + 1) * sizeof(u128)); uint8_t *vb = malloc(maxlen + 1);
    size_t nk = 0;
    for (uint64_t q = r0; q < r1; q++) {
        const size_t len = r->read_off[q + 1] - r->read_off[q];
        const size_t m = kmers_impl128(r->seqs + r->read_off[q], len, k, kb, vb);
        for (size_t i = 0; i < m; i++) if (vb[i]) keys[nk++] = kb[i];
    }
    free(kb); free(vb);
    radix_sort_u128(keys, tmp, nk, 2 * k);
    free(tmp);
    size_t nd = 0;
    for (size_t i = 0; i < nk;) { size_t j = i + 1; while (j < nk && keys[j] == keys[i]) j++; nd++; i = j; }
    sm->kmers = malloc((nd ? nd : 1) * 8);
    sm->kmers_hi = malloc((nd ? nd : 1) * 8);
    sm->counts = malloc((nd ? nd : 1) * 4);
    size_t o = 0;
    sm->nb_distinct = sm->nb_kmers = sm->chord_n2 = 0;
    for (size_t i = 0; i < nk;) {
        size_t j = i + 1; while (j < nk && keys[j] == keys[i]) j++;
        int64_t c = (int64_t)(j - i);
        if (!(c < r->amin || c > r->amax)) {
            sm->kmers[o] = (uint64_t)keys[i]; sm->kmers_hi[o] = (uint64_t)(keys[i] >> 64); sm->counts[o] = (uint32_t)c; o++;
            sm->nb_distinct += 1; sm->nb_kmers += (uint64_t)c;
            sm->chord_n2 += (uint64_t)pow((double)c, 2);
        }
        i = j;
    }
    sm->n = o;
    sm->nb_reads = r1 - r0;
    free(keys);
}

static void count_sample(orc_run *r, int s) {
    if (r->k > 31) { count_sample128(r, s); return; }
    sample_t *sm = &r->samples[s];
    const int k = r->k;
    uint64_t r0 = r->sample_read_off[s], r1 = r->sample_read_off[s + 1];
    size_t cap = 0, last_cap = 0;
    for (uint64_t q = r0; q < r1; q++) {
        uint64_t len = r->read_off[q + 1] - r->read_off[q];
        if (len >= (uint64_t)k) { cap += len - k + 1; last_cap = len - k + 1; }
        else if (k <= 15) cap += last_cap;  /* MiniKC replays the previous read, see below */
    }
    uint64_t *keys = malloc((cap ? cap : 1) * 8), *tmp = malloc((cap ? cap : 1) * 8);
    size_t maxlen = 0;
    for (uint64_t q = r0; q < r1; q++) { uint64_t l = r->read_off[q + 1] - r->read_off[q]; if (l > maxlen) maxlen = l; }
    uint64_t *kb = malloc((maxlen + 1) * 8); uint8_t *vb = malloc(maxlen + 1);
    size_t nk = 0, last_m = 0;
    for (uint64_t q = r0; q < r1; q++) {
        const char *seq = r->seqs + r->read_off[q];
        size_t len = r->read_off[q + 1] - r->read_off[q];
        size_t m;
        /* invalid windows are skipped -- except for k <= 15, where simkaCount switches to MiniKC (src/SimkaCount.cpp:283-288),
         * whose counting loop does not test isValid() (src/minikc/MiniKC.hpp:219-231): every window is counted, a non-ACGT
         * character contributing its (c>>1)&3 code.  Second MiniKC quirk, pinned on reference runs (tests/golden/truth_ref/
         * example_n_results_k{12,15}_*): its k-mer iterator is refilled by Model::Iterator::setData -> Model::build
         * (GC/kmer/impl/Model.hpp:603-610,499-513), which returns early WITHOUT touching the k-mer vector when the read is
         * shorter than k, so MiniKC::count (single-threaded, reads in bank order) walks the PREVIOUS read's k-mers again. */
        if (k <= 15 && len < (size_t)k) m = last_m;           /* kb still holds the previous long-enough read's k-mers */
        else { m = orc_canonical_kmers(seq, len, k, kb, vb); last_m = m; }
        for (size_t i = 0; i < m; i++) if (vb[i] || k <= 15) keys[nk++] = kb[i];
    }
    free(kb); free(vb);
    radix_sort_u64(keys, tmp, nk, 2 * k);
    free(tmp);
    /* run-length count, filter (S/minikc/MiniKC.hpp:56), totals (:60-62; pow(c,2) -> u64) */
    size_t nd = 0;
    for (size_t i = 0; i < nk;) { size_t j = i + 1; while (j < nk && keys[j] == keys[i]) j++; nd++; i = j; }
    sm->kmers = malloc((nd ? nd : 1) * 8);
    sm->counts = malloc((nd ? nd : 1) * 4);
    size_t o = 0;
    sm->nb_distinct = sm->nb_kmers = sm->chord_n2 = 0;
    for (size_t i = 0; i < nk;) {
        size_t j = i + 1; while (j < nk && keys[j] == keys[i]) j++;
        int64_t c = (int64_t)(j - i); /* CountNumber is int32 (GC/system/api/types.hpp:49-52) */
        if (!(c < r->amin || c > r->amax)) {
            sm->kmers[o] = keys[i]; sm->counts[o] = (uint32_t)c; o++;
            sm->nb_distinct += 1; sm->nb_kmers += (uint64_t)c;
            sm->chord_n2 += (uint64_t)pow((double)c, 2);
        }
        i = j;
    }
    sm->n = o;
    sm->nb_reads = r1 - r0;
    free(keys);
}

void orc_run_count(orc_run *r) {
    if (r->counted) return;
#pragma omp parallel for schedule(dynamic, 1) num_threads(r->nthreads)
    for (int s = 0; s < r->n; s++) count_sample(r, s);
    r->counted = 1;
}

/* ------------------------------------------------------------------------------------------
 * stage 2: N-way merge + accumulator update
 * ---------------------------------------------------------------------------------------- */

typedef struct { u128 kmer; int s; } heap_item;
#define SAMPLE_KMER(sm, i) ((sm)->kmers_hi ? (((u128)(sm)->kmers_hi[i] << 64) | (sm)->kmers[i]) : (u128)(sm)->kmers[i])
static inline int heap_less(heap_item a, heap_item b) { return a.kmer < b.kmer || (a.kmer == b.kmer && a.s < b.s); }
static void heap_push(heap_item *h, int *n, heap_item x) {
    int i = (*n)++; h[i] = x;
    while (i > 0) { int p = (i - 1) / 2; if (!heap_less(h[i], h[p])) break; heap_item t = h[i]; h[i] = h[p]; h[p] = t; i = p; }
}
static heap_item heap_pop(heap_item *h, int *n) {
    heap_item top = h[0]; h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, rr = l + 1, m = i;
        if (l < *n && heap_less(h[l], h[m])) m = l;
        if (rr < *n && heap_less(h[rr], h[m])) m = rr;
        if (m == i) break;
        heap_item t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
    return top;
}

/* S/core/SimkaAlgorithm.hpp:341-516 -- updateDistance{Default,Simple,Complex} on one abundance vector.
 * `shared` = indices with count>0 in increasing order (:342-345). */
static void update_distance(const orc_run *r, stats_t *st, const int32_t *counts, const uint16_t *shared, int ns) {
    const int n = r->n;
    /* default (:356-378) */
    for (int ii = 0; ii < ns; ii++)
        for (int jj = ii + 1; jj < ns; jj++) {
            int i = shared[ii], j = shared[jj];
            size_t idx = sym_index(n, i, j);
            uint64_t ai = (uint64_t)counts[i], aj = (uint64_t)counts[j];
            st->nb_shared[(size_t)i * n + j] += ai;
            st->nb_shared[(size_t)j * n + i] += aj;
            st->distinct_shared[idx] += 1;
            st->braycurtis_num[idx] += ai < aj ? ai : aj;
        }
    /* simple (:381-402): chord += u64 product into long double; hellinger is a u64 fed with a double */
    if (r->simple)
        for (int ii = 0; ii < ns; ii++)
            for (int jj = ii + 1; jj < ns; jj++) {
                int i = shared[ii], j = shared[jj];
                uint64_t ai = (uint64_t)counts[i], aj = (uint64_t)counts[j];
                size_t c = (size_t)i * n + j;
                st->chord_ninj[c] += ai * aj;
                st->hellinger[c] = (uint64_t)((double)st->hellinger[c] + sqrt((double)(ai * aj)));
                st->kulczynski[c] += ai < aj ? ai : aj;
            }
    /* complex (:404-516): every i<j with c_i>0 or c_j>0; N_i = global solid k-mers of sample i */
    if (r->complex_) {
        for (int i = 0; i < n; i++) {
            if (counts[i]) {
                for (int j = i + 1; j < n; j++) {
                    double ai = counts[i], aj = counts[j], d1, d2 = 0;
                    double Ni = (double)r->samples[i].nb_kmers, Nj = (double)r->samples[j].nb_kmers;
                    double yX = aj * Ni, xY = ai * Nj;
                    double xi = ai / Ni;
                    d1 = xi * log((2 * xY) / (xY + yX));
                    if (aj) { double xj = aj / Nj; d2 = xj * log((2 * yX) / (xY + yX)); }
                    size_t c = (size_t)i * n + j;
                    st->kl[c] += d1 + d2;
                    st->canberra[c] = (uint64_t)((double)st->canberra[c] + fabs(ai - aj) / (ai + aj));
                    int d = (int)((uint64_t)(ai * Nj) - (uint64_t)(aj * Ni));
                    d = (int)(d < 0 ? 0u - (unsigned)d : (unsigned)d); /* abs(int), wrapping like the x86 build */
                    st->whittaker[c] += (uint64_t)(int64_t)d;
                }
            } else {
                for (int jj = 0; jj < ns; jj++) {
                    int j = shared[jj];
                    if (i > j) continue;
                    double ai = 0, aj = counts[j];
                    double Ni = (double)r->samples[i].nb_kmers, Nj = (double)r->samples[j].nb_kmers;
                    double xY = ai * Nj, yX = aj * Ni;
                    double xj = aj / Nj;
                    double d2 = xj * log((2 * yX) / (xY + yX));
                    size_t c = (size_t)i * n + j;
                    st->kl[c] += 0 + d2;
                    st->canberra[c] = (uint64_t)((double)st->canberra[c] + fabs(ai - aj) / (ai + aj));
                    int d = (int)((uint64_t)(ai * Nj) - (uint64_t)(aj * Ni));
                    d = (int)(d < 0 ? 0u - (unsigned)d : (unsigned)d);
                    st->whittaker[c] += (uint64_t)(int64_t)d;
                }
            }
        }
    }
}

/* S/SimkaMerge.cpp:1164-1264 (k-way merge) + :1307-1326 (insert gate) over the k-mer range [lo,hi) */
static void merge_range(const orc_run *r, stats_t *st, u128 lo, u128 hi, int last) {
    const int n = r->n;
    size_t *pos = malloc((size_t)n * sizeof(size_t)), *end = malloc((size_t)n * sizeof(size_t));
    heap_item *heap = malloc((size_t)n * sizeof(heap_item));
    int32_t *counts = calloc((size_t)n, 4);
    uint16_t *shared = malloc((size_t)n * 2);
    int hn = 0;
    for (int s = 0; s < n; s++) {
        const sample_t *sm = &r->samples[s];
        size_t a = 0, b = sm->n; /* lower_bound(lo) */
        while (a < b) { size_t m = (a + b) / 2; if (SAMPLE_KMER(sm, m) < lo) a = m + 1; else b = m; }
        pos[s] = a;
        if (last) end[s] = sm->n;
        else { size_t c = a, d = sm->n; while (c < d) { size_t m = (c + d) / 2; if (SAMPLE_KMER(sm, m) < hi) c = m + 1; else d = m; } end[s] = c; }
        if (pos[s] < end[s]) heap_push(heap, &hn, (heap_item){SAMPLE_KMER(sm, pos[s]), s});
    }
    while (hn > 0) {
        u128 kmer = heap[0].kmer;
        int ns = 0;
        while (hn > 0 && heap[0].kmer == kmer) {
            heap_item it = heap_pop(heap, &hn);
            int s = it.s;
            counts[s] += (int32_t)r->samples[s].counts[pos[s]];
            shared[ns++] = (uint16_t)s; /* popped in increasing sample order for equal k-mers */
            pos[s]++;
            if (pos[s] < end[s]) heap_push(heap, &hn, (heap_item){SAMPLE_KMER(&r->samples[s], pos[s]), s});
        }
        st->nb_distinct_kmers += 1;
        if (r->complex_ || ns > 1) {
            if (ns > 1) st->nb_shared_kmers += 1;
            update_distance(r, st, counts, shared, ns);
        }
        for (int q = 0; q < ns; q++) counts[shared[q]] = 0;
    }
    free(pos); free(end); free(heap); free(counts); free(shared);
}

void orc_run_merge(orc_run *r) {
    if (r->merged) return;
    if (!r->counted) orc_run_count(r);
    /* the reference merges one minimizer partition per process and sums the per-partition
     * statistics (S/SimkaPotara.hpp:1161-1174); any partition of the k-mer space gives the same
     * integers (tests/simple_test.py:122-129).  We split on the top 8 bits of the 2k-bit k-mer. */
    const int nparts = 256;
    const int kb = 2 * r->k;
    stats_t *part = calloc((size_t)r->nthreads, sizeof(stats_t));
    for (int t = 0; t < r->nthreads; t++) stats_alloc(&part[t], r->n);
#pragma omp parallel for schedule(dynamic, 1) num_threads(r->nthreads)
    for (int p = 0; p < nparts; p++) {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        u128 lo, hi;
        if (kb >= 8) { lo = (u128)p << (kb - 8); hi = (u128)(p + 1) << (kb - 8); }
        else { lo = 0; hi = ~(u128)0; if (p > 0) continue; }
        merge_range(r, &part[t], lo, hi, p == nparts - 1 || kb < 8);
    }
    for (int t = 0; t < r->nthreads; t++) { stats_add(&r->st, &part[t], r->n); stats_free(&part[t]); }
    free(part);
    r->merged = 1;
}

size_t orc_sample_size(const orc_run *r, int s) { return r->samples[s].n; }
const uint64_t *orc_sample_kmers(const orc_run *r, int s) { return r->samples[s].kmers; }
const uint64_t *orc_sample_kmers_hi(const orc_run *r, int s) { return r->samples[s].kmers_hi; }  /* NULL for k <= 31 */
const uint32_t *orc_sample_counts(const orc_run *r, int s) { return r->samples[s].counts; }
void orc_sample_totals(const orc_run *r, int s, uint64_t out[4]) {
    const sample_t *sm = &r->samples[s];
    out[0] = sm->nb_reads; out[1] = sm->nb_distinct; out[2] = sm->nb_kmers; out[3] = sm->chord_n2;
}
uint64_t orc_nb_distinct_kmers(const orc_run *r) { return r->st.nb_distinct_kmers; }
uint64_t orc_nb_shared_kmers(const orc_run *r) { return r->st.nb_shared_kmers; }

int orc_get_u64(const orc_run *r, const char *name, uint64_t *out) {
    size_t nn = (size_t)r->n * r->n, ns = sym_size(r->n);
    const stats_t *s = &r->st;
    if (!strcmp(name, "distinct_shared")) memcpy(out, s->distinct_shared, ns * 8);
    else if (!strcmp(name, "braycurtis_num")) memcpy(out, s->braycurtis_num, ns * 8);
    else if (!strcmp(name, "nb_shared")) memcpy(out, s->nb_shared, nn * 8);
    else if (!strcmp(name, "hellinger")) memcpy(out, s->hellinger, nn * 8);
    else if (!strcmp(name, "kulczynski")) memcpy(out, s->kulczynski, nn * 8);
    else if (!strcmp(name, "canberra")) memcpy(out, s->canberra, nn * 8);
    else if (!strcmp(name, "whittaker")) memcpy(out, s->whittaker, nn * 8);
    else if (!strcmp(name, "chord_ninj")) for (size_t i = 0; i < nn; i++) out[i] = (uint64_t)s->chord_ninj[i];
    else return -1;
    return 0;
}
int orc_get_f64(const orc_run *r, const char *name, double *out) {
    size_t nn = (size_t)r->n * r->n;
    if (!strcmp(name, "kl")) for (size_t i = 0; i < nn; i++) out[i] = (double)r->st.kl[i];
    else if (!strcmp(name, "chord_ninj")) for (size_t i = 0; i < nn; i++) out[i] = (double)r->st.chord_ninj[i];
    else return -1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * distances: S/core/SimkaDistance.cpp:920-1226, matrices S/core/SimkaDistance.hpp:155-475
 * ---------------------------------------------------------------------------------------- */

typedef struct {
    int n, simple, complex_;
    const uint64_t *nb_distinct, *nb_kmers;     /* _nbSolidDistinctKmersPerBank, _nbSolidKmersPerBank */
    long double *chord_sqrt_n2;                 /* sqrt(sum c^2): S/core/SimkaDistance.cpp:132 */
    const uint64_t *distinct_shared, *braycurtis_num, *nb_shared;
    const long double *chord_ninj;
    const uint64_t *hellinger, *kulczynski, *canberra, *whittaker;
    const long double *kl;
} dist_in;

static void get_abc(const dist_in *d, int i, int j, uint64_t *a, uint64_t *b, uint64_t *c) { /* :920-926 */
    *a = d->distinct_shared[sym_index(d->n, i, j)];
    *b = d->nb_distinct[i] - *a;
    *c = d->nb_distinct[j] - *a;
}
#define NS(i) (d->nb_kmers[i])
#define M2(m, i, j) ((m)[(size_t)(i) * d->n + (j)])

static double d_ab_braycurtis(const dist_in *d, int i, int j) { /* :928-939 */
    double union_ = NS(i) + NS(j);
    if (union_ == 0) return 1;
    double inter = 2 * d->braycurtis_num[sym_index(d->n, i, j)];
    return 1 - inter / union_;
}
static double d_ab_chord(const dist_in *d, int i, int j) { /* :942-959 */
    double den = d->chord_sqrt_n2[i] * d->chord_sqrt_n2[j];
    if (den == 0) return sqrt(2);
    long double v = sqrtl(2 - 2 * M2(d->chord_ninj, i, j) / den);
    return v;
}
static double d_ab_hellinger(const dist_in *d, int i, int j) { /* :962-972 */
    double union_ = sqrt(NS(i)) * sqrt(NS(j));
    if (union_ == 0) return sqrt(2);
    double inter = 2 * M2(d->hellinger, i, j);
    return sqrt(2 - (inter / union_));
}
static double d_ab_whittaker(const dist_in *d, int i, int j) { /* :988-998 */
    long double union_ = NS(i) * NS(j);
    if (union_ == 0) return 1;
    long double inter = M2(d->whittaker, i, j);
    double w = 0.5 * (inter / union_);
    return w;
}
static double d_ab_kl(const dist_in *d, int i, int j) { /* :1001-1007 */
    if (M2(d->kl, i, j) == 0) return 1;
    return sqrt(0.5 * M2(d->kl, i, j));
}
static double d_ab_canberra(const dist_in *d, int i, int j) { /* :1010-1021 */
    uint64_t ua, ub, uc; get_abc(d, i, j, &ua, &ub, &uc);
    double a = (double)ua, b = (double)ub, c = (double)uc;
    if ((a + b + c) == 0) return 1;
    return (1 / (a + b + c)) * M2(d->canberra, i, j);
}
static double d_ab_kulczynski(const dist_in *d, int i, int j) { /* :1024-1038; [j][i] is never written */
    if (NS(i) == 0 || NS(j) == 0) return 1;
    long double n1 = (double)M2(d->kulczynski, i, j) / (double)NS(i);
    long double n2 = (double)M2(d->kulczynski, j, i) / (double)NS(j);
    double v = 1 - 0.5 * (n1 + n2);
    return v;
}
static double d_ab_jaccard_simka(const dist_in *d, int i, int j, int sym) { /* :1041-1065 */
    double A1 = M2(d->nb_shared, i, j), B1 = M2(d->nb_shared, j, i), A0 = NS(i), B0 = NS(j);
    double num, den;
    if (sym) { num = A1 + B1; den = A0 + B0; } else { num = A1; den = A0; }
    if (den == 0) return 1;
    return 1 - num / den;
}
static double d_ab_ochiai(const dist_in *d, int i, int j) { /* :1068-1078 */
    double A1 = M2(d->nb_shared, i, j), B1 = M2(d->nb_shared, j, i), A0 = NS(i), B0 = NS(j);
    if (A0 == 0 || B0 == 0) return 1;
    return 1 - sqrt(A1 / A0) * sqrt(B1 / B0);
}
static double d_ab_sorensen(const dist_in *d, int i, int j) { /* :1081-1096 */
    double A1 = M2(d->nb_shared, i, j), B1 = M2(d->nb_shared, j, i), A0 = NS(i), B0 = NS(j);
    double num = 2 * A1 * B1, den = A0 * B1 + A1 * B0;
    if (den == 0) return 1;
    return 1 - num / den;
}
static double d_ab_jaccard(const dist_in *d, int i, int j) { /* :1099-1115 */
    double A1 = M2(d->nb_shared, i, j), B1 = M2(d->nb_shared, j, i), A0 = NS(i), B0 = NS(j);
    double num = A1 * B1, den = A0 * B1 + A1 * B0 - A1 * B1;
    if (den == 0) return 1;
    return 1 - num / den;
}
static double d_pa_chord(double a, double b, double c) { /* :1117-1127 */
    double p1 = sqrt((a + b) * (a + c));
    if (p1 == 0) return sqrt(2);
    return sqrt(2 * (1 - a / p1));
}
static double d_pa_whittaker(double a, double b, double c) { /* :1129-1145 */
    if (a + b == 0 || a + c == 0) return 1;
    double p1 = b / (a + b), p2 = c / (a + c), p3 = a / (a + b), p4 = a / (a + c);
    return 0.5 * (p1 + p2 + fabs(p3 - p4));
}
static double d_pa_kulczynski(double a, double b, double c) { /* :1156-1170 */
    if (a + b == 0 || a + c == 0) return 1;
    double p1 = a / (a + b), p2 = a / (a + c);
    return 1 - 0.5 * (p1 + p2);
}
static double d_pa_braycurtis(double a, double b, double c) { /* :1172-1183 */
    if ((2 * a + b + c) == 0) return 1;
    return (b + c) / (2 * a + b + c);
}
static double d_pa_ochiai(double a, double b, double c) { /* :1188-1199; denominator rounded to float */
    float val = sqrt((a + b) * (a + c));
    if (val == 0) return 1;
    return 1 - (a / val);
}
static double d_pa_jaccard(double a, double b, double c) { /* :1201-1209 */
    if ((a + b + c) == 0) return 1;
    return (b + c) / (a + b + c);
}
static double d_pa_jaccard_simka(const dist_in *d, int i, int j, size_t idx, int sym) { /* :1211-1226 */
    double num, den;
    if (sym) { num = 2 * d->distinct_shared[idx]; den = d->nb_distinct[i] + d->nb_distinct[j]; }
    else { num = d->distinct_shared[idx]; den = d->nb_distinct[i]; }
    if (den == 0) return 1;
    return 1 - num / den;
}

/* order and names: S/core/SimkaDistance.cpp:617-648 */
static const char *const MATRIX_NAMES[] = {
    "mat_presenceAbsence_chord", "mat_presenceAbsence_whittaker", "mat_presenceAbsence_kulczynski",
    "mat_presenceAbsence_braycurtis", "mat_presenceAbsence_jaccard", "mat_presenceAbsence_simka-jaccard",
    "mat_presenceAbsence_simka-jaccard_asym", "mat_presenceAbsence_ochiai",
    "mat_abundance_simka-jaccard", "mat_abundance_simka-jaccard_asym", "mat_abundance_ab-ochiai",
    "mat_abundance_ab-sorensen", "mat_abundance_ab-jaccard", "mat_abundance_braycurtis", "mat_abundance_jaccard",
    "mat_abundance_chord", "mat_abundance_hellinger", "mat_abundance_kulczynski",
    "mat_abundance_whittaker", "mat_abundance_jensenshannon", "mat_abundance_canberra"};

static int matrix_id(int simple, int complex_, int m) {
    if (m < 15) return m;
    if (simple) { if (m < 18) return m; m -= 3; }
    if (complex_ && m >= 15 && m < 18) return m + 3;
    return -1;
}

static void compute_matrix(const dist_in *d, int id, float *out) {
    const int n = d->n;
    memset(out, 0, (size_t)n * n * sizeof(float)); /* createSquaredMatrix: zero diagonal */
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            uint64_t ua, ub, uc; get_abc(d, i, j, &ua, &ub, &uc);
            double a = (double)ua, b = (double)ub, c = (double)uc;
            size_t idx = sym_index(n, i, j);
            double v = 0, w = 0; int asym = 0;
            switch (id) {
            case 0: v = d_pa_chord(a, b, c); break;
            case 1: v = d_pa_whittaker(a, b, c); break;
            case 2: v = d_pa_kulczynski(a, b, c); break;
            case 3: v = d_pa_braycurtis(a, b, c); break;
            case 4: v = d_pa_jaccard(a, b, c); break;
            case 5: v = d_pa_jaccard_simka(d, i, j, idx, 1); break;
            case 6: v = d_pa_jaccard_simka(d, i, j, idx, 0); w = d_pa_jaccard_simka(d, j, i, idx, 0); asym = 1; break;
            case 7: v = d_pa_ochiai(a, b, c); break;
            case 8: v = d_ab_jaccard_simka(d, i, j, 1); break;
            case 9: v = d_ab_jaccard_simka(d, i, j, 0); w = d_ab_jaccard_simka(d, j, i, 0); asym = 1; break;
            case 10: v = d_ab_ochiai(d, i, j); break;
            case 11: v = d_ab_sorensen(d, i, j); break;
            case 12: v = d_ab_jaccard(d, i, j); break;
            case 13: case 14: v = d_ab_braycurtis(d, i, j); break;
            case 15: v = d_ab_chord(d, i, j); break;
            case 16: v = d_ab_hellinger(d, i, j); break;
            case 17: v = d_ab_kulczynski(d, i, j); break;
            case 18: v = d_ab_whittaker(d, i, j); break;
            case 19: v = d_ab_kl(d, i, j); break;
            case 20: v = d_ab_canberra(d, i, j); break;
            }
            out[(size_t)i * n + j] = (float)v;
            out[(size_t)j * n + i] = asym ? (float)w : (float)v;
        }
    if (id == 14) { /* computeJaccardDistanceFromBrayCurtis (S/core/SimkaDistance.hpp:463-475): from the FLOAT matrix */
        for (size_t q = 0; q < (size_t)n * n; q++) { double B = out[q]; double J = (2 * B) / (1 + B); out[q] = (float)J; }
    }
}

static void dist_in_from_run(const orc_run *r, dist_in *d, uint64_t **tmp_nd, uint64_t **tmp_nk) {
    const int n = r->n;
    d->n = n; d->simple = r->simple; d->complex_ = r->complex_;
    uint64_t *nd = malloc((size_t)n * 8), *nk = malloc((size_t)n * 8);
    d->chord_sqrt_n2 = malloc((size_t)n * sizeof(long double));
    for (int i = 0; i < n; i++) {
        nd[i] = r->samples[i].nb_distinct; nk[i] = r->samples[i].nb_kmers;
        d->chord_sqrt_n2[i] = sqrt((double)r->samples[i].chord_n2); /* sqrt(u64) -> double, stored long double */
    }
    d->nb_distinct = nd; d->nb_kmers = nk; *tmp_nd = nd; *tmp_nk = nk;
    d->distinct_shared = r->st.distinct_shared; d->braycurtis_num = r->st.braycurtis_num;
    d->nb_shared = r->st.nb_shared; d->chord_ninj = r->st.chord_ninj; d->hellinger = r->st.hellinger;
    d->kulczynski = r->st.kulczynski; d->canberra = r->st.canberra; d->whittaker = r->st.whittaker; d->kl = r->st.kl;
}

int orc_nb_matrices(const orc_run *r) { return 15 + (r->simple ? 3 : 0) + (r->complex_ ? 3 : 0); }
const char *orc_matrix_name(const orc_run *r, int m) {
    int id = matrix_id(r->simple, r->complex_, m);
    return id < 0 ? NULL : MATRIX_NAMES[id];
}
void orc_matrix(const orc_run *r, int m, float *out) {
    dist_in d; uint64_t *a, *b;
    dist_in_from_run(r, &d, &a, &b);
    compute_matrix(&d, matrix_id(r->simple, r->complex_, m), out);
    free(a); free(b); free(d.chord_sqrt_n2);
}

void orc_matrix_from_accumulators(const orc_accumulators *acc, int m, float *out) {
    const int n = acc->n; size_t nn = (size_t)n * n;
    dist_in d; memset(&d, 0, sizeof d);
    d.n = n; d.simple = acc->simple; d.complex_ = acc->complex_;
    d.nb_distinct = acc->nb_distinct; d.nb_kmers = acc->nb_kmers;
    d.chord_sqrt_n2 = malloc((size_t)n * sizeof(long double));
    for (int i = 0; i < n; i++) d.chord_sqrt_n2[i] = acc->chord_n2 ? sqrt((double)acc->chord_n2[i]) : 0;
    d.distinct_shared = acc->distinct_shared; d.braycurtis_num = acc->braycurtis_num; d.nb_shared = acc->nb_shared;
    long double *ch = calloc(nn, sizeof(long double)), *kl = calloc(nn, sizeof(long double));
    if (acc->chord_ninj) for (size_t i = 0; i < nn; i++) ch[i] = acc->chord_ninj[i];
    if (acc->kl) for (size_t i = 0; i < nn; i++) kl[i] = acc->kl[i];
    d.chord_ninj = ch; d.kl = kl;
    d.hellinger = acc->hellinger; d.kulczynski = acc->kulczynski; d.canberra = acc->canberra; d.whittaker = acc->whittaker;
    compute_matrix(&d, matrix_id(acc->simple, acc->complex_, m), out);
    free(ch); free(kl); free(d.chord_sqrt_n2);
}

/* S/core/SimkaDistance.cpp:653-699 */
size_t orc_matrix_csv(const orc_run *r, int m, const char *const *ids, char *buf, size_t cap) {
    const int n = r->n;
    float *mat = malloc((size_t)n * n * sizeof(float));
    orc_matrix(r, m, mat);
    size_t need = 0;
    char num[64];
#define PUT(s, l) do { if (need + (l) < cap) memcpy(buf + need, (s), (l)); need += (l); } while (0)
    for (int i = 0; i < n; i++) { PUT(";", 1); PUT(ids[i], strlen(ids[i])); }
    PUT("\n", 1);
    for (int i = 0; i < n; i++) {
        PUT(ids[i], strlen(ids[i]));
        for (int j = 0; j < n; j++) {
            int l = snprintf(num, sizeof num, ";%f", mat[(size_t)i * n + j]);
            PUT(num, (size_t)l);
        }
        PUT("\n", 1);
    }
#undef PUT
    if (need < cap) buf[need] = 0;
    free(mat);
    return need;
}

/* ------------------------------------------------------------------------------------------
 * SimkaMin distance: S/simkaMin/SimkaMinDistance.hpp:190-279 (computeDistance_unsynch)
 * ---------------------------------------------------------------------------------------- */
static void simkamin_pair(const uint64_t *h1, const uint32_t *c1, size_t n1,
                          const uint64_t *h2, const uint32_t *c2, size_t n2, uint32_t sketch_size,
                          float *jac, float *bc) {
    uint64_t nb_shared_distinct = 0, nb_distinct = 0, nb_shared_kmers = 0, nb_kmers = 0;
    /* sketchSize = min of the two (zero-stripped) sizes (:204); iterator = (next, isDone) (:73-101) */
    uint64_t ss = n1 < n2 ? n1 : n2;
    (void)sketch_size;
    if (n1 > 0 && n2 > 0) {
        size_t i = 0, j = 0; /* items already fetched from each side */
        uint64_t k1 = h1[i], k2 = h2[j]; uint32_t a1 = c1[i], a2 = c2[j]; i++; j++;
        while (nb_distinct < ss) {
            if (k1 > k2) {
                nb_distinct += 1; nb_kmers += a2;
                if (j >= n2) break;
                k2 = h2[j]; a2 = c2[j]; j++;
            } else if (k1 < k2) {
                nb_distinct += 1; nb_kmers += a1;
                if (i >= n1) break;
                k1 = h1[i]; a1 = c1[i]; i++;
            } else {
                nb_distinct += 1; nb_kmers += (uint64_t)a1 + a2;
                nb_shared_distinct += 1; nb_shared_kmers += a1 < a2 ? a1 : a2;
                if (j >= n2 || i >= n1) break;
                k1 = h1[i]; a1 = c1[i]; i++;
                k2 = h2[j]; a2 = c2[j]; j++;
            }
        }
    }
    if (nb_distinct == 0) *jac = 1; else
    *jac = (float)(1 - (long double)nb_shared_distinct / (long double)nb_distinct);
    if (nb_kmers == 0) *bc = 1; else
    *bc = (float)(1 - (long double)(2 * nb_shared_kmers) / (long double)nb_kmers);
}

void orc_simkamin_distance(const uint64_t *hashes, const uint32_t *counts, int n, uint32_t s,
                           float *jaccard, float *braycurtis, int nthreads) {
    size_t *start = malloc((size_t)n * sizeof(size_t));
    for (int i = 0; i < n; i++) { /* loadSketches strips the zero padding at the front (:521-586) */
        size_t q = 0; while (q < s && hashes[(size_t)i * s + q] == 0) q++;
        start[i] = q;
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    for (int i = 0; i < n; i++) { jaccard[(size_t)i * n + i] = 0; braycurtis[(size_t)i * n + i] = 0; }
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            float ja, b;
            simkamin_pair(hashes + (size_t)i * s + start[i], counts + (size_t)i * s + start[i], s - start[i],
                          hashes + (size_t)j * s + start[j], counts + (size_t)j * s + start[j], s - start[j], s, &ja, &b);
            jaccard[(size_t)i * n + j] = jaccard[(size_t)j * n + i] = ja;
            braycurtis[(size_t)i * n + j] = braycurtis[(size_t)j * n + i] = b;
        }
    free(start);
}

/* ------------------------------------------------------------------------------------------
 * SimkaMin sketch: src/simkaMin/SimkaMinCount.hpp:217-266 (SelectKmersCommand::operator()), :311-338
 * (processUnfiltered), :168-189 (heap -> ascending array, zero padding at the front), MurmurHash3_x64_128
 * (src/simkaMin/MurmurHash3.cpp:246-330; Austin Appleby's public-domain algorithm) on the 8 bytes of the
 * canonical k-mer value, out[0] kept.
 * Direct simulation of the reference's containers: a max-heap of the selected hashes + a hash -> count map.
 * Note the reference's order dependence: once the heap is full an occurrence of the CURRENT maximum is not
 * counted (the test is `hash < top`).
 * ---------------------------------------------------------------------------------------- */
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
uint64_t orc_murmur3_x64_128_u64(uint64_t key, uint32_t seed) {
    uint64_t h1 = seed, h2 = seed;
    uint64_t k1 = key;  /* tail of 8 bytes, little endian */
    k1 *= 0x87c37b91114253d5ull; k1 = rotl64(k1, 31); k1 *= 0x4cf5ad432745937full; h1 ^= k1;
    h1 ^= 8; h2 ^= 8;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

typedef struct { uint64_t *key; uint32_t *val; uint8_t *used; size_t cap, n; } u64map;
static size_t map_slot(const u64map *m, uint64_t k) { return (size_t)((k * 0x9E3779B97F4A7C15ull) >> 20) & (m->cap - 1); }
static long map_find(const u64map *m, uint64_t k) {
    size_t i = map_slot(m, k);
    while (m->used[i]) { if (m->key[i] == k) return (long)i; i = (i + 1) & (m->cap - 1); }
    return -1;
}
static void map_put(u64map *m, uint64_t k, uint32_t v) {
    size_t i = map_slot(m, k);
    while (m->used[i]) i = (i + 1) & (m->cap - 1);
    m->used[i] = 1; m->key[i] = k; m->val[i] = v; m->n++;
}
static void map_erase(u64map *m, uint64_t k) {  /* backward-shift deletion */
    long f = map_find(m, k);
    if (f < 0) return;
    size_t i = (size_t)f, j = i;
    m->used[i] = 0; m->n--;
    for (;;) {
        j = (j + 1) & (m->cap - 1);
        if (!m->used[j]) break;
        size_t h = map_slot(m, m->key[j]);
        if ((i <= j) ? (i < h && h <= j) : (i < h || h <= j)) continue;
        m->used[i] = 1; m->key[i] = m->key[j]; m->val[i] = m->val[j]; m->used[j] = 0;
        i = j;
    }
}
static void sk_heap_push(uint64_t *h, size_t *n, uint64_t v) {
    size_t i = (*n)++;
    h[i] = v;
    while (i > 0 && h[(i - 1) / 2] < h[i]) { uint64_t t = h[i]; h[i] = h[(i - 1) / 2]; h[(i - 1) / 2] = t; i = (i - 1) / 2; }
}
static void sk_heap_pop(uint64_t *h, size_t *n) {
    h[0] = h[--(*n)];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, b = i;
        if (l < *n && h[l] > h[b]) b = l;
        if (r < *n && h[r] > h[b]) b = r;
        if (b == i) break;
        uint64_t t = h[i]; h[i] = h[b]; h[b] = t; i = b;
    }
}

/* hashes_out / counts_out: sketch_size entries, ascending hashes, zero padding at the FRONT when the sample has
 * fewer distinct k-mers.  Returns the number of selected hashes. */
uint32_t orc_simkamin_sketch(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int k, uint32_t sketch_size,
                             uint32_t seed, uint64_t *hashes_out, uint32_t *counts_out) {
    u64map m;
    m.cap = 16; while (m.cap < (size_t)sketch_size * 4 + 16) m.cap <<= 1;
    m.key = calloc(m.cap, 8); m.val = calloc(m.cap, 4); m.used = calloc(m.cap, 1); m.n = 0;
    uint64_t *heap = malloc(((size_t)sketch_size + 1) * 8);
    size_t hn = 0;
    uint64_t *canon = NULL; uint8_t *valid = NULL; size_t cap = 0;
    /* A read shorter than k: SelectKmersCommand refills its k-mer vector with Model::build, which leaves it untouched, and walks
     * it again (src/simkaMin/SimkaMinCount.hpp:217-263; GC/kmer/impl/Model.hpp:603-610) -- the previous long-enough read is counted
     * once more.  (Single-thread order: the reference hands groups of 1000 reads to its threads, each with its own vector.) */
    int64_t last = -1;
    for (uint64_t r0 = 0; r0 < n_reads; r0++) {
        uint64_t r = r0;
        if ((size_t)(read_off[r0 + 1] - read_off[r0]) < (size_t)k) { if (last < 0) continue; r = (uint64_t)last; } else last = (int64_t)r0;
        const size_t len = (size_t)(read_off[r + 1] - read_off[r]);
        if (len > cap) { cap = len * 2; canon = realloc(canon, cap * 8); valid = realloc(valid, cap); }
        const size_t nk = orc_canonical_kmers(seqs + read_off[r], len, k, canon, valid);
        for (size_t i = 0; i < nk; i++) {
            if (!valid[i]) continue;
            const uint64_t h = orc_murmur3_x64_128_u64(canon[i], seed);
            if (hn < sketch_size) {                       /* SimkaMinCount.hpp:313-322 */
                long f = map_find(&m, h);
                if (f < 0) { sk_heap_push(heap, &hn, h); map_put(&m, h, 1); } else m.val[f] += 1;
            } else if (h < heap[0]) {                      /* :324-336 */
                long f = map_find(&m, h);
                if (f < 0) { map_erase(&m, heap[0]); sk_heap_pop(heap, &hn); sk_heap_push(heap, &hn, h); map_put(&m, h, 1); }
                else m.val[f] += 1;
            }
        }
    }
    const uint32_t n_sel = (uint32_t)hn;
    for (uint32_t i = 0; i < sketch_size; i++) { hashes_out[i] = 0; counts_out[i] = 0; }
    for (size_t i = 0; i < n_sel; i++) {                   /* :174-189: pop the maximum into the last free cell */
        const uint64_t h = heap[0];
        hashes_out[sketch_size - 1 - i] = h;
        counts_out[sketch_size - 1 - i] = m.val[map_find(&m, h)];
        sk_heap_pop(heap, &hn);
    }
    free(heap); free(m.key); free(m.val); free(m.used); free(canon); free(valid);
    return n_sel;
}

/* The `-filter` variant: processFiltered (src/simkaMin/SimkaMinCount.hpp:341-380).  A k-mer that passes the heap test
 * (heap not full, or hash < top) is first remembered in a Bloom filter; from its second passing occurrence on it lives in
 * the heap / map with a count that starts at 2.  The reference's filter is a BloomCacheCoherent of max-memory * 8 / cores
 * bits with 7 hash functions (:1155-1161): 6.4e10 bits at the wrapper's default memory, where a false positive is
 * out of reach for any input that fits a machine, so this restatement remembers the k-mers EXACTLY (a set of hashes; the
 * 64-bit MurmurHash3 value stands for the k-mer, as it does in the reference's own count map).  It diverges from the
 * reference only where the reference's Bloom filter would have answered a false positive (memory-starved runs). */
uint32_t orc_simkamin_sketch_filtered(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int k, uint32_t sketch_size,
                                      uint32_t seed, uint64_t *hashes_out, uint32_t *counts_out) {
    u64map m, seen;
    m.cap = 16; while (m.cap < (size_t)sketch_size * 4 + 16) m.cap <<= 1;
    m.key = calloc(m.cap, 8); m.val = calloc(m.cap, 4); m.used = calloc(m.cap, 1); m.n = 0;
    seen.cap = 1024; seen.key = calloc(seen.cap, 8); seen.val = calloc(seen.cap, 4); seen.used = calloc(seen.cap, 1); seen.n = 0;
    uint64_t *heap = malloc(((size_t)sketch_size + 1) * 8);
    size_t hn = 0;
    uint64_t *canon = NULL; uint8_t *valid = NULL; size_t cap = 0;
    /* A read shorter than k: SelectKmersCommand refills its k-mer vector with Model::build, which leaves it untouched, and walks
     * it again (src/simkaMin/SimkaMinCount.hpp:217-263; GC/kmer/impl/Model.hpp:603-610) -- the previous long-enough read is counted
     * once more.  (Single-thread order: the reference hands groups of 1000 reads to its threads, each with its own vector.) */
    int64_t last = -1;
    for (uint64_t r0 = 0; r0 < n_reads; r0++) {
        uint64_t r = r0;
        if ((size_t)(read_off[r0 + 1] - read_off[r0]) < (size_t)k) { if (last < 0) continue; r = (uint64_t)last; } else last = (int64_t)r0;
        const size_t len = (size_t)(read_off[r + 1] - read_off[r]);
        if (len > cap) { cap = len * 2; canon = realloc(canon, cap * 8); valid = realloc(valid, cap); }
        const size_t nk = orc_canonical_kmers(seqs + read_off[r], len, k, canon, valid);
        for (size_t i = 0; i < nk; i++) {
            if (!valid[i]) continue;
            const uint64_t h = orc_murmur3_x64_128_u64(canon[i], seed);
            const int full = hn >= sketch_size;
            if (full && !(h < heap[0])) continue;               /* :362 */
            if (map_find(&seen, h) >= 0) {                       /* _bloomFilter->contains(kmer) :344,363 */
                long f = map_find(&m, h);
                if (f >= 0) m.val[f] += 1;
                else {
                    if (full) { map_erase(&m, heap[0]); sk_heap_pop(heap, &hn); }
                    sk_heap_push(heap, &hn, h);
                    map_put(&m, h, 2);
                }
            } else {                                             /* _bloomFilter->insert(kmer) :355,376 */
                if ((seen.n + 1) * 2 > seen.cap) {               /* grow the exact set */
                    u64map g; g.cap = seen.cap * 2; g.key = calloc(g.cap, 8); g.val = calloc(g.cap, 4); g.used = calloc(g.cap, 1); g.n = 0;
                    for (size_t q = 0; q < seen.cap; q++) if (seen.used[q]) map_put(&g, seen.key[q], 0);
                    free(seen.key); free(seen.val); free(seen.used);
                    seen = g;
                }
                map_put(&seen, h, 0);
            }
        }
    }
    const uint32_t n_sel = (uint32_t)hn;
    for (uint32_t i = 0; i < sketch_size; i++) { hashes_out[i] = 0; counts_out[i] = 0; }
    for (size_t i = 0; i < n_sel; i++) {
        const uint64_t h = heap[0];
        hashes_out[sketch_size - 1 - i] = h;
        counts_out[sketch_size - 1 - i] = m.val[map_find(&m, h)];
        sk_heap_pop(heap, &hn);
    }
    free(heap); free(m.key); free(m.val); free(m.used); free(seen.key); free(seen.val); free(seen.used); free(canon); free(valid);
    return n_sel;
}

/* ------------------------------------------------------------------------------------------
 * FASTA / FASTQ (.gz) reader: GC/bank/impl/BankFasta.cpp:485-572 (kseq-style)
 * ---------------------------------------------------------------------------------------- */
typedef struct { gzFile f; unsigned char buf[1 << 16]; int beg, end, eof; } gzbuf;
static int gb_getc(gzbuf *b) {
    if (b->beg >= b->end) {
        if (b->eof) return -1;
        b->beg = 0; b->end = gzread(b->f, b->buf, sizeof b->buf);
        if (b->end < (int)sizeof b->buf) b->eof = 1;
        if (b->end <= 0) return -1;
    }
    return b->buf[b->beg++];
}
static void push_char(char **buf, uint64_t *len, uint64_t *cap, char c) {
    if (*len + 1 > *cap) { *cap = *cap ? *cap * 2 : (1 << 20); *buf = realloc(*buf, *cap); }
    (*buf)[(*len)++] = c;
}
/* appends the rest of the current line (without '\n', a trailing '\r' dropped when the total is >1 chars) */
static int gb_getline_append(gzbuf *b, char **buf, uint64_t *len, uint64_t *cap, uint64_t start) {
    int c, any = 0;
    while ((c = gb_getc(b)) != -1 && c != '\n') { push_char(buf, len, cap, (char)c); any = 1; }
    if (*len - start > 1 && (*buf)[*len - 1] == '\r') (*len)--;
    return (c == -1 && !any) ? -1 : 0;
}

int64_t orc_read_fastx(const char *path, char **buf, uint64_t *buf_len, uint64_t *buf_cap,
                       uint64_t **off, uint64_t *noff, uint64_t *off_cap) {
    gzbuf *b = calloc(1, sizeof *b);
    b->f = gzopen(path, "rb");
    if (!b->f) { free(b); return -1; }
    int64_t nreads = 0;
    int last = 0, c;
    if (*noff == 0) {
        if (*off_cap < 1) { *off_cap = 1024; *off = realloc(*off, *off_cap * 8); }
        (*off)[(*noff)++] = *buf_len;
    }
    for (;;) {
        if (last == 0) {
            while ((c = gb_getc(b)) != -1 && c != '>' && c != '@') ;
            if (c == -1) break;
            last = c;
        }
        /* header line: skipped */
        while ((c = gb_getc(b)) != -1 && c != '\n') ;
        if (c == -1 && b->beg >= b->end && b->eof) { /* header without sequence at EOF: still a (empty) read */ }
        uint64_t start = *buf_len;
        while ((c = gb_getc(b)) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            push_char(buf, buf_len, buf_cap, (char)c);
            gb_getline_append(b, buf, buf_len, buf_cap, start);
        }
        last = (c == '>' || c == '@') ? c : 0;
        if (c == '+') {
            uint64_t rlen = *buf_len - start, qlen = 0;
            while ((c = gb_getc(b)) != -1 && c != '\n') ;
            /* quality lines until at least rlen chars have been read */
            while (c != -1) {
                int any = 0;
                while ((c = gb_getc(b)) != -1 && c != '\n') { qlen++; any = 1; }
                if (c == -1 && !any) break;
                if (qlen >= rlen) break;
            }
            last = 0;
        }
        if (*noff + 1 > *off_cap) { *off_cap = *off_cap * 2; *off = realloc(*off, *off_cap * 8); }
        (*off)[(*noff)++] = *buf_len;
        nreads++;
        if (c == -1 && last == 0) break;
    }
    gzclose(b->f);
    free(b);
    return nreads;
}
