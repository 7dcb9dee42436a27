// simka_host.cpp -- host-side pieces of the C ABI that need no GPU: read packing, the distance
// closed forms and the CSV writer.  The distance formulas are the cold N^2 tail of the path
// (SURVEY.md 3.4); they are evaluated on the host in the reference's own precision (long double
// where it uses long double) so that the %f output is byte-identical.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simka_b200.h"
#include "host/bank.hpp"
#include "host/repartitor.hpp"
#include "host/kmer_model.hpp"
#include "host/simka_input.hpp"

namespace {

inline size_t sym_index(size_t n, size_t i, size_t j) { return j + (n - 1) * i - (i * (i - 1) / 2); }  // src/core/SimkaAlgorithm.hpp:364

// ---- SimkaDistance (src/core/SimkaDistance.cpp:920-1226) -----------------------------------------
struct Dist {
    const simka_stats_view &s;
    size_t n;
    explicit Dist(const simka_stats_view &v) : s(v), n((size_t)v.nb_samples) {}
    uint64_t D(size_t i) const { return s.totals[i].nb_distinct; }   // _nbSolidDistinctKmersPerBank
    uint64_t NS(size_t i) const { return s.totals[i].nb_kmers; }     // _nbSolidKmersPerBank
    uint64_t sq(const uint64_t *m, size_t i, size_t j) const { return m ? m[i * n + j] : 0; }
    void abc(size_t i, size_t j, double &a, double &b, double &c) const {  // get_abc :920-926
        uint64_t ua = s.distinct_shared[sym_index(n, i, j)];
        a = (double)ua; b = (double)(D(i) - ua); c = (double)(D(j) - ua);
    }
    double ab_braycurtis(size_t i, size_t j) const {  // :928-939
        double union_ = NS(i) + NS(j);
        if (union_ == 0) return 1;
        double inter = 2 * s.braycurtis_num[sym_index(n, i, j)];
        return 1 - inter / union_;
    }
    double ab_chord(size_t i, size_t j) const {  // :942-959; _chord_sqrt_N2 = sqrt(u64) kept as long double (:132)
        long double si = sqrt((double)s.totals[i].chord_n2), sj = sqrt((double)s.totals[j].chord_n2);
        double den = si * sj;
        if (den == 0) return sqrt(2);
        long double ninj = (long double)sq(s.chord_ninj, i, j);
        long double v = sqrtl(2 - 2 * ninj / den);
        return (double)v;
    }
    double ab_hellinger(size_t i, size_t j) const {  // :962-972
        double union_ = sqrt((double)NS(i)) * sqrt((double)NS(j));
        if (union_ == 0) return sqrt(2);
        double inter = 2 * sq(s.hellinger, i, j);
        return sqrt(2 - (inter / union_));
    }
    double ab_whittaker(size_t i, size_t j) const {  // :988-998
        long double union_ = NS(i) * NS(j);
        if (union_ == 0) return 1;
        long double inter = sq(s.whittaker, i, j);
        double w = 0.5 * (inter / union_);
        return w;
    }
    double ab_kl(size_t i, size_t j) const {  // :1001-1007
        long double v = s.kullback_leibler ? (long double)s.kullback_leibler[i * n + j] : 0;
        if (v == 0) return 1;
        return sqrt(0.5 * v);
    }
    double ab_canberra(size_t i, size_t j) const {  // :1010-1021
        double a, b, c; abc(i, j, a, b, c);
        if ((a + b + c) == 0) return 1;
        return (1 / (a + b + c)) * sq(s.canberra, i, j);
    }
    double ab_kulczynski(size_t i, size_t j) const {  // :1024-1038 ([j][i] is never written by the update)
        if (NS(i) == 0 || NS(j) == 0) return 1;
        long double n1 = (double)sq(s.kulczynski, i, j) / (double)NS(i);
        long double n2 = (double)sq(s.kulczynski, j, i) / (double)NS(j);
        double v = 1 - 0.5 * (n1 + n2);
        return v;
    }
    double ab_jaccard_simka(size_t i, size_t j, bool sym) const {  // :1041-1065
        double A1 = sq(s.nb_shared, i, j), B1 = sq(s.nb_shared, j, i), A0 = NS(i), B0 = NS(j);
        double num = sym ? A1 + B1 : A1, den = sym ? A0 + B0 : A0;
        if (den == 0) return 1;
        return 1 - num / den;
    }
    double ab_ochiai(size_t i, size_t j) const {  // :1068-1078
        double A1 = sq(s.nb_shared, i, j), B1 = sq(s.nb_shared, j, i), A0 = NS(i), B0 = NS(j);
        if (A0 == 0 || B0 == 0) return 1;
        return 1 - sqrt(A1 / A0) * sqrt(B1 / B0);
    }
    double ab_sorensen(size_t i, size_t j) const {  // :1081-1096
        double A1 = sq(s.nb_shared, i, j), B1 = sq(s.nb_shared, j, i), A0 = NS(i), B0 = NS(j);
        double num = 2 * A1 * B1, den = A0 * B1 + A1 * B0;
        if (den == 0) return 1;
        return 1 - num / den;
    }
    double ab_jaccard(size_t i, size_t j) const {  // :1099-1115
        double A1 = sq(s.nb_shared, i, j), B1 = sq(s.nb_shared, j, i), A0 = NS(i), B0 = NS(j);
        double num = A1 * B1, den = A0 * B1 + A1 * B0 - A1 * B1;
        if (den == 0) return 1;
        return 1 - num / den;
    }
    static double pa_chord(double a, double b, double c) {  // :1117-1127
        double p1 = sqrt((a + b) * (a + c));
        if (p1 == 0) return sqrt(2);
        return sqrt(2 * (1 - a / p1));
    }
    static double pa_whittaker(double a, double b, double c) {  // :1129-1145
        if (a + b == 0 || a + c == 0) return 1;
        double p1 = b / (a + b), p2 = c / (a + c), p3 = a / (a + b), p4 = a / (a + c);
        return 0.5 * (p1 + p2 + fabs(p3 - p4));
    }
    static double pa_kulczynski(double a, double b, double c) {  // :1156-1170
        if (a + b == 0 || a + c == 0) return 1;
        return 1 - 0.5 * (a / (a + b) + a / (a + c));
    }
    static double pa_braycurtis(double a, double b, double c) {  // :1172-1183
        if ((2 * a + b + c) == 0) return 1;
        return (b + c) / (2 * a + b + c);
    }
    static double pa_ochiai(double a, double b, double c) {  // :1188-1199 (float denominator)
        float val = sqrt((a + b) * (a + c));
        if (val == 0) return 1;
        return 1 - (a / val);
    }
    static double pa_jaccard(double a, double b, double c) {  // :1201-1209
        if ((a + b + c) == 0) return 1;
        return (b + c) / (a + b + c);
    }
    double pa_jaccard_simka(size_t i, size_t j, size_t idx, bool sym) const {  // :1211-1226
        double num = sym ? 2 * s.distinct_shared[idx] : s.distinct_shared[idx];
        double den = sym ? D(i) + D(j) : D(i);
        if (den == 0) return 1;
        return 1 - num / den;
    }
};

// order and names of outputMatrix (src/core/SimkaDistance.cpp:617-648)
const char *const NAMES[21] = {
    "mat_presenceAbsence_chord", "mat_presenceAbsence_whittaker", "mat_presenceAbsence_kulczynski",
    "mat_presenceAbsence_braycurtis", "mat_presenceAbsence_jaccard", "mat_presenceAbsence_simka-jaccard",
    "mat_presenceAbsence_simka-jaccard_asym", "mat_presenceAbsence_ochiai",
    "mat_abundance_simka-jaccard", "mat_abundance_simka-jaccard_asym", "mat_abundance_ab-ochiai",
    "mat_abundance_ab-sorensen", "mat_abundance_ab-jaccard", "mat_abundance_braycurtis", "mat_abundance_jaccard",
    "mat_abundance_chord", "mat_abundance_hellinger", "mat_abundance_kulczynski",
    "mat_abundance_whittaker", "mat_abundance_jensenshannon", "mat_abundance_canberra"};

int matrix_id(const simka_stats_view *s, int m) {
    if (m < 0) return -1;
    if (m < 15) return m;
    if (s->simple_dist) { if (m < 18) return m; m -= 3; }
    if (s->complex_dist && m >= 15 && m < 18) return m + 3;
    return -1;
}

}  // namespace

extern "C" {

int simka_nb_matrices(const simka_stats_view *s) { return s ? 15 + (s->simple_dist ? 3 : 0) + (s->complex_dist ? 3 : 0) : 0; }

const char *simka_matrix_name(const simka_stats_view *s, int m) {
    int id = s ? matrix_id(s, m) : -1;
    return id < 0 ? nullptr : NAMES[id];
}

int simka_distance_matrix(const simka_stats_view *s, int m, float *out) {
    if (!s || !out || !s->totals || !s->distinct_shared || !s->braycurtis_num || !s->nb_shared) return SIMKA_ERR_INVALID;
    const int id = matrix_id(s, m);
    if (id < 0) return SIMKA_ERR_INVALID;
    Dist d(*s);
    const size_t n = d.n;
    memset(out, 0, n * n * sizeof(float));
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++) {
            double a, b, c; d.abc(i, j, a, b, c);
            const size_t idx = sym_index(n, i, j);
            double v = 0, w = 0; bool asym = false;
            switch (id) {
            case 0: v = Dist::pa_chord(a, b, c); break;
            case 1: v = Dist::pa_whittaker(a, b, c); break;
            case 2: v = Dist::pa_kulczynski(a, b, c); break;
            case 3: v = Dist::pa_braycurtis(a, b, c); break;
            case 4: v = Dist::pa_jaccard(a, b, c); break;
            case 5: v = d.pa_jaccard_simka(i, j, idx, true); break;
            case 6: v = d.pa_jaccard_simka(i, j, idx, false); w = d.pa_jaccard_simka(j, i, idx, false); asym = true; break;
            case 7: v = Dist::pa_ochiai(a, b, c); break;
            case 8: v = d.ab_jaccard_simka(i, j, true); break;
            case 9: v = d.ab_jaccard_simka(i, j, false); w = d.ab_jaccard_simka(j, i, false); asym = true; break;
            case 10: v = d.ab_ochiai(i, j); break;
            case 11: v = d.ab_sorensen(i, j); break;
            case 12: v = d.ab_jaccard(i, j); break;
            case 13: case 14: v = d.ab_braycurtis(i, j); break;
            case 15: v = d.ab_chord(i, j); break;
            case 16: v = d.ab_hellinger(i, j); break;
            case 17: v = d.ab_kulczynski(i, j); break;
            case 18: v = d.ab_whittaker(i, j); break;
            case 19: v = d.ab_kl(i, j); break;
            case 20: v = d.ab_canberra(i, j); break;
            }
            out[i * n + j] = (float)v;
            out[j * n + i] = asym ? (float)w : (float)v;
        }
    if (id == 14)  // computeJaccardDistanceFromBrayCurtis works on the float matrix (src/core/SimkaDistance.hpp:463-475)
        for (size_t q = 0; q < n * n; q++) { double B = out[q]; double J = (2 * B) / (1 + B); out[q] = (float)J; }
    return SIMKA_OK;
}

size_t simka_matrix_csv(const simka_stats_view *s, int m, const char *const *ids, char *buf, size_t cap) {
    if (!s || !ids) return 0;
    const size_t n = (size_t)s->nb_samples;
    std::vector<float> mat(n * n);
    if (simka_distance_matrix(s, m, mat.data()) != SIMKA_OK) return 0;
    std::string str;
    for (size_t i = 0; i < n; i++) { str += ";"; str += ids[i]; }
    str += '\n';
    char num[64];
    for (size_t i = 0; i < n; i++) {
        str += ids[i];
        for (size_t j = 0; j < n; j++) { snprintf(num, sizeof num, ";%f", mat[i * n + j]); str += num; }
        str += '\n';
    }
    if (buf && cap > str.size()) memcpy(buf, str.c_str(), str.size() + 1);
    return str.size();
}

int simka_write_matrices(const simka_stats_view *s, const char *const *ids, const char *out_dir) {
    if (!s || !ids || !out_dir) return SIMKA_ERR_INVALID;
    const int nm = simka_nb_matrices(s);
    for (int m = 0; m < nm; m++) {
        size_t need = simka_matrix_csv(s, m, ids, nullptr, 0);
        std::vector<char> buf(need + 1);
        simka_matrix_csv(s, m, ids, buf.data(), need + 1);
        std::string path = std::string(out_dir) + "/" + simka_matrix_name(s, m) + ".csv.gz";
        gzFile f = gzopen(path.c_str(), "wb");
        if (!f) return SIMKA_ERR_IO;
        gzwrite(f, buf.data(), (unsigned)need);
        gzclose(f);
    }
    return SIMKA_OK;
}

// ---- read packing ----------------------------------------------------------------------------------

// 32 characters -> 32 two-bit codes (first character most significant) + 32 bad bits, with AVX2 + BMI2 when the CPU has them
// (run-time dispatch): the bytes are reversed so that movemask / pext, which put byte 0 in the low bits, deliver the first
// character on top; code = (c >> 1) & 3 is bits 1-2 of every byte, gathered by pext; bad = not in ACGTacgt.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define SIMKA_PACK_SIMD 1
__attribute__((target("avx2,bmi2"))) static inline void pack32_avx2(const unsigned char *sp, uint64_t *v, uint32_t *bb) {
    const __m256i rev16 = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(sp));
    x = _mm256_shuffle_epi8(x, rev16);
    x = _mm256_permute2x128_si256(x, x, 1);  // byte i = character 31 - i
    const __m256i u = _mm256_and_si256(x, _mm256_set1_epi8((char)0xDF));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
    *bb = ~(uint32_t)_mm256_movemask_epi8(ok);
    const uint64_t m = 0x0606060606060606ull;
    *v = _pext_u64((uint64_t)_mm256_extract_epi64(x, 0), m) | (_pext_u64((uint64_t)_mm256_extract_epi64(x, 1), m) << 16) |
         (_pext_u64((uint64_t)_mm256_extract_epi64(x, 2), m) << 32) | (_pext_u64((uint64_t)_mm256_extract_epi64(x, 3), m) << 48);
}
static bool pack_simd_available() {
    static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !getenv("SIMKA_B200_PACK_SCALAR");
    return ok;
}
#else
#define SIMKA_PACK_SIMD 0
#endif

uint64_t simka_pack_words(uint64_t total_bases, uint64_t n_reads) { return (total_bases + n_reads + 31) / 32; }

uint64_t simka_pack_reads(const char *seqs, const uint64_t *read_off, uint64_t n_reads, uint64_t *packed, uint32_t *bad, int nthreads) {
    return simka_pack_reads_k(seqs, read_off, n_reads, packed, bad, nthreads, 31);
}

uint64_t simka_pack_reads_k(const char *seqs, const uint64_t *read_off, uint64_t n_reads, uint64_t *packed, uint32_t *bad, int nthreads,
                            int kmer_size) {
    // k <= 15: the reference counts with MiniKC, which does not skip windows that contain a non-ACGT character
    // (src/minikc/MiniKC.hpp:219-231): only the read separators are marked bad
    const bool badchars = kmer_size > 15;
    const uint64_t base0 = n_reads ? read_off[0] : 0;
    const uint64_t total = n_reads ? read_off[n_reads] - base0 : 0;
    const uint64_t n_words = simka_pack_words(total, n_reads);
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n_words / 4096 + 1) nthreads = (int)(n_words / 4096 + 1);
    auto work = [&](uint64_t w0, uint64_t w1) {
        // stream position p = (read_off[r]-base0) + r + o; find the read containing position 32*w0
        const uint64_t p0 = w0 * 32;
        uint64_t r = 0;
        if (n_reads) {
            uint64_t lo = 0, hi = n_reads;  // start(lo) <= p0 < start(hi)
            if (total + n_reads <= p0) r = n_reads;
            else {
                while (hi - lo > 1) { uint64_t mid = (lo + hi) >> 1; if (read_off[mid] - base0 + mid <= p0) lo = mid; else hi = mid; }
                r = lo;
            }
        }
        uint64_t rbeg = 0, rend = 0, o = 0;
        if (r < n_reads) { rbeg = read_off[r]; rend = read_off[r + 1]; o = p0 - (rbeg - base0 + r); }
        // bad[c] = 1 for every character outside ACGTacgt (GC/tools/misc/api/Data.hpp:179-183)
        static const struct BadTab {
            uint8_t t[256];
            BadTab() { for (int c = 0; c < 256; c++) { int u = c & 0xDF; t[c] = !(u == 'A' || u == 'C' || u == 'G' || u == 'T'); } }
        } bt_all;
        static const struct NoBadTab { uint8_t t[256]; NoBadTab() { memset(t, 0, sizeof t); } } bt_none;
        const uint8_t *btt = badchars ? bt_all.t : bt_none.t;
        uint64_t w = w0, acc = 0;
        uint32_t bacc = 0;
        int cnt = 0;  // positions already in (acc, bacc)
#if SIMKA_PACK_SIMD
        const bool simd = pack_simd_available();
        const unsigned char *const simd_end = reinterpret_cast<const unsigned char *>(seqs) + (n_reads ? read_off[n_reads] : 0);  // loads of 32 bytes stay below
#endif
        while (w < w1) {
            if (r >= n_reads) {  // past the last read: bad padding
                acc <<= 2 * (32 - cnt);
                bacc = cnt ? ((bacc << (32 - cnt)) | (0xFFFFFFFFu >> cnt)) : 0xFFFFFFFFu;
                packed[w] = acc; bad[w] = bacc; w++;
                acc = 0; bacc = 0; cnt = 0;
                continue;
            }
            if (rbeg + o >= rend) {  // the separator position after read r
                acc <<= 2; bacc = (bacc << 1) | 1u; cnt++;
                r++;
                if (r < n_reads) { rbeg = rend; rend = read_off[r + 1]; }
                o = 0;
            } else {  // a run of bases of read r: up to the end of the current word
                const unsigned char *sp = reinterpret_cast<const unsigned char *>(seqs) + rbeg + o;
                uint64_t avail = rend - (rbeg + o);
#if SIMKA_PACK_SIMD
                if (simd && sp + 32 <= simd_end) {  // 32 characters at once, of which the first `take` belong to this word
                    const int take = (int)std::min<uint64_t>(avail, (uint64_t)(32 - cnt));
                    uint64_t v; uint32_t bb;
                    pack32_avx2(sp, &v, &bb);
                    if (!badchars) bb = 0;
                    if (take == 32) { acc = v; bacc = bb; }
                    else { acc = (acc << (2 * take)) | (v >> (2 * (32 - take))); bacc = (bacc << take) | (bb >> (32 - take)); }
                    cnt += take;
                    o += take;
                    if (cnt == 32) { packed[w] = acc; bad[w] = bacc; w++; acc = 0; bacc = 0; cnt = 0; }
                    continue;
                }
#endif
                if (cnt == 0 && avail >= 32) {  // fast path: a whole word inside the read
                    uint64_t v = 0; uint32_t bb = 0;
                    for (int j = 0; j < 32; j++) { const unsigned c = sp[j]; v = (v << 2) | ((c >> 1) & 3u); bb = (bb << 1) | btt[c]; }
                    packed[w] = v; bad[w] = bb; w++;
                    o += 32;
                    continue;
                }
                const int take = (int)std::min<uint64_t>(avail, (uint64_t)(32 - cnt));
                for (int j = 0; j < take; j++) { const unsigned c = sp[j]; acc = (acc << 2) | ((c >> 1) & 3u); bacc = (bacc << 1) | btt[c]; }
                cnt += take;
                o += take;
            }
            if (cnt == 32) { packed[w] = acc; bad[w] = bacc; w++; acc = 0; bacc = 0; cnt = 0; }
        }
    };
    if (nthreads == 1) work(0, n_words);
    else {
        std::vector<std::thread> th;
        const uint64_t per = (n_words + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; t++) {
            uint64_t w0 = (uint64_t)t * per, w1 = std::min<uint64_t>(n_words, w0 + per);
            if (w0 < w1) th.emplace_back(work, w0, w1);
        }
        for (auto &t : th) t.join();
    }
    packed[n_words] = 0;
    bad[n_words] = 0xFFFFFFFFu;
    return n_words;
}


static int short_read_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out, uint64_t **read_off_out);

int simka_minikc_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out, uint64_t **read_off_out) {
    if (!seqs_out || !read_off_out || (n_reads && (!seqs || !read_off))) return SIMKA_ERR_INVALID;
    *seqs_out = nullptr; *read_off_out = nullptr;
    if (kmer_size > 15 || kmer_size < 1) return SIMKA_OK;
    return short_read_replay(seqs, read_off, n_reads, kmer_size, seqs_out, read_off_out);
}

int simkamin_short_read_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out, uint64_t **read_off_out) {
    if (!seqs_out || !read_off_out || (n_reads && (!seqs || !read_off))) return SIMKA_ERR_INVALID;
    *seqs_out = nullptr; *read_off_out = nullptr;
    if (kmer_size < 1) return SIMKA_OK;
    return short_read_replay(seqs, read_off, n_reads, kmer_size, seqs_out, read_off_out);
}

static int short_read_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out, uint64_t **read_off_out) try {
    const uint64_t k = (uint64_t)kmer_size;
    bool any = false;
    uint64_t total = 0, last = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint64_t len = read_off[r + 1] - read_off[r];
        if (len < k) { any = any || last != len; total += last; } else { last = len; total += len; }
    }
    if (!any) return SIMKA_OK;
    char *out = (char *)malloc(total ? total : 1);
    uint64_t *off = (uint64_t *)malloc((n_reads + 1) * 8);
    if (!out || !off) { free(out); free(off); return SIMKA_ERR_NOMEM; }
    uint64_t pos = 0, lbeg = 0, llen = 0;
    off[0] = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint64_t len = read_off[r + 1] - read_off[r];
        if (len >= k) { lbeg = read_off[r]; llen = len; }
        memcpy(out + pos, seqs + lbeg, llen);  // a short read: the previous long-enough read again (nothing before the first one)
        pos += llen;
        off[r + 1] = pos;
    }
    *seqs_out = out; *read_off_out = off;
    return SIMKA_OK;
} catch (...) { return SIMKA_ERR_NOMEM; }

// ---- hooks over the host-side mirror classes (tests/test_host_cpu.py compares them with the oracle) ----------

uint64_t simka_host_kmers(const char *seq, uint64_t len, int k, int canonical, uint64_t *out, uint8_t *valid) {
    using simka_b200::Kmer;
    try {
        if (canonical) {
            Kmer<>::ModelCanonical model((size_t)k);
            Kmer<>::ModelCanonical::Iterator it(model);
            it.setData(seq, (size_t)len);
            uint64_t n = 0;
            for (it.first(); !it.isDone(); it.next(), n++) { out[n] = it->value().getVal(); valid[n] = it->isValid(); }
            return n;
        }
        Kmer<>::ModelDirect model((size_t)k);
        std::vector<Kmer<>::KmerDirect> v;
        model.build(seq, (size_t)len, v);
        for (size_t i = 0; i < v.size(); i++) { out[i] = v[i].value().getVal(); valid[i] = v[i].isValid(); }
        return v.size();
    } catch (...) { return 0; }
}

uint64_t simka_host_minimizer(uint64_t kmer, int k, int m) {
    try {
        simka_b200::Kmer<>::ModelMinimizer<> model((size_t)k, (size_t)m);
        return model.getMinimizerValue(simka_b200::Kmer<>::Type(kmer));
    } catch (...) { return ~0ull; }
}

// partitions of canonical k-mers under the repartition table of a reference run's config.h5 (host/repartitor.hpp); returns the
// number of partitions, -1 when the file cannot be read as one
int simka_host_repartition(const char *config_h5, const uint64_t *kmers, uint64_t n, int k, uint16_t *parts_out) {
    try {
        const simka_b200::Repartitor r = simka_b200::Repartitor::load(config_h5);
        simka_b200::Kmer<>::ModelMinimizer<> model((size_t)k, (size_t)r.minimizer_size);
        for (uint64_t i = 0; i < n; i++) parts_out[i] = r.table[(size_t)model.getMinimizerValue(simka_b200::Kmer<>::Type(kmers[i]))];
        return (int)r.nbpart;
    } catch (...) { return -1; }
}

float simka_host_shannon_index(const char *seq, uint64_t len) { return simka_b200::SimkaSequenceFilter::getShannonIndex(seq, (size_t)len); }

// reads a bank (fasta / fastq / gz / album) through Bank::open + Iterator<Sequence>; malloc'ed outputs
int64_t simka_host_read_bank(const char *uri, char **seqs, uint64_t **read_off) {
    using namespace simka_b200::bank;
    try {
        std::unique_ptr<IBank> b(Bank::open(uri));
        std::unique_ptr<Iterator<Sequence>> it(b->iterator());
        std::string all;
        std::vector<uint64_t> off(1, 0);
        for (it->first(); !it->isDone(); it->next()) {
            all.append(it->item().getDataBuffer() ? it->item().getDataBuffer() : "", it->item().getDataSize());
            off.push_back(all.size());
        }
        *seqs = (char *)malloc(all.size() + 1);
        *read_off = (uint64_t *)malloc(off.size() * 8);
        if (!*seqs || !*read_off) return -1;
        memcpy(*seqs, all.data(), all.size());
        memcpy(*read_off, off.data(), off.size() * 8);
        return (int64_t)off.size() - 1;
    } catch (...) { return -1; }
}

// IBank::estimateNbItems of a bank (fasta / fastq / gz / album): the reference's head-of-file extrapolation, -1 on error
int64_t simka_host_estimate_nb_items(const char *uri) {
    using namespace simka_b200::bank;
    try {
        std::unique_ptr<IBank> b(Bank::open(uri));
        return b->estimateNbItems();
    } catch (...) { return -1; }
}

void simka_host_free(void *p) { free(p); }

}  // extern "C"
