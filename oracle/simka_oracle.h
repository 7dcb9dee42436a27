/*
 * simka_oracle.h -- CPU restatement of the Simka hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for simka_b200: a plain-C restatement of the reference's
 * algorithm (reads -> canonical k-mer counting -> abundance filter -> multi-sample merge
 * -> N x N accumulators -> distance matrices -> CSV text).  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The shipped path (simka_b200/csrc) never links or calls anything in here.
 *
 * Parity pin: reproduces byte-for-byte the 4 x 20 truth matrices of the reference's own
 * end-to-end tests (tests/truth/results_k{21,31}_t{0,2} of the reference, copied as data
 * fixtures to tests/golden/truth/), the TestKmer known-answer vectors and the
 * count_synchro totals recorded in SURVEY.md 8(c).  See tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root; GC/ = thirdparty/gatb-core/gatb-core/src/gatb/).
 */
#ifndef SIMKA_ORACLE_H
#define SIMKA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- k-mer model (GC/kmer/impl/Model.hpp:637-657,726-765,858-884; Data.hpp:179-183) ---- */

/* Writes the len-k+1 canonical k-mers of `seq` (min(fwd, revcomp) as unsigned ints) and their
 * validity (1 = all k chars in ACGTacgt).  Returns the number of windows (0 if len < k). */
size_t orc_canonical_kmers(const char *seq, size_t len, int k, uint64_t *canon, uint8_t *valid);
/* Same but forward-strand (ModelDirect) values. */
size_t orc_direct_kmers(const char *seq, size_t len, int k, uint64_t *fwd, uint8_t *valid);
uint64_t orc_revcomp(uint64_t x, int k); /* GC/tools/math/LargeInt1.pri:137-155 */
/* minimizer of a canonical k-mer, m-mers with the AA ban (GC/kmer/impl/Model.hpp:1010-1070,1220-1287) */
uint32_t orc_minimizer(uint64_t kmer_fwd, int k, int m);

/* ---- run handle --------------------------------------------------------------------- */

typedef struct orc_run orc_run;

/* Sequences of all samples are given concatenated in `seqs`; read r spans
 * [read_off[r], read_off[r+1]); sample s owns reads [sample_read_off[s], sample_read_off[s+1]).
 * abundance filter: keep count c iff amin <= c <= amax  (src/minikc/MiniKC.hpp:54-56).
 * nthreads <= 0: all cores.  Returns NULL on allocation failure. */
orc_run *orc_run_create(const char *seqs, const uint64_t *read_off, const uint64_t *sample_read_off,
                        int nsamples, int k, int64_t amin, int64_t amax, int simple, int complex_,
                        int nthreads);
/* stage 1 only (count + filter) / stage 2 only (merge + accumulate); orc_run_create does not compute */
void orc_run_count(orc_run *r);
void orc_run_merge(orc_run *r);
void orc_run_destroy(orc_run *r);

/* stage-1 outputs (src/SimkaCount.cpp:303-317): per sample sorted (kmer,count) + 4 totals */
size_t orc_sample_size(const orc_run *r, int s);
const uint64_t *orc_sample_kmers(const orc_run *r, int s);
const uint64_t *orc_sample_kmers_hi(const orc_run *r, int s);   /* high words of the k-mers, k > 31 only (else NULL) */
size_t orc_canonical_kmers128(const char *seq, size_t len, int k, uint64_t *lo, uint64_t *hi, uint8_t *valid);
const uint32_t *orc_sample_counts(const orc_run *r, int s);
/* out[0..3] = nbReads, nbDistinct, sum c, sum c^2 */
void orc_sample_totals(const orc_run *r, int s, uint64_t out[4]);

/* stage-2 accumulators (src/core/SimkaDistance.hpp:86-125).  Matrices are N*N row-major,
 * symmetric-index vectors have N(N+1)/2 entries, idx = j + (N-1) i - i(i-1)/2.
 * names: "distinct_shared" (sym), "braycurtis_num" (sym), "nb_shared" (N*N),
 *        "chord_ninj", "hellinger", "kulczynski" (N*N, simple), "canberra", "whittaker" (N*N, complex) */
int orc_get_u64(const orc_run *r, const char *name, uint64_t *out);
/* "kl" (N*N, complex), "chord_ninj" as double */
int orc_get_f64(const orc_run *r, const char *name, double *out);
uint64_t orc_nb_distinct_kmers(const orc_run *r);
uint64_t orc_nb_shared_kmers(const orc_run *r);

/* ---- distances (src/core/SimkaDistance.cpp:603-649,920-1226) -------------------------- */

/* number of matrices that outputMatrix would write (15, +3 simple, +3 complex) */
int orc_nb_matrices(const orc_run *r);
const char *orc_matrix_name(const orc_run *r, int m);     /* e.g. "mat_abundance_braycurtis" */
void orc_matrix(const orc_run *r, int m, float *out);     /* N*N floats */
/* CSV text exactly as SimkaStatistics::dumpMatrix (src/core/SimkaDistance.cpp:653-699);
 * ids = N NUL-terminated strings.  Returns bytes written (excluding NUL), or needed size if cap too small */
size_t orc_matrix_csv(const orc_run *r, int m, const char *const *ids, char *buf, size_t cap);

/* ---- standalone finalisation from raw accumulators (used to cross-check the product's host
 *      finaliser): layout identical to orc_get_* ------------------------------------------------ */
typedef struct {
    int n, simple, complex_;
    const uint64_t *nb_reads, *nb_distinct, *nb_kmers, *chord_n2;       /* per sample */
    const uint64_t *distinct_shared, *braycurtis_num;                   /* sym */
    const uint64_t *nb_shared;                                          /* N*N */
    const uint64_t *chord_ninj, *hellinger, *kulczynski;                /* N*N (simple) */
    const uint64_t *canberra, *whittaker;                               /* N*N (complex) */
    const double *kl;                                                   /* N*N (complex) */
} orc_accumulators;
void orc_matrix_from_accumulators(const orc_accumulators *a, int m, float *out);

/* ---- SimkaMin distance (src/simkaMin/SimkaMinDistance.hpp:190-279) -------------------- */
/* sketches: n x s (hash u64, count u32) in two arrays; leading zero hashes are padding
 * (loadSketches :521-586 strips them).  Writes n*n float32 Jaccard and Bray-Curtis. */
/* SimkaMin sketch of one sample (src/simkaMin/SimkaMinCount.hpp:217-338): ascending bottom-s MurmurHash3 values of the
 * valid canonical k-mers with their counts, zero padding at the front.  Returns the number of selected hashes. */
uint64_t orc_murmur3_x64_128_u64(uint64_t key, uint32_t seed);
uint32_t orc_simkamin_sketch(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int k, uint32_t sketch_size,
                             uint32_t seed, uint64_t *hashes_out, uint32_t *counts_out);
uint32_t orc_simkamin_sketch_filtered(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int k, uint32_t sketch_size,
                             uint32_t seed, uint64_t *hashes_out, uint32_t *counts_out);
void orc_simkamin_distance(const uint64_t *hashes, const uint32_t *counts, int n, uint32_t s,
                           float *jaccard, float *braycurtis, int nthreads);

/* ---- input files ---------------------------------------------------------------------- */
/* Parse one FASTA/FASTQ(.gz) file with the kseq-style rules of GC/bank/impl/BankFasta.cpp:485-572.
 * Appends sequences to *buf (realloc'd) and offsets to *off.  Returns #reads or -1. */
int64_t orc_read_fastx(const char *path, char **buf, uint64_t *buf_len, uint64_t *buf_cap,
                       uint64_t **off, uint64_t *noff, uint64_t *off_cap);

#ifdef __cplusplus
}
#endif
#endif
