/*
 * simka_b200.h -- C ABI of the B200-native Simka hot path (libsimka_b200.so).
 *
 * The reference (GATB/simka) has no in-process plugin ABI for this path: its seams are the
 * simkaCount / simkaMerge process CLIs plus temp files, and the GATB-core C++ template
 * interfaces.  This header is the boundary a reference maintainer binds to
 * replace those stages in-process; every entry point names the reference code it replaces
 * (paths relative to the reference root; GC/ = thirdparty/gatb-core/gatb-core/src/gatb/).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * SIMKA_ERR_* code and never throws; the caller owns all host buffers; the library owns all
 * device memory; one context per GPU; a context is not thread-safe.  There is NO CPU fallback:
 * if no CUDA device is usable, simka_gpu_create fails with SIMKA_ERR_CUDA.
 */
#ifndef SIMKA_B200_H
#define SIMKA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMKA_OK 0
#define SIMKA_ERR_INVALID (-1)   /* bad argument / call order                         */
#define SIMKA_ERR_CUDA (-2)      /* CUDA runtime failure (see simka_gpu_last_error)   */
#define SIMKA_ERR_OVERFLOW (-3)  /* an on-chip table overflowed (should never happen) */
#define SIMKA_ERR_NOMEM (-4)
#define SIMKA_ERR_IO (-5)

typedef struct simka_ctx simka_ctx;

/* Replaces the argv of simkaCount / simkaMerge (src/SimkaCount.cpp:80-98, src/SimkaMerge.cpp:1530-1544):
 * -kmer-size, -nb-datasets, -abundance-min/-max, -simple-dist, -complex-dist. */
typedef struct {
    int32_t kmer_size;        /* 2..63: span 32 (k <= 31, one word) and span 64 (k <= 63, LargeInt<2>) */
    int32_t nb_samples;       /* N, 1..8192                                                        */
    int64_t abundance_min;    /* keep a k-mer of a sample iff min <= count <= max                  */
    int64_t abundance_max;    /*   (src/minikc/MiniKC.hpp:56)                                       */
    int32_t simple_dist;      /* chord / hellinger / kulczynski accumulators                       */
    int32_t complex_dist;     /* whittaker / jensen-shannon / canberra accumulators                */
    int32_t device;           /* CUDA device ordinal                                               */
    int32_t shard_rank;       /* multi-GPU: this context keeps hash-buckets rank of shard_count    */
    int32_t shard_count;      /*   (the reference's one-merge-process-per-partition split,         */
                              /*    src/SimkaPotara.hpp:991-1058); 0 or 1 = no sharding            */
    int32_t reserved;
    uint64_t expected_kmers_per_sample; /* sizing hint for the merge granularity, 0 = derive       */
    int32_t sketch_mode;      /* 1: count MurmurHash3_x64_128(canonical k-mer, sketch_seed) instead of    */
    uint32_t sketch_seed;     /*    the k-mer itself (SimkaMin sketches; stage 1 only)                    */
} simka_params;

/* Replaces the 4 lines of count_synchro/<id>.ok (src/SimkaCount.cpp:303-317,355-368). */
typedef struct {
    uint64_t nb_reads;
    uint64_t nb_distinct;     /* distinct solid k-mers         */
    uint64_t nb_kmers;        /* sum of solid abundances       */
    uint64_t chord_n2;        /* sum of squared abundances     */
} simka_sample_totals;

/* Replaces SimkaStatistics (src/core/SimkaDistance.hpp:68-139).  Host arrays owned by the caller.
 * sym = N(N+1)/2 entries, idx = j + (N-1) i - i(i-1)/2 for i<j; sq = N*N row-major.
 * Arrays of a disabled distance family may be NULL. */
typedef struct {
    int32_t nb_samples, simple_dist, complex_dist, reserved;
    uint64_t nb_distinct_kmers, nb_shared_kmers;  /* _nbDistinctKmers, _nbSharedKmers              */
    simka_sample_totals *totals;                  /* [N]                                           */
    uint64_t *distinct_shared;                    /* sym  _matrixNbDistinctSharedKmers             */
    uint64_t *braycurtis_num;                     /* sym  _brayCurtisNumerator                     */
    uint64_t *nb_shared;                          /* sq   _matrixNbSharedKmers                     */
    uint64_t *chord_ninj;                         /* sq   _chord_NiNj (integer-valued)             */
    uint64_t *hellinger;                          /* sq   _hellinger_SqrtNiNj                      */
    uint64_t *kulczynski;                         /* sq   _kulczynski_minNiNj (upper triangle)     */
    uint64_t *canberra;                           /* sq   _canberra                                */
    uint64_t *whittaker;                          /* sq   _whittaker_minNiNj                       */
    double *kullback_leibler;                     /* sq   _kullbackLeibler                         */
} simka_stats_view;

/* ---- life cycle -------------------------------------------------------------------------- */
int simka_gpu_create(simka_ctx **ctx, const simka_params *params);
void simka_gpu_destroy(simka_ctx *ctx);
/* forget all samples and accumulators but keep the device work buffers (next run of the same shape) */
int simka_gpu_reset(simka_ctx *ctx);
const char *simka_gpu_last_error(const simka_ctx *ctx);
/* pinned host buffers for the streaming path (cudaHostAlloc / cudaFreeHost) */
int simka_gpu_host_alloc(void **ptr, size_t bytes);
int simka_gpu_host_free(void *ptr);

/* ---- stage 1: replaces `simkaCount` for one sample -----------------------------------------
 * (src/SimkaCount.cpp:174-353 -> SortingCountAlgorithm, GC/kmer/impl/SortingCountAlgorithm.cpp:600-653,
 *  -> SimkaCompressedProcessor::process, src/minikc/MiniKC.hpp:54-79)
 *
 * Reads of a sample arrive as a 2-bit packed base stream + a 1-bit/base "bad" mask:
 *   base i of the stream lives in packed[i/32] at bits [62-2*(i%32), 63-2*(i%32)] (first base most
 *   significant, code A=0 C=1 T=2 G=3 = (ascii>>1)&3, GC/tools/misc/api/Data.hpp:179);
 *   bad[i/32] bit (31-i%32) is 1 for any char outside ACGTacgt and for the ONE separator position
 *   that must follow every read; positions >= n_bases of the last word are bad.
 * Both arrays must have n_words+1 entries (the last entry: packed 0, bad 0xFFFFFFFF).
 * simka_pack_reads() below produces exactly this layout.  A sample may be pushed in several
 * calls (k-mers never span two pushes).  The copy is asynchronous when the buffers are pinned;
 * they must stay valid until simka_gpu_finish_sample(sample_id) returns.  Two samples may be in
 * flight: pushing sample s+1 before finishing sample s overlaps its copy with the kernels of s. */
int simka_gpu_push_reads(simka_ctx *ctx, int sample_id, const uint64_t *packed, const uint32_t *bad,
                         uint64_t n_words, uint64_t n_reads);
/* Same layout, but the buffers already live in device memory of this context's GPU (zero copy; they
 * must stay valid and unchanged until simka_gpu_finish_sample returns). */
int simka_gpu_push_reads_device(simka_ctx *ctx, int sample_id, const uint64_t *d_packed, const uint32_t *d_bad,
                                uint64_t n_words, uint64_t n_reads);
/* Same, from concatenated ASCII reads (read r = seqs[read_off[r] .. read_off[r+1])); the 2-bit
 * packing runs on the GPU (replaces the per-char encode of GC/kmer/impl/Model.hpp:726-765). */
int simka_gpu_push_ascii(simka_ctx *ctx, int sample_id, const char *seqs, const uint64_t *read_off,
                         uint64_t n_reads);
/* Counts the pushed reads of the sample: canonical k-mers, exact multiset count, abundance filter,
 * per-sample totals.  The filtered (k-mer,count) list stays resident in HBM. */
int simka_gpu_finish_sample(simka_ctx *ctx, int sample_id, simka_sample_totals *out);

/* parity hook: the sample's solid k-mers in increasing k-mer order with their counts, i.e. the
 * decoded content of the solid/part_P/__p__I.gz files (src/SimkaCount.cpp:248-257).  Call with kmers=NULL to
 * get the size.  Only this shard's k-mers are returned when sharded. */
int simka_gpu_dump_sample_counts(simka_ctx *ctx, int sample_id, uint64_t *kmers, uint32_t *counts, uint64_t *n);
/* the same for 32 <= kmer_size <= 63 (Kmer<64>, LargeInt<2>: GC/tools/math/LargeInt2.pri): k-mers as (low word, high word),
 * ascending as 128-bit integers */
int simka_gpu_dump_sample_counts128(simka_ctx *ctx, int sample_id, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *counts, uint64_t *n);

/* ---- barrier between the stages: global per-sample totals -----------------------------------
 * (SimkaStatistics ctor reads count_synchro/<id>.ok of ALL samples, src/core/SimkaDistance.cpp:116-151)
 * Single GPU: nothing to do.  Sharded: all-reduce(sum) the three counters over the shards
 * (nb_reads is per sample, not summed) and hand the global values back before the merge. */
int simka_gpu_get_totals(simka_ctx *ctx, simka_sample_totals *out /* [N] */);
int simka_gpu_set_totals(simka_ctx *ctx, const simka_sample_totals *in /* [N] */);

/* ---- multi-GPU, sample-sharded counting: the exchange between the stages ---------------------------
 * The reference runs one simkaCount process per SAMPLE and one simkaMerge process per PARTITION; what
 * crosses between them is the per-sample, per-partition list of solid (k-mer, count) records
 * (solid/part_P/__p__I.gz, src/SimkaCount.cpp:248-257 -> src/SimkaMerge.cpp:1082-1103).  Same here: a rank
 * counts its own samples (all hash buckets), exports each sample's list -- grouped by "fine range" = the
 * top simka_gpu_fine_bits() bits of the hashed k-mer, offsets in d_off[2^bits+1] -- and the rank that owns a
 * contiguous block of fine ranges imports that slice of EVERY sample (any transport: NCCL all-to-all in
 * simka_b200/dist.py) into a second context, sets the global totals and runs simka_gpu_merge_and_accumulate.
 * Exported pointers are device memory owned by the context; imported arrays are device memory owned by the
 * caller and must stay valid until the merge has run.  d_off of an import is re-based to the slice
 * (d_off[0] = 0) and has 2^fine_bits+1 entries for the slice's own fine_bits. */
int simka_gpu_fine_bits(simka_ctx *ctx);
int simka_gpu_export_sample(simka_ctx *ctx, int sample_id, const uint64_t **d_kmers, const uint32_t **d_counts,
                            const uint32_t **d_off, uint64_t *n);
int simka_gpu_import_sample(simka_ctx *ctx, int sample_id, const uint64_t *d_kmers, const uint32_t *d_counts, uint64_t n,
                            const uint32_t *d_off, int fine_bits, uint64_t nb_reads);

/* ---- multi-GPU inside the library: NCCL over NVLink ------------------------------------------------
 * One context per GPU (one process per GPU under torchrun / mpirun, or one host thread per GPU in the `simka -nb-gpus` command
 * line).  Rank r counts the samples s with s % world == r through the calls above; then ONE call does everything the reference
 * does between and after its two stages (src/SimkaPotara.hpp:991-1058 barrier + merge jobs, :1161-1174 sum of the
 * per-partition statistics): the per-sample totals become global (ncclAllReduce), every sample's list slice goes straight from
 * the list it was counted into to the rank that owns that block of hash ranges (ncclSend / ncclRecv in one group, no staging
 * copy), that rank merges its block over all samples, and the N x N accumulators are summed over the ranks (ncclAllReduce:
 * u64 block, f64 block).  Afterwards simka_gpu_get_stats(ctx) returns the GLOBAL statistics on every rank.
 * NCCL is loaded at run time (libnccl.so.2); without it these entries fail with SIMKA_ERR_INVALID. */
#define SIMKA_COMM_ID_BYTES 128
int simka_comm_unique_id(uint8_t id[SIMKA_COMM_ID_BYTES]);                       /* rank 0; hand the bytes to every rank */
int simka_gpu_comm_init(simka_ctx *ctx, const uint8_t id[SIMKA_COMM_ID_BYTES], int rank, int world);
int simka_gpu_exchange_and_merge(simka_ctx *ctx);
/* device time of the last simka_gpu_exchange_and_merge, in ms: [0] totals + sizes all-reduce, [1] list exchange, [2] merge, [3] accumulator all-reduce */
int simka_gpu_exchange_timings(simka_ctx *ctx, float ms[4]);

/* Registers a sample from its solid (canonical k-mer, count) list in HOST memory -- the decoded content of the
 * reference-format files solid/part_P/__p__I.gz (src/SimkaCount.cpp:248-257, records read by
 * src/SimkaMerge.cpp:1082-1103) -- instead of counting its reads.  Any order; counts already abundance-filtered.
 * Used by the simkaMerge executable of this build and to resume from kept temp files. */
int simka_gpu_import_counts(simka_ctx *ctx, int sample_id, const uint64_t *kmers, const uint32_t *counts, uint64_t n,
                            uint64_t nb_reads);

/* ---- stage 2: replaces `simkaMerge` over all of this shard's buckets ---------------------------
 * (src/SimkaMerge.cpp:1164-1326 k-way merge + insert gate, SimkaCountProcessorSimple::process /
 *  updateDistance{Default,Simple,Complex}, src/core/SimkaAlgorithm.hpp:200-516) */
int simka_gpu_merge_and_accumulate(simka_ctx *ctx);
/* Accumulators of this shard; element-wise sum over shards == the reference's mainStats
 * (src/SimkaPotara.hpp:1161-1174, src/core/SimkaDistance.cpp:156-213). */
int simka_gpu_get_stats(simka_ctx *ctx, simka_stats_view *out);

/* device-side time per kernel, in ms, summed over the samples since create/reset (CUDA events on
 * the context's stream, recorded around each launch) */
typedef struct {
    float extract_count_ms, extract_scatter_ms; /* K1 canonical k-mer extraction: count pass, scatter pass   */
    float part_count_ms, part_scatter_ms;       /* K2 level-2 partition: count pass, scatter pass            */
    float bin_count_ms;                         /* K3 per-bin exact count + filter                            */
    float compact_ms;                           /* survivors -> per-sample list                               */
    float merge_ms;                             /* K4 merge + N x N accumulate (+ Whittaker one-sided)        */
    float samples_ms;                           /* stage 1 total (first to last kernel of every sample)       */
    float total_ms;                             /* samples_ms + merge_ms                                      */
    uint64_t kmers_in;                          /* valid k-mer occurrences kept by this shard                 */
    uint64_t entries;                           /* solid (sample,k-mer) entries after the abundance filter    */
    uint64_t distinct_kmers;                    /* distinct solid k-mers over all samples                     */
    uint64_t pairs;                             /* sample pairs enumerated by the accumulate kernel           */
    uint64_t kernel_launches;
    uint64_t exact_retries;                     /* samples redone with the exact two-pass K1 (a bucket overflowed)  */
    uint64_t superk_samples;                    /* samples counted through the experimental super-k-mer front end
                                                 * (SIMKA_B200_SUPERK=1; 0 on the default path)                      */
    uint64_t superk_fallbacks;                  /* ... that overflowed the record list and were redone per k-mer    */
} simka_timings;
int simka_gpu_get_timings(simka_ctx *ctx, simka_timings *out);
/* the CUDA stream all of this context's kernels run on (for external CUDA-event timing) */
void *simka_gpu_stream(simka_ctx *ctx);

/* ---- host helpers (no GPU needed) ---------------------------------------------------------------- */

/* words needed for reads of the given total length (one separator per read + pad word) */
uint64_t simka_pack_words(uint64_t total_bases, uint64_t n_reads);
/* 2-bit pack + bad mask in the layout simka_gpu_push_reads expects; multi-threaded.
 * packed/bad must hold simka_pack_words()+1 entries.  Returns the number of words used. */
uint64_t simka_pack_reads(const char *seqs, const uint64_t *read_off, uint64_t n_reads,
                          uint64_t *packed, uint32_t *bad, int nthreads);
/* Same, for a given -kmer-size: with k <= 15 the reference counts with MiniKC, which does NOT skip windows containing a
 * non-ACGT character (src/SimkaCount.cpp:283-288, src/minikc/MiniKC.hpp:219-231; the character contributes its
 * (c>>1)&3 code), so only the read separators are marked bad.  simka_pack_reads == kmer_size 31.
 * simka_gpu_push_ascii applies the same rule from the context's kmer_size. */
uint64_t simka_pack_reads_k(const char *seqs, const uint64_t *read_off, uint64_t n_reads,
                            uint64_t *packed, uint32_t *bad, int nthreads, int kmer_size);

/* k <= 15 only -- the second MiniKC quirk (pinned on reference runs, tests/golden/truth_ref/example_n_results_k{12,15}_*):
 * MiniKC::count (src/minikc/MiniKC.hpp:204-231) refills its k-mer iterator through Model::Iterator::setData -> Model::build
 * (GC/kmer/impl/Model.hpp:603-610,499-513), which leaves the k-mer vector untouched for a read shorter than k, so the
 * previous long-enough read of the sample is counted once more.  This returns the read set with every such short read
 * replaced by a copy of that previous read (a short read holds no k-mer of its own): pack / push THAT for k <= 15.
 * *seqs_out == NULL on return means "nothing to replace, use the input" (always the case for kmer_size > 15);
 * otherwise both outputs are malloc'ed (simka_host_free).  The host readers of the command lines do the same while parsing. */
int simka_minikc_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out,
                        uint64_t **read_off_out);
/* SimkaMin has the same stale-vector behaviour for EVERY k: SelectKmersCommand refills its k-mer vector with Model::build, which
 * leaves it untouched for a read shorter than k, and then walks it again (src/simkaMin/SimkaMinCount.hpp:217-263): the previous
 * long-enough read is counted once more.  Same contract as simka_minikc_replay, for any kmer_size; apply it to the reads of a
 * sample before packing them for simkamin_gpu_sketch.  (The reference hands the reads to its threads in groups of 1000 and every
 * thread keeps its own vector, so for a short read that STARTS such a group the replayed read depends on the thread schedule;
 * this is the single-thread order, `-nb-cores 1`.) */
int simkamin_short_read_replay(const char *seqs, const uint64_t *read_off, uint64_t n_reads, int kmer_size, char **seqs_out, uint64_t **read_off_out);

/* Replaces SimkaDistance + SimkaStatistics::outputMatrix (src/core/SimkaDistance.cpp:603-649,920-1226).
 * Matrix order = the reference's output order: 15 default, +3 if simple, +3 if complex. */
int simka_nb_matrices(const simka_stats_view *stats);
const char *simka_matrix_name(const simka_stats_view *stats, int m);
int simka_distance_matrix(const simka_stats_view *stats, int m, float *out /* N*N */);
/* CSV text of dumpMatrix (src/core/SimkaDistance.cpp:653-699).  Returns the byte count needed. */
size_t simka_matrix_csv(const simka_stats_view *stats, int m, const char *const *ids, char *buf, size_t cap);
/* writes <dir>/<name>.csv.gz for every matrix, like outputMatrix */
int simka_write_matrices(const simka_stats_view *stats, const char *const *ids, const char *out_dir);

/* ---- SimkaMin sketch of one sample: replaces SelectKmersCommand (src/simkaMin/SimkaMinCount.hpp:217-338) --------
 * Reads in the packed layout of simka_pack_reads (host memory).  hashes_out / counts_out: sketch_size entries, the
 * bottom-s distinct MurmurHash3_x64_128(canonical k-mer, seed) values in ascending order with their counts, zero
 * padding at the FRONT when the sample has fewer distinct k-mers (:174-189); *n_selected = entries that are not
 * padding.  Reproduces the reference's order dependence: occurrences of the largest selected hash are counted only
 * until the sketch is complete (the reference tests `hash < top`). */
int simkamin_gpu_sketch(int device, int kmer_size, uint32_t sketch_size, uint32_t seed, const uint64_t *packed, const uint32_t *bad,
                        uint64_t n_words, uint64_t n_reads, uint64_t *hashes_out, uint32_t *counts_out, uint32_t *n_selected);
/* the same with the reference's -filter option (processFiltered, src/simkaMin/SimkaMinCount.hpp:341-380): only k-mers seen at
 * least twice are sketched, counts start at 2 on the second occurrence.  The reference remembers first occurrences in a Bloom
 * filter (BloomCacheCoherent of max-memory * 8 / cores bits, 7 hashes, :1155-1161); this entry counts exactly, i.e. it is
 * identical to the reference wherever that filter gives no false positive (it reproduces the reference's 8 -filter truth sets). */
int simkamin_gpu_sketch_filter(int device, int kmer_size, uint32_t sketch_size, uint32_t seed, const uint64_t *packed, const uint32_t *bad,
                        uint64_t n_words, uint64_t n_reads, uint64_t *hashes_out, uint32_t *counts_out, uint32_t *n_selected);

/* How the calling thread's last simkamin_gpu_sketch / simkamin_gpu_sketch_filter call got its sketch.  The bottom-s selection of
 * the reference is a heap over every k-mer of the sample (processUnfiltered, src/simkaMin/SimkaMinCount.hpp:311-338); here only
 * hashes below a bound are extracted and counted, the bound growing from 16 * s / positions of the hash range by factors of 8
 * until >= s distinct hashes lie below it (exact: every occurrence of a kept hash is counted).  attempts = counting passes run,
 * keys_counted = k-mer occurrences that passed the bound over all attempts, distinct_dumped = entries the last attempt
 * brought back to the host (the counted list is grouped by ranges of the hash: only the leading ranges, up to the one holding
 * the s-th smallest hash, are copied and sorted).  SIMKA_B200_SKETCH_PREFILTER=0 takes the whole hash range at once. */
typedef struct simkamin_sketch_info {
    uint64_t attempts, keys_counted, distinct_dumped;
} simkamin_sketch_info;
int simkamin_gpu_last_sketch_info(simkamin_sketch_info *out);
/* The sketch entries keep one stage-1 context (streams, events, device work buffers) between calls with the same device, k, seed
 * and filter, so that a run over many datasets does not allocate and free them per dataset (calls are serialised on it;
 * SIMKA_B200_SKETCH_CTX_CACHE=0: a fresh context per call).  This frees it. */
void simkamin_gpu_release_cache(void);

/* ---- SimkaMin: replaces ComputeDistanceManager::computeDistance_unsynch over all pairs ------------
 * (src/simkaMin/SimkaMinDistance.hpp:190-303).  hashes/counts: n sketches of s entries, ascending,
 * zero hashes at the front are padding (loadSketches :521-586).  Outputs n*n float32 row-major. */
int simkamin_gpu_distance(int device, const uint64_t *hashes, const uint32_t *counts, uint32_t n,
                          uint32_t s, float *jaccard, float *braycurtis, float *kernel_ms);

/* ---- hooks over the host-side mirror of the GATB input surface (simka_b200/csrc/host/*.hpp) ----------------
 * Kmer<32>::ModelCanonical / ModelDirect iteration (GC/kmer/impl/Model.hpp:726-765): k-mer values + validity of
 * every window; returns the number of windows.  ModelMinimizer::getMinimizerValue (:1144-1148).
 * SimkaSequenceFilter::getShannonIndex (src/core/SimkaCommons.hpp:398-432).  Bank::open + Iterator<Sequence>
 * over a fasta / fastq / gz / album uri: concatenated reads + offsets, malloc'ed (simka_host_free). */
uint64_t simka_host_kmers(const char *seq, uint64_t len, int k, int canonical, uint64_t *out, uint8_t *valid);
uint64_t simka_host_minimizer(uint64_t kmer, int k, int m);
float simka_host_shannon_index(const char *seq, uint64_t len);
/* Repartitor (GC/kmer/impl/PartiInfo.cpp:228-261) of a reference run, read from its config.h5 (S/SimkaPotara.hpp:573-606,767): the
 * partition table[minimizer(kmer)] of n canonical k-mers.  Returns the number of partitions, -1 if the file cannot be read. */
int simka_host_repartition(const char *config_h5, const uint64_t *kmers, uint64_t n, int k, uint16_t *parts_out);
int64_t simka_host_read_bank(const char *uri, char **seqs, uint64_t **read_off);
/* IBank::estimateNbItems (GC/bank/impl/BankFasta.cpp:728-775, AbstractBank.hpp:71-74): sequences at the head of every file
 * extrapolated to the file size (a gz file = 4 x its compressed size); what `simka -max-reads 0` derives its limit from. */
int64_t simka_host_estimate_nb_items(const char *uri);
void simka_host_free(void *p);

/* The two SimkaMin distance entry points keep their device buffers between calls (one call at a time); this frees them. */
int simkamin_gpu_release(void);

/* all n1 x n2 pairs between two sets of sketches: the rectangular blocks of `simkaMinCore distance -start-i/-start-j`
 * (src/simkaMin/SimkaMinDistance.hpp:642-753).  Outputs n1*n2 float32 row-major. */
int simkamin_gpu_distance_rect(int device, const uint64_t *hashes1, const uint32_t *counts1, uint32_t n1, const uint64_t *hashes2,
                               const uint32_t *counts2, uint32_t n2, uint32_t s, float *jaccard, float *braycurtis, float *kernel_ms);

const char *simka_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
