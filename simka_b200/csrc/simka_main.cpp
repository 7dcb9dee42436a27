// simka_main.cpp -- the `simka` command line on top of libsimka_b200.so.
//
// Same flags, input-file grammar, output file names and CSV format as the reference's `simka`
// (src/core/Simka.cpp:25-117, src/SimkaPotara.cpp:36-53, src/core/SimkaDistance.cpp:603-699); the per-sample
// simkaCount processes and per-partition simkaMerge processes of the reference (src/SimkaPotara.hpp:813-1124)
// are replaced by in-process calls of the C ABI: every sample is parsed on the host (IBank surface, bank.hpp),
// 2-bit packed into pinned chunks and streamed to the GPU while the previous chunk is being counted.
// Errors print "EXCEPTION: <msg>" and return EXIT_FAILURE like the reference's main (src/SimkaPotara.cpp:147-163).
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/simka_b200.h"
#include "host/bank.hpp"
#include "host/simka_input.hpp"

using namespace simka_b200;

namespace {

struct Options {
    std::string in, out = "./simka_results", outTmp;
    bool keepTmp = false, dataInfo = false, simple = false, complex_ = false;
    int kmerSize = 21;
    long long abundanceMin = 2, abundanceMax = 999999999;
    long long maxReads = -1;
    size_t minReadSize = 0;
    double minShannon = 0;
    int nbCores = 0, verbose = 1, device = 0, nbGpus = 1;
    long long maxMemory = 5000;
};

const char *kOneParam[] = {"-in", "-out", "-out-tmp", "-kmer-size", "-abundance-min", "-abundance-max", "-kmer-shannon-index", "-max-reads",
                           "-min-read-size", "-min-shannon-index", "-nb-cores", "-max-memory", "-max-count", "-max-merge", "-count-cmd",
                           "-merge-cmd", "-count-file", "-merge-file", "-verbose", "-minimizer-size", "-minimizer-type", "-repartition-type",
                           "-solidity-kind", "-max-disk", "-storage-type", "-histo-max", "-solidity-custom", "-nb-gpus", "-device", nullptr};
const char *kNoParam[] = {"-keep-tmp", "-data-info", "-simple-dist", "-complex-dist", "-help", "-version", nullptr};

bool in_list(const char *const *l, const std::string &s) { for (; *l; l++) if (s == *l) return true; return false; }

void usage() {
    printf("\n[Simka options] (B200 build: counting and merging run in-process on the GPU)\n"
           "       -in                 (1 arg) :    input file of samples. One sample per line: id1: filename1...\n"
           "       -out                (1 arg) :    output directory for result files (distance matrices)  [default './simka_results']\n"
           "       -out-tmp            (1 arg) :    output directory for temporary files\n"
           "       -keep-tmp           (0 arg) :    keep temporary files\n"
           "       -data-info          (0 arg) :    compute (and display) information before running Simka, such as the number of reads per dataset\n"
           "   [distance options]\n"
           "       -simple-dist        (0 arg) :    compute all simple distances (Chord, Hellinger...)\n"
           "       -complex-dist       (0 arg) :    compute all complex distances (Jensen-Shannon...)\n"
           "   [kmer options]\n"
           "       -kmer-size          (1 arg) :    size of a kmer  [default '21']\n"
           "       -abundance-min      (1 arg) :    min abundance a kmer need to be considered  [default '2']\n"
           "       -abundance-max      (1 arg) :    max abundance a kmer can have to be considered  [default '999999999']\n"
           "       -kmer-shannon-index (1 arg) :    minimal Shannon index a kmer should have to be kept (no effect, as in the reference)  [default '0']\n"
           "   [read options]\n"
           "       -max-reads          (1 arg) :    maximum number of reads per sample to process. Can be -1: use all reads. Can be 0: estimate it  [default '-1']\n"
           "       -min-read-size      (1 arg) :    minimal size a read should have to be kept  [default '0']\n"
           "       -min-shannon-index  (1 arg) :    minimal Shannon index a read should have to be kept. Float in [0,2]  [default '0']\n"
           "   [core options]\n"
           "       -nb-cores           (1 arg) :    number of host threads for parsing / packing  [default '0' = all]\n"
           "       -max-memory         (1 arg) :    max memory (MB); accepted for compatibility  [default '5000']\n"
           "       -device             (1 arg) :    CUDA device ordinal (the first one with -nb-gpus)  [default '0']\n"
           "       -nb-gpus            (1 arg) :    GPUs to use: sample i is counted on GPU (i mod nb-gpus), every GPU merges its block of hash ranges,\n"
           "                                        lists and accumulators cross NVLink through NCCL  [default '1']\n"
           "   [cluster options] -max-count -max-merge -count-cmd -merge-cmd -count-file -merge-file : accepted and ignored\n"
           "       -verbose            (1 arg) :    verbosity level  [default '1']\n\n");
}

Options parse_args(int argc, char **argv) {
    Options o;
    std::map<std::string, std::string> kv;
    // Command-line mistakes are handled the way the reference's GATB `Tool` does (GC/tools/misc/impl/Tool.cpp, OptionFailure): every
    // problem is listed as "ERROR: ...", the option list follows, and the process ends with EXIT_SUCCESS -- only failures of the run
    // itself ("EXCEPTION: ...") give EXIT_FAILURE.
    std::vector<std::string> errs;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (in_list(kNoParam, a)) { kv[a] = "1"; continue; }
        if (in_list(kOneParam, a)) {
            if (i + 1 >= argc) { errs.push_back("Too few arguments for the " + a + " option..."); continue; }
            kv[a] = argv[++i];
            continue;
        }
        errs.push_back("Unknown parameter '" + a + "'");
    }
    if (kv.count("-help")) { usage(); exit(EXIT_SUCCESS); }
    if (kv.count("-version")) { printf("%s\n", simka_b200_version()); exit(EXIT_SUCCESS); }
    if (!kv.count("-in")) errs.push_back("Option '-in' is mandatory");
    if (!kv.count("-out-tmp")) errs.push_back("Option '-out-tmp' is mandatory");
    if (!errs.empty()) {
        for (const std::string &e : errs) printf("ERROR: %s\n", e.c_str());
        printf("\n");
        usage();
        exit(EXIT_SUCCESS);
    }
    o.in = kv["-in"];
    o.outTmp = kv["-out-tmp"];
    if (kv.count("-out")) o.out = kv["-out"];
    o.keepTmp = kv.count("-keep-tmp");
    o.dataInfo = kv.count("-data-info");
    o.simple = kv.count("-simple-dist");
    o.complex_ = kv.count("-complex-dist");
    if (kv.count("-kmer-size")) o.kmerSize = atoi(kv["-kmer-size"].c_str());
    if (kv.count("-abundance-min")) o.abundanceMin = atoll(kv["-abundance-min"].c_str());
    if (kv.count("-abundance-max")) o.abundanceMax = atoll(kv["-abundance-max"].c_str());
    if (kv.count("-max-reads")) o.maxReads = atoll(kv["-max-reads"].c_str());
    if (kv.count("-min-read-size")) o.minReadSize = (size_t)atoll(kv["-min-read-size"].c_str());
    if (kv.count("-min-shannon-index")) o.minShannon = atof(kv["-min-shannon-index"].c_str());
    if (kv.count("-nb-cores")) o.nbCores = atoi(kv["-nb-cores"].c_str());
    if (kv.count("-max-memory")) o.maxMemory = atoll(kv["-max-memory"].c_str());
    if (kv.count("-verbose")) o.verbose = atoi(kv["-verbose"].c_str());
    if (kv.count("-device")) o.device = atoi(kv["-device"].c_str());
    if (kv.count("-nb-gpus")) o.nbGpus = std::max(1, atoi(kv["-nb-gpus"].c_str()));
    // src/SimkaPotara.hpp:383-387
    if (o.maxMemory < 500) throw bank::Exception("Please allocate more memory (minimum is 500 MB) with the option -max-memory");
    o.minShannon = std::max(o.minShannon, 0.0);
    o.minShannon = std::min(o.minShannon, 2.0);  // src/core/SimkaAlgorithm.cpp:202-204
    if (o.kmerSize < 2 || o.kmerSize > 63)
        throw bank::Exception("this build supports -kmer-size 2..63 (spans 32 and 64); got " + std::to_string(o.kmerSize));
    if (o.kmerSize > 31 && o.nbGpus > 1) throw bank::Exception("-nb-gpus needs -kmer-size <= 31");
    return o;
}

void mkdir_p(const std::string &d) {
    std::string cur;
    for (size_t i = 0; i <= d.size(); i++) {
        if (i == d.size() || d[i] == '/') {
            if (!cur.empty() && cur != "." && cur != "..") {
                if (mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) throw bank::Exception("Cannot create dir " + cur);
            }
        }
        if (i < d.size()) cur.push_back(d[i]);
    }
}
void rm_rf(const std::string &d) {
    if (d.empty() || d == "/") return;
    std::string cmd = "rm -rf '" + d + "'";
    if (system(cmd.c_str()) != 0) {}
}
void write_text(const std::string &path, const std::string &text) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw bank::Exception("cannot write " + path);
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
}

// A chunk of packed reads on the host.  Pinned (cudaHostAlloc) once the GPU context exists; chunks filled BEFORE that -- the
// parser threads start at once while the CUDA context takes 1-2 s to come up, and a pinned allocation would wait for it -- are
// plain page-aligned memory, which simka_gpu_push_reads copies through the driver's staging buffers.
struct PinnedChunk {
    uint64_t *packed = nullptr;
    uint32_t *bad = nullptr;
    uint64_t cap_words = 0;
    bool pinned = false;
    void reserve(uint64_t words, bool may_pin) {
        if (words + 1 <= cap_words) return;
        release();
        cap_words = words + words / 8 + 1024;
        void *p = nullptr, *b = nullptr;
        if (may_pin) {
            if (simka_gpu_host_alloc(&p, cap_words * 8) != SIMKA_OK || simka_gpu_host_alloc(&b, cap_words * 4) != SIMKA_OK)
                throw bank::Exception("cannot allocate pinned host memory");
        } else {
            p = aligned_alloc(4096, (cap_words * 8 + 4095) & ~(size_t)4095);
            b = aligned_alloc(4096, (cap_words * 4 + 4095) & ~(size_t)4095);
            if (!p || !b) { free(p); free(b); throw bank::Exception("cannot allocate host memory"); }
        }
        pinned = may_pin;
        packed = (uint64_t *)p;
        bad = (uint32_t *)b;
    }
    void release() {
        if (pinned) {
            if (packed) simka_gpu_host_free(packed);
            if (bad) simka_gpu_host_free(bad);
        } else {
            free(packed);
            free(bad);
        }
        packed = nullptr; bad = nullptr; cap_words = 0;
    }
};

// bytes of one sample's input files (gzip files weighted by a typical 4x expansion)
uint64_t sample_input_bytes(const SampleSpec &s) {
    uint64_t total = 0;
    for (auto &g : s.groups)
        for (auto &f : g) {
            struct stat st;
            if (stat(f.c_str(), &st) != 0) continue;
            const bool gz = f.size() > 3 && f.compare(f.size() - 3, 3, ".gz") == 0;
            total += (uint64_t)st.st_size * (gz ? 4 : 1);
        }
    return total;
}

void check(simka_ctx *ctx, int rc, const char *what) {
    if (rc != SIMKA_OK) throw bank::Exception(std::string(what) + ": " + (ctx ? simka_gpu_last_error(ctx) : "failed") + " (code " + std::to_string(rc) + ")");
}

int run(int argc, char **argv) {
    Options o = parse_args(argc, argv);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<SampleSpec> samples = parse_simka_input(o.in);
    if (samples.empty()) throw bank::Exception("no sample in " + o.in);
    const size_t N = samples.size();
    const int nthreads = o.nbCores > 0 ? o.nbCores : (int)std::thread::hardware_concurrency();

    // directory layout of the reference (src/core/SimkaAlgorithm.cpp:214-239, Appendix A of SURVEY.md)
    const std::string tmp = o.outTmp + "/simka_output_temp";
    mkdir_p(o.out);
    mkdir_p(tmp + "/input");
    mkdir_p(tmp + "/count_synchro");
    {
        std::string ids, list;
        for (const SampleSpec &s : samples) {
            ids += s.id + "\n";
            std::string album;
            for (auto &g : s.groups) for (auto &f : g) album += f + "\n";
            if (!album.empty()) album.pop_back();
            write_text(tmp + "/input/" + s.id, album);
            list += tmp + "/input//" + s.id + "\n";
        }
        if (!list.empty()) list.pop_back();
        write_text(tmp + "/input/__input_simka__", list);
        write_text(tmp + "/datasetIds", ids);
    }
    SimkaSequenceFilter filter(o.minReadSize, o.minShannon);

    // -max-reads: -1 = all, N = at most N per group, 0 = (min + mean) / 2 over the samples
    // (src/core/SimkaAlgorithm.cpp:377-445).  The reference takes the read counts from IBank::estimateNbItems -- the head of every
    // file extrapolated to the file size, a gz file counted as four times its compressed size -- and so does this: the limit
    // must be the same number for the same files (bank::BankFasta::estimate)
    uint64_t maxReads = 0;
    if (o.maxReads > 0) maxReads = (uint64_t)o.maxReads;
    if (o.maxReads == 0 || o.dataInfo) {
        uint64_t mn = ~0ull, sum = 0;
        for (const SampleSpec &s : samples) {
            uint64_t n = 0;
            for (auto &g : s.groups)
                for (auto &f : g) { uint64_t ni, t, m; bank::BankFasta(f).estimate(ni, t, m); n += ni; }
            if (o.dataInfo || o.verbose) printf("\t%s: about %llu reads\n", s.id.c_str(), (unsigned long long)n);
            n /= std::max<size_t>(1, s.groups.size());  // the limit applies per paired group (src/core/SimkaAlgorithm.cpp:405, nbReads /= _nbBankPerDataset[i])
            mn = std::min(mn, n);
            sum += n;
        }
        if (o.maxReads == 0) {
            maxReads = (mn + sum / N) / 2;
            if (o.verbose) printf("Reads per sample used up to: %llu\n\n", (unsigned long long)maxReads);
        }
    }

    // stage 1.  Host ingest pipeline: worker threads parse + 2-bit pack whole samples into pinned chunks (one sample per
    // worker at a time: files are independent, so the parse scales with the cores), the main thread feeds the GPU in
    // sample order: push (async H2D) of sample i+1 overlaps the counting of sample i.  Pinned chunks come from a pool and
    // go back to it when their sample is finished; at most `inflight` samples are parsed ahead.
    const uint64_t CHUNK_BASES = 1ull << 29;
    struct Parsed {
        std::vector<PinnedChunk *> chunks;
        std::vector<uint64_t> words, reads;
        uint64_t nreads = 0;
        bool ready = false;
        std::string error;
    };
    std::vector<Parsed> parsed(N);
    std::vector<PinnedChunk *> pool;
    std::mutex mu;
    std::condition_variable cv;
    size_t next_sample = 0, consumed = 0;
    const size_t workers = std::max<size_t>(1, std::min<size_t>({(size_t)nthreads, N, (size_t)16}));
    // samples parsed ahead of the GPU: as many as 80 % of -max-memory holds in packed form (0.375 B per base, about 0.4 B per
    // input byte), at least workers + 2 -- with the default 5000 MB all 32 samples of a 5 GB run are packed while CUDA starts
    size_t inflight = workers + 2;
    {
        uint64_t in_bytes = 0;
        for (const SampleSpec &sp : samples) in_bytes += sample_input_bytes(sp);
        const double per_sample = 0.4 * (double)in_bytes / (double)N + 1.0;
        const double fit = 0.8 * (double)o.maxMemory * 1048576.0 / per_sample;
        inflight = std::max(inflight, (size_t)std::min<double>((double)N, fit));
    }
    std::atomic<bool> gpu_ready{false};
    const bool pageable_only = getenv("SIMKA_B200_PAGEABLE_INPUT") != nullptr;  // tests: every chunk takes the before-the-context path
    double t_packed = 0;  // when the last sample so far came out of the parser threads (under mu)
    auto get_chunk = [&](uint64_t words) {
        PinnedChunk *c = nullptr;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t q = 0; q < pool.size(); q++)
                if (pool[q]->cap_words >= words + 1) { c = pool[q]; pool.erase(pool.begin() + q); break; }
            if (!c && !pool.empty()) { c = pool.back(); pool.pop_back(); }
        }
        if (!c) c = new PinnedChunk();
        const bool may_pin = gpu_ready.load() && !pageable_only;
        if (may_pin && !c->pinned) c->release();  // a recycled before-the-context chunk: replaced by a pinned one, once
        c->reserve(words, may_pin);
        return c;
    };
    auto worker = [&]() {
        std::string seqs;
        std::vector<uint64_t> off;
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next_sample >= N || next_sample < consumed + inflight; });
                if (next_sample >= N) return;
                i = next_sample++;
            }
            Parsed &P = parsed[i];
            try {
                seqs.clear();
                off.assign(1, 0);
                auto flush = [&]() {
                    if (off.size() <= 1) return;
                    const uint64_t nr = off.size() - 1;
                    const uint64_t words = simka_pack_words(seqs.size(), nr);
                    PinnedChunk *c = get_chunk(words);
                    simka_pack_reads_k(seqs.data(), off.data(), nr, c->packed, c->bad, 1, o.kmerSize);
                    P.chunks.push_back(c);
                    P.words.push_back(words);
                    P.reads.push_back(nr);
                    seqs.clear();
                    off.assign(1, 0);
                };
                P.nreads = read_sample_into(samples[i], filter, maxReads, seqs, off, [&]() { if (seqs.size() >= CHUNK_BASES) flush(); }, o.kmerSize);
                flush();
            } catch (const std::exception &e) { P.error = e.what(); if (P.error.empty()) P.error = "read error"; }
            {
                std::lock_guard<std::mutex> lk(mu);
                P.ready = true;
                t_packed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool_threads;
    for (size_t w = 0; w < workers; w++) pool_threads.emplace_back(worker);
    // the GPU context is created while the parser threads already work on the first samples
    simka_params prm;
    memset(&prm, 0, sizeof prm);
    prm.kmer_size = o.kmerSize;
    prm.nb_samples = (int32_t)N;
    prm.abundance_min = o.abundanceMin;
    prm.abundance_max = o.abundanceMax;
    prm.simple_dist = o.simple;
    prm.complex_dist = o.complex_;
    prm.device = o.device;
    // one context per GPU; with -nb-gpus G > 1 sample i belongs to GPU i % G, one host thread drives each GPU
    const size_t G = std::min<size_t>((size_t)o.nbGpus, N);
    if (G > 1) {
        // every context must cut the hash space into the same fine ranges: the geometry comes from the expected size of a sample,
        // taken here from the first parsed one (a single context derives it from the first sample it counts)
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return parsed[0].ready; });
        uint64_t w = 0;
        for (uint64_t x : parsed[0].words) w += x;
        prm.expected_kmers_per_sample = std::max<uint64_t>(1, w * 32);
    }
    std::vector<simka_ctx *> ctxs(G, nullptr);
    int rc = SIMKA_OK;
    for (size_t g = 0; g < G && rc == SIMKA_OK; g++) {
        simka_params pg = prm;
        pg.device = o.device + (int)g;
        rc = simka_gpu_create(&ctxs[g], &pg);
    }
    simka_ctx *ctx = ctxs[0];
    const double t_create = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != SIMKA_OK) { { std::lock_guard<std::mutex> lk(mu); next_sample = N; } cv.notify_all(); for (auto &t : pool_threads) t.join(); for (simka_ctx *c : ctxs) simka_gpu_destroy(c); throw bank::Exception("simka_gpu_create failed (code " + std::to_string(rc) + "): " + std::to_string(G) + " CUDA device(s) are required, there is no CPU path"); }
    if (G > 1) {
        uint8_t id[SIMKA_COMM_ID_BYTES];
        if (simka_comm_unique_id(id) != SIMKA_OK) throw bank::Exception("-nb-gpus: NCCL (libnccl.so.2) is not available");
        std::vector<std::thread> th;
        std::vector<int> rcs(G, SIMKA_OK);
        for (size_t g = 0; g < G; g++) th.emplace_back([&, g] { rcs[g] = simka_gpu_comm_init(ctxs[g], id, (int)g, (int)G); });  // (collective: all ranks at once)
        for (auto &t : th) t.join();
        for (size_t g = 0; g < G; g++) if (rcs[g] != SIMKA_OK) throw bank::Exception(std::string("simka_gpu_comm_init: ") + simka_gpu_last_error(ctxs[g]));
    }

    gpu_ready.store(true);

    std::vector<simka_sample_totals> totals(N);
    std::string failure;
    std::mutex fail_mu;
    auto failed = [&] { std::lock_guard<std::mutex> lk(fail_mu); return !failure.empty(); };
    auto fail = [&](const std::string &m) { std::lock_guard<std::mutex> lk(fail_mu); if (failure.empty()) failure = m; };
    size_t consumed_count = 0;  // samples whose chunks went back to the pool (under mu)
    auto finish = [&](simka_ctx *cx, size_t i) {  // counts sample i (already pushed), recycles its pinned chunks, writes its .ok file
        Parsed &P = parsed[i];
        if (!failed()) {
            const int rc2 = simka_gpu_finish_sample(cx, (int)i, &totals[i]);
            if (rc2 != SIMKA_OK) fail(std::string("simka_gpu_finish_sample: ") + simka_gpu_last_error(cx) + " (code " + std::to_string(rc2) + ")");
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            for (PinnedChunk *c : P.chunks) pool.push_back(c);
            P.chunks.clear();
            consumed = ++consumed_count;
        }
        cv.notify_all();
        if (failed()) return;
        totals[i].nb_reads = P.nreads;
        // count_synchro/<id>.ok: nbReads, nbDistinct, sum c, sum c^2 (src/SimkaCount.cpp:355-368)
        char buf[256];
        snprintf(buf, sizeof buf, "%llu\n%llu\n%llu\n%llu\n", (unsigned long long)totals[i].nb_reads, (unsigned long long)totals[i].nb_distinct,
                 (unsigned long long)totals[i].nb_kmers, (unsigned long long)totals[i].chord_n2);
        write_text(tmp + "/count_synchro/" + samples[i].id + ".ok", buf);
        if (o.verbose) printf("Counting dataset %zu (%s): %llu reads, %llu distinct solid k-mers, %llu solid k-mers\n", i, samples[i].id.c_str(),
                              (unsigned long long)totals[i].nb_reads, (unsigned long long)totals[i].nb_distinct, (unsigned long long)totals[i].nb_kmers);
    };
    auto drive = [&](size_t g) {  // the samples of GPU g, in order
        simka_ctx *cx = ctxs[g];
        size_t prev = N;
        for (size_t i = g; i < N; i += G) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return parsed[i].ready; });
            }
            Parsed &P = parsed[i];
            if (!P.error.empty()) fail(P.error);
            // push sample i (asynchronous copies) BEFORE finishing the previous one: its H2D overlaps those kernels
            for (size_t c = 0; c < P.chunks.size() && !failed(); c++) {
                const int rc2 = simka_gpu_push_reads(cx, (int)i, P.chunks[c]->packed, P.chunks[c]->bad, P.words[c], P.reads[c]);
                if (rc2 != SIMKA_OK) fail(std::string("simka_gpu_push_reads: ") + simka_gpu_last_error(cx) + " (code " + std::to_string(rc2) + ")");
            }
            if (prev < N) finish(cx, prev);
            prev = i;
        }
        if (prev < N) finish(cx, prev);
    };
    if (G == 1) drive(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(drive, g);
        for (auto &t : th) t.join();
    }
    for (auto &t : pool_threads) t.join();
    if (!failure.empty()) { for (simka_ctx *c : ctxs) simka_gpu_destroy(c); throw bank::Exception(failure); }
    const double t_stage1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // stage 2 + final statistics
    if (G == 1) {
        check(ctx, simka_gpu_set_totals(ctx, totals.data()), "simka_gpu_set_totals");
        check(ctx, simka_gpu_merge_and_accumulate(ctx), "simka_gpu_merge_and_accumulate");
    } else {
        // totals all-reduce, list exchange, merge of every GPU's block of hash ranges, accumulator all-reduce (NCCL, one thread per GPU)
        std::vector<std::thread> th;
        std::vector<int> rcs(G, SIMKA_OK);
        for (size_t g = 0; g < G; g++) th.emplace_back([&, g] { rcs[g] = simka_gpu_exchange_and_merge(ctxs[g]); });
        for (auto &t : th) t.join();
        for (size_t g = 0; g < G; g++) check(ctxs[g], rcs[g], "simka_gpu_exchange_and_merge");
    }
    const size_t nn = N * N, ns = N * (N + 1) / 2;
    std::vector<uint64_t> u(2 * ns + 6 * nn);
    std::vector<double> kl(nn);
    std::vector<simka_sample_totals> gt(N);
    simka_stats_view v;
    memset(&v, 0, sizeof v);
    v.totals = gt.data();
    v.distinct_shared = &u[0];
    v.braycurtis_num = &u[ns];
    v.nb_shared = &u[2 * ns];
    v.chord_ninj = &u[2 * ns + nn];
    v.hellinger = &u[2 * ns + 2 * nn];
    v.kulczynski = &u[2 * ns + 3 * nn];
    v.canberra = &u[2 * ns + 4 * nn];
    v.whittaker = &u[2 * ns + 5 * nn];
    v.kullback_leibler = kl.data();
    check(ctx, simka_gpu_get_stats(ctx, &v), "simka_gpu_get_stats");
    std::vector<const char *> ids;
    for (const SampleSpec &s : samples) ids.push_back(s.id.c_str());
    rc = simka_write_matrices(&v, ids.data(), o.out.c_str());
    if (rc != SIMKA_OK) throw bank::Exception("cannot write the matrices to " + o.out);
    if (o.verbose) {
        simka_timings tm;
        simka_gpu_get_timings(ctx, &tm);
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("\nNb distinct kmers: %llu   shared by >= 2 samples: %llu\n", (unsigned long long)v.nb_distinct_kmers, (unsigned long long)v.nb_shared_kmers);
        printf("GPU time: %.1f ms counting + %.1f ms merging, %llu kernel launches; wall %.2f s (GPU context ready at %.2f s, all samples parsed + packed at %.2f s, read + count done at %.2f s, %zu parser threads, %zu samples ahead)\n", tm.samples_ms,
               tm.merge_ms, (unsigned long long)tm.kernel_launches, wall, t_create, t_packed, t_stage1, workers, inflight);
        printf("Output dir: %s\n", o.out.c_str());
    }
    for (simka_ctx *c : ctxs) simka_gpu_destroy(c);
    for (PinnedChunk *c : pool) { c->release(); delete c; }
    if (!o.keepTmp) rm_rf(tmp);
    return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char **argv) {
    try {
        return run(argc, argv);
    } catch (const std::exception &e) {
        std::cout << "EXCEPTION: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
}
