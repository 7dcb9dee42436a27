// stage_files.hpp -- the temp-dir wire formats of a reference run, read and written by
// the simkaCount / simkaMerge executables of this build so that they interoperate with the reference's
// orchestrator and stage binaries:
//   solid/part_<p>/__p__<i>.gz      gzip-1 stream of 24-byte {u64 kmer; u32 bank; u32 pad; u64 count}, ascending k-mer
//                                   (Kmer_BankId_Count, src/minikc/MiniKC.hpp:24-38; BagGzFile "wb1", GC/tools/collections/impl/BagFile.hpp:116-123)
//   count_synchro/<id>.ok           4 decimal lines: nbReads, nbDistinctSolid, sum c, sum c^2   (src/SimkaCount.cpp:355-368)
//   kmercount_per_partition/<id>.txt  P lines: distinct solid k-mers per partition              (src/SimkaCount.cpp:338-345)
//   datasetIds                      one sample id per line                                      (src/SimkaPotara.hpp:441-458)
//   stats/part_<p>.gz               gzip stream of x87 long double in the order of SimkaStatistics::save (src/core/SimkaDistance.cpp:455-513)
//   merge_synchro/<p>.ok            empty                                                       (src/SimkaMerge.cpp:1472-1476)
#pragma once
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../../include/simka_b200.h"
#include "bank.hpp"

namespace simka_b200 {
namespace stage {

struct Record {  // 24 bytes, as the reference struct is laid out for span 32
    uint64_t kmer;
    uint32_t bank;
    uint32_t pad;
    uint64_t count;
};
static_assert(sizeof(Record) == 24, "Kmer_BankId_Count is 24 bytes");
static_assert(sizeof(long double) == 16, "the stats stream is made of 16-byte x87 long doubles");

inline void mkdirs(const std::string &d) {
    std::string cur;
    for (size_t i = 0; i <= d.size(); i++) {
        if (i == d.size() || d[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); }
        if (i < d.size()) cur.push_back(d[i]);
    }
}

inline std::vector<std::string> listdir(const std::string &d) {  // plain entry names, "." and ".." left out
    std::vector<std::string> out;
    if (DIR *dp = opendir(d.c_str())) {
        while (struct dirent *e = readdir(dp)) {
            const std::string n = e->d_name;
            if (n != "." && n != "..") out.push_back(n);
        }
        closedir(dp);
    }
    return out;
}

inline std::string part_file(const std::string &dir, size_t p, size_t bank) {
    return dir + "/solid/part_" + std::to_string(p) + "/__p__" + std::to_string(bank) + ".gz";
}

// gzwrite takes an `unsigned` length and reports errors by returning 0: write in chunks of at most 1 GiB, check every call
// (a truncated stats/part_P.gz followed by its merge_synchro token would be read as a complete result)
inline void gz_write_all(gzFile f, const void *data, size_t bytes, const std::string &path) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    while (bytes) {
        const unsigned chunk = (unsigned)std::min<size_t>(bytes, (size_t)1 << 30);
        if (gzwrite(f, p, chunk) != (int)chunk) { gzclose(f); throw bank::Exception("write error on " + path); }
        p += chunk;
        bytes -= chunk;
    }
}

inline void write_solid_part(const std::string &path, uint32_t bank, const uint64_t *kmers, const uint32_t *counts, size_t n) {
    gzFile f = gzopen(path.c_str(), "wb1");
    if (!f) throw bank::Exception("cannot write " + path);
    std::vector<Record> buf;
    buf.reserve(10000);
    for (size_t i = 0; i < n; i++) {
        Record r;
        r.kmer = kmers[i]; r.bank = bank; r.pad = 0; r.count = counts[i];
        buf.push_back(r);
        if (buf.size() == 10000 || i + 1 == n) { gz_write_all(f, buf.data(), buf.size() * sizeof(Record), path); buf.clear(); }
    }
    if (gzclose(f) != Z_OK) throw bank::Exception("error closing " + path);
}

inline void read_solid_part(const std::string &path, std::vector<uint64_t> &kmers, std::vector<uint32_t> &counts) {
    kmers.clear();
    counts.clear();
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return;  // a sample without solid k-mers in this partition may have no file
    std::vector<Record> buf(10000);
    for (;;) {
        const int got = gzread(f, buf.data(), (unsigned)(buf.size() * sizeof(Record)));
        if (got <= 0) break;
        const size_t nr = (size_t)got / sizeof(Record);
        for (size_t i = 0; i < nr; i++) {
            if (buf[i].count > 0xFFFFFFFFull) { gzclose(f); throw bank::Exception("count above 2^32 in " + path); }
            kmers.push_back(buf[i].kmer);
            counts.push_back((uint32_t)buf[i].count);
        }
    }
    gzclose(f);
}

inline void write_ok(const std::string &path, const simka_sample_totals &t) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) throw bank::Exception("cannot write " + path);
    fprintf(f, "%llu\n%llu\n%llu\n%llu\n", (unsigned long long)t.nb_reads, (unsigned long long)t.nb_distinct, (unsigned long long)t.nb_kmers,
            (unsigned long long)t.chord_n2);
    fclose(f);
}

inline simka_sample_totals read_ok(const std::string &path) {
    std::ifstream in(path.c_str());
    if (!in) throw bank::Exception("missing " + path);
    std::vector<unsigned long long> v;
    std::string line;
    while (std::getline(in, line)) if (!line.empty()) v.push_back(strtoull(line.c_str(), nullptr, 10));
    if (v.size() < 4) throw bank::Exception("bad " + path);
    simka_sample_totals t;
    t.nb_reads = v[0]; t.nb_distinct = v[1]; t.nb_kmers = v[2]; t.chord_n2 = v[3];
    return t;
}

inline std::vector<std::string> read_lines(const std::string &path) {
    std::ifstream in(path.c_str());
    if (!in) throw bank::Exception("missing " + path);
    std::vector<std::string> v;
    std::string line;
    while (std::getline(in, line)) if (!line.empty()) v.push_back(line);
    return v;
}

// SimkaStatistics::save (src/core/SimkaDistance.cpp:455-513)
inline void write_stats(const std::string &path, const simka_stats_view &v) {
    const size_t N = (size_t)v.nb_samples, ns = N * (N + 1) / 2;
    std::vector<long double> s;
    auto put = [&](long double x) { s.push_back(x); };
    put(v.simple_dist ? 1 : 0);
    put(v.complex_dist ? 1 : 0);
    put(0);                              // _nbKmers            (only printed)
    put(0);                              // _nbErroneousKmers   (only printed)
    put((long double)v.nb_distinct_kmers);
    put(0);                              // _nbSolidKmers       (only printed)
    put((long double)v.nb_shared_kmers);
    for (size_t i = 0; i < N; i++) put((long double)v.totals[i].nb_distinct);  // _nbSolidDistinctKmersPerBank (global, from the .ok files)
    for (size_t i = 0; i < N; i++) put(0);                                     // _nbKmersPerBank (only printed)
    for (size_t i = 0; i < N; i++) put((long double)v.totals[i].nb_kmers);     // _nbSolidKmersPerBank (global)
    for (size_t i = 0; i < N; i++) for (size_t j = 0; j < N; j++) put((long double)v.nb_shared[i * N + j]);
    for (size_t i = 0; i < ns; i++) { put((long double)v.distinct_shared[i]); put((long double)v.braycurtis_num[i]); }
    if (v.simple_dist) {
        for (size_t i = 0; i < N; i++) put((long double)sqrt((double)v.totals[i].chord_n2));  // _chord_sqrt_N2 (SimkaDistance.cpp:138)
        for (size_t i = 0; i < N; i++) {
            for (size_t j = 0; j < N; j++) put((long double)v.chord_ninj[i * N + j]);
            for (size_t j = 0; j < N; j++) put((long double)v.hellinger[i * N + j]);
            for (size_t j = 0; j < N; j++) put((long double)v.kulczynski[i * N + j]);
        }
    }
    if (v.complex_dist) {
        for (size_t i = 0; i < N; i++) {
            for (size_t j = 0; j < N; j++) put((long double)v.canberra[i * N + j]);
            for (size_t j = 0; j < N; j++) put((long double)v.whittaker[i * N + j]);
            for (size_t j = 0; j < N; j++) put((long double)v.kullback_leibler[i * N + j]);
        }
    }
    // the padding bytes of an x87 long double are unspecified: zero them so that the files are reproducible
    std::vector<unsigned char> raw(s.size() * 16, 0);
    for (size_t i = 0; i < s.size(); i++) memcpy(&raw[i * 16], &s[i], 10);
    gzFile f = gzopen(path.c_str(), "wb1");
    if (!f) throw bank::Exception("cannot write " + path);
    gz_write_all(f, raw.data(), raw.size(), path);
    if (gzclose(f) != Z_OK) throw bank::Exception("error closing " + path);
}

// "-flag value" / "-flag" command lines of the stage binaries
inline std::map<std::string, std::string> parse_flags(int argc, char **argv, const char *const *no_param) {
    std::map<std::string, std::string> kv;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        bool np = false;
        for (const char *const *q = no_param; *q; q++) if (a == *q) np = true;
        if (np || i + 1 >= argc) { kv[a] = "1"; continue; }
        kv[a] = argv[++i];
    }
    return kv;
}

// the partition of a canonical k-mer in the files this build writes (any function works for the merge as long as
// every sample uses the same one; the reference uses its minimizer repartition table)
inline size_t partition_of(uint64_t kmer, size_t nb_partitions) {
    uint64_t x = kmer * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    return (size_t)((x >> 32) % nb_partitions);
}

}  // namespace stage
}  // namespace simka_b200
