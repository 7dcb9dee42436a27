// simka_input.hpp -- the input side of `simka`: input-file grammar, read filter, per-sample read selection.
//
// Restates (does not port) src/core/SimkaAlgorithm.cpp:245-351 (layoutInputFilename: "ID: f1 , f2 ; g1 , g2"),
// src/core/SimkaCommons.hpp:317-436 (SimkaSequenceFilter: minimal read size, Shannon index) and :150-314
// (SimkaInputIterator: the files of a sample are read group by group -- a group is one ';' part -- and at most
// `maxReads` filtered reads are taken from each group).
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "bank.hpp"

namespace simka_b200 {

struct SampleSpec {
    std::string id;
    std::vector<std::vector<std::string>> groups;  // ';' separated groups of ',' separated files (absolute paths)
    size_t nbFiles() const { size_t n = 0; for (auto &g : groups) n += g.size(); return n; }
};

inline std::string dirname_of(const std::string &p) {
    size_t s = p.rfind('/');
    return s == std::string::npos ? std::string(".") : (s == 0 ? std::string("/") : p.substr(0, s));
}

inline std::string real_path(const std::string &p) {
    char buf[4096];
    if (realpath(p.c_str(), buf)) return std::string(buf);
    return p;
}

// src/core/SimkaAlgorithm.cpp:267-335: spaces removed, "id:groups", groups split on ';', files on ',';
// relative file names are resolved against the directory of the input file
inline std::vector<SampleSpec> parse_simka_input(const std::string &filename) {
    std::ifstream in(filename.c_str());
    if (!in) throw bank::Exception("Input filename does not exist  (" + filename + ")");
    const std::string dir = dirname_of(real_path(filename));
    std::vector<SampleSpec> out;
    std::string line;
    while (std::getline(in, line)) {
        std::string l;
        for (char c : line) if (c != ' ' && c != '\r' && c != '\t') l.push_back(c);
        if (l.empty()) continue;
        size_t colon = l.find(':');
        if (colon == std::string::npos) throw bank::Exception("bad line in " + filename + ": " + line);
        SampleSpec s;
        s.id = l.substr(0, colon);
        std::string rest = l.substr(colon + 1);
        size_t c2 = rest.find(':');  // the reference splits on every ':' and keeps part 1
        if (c2 != std::string::npos) rest = rest.substr(0, c2);
        std::stringstream gs(rest);
        std::string grp;
        while (std::getline(gs, grp, ';')) {
            std::vector<std::string> files;
            std::stringstream fs(grp);
            std::string f;
            while (std::getline(fs, f, ',')) {
                if (f.empty()) continue;
                files.push_back(f[0] == '/' ? f : dir + "/" + f);
            }
            s.groups.push_back(files);
        }
        out.push_back(s);
    }
    return out;
}

// src/core/SimkaCommons.hpp:317-436
struct SimkaSequenceFilter {
    SimkaSequenceFilter(size_t minReadSize = 0, double minShannonIndex = 0) : _minReadSize(minReadSize), _minShannonIndex(minShannonIndex) {}
    bool operator()(const char *seq, size_t len) const {
        if (_minReadSize != 0 && len < _minReadSize) return false;
        if (_minShannonIndex != 0 && getShannonIndex(seq, len) < _minShannonIndex) return false;
        return true;
    }
    bool operator()(bank::Sequence &s) const { return (*this)(s.getDataBuffer(), s.getDataSize()); }
    // frequencies of A, C, T, G, N (every other character, lower case included, counts as 'A': the reference's
    // nt2binTab is indexed by the raw character), float accumulation, base-2 logarithm, absolute value
    static float getShannonIndex(const char *seq, size_t len) {
        float freqs[5] = {0, 0, 0, 0, 0};
        for (size_t i = 0; i < len; i++) {
            int b = 0;
            switch ((unsigned char)seq[i]) { case 'C': b = 1; break; case 'T': b = 2; break; case 'G': b = 3; break; case 'N': b = 4; break; default: b = 0; }
            freqs[b] += 1.0f;
        }
        float index = 0;
        for (int i = 0; i < 5; i++) {
            freqs[i] /= (float)len;
            if (freqs[i] != 0) index += freqs[i] * std::log(freqs[i]) / std::log(2);  // log(float) / log(int -> double), as the reference's overloads resolve
        }
        return std::fabs(index);
    }
    size_t _minReadSize;
    double _minShannonIndex;
};

// Read selection of SimkaInputIterator (src/core/SimkaCommons.hpp:150-314), restated: the files of a sample are read group
// by group (a group = one ';' part of the input line).  Inside a group a counter is incremented for every filtered read
// EXCEPT the first read of each file (first() does not count); the read that makes the counter reach maxReads is dropped
// and the group ends.  So a group yields maxReads reads, plus one for every file switch that happens before the limit
// (maxReads == 0: everything).
struct GroupReadCounter {
    uint64_t maxReads, count = 0;
    bool first_of_file = true;
    explicit GroupReadCounter(uint64_t m) : maxReads(m) {}
    void new_file() { first_of_file = true; }
    // call for every filtered read; false = drop it and stop the group
    bool take() {
        if (first_of_file) { first_of_file = false; return true; }
        count++;
        return !(maxReads && count >= maxReads);
    }
};

// Reads one sample; `sink(ptr, len)` receives each selected read.  Returns the number of selected reads.
template <class Sink>
inline uint64_t read_sample(const SampleSpec &spec, const SimkaSequenceFilter &filter, uint64_t maxReads, Sink sink) {
    uint64_t total = 0;
    std::string data;
    for (const auto &group : spec.groups) {
        GroupReadCounter ctr(maxReads);
        bool stop = false;
        for (const std::string &file : group) {
            if (stop) break;
            bank::FastxParser parser(file);
            ctr.new_file();
            for (;;) {
                data.clear();
                if (!parser.nextRead(data)) break;
                if (!filter(data.data(), data.size())) continue;
                if (!ctr.take()) { stop = true; break; }
                sink(data.data(), data.size());
                total++;
            }
        }
    }
    return total;
}

// Batch variant for the ingest pipeline: reads are parsed straight into the concatenated buffer `seqs` (offsets in
// `off`, off[0] = 0 on entry), a rejected read is rolled back; `after_read()` is called after every accepted read (the
// caller flushes the batch there when it is large enough).
// MiniKC quirk (k <= 15 only; pinned on reference runs, tests/golden/truth_ref/example_n_results_k{12,15}_*): MiniKC::count
// (src/minikc/MiniKC.hpp:204-231) refills its k-mer iterator with Model::Iterator::setData -> Model::build
// (GC/kmer/impl/Model.hpp:603-610,499-513), which returns early without touching the k-mer vector when the read is shorter than
// k -- so the k-mers of the PREVIOUS long-enough read of the sample are counted once more.  A short read holds no k-mer of its
// own, hence replacing it by a copy of that previous read gives exactly the reference's counts.
struct MiniKCReplay {
    size_t k = 0;  // 0: off (k > 15)
    std::string last;
    // any_k: SimkaMin's SelectKmersCommand has the same stale vector for every k (src/simkaMin/SimkaMinCount.hpp:217-263)
    explicit MiniKCReplay(int kmer_size = 0, bool any_k = false) : k(kmer_size > 0 && (kmer_size <= 15 || any_k) ? (size_t)kmer_size : 0) {}
    // the read just appended to `seqs` at [start, seqs.size())
    void apply(std::string &seqs, size_t start) {
        if (!k) return;
        if (seqs.size() - start < k) { seqs.resize(start); seqs.append(last); }
        else last.assign(seqs, start, std::string::npos);
    }
};

template <class AfterRead>
inline uint64_t read_sample_into(const SampleSpec &spec, const SimkaSequenceFilter &filter, uint64_t maxReads, std::string &seqs,
                                 std::vector<uint64_t> &off, AfterRead after_read, int kmer_size = 0, bool replay_any_k = false) {
    uint64_t total = 0;
    MiniKCReplay replay(kmer_size, replay_any_k);
    {   // room for the whole sample up front (bases < file bytes; gzip files weighted 4x; at most 1 GiB, the caller flushes in
        // batches anyway): growing a 150 MB string by doubling re-copies it and touches twice the pages -- 0.75 -> 1.35 GB/s
        // of FASTA per thread
        uint64_t bytes = 0;
        for (const auto &group : spec.groups)
            for (const std::string &file : group) {
                struct stat st;
                if (::stat(file.c_str(), &st) != 0) continue;
                const bool gz = file.size() > 3 && file.compare(file.size() - 3, 3, ".gz") == 0;
                bytes += (uint64_t)st.st_size * (gz ? 4 : 1);
            }
        bytes = std::min<uint64_t>(bytes, (uint64_t)1 << 30);
        if (maxReads == 0 && seqs.capacity() < seqs.size() + bytes) {
            seqs.reserve(seqs.size() + bytes);
            off.reserve(off.size() + bytes / 128);
        }
    }
    for (const auto &group : spec.groups) {
        GroupReadCounter ctr(maxReads);
        bool stop = false;
        for (const std::string &file : group) {
            if (stop) break;
            bank::FastxParser parser(file);
            ctr.new_file();
            for (;;) {
                const size_t start = seqs.size();
                if (!parser.nextRead(seqs)) break;
                if (!filter(seqs.data() + start, seqs.size() - start)) { seqs.resize(start); continue; }
                if (!ctr.take()) { seqs.resize(start); stop = true; break; }
                replay.apply(seqs, start);
                off.push_back(seqs.size());
                after_read();
                total++;
            }
        }
    }
    return total;
}

}  // namespace simka_b200
