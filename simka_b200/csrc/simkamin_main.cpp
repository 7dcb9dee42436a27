// simkamin_main.cpp -- `simkaMinCore sketch | distance | export | append | info | matrix-update` with the reference's command
// lines and file formats (src/simkaMin/SimkaMin.cpp:30-110; sketch: SimkaMinCount.hpp:1347-1381,713-1127; distance:
//  SimkaMinDistance.hpp:414-640,955-998; export: SimkaMinDistanceMatrixExporter.hpp:330-404; append: SimkaMinAppend.hpp:60-160;
//  info: SimkaMinInfos.hpp:58-100; matrix-update (hidden, used by simkaMin_update.py): SimkaMinDistanceMatrixMerger.hpp:87-127),
// so that the reference's simkaMin.py / simkaMin_update.py wrappers can drive it.
//   sketch file:  u8 kmerSize, u32 sketchSize, u32 seed, u32 nbDatasets (13 bytes, SimkaMinCommons.hpp:28), then per dataset
//                 sketchSize x {u64 hash, u32 count, 4 pad} ascending (zero padding first), then the ids as u8 length + bytes
//   distance:     mat_presenceAbsence_jaccard.bin / mat_abundance_braycurtis.bin, float32 at (i * nbDatasets2 + j) * 4
//   export:       <name>.csv.gz, header ";id..." and rows "id;%f;..."
// -filter (SimkaMinCount.hpp:341-380): first occurrences are remembered exactly instead of in the reference's Bloom filter.
#include <sys/stat.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host/simka_input.hpp"
#include "host/stage_files.hpp"

using namespace simka_b200;

namespace {

struct Entry { uint64_t hash; uint32_t count; uint32_t pad; };
static_assert(sizeof(Entry) == 16, "KmerAndCountType is 16 bytes");
const uint64_t HEADER = 13;

struct SketchHeader { uint8_t k; uint32_t s, seed, n; };

SketchHeader read_header(const std::string &path) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) throw bank::Exception("cannot open sketch file " + path);
    SketchHeader h;
    f.read((char *)&h.k, 1); f.read((char *)&h.s, 4); f.read((char *)&h.seed, 4); f.read((char *)&h.n, 4);
    if (!f) throw bank::Exception("bad sketch file " + path);
    return h;
}

std::vector<std::string> read_ids(const std::string &path, const SketchHeader &h) {
    std::ifstream f(path.c_str(), std::ios::binary);
    f.seekg((std::streamoff)(HEADER + (uint64_t)h.n * h.s * sizeof(Entry)));
    std::vector<std::string> ids;
    for (uint32_t i = 0; i < h.n; i++) {
        uint8_t len = 0;
        f.read((char *)&len, 1);
        std::string s(len, ' ');
        f.read(&s[0], len);
        ids.push_back(s);
    }
    return ids;
}

long long geti(std::map<std::string, std::string> &kv, const char *k, long long def) { return kv.count(k) ? atoll(kv[k].c_str()) : def; }

int do_sketch(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in")) throw bank::Exception("Option '-in' is mandatory");
    const bool abundance_filter = kv.count("-filter") != 0;
    const std::string out = kv.count("-out") ? kv["-out"] : "./simkaMin_kmers.bin";
    const int k = (int)geti(kv, "-kmer-size", 21);
    const uint32_t s = (uint32_t)geti(kv, "-nb-kmers", 100000), seed = (uint32_t)geti(kv, "-seed", 100);
    const uint64_t max_reads = (uint64_t)std::max(0LL, geti(kv, "-max-reads", 0));
    if (k < 2 || k > 31) throw bank::Exception("this build supports -kmer-size 2..31");
    SimkaSequenceFilter filter((size_t)geti(kv, "-min-read-size", 0), kv.count("-min-shannon-index") ? atof(kv["-min-shannon-index"].c_str()) : 0);
    std::vector<SampleSpec> samples = parse_simka_input(kv["-in"]);
    FILE *f = fopen(out.c_str(), "wb");
    if (!f) throw bank::Exception("cannot write " + out);
    const uint8_t k8 = (uint8_t)k;
    const uint32_t n = (uint32_t)samples.size();
    fwrite(&k8, 1, 1, f); fwrite(&s, 4, 1, f); fwrite(&seed, 4, 1, f); fwrite(&n, 4, 1, f);
    std::vector<uint64_t> hashes(s);
    std::vector<uint32_t> counts(s);
    std::vector<Entry> rec(s);
    // Host pipeline: worker threads parse + 2-bit pack whole datasets (files are independent, so this scales with the cores;
    // the reference sketches its datasets in parallel the same way, SimkaMinCount.hpp:1129-1260), the main thread sketches
    // them on the GPU in input order and writes the records; at most `window` datasets are packed ahead.
    struct Prepared {
        std::vector<uint64_t> packed;
        std::vector<uint32_t> bad;
        uint64_t words = 0, reads = 0;
        bool ready = false;
        std::string error;
    };
    std::vector<Prepared> prep(n);
    std::mutex mu;
    std::condition_variable cv;
    size_t next = 0, consumed = 0;
    const long long cores = geti(kv, "-nb-cores", 0);
    const size_t hw = cores > 0 ? (size_t)cores : (size_t)std::max(1u, std::thread::hardware_concurrency());
    const size_t workers = std::max<size_t>(1, std::min<size_t>({hw, (size_t)n, (size_t)16}));
    const size_t window = workers + 2;
    const int pack_threads = (int)std::max<size_t>(1, hw / workers);
    auto worker = [&]() {
        std::string seqs;
        std::vector<uint64_t> off;
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next >= n || next < consumed + window; });
                if (next >= n) return;
                i = next++;
            }
            Prepared &P = prep[i];
            try {
                seqs.clear();
                off.assign(1, 0);
                read_sample_into(samples[i], filter, max_reads, seqs, off, [] {}, k, true);  // (short reads replay the previous read)
                P.reads = off.size() - 1;
                P.words = simka_pack_words(seqs.size(), P.reads);
                P.packed.resize(P.words + 1);
                P.bad.resize(P.words + 1);
                simka_pack_reads(seqs.data(), off.data(), P.reads, P.packed.data(), P.bad.data(), pack_threads);
            } catch (const std::exception &e) { P.error = e.what(); if (P.error.empty()) P.error = "read error"; }
            {
                std::lock_guard<std::mutex> lk(mu);
                P.ready = true;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> threads;
    for (size_t w = 0; w < workers; w++) threads.emplace_back(worker);
    std::string failure;
    for (size_t i = 0; i < n && failure.empty(); i++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return prep[i].ready; });
        }
        Prepared &P = prep[i];
        if (!P.error.empty()) failure = P.error;
        else {
            uint32_t nsel = 0;
            const int rc = (abundance_filter ? simkamin_gpu_sketch_filter : simkamin_gpu_sketch)((int)geti(kv, "-device", 0), k, s, seed, P.packed.data(), P.bad.data(), P.words,
                                                                                       P.reads, hashes.data(), counts.data(), &nsel);
            if (rc != SIMKA_OK) failure = "simkamin_gpu_sketch failed (code " + std::to_string(rc) + "): a CUDA device is required";
            else {
                for (uint32_t q = 0; q < s; q++) { rec[q].hash = hashes[q]; rec[q].count = counts[q]; rec[q].pad = 0; }
                fwrite(rec.data(), sizeof(Entry), s, f);
            }
        }
        std::vector<uint64_t>().swap(P.packed);
        std::vector<uint32_t>().swap(P.bad);
        {
            std::lock_guard<std::mutex> lk(mu);
            consumed = i + 1;
            if (!failure.empty()) next = n;  // stop the workers
        }
        cv.notify_all();
    }
    for (auto &t : threads) t.join();
    if (!failure.empty()) { fclose(f); throw bank::Exception(failure); }
    for (const SampleSpec &sp : samples) {
        const uint8_t len = (uint8_t)sp.id.size();
        fwrite(&len, 1, 1, f);
        fwrite(sp.id.data(), 1, len, f);
    }
    fclose(f);
    return EXIT_SUCCESS;
}

void load_block(const std::string &path, const SketchHeader &h, uint32_t start, uint32_t n, uint32_t s_use, std::vector<uint64_t> &hashes,
                std::vector<uint32_t> &counts) {
    std::ifstream f(path.c_str(), std::ios::binary);
    hashes.assign((size_t)n * s_use, 0);
    counts.assign((size_t)n * s_use, 0);
    std::vector<Entry> rec(s_use);
    for (uint32_t i = 0; i < n; i++) {
        f.seekg((std::streamoff)(HEADER + (uint64_t)(start + i) * h.s * sizeof(Entry)));
        f.read((char *)rec.data(), (std::streamsize)(sizeof(Entry) * s_use));  // the first min(s1,s2) entries (SimkaMinDistance.hpp:536-552)
        if (!f) throw bank::Exception("short read in " + path);
        for (uint32_t q = 0; q < s_use; q++) { hashes[(size_t)i * s_use + q] = rec[q].hash; counts[(size_t)i * s_use + q] = rec[q].count; }
    }
}

int do_distance(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in1") || !kv.count("-in2")) throw bank::Exception("Options '-in1' and '-in2' are mandatory");
    const std::string in1 = kv["-in1"], in2 = kv["-in2"], out = kv.count("-out") ? kv["-out"] : "./simkaMin_results";
    const SketchHeader h1 = read_header(in1), h2 = read_header(in2);
    if (h1.k != h2.k) { std::cerr << "ERROR: can't compare both sketches because of different kmer sizes (" << (int)h1.k << " and " << (int)h2.k << ")" << std::endl; return 1; }
    if (h1.seed != h2.seed) { std::cerr << "ERROR: can't compare both sketches because of different seeds (" << h1.seed << " and " << h2.seed << ")" << std::endl; return 1; }
    uint32_t si = (uint32_t)geti(kv, "-start-i", 0), sj = (uint32_t)geti(kv, "-start-j", 0);
    uint32_t ni = (uint32_t)geti(kv, "-n-i", 0), nj = (uint32_t)geti(kv, "-n-j", 0);
    if (ni == 0) ni = h1.n;
    if (nj == 0) nj = h2.n;
    if (si + ni > h1.n || sj + nj > h2.n) throw bank::Exception("block outside the sketch files");
    const uint32_t s = std::min(h1.s, h2.s);
    stage::mkdirs(out);
    const bool symmetric = in1 == in2 && si == sj;
    std::vector<uint64_t> ha, hb;
    std::vector<uint32_t> ca, cb;
    load_block(in1, h1, si, ni, s, ha, ca);
    std::vector<float> jac, bc;
    int rc;
    const int device = (int)geti(kv, "-device", 0);
    if (symmetric) {
        jac.assign((size_t)ni * ni, 0); bc.assign((size_t)ni * ni, 0);
        rc = simkamin_gpu_distance(device, ha.data(), ca.data(), ni, s, jac.data(), bc.data(), nullptr);
        nj = ni;
    } else {
        load_block(in2, h2, sj, nj, s, hb, cb);
        jac.assign((size_t)ni * nj, 0); bc.assign((size_t)ni * nj, 0);
        rc = simkamin_gpu_distance_rect(device, ha.data(), ca.data(), ni, hb.data(), cb.data(), nj, s, jac.data(), bc.data(), nullptr);
    }
    if (rc != SIMKA_OK) throw bank::Exception("distance kernel failed (code " + std::to_string(rc) + "): a CUDA device is required");
    // float32 at (i * nbDatasets2 + j) * 4 (and mirrored at (j * nbDatasets1 + i) * 4 for a symmetric block), :281-303,617-628
    for (int m = 0; m < 2; m++) {
        const std::string path = out + (m == 0 ? "/mat_presenceAbsence_jaccard.bin" : "/mat_abundance_braycurtis.bin");
        FILE *f = fopen(path.c_str(), "r+b");
        if (!f) f = fopen(path.c_str(), "w+b");
        if (!f) throw bank::Exception("cannot write " + path);
        const std::vector<float> &v = m == 0 ? jac : bc;
        for (uint32_t i = 0; i < ni; i++)
            for (uint32_t j = 0; j < nj; j++) {
                if (symmetric && j < i) continue;
                const float d = (symmetric && i == j) ? 0.0f : v[(size_t)i * nj + j];
                fseeko(f, (off_t)(((uint64_t)(si + i) * h2.n + (sj + j)) * 4), SEEK_SET);
                fwrite(&d, 4, 1, f);
                if (symmetric && j > i) {
                    fseeko(f, (off_t)(((uint64_t)(sj + j) * h1.n + (si + i)) * 4), SEEK_SET);
                    fwrite(&d, 4, 1, f);
                }
            }
        fclose(f);
    }
    return EXIT_SUCCESS;
}

int do_export(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in") || !kv.count("-in1") || !kv.count("-in2")) throw bank::Exception("Options '-in', '-in1' and '-in2' are mandatory");
    const std::string dir = kv["-in"], out = kv.count("-out") ? kv["-out"] : "./simkaMin_results";
    const SketchHeader h1 = read_header(kv["-in1"]), h2 = read_header(kv["-in2"]);
    const std::vector<std::string> ids1 = read_ids(kv["-in1"], h1), ids2 = read_ids(kv["-in2"], h2);
    stage::mkdirs(out);
    for (const char *name : {"mat_abundance_braycurtis", "mat_presenceAbsence_jaccard"}) {
        std::ifstream f((dir + "/" + name + ".bin").c_str(), std::ios::binary);
        if (!f) continue;
        gzFile g = gzopen((out + "/" + name + ".csv.gz").c_str(), "wb");
        if (!g) throw bank::Exception("cannot write " + out + "/" + name + ".csv.gz");
        std::string line;
        for (const std::string &id : ids2) line += ";" + id;
        line += "\n";
        gzwrite(g, line.data(), (unsigned)line.size());
        std::vector<float> row(ids2.size());
        for (size_t i = 0; i < ids1.size(); i++) {
            f.read((char *)row.data(), (std::streamsize)(row.size() * 4));
            line = ids1[i] + ";";
            char buf[64];
            for (size_t j = 0; j < ids2.size(); j++) { snprintf(buf, sizeof buf, "%f", row[j]); line += buf; line += ";"; }
            line.erase(line.size() - 1);
            line += "\n";
            gzwrite(g, line.data(), (unsigned)line.size());
        }
        gzclose(g);
    }
    return EXIT_SUCCESS;
}

// append: the sketches of -in2 are added to -in1 in place (same k, sketch size and seed required): header count updated,
// sketches of -in2 written where the ids of -in1 began, then the ids of -in1 and of -in2
int do_append(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in1") || !kv.count("-in2")) throw bank::Exception("Options '-in1' and '-in2' are mandatory");
    const std::string in1 = kv["-in1"], in2 = kv["-in2"];
    for (const std::string &f : {in1, in2}) {
        struct stat sb;
        if (stat(f.c_str(), &sb) != 0) { std::cerr << "Error: input does not exist (" << f << ")" << std::endl; return 1; }
    }
    const SketchHeader h1 = read_header(in1), h2 = read_header(in2);
    if (h1.k != h2.k) { std::cerr << "Error: can't merge sketches with different kmer sizes (" << (int)h1.k << " vs " << (int)h2.k << ")" << std::endl; return 1; }
    if (h1.s != h2.s) { std::cerr << "Error: can't merge sketches with different sketch sizes (" << h1.s << " vs " << h2.s << ")" << std::endl; return 1; }
    if (h1.seed != h2.seed) { std::cerr << "Error: can't merge sketches with different seeds (" << h1.seed << " vs " << h2.seed << ")" << std::endl; return 1; }
    const std::vector<std::string> ids1 = read_ids(in1, h1), ids2 = read_ids(in2, h2);
    FILE *out = fopen(in1.c_str(), "r+b");
    FILE *src = fopen(in2.c_str(), "rb");
    if (!out || !src) throw bank::Exception("cannot open the sketch files");
    const uint32_t n = h1.n + h2.n;
    fseeko(out, 9, SEEK_SET);  // u8 k, u32 sketch size, u32 seed, then the dataset count
    fwrite(&n, 4, 1, out);
    fseeko(out, (off_t)(HEADER + (uint64_t)h1.n * h1.s * sizeof(Entry)), SEEK_SET);
    fseeko(src, (off_t)HEADER, SEEK_SET);
    std::vector<char> buf(1 << 20);
    for (uint64_t left = (uint64_t)h2.n * h2.s * sizeof(Entry); left > 0;) {
        const size_t want = (size_t)std::min<uint64_t>(left, buf.size());
        if (fread(buf.data(), 1, want, src) != want) throw bank::Exception("short read in " + in2);
        fwrite(buf.data(), 1, want, out);
        left -= want;
    }
    for (const std::vector<std::string> *ids : {&ids1, &ids2})
        for (const std::string &id : *ids) {
            const uint8_t len = (uint8_t)id.size();
            fwrite(&len, 1, 1, out);
            fwrite(id.data(), 1, len, out);
        }
    fclose(src);
    fclose(out);
    return EXIT_SUCCESS;
}

int do_info(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in")) throw bank::Exception("Option '-in' is mandatory");
    const std::string in = kv["-in"];
    struct stat sb;
    if (stat(in.c_str(), &sb) != 0) { std::cerr << "Error: input does not exist (" << in << ")" << std::endl; return 1; }
    const SketchHeader h = read_header(in);
    std::cout << "Sketch info: " << in << "\n\n"
              << "k-mer size  : " << (uint32_t)h.k << "\n"
              << "Sketch size : " << h.s << "\n"
              << "Seed        : " << h.seed << "\n\n"
              << "Nb Datasets: " << h.n << "\n";
    for (const std::string &id : read_ids(in, h)) std::cout << id << "\n";
    std::cout << std::endl;
    return EXIT_SUCCESS;
}

// matrix-update: the binary matrices of -in (E x E, E = datasets of -in1) grow to (E+N) x (E+N) (N = datasets of -in2) from
// -in/existingVsNew/<name> (E x N) and -in/newVsNew/<name> (N x N):  [[existing, eVsN], [eVsN^T, nVsN]]
// (SimkaMinDistanceMatrixExporter.hpp:44-121), written to <name>.temp and renamed over the existing file.
int do_matrix_update(std::map<std::string, std::string> &kv) {
    if (!kv.count("-in") || !kv.count("-in1") || !kv.count("-in2")) throw bank::Exception("Options '-in', '-in1' and '-in2' are mandatory");
    const std::string dir = kv["-in"] + "/";
    const uint64_t E = read_header(kv["-in1"]).n, N = read_header(kv["-in2"]).n, T = E + N;
    auto load = [](const std::string &path, uint64_t rows, uint64_t cols) {
        std::vector<float> m(rows * cols, 0.0f);  // a short file leaves zeros, like the reference's unchecked reads into zeroed rows
        std::ifstream f(path.c_str(), std::ios::binary);
        if (f) f.read((char *)m.data(), (std::streamsize)(m.size() * 4));
        return m;
    };
    for (const std::string &name : stage::listdir(dir)) {
        if (name.find("mat_") == std::string::npos || name.find(".bin") == std::string::npos || name.find(".temp") != std::string::npos) continue;
        std::cout << name << std::endl;
        const std::vector<float> ee = load(dir + name, E, E), en = load(dir + "/existingVsNew/" + name, E, N), nn = load(dir + "/newVsNew/" + name, N, N);
        std::vector<float> out(T * T, 0.0f);
        for (uint64_t i = 0; i < E; i++) {
            for (uint64_t j = 0; j < E; j++) out[i * T + j] = ee[i * E + j];
            for (uint64_t j = 0; j < N; j++) out[i * T + E + j] = out[(E + j) * T + i] = en[i * N + j];
        }
        for (uint64_t i = 0; i < N; i++)
            for (uint64_t j = 0; j < N; j++) out[(E + i) * T + E + j] = nn[i * N + j];
        const std::string tmp = dir + name + ".temp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(out.data(), 4, out.size(), f) != out.size()) { if (f) fclose(f); throw bank::Exception("cannot write " + tmp); }
        fclose(f);
        if (rename(tmp.c_str(), (dir + name).c_str()) != 0) throw bank::Exception("cannot replace " + dir + name);
    }
    return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char **argv) {
    try {
        if (argc < 2) { std::cout << "usage: simkaMinCore sketch|distance|export|append|info|matrix-update [options]" << std::endl; return EXIT_FAILURE; }
        static const char *no_param[] = {"-filter", nullptr};
        auto kv = stage::parse_flags(argc - 1, argv + 1, no_param);
        const std::string cmd = argv[1];
        if (cmd == "sketch") return do_sketch(kv);
        if (cmd == "distance") return do_distance(kv);
        if (cmd == "export") return do_export(kv);
        if (cmd == "append") return do_append(kv);
        if (cmd == "info") return do_info(kv);
        if (cmd == "matrix-update") return do_matrix_update(kv);
        throw bank::Exception("unknown command '" + cmd + "' (sketch, distance, export, append, info, matrix-update)");
    } catch (const std::exception &e) {
        std::cout << "EXCEPTION: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
}
