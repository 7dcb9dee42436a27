// simka_count_main.cpp -- `simkaCount`: stage 1 for ONE sample, with the reference's command line and temp files
// (src/SimkaCount.cpp:80-98,174-368): reads <out-tmp-simka>/input/<bank-name>, counts on the GPU, writes
// solid/part_<p>/__p__<bank-index>.gz, kmercount_per_partition/<bank-name>.txt and count_synchro/<bank-name>.ok.
// When the temp directory holds the config.h5 of a reference orchestrator, the k-mers go to the partitions ITS simkaCount would
// send them to (repartition table of minimizers, host/repartitor.hpp), so samples counted here and by the reference can share one
// temp directory; without it the partition is a fixed hash (any function merges correctly as long as every sample uses it).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>

#include "host/kmer_model.hpp"
#include "host/repartitor.hpp"
#include "host/simka_input.hpp"
#include "host/stage_files.hpp"

using namespace simka_b200;

static int run(int argc, char **argv) {
    static const char *no_param[] = {nullptr};
    auto kv = stage::parse_flags(argc, argv, no_param);
    auto need = [&](const char *k) { if (!kv.count(k)) throw bank::Exception(std::string("Option '") + k + "' is mandatory"); return kv[k]; };
    const int k = atoi(need("-kmer-size").c_str());
    const std::string dir = need("-out-tmp-simka"), name = need("-bank-name");
    const uint32_t bank = (uint32_t)atoi(need("-bank-index").c_str());
    const size_t nb_groups = (size_t)std::max(1, atoi(need("-nb-datasets").c_str()));
    const size_t P = (size_t)std::max(1, atoi(need("-nb-partitions").c_str()));
    const long long amin = kv.count("-abundance-min") ? atoll(kv["-abundance-min"].c_str()) : 2;
    const long long amax = kv.count("-abundance-max") ? atoll(kv["-abundance-max"].c_str()) : 999999999;
    const uint64_t max_reads = kv.count("-max-reads") ? (uint64_t)std::max(0LL, atoll(kv["-max-reads"].c_str())) : 0;
    SimkaSequenceFilter filter(kv.count("-min-read-size") ? (size_t)atoll(kv["-min-read-size"].c_str()) : 0,
                               kv.count("-min-shannon-index") ? atof(kv["-min-shannon-index"].c_str()) : 0);
    if (k < 2 || k > 31) throw bank::Exception("this build supports -kmer-size 2..31");
    // the sample's album lists all its files; the paired groups are equal slices of it (src/core/SimkaCommons.hpp:166)
    SampleSpec spec;
    spec.id = name;
    std::vector<std::string> files = stage::read_lines(dir + "/input/" + name);
    const size_t per = std::max<size_t>(1, files.size() / nb_groups);
    for (size_t g = 0; g * per < files.size(); g++)
        spec.groups.emplace_back(files.begin() + g * per, files.begin() + std::min(files.size(), (g + 1) * per));

    simka_params prm;
    memset(&prm, 0, sizeof prm);
    prm.kmer_size = k; prm.nb_samples = 1; prm.abundance_min = amin; prm.abundance_max = amax;
    prm.device = kv.count("-device") ? atoi(kv["-device"].c_str()) : 0;
    simka_ctx *ctx = nullptr;
    if (simka_gpu_create(&ctx, &prm) != SIMKA_OK) throw bank::Exception("simka_gpu_create failed: a CUDA device is required, there is no CPU path");
    std::string seqs;
    std::vector<uint64_t> off(1, 0);
    const uint64_t nreads = read_sample_into(spec, filter, max_reads, seqs, off, [] {}, k);
    const uint64_t words = simka_pack_words(seqs.size(), off.size() - 1);
    std::vector<uint64_t> packed(words + 1);
    std::vector<uint32_t> bad(words + 1);
    simka_pack_reads_k(seqs.data(), off.data(), off.size() - 1, packed.data(), bad.data(), kv.count("-nb-cores") ? atoi(kv["-nb-cores"].c_str()) : 0, k);
    simka_sample_totals tot;
    if (simka_gpu_push_reads(ctx, 0, packed.data(), bad.data(), words, off.size() - 1) != SIMKA_OK || simka_gpu_finish_sample(ctx, 0, &tot) != SIMKA_OK)
        throw bank::Exception(std::string("counting failed: ") + simka_gpu_last_error(ctx));
    tot.nb_reads = nreads;
    uint64_t n = 0;
    simka_gpu_dump_sample_counts(ctx, 0, nullptr, nullptr, &n);
    std::vector<uint64_t> kmers(std::max<uint64_t>(n, 1));
    std::vector<uint32_t> counts(std::max<uint64_t>(n, 1));
    if (simka_gpu_dump_sample_counts(ctx, 0, kmers.data(), counts.data(), &n) != SIMKA_OK) throw bank::Exception("dump failed");
    simka_gpu_destroy(ctx);
    // split the ascending list into the P partition files (each stays ascending)
    std::vector<std::vector<uint64_t>> pk(P);
    std::vector<std::vector<uint32_t>> pc(P);
    Repartitor rep;
    {
        std::ifstream probe((dir + "/config.h5").c_str(), std::ios::binary);
        if (probe) {
            probe.close();
            rep = Repartitor::load(dir + "/config.h5");  // (a damaged file is an error, not a reason to fall back silently)
            if (rep.nbpart != P) throw bank::Exception("config.h5 holds " + std::to_string(rep.nbpart) + " partitions, -nb-partitions says " + std::to_string(P));
        }
    }
    if (rep.nbpart) {
        Kmer<>::ModelMinimizer<> model((size_t)k, (size_t)rep.minimizer_size);
        for (uint64_t i = 0; i < n; i++) {
            const size_t p = rep.table[(size_t)model.getMinimizerValue(Kmer<>::Type(kmers[i]))];
            pk[p].push_back(kmers[i]);
            pc[p].push_back(counts[i]);
        }
    } else {
        for (uint64_t i = 0; i < n; i++) { const size_t p = stage::partition_of(kmers[i], P); pk[p].push_back(kmers[i]); pc[p].push_back(counts[i]); }
    }
    std::string per_part;
    for (size_t p = 0; p < P; p++) {
        stage::mkdirs(dir + "/solid/part_" + std::to_string(p));
        stage::write_solid_part(stage::part_file(dir, p, bank), bank, pk[p].data(), pc[p].data(), pk[p].size());
        per_part += std::to_string(pk[p].size()) + "\n";
    }
    stage::mkdirs(dir + "/kmercount_per_partition");
    stage::mkdirs(dir + "/count_synchro");
    { FILE *f = fopen((dir + "/kmercount_per_partition/" + name + ".txt").c_str(), "w"); if (f) { fputs(per_part.c_str(), f); fclose(f); } }
    stage::write_ok(dir + "/count_synchro/" + name + ".ok", tot);  // last: it is the barrier token
    return EXIT_SUCCESS;
}

int main(int argc, char **argv) {
    try { return run(argc, argv); }
    catch (const std::exception &e) { std::cout << "EXCEPTION: " << e.what() << std::endl; return EXIT_FAILURE; }
}
