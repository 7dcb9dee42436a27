// simka_merge_main.cpp -- `simkaMerge`: stage 2 for ONE partition, with the reference's command line and temp files
// (src/SimkaMerge.cpp:1530-1544,940-1305,1451-1476): reads datasetIds, count_synchro/*.ok and
// solid/part_<p>/__p__<i>.gz of every sample (written by this build's simkaCount or by the reference's), merges and
// accumulates on the GPU, writes stats/part_<p>.gz and merge_synchro/<p>.ok.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "host/stage_files.hpp"

using namespace simka_b200;

static int run(int argc, char **argv) {
    static const char *no_param[] = {"-simple-dist", "-complex-dist", nullptr};
    auto kv = stage::parse_flags(argc, argv, no_param);
    auto need = [&](const char *k) { if (!kv.count(k)) throw bank::Exception(std::string("Option '") + k + "' is mandatory"); return kv[k]; };
    const int k = atoi(need("-kmer-size").c_str());
    const std::string dir = need("-out-tmp-simka");
    const size_t p = (size_t)atoi(need("-partition-id").c_str());
    if (k < 2 || k > 31) throw bank::Exception("this build supports -kmer-size 2..31");
    const std::vector<std::string> ids = stage::read_lines(dir + "/datasetIds");
    const size_t N = ids.size();
    if (!N) throw bank::Exception("no dataset in " + dir + "/datasetIds");
    std::vector<simka_sample_totals> totals(N);
    for (size_t i = 0; i < N; i++) totals[i] = stage::read_ok(dir + "/count_synchro/" + ids[i] + ".ok");

    simka_params prm;
    memset(&prm, 0, sizeof prm);
    prm.kmer_size = k; prm.nb_samples = (int32_t)N; prm.abundance_min = 0; prm.abundance_max = 0x7FFFFFFFFFFFFFFFll;
    prm.simple_dist = kv.count("-simple-dist"); prm.complex_dist = kv.count("-complex-dist");
    prm.device = kv.count("-device") ? atoi(kv["-device"].c_str()) : 0;
    simka_ctx *ctx = nullptr;
    if (simka_gpu_create(&ctx, &prm) != SIMKA_OK) throw bank::Exception("simka_gpu_create failed: a CUDA device is required, there is no CPU path");
    std::vector<uint64_t> kmers;
    std::vector<uint32_t> counts;
    for (size_t i = 0; i < N; i++) {
        stage::read_solid_part(stage::part_file(dir, p, i), kmers, counts);
        if (simka_gpu_import_counts(ctx, (int)i, kmers.data(), counts.data(), kmers.size(), totals[i].nb_reads) != SIMKA_OK)
            throw bank::Exception(std::string("import failed: ") + simka_gpu_last_error(ctx));
    }
    if (simka_gpu_set_totals(ctx, totals.data()) != SIMKA_OK || simka_gpu_merge_and_accumulate(ctx) != SIMKA_OK)
        throw bank::Exception(std::string("merge failed: ") + simka_gpu_last_error(ctx));
    const size_t nn = N * N, ns = N * (N + 1) / 2;
    std::vector<uint64_t> u(2 * ns + 6 * nn);
    std::vector<double> kl(nn);
    std::vector<simka_sample_totals> gt(N);
    simka_stats_view v;
    memset(&v, 0, sizeof v);
    v.totals = gt.data();
    v.distinct_shared = &u[0]; v.braycurtis_num = &u[ns]; v.nb_shared = &u[2 * ns]; v.chord_ninj = &u[2 * ns + nn];
    v.hellinger = &u[2 * ns + 2 * nn]; v.kulczynski = &u[2 * ns + 3 * nn]; v.canberra = &u[2 * ns + 4 * nn]; v.whittaker = &u[2 * ns + 5 * nn];
    v.kullback_leibler = kl.data();
    if (simka_gpu_get_stats(ctx, &v) != SIMKA_OK) throw bank::Exception(std::string("get_stats failed: ") + simka_gpu_last_error(ctx));
    simka_gpu_destroy(ctx);
    stage::mkdirs(dir + "/stats");
    stage::mkdirs(dir + "/merge_synchro");
    stage::write_stats(dir + "/stats/part_" + std::to_string(p) + ".gz", v);
    { FILE *f = fopen((dir + "/merge_synchro/" + std::to_string(p) + ".ok").c_str(), "w"); if (f) fclose(f); }
    return EXIT_SUCCESS;
}

int main(int argc, char **argv) {
    try { return run(argc, argv); }
    catch (const std::exception &e) { std::cout << "EXCEPTION: " << e.what() << std::endl; return EXIT_FAILURE; }
}
