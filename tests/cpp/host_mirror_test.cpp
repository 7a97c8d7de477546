// C++ host-mirror test driver: reads like the reference's flow in src/main.rs:109-163 and checks
// the GPU result against the CPU oracle (oracle/bsr_oracle.c, test infrastructure) bit for bit.
//   usage: host_mirror_test cpu   -> host logic only (no device)
//          host_mirror_test gpu   -> full search on cuda:0
#include <cstdio>
#include <cstring>
#include <string>

#include "../../better_search_rag_rust_b200/csrc/bsr_host.hpp"

extern "C" {
void bsr_oracle_interval_by_rank(int32_t, int32_t, uint64_t, uint64_t*, uint64_t*);
void bsr_oracle_generate_rows(uint64_t, uint64_t, uint64_t, uint32_t, float*);
uint64_t bsr_oracle_parallel_top_k_similarity_search(const float*, uint64_t, uint32_t, int32_t, uint64_t,
                                                     const float*, uint32_t, int32_t, uint64_t*, float*);
}

using namespace bsr_host;

#define REQUIRE(c)                                                      \
    do {                                                                \
        if (!(c)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);  \
            return 1;                                                   \
        }                                                               \
    } while (0)

static int test_cpu() {
    for (size_t count : {0ul, 1ul, 5ul, 9ul, 10ul, 1027ul, 100000000ul})
        for (int size : {1, 2, 4, 7, 8, 41})
            for (int rank = 0; rank < size; rank++) {
                uint64_t s, e;
                bsr_oracle_interval_by_rank(rank, size, count, &s, &e);
                RankInterval iv = interval_by_rank(rank, size, count);
                REQUIRE(iv.start_index == s && iv.end_index == e);
            }
    REQUIRE(interval_by_rank(3, 4, 5).rows(5).second == 0);
    TopK top = {{4, 0.1f}, {0, 0.2f}};
    Accuracy a = calculate_accuracy_metrics(top, 0, 50);
    REQUIRE(a.mrr == 0.5f && a.recall == 1.0f && a.overlap == 1.0f);
    a = calculate_accuracy_metrics(top, 9, 50);
    REQUIRE(a.mrr == 0.0f && a.recall == 0.0f && a.overlap == 0.0f);
    // without a device the store constructor must throw, not fall back
    int n_dev = 0;
    if (bsr_device_count(&n_dev) != BSR_OK || n_dev == 0) {
        bool threw = false;
        try {
            GpuVectorstore s(0, 768);
        } catch (const Error& e) {
            threw = e.status == BSR_ERR_CUDA;
        }
        REQUIRE(threw);
    }
    std::printf("cpu ok\n");
    return 0;
}

static int test_gpu() {
    const uint64_t seed = 0x5EED0001, n = 20011;
    const uint32_t dim = 768;
    const size_t top_k = 50;  // src/main.rs:110
    std::vector<float> rows(n * dim);
    bsr_oracle_generate_rows(seed, 0, n, dim, rows.data());

    GpuVectorstore store(0, dim);
    std::vector<std::vector<float>> first(1000, std::vector<float>(dim));
    for (size_t i = 0; i < 1000; i++) std::memcpy(first[i].data(), &rows[i * dim], dim * 4);
    store.append_many(first);                           // host rows
    store.append_synthetic(seed, 1000, n - 1000);       // the rest generated on device
    REQUIRE(store.get_count() == n);
    auto back = store.get_many(SliceArgs{19990, 100});  // clamped
    REQUIRE(back.size() == 21 && std::memcmp(back[20].data(), &rows[(n - 1) * dim], dim * 4) == 0);

    LocalWorld world;
    std::vector<float> target = store.get(0);           // src/main.rs:120
    auto top = parallel_top_k_similarity_search(world, 0, 1, store, top_k, target);
    REQUIRE(top.has_value() && top->size() == top_k);
    std::vector<uint64_t> wi(top_k);
    std::vector<float> wd(top_k);
    uint64_t wn = bsr_oracle_parallel_top_k_similarity_search(rows.data(), n, dim, 1, top_k, target.data(),
                                                              dim, 1, wi.data(), wd.data());
    REQUIRE(wn == top_k);
    for (size_t i = 0; i < top_k; i++) {
        REQUIRE((*top)[i].first == wi[i]);
        REQUIRE(std::memcmp(&(*top)[i].second, &wd[i], 4) == 0);
    }
    Accuracy a = calculate_accuracy_metrics(*top, 0, top_k);
    REQUIRE(a.mrr == 1.0f && a.recall == 1.0f && a.overlap == 1.0f);

    // six "ranks" on one GPU: shards by interval_by_rank, gather in rank order, merge on device
    std::vector<size_t> gi;
    std::vector<float> gd;
    for (int r = 0; r < 6; r++) {
        RankInterval iv = interval_by_rank(r, 6, n);
        auto [f, c] = iv.rows(n);
        GpuVectorstore shard(0, dim, iv.start_index, c);
        shard.append_synthetic(seed, f, c);
        for (auto& p : compute_local_top_k(shard, top_k, target)) {
            gi.push_back(p.first);
            gd.push_back(p.second);
        }
    }
    TopK merged = compute_global_top_k(store, gi, gd, top_k);
    wn = bsr_oracle_parallel_top_k_similarity_search(rows.data(), n, dim, 6, top_k, target.data(), dim, 2,
                                                     wi.data(), wd.data());
    REQUIRE(merged.size() == wn);
    for (size_t i = 0; i < wn; i++)
        REQUIRE(merged[i].first == wi[i] && std::memcmp(&merged[i].second, &wd[i], 4) == 0);
    print_top_k_results(TopK(merged.begin(), merged.begin() + 3));
    std::printf("gpu ok\n");
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    try {
        return mode == "gpu" ? test_gpu() : test_cpu();
    } catch (const Error& e) {
        std::printf("bsr error %d: %s\n", e.status, e.what());
        return 2;
    }
}
