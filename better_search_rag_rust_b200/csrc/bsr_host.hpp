// bsr_host.hpp -- C++17 host-side mirror of the reference's search interface, above the C ABI.
//
// The reference is compiled Rust and this image has no Rust toolchain, so this header is the
// compiled-language host for libbsr_b200.so: same names, argument meaning and error behaviour as
// the Rust items it mirrors (citations relative to the reference tree).  Header-only; link with
// -lbsr_b200.  Parquet decode stays outside (polars in the Rust host, pyarrow in the Python one).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bsr.h"

namespace bsr_host {

// PolarsError stand-in: every failed C-ABI call becomes one of these (never a CPU fallback).
struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};
inline void check(int status) {
    if (status != BSR_OK) throw Error(status, bsr_last_error());
}

constexpr int ROOT = 0;  // src/mpi_helpers/mod.rs
inline bool is_root(int rank) { return rank == ROOT; }

// ---- src/mpi_helpers/load_balance.rs:8-42 --------------------------------------------------------
struct RankInterval {
    size_t start_index, end_index;
    size_t get_count() const { return end_index - start_index; }  // wraps like a release build
    // rows the polars slice really yields (src/mpi_helpers/metrics.rs:30-33, clamping slice)
    std::pair<size_t, size_t> rows(size_t count) const {
        if (start_index >= count) return {count, 0};
        return {start_index, std::min(get_count(), count - start_index)};
    }
};
inline RankInterval interval_by_rank(int rank, int size, size_t count) {
    const size_t usize = (size_t)size;
    const size_t per_rank = usize > count ? 1 : (count + usize - 1) / usize;
    const size_t start = per_rank * (size_t)rank;
    const size_t end = rank == size - 1 ? count : std::min(start + per_rank, count);
    return {start, end};
}

// ---- src/vectorstore/polars.rs:7-247 ---------------------------------------------------------------
struct SliceArgs {
    int32_t offset;
    size_t length;
};

class GpuVectorstore {
  public:
    // PolarsVectorstore::new(path, empty = true); row_base = RankInterval::start_index of the rank
    GpuVectorstore(int device, uint32_t dim, uint64_t row_base = 0, uint64_t capacity_rows = 0)
        : dim_(dim) {
        check(bsr_store_create(device, dim, BSR_F32, row_base, capacity_rows, &h_));
    }
    ~GpuVectorstore() { bsr_store_destroy(h_); }
    GpuVectorstore(const GpuVectorstore&) = delete;
    GpuVectorstore& operator=(const GpuVectorstore&) = delete;

    // BSR_COSINE: the reference's cosine_distance (default); BSR_DOT: 1 - dot product
    void set_metric(bsr_metric m) { check(bsr_store_set_metric(h_, m)); }
    bsr_metric metric() const {
        bsr_metric m;
        check(bsr_store_metric(h_, &m));
        return m;
    }
    void set_option(const char* name, int64_t value) { check(bsr_store_set_option(h_, name, value)); }
    void reset() { check(bsr_store_reset(h_)); }                                  // :93-95
    void append(const std::vector<float>& v) { append_many({v}); }                // :97-99
    void append_many(const std::vector<std::vector<float>>& vectors) {            // :101-119
        if (vectors.empty()) return;
        std::vector<float> flat;
        flat.reserve(vectors.size() * dim_);
        for (const auto& v : vectors) {
            if (v.size() != dim_) throw Error(BSR_ERR_INVALID, "row length does not match the store");
            flat.insert(flat.end(), v.begin(), v.end());
        }
        check(bsr_store_append(h_, flat.data(), vectors.size()));
    }
    void append_synthetic(uint64_t seed, uint64_t first_row, uint64_t n) {        // src/utils.rs:14-35
        check(bsr_store_append_synthetic(h_, seed, first_row, n));
    }
    size_t get_count() const {                                                    // :243-246
        uint64_t n = 0;
        check(bsr_store_count(h_, &n));
        return (size_t)n;
    }
    std::vector<std::vector<float>> get_many(std::optional<SliceArgs> slice = std::nullopt) const {
        const int64_t count = (int64_t)get_count();                               // :121-156
        int64_t start = 0, len = count;
        if (slice) {  // DataFrame::slice clamps; a negative offset counts from the end
            const int64_t off = slice->offset;
            start = off < 0 ? std::max<int64_t>(0, count + off) : std::min(off, count);
            len = std::max<int64_t>(0, std::min<int64_t>((int64_t)slice->length, count - start));
        }
        std::vector<float> flat((size_t)len * dim_);
        if (len) check(bsr_store_get_many(h_, (uint64_t)start, (uint64_t)len, flat.data()));
        std::vector<std::vector<float>> out((size_t)len);
        for (int64_t i = 0; i < len; i++) out[i].assign(flat.begin() + i * dim_, flat.begin() + (i + 1) * dim_);
        return out;
    }
    std::vector<float> get(size_t index) const {                                  // :158-169
        auto rows = get_many(SliceArgs{(int32_t)index, 1});
        if (rows.empty()) throw Error(BSR_ERR_RANGE, "Index not found");
        return rows[0];
    }
    bsr_store* handle() const { return h_; }
    uint32_t dim() const { return dim_; }

  private:
    bsr_store* h_ = nullptr;
    uint32_t dim_;
};

using TopK = std::vector<std::pair<size_t, float>>;

// ---- src/mpi_helpers/metrics.rs:16-53, over the rank's HBM-resident shard --------------------------
inline TopK compute_local_top_k(GpuVectorstore& shard, size_t top_k, const std::vector<float>& target) {
    const size_t n = shard.get_count();
    TopK out;
    if (n == 0) return out;
    if (target.size() != shard.dim()) {  // cosine_distance == 1.0 for every row (src/metrics.rs:144-146)
        uint64_t base = 0;
        check(bsr_store_row_base(shard.handle(), &base));
        for (size_t i = 0; i < std::min(n, top_k); i++) out.emplace_back((size_t)base + i, 1.0f);
        return out;
    }
    std::vector<uint64_t> idx(top_k);
    std::vector<float> dist(top_k);
    uint32_t count = 0;
    check(bsr_search(shard.handle(), target.data(), 1, (uint32_t)top_k, idx.data(), dist.data(), &count));
    for (uint32_t i = 0; i < count; i++) out.emplace_back((size_t)idx[i], dist[i]);
    return out;
}

// ---- the rsmpi Communicator as an interface; one process per GPU implements it with its own
//      transport (MPI, NCCL, ...).  LocalWorld is the size-1 case. ---------------------------------
struct Communicator {
    virtual ~Communicator() = default;
    virtual int rank() const = 0;
    virtual int size() const = 0;
    // every rank contributes `bytes`; root receives size()*bytes in rank order, others may get nothing
    virtual std::vector<uint8_t> gather_to_root(const void* data, size_t bytes) = 0;
    virtual void barrier() {}
};
struct LocalWorld : Communicator {
    int rank() const override { return 0; }
    int size() const override { return 1; }
    std::vector<uint8_t> gather_to_root(const void* data, size_t bytes) override {
        const uint8_t* p = (const uint8_t*)data;
        return std::vector<uint8_t>(p, p + bytes);
    }
};

// ---- src/mpi_helpers/metrics.rs:56-138: rank-ordered concatenation at root -------------------------
inline std::pair<std::vector<size_t>, std::vector<float>> gather_top_k_results(
    Communicator& world, int rank, const TopK& local_top_k, size_t top_k) {
    // fixed slots per rank: [count | idx[top_k] | dist[top_k]] so one gather replaces the count
    // all-gather + two sends per rank (:68,:110-131)
    std::vector<uint64_t> slot(1 + top_k + (top_k + 1) / 2, 0);
    slot[0] = local_top_k.size();
    float* d = reinterpret_cast<float*>(slot.data() + 1 + top_k);
    for (size_t i = 0; i < local_top_k.size(); i++) {
        slot[1 + i] = local_top_k[i].first;
        d[i] = local_top_k[i].second;
    }
    auto all = world.gather_to_root(slot.data(), slot.size() * sizeof(uint64_t));
    std::vector<size_t> gi;
    std::vector<float> gd;
    if (!is_root(rank)) return {gi, gd};
    for (int r = 0; r < world.size(); r++) {
        const uint64_t* s = reinterpret_cast<const uint64_t*>(all.data()) + (size_t)r * slot.size();
        const float* sd = reinterpret_cast<const float*>(s + 1 + top_k);
        for (uint64_t i = 0; i < s[0]; i++) {
            gi.push_back((size_t)s[1 + i]);
            gd.push_back(sd[i]);
        }
    }
    return {gi, gd};
}

// ---- src/mpi_helpers/metrics.rs:141-171: sorted on the GPU ------------------------------------------
inline TopK compute_global_top_k(GpuVectorstore& scratch, const std::vector<size_t>& global_indices,
                                 const std::vector<float>& global_distances, size_t top_k) {
    TopK out;
    if (global_indices.empty()) return out;
    std::vector<uint64_t> gi(global_indices.begin(), global_indices.end()), idx(top_k);
    std::vector<float> dist(top_k);
    uint32_t count = 0;
    check(bsr_merge_host(scratch.handle(), gi.data(), global_distances.data(), gi.size(), (uint32_t)top_k,
                         idx.data(), dist.data(), &count));
    for (uint32_t i = 0; i < count; i++) out.emplace_back((size_t)idx[i], dist[i]);
    return out;
}

// ---- src/mpi_helpers/metrics.rs:174-206: Some(top_k) on root, nullopt elsewhere ---------------------
inline std::optional<TopK> parallel_top_k_similarity_search(Communicator& world, int rank, int /*size*/,
                                                            GpuVectorstore& shard, size_t top_k,
                                                            const std::vector<float>& target_vector) {
    TopK local;
    try {
        local = compute_local_top_k(shard, top_k, target_vector);
    } catch (const Error& e) {
        if (e.status == BSR_ERR_CUDA || e.status == BSR_ERR_OOM) throw;  // no silent degradation
        std::printf("[Rank %d] Error in similarity search: %s\n", rank, e.what());  // :185-191
    }
    auto [gi, gd] = gather_top_k_results(world, rank, local, top_k);
    world.barrier();
    if (!is_root(rank)) return std::nullopt;
    return compute_global_top_k(shard, gi, gd, top_k);
}

// ---- the same with the library's own NCCL communicator (bsr_comm_*): scan, all-gather of the
//      candidate slots over NVLink, on-device merge -- one C-ABI call per rank -------------------------
class DeviceWorld {
  public:
    // rank 0 creates the id; ship the 128 bytes to every rank over the host's channel
    static std::vector<uint8_t> unique_id() {
        std::vector<uint8_t> id(BSR_COMM_ID_BYTES);
        check(bsr_comm_unique_id(id.data()));
        return id;
    }
    DeviceWorld(const std::vector<uint8_t>& id, int n_ranks, int rank, int device) : rank_(rank), size_(n_ranks) {
        if (id.size() != BSR_COMM_ID_BYTES) throw Error(BSR_ERR_INVALID, "unique id must be 128 bytes");
        check(bsr_comm_create(id.data(), n_ranks, rank, device, &c_));
    }
    ~DeviceWorld() { bsr_comm_destroy(c_); }
    DeviceWorld(const DeviceWorld&) = delete;
    DeviceWorld& operator=(const DeviceWorld&) = delete;
    int rank() const { return rank_; }
    int size() const { return size_; }
    bsr_comm* handle() const { return c_; }
    // single-query top_k <= 32 searches go through NVLink peer stores fused into the scan kernel
    bool uses_mailbox() const {
        int yes = 0;
        check(bsr_comm_uses_mailbox(c_, &yes));
        return yes != 0;
    }

  private:
    bsr_comm* c_ = nullptr;
    int rank_, size_;
};

inline std::optional<TopK> parallel_top_k_similarity_search(DeviceWorld& world, int rank, int /*size*/,
                                                            GpuVectorstore& shard, size_t top_k,
                                                            const std::vector<float>& target_vector) {
    if (target_vector.size() != shard.dim()) throw Error(BSR_ERR_INVALID, "query length differs from the store");
    std::vector<uint64_t> idx(top_k);
    std::vector<float> dist(top_k);
    uint32_t count = 0;
    check(bsr_search_parallel(shard.handle(), world.handle(), target_vector.data(), 1, (uint32_t)top_k,
                              is_root(rank) ? idx.data() : nullptr, is_root(rank) ? dist.data() : nullptr,
                              is_root(rank) ? &count : nullptr));
    if (!is_root(rank)) return std::nullopt;
    TopK out;
    for (uint32_t i = 0; i < count; i++) out.emplace_back((size_t)idx[i], dist[i]);
    return out;
}

// ---- src/mpi_helpers/metrics.rs:209-249 --------------------------------------------------------------
inline void print_top_k_results(const TopK& r) {
    std::printf("Global top-%zu results:\n", r.size());
    for (size_t i = 0; i < r.size(); i++)
        std::printf("  %zu. Index: %zu, Distance: %.9g\n", i + 1, r[i].first, r[i].second);
}
struct Accuracy {
    float mrr, recall, overlap;
};
inline Accuracy calculate_accuracy_metrics(const TopK& r, size_t query_idx, size_t top_k) {
    size_t position = 0;
    for (size_t i = 0; i < r.size(); i++)
        if (r[i].first == query_idx) {
            position = i + 1;
            break;
        }
    return {position ? 1.0f / (float)position : 0.0f, (position && position <= top_k) ? 1.0f : 0.0f,
            position ? 1.0f : 0.0f};
}

}  // namespace bsr_host
