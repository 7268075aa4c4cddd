// hannoy_b200.hpp — C++17 host-side mirror of hannoy's search API over the C-ABI (hannoy_b200.h).
//
// hannoy is a Rust crate; this image has no Rust toolchain, so the compiled host side that sits above
// the C-ABI is this header (the Rust binding itself is shipped as source under rust/hannoy-b200/).
// Names, argument meaning and error behaviour follow the reference:
//   hannoy::Reader<D>::open            src/reader.rs:387-431   (MissingMetadata / UnmatchingDistance / NeedBuild)
//   Reader::nns(count) -> QueryBuilder src/reader.rs:611-620
//   QueryBuilder::ef_search/candidates/linear_below/linear_below_ratio   src/reader.rs:200-260
//   QueryBuilder::by_vector / by_item  src/reader.rs:81-89,132-148  (+ batched by_vectors / by_items)
//   Searched{nns, did_cancel}::into_nns src/reader.rs:36-57
//   distances::{Euclidean, Cosine, Manhattan, Hamming, BinaryQuantized*}  src/distance/*.rs (D::name())
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hannoy_b200.h"

namespace hannoy {

using ItemId = uint32_t;  // src/lib.rs:134

// ---- src/error.rs ---------------------------------------------------------------------------------
struct Error : std::runtime_error {
    hb_status status;
    Error(hb_status st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};
struct InvalidVecDimension : Error { using Error::Error; };
struct MissingMetadata : Error { using Error::Error; };
struct UnmatchingDistance : Error { using Error::Error; };
struct NeedBuild : Error { using Error::Error; };

inline void check(hb_status st) {
    if (st == HB_OK) return;
    std::string msg = hb_last_error();
    switch (st) {
        case HB_EDIM: throw InvalidVecDimension(st, msg);
        case HB_EMISSING_METADATA: throw MissingMetadata(st, msg);
        case HB_EUNMATCHING_DISTANCE: throw UnmatchingDistance(st, msg);
        case HB_ENEED_BUILD: throw NeedBuild(st, msg);
        default: throw Error(st, msg);
    }
}

// ---- src/distance/*.rs: tag types, `D::name()` is what the metadata stores -----------------------------
namespace distances {
struct Euclidean { static constexpr hb_metric metric = HB_EUCLIDEAN; };
struct Cosine { static constexpr hb_metric metric = HB_COSINE; };
struct Manhattan { static constexpr hb_metric metric = HB_MANHATTAN; };
struct Hamming { static constexpr hb_metric metric = HB_HAMMING; };
struct BinaryQuantizedCosine { static constexpr hb_metric metric = HB_BQ_COSINE; };
struct BinaryQuantizedEuclidean { static constexpr hb_metric metric = HB_BQ_EUCLIDEAN; };
struct BinaryQuantizedManhattan { static constexpr hb_metric metric = HB_BQ_MANHATTAN; };
template <class D>
inline const char* name() { return hb_metric_name(D::metric); }
}  // namespace distances

// ---- src/reader.rs:36-57 ---------------------------------------------------------------------------------
struct Searched {
    std::vector<std::pair<ItemId, float>> nns;
    bool did_cancel_ = false;  // set when the query carried a CancelToken / poll budget and was interrupted
    bool did_cancel() const { return did_cancel_; }
    std::vector<std::pair<ItemId, float>> into_nns() && { return std::move(nns); }
};

template <class D>
class Reader;

// What a `cancel_fn` closure (reader.rs:108-118,167-188) becomes on this side of the ABI: a flag on the device that any
// host thread may trip while searches carrying it are in flight.
class CancelToken {
  public:
    explicit CancelToken(int device = 0) { check(hb_cancel_token_create(device, &t_)); }
    CancelToken(const CancelToken&) = delete;
    ~CancelToken() { hb_cancel_token_free(t_); }
    void cancel() { check(hb_cancel_token_cancel(t_)); }
    void reset() { check(hb_cancel_token_reset(t_)); }
    bool is_cancelled() const { return hb_cancel_token_is_cancelled(t_) != 0; }
    const hb_cancel_token* raw() const { return t_; }

  private:
    hb_cancel_token* t_ = nullptr;
};

// ---- src/reader.rs:60-261 ----------------------------------------------------------------------------------
template <class D>
class QueryBuilder {
  public:
    QueryBuilder& ef_search(size_t ef) { ef_ = ef > count_ ? ef : count_; return *this; }  // reader.rs:217-220
    QueryBuilder& candidates(const std::vector<ItemId>& c) { cand_ = &c; return *this; }
    QueryBuilder& linear_below(size_t threshold) { linear_below_ = threshold; return *this; }
    QueryBuilder& linear_below_ratio(float ratio) { linear_below_ratio_ = ratio; return *this; }
    // *_with_cancellation: the searches of this builder poll `token` where the reference polls cancel_fn (reader.rs:330,684)
    QueryBuilder& with_cancellation(const CancelToken& token) { cancel_ = token.raw(); return *this; }
    // the deterministic closure "true from its n-th call on" (0 = never)
    QueryBuilder& cancel_after_polls(uint64_t n) { cancel_after_ = n; return *this; }

    // nq x dimensions row-major; one kernel launch for the whole batch
    std::vector<Searched> by_vectors(const float* q, size_t nq, size_t dims) const {
        Out o(nq, count_);
        hb_query_opts opts = make_opts();
        check(hb_search_by_vector(reader_->raw(), q, nq, (uint32_t)dims, (uint32_t)count_, (uint32_t)ef_, &opts, o.ids.data(),
                                  o.dist.data(), o.len.data(), nullptr));
        std::vector<Searched> res(nq);
        for (size_t i = 0; i < nq; ++i) res[i] = o.get(i, count_);
        return res;
    }
    Searched by_vector(const std::vector<float>& v) const { return std::move(by_vectors(v.data(), 1, v.size())[0]); }

    std::vector<std::optional<Searched>> by_items(const std::vector<ItemId>& items) const {
        Out o(items.size(), count_);
        hb_query_opts opts = make_opts();
        check(hb_search_by_item(reader_->raw(), items.data(), items.size(), (uint32_t)count_, (uint32_t)ef_, &opts, o.ids.data(),
                                o.dist.data(), o.len.data(), nullptr));
        std::vector<std::optional<Searched>> res(items.size());
        for (size_t i = 0; i < items.size(); ++i)
            if (o.len[i] != HB_LEN_NONE) res[i] = o.get(i, count_);  // reader.rs:826: None if the item is absent
        return res;
    }
    std::optional<Searched> by_item(ItemId item) const { return std::move(by_items({item})[0]); }

  private:
    friend class Reader<D>;
    struct Out {
        std::vector<uint32_t> ids, len;
        std::vector<float> dist;
        Out(size_t nq, size_t k) : ids(nq * k), len(nq), dist(nq * k) {}
        Searched get(size_t i, size_t k) const {
            Searched s;
            s.did_cancel_ = (len[i] & HB_LEN_CANCELLED) != 0;
            for (uint32_t j = 0; j < (len[i] & ~HB_LEN_CANCELLED); ++j) s.nns.emplace_back(ids[i * k + j], dist[i * k + j]);
            return s;
        }
    };
    QueryBuilder(const Reader<D>* r, size_t count) : reader_(r), count_(count) {}
    hb_query_opts make_opts() const {
        hb_query_opts o{};
        o.candidates = cand_ ? cand_->data() : nullptr;
        o.n_candidates = cand_ ? cand_->size() : 0;
        o.has_candidates = cand_ != nullptr;
        o.linear_below = (uint32_t)linear_below_;
        o.linear_below_ratio = linear_below_ratio_;
        o.cancel = cancel_;
        o.cancel_after_polls = cancel_after_;
        return o;
    }
    const Reader<D>* reader_;
    const std::vector<ItemId>* cand_ = nullptr;
    size_t count_;
    size_t ef_ = 100;               // DEFAULT_EF_SEARCH, reader.rs:23
    size_t linear_below_ = 1000;    // reader.rs:29
    float linear_below_ratio_ = 1.0f;  // reader.rs:32
    const hb_cancel_token* cancel_ = nullptr;
    uint64_t cancel_after_ = 0;
};

// ---- src/reader.rs:374-431,545-620 ---------------------------------------------------------------------------
template <class D>
class Reader {
  public:
    // `KvCursor` yields the raw (key, value) byte pairs of the LMDB read transaction, e.g. a
    // std::vector<std::pair<std::string, std::string>>; the transaction is read exactly once.
    template <class KvCursor>
    static Reader open(const KvCursor& kv, uint16_t index, int device = 0) {
        Reader r;
        check(hb_index_begin(D::metric, index, &r.ix_));
        for (const auto& [k, v] : kv)
            check(hb_index_push_kv(r.ix_, (const uint8_t*)k.data(), k.size(), (const uint8_t*)v.data(), v.size()));
        check(hb_index_finalize(r.ix_, device));
        return r;
    }
    // Reader::open from the LMDB environment on disk (directory holding data.mdb, or the data file itself); db_name =
    // nullptr for the unnamed database.  The library walks the B+tree itself: no liblmdb, no transaction object.
    static Reader open_path(const char* path, const char* db_name, uint16_t index, int device = 0) {
        Reader r;
        check(hb_index_open_lmdb(path, db_name, D::metric, index, device, &r.ix_));
        return r;
    }
    // HannoyBuilder::build on the device (hb_index_build_graph) over the Item pairs of a database (metadata optional: pass
    // opts.dimensions for a database that was never built), then Reader::open on the result.  export_kv() hands the
    // Metadata / Links pairs back in the reference encoding for `database.put`.
    template <class KvCursor>
    static Reader build(const KvCursor& kv, uint16_t index, const hb_build_opts& opts, int device = 0) {
        Reader r;
        check(hb_index_begin(D::metric, index, &r.ix_));
        for (const auto& [k, v] : kv)
            check(hb_index_push_kv(r.ix_, (const uint8_t*)k.data(), k.size(), (const uint8_t*)v.data(), v.size()));
        check(hb_index_build_graph(r.ix_, &opts, device, nullptr));
        check(hb_index_finalize(r.ix_, device));
        return r;
    }
    std::vector<std::pair<std::string, std::string>> export_kv(bool with_items) const {
        std::vector<std::pair<std::string, std::string>> out;
        auto cb = [](void* u, const uint8_t* k, size_t kl, const uint8_t* v, size_t vl) -> int {
            ((std::vector<std::pair<std::string, std::string>>*)u)->emplace_back(std::string((const char*)k, kl), std::string((const char*)v, vl));
            return 0;
        };
        check(hb_index_export_kv(ix_, with_items ? 1 : 0, cb, &out));
        return out;
    }
    Reader(Reader&& o) noexcept : ix_(o.ix_) { o.ix_ = nullptr; }
    Reader& operator=(Reader&& o) noexcept { std::swap(ix_, o.ix_); return *this; }
    Reader(const Reader&) = delete;
    ~Reader() { hb_index_free(ix_); }

    size_t dimensions() const { return hb_index_dimensions(ix_); }
    uint64_t n_items() const { return hb_index_n_items(ix_); }
    bool is_empty() const { return n_items() == 0; }
    bool contains_item(ItemId item) const { return hb_index_contains_item(ix_, item) != 0; }
    std::vector<ItemId> item_ids() const {
        std::vector<ItemId> ids(n_items());
        hb_index_item_ids(ix_, ids.data(), ids.size());
        return ids;
    }
    std::optional<std::vector<float>> item_vector(ItemId item) const {
        std::vector<float> v(dimensions());
        if (hb_index_item_vector(ix_, item, v.data()) != HB_OK) return std::nullopt;
        return v;
    }
    // Replicas on further GPUs (hb_index_replicate): from then on ONE by_vector / by_item batch is partitioned into contiguous
    // slices, one host thread + stream per device — the reference's rayon workers over one Reader (src/parallel.rs:18-38).
    void replicate(const std::vector<int>& devices) { check(hb_index_replicate(ix_, devices.data(), (int)devices.size())); }
    int n_devices() const { return hb_index_n_devices(ix_); }
    // The graph as CSR over item ids (Reader::get_links for every item of a layer at once, reader.rs:966-976).
    uint32_t n_layers() const { return hb_index_n_layers(ix_); }
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> layer_csr(uint32_t layer) const {
        uint64_t nnz = 0;
        std::vector<uint64_t> off(n_items() + 1);
        check(hb_index_layer_csr(ix_, layer, off.data(), nullptr, 0, &nnz));
        std::vector<uint32_t> nbr(nnz);
        check(hb_index_layer_csr(ix_, layer, off.data(), nbr.data(), nnz, &nnz));
        return {std::move(off), std::move(nbr)};
    }
    QueryBuilder<D> nns(size_t count) const { return QueryBuilder<D>(this, count); }
    const hb_index* raw() const { return ix_; }

  private:
    Reader() = default;
    hb_index* ix_ = nullptr;
};

}  // namespace hannoy
