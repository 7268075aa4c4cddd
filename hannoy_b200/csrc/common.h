// common.h — structures shared by the host side (snapshot.cpp, capi.cu) and the kernels.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/hannoy_b200.h"

namespace hb {

// How a row is laid out in HBM and which distance routine walks it.
enum RowKind : int {
    // f32, dims >= 32, Euclidean / Cosine.  Reproduces the AVX+FMA summation order of
    // src/spaces/simple_avx.rs:17-110 with one warp per row: the row is stored in 128-float chunks,
    // permuted so that lane j's float4 of chunk c holds elements {128c + 32t + j | t = 0..3};
    // the n % 32 tail follows in natural order.
    KIND_F32_WARP = 0,
    // f32 in natural order, one lane per row, sequential arithmetic: Euclidean / Cosine with
    // dims < 32 (SSE path simple_sse.rs:18-110 for 16..31, scalar simple.rs:49-51,81-83 below) and
    // Manhattan at any dims (manhattan.rs:41-43 is a strictly sequential sum).
    KIND_F32_LANE = 1,
    // packed u64 words (Binary / BinaryQuantized codecs), popcount distances.
    KIND_BIN = 2,
};

constexpr int MAX_LEVELS = 64;  // max_level is a u8 in the metadata; P(level >= 16) < 1e-19 for M >= 4

// Device view of one snapshotted index.  All pointers are device pointers.
struct DevIndex {
    uint32_t n = 0;        // number of items; slots are dense ranks 0..n-1 in ascending ItemId order
    uint32_t dims = 0;
    int metric = 0;
    int kind = 0;
    uint32_t row_stride = 0;  // bytes between rows (multiple of 16)
    uint32_t n_chunks = 0;    // KIND_F32_WARP: 128-float chunks in the main part
    uint32_t tail = 0;        // KIND_F32_WARP: dims % 32
    uint32_t tail_off = 0;    // KIND_F32_WARP: float offset of the tail inside a row
    uint32_t n_words = 0;     // KIND_BIN: u64 words per row = ceil(dims/64)
    const uint8_t* rows = nullptr;
    const float* hdr = nullptr;      // per-slot norm (Cosine, BQ-Cosine) or nullptr
    int hdr_uniform = 0;             // every item carries the same header (always so for BQ-Cosine: sqrt(bq_dot(v, v)) = sqrt(padded
    float hdr_value = 0.0f;          // length), binary_quantized_cosine.rs:36-38) -> the walk uses hdr_value instead of loading hdr[s]
    const uint32_t* ids = nullptr;   // slot -> ItemId
    uint32_t n_layers = 0;
    uint32_t max_level = 0;
    const uint32_t* off[MAX_LEVELS] = {};  // per layer: n+1 CSR offsets (u32; nnz < 2^32 checked at finalize)
    const uint32_t* nbr[MAX_LEVELS] = {};  // per layer: neighbour SLOTS, ascending in each list
    const uint32_t* eps = nullptr;         // entry point slots, metadata order
    uint32_t n_ep = 0;
    // layer 0 again, fixed stride: nbr0_stride (32 or 64) slots per item, ascending, padded with UINT32_MAX — one
    // aligned 128-byte line per expansion (two for graphs with 32 < M0 <= 64, the second only read when the first is
    // full) instead of two dependent CSR reads.  nullptr if some layer-0 list is longer than FIXED_DEG_MAX.
    const uint32_t* nbr0x = nullptr;
    uint32_t nbr0_stride = 32;
};
constexpr uint32_t FIXED_DEG = 32;       // one adjacency line
constexpr uint32_t FIXED_DEG_MAX = 64;   // widest fixed-stride layer-0 list (M0 = 48 / 64 of the reference's bindings, python.rs:280)
constexpr int HB_MAX_SHARDS = 16;

// One search call.
struct SearchParams {
    DevIndex ix;
    const float* q = nullptr;          // nq x dims raw f32 queries (by_vector)
    const uint32_t* q_slots = nullptr; // by_item: slot per query, UINT32_MAX = absent
    uint64_t nq = 0;
    uint32_t count = 0, ef_raw = 0;
    int mode = 0;  // 0 by_vector (hnsw), 1 by_item (hnsw), 2 by_vector linear scan, 3 by_item linear scan
    const uint32_t* cand_slots = nullptr;  // linear scan: candidate slots, ascending
    uint32_t n_cand_slots = 0;
    const uint32_t* cand_bits = nullptr;  // dense candidates bitset over slots, or nullptr
    uint32_t* out_ids = nullptr;
    float* out_dist = nullptr;
    uint32_t* out_len = nullptr;
    uint64_t* out_ctr = nullptr;  // nq x HB_N_CTR or nullptr
    // workspace
    uint32_t* visited = nullptr;       // n_slots x vis_words
    uint32_t vis_words = 0;
    uint32_t* touched = nullptr;       // n_slots x touched_cap
    uint32_t touched_cap = 0;
    unsigned long long* work_counter = nullptr;  // [2]: one per pass
    uint32_t* overflow_list = nullptr; // queries that need the slow path
    uint32_t* n_overflow = nullptr;
    uint32_t res_cap = 0, q_cap = 0;   // heap capacities (entries) for this pass
    unsigned long long* gheap = nullptr;  // slow pass: n_slots x (res_cap + q_cap) u64 in global memory
    int pass = 0;                      // 0 = shared-memory heaps, 1 = global-memory heaps over overflow_list
    uint32_t n_work = 0;               // pass 0: nq ; pass 1: read from *n_overflow on device
    uint32_t q_smem_bytes = 0;         // per-warp query staging bytes
    uint32_t ring_slots = 0;           // KIND_F32_WARP: rows in flight per warp (multiple of ROW_GROUP), ring.cuh
    uint32_t ring_stride = 0;          // bytes between ring slots
    int vis_atomic = 0;                // graph builder: adjacency lists may hold duplicates, the visited test must serialise the lanes of a chunk
    int bin_wide = 0;                  // binary kernel: heaps so large (ef ~ 800) that shared memory limits the SM to four CTAs: the 128-register instantiation
    int ring_short = 0;                // short rows: the instantiation of the ring kernel compiled for one more resident CTA per SM
    // id-sharded search with the all-gather fused into the epilogue: every query's padded top-k is stored straight into
    // each peer's gather buffer [shard][nq][count] over NVLink (peer_ids[p] / peer_dist[p] are peer-mapped device pointers)
    uint32_t* peer_ids[HB_MAX_SHARDS] = {};
    float* peer_dist[HB_MAX_SHARDS] = {};
    int n_peers = 0, shard_rank = 0;
    // cancellation (reader.rs:91-188,330): `cancel_fn` is polled before every pop of a layer-0 visit and before every
    // linear-scan chunk.  cancel_flag = device word set asynchronously by hb_cancel_token_cancel; cancel_after = the
    // deterministic closure "true from its cancel_after-th call on" (0 = never; dead-entry pruning is then off so
    // that the poll count equals the reference's); linear_cancelled = the host already cut the linear scan short.
    const uint32_t* cancel_flag = nullptr;
    uint32_t cancel_after = 0;
    int linear_cancelled = 0;
    int no_trim = 0;                   // graph builder, metrics with negative distances: shared-memory heaps without dead-entry trimming,
                                       // no deferred pops (the literal reference loop); a negative distance does not end the walk
    int defer = 1;                     // layer 0, pass 0: overlap a chunk's heap update with the next pop's adjacency / visited traffic
    // f32 rows of at most 1 KB: the ring is fed four rows per instruction through a 2-D tensor map of the row array
    // (TMA tile::gather4; box = one row).  The 128-byte CUtensorMap is carried by value (it must live in param space).
    alignas(64) unsigned char rows_tmap[128] = {};
    int gather4 = 0;
    int team = 1;                      // f32 ring kernel: warps of a CTA that ran out of queries gather rows for the ones still walking
    uint32_t n_static = 0;             // queries handed out by position (warp w of CTA b starts with query w * gridDim + b), the rest by counter
};
// One batch of the device graph builder's candidate search (search.cu build_search_kernel): `walk_layer` of
// src/hnsw.rs:460-519 for every item of the batch on layer `level` with ef = efc, optionally preceded by the greedy
// ef = 1 descent of `insert` (hnsw.rs:304-309).  The graph in `sp.ix` is not modified while the kernel runs.
struct BuildSearchParams {
    SearchParams sp;                 // index view, visited workspace, heap capacities, ring geometry; sp.q_slots = the batch's item slots
    uint32_t n_items = 0;
    uint32_t level = 0, efc = 0;
    int descend = 0;                 // 1: start at sp.ix.eps and descend max_level .. level + 1; 0: start at eps_in
    const uint32_t* eps_in = nullptr;  // [n_items][eps_stride] slots, UINT32_MAX padded (the neighbours selected one layer above)
    uint32_t eps_stride = 0;
    unsigned long long* cand = nullptr;  // [n_items][efc] ascending (distance bits << 32 | slot)
    uint32_t* cand_len = nullptr;        // [n_items]
    unsigned long long* n_cut = nullptr; // walks that stopped early (heap capacity / negative distance in the pruning pass)
};

#ifndef HB_ROW_GROUP
#define HB_ROW_GROUP 4
#endif
constexpr int ROW_GROUP = HB_ROW_GROUP;  // rows reduced together by one warp

// ---- host-side snapshot -----------------------------------------------------------------------
struct HostLayer {
    std::vector<uint64_t> off;  // n+1
    std::vector<uint32_t> nbr;  // slots
};

struct Workspace {
    int n_slots = 0;
    int n_sm = 0;
    uint32_t* visited = nullptr;
    uint32_t vis_words = 0;
    uint32_t* touched = nullptr;
    uint32_t touched_cap = 0;
    unsigned long long* work_counter = nullptr;
    uint32_t* overflow_list = nullptr;
    uint64_t overflow_cap = 0;
    uint32_t* n_overflow = nullptr;
    uint64_t* gheap = nullptr;
    uint64_t gheap_entries_per_slot = 0;
    int slow_slots = 0;
    void* stream = nullptr;  // cudaStream_t owned by the workspace (host API)
    void* busy = nullptr;    // cudaEvent_t: recorded behind the last asynchronous use (device API); reusable once it has fired
    void* last_stream = nullptr;  // ... or at once by a call on the same stream (stream order)
    bool async_used = false;
    // staging buffers for the host API (grown on demand)
    void* d_q = nullptr; size_t d_q_bytes = 0;
    void* d_out = nullptr; size_t d_out_bytes = 0;
    void* d_cand = nullptr; size_t d_cand_bytes = 0;
};

}  // namespace hb

struct hb_index {
    hb_metric metric = HB_EUCLIDEAN;
    uint16_t index = 0;
    bool finalized = false;
    int device = -1;
    // --- staging from push_kv ---
    bool have_metadata = false;
    std::string meta_distance;
    uint32_t meta_dims = 0;
    std::vector<uint32_t> meta_items;
    std::vector<uint32_t> meta_eps;
    uint32_t meta_max_level = 0;
    uint32_t version[3] = {0, 0, 0};
    bool need_build = false;
    std::map<uint32_t, std::vector<uint8_t>> kv_items;                    // id -> header||vector bytes (pairs seen before the metadata)
    bool direct_rows = false;        // metadata seen first (always, in LMDB key order): items go straight into host_rows
    std::vector<uint8_t> row_seen;   // per slot
    std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> kv_links;  // (id, layer) -> ids
    // --- canonical host snapshot ---
    uint32_t dims = 0;
    std::vector<uint32_t> ids;      // ascending
    size_t host_row_bytes = 0;      // natural encoding: 4*dims or 8*ceil(dims/64)
    std::vector<uint8_t> host_rows; // natural order (for item_vector and re-layout)
    std::vector<float> host_hdr;
    std::vector<hb::HostLayer> layers;
    std::vector<uint32_t> eps;      // slots
    uint32_t max_level = 0;
    std::vector<uint32_t> node_level;  // per slot, set by the device graph builder: the item has a Links node on layers 0..node_level
    // --- device ---
    // A replica (hb_index_replicate) is an hb_index that holds device state only: `primary` owns the host snapshot
    // (ids, dims, eps, ...) that the host side of a search reads; `replicas` of the primary are searched by
    // hb_search_by_vector / hb_search_by_item on contiguous slices of the batch, one host thread + stream per device.
    const hb_index* primary = nullptr;
    std::vector<hb_index*> replicas;
    const hb_index* host() const { return primary ? primary : this; }
    hb::DevIndex dev;
    alignas(64) unsigned char rows_tmap[128] = {};   // CUtensorMap over dev.rows (one row per box) for the gather4 ring, if have_rows_tmap
    bool have_rows_tmap = false;
    std::vector<void*> dev_allocs;
    std::vector<size_t> dev_alloc_bytes;  // parallel to dev_allocs (replication copies buffer by buffer)
    std::mutex ws_mu;
    std::vector<hb::Workspace*> ws_free;
    std::vector<hb::Workspace*> ws_all;
};

namespace hb {
void set_error(const char* fmt, ...);
// snapshot.cpp
hb_status decode_kv(hb_index* ix, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen);
hb_status build_host_snapshot_from_kv(hb_index* ix);
hb_status build_host_items_for_build(hb_index* ix, uint32_t dims_opt);
bool roaring_decode(const uint8_t* p, size_t len, std::vector<uint32_t>& out);
int64_t slot_of(const hb_index* ix, uint32_t id);
hb_status snapshot_save(const hb_index* ix, const char* path);
void roaring_encode(const uint32_t* sorted_ids, size_t n, std::vector<uint8_t>& out);
typedef int (*kv_emit_fn)(void* user, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen);
hb_status export_kv(const hb_index* ix, bool with_items, kv_emit_fn fn, void* user);
// capi.cu / build.cu
hb_status setup_dev_rows(hb_index* ix, DevIndex& d, std::vector<void*>& allocs);
hb_status build_graph_on_device(hb_index* ix, uint32_t M, uint32_t M0, uint32_t efc, float alpha, uint64_t seed, uint32_t batch_max, int device,
                                uint64_t* stats);
hb_status snapshot_load(hb_index* ix, const uint8_t* data, size_t size);
// lmdb_walk.cpp: in-order walk of one database of an LMDB data file, restricted to keys starting with `prefix`
typedef hb_status (*lmdb_visit_fn)(void* user, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen, unsigned node_flags);
hb_status lmdb_scan(const char* path, const char* db_name, const uint8_t* prefix, size_t prefix_len, lmdb_visit_fn fn, void* user,
                    uint64_t* txnid_out);
// exact_tc.cu: a CUtensorMap over a row-major f32 array whose box is one whole row of `kfloats` (<= 256) floats, unswizzled
bool make_row_gather_map(void* out128, const void* base, uint64_t rows, uint64_t kfloats, uint64_t row_bytes);
// device row layout
uint32_t device_row_stride(int kind, uint32_t dims);
void layout_row(int kind, uint32_t dims, const uint8_t* natural, uint8_t* out);
int kind_for(hb_metric m, uint32_t dims);
// search.cu
constexpr int SEARCH_WARPS_PER_BLOCK = 4;
constexpr int SEARCH_MAX_SMEM = 224 * 1024;  // dynamic shared memory a search CTA may ask for: 227 KB per CTA less the kernel's static TeamShared block
hb_status launch_search(const SearchParams& fast, const SearchParams& slow, int blocks_fast, int blocks_slow, void* stream);
size_t search_smem_per_warp(const SearchParams& p);
int search_blocks_per_sm(const SearchParams& p);  // resident CTAs per SM for these parameters
hb_status launch_build_search(const BuildSearchParams& bp, int blocks, void* stream);
int build_search_blocks_per_sm(const SearchParams& p);
// runtime tunables (hb_tune): ring bytes per warp, max resident CTAs per SM, ...
int tunable(const char* key, int dflt);
// exact.cu
hb_status launch_exact_knn(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream);
hb_status launch_merge_topk(const uint32_t* d_ids, const float* d_dist, uint32_t n_parts, uint64_t nq, uint32_t k,
                            uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream);
extern std::atomic<unsigned long long> g_launches;
void read_phases(unsigned long long* out);
hb_status launch_peer_signal_wait(uint32_t* const* peer_flags, int n_peers, int shard_rank, uint32_t* my_flags, uint32_t epoch, void* stream);
uint32_t read_trace(unsigned long long* out, uint32_t cap);
}  // namespace hb
