/* hannoy_b200.h — C-ABI of libhannoy_b200.so, the B200-native batched HNSW search engine that
 * sits behind hannoy's Reader / QueryBuilder / Distance surface.
 *
 * hannoy (nnethercott/hannoy v0.1.3) has no FFI of its own: its boundary is the public Rust API
 * (src/lib.rs:99-125).  Every entry point below names the reference interface it replaces;
 * `file:line` is relative to the reference repository.  The Rust-side binding a maintainer would
 * add is shown in INTEGRATION.md (crate source under rust/hannoy-b200/).
 *
 * Conventions
 *  - plain pointers and sizes; no C++ / torch types; no exception crosses the boundary.
 *  - every call returns hb_status; hb_last_error() gives a thread-local message for the last failure.
 *  - caller owns all host buffers and they only need to live for the duration of the call;
 *    the library owns all device memory.  Outputs are caller-allocated.
 *  - an hb_index is immutable after hb_index_finalize (mirrors `Reader: Send + Sync` over an MVCC
 *    snapshot); hb_search_* may be called concurrently from several host threads.
 *  - results carry the ORIGINAL sparse ItemIds (u32), distances are the reference's f32 values.
 *  - there is NO CPU fallback: without a CUDA device every compute call returns HB_ECUDA.
 */
#ifndef HANNOY_B200_H
#define HANNOY_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_index hb_index; /* opaque */

typedef enum {
    HB_OK = 0,
    HB_EINVAL = 1,             /* bad argument */
    HB_EDIM = 2,               /* Error::InvalidVecDimension      src/error.rs, reader.rs:133-138 */
    HB_EFORMAT = 3,            /* undecodable key/value           src/key.rs, src/node.rs:151-174 */
    HB_ECUDA = 4,              /* CUDA runtime error / no device */
    HB_ENOMEM = 5,
    HB_ENCCL = 6,
    HB_EMISSING_METADATA = 7,  /* Error::MissingMetadata          reader.rs:390-393 */
    HB_EUNMATCHING_DISTANCE = 8, /* Error::UnmatchingDistance     reader.rs:400-405 */
    HB_ENEED_BUILD = 9,        /* Error::NeedBuild                reader.rs:407-416 */
    HB_ESTATE = 10             /* call order violated (e.g. search before finalize) */
} hb_status;

/* `D: Distance` (src/distance/mod.rs:26-48), identified by D::name() */
typedef enum {
    HB_EUCLIDEAN = 0,    /* "euclidean"                  src/distance/euclidean.rs  (squared L2) */
    HB_COSINE = 1,       /* "cosine"                     src/distance/cosine.rs */
    HB_MANHATTAN = 2,    /* "manhattan"                  src/distance/manhattan.rs */
    HB_HAMMING = 3,      /* "hamming"                    src/distance/hamming.rs (codec Binary) */
    HB_BQ_COSINE = 4,    /* "binary quantized cosine"    src/distance/binary_quantized_cosine.rs */
    HB_BQ_EUCLIDEAN = 5, /* "binary quantized euclidean" src/distance/binary_quantized_euclidean.rs */
    HB_BQ_MANHATTAN = 6  /* "binary quantized manhattan" src/distance/binary_quantized_manhattan.rs */
} hb_metric;

/* D::name() for a metric, and the reverse (returns -1 if unknown). */
const char* hb_metric_name(hb_metric m);
int hb_metric_from_name(const char* name);

/* ---- Reader::open (src/reader.rs:387-431): snapshot an index out of the KV store ------------- */

/* Start a snapshot of hannoy index `index` (the u16 key prefix, src/key.rs:19-23) for distance `m`. */
hb_status hb_index_begin(hb_metric m, uint16_t index, hb_index** out);

/* Route (a): feed raw LMDB pairs exactly as a heed cursor yields them (any order; pairs of other
 * indexes are ignored).  Decodes KeyCodec (src/key.rs:54-82), NodeCodec Item/Links
 * (src/node.rs:130-174), MetadataCodec (src/metadata.rs:49-73), VersionCodec (src/version.rs:48-59),
 * Updated stones (src/update_status.rs) and the Roaring portable format (src/roaring.rs:14-32). */
hb_status hb_index_push_kv(hb_index*, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen);

/* Route (a'): read the LMDB environment on disk directly — what `Reader::open(&rtxn, index, database)` does through
 * heed (src/reader.rs:387-431; environment opened as in src/tests/mod.rs:108-111, benches/speed.rs:39-53), without
 * liblmdb or Rust.  `path` = the environment directory (holding data.mdb) or the data file itself (MDB_NOSUBDIR);
 * `db_name` = NULL/"" for the unnamed database (`env.create_database(&mut wtxn, None)`) or the name of a named one.
 * The B+tree is walked once in key order restricted to the index's 2-byte big-endian key prefix (src/key.rs:57-60,
 * the range `Prefix::all(index)` iterates, src/key.rs:94-127); every pair goes through the same decoder as
 * hb_index_push_kv.  The newest committed transaction is read; if a writer commits while the file is being walked the
 * call returns HB_ESTATE (hb_index_open_lmdb retries with a fresh snapshot).  n_pairs_out (optional) = pairs read. */
hb_status hb_index_push_lmdb(hb_index*, const char* path, const char* db_name, uint64_t* n_pairs_out);
/* hb_index_begin + hb_index_push_lmdb + hb_index_finalize: Reader::open from a path. */
hb_status hb_index_open_lmdb(const char* path, const char* db_name, hb_metric m, uint16_t index, int device, hb_index** out);
/* The walker on its own: visits every (key, value) of the database whose key starts with `prefix` (prefix_len may be
 * 0), in key order; the pointers are only valid during the callback; a non-zero return stops the scan (HB_ESTATE).
 * txnid_out (optional) = id of the LMDB transaction that was read. */
typedef int (*hb_kv_visit)(void* user, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen);
hb_status hb_lmdb_scan(const char* path, const char* db_name, const uint8_t* prefix, size_t prefix_len, hb_kv_visit fn,
                       void* user, uint64_t* txnid_out);

/* Route (c): the flat-file snapshot cache.  hb_index_save writes the decoded snapshot of a finalized index (ids, rows,
 * header norms, per-layer CSR, entry points, the index Version of src/version.rs) to `path` (atomically: tmp + rename,
 * FNV-1a trailer); hb_index_load fills an index created by hb_index_begin from such a file instead of push_kv /
 * push_lmdb — HB_EUNMATCHING_DISTANCE if the file was written for another distance, HB_EFORMAT if it is corrupt —
 * and hb_index_finalize uploads it.  The file is a cache: the caller keys it (environment, index, the LMDB
 * transaction id reported by hb_lmdb_scan) and drops it when the index is rebuilt.  No reference counterpart. */
hb_status hb_index_save(const hb_index*, const char* path);
hb_status hb_index_load(hb_index*, const char* path);

/* ---- HannoyBuilder::build on the device (writer.rs:521-603 -> hnsw.rs:122-216; the GPU build the reference lists
 * as missing, README.md:23-24) ---------------------------------------------------------------------------------
 * For an index that holds items but no graph yet (hb_index_from_arrays with n_layers = 0, or the Item / Metadata pairs
 * of a database whose `Writer` has not built): samples a level per item from the reference's distribution
 * (hnsw.rs:94-120), inserts the items level group by level group in batches — candidate search (`walk_layer`),
 * `robust_prune` and `add_link` in both directions run as kernels — and leaves per-layer links, entry points and
 * max_level in the index.  hb_index_finalize then uploads it for searching; hb_index_export_kv hands it back in the
 * reference's on-disk encoding.  Like the reference's rayon build the result depends on insertion interleaving: a
 * valid hannoy graph of the same quality, not a bit-copy of a CPU build.  M <= 32, M0 <= 64 (the reference's bindings
 * instantiate (16,32), (24,48) and (32,64), src/python.rs:280).
 * stats_out (optional): u64[8] = batches, kernel launches, items, max_level, reverse links dropped because more than 32
 * arrived for one node in one batch, candidate walks cut short, milliseconds in the batch loop (kernels), milliseconds in
 * the whole call (row upload, kernels, download and CSR). */
typedef struct {
    uint32_t M, M0;            /* HannoyBuilder::<M, M0> const generics, 16 / 32 in the reference's docs and benches */
    uint32_t ef_construction;  /* HannoyBuilder::ef_construction, default 100 (writer.rs) */
    float alpha;               /* HannoyBuilder::alpha, default 1.0 */
    uint64_t seed;             /* the `rng` argument of build() */
    uint32_t batch_max;        /* most items in flight at once (0 = 4096); never more than 1/64 of the items already linked */
    uint32_t dimensions;       /* only for a database that was never built (no metadata pair yet): Writer::new(.., dimensions);
                                  0 = take them from the metadata.  With metadata present the items are metadata.items. */
} hb_build_opts;
hb_status hb_index_build_graph(hb_index*, const hb_build_opts* opts /* NULL = defaults */, int device, uint64_t* stats_out);
/* Every pair of the index in LMDB key order, in the encodings `Writer::build` writes (Metadata, Version, one Links node
 * per (item, layer), and the Item nodes if with_items): `database.put(wtxn, key, value)` them and the CPU `Reader`
 * opens the graph.  The pointers are valid during the callback only; a non-zero return stops the export. */
hb_status hb_index_export_kv(const hb_index*, int with_items, hb_kv_visit fn, void* user);

/* Route (b): flat arrays (bench / tests).  ids ascending & unique; rows = n x dims f32 (float
 * metrics) or n x ceil(dims/64) u64 code words (binary metrics); hdr = n header norms (Cosine,
 * BQ-Cosine) or NULL; per layer l: offsets[l] has n+1 u64 entries, nbrs[l] holds neighbour ITEM IDS,
 * ascending inside each list; entry_points are item ids in metadata order. */
hb_status hb_index_from_arrays(hb_index*, uint32_t dims, const uint32_t* ids, uint64_t n, const void* rows,
                               const float* hdr, uint32_t n_layers, const uint64_t* const* offsets,
                               const uint32_t* const* nbrs, const uint32_t* entry_points, uint32_t n_ep,
                               uint32_t max_level);

/* Runs the Reader::open checks (MissingMetadata, UnmatchingDistance, NeedBuild), flattens the
 * roaring edge lists into per-layer CSR over dense ranks, repacks rows into the 16-byte aligned
 * device layout and uploads everything to `device`.  After this the index is immutable. */
hb_status hb_index_finalize(hb_index*, int device);

/* Replicas on further GPUs (the reference's `Reader` is `Send + Sync` and is shared by the threads of a
 * rayon pool, src/parallel.rs:18-38 — here the "threads" are devices).  hb_index_replicate copies the finalized device
 * snapshot to each of `devices` (device-to-device); hb_index_finalize_replicated = hb_index_finalize(devices[0]) +
 * hb_index_replicate(devices + 1, n_dev - 1).  From then on ONE hb_search_by_vector / hb_search_by_item call partitions
 * its batch into contiguous nq / n_devices slices, one host thread + stream per device, no collective; results are
 * identical to the single-device call.  The *_device entry points keep running on device 0 of the index. */
hb_status hb_index_replicate(hb_index*, const int* devices, int n_dev);
hb_status hb_index_finalize_replicated(hb_index*, const int* devices, int n_dev);
int hb_index_n_devices(const hb_index*);
int hb_index_device(const hb_index*, int i);   /* CUDA ordinal of the i-th copy, -1 if out of range */

void hb_index_free(hb_index*);

/* Reader accessors (src/reader.rs:545-573) */
uint32_t hb_index_dimensions(const hb_index*);
uint64_t hb_index_n_items(const hb_index*);
uint32_t hb_index_n_entry_points(const hb_index*);
uint32_t hb_index_max_level(const hb_index*);
hb_status hb_index_version(const hb_index*, uint32_t* major, uint32_t* minor, uint32_t* patch);
/* Reader::item_ids: copies min(cap, n) ascending ids */
uint64_t hb_index_item_ids(const hb_index*, uint32_t* out, uint64_t cap);
/* Reader::contains_item (reader.rs:595-601) */
int hb_index_contains_item(const hb_index*, uint32_t item);
/* The graph of the snapshot as flat arrays — the inverse of hb_index_from_arrays, what iterating `Links` nodes
 * (src/node.rs:162-165, reader.rs:966-976 get_links) gives: number of layers; entry-point item ids in metadata order
 * (returns how many there are, copies min(cap, that)); one layer as CSR: offsets (n+1, may be NULL), neighbour ITEM
 * IDS ascending inside each list (room for `cap` entries, may be NULL), *nnz_out = edges of the layer.  Lets a host-side
 * reader (the oracle, a CPU hannoy) take over a graph that hb_index_build_graph built without going through
 * per-pair export callbacks. */
uint32_t hb_index_n_layers(const hb_index*);
uint32_t hb_index_entry_points(const hb_index*, uint32_t* out, uint32_t cap);
hb_status hb_index_layer_csr(const hb_index*, uint32_t layer, uint64_t* offsets, uint32_t* nbr_ids, uint64_t cap, uint64_t* nnz_out);
/* Reader::item_vector (reader.rs:581-587): decoded f32 vector truncated to `dimensions`; HB_EINVAL if absent */
hb_status hb_index_item_vector(const hb_index*, uint32_t item, float* out);

/* ---- QueryBuilder (src/reader.rs:60-261) ------------------------------------------------------ */
struct hb_cancel_token;
typedef struct {
    const uint32_t* candidates; /* QueryBuilder::candidates (reader.rs:200-203): item ids, any order; NULL = none */
    uint64_t n_candidates;
    int has_candidates;         /* distinguishes "no bitmap" from "empty bitmap" */
    uint32_t linear_below;      /* QueryBuilder::linear_below, default 1000 (reader.rs:29,234-237) */
    float linear_below_ratio;   /* QueryBuilder::linear_below_ratio, default 1.0 (reader.rs:32,252-260) */
    /* by_vector_with_cancellation / by_item_with_cancellation (reader.rs:91-188): the `cancel_fn` closure cannot cross
     * the ABI; its two practical shapes can.  `cancel` = a token another host thread trips (hb_cancel_token_cancel —
     * what `|| Instant::now() > deadline` or an abort flag amount to); the kernels poll it where the reference calls
     * cancel_fn (before every layer-0 pop, reader.rs:330; once per 32 candidates of a linear scan, reader.rs:684).
     * `cancel_after_polls` = the deterministic closure "true from its N-th call on" (0 = never), polled exactly like
     * the reference polls cancel_fn — used to test parity of the Cancelled(..) results.  Both NULL/0: by_vector. */
    const struct hb_cancel_token* cancel;
    uint64_t cancel_after_polls;
} hb_query_opts;

/* A cancellation flag resident on `device`.  cancel / reset / is_cancelled may be called from any host thread while
 * searches that carry the token are running. */
typedef struct hb_cancel_token hb_cancel_token;
hb_status hb_cancel_token_create(int device, hb_cancel_token** out);
hb_status hb_cancel_token_cancel(hb_cancel_token*);
hb_status hb_cancel_token_reset(hb_cancel_token*);
int hb_cancel_token_is_cancelled(const hb_cancel_token*);
void hb_cancel_token_free(hb_cancel_token*);

/* per-query counters written by the search kernels (u64 each) */
enum { HB_CTR_DIST_UPPER = 0, HB_CTR_DIST_L0 = 1, HB_CTR_EXP_UPPER = 2, HB_CTR_EXP_L0 = 3,
       HB_CTR_DEG_UPPER = 4, HB_CTR_DEG_L0 = 5, HB_CTR_FLAGS = 6, HB_CTR_RESERVED = 7, HB_N_CTR = 8 };
enum { HB_FLAG_FALLBACK = 1, HB_FLAG_LINEAR = 2, HB_FLAG_SLOW_PATH = 4, HB_FLAG_CANCELLED = 8 };
/* Searched::did_cancel (reader.rs:36-57): bit 31 of out_len[i] when the query was cancelled (only possible when the
 * call passed a cancel token or cancel_after_polls); the low 31 bits stay the number of results. */
#define HB_LEN_CANCELLED 0x80000000u
#define HB_LEN_NONE 0xFFFFFFFFu

/* reader.nns(count).ef_search(..).by_vector(..) for a batch of nq queries (reader.rs:132-148 ->
 * 642-665 -> 722-800).  `count` = nns(count); `ef` = the QueryBuilder.ef field as the reference
 * holds it (100 by default, max(ef,count) after ef_search()).  q = nq x dims f32, row-major, host.
 * out_ids/out_dist: nq x count; out_len: nq (number of valid results per query);
 * out_counters: nq x HB_N_CTR u64 or NULL.  opts may be NULL.  Returns HB_EDIM if dims mismatch. */
hb_status hb_search_by_vector(const hb_index*, const float* q, uint64_t nq, uint32_t dims, uint32_t count, uint32_t ef,
                              const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                              uint64_t* out_counters);

/* reader.nns(count).by_item(..) for a batch (reader.rs:81-89 -> 809-894).  out_len[i] = UINT32_MAX
 * encodes `Ok(None)` (item absent / nothing to search). */
hb_status hb_search_by_item(const hb_index*, const uint32_t* items, uint64_t nq, uint32_t count, uint32_t ef,
                            const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                            uint64_t* out_counters);

/* Same search with every buffer already resident on the index's device, enqueued on `stream`
 * (a cudaStream_t) with no host synchronisation: used to time the kernels alone.  d_out_counters may be NULL. */
hb_status hb_search_by_vector_device(const hb_index*, const float* d_q, uint64_t nq, uint32_t count, uint32_t ef,
                                     uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len,
                                     uint64_t* d_out_counters, void* stream);

/* Exact k-NN in the index metric over all items (recall ground truth; the reference's analogue is
 * brute_force_search over every id, reader.rs:668-711).  Ties broken by (distance bits, id). */
hb_status hb_exact_knn(const hb_index*, const float* q, uint64_t nq, uint32_t dims, uint32_t k, uint32_t* out_ids, float* out_dist);

/* Merge `n_parts` per-shard top-k lists (each nq x k, ids + distances, padded with id=UINT32_MAX)
 * into the global top-k by (distance bits, id) — the step after the all-gather when an index is
 * sharded by item id.  Buffers are DEVICE pointers laid out [part][nq][k]; runs on `stream`. */
hb_status hb_merge_topk_device(int device, const uint32_t* d_ids, const float* d_dist, uint32_t n_parts, uint64_t nq,
                               uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream);

/* ---- id-sharded index with the all-gather fused into the search kernel (no reference counterpart: hannoy is
 * single-node; shard s = hannoy index s, src/key.rs:19-23) -------------------------------------------------------
 * One process per GPU.  Each rank creates its exchange buffer and gets a 64-byte CUDA-IPC handle; the ranks exchange
 * the handles (any host all-gather) and connect.  hb_search_sharded_device is then a collective: every rank passes
 * ALL queries (device-resident); the search kernel's epilogue stores each query's padded per-shard top-k straight into
 * every peer's gather buffer over NVLink, a flag exchange orders it, and the merge kernel leaves the global top-k by
 * (distance bits, id) in d_out_* on every rank.  Capacity: nq <= nq_cap, count <= k_cap, n_shards <= 16. */
typedef struct hb_shard_group hb_shard_group;
hb_status hb_shard_group_create(int device, int n_shards, int rank, uint64_t nq_cap, uint32_t k_cap, hb_shard_group** out,
                                uint8_t handle_out[64]);
hb_status hb_shard_group_connect(hb_shard_group*, const uint8_t* handles /* n_shards x 64 bytes, rank order */);
hb_status hb_search_sharded_device(const hb_index*, hb_shard_group*, const float* d_q, uint64_t nq, uint32_t count, uint32_t ef,
                                   uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream);
void hb_shard_group_free(hb_shard_group*);

/* number of kernel launches issued by this library in this process (for bench gpu_launches) */
uint64_t hb_launch_count(void);

const char* hb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
