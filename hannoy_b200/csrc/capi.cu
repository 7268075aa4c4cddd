// capi.cu — the C-ABI of libhannoy_b200.so (include/hannoy_b200.h): index lifecycle, upload, workspaces
// and the host-facing search entry points.  No CPU compute path exists: without a device every compute
// call fails with HB_ECUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.h"

namespace hb {
const char* last_error();
}
using namespace hb;

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                   \
            return HB_ECUDA;                                                             \
        }                                                                                \
    } while (0)

// ---- runtime tunables: hb_tune(key, value) wins over the environment variable HB_<KEY> ---------------------
static std::mutex g_tune_mu;
static std::map<std::string, int> g_tune;
namespace hb {
int tunable(const char* key, int dflt) {
    {
        std::lock_guard<std::mutex> g(g_tune_mu);
        auto it = g_tune.find(key);
        if (it != g_tune.end()) return it->second;
    }
    std::string env = "HB_";
    for (const char* c = key; *c; ++c) env.push_back((char)std::toupper((unsigned char)*c));
    const char* v = std::getenv(env.c_str());
    return v ? std::atoi(v) : dflt;
}
}  // namespace hb

static const char* kMetricNames[] = {"euclidean", "cosine", "manhattan", "hamming", "binary quantized cosine",
                                     "binary quantized euclidean", "binary quantized manhattan"};

// device-side re-layout of natural rows into the search layout (see RowKind in common.h)
__global__ void layout_rows_kernel(const uint8_t* __restrict__ nat, size_t nat_stride, uint8_t* __restrict__ out,
                                   uint32_t out_stride, uint64_t n, uint32_t dims, int kind) {
    uint64_t row = blockIdx.x;
    if (row >= n) return;
    const uint8_t* src = nat + row * nat_stride;
    uint8_t* dst = out + row * (size_t)out_stride;
    for (uint32_t i = threadIdx.x; i < out_stride / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dst)[i] = 0;
    __syncthreads();
    if (kind == KIND_F32_WARP) {
        uint32_t blocks = dims / 32, chunks = (blocks + 3) / 4, main = blocks * 32;
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (uint32_t e = threadIdx.x; e < dims; e += blockDim.x) {
            uint32_t v = s[e];
            if (e < main) {
                uint32_t blk = e >> 5, j = e & 31;
                d[(blk >> 2) * 128 + j * 4 + (blk & 3)] = v;
            } else {
                d[chunks * 128 + (e - main)] = v;
            }
        }
    } else {
        for (size_t i = threadIdx.x; i < nat_stride / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    }
}

extern "C" {

const char* hb_metric_name(hb_metric m) { return ((int)m >= 0 && (int)m < 7) ? kMetricNames[(int)m] : "unknown"; }
int hb_metric_from_name(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < 7; ++i)
        if (!std::strcmp(name, kMetricNames[i])) return i;
    return -1;
}
const char* hb_last_error(void) { return last_error(); }
uint64_t hb_launch_count(void) { return g_launches; }
void hb_debug_phases(uint64_t* out16) { if (out16) read_phases((unsigned long long*)out16); }
uint32_t hb_debug_trace(uint64_t* out, uint32_t cap) { return out ? read_trace((unsigned long long*)out, cap) : 0; }
hb_status hb_tune(const char* key, int value) {
    if (!key) return HB_EINVAL;
    std::lock_guard<std::mutex> g(g_tune_mu);
    g_tune[key] = value;
    return HB_OK;
}

hb_status hb_index_begin(hb_metric m, uint16_t index, hb_index** out) {
    if (!out || (int)m < 0 || (int)m > 6) { set_error("hb_index_begin: bad arguments"); return HB_EINVAL; }
    hb_index* ix = new (std::nothrow) hb_index();
    if (!ix) return HB_ENOMEM;
    ix->metric = m;
    ix->index = index;
    *out = ix;
    return HB_OK;
}

hb_status hb_index_push_kv(hb_index* ix, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen) {
    if (!ix || !key || (!val && vlen)) { set_error("hb_index_push_kv: null argument"); return HB_EINVAL; }
    if (ix->finalized) { set_error("index already finalized"); return HB_ESTATE; }
    try {
        return decode_kv(ix, key, klen, val, vlen);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

// ---- LMDB route: the environment on disk is read directly (lmdb_walk.cpp) ---------------------------------------
struct PushCtx { hb_index* ix; uint64_t n; };
static hb_status push_cb(void* u, const uint8_t* k, size_t kl, const uint8_t* v, size_t vl, unsigned) {
    PushCtx* c = (PushCtx*)u;
    ++c->n;
    return decode_kv(c->ix, k, kl, v, vl);
}
struct UserCtx { hb_kv_visit fn; void* user; };
static hb_status user_cb(void* u, const uint8_t* k, size_t kl, const uint8_t* v, size_t vl, unsigned) {
    UserCtx* c = (UserCtx*)u;
    if (c->fn(c->user, k, kl, v, vl) != 0) { set_error("hb_lmdb_scan: stopped by the visitor"); return HB_ESTATE; }
    return HB_OK;
}

hb_status hb_lmdb_scan(const char* path, const char* db_name, const uint8_t* prefix, size_t prefix_len, hb_kv_visit fn, void* user,
                       uint64_t* txnid_out) {
    if (!path || !fn || (!prefix && prefix_len)) { set_error("hb_lmdb_scan: null argument"); return HB_EINVAL; }
    try {
        UserCtx c{fn, user};
        return lmdb_scan(path, db_name, prefix, prefix_len, user_cb, &c, txnid_out);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

hb_status hb_index_push_lmdb(hb_index* ix, const char* path, const char* db_name, uint64_t* n_pairs_out) {
    if (!ix || !path) { set_error("hb_index_push_lmdb: null argument"); return HB_EINVAL; }
    if (ix->finalized) { set_error("index already finalized"); return HB_ESTATE; }
    try {
        const uint8_t prefix[2] = {(uint8_t)(ix->index >> 8), (uint8_t)(ix->index & 0xff)};  // KeyCodec: index is big-endian, src/key.rs:57-60
        PushCtx c{ix, 0};
        hb_status st = lmdb_scan(path, db_name, prefix, 2, push_cb, &c, nullptr);
        if (n_pairs_out) *n_pairs_out = c.n;
        return st;
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

hb_status hb_index_open_lmdb(const char* path, const char* db_name, hb_metric m, uint16_t index, int device, hb_index** out) {
    if (!out) { set_error("hb_index_open_lmdb: null argument"); return HB_EINVAL; }
    *out = nullptr;
    hb_status st = HB_OK;
    for (int attempt = 0; attempt < 4; ++attempt) {  // HB_ESTATE = a writer committed under the walk: take a fresh snapshot
        hb_index* ix = nullptr;
        if ((st = hb_index_begin(m, index, &ix)) != HB_OK) return st;
        st = hb_index_push_lmdb(ix, path, db_name, nullptr);
        if (st == HB_OK) st = hb_index_finalize(ix, device);
        if (st == HB_OK) { *out = ix; return HB_OK; }
        hb_index_free(ix);
        if (st != HB_ESTATE) break;
    }
    return st;
}

// ---- device graph builder (build.cu) and write-back ------------------------------------------------------------------
hb_status hb_index_build_graph(hb_index* ix, const hb_build_opts* opts, int device, uint64_t* stats_out) {
    if (!ix) { set_error("hb_index_build_graph: null argument"); return HB_EINVAL; }
    if (ix->finalized) { set_error("index already finalized"); return HB_ESTATE; }
    try {
        hb_build_opts o = {16, 32, 100, 1.0f, 42, 0, 0};
        if (opts) o = *opts;
        if (ix->ids.empty() && (ix->have_metadata || !ix->kv_items.empty())) {   // items came through push_kv / push_lmdb
            // a database that was never built has no metadata yet (the Writer writes it in build(), writer.rs:521-603): the
            // dimensions then come from the caller.  Either way the item set is the Item nodes present, not the stored bitmap.
            hb_status st = build_host_items_for_build(ix, o.dimensions);
            if (st != HB_OK) return st;
        }
        if (ix->version[0] == 0 && ix->version[1] == 0 && ix->version[2] == 0) { ix->version[1] = 1; ix->version[2] = 3; }
        return build_graph_on_device(ix, o.M, o.M0, o.ef_construction, o.alpha, o.seed, o.batch_max, device, stats_out);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

struct ExportCtx { hb_kv_visit fn; void* user; };
static int export_cb(void* u, const uint8_t* k, size_t kl, const uint8_t* v, size_t vl) {
    ExportCtx* c = (ExportCtx*)u;
    return c->fn(c->user, k, kl, v, vl);
}
hb_status hb_index_export_kv(const hb_index* ix, int with_items, hb_kv_visit fn, void* user) {
    if (!ix || !fn) { set_error("hb_index_export_kv: null argument"); return HB_EINVAL; }
    if (!ix->finalized && (ix->ids.empty() || !ix->kv_items.empty() || !ix->kv_links.empty())) {
        set_error("hb_index_export_kv: the index holds no decoded snapshot yet");
        return HB_ESTATE;
    }
    try {
        ExportCtx c{fn, user};
        hb_status st = export_kv(ix, with_items != 0, export_cb, &c);
        if (st == HB_ESTATE) set_error("hb_index_export_kv: stopped by the visitor");
        return st;
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

// ---- flat-file snapshot cache (snapshot.cpp) ----------------------------------------------------------------------
hb_status hb_index_save(const hb_index* ix, const char* path) {
    if (!ix || !path) { set_error("hb_index_save: null argument"); return HB_EINVAL; }
    if (!ix->finalized && (ix->ids.empty() || !ix->kv_items.empty() || !ix->kv_links.empty())) {
        set_error("hb_index_save: the index holds no decoded snapshot yet (finalize it first)");
        return HB_ESTATE;
    }
    try {
        return snapshot_save(ix, path);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

hb_status hb_index_load(hb_index* ix, const char* path) {
    if (!ix || !path) { set_error("hb_index_load: null argument"); return HB_EINVAL; }
    if (ix->finalized || !ix->ids.empty() || ix->have_metadata) { set_error("hb_index_load: the index is not empty"); return HB_ESTATE; }
    // the file is mapped, not read: the arrays are copied once, straight into the snapshot
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) { set_error("snapshot: cannot open %s", path); return HB_EINVAL; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { close(fd); set_error("snapshot: cannot stat %s", path); return HB_EFORMAT; }
    void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); set_error("snapshot: mmap of %s failed", path); return HB_ENOMEM; }
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    hb_status st = HB_OK;
    try {
        st = snapshot_load(ix, (const uint8_t*)m, (size_t)sb.st_size);
    } catch (const std::bad_alloc&) {
        st = HB_ENOMEM;
    }
    munmap(m, (size_t)sb.st_size);
    close(fd);
    return st;
}

hb_status hb_index_from_arrays(hb_index* ix, uint32_t dims, const uint32_t* ids, uint64_t n, const void* rows,
                               const float* hdr, uint32_t n_layers, const uint64_t* const* offsets,
                               const uint32_t* const* nbrs, const uint32_t* entry_points, uint32_t n_ep,
                               uint32_t max_level) {
    if (!ix) return HB_EINVAL;
    if (ix->finalized) { set_error("index already finalized"); return HB_ESTATE; }
    if (n >= 0xffffffffull) { set_error("too many items"); return HB_EINVAL; }
    if (n && (!ids || !rows)) { set_error("hb_index_from_arrays: null ids/rows"); return HB_EINVAL; }
    // n_layers == 0: items only, the graph is still to be built (hb_index_build_graph)
    if (n_layers > (uint32_t)MAX_LEVELS || (n && n_layers && max_level >= n_layers) || (!n_layers && (n_ep || max_level))) { set_error("bad layer count"); return HB_EINVAL; }
    try {
        ix->dims = dims;
        ix->ids.assign(ids, ids + n);
        for (uint64_t i = 1; i < n; ++i)
            if (ids[i] <= ids[i - 1]) { set_error("ids must be strictly ascending"); return HB_EINVAL; }
        bool bin = ix->metric >= HB_HAMMING;
        ix->host_row_bytes = bin ? 8 * (((size_t)dims + 63) / 64) : 4 * (size_t)dims;
        ix->host_rows.assign((const uint8_t*)rows, (const uint8_t*)rows + n * ix->host_row_bytes);
        if (hdr) ix->host_hdr.assign(hdr, hdr + n);
        else ix->host_hdr.assign(n, 0.0f);
        ix->layers.assign(n_layers, HostLayer());
        if (n_layers && (!offsets || !nbrs)) { set_error("hb_index_from_arrays: null offsets/nbrs"); return HB_EINVAL; }
        for (uint32_t l = 0; l < n_layers; ++l) {
            // the arrays are trusted by the kernels: same checks as hb_index_load (snapshot.cpp)
            HostLayer& hl = ix->layers[l];
            if (!offsets[l]) { set_error("layer %u: null offsets", l); return HB_EINVAL; }
            hl.off.assign(offsets[l], offsets[l] + n + 1);
            if (hl.off[0] != 0) { set_error("layer %u: offsets must start at 0", l); return HB_EFORMAT; }
            for (uint64_t i = 0; i < n; ++i)
                if (hl.off[i + 1] < hl.off[i]) { set_error("layer %u: offsets are not monotone at item %llu", l, (unsigned long long)i); return HB_EFORMAT; }
            uint64_t nnz = n ? hl.off[n] : 0;
            if (nnz >= 0xffffffffull) { set_error("layer %u has too many edges for 32-bit offsets", l); return HB_EFORMAT; }
            if (nnz && !nbrs[l]) { set_error("layer %u: null neighbour array", l); return HB_EINVAL; }
            hl.nbr.resize(nnz);
            for (uint64_t e = 0; e < nnz; ++e) {
                int64_t s = slot_of(ix, nbrs[l][e]);
                if (s < 0) { set_error("layer %u: neighbour id %u is not an item", l, nbrs[l][e]); return HB_EFORMAT; }
                hl.nbr[e] = (uint32_t)s;
            }
            for (uint64_t i = 0; i < n; ++i)      // roaring iteration order (reader.rs:342): strictly ascending in every list
                for (uint64_t e = hl.off[i] + 1; e < hl.off[i + 1]; ++e)
                    if (hl.nbr[e] <= hl.nbr[e - 1]) { set_error("layer %u: neighbours of item %u are not strictly ascending", l, ids[i]); return HB_EFORMAT; }
        }
        ix->eps.clear();
        for (uint32_t i = 0; i < n_ep; ++i) {
            int64_t s = slot_of(ix, entry_points[i]);
            if (s < 0) { set_error("entry point %u is not an item", entry_points[i]); return HB_EFORMAT; }
            ix->eps.push_back((uint32_t)s);
        }
        ix->max_level = max_level;
        ix->have_metadata = true;
        ix->meta_distance = hb_metric_name(ix->metric);
        ix->meta_dims = dims;
        ix->version[0] = 0; ix->version[1] = 1; ix->version[2] = 3;
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
    return HB_OK;
}

}  // extern "C"


// Geometry of the device row layout + the rows themselves (natural encoding uploaded in slabs, re-laid out on the
// device).  Shared by hb_index_finalize and the device graph builder.  The current device must be set.
hb_status hb::setup_dev_rows(hb_index* ix, DevIndex& d, std::vector<void*>& allocs) {
    size_t n = ix->ids.size();
    d.n = (uint32_t)n;
    d.dims = ix->dims;
    d.metric = (int)ix->metric;
    d.kind = kind_for(ix->metric, ix->dims);
    d.row_stride = device_row_stride(d.kind, ix->dims);
    if (d.kind == KIND_F32_WARP) {
        uint32_t blocks = ix->dims / 32;
        d.n_chunks = (blocks + 3) / 4;
        d.tail = ix->dims % 32;
        d.tail_off = d.n_chunks * 128;
    }
    d.n_words = (ix->dims + 63) / 64;
    if (n) {
        void* drows = nullptr;
        CUDA_TRY(cudaMalloc(&drows, std::max<size_t>(n * (size_t)d.row_stride, 16)));
        allocs.push_back(drows);
        const size_t slab_rows = std::max<size_t>(1, (256u << 20) / std::max<size_t>(ix->host_row_bytes, 1));
        void* stage = nullptr;
        CUDA_TRY(cudaMalloc(&stage, slab_rows * ix->host_row_bytes));
        for (size_t r0 = 0; r0 < n; r0 += slab_rows) {
            size_t nr = std::min(slab_rows, n - r0);
            if (cudaMemcpy(stage, ix->host_rows.data() + r0 * ix->host_row_bytes, nr * ix->host_row_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(stage);
                set_error("row upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                return HB_ECUDA;
            }
            layout_rows_kernel<<<(unsigned)nr, 128>>>((const uint8_t*)stage, ix->host_row_bytes,
                                                      (uint8_t*)drows + r0 * (size_t)d.row_stride, d.row_stride, nr, ix->dims, d.kind);
            ++g_launches;
            if (cudaDeviceSynchronize() != cudaSuccess) {
                cudaFree(stage);
                set_error("row layout failed: %s", cudaGetErrorString(cudaGetLastError()));
                return HB_ECUDA;
            }
        }
        cudaFree(stage);
        d.rows = (const uint8_t*)drows;
    }
    return HB_OK;
}

template <class T>
static hb_status upload(hb_index* ix, const T* host, size_t count, const T** out) {
    void* d = nullptr;
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CUDA_TRY(cudaMalloc(&d, bytes));
    ix->dev_allocs.push_back(d);
    ix->dev_alloc_bytes.push_back(bytes);
    if (count) CUDA_TRY(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)d;
    return HB_OK;
}

extern "C" {

hb_status hb_index_finalize(hb_index* ix, int device) {
    if (!ix) return HB_EINVAL;
    if (ix->finalized) { set_error("index already finalized"); return HB_ESTATE; }
    hb_status st;
    try {
        if (!ix->kv_items.empty() || !ix->kv_links.empty() || ix->ids.empty()) {
            // KV route (or an empty index): run the Reader::open checks and flatten
            if (ix->ids.empty()) {
                st = build_host_snapshot_from_kv(ix);
                if (st != HB_OK) return st;
            }
        }
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (libhannoy_b200 has no CPU path)"); return HB_ECUDA; }
    if (device < 0 || device >= ndev) { set_error("bad device %d", device); return HB_EINVAL; }
    CUDA_TRY(cudaSetDevice(device));
    ix->device = device;
    DevIndex& d = ix->dev;
    size_t n = ix->ids.size();
    if ((st = setup_dev_rows(ix, d, ix->dev_allocs)) != HB_OK) return st;
    while (ix->dev_alloc_bytes.size() < ix->dev_allocs.size()) ix->dev_alloc_bytes.push_back(std::max<size_t>(n * (size_t)d.row_stride, 16));
    d.max_level = ix->max_level;
    d.n_layers = (uint32_t)ix->layers.size();
    if (d.n_layers > (uint32_t)MAX_LEVELS) { set_error("too many layers"); return HB_EINVAL; }
    if ((st = upload(ix, ix->host_hdr.data(), n, &d.hdr)) != HB_OK) return st;
    d.hdr_uniform = 0;
    if (n && d.kind == KIND_BIN) {   // one scattered 4-byte load (a 32-byte DRAM sector) per distance saved when the headers are all alike
        uint32_t h0, hi;
        std::memcpy(&h0, &ix->host_hdr[0], 4);
        bool same = true;
        for (size_t i = 1; i < n && same; ++i) { std::memcpy(&hi, &ix->host_hdr[i], 4); same = hi == h0; }
        d.hdr_uniform = same ? 1 : 0;
        d.hdr_value = ix->host_hdr[0];
    }
    if ((st = upload(ix, ix->ids.data(), n, &d.ids)) != HB_OK) return st;
    for (uint32_t l = 0; l < d.n_layers; ++l) {
        HostLayer& hl = ix->layers[l];
        if (hl.off.size() != n + 1) hl.off.assign(n + 1, 0);
        if (hl.off[n] >= 0xffffffffull) { set_error("layer %u has too many edges for 32-bit offsets", l); return HB_EINVAL; }
        std::vector<uint32_t> off32(hl.off.begin(), hl.off.end());
        if ((st = upload(ix, off32.data(), off32.size(), &d.off[l])) != HB_OK) return st;
        if ((st = upload(ix, hl.nbr.data(), hl.nbr.size(), &d.nbr[l])) != HB_OK) return st;
    }
    if (d.n_layers && n && tunable("fixed_adjacency", 1)) {
        const HostLayer& l0 = ix->layers[0];
        uint64_t max_deg = 0;
        for (size_t i = 0; i < n; ++i) max_deg = std::max<uint64_t>(max_deg, l0.off[i + 1] - l0.off[i]);
        if (max_deg <= FIXED_DEG_MAX) {
            // one 128-byte line per item, two for graphs built with 32 < M0 <= 64 (longer lists stay on the CSR path)
            const size_t stride = max_deg <= FIXED_DEG ? FIXED_DEG : FIXED_DEG_MAX;
            std::vector<uint32_t> fx(n * stride, 0xffffffffu);
            for (size_t i = 0; i < n; ++i)
                std::copy(l0.nbr.begin() + l0.off[i], l0.nbr.begin() + l0.off[i + 1], fx.begin() + i * stride);
            if ((st = upload(ix, fx.data(), fx.size(), &d.nbr0x)) != HB_OK) return st;
            d.nbr0_stride = (uint32_t)stride;
        }
    }
    if ((st = upload(ix, ix->eps.data(), ix->eps.size(), &d.eps)) != HB_OK) return st;
    d.n_ep = (uint32_t)ix->eps.size();
    ix->have_rows_tmap = n && d.kind == KIND_F32_WARP && d.row_stride <= 1024 && d.row_stride % 128 == 0 /* tensor copies land on 128-byte boundaries */ && make_row_gather_map(ix->rows_tmap, d.rows, n, d.row_stride / 4, d.row_stride);
    ix->finalized = true;
    return HB_OK;
}

// ---- replicas on further devices (SURVEY §8b `hb_index_finalize(ix, devices, n_dev, ..)`, §8e) ----------------------------
// Every device buffer of the finalized index is copied device-to-device (NVLink when the GPUs are peers), and the pointers
// of the DevIndex are re-based onto the copies.
static hb_status replicate_one(const hb_index* ix, int device, hb_index** out) {
    hb_index* r = new (std::nothrow) hb_index();
    if (!r) return HB_ENOMEM;
    r->metric = ix->metric; r->index = ix->index; r->dims = ix->dims; r->finalized = true; r->device = device; r->primary = ix;
    std::vector<std::pair<const uint8_t*, size_t>> ranges;
    auto fail = [&](hb_status st) { hb_index_free(r); return st; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError())); return fail(HB_ECUDA); }
    for (size_t a = 0; a < ix->dev_allocs.size(); ++a) {
        void* base = ix->dev_allocs[a];
        const size_t bytes = ix->dev_alloc_bytes[a];
        void* cp = nullptr;
        if (cudaMalloc(&cp, bytes) != cudaSuccess) { set_error("device %d: allocation of %zu bytes failed", device, bytes); cudaGetLastError(); return fail(HB_ENOMEM); }
        r->dev_allocs.push_back(cp);
        if (cudaMemcpyPeer(cp, device, base, ix->device, bytes) != cudaSuccess) { set_error("cudaMemcpyPeer %d -> %d failed: %s", ix->device, device, cudaGetErrorString(cudaGetLastError())); return fail(HB_ECUDA); }
        ranges.push_back({(const uint8_t*)base, bytes});
    }
    auto rebase = [&](const void* p) -> const void* {
        if (!p) return nullptr;
        for (size_t i = 0; i < ranges.size(); ++i)
            if ((const uint8_t*)p >= ranges[i].first && (const uint8_t*)p < ranges[i].first + ranges[i].second)
                return (const uint8_t*)r->dev_allocs[i] + ((const uint8_t*)p - ranges[i].first);
        return nullptr;
    };
    r->dev = ix->dev;
    DevIndex& d = r->dev;
    d.rows = (const uint8_t*)rebase(ix->dev.rows); d.hdr = (const float*)rebase(ix->dev.hdr); d.ids = (const uint32_t*)rebase(ix->dev.ids);
    d.eps = (const uint32_t*)rebase(ix->dev.eps); d.nbr0x = (const uint32_t*)rebase(ix->dev.nbr0x);
    for (uint32_t l = 0; l < d.n_layers; ++l) { d.off[l] = (const uint32_t*)rebase(ix->dev.off[l]); d.nbr[l] = (const uint32_t*)rebase(ix->dev.nbr[l]); }
    r->have_rows_tmap = ix->have_rows_tmap && make_row_gather_map(r->rows_tmap, d.rows, d.n, d.row_stride / 4, d.row_stride);
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("replication to device %d failed: %s", device, cudaGetErrorString(cudaGetLastError())); return fail(HB_ECUDA); }
    *out = r;
    return HB_OK;
}

extern "C" {

hb_status hb_index_replicate(hb_index* ix, const int* devices, int n_dev) {
    if (!ix || (n_dev > 0 && !devices) || n_dev < 0) { set_error("hb_index_replicate: bad arguments"); return HB_EINVAL; }
    if (!ix->finalized || ix->primary) { set_error("hb_index_replicate: finalize the index first"); return HB_ESTATE; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (libhannoy_b200 has no CPU path)"); return HB_ECUDA; }
    for (int i = 0; i < n_dev; ++i) {
        // (a device may be named more than once: it then holds several copies, each with its own stream and workspaces)
        if (devices[i] < 0 || devices[i] >= ndev) { set_error("bad device %d", devices[i]); return HB_EINVAL; }
    }
    try {
        for (int i = 0; i < n_dev; ++i) {
            hb_index* r = nullptr;
            hb_status st = replicate_one(ix, devices[i], &r);
            if (st != HB_OK) return st;
            ix->replicas.push_back(r);
        }
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
    cudaSetDevice(ix->device);
    return HB_OK;
}

hb_status hb_index_finalize_replicated(hb_index* ix, const int* devices, int n_dev) {
    if (!ix || !devices || n_dev < 1) { set_error("hb_index_finalize_replicated: bad arguments"); return HB_EINVAL; }
    hb_status st = hb_index_finalize(ix, devices[0]);
    if (st != HB_OK) return st;
    return hb_index_replicate(ix, devices + 1, n_dev - 1);
}

int hb_index_n_devices(const hb_index* ix) { return ix ? 1 + (int)ix->replicas.size() : 0; }
int hb_index_device(const hb_index* ix, int i) {
    if (!ix || i < 0 || i > (int)ix->replicas.size()) return -1;
    return i == 0 ? ix->device : ix->replicas[i - 1]->device;
}

}  // extern "C"

static void free_workspace(Workspace* w) {
    if (!w) return;
    cudaFree(w->visited); cudaFree(w->touched); cudaFree(w->work_counter); cudaFree(w->overflow_list);
    cudaFree(w->n_overflow); cudaFree(w->gheap); cudaFree(w->d_q); cudaFree(w->d_out); cudaFree(w->d_cand);
    if (w->stream) cudaStreamDestroy((cudaStream_t)w->stream);
    if (w->busy) cudaEventDestroy((cudaEvent_t)w->busy);
    delete w;
}

void hb_index_free(hb_index* ix) {
    if (!ix) return;
    for (hb_index* r : ix->replicas) hb_index_free(r);
    ix->replicas.clear();
    if (ix->device >= 0) cudaSetDevice(ix->device);
    for (Workspace* w : ix->ws_all) free_workspace(w);
    for (void* p : ix->dev_allocs) cudaFree(p);
    delete ix;
}

uint32_t hb_index_dimensions(const hb_index* ix) { return ix ? ix->dims : 0; }
uint64_t hb_index_n_items(const hb_index* ix) { return ix ? ix->ids.size() : 0; }
uint32_t hb_index_n_entry_points(const hb_index* ix) { return ix ? (uint32_t)ix->eps.size() : 0; }
uint32_t hb_index_max_level(const hb_index* ix) { return ix ? ix->max_level : 0; }
hb_status hb_index_version(const hb_index* ix, uint32_t* major, uint32_t* minor, uint32_t* patch) {
    if (!ix) return HB_EINVAL;
    if (major) *major = ix->version[0];
    if (minor) *minor = ix->version[1];
    if (patch) *patch = ix->version[2];
    return HB_OK;
}
uint64_t hb_index_item_ids(const hb_index* ix, uint32_t* out, uint64_t cap) {
    if (!ix) return 0;
    uint64_t m = std::min<uint64_t>(cap, ix->ids.size());
    if (out) std::copy(ix->ids.begin(), ix->ids.begin() + m, out);
    return ix->ids.size();
}
int hb_index_contains_item(const hb_index* ix, uint32_t item) { return ix && slot_of(ix, item) >= 0; }
uint32_t hb_index_n_layers(const hb_index* ix) { return ix ? (uint32_t)ix->layers.size() : 0; }
uint32_t hb_index_entry_points(const hb_index* ix, uint32_t* out, uint32_t cap) {
    if (!ix) return 0;
    for (size_t i = 0; i < ix->eps.size() && i < cap && out; ++i) out[i] = ix->ids[ix->eps[i]];
    return (uint32_t)ix->eps.size();
}
hb_status hb_index_layer_csr(const hb_index* ix, uint32_t layer, uint64_t* offsets, uint32_t* nbr_ids, uint64_t cap, uint64_t* nnz_out) {
    if (!ix) return HB_EINVAL;
    if (layer >= ix->layers.size()) { set_error("hb_index_layer_csr: the index has %zu layers", ix->layers.size()); return HB_EINVAL; }
    const HostLayer& hl = ix->layers[layer];
    const size_t n = ix->ids.size();
    const uint64_t nnz = hl.off.size() == n + 1 ? hl.off[n] : 0;
    if (nnz_out) *nnz_out = nnz;
    if (offsets) {
        if (hl.off.size() == n + 1) std::copy(hl.off.begin(), hl.off.end(), offsets);
        else std::fill(offsets, offsets + n + 1, 0);
    }
    if (nbr_ids) {
        if (cap < nnz) { set_error("hb_index_layer_csr: room for %llu neighbours, the layer has %llu", (unsigned long long)cap, (unsigned long long)nnz); return HB_EINVAL; }
        for (uint64_t e = 0; e < nnz; ++e) nbr_ids[e] = ix->ids[hl.nbr[e]];
    }
    return HB_OK;
}
hb_status hb_index_item_vector(const hb_index* ix, uint32_t item, float* out) {
    if (!ix || !out) return HB_EINVAL;
    int64_t s = slot_of(ix, item);
    if (s < 0) { set_error("item %u not found", item); return HB_EINVAL; }
    const uint8_t* row = ix->host_rows.data() + (size_t)s * ix->host_row_bytes;
    if (ix->metric >= HB_HAMMING) {
        // Binary::to_vec gives 0.0 / 1.0, BinaryQuantized::to_vec gives -1.0 / 1.0 (unaligned_vector/binary*.rs)
        for (uint32_t e = 0; e < ix->dims; ++e) {
            uint64_t w;
            std::memcpy(&w, row + 8 * (e / 64), 8);
            bool bit = (w >> (e % 64)) & 1;
            out[e] = ix->metric == HB_HAMMING ? (bit ? 1.0f : 0.0f) : (bit ? 1.0f : -1.0f);
        }
    } else {
        std::memcpy(out, row, 4 * (size_t)ix->dims);
    }
    return HB_OK;
}

}  // extern "C"

// ---- workspaces ---------------------------------------------------------------------------------------
static hb_status make_workspace(hb_index* ix, Workspace** out) {
    Workspace* w = new Workspace();
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ix->device));
    int bps = 64 / SEARCH_WARPS_PER_BLOCK;  // one visited bitset per warp the hardware can keep resident
    w->n_sm = prop.multiProcessorCount;
    w->n_slots = prop.multiProcessorCount * bps * SEARCH_WARPS_PER_BLOCK;
    size_t n = ix->host()->ids.size();
    w->vis_words = (uint32_t)(((n + 31) / 32 + 31) / 32 * 32);
    if (w->vis_words == 0) w->vis_words = 32;
    w->touched_cap = (uint32_t)std::max(1024, tunable("touched_cap", 16384));
    CUDA_TRY(cudaMalloc(&w->visited, (size_t)w->n_slots * w->vis_words * 4));
    CUDA_TRY(cudaMemset(w->visited, 0, (size_t)w->n_slots * w->vis_words * 4));
    CUDA_TRY(cudaMalloc(&w->touched, (size_t)w->n_slots * w->touched_cap * 4));
    CUDA_TRY(cudaMalloc(&w->work_counter, 16));
    CUDA_TRY(cudaMalloc(&w->n_overflow, 16));
    CUDA_TRY(cudaMemset(w->n_overflow, 0, 16));
    w->gheap_entries_per_slot = 2 * (uint64_t)(n + ix->host()->eps.size() + 64);
    uint64_t per_slot = w->gheap_entries_per_slot * 8;
    int slow = (int)std::min<uint64_t>(64, std::max<uint64_t>(4, (512ull << 20) / per_slot));
    slow = slow / SEARCH_WARPS_PER_BLOCK * SEARCH_WARPS_PER_BLOCK;
    w->slow_slots = std::min(slow, w->n_slots);
    CUDA_TRY(cudaMalloc(&w->gheap, (size_t)w->slow_slots * per_slot));
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    w->stream = s;
    cudaEvent_t ev;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    w->busy = ev;
    *out = w;
    return HB_OK;
}

static hb_status acquire_ws(const hb_index* cix, Workspace** out, void* for_stream = nullptr) {
    hb_index* ix = const_cast<hb_index*>(cix);
    {
        // a workspace released by the device API may still be in use by a kernel on the caller's stream: take
        // one whose `busy` event has fired, else make a new one
        std::lock_guard<std::mutex> g(ix->ws_mu);
        for (size_t i = ix->ws_free.size(); i-- > 0;) {
            Workspace* w = ix->ws_free[i];
            if (!w->async_used || (for_stream && w->last_stream == for_stream) || cudaEventQuery((cudaEvent_t)w->busy) == cudaSuccess) {
                ix->ws_free.erase(ix->ws_free.begin() + i);
                *out = w;
                return HB_OK;
            }
            cudaGetLastError();  // cudaErrorNotReady is not sticky, but clear it
        }
    }
    Workspace* w = nullptr;
    hb_status st = make_workspace(ix, &w);
    if (st != HB_OK) { free_workspace(w); return st; }
    std::lock_guard<std::mutex> g(ix->ws_mu);
    ix->ws_all.push_back(w);
    *out = w;
    return HB_OK;
}
static void release_ws(const hb_index* cix, Workspace* w) {
    hb_index* ix = const_cast<hb_index*>(cix);
    std::lock_guard<std::mutex> g(ix->ws_mu);
    ix->ws_free.push_back(w);
}
static hb_status grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return HB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    CUDA_TRY(cudaMalloc(p, need));
    *have = need;
    return HB_OK;
}

// Fill the two passes' parameters and launch.  `ov_cap` entries must be available in w->overflow_list.
static hb_status run_search(const hb_index* ix, Workspace* w, SearchParams base, void* stream) {
    const DevIndex& d = ix->dev;
    uint64_t nq = base.nq;
    if (nq > 0xfffffff0ull) { set_error("batch too large"); return HB_EINVAL; }
    if (w->overflow_cap < nq) {
        if (w->overflow_list) cudaFree(w->overflow_list);
        w->overflow_list = nullptr; w->overflow_cap = 0;
        CUDA_TRY(cudaMalloc(&w->overflow_list, std::max<uint64_t>(nq, 1024) * 4));
        w->overflow_cap = std::max<uint64_t>(nq, 1024);
    }
    base.ix = d;
    base.visited = w->visited; base.vis_words = w->vis_words;
    base.touched = w->touched; base.touched_cap = w->touched_cap;
    base.work_counter = w->work_counter;
    base.overflow_list = w->overflow_list; base.n_overflow = w->n_overflow;
    base.n_work = (uint32_t)nq;
    base.q_smem_bytes = (d.row_stride + 15) & ~15u;
    base.defer = tunable("defer", 1);
    // visited set by read + conditional reduction in the binary kernel (C4s 39.3 -> 36.5 ms: a lookup of a visited point leaves its
    // sector clean); the f32 kernels keep the single atomic (one operation less on a lone walk's chain)
    base.vis_atomic = tunable("vis_atomic", d.kind == KIND_BIN ? 0 : 1);
    base.team = tunable("team", 1);
    const uint32_t ef0 = std::max(base.ef_raw, base.count);
    if (d.kind == KIND_F32_WARP && d.row_stride >= (uint32_t)tunable("ring_min_row", 0)) {
        // rows in flight per warp: as many as fit the ring budget, in whole reduction groups (512-byte rows: the
        // whole neighbour list of an expansion in one shot).  "ring_min_row" routes shorter rows to the plain-load
        // gather (search.cu KIND_F32_DIRECT) instead; measured slower on C2 (1.57M vs 1.84M QPS), so off by default.
        // Short rows (a reduction group of four fits 8 KB) take an 8 KB ring: four CTAs per SM then fit and the kernel's
        // instantiation compiled for four is used (C2, 512-byte rows: 12 KB x 3 CTAs 1.65 M QPS, 8 KB x 4 CTAs 1.81 M).
        const bool short_rows = d.row_stride * ROW_GROUP <= 8192;
        uint32_t budget = (uint32_t)std::max(0, tunable("ring_bytes", short_rows ? 8192 : 12288));
        uint32_t slots = budget / d.row_stride / ROW_GROUP * ROW_GROUP;
        slots = std::max<uint32_t>(ROW_GROUP, std::min<uint32_t>(slots, 32));
        base.ring_slots = slots;
        base.ring_stride = d.row_stride;
        base.ring_short = tunable("ring_short", short_rows ? 1 : 0);
        if (ix->have_rows_tmap && ROW_GROUP % 4 == 0 && tunable("gather4", 1)) { base.gather4 = 1; std::memcpy(base.rows_tmap, ix->rows_tmap, 128); }
        SearchParams probe = base;
        probe.pass = 1;
        if (search_smem_per_warp(probe) * SEARCH_WARPS_PER_BLOCK > (size_t)SEARCH_MAX_SMEM) {
            base.ring_slots = 0;  // rows too long to stage: direct global-memory gather
            base.ring_stride = 0;
            base.gather4 = 0;
            base.ring_short = 0;
        }
    }
    SearchParams fast = base, slow = base;
    if (base.mode >= 2) {
        fast.res_cap = (base.count + 32 + 31) & ~31u;
        fast.q_cap = 0;
    } else {
        // live queue entries are a subset of the result set (+ distance ties) once the result set is full:
        // dead ones are dropped by queue_push, so the queue needs little more room than the result set
        fast.res_cap = (std::max(ef0, d.n_ep) + 32 + 31) & ~31u;
        fast.q_cap = (std::max(ef0, d.n_ep) + 64 + 31) & ~31u;
    }
    fast.pass = 0;
    int bps_fast = search_blocks_per_sm(fast);
    if (d.kind == KIND_BIN && bps_fast > 0 && bps_fast <= 4 && tunable("bin_wide", 1)) {
        fast.bin_wide = 1;                 // shared memory, not registers, limits the occupancy: take the instantiation that spills nothing
        bps_fast = search_blocks_per_sm(fast);
    }
    if (bps_fast == 0) {
        fast.res_cap = 0; fast.q_cap = 0;  // every query takes the global-memory pass
    }
    slow.pass = 1;
    slow.gheap = (unsigned long long*)w->gheap;
    uint64_t half = w->gheap_entries_per_slot / 2;
    slow.res_cap = (uint32_t)half; slow.q_cap = (uint32_t)half;
    int bps_slow = search_blocks_per_sm(slow);
    if (bps_slow == 0) { set_error("dimension too large"); return HB_EINVAL; }
    int bps_cap = std::max(1, tunable("blocks_per_sm", 64));
    int blocks_fast = std::min(w->n_slots / SEARCH_WARPS_PER_BLOCK, w->n_sm * std::max(std::min(bps_fast, bps_cap), 1));
    int blocks_slow = w->slow_slots / SEARCH_WARPS_PER_BLOCK;
    return launch_search(fast, slow, blocks_fast, blocks_slow, stream);
}

struct CandInfo {
    std::vector<uint32_t> slots;  // candidates ∩ items, ascending slots
    std::vector<uint32_t> bits;
};

static bool should_linear_scan(size_t n_items, size_t cand_in_db, const hb_query_opts* o) {  // reader.rs:622-640
    if (n_items == 0 || !o || !o->has_candidates) return false;
    bool below_threshold = (uint64_t)cand_in_db < (uint64_t)o->linear_below;
    bool below_ratio = ((float)cand_in_db / (float)n_items) <= o->linear_below_ratio;
    return below_threshold && below_ratio;
}

struct hb_cancel_token {
    int device = 0;
    uint32_t* d_flag = nullptr;   // polled by the search kernels
    uint32_t* h_vals = nullptr;   // pinned {1, 0}: sources of the flag copies
    cudaStream_t stream = nullptr;  // non-blocking: the copy overtakes running kernels
    std::atomic<int> cancelled{0};
};

// One device: `dx` holds the device state (the primary itself or one of its replicas), its host() the snapshot.
static hb_status search_host_one(const hb_index* dx, const float* q, const uint32_t* items, uint64_t nq, uint32_t count, uint32_t ef,
                                 const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len, uint64_t* out_ctr) {
    const hb_index* ix = dx->host();
    const bool by_item = items != nullptr;
    const size_t n = ix->ids.size();
    if (nq == 0) return HB_OK;
    if (!out_len || (count && (!out_ids || !out_dist))) { set_error("null output buffer"); return HB_EINVAL; }
    auto none = [&](uint32_t v) {
        for (uint64_t i = 0; i < nq; ++i) out_len[i] = v;
        if (out_ids && count) std::memset(out_ids, 0, nq * (size_t)count * 4);
        if (out_dist && count) std::memset(out_dist, 0, nq * (size_t)count * 4);
        if (out_ctr) std::memset(out_ctr, 0, nq * HB_N_CTR * 8);
    };
    // reader.rs:654-656 / 822-824
    CandInfo ci;
    bool has_cand = opts && opts->has_candidates;
    if (has_cand) {
        std::vector<uint32_t> c(opts->candidates, opts->candidates + opts->n_candidates);
        std::sort(c.begin(), c.end());
        c.erase(std::unique(c.begin(), c.end()), c.end());
        for (uint32_t id : c) { int64_t s = slot_of(ix, id); if (s >= 0) ci.slots.push_back((uint32_t)s); }
    }
    if (n == 0 || (has_cand && ci.slots.empty())) { none(by_item ? 0xffffffffu : 0u); return HB_OK; }
    bool linear = has_cand && should_linear_scan(n, ci.slots.size(), opts);
    const hb_cancel_token* tok = opts ? opts->cancel : nullptr;
    const uint64_t cancel_after = opts ? opts->cancel_after_polls : 0;
    if (tok && tok->device != dx->device) { set_error("the cancel token lives on device %d, the index on %d", tok->device, dx->device); return HB_EINVAL; }
    int linear_cancelled = 0;
    std::vector<uint32_t> scan_slots;  // linear scan under a poll budget: cancel_fn is called once per candidate id,
    if (linear && cancel_after) {      // present or not (reader.rs:683-687) -> the scan stops before candidate number cancel_after
        std::vector<uint32_t> c(opts->candidates, opts->candidates + opts->n_candidates);
        std::sort(c.begin(), c.end());
        c.erase(std::unique(c.begin(), c.end()), c.end());
        if (c.size() >= cancel_after) { c.resize(cancel_after - 1); linear_cancelled = 1; }
        for (uint32_t id : c) { int64_t s = slot_of(ix, id); if (s >= 0) scan_slots.push_back((uint32_t)s); }
    }
    const std::vector<uint32_t>& lin_slots = (linear && cancel_after) ? scan_slots : ci.slots;
    if (count == 0 && !by_item) { none(0); return HB_OK; }

    CUDA_TRY(cudaSetDevice(dx->device));
    Workspace* w = nullptr;
    hb_status st = acquire_ws(dx, &w);
    if (st != HB_OK) return st;
    cudaStream_t stream = (cudaStream_t)w->stream;
    auto fail = [&](hb_status s) { release_ws(dx, w); return s; };

    size_t q_bytes = by_item ? nq * 4 : nq * (size_t)ix->dims * 4;
    if ((st = grow(&w->d_q, &w->d_q_bytes, q_bytes)) != HB_OK) return fail(st);
    size_t ids_b = nq * (size_t)count * 4, len_b = nq * 4, ctr_b = out_ctr ? nq * HB_N_CTR * 8 : 0;
    size_t off_dist = (ids_b + 255) & ~(size_t)255, off_len = (off_dist + ids_b + 255) & ~(size_t)255;
    size_t off_ctr = (off_len + len_b + 255) & ~(size_t)255;
    if ((st = grow(&w->d_out, &w->d_out_bytes, off_ctr + ctr_b + 256)) != HB_OK) return fail(st);
    std::vector<uint32_t> qslots;
    if (by_item) {
        qslots.resize(nq);
        for (uint64_t i = 0; i < nq; ++i) { int64_t s = slot_of(ix, items[i]); qslots[i] = s < 0 ? 0xffffffffu : (uint32_t)s; }
        if (cudaMemcpyAsync(w->d_q, qslots.data(), q_bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) { set_error("H2D failed"); return fail(HB_ECUDA); }
    } else {
        if (cudaMemcpyAsync(w->d_q, q, q_bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) { set_error("H2D failed"); return fail(HB_ECUDA); }
    }
    SearchParams p;
    if (has_cand) {
        size_t words = (n + 31) / 32;
        ci.bits.assign(words, 0);
        for (uint32_t s : ci.slots) ci.bits[s >> 5] |= 1u << (s & 31);
        size_t cb = words * 4, off_slots = (cb + 255) & ~(size_t)255;
        if ((st = grow(&w->d_cand, &w->d_cand_bytes, off_slots + lin_slots.size() * 4 + 256)) != HB_OK) return fail(st);
        cudaMemcpyAsync(w->d_cand, ci.bits.data(), cb, cudaMemcpyHostToDevice, stream);
        if (!lin_slots.empty()) cudaMemcpyAsync((uint8_t*)w->d_cand + off_slots, lin_slots.data(), lin_slots.size() * 4, cudaMemcpyHostToDevice, stream);
        p.cand_bits = (const uint32_t*)w->d_cand;
        p.cand_slots = (const uint32_t*)((uint8_t*)w->d_cand + off_slots);
        p.n_cand_slots = (uint32_t)lin_slots.size();
    }
    p.q = by_item ? nullptr : (const float*)w->d_q;
    p.q_slots = by_item ? (const uint32_t*)w->d_q : nullptr;
    p.nq = nq; p.count = count; p.ef_raw = ef;
    p.mode = (by_item ? 1 : 0) | (linear ? 2 : 0);
    p.cancel_flag = tok ? tok->d_flag : nullptr;
    p.cancel_after = (uint32_t)std::min<uint64_t>(cancel_after, 0xffffffffull);
    p.linear_cancelled = linear_cancelled;
    uint8_t* ob = (uint8_t*)w->d_out;
    p.out_ids = (uint32_t*)ob; p.out_dist = (float*)(ob + off_dist); p.out_len = (uint32_t*)(ob + off_len);
    p.out_ctr = out_ctr ? (uint64_t*)(ob + off_ctr) : nullptr;
    if ((st = run_search(dx, w, p, stream)) != HB_OK) return fail(st);
    if (ids_b) {
        cudaMemcpyAsync(out_ids, p.out_ids, ids_b, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(out_dist, p.out_dist, ids_b, cudaMemcpyDeviceToHost, stream);
    }
    cudaMemcpyAsync(out_len, p.out_len, len_b, cudaMemcpyDeviceToHost, stream);
    if (out_ctr) cudaMemcpyAsync(out_ctr, p.out_ctr, ctr_b, cudaMemcpyDeviceToHost, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("search failed: %s", cudaGetErrorString(e)); return fail(HB_ECUDA); }
    w->async_used = false;  // drained
    release_ws(dx, w);
    return HB_OK;
}

// hb_search_by_vector / hb_search_by_item.  A replicated index (hb_index_replicate; SURVEY §8e "replicas only, partition
// the query batch") splits the batch into contiguous slices, one per device, each searched by its own host thread on its
// own stream; there is no collective, the slices land in the caller's buffers directly.
static hb_status search_host(const hb_index* ix, const float* q, const uint32_t* items, uint64_t nq, uint32_t count, uint32_t ef,
                             const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len, uint64_t* out_ctr) {
    const size_t G = 1 + ix->replicas.size();
    if (G == 1 || nq < 2) return search_host_one(ix, q, items, nq, count, ef, opts, out_ids, out_dist, out_len, out_ctr);
    if (opts && opts->cancel) {  // a cancel token lives on one device: the whole batch runs there
        const hb_index* on = ix->device == opts->cancel->device ? ix : nullptr;
        for (const hb_index* r : ix->replicas) if (!on && r->device == opts->cancel->device) on = r;
        if (!on) { set_error("the cancel token lives on device %d, which holds no replica of the index", opts->cancel->device); return HB_EINVAL; }
        return search_host_one(on, q, items, nq, count, ef, opts, out_ids, out_dist, out_len, out_ctr);
    }
    std::vector<hb_status> st(G, HB_OK);
    std::vector<std::string> msg(G);
    const size_t dims = ix->dims;
    auto run = [&](size_t g) {
        const uint64_t base = nq / G, rem = nq % G;
        const uint64_t lo = g * base + std::min<uint64_t>(g, rem), cnt = base + (g < rem ? 1 : 0);
        if (cnt == 0) return;
        const hb_index* dx = g == 0 ? ix : ix->replicas[g - 1];
        try {
            st[g] = search_host_one(dx, q ? q + lo * dims : nullptr, items ? items + lo : nullptr, cnt, count, ef, opts,
                                    out_ids ? out_ids + lo * count : nullptr, out_dist ? out_dist + lo * count : nullptr, out_len + lo,
                                    out_ctr ? out_ctr + lo * HB_N_CTR : nullptr);
        } catch (const std::bad_alloc&) {
            st[g] = HB_ENOMEM;
        }
        if (st[g] != HB_OK) msg[g] = last_error();
    };
    if (!out_len) { set_error("null output buffer"); return HB_EINVAL; }
    std::vector<std::thread> th;
    for (size_t g = 1; g < G; ++g) th.emplace_back(run, g);
    run(0);
    for (auto& t : th) t.join();
    for (size_t g = 0; g < G; ++g)
        if (st[g] != HB_OK) { set_error("device %d: %s", (g == 0 ? ix : ix->replicas[g - 1])->device, msg[g].c_str()); return st[g]; }
    return HB_OK;
}

extern "C" {

hb_status hb_cancel_token_create(int device, hb_cancel_token** out) {
    if (!out) { set_error("hb_cancel_token_create: null argument"); return HB_EINVAL; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (libhannoy_b200 has no CPU path)"); return HB_ECUDA; }
    if (device < 0 || device >= ndev) { set_error("bad device %d", device); return HB_EINVAL; }
    CUDA_TRY(cudaSetDevice(device));
    hb_cancel_token* t = new (std::nothrow) hb_cancel_token();
    if (!t) return HB_ENOMEM;
    t->device = device;
    if (cudaMalloc(&t->d_flag, 4) != cudaSuccess || cudaMemset(t->d_flag, 0, 4) != cudaSuccess ||
        cudaHostAlloc(&t->h_vals, 8, cudaHostAllocDefault) != cudaSuccess ||
        cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cancel token: %s", cudaGetErrorString(cudaGetLastError()));
        hb_cancel_token_free(t);
        return HB_ECUDA;
    }
    t->h_vals[0] = 1; t->h_vals[1] = 0;
    *out = t;
    return HB_OK;
}
static hb_status token_write(hb_cancel_token* t, int v) {
    if (!t) { set_error("null cancel token"); return HB_EINVAL; }
    t->cancelled.store(v);
    CUDA_TRY(cudaSetDevice(t->device));
    CUDA_TRY(cudaMemcpyAsync(t->d_flag, &t->h_vals[v ? 0 : 1], 4, cudaMemcpyHostToDevice, t->stream));
    CUDA_TRY(cudaStreamSynchronize(t->stream));  // the copy engine runs beside the kernels: this returns in microseconds
    return HB_OK;
}
hb_status hb_cancel_token_cancel(hb_cancel_token* t) { return token_write(t, 1); }
hb_status hb_cancel_token_reset(hb_cancel_token* t) { return token_write(t, 0); }
int hb_cancel_token_is_cancelled(const hb_cancel_token* t) { return t ? t->cancelled.load() : 0; }
void hb_cancel_token_free(hb_cancel_token* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamDestroy(t->stream);
    if (t->d_flag) cudaFree(t->d_flag);
    if (t->h_vals) cudaFreeHost(t->h_vals);
    delete t;
}

hb_status hb_search_by_vector(const hb_index* ix, const float* q, uint64_t nq, uint32_t dims, uint32_t count, uint32_t ef,
                              const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                              uint64_t* out_counters) {
    if (!ix || (!q && nq)) { set_error("hb_search_by_vector: null argument"); return HB_EINVAL; }
    if (!ix->finalized) { set_error("index not finalized"); return HB_ESTATE; }
    if (dims != ix->dims) {  // Error::InvalidVecDimension — reader.rs:133-138
        set_error("Invalid vector dimensions. Got %u but expected %u", dims, ix->dims);
        return HB_EDIM;
    }
    try {
        return search_host(ix, q, nullptr, nq, count, ef, opts, out_ids, out_dist, out_len, out_counters);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

hb_status hb_search_by_item(const hb_index* ix, const uint32_t* items, uint64_t nq, uint32_t count, uint32_t ef,
                            const hb_query_opts* opts, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                            uint64_t* out_counters) {
    if (!ix || (!items && nq)) { set_error("hb_search_by_item: null argument"); return HB_EINVAL; }
    if (!ix->finalized) { set_error("index not finalized"); return HB_ESTATE; }
    try {
        return search_host(ix, nullptr, items, nq, count, ef, opts, out_ids, out_dist, out_len, out_counters);
    } catch (const std::bad_alloc&) {
        return HB_ENOMEM;
    }
}

hb_status hb_search_by_vector_device(const hb_index* ix, const float* d_q, uint64_t nq, uint32_t count, uint32_t ef,
                                     uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len,
                                     uint64_t* d_out_counters, void* stream) {
    if (!ix || !d_q || !d_out_ids || !d_out_dist || !d_out_len) { set_error("null argument"); return HB_EINVAL; }
    if (!ix->finalized) { set_error("index not finalized"); return HB_ESTATE; }
    if (ix->ids.empty() || nq == 0 || count == 0) {
        if (nq) cudaMemsetAsync(d_out_len, 0, nq * 4, (cudaStream_t)stream);
        return HB_OK;
    }
    CUDA_TRY(cudaSetDevice(ix->device));
    Workspace* w = nullptr;
    hb_status st = acquire_ws(ix, &w, stream ? stream : (void*)1);
    if (st != HB_OK) return st;
    SearchParams p;
    p.q = d_q; p.nq = nq; p.count = count; p.ef_raw = ef; p.mode = 0;
    p.out_ids = d_out_ids; p.out_dist = d_out_dist; p.out_len = d_out_len; p.out_ctr = d_out_counters;
    st = run_search(ix, w, p, stream);
    // the workspace becomes reusable once everything enqueued so far on `stream` has run
    cudaEventRecord((cudaEvent_t)w->busy, (cudaStream_t)stream);
    w->last_stream = stream ? stream : (void*)1;  // (void*)1 stands for the legacy default stream
    w->async_used = true;
    release_ws(ix, w);
    return st;
}

// ---- id-sharded search with the all-gather fused into the search epilogue (peer memory over NVLink) ----------
struct hb_shard_group {
    int device = -1, n_shards = 0, rank = 0;
    uint64_t nq_cap = 0;
    uint32_t k_cap = 0;
    void* local = nullptr;                 // this rank's exchange buffer (cudaMalloc, exported through CUDA IPC)
    size_t bytes = 0, off_dist[2] = {0, 0}, off_ids[2] = {0, 0}, off_flags = 0;
    void* peer_base[HB_MAX_SHARDS] = {};   // every shard's exchange buffer mapped here (own entry = local)
    bool opened[HB_MAX_SHARDS] = {};
    uint32_t** d_peer_flags = nullptr;     // device array of the peers' flag arrays
    uint32_t epoch = 0;
};

hb_status hb_shard_group_create(int device, int n_shards, int rank, uint64_t nq_cap, uint32_t k_cap, hb_shard_group** out,
                                uint8_t handle_out[64]) {
    if (!out || !handle_out || n_shards < 1 || n_shards > HB_MAX_SHARDS || rank < 0 || rank >= n_shards) { set_error("bad shard group arguments"); return HB_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(device));
    hb_shard_group* g = new hb_shard_group();
    g->device = device; g->n_shards = n_shards; g->rank = rank; g->nq_cap = nq_cap; g->k_cap = k_cap;
    size_t part = ((size_t)n_shards * nq_cap * k_cap * 4 + 255) & ~(size_t)255;
    g->off_ids[0] = 0; g->off_dist[0] = part; g->off_ids[1] = 2 * part; g->off_dist[1] = 3 * part;  // two epochs' buffers
    g->off_flags = 4 * part;
    g->bytes = g->off_flags + 256;
    if (cudaMalloc(&g->local, g->bytes) != cudaSuccess) { delete g; set_error("cudaMalloc of the exchange buffer failed"); return HB_ENOMEM; }
    cudaMemset(g->local, 0, g->bytes);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, g->local) != cudaSuccess) { set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError())); cudaFree(g->local); delete g; return HB_ECUDA; }
    std::memcpy(handle_out, &h, 64);
    g->peer_base[rank] = g->local;
    *out = g;
    return HB_OK;
}

// `handles`: n_shards x 64 bytes, the handle every rank got from hb_shard_group_create, in rank order (exchange them
// with any host-side all-gather).  Maps every peer's exchange buffer into this process.
hb_status hb_shard_group_connect(hb_shard_group* g, const uint8_t* handles) {
    if (!g || !handles) return HB_EINVAL;
    CUDA_TRY(cudaSetDevice(g->device));
    for (int r = 0; r < g->n_shards; ++r) {
        if (r == g->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * r, 64);
        if (cudaIpcOpenMemHandle(&g->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle (shard %d) failed: %s", r, cudaGetErrorString(cudaGetLastError()));
            return HB_ECUDA;
        }
        g->opened[r] = true;
    }
    std::vector<uint32_t*> flags(g->n_shards);
    for (int r = 0; r < g->n_shards; ++r) flags[r] = (uint32_t*)((uint8_t*)g->peer_base[r] + g->off_flags);
    CUDA_TRY(cudaMalloc(&g->d_peer_flags, sizeof(uint32_t*) * g->n_shards));
    CUDA_TRY(cudaMemcpy(g->d_peer_flags, flags.data(), sizeof(uint32_t*) * g->n_shards, cudaMemcpyHostToDevice));
    return HB_OK;
}

void hb_shard_group_free(hb_shard_group* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < g->n_shards; ++r)
        if (g->opened[r]) cudaIpcCloseMemHandle(g->peer_base[r]);
    cudaFree(g->d_peer_flags);
    cudaFree(g->local);
    delete g;
}

// Collective: every rank calls it with the same (nq, count, ef) and the same queries (all of them, device-resident).
// Searches this rank's shard; the kernel's epilogue stores each query's padded top-k into every rank's gather buffer
// over NVLink (no NCCL all-gather); a flag exchange orders it; the merge kernel then produces the global top-k in
// d_out_* on every rank.  All on `stream`, no host synchronisation.
hb_status hb_search_sharded_device(const hb_index* ix, hb_shard_group* g, const float* d_q, uint64_t nq, uint32_t count, uint32_t ef,
                                   uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream) {
    if (!ix || !g || !d_q || !d_out_ids || !d_out_dist || !d_out_len) { set_error("null argument"); return HB_EINVAL; }
    if (!ix->finalized) { set_error("index not finalized"); return HB_ESTATE; }
    if (!g->d_peer_flags) { set_error("shard group not connected"); return HB_ESTATE; }
    if (nq > g->nq_cap || count > g->k_cap || nq == 0 || count == 0) { set_error("batch exceeds the shard group's capacity"); return HB_EINVAL; }
    if (ix->ids.empty()) { set_error("empty shard"); return HB_EINVAL; }
    CUDA_TRY(cudaSetDevice(ix->device));
    Workspace* w = nullptr;
    hb_status st = acquire_ws(ix, &w, stream ? stream : (void*)1);
    if (st != HB_OK) return st;
    const uint32_t epoch = ++g->epoch;
    const int par = epoch & 1;
    SearchParams p;
    p.q = d_q; p.nq = nq; p.count = count; p.ef_raw = ef; p.mode = 0;
    p.out_ids = d_out_ids; p.out_dist = d_out_dist; p.out_len = d_out_len;   // this shard's own list also lands here (scratch)
    p.n_peers = g->n_shards; p.shard_rank = g->rank;
    for (int r = 0; r < g->n_shards; ++r) {
        p.peer_ids[r] = (uint32_t*)((uint8_t*)g->peer_base[r] + g->off_ids[par]);
        p.peer_dist[r] = (float*)((uint8_t*)g->peer_base[r] + g->off_dist[par]);
    }
    st = run_search(ix, w, p, stream);
    cudaEventRecord((cudaEvent_t)w->busy, (cudaStream_t)stream);
    w->last_stream = stream ? stream : (void*)1;
    w->async_used = true;
    release_ws(ix, w);
    if (st != HB_OK) return st;
    uint32_t* my_flags = (uint32_t*)((uint8_t*)g->local + g->off_flags);
    if ((st = launch_peer_signal_wait(g->d_peer_flags, g->n_shards, g->rank, my_flags, epoch, stream)) != HB_OK) return st;
    return launch_merge_topk((const uint32_t*)((uint8_t*)g->local + g->off_ids[par]), (const float*)((uint8_t*)g->local + g->off_dist[par]),
                             (uint32_t)g->n_shards, nq, count, d_out_ids, d_out_dist, d_out_len, stream);
}

hb_status hb_exact_knn(const hb_index* ix, const float* q, uint64_t nq, uint32_t dims, uint32_t k, uint32_t* out_ids, float* out_dist) {
    if (!ix || (!q && nq) || !out_ids || !out_dist) { set_error("null argument"); return HB_EINVAL; }
    if (!ix->finalized) { set_error("index not finalized"); return HB_ESTATE; }
    if (dims != ix->dims) { set_error("Invalid vector dimensions. Got %u but expected %u", dims, ix->dims); return HB_EDIM; }
    if (nq == 0 || k == 0) return HB_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    float* dq = nullptr; uint32_t* dids = nullptr; float* dd = nullptr;
    CUDA_TRY(cudaMalloc(&dq, nq * (size_t)dims * 4));
    CUDA_TRY(cudaMalloc(&dids, nq * (size_t)k * 4));
    CUDA_TRY(cudaMalloc(&dd, nq * (size_t)k * 4));
    hb_status st = HB_OK;
    if (cudaMemcpy(dq, q, nq * (size_t)dims * 4, cudaMemcpyHostToDevice) != cudaSuccess) st = HB_ECUDA;
    if (st == HB_OK) st = launch_exact_knn(ix->dev, dq, nq, k, dids, dd, nullptr);
    if (st == HB_OK && cudaDeviceSynchronize() != cudaSuccess) { set_error("exact_knn kernel failed: %s", cudaGetErrorString(cudaGetLastError())); st = HB_ECUDA; }
    if (st == HB_OK) {
        cudaMemcpy(out_ids, dids, nq * (size_t)k * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out_dist, dd, nq * (size_t)k * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(dq); cudaFree(dids); cudaFree(dd);
    return st;
}

hb_status hb_merge_topk_device(int device, const uint32_t* d_ids, const float* d_dist, uint32_t n_parts, uint64_t nq,
                               uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream) {
    if (!d_ids || !d_dist || !d_out_ids || !d_out_dist) { set_error("null argument"); return HB_EINVAL; }
    CUDA_TRY(cudaSetDevice(device));
    return launch_merge_topk(d_ids, d_dist, n_parts, nq, k, d_out_ids, d_out_dist, d_out_len, stream);
}

}  // extern "C"
