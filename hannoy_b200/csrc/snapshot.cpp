// snapshot.cpp — host side of Reader::open: decode the reference's LMDB key/value encoding, flatten the
// roaring edge lists into CSR over dense ranks, and lay rows out for the device.
// Product code: shares nothing with oracle/ (which has its own, independent, encoder+decoder).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.h"

namespace hb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

static inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

// Roaring "portable" serialization (RoaringFormatSpec), what roaring 0.10 writes for
// RoaringBitmap::serialize_into and accepts in deserialize_from (src/roaring.rs:20,29; src/node.rs:143,164;
// src/metadata.rs:40,58).  Containers: array (<= 4096 sorted u16), bitmap (1024 x u64), run (with cookie 12347).
bool roaring_decode(const uint8_t* p, size_t len, std::vector<uint32_t>& out) {
    const uint32_t SERIAL_COOKIE_NO_RUN = 12346, SERIAL_COOKIE = 12347, NO_OFFSET_THRESHOLD = 4;
    if (len < 4) return false;
    uint32_t cookie = le32(p);
    size_t pos = 4, n = 0;
    const uint8_t* run_bitmap = nullptr;
    if ((cookie & 0xffff) == SERIAL_COOKIE) {
        n = (size_t)(cookie >> 16) + 1;
        size_t rb = (n + 7) / 8;
        if (len < pos + rb) return false;
        run_bitmap = p + pos;
        pos += rb;
    } else if (cookie == SERIAL_COOKIE_NO_RUN) {
        if (len < 8) return false;
        n = le32(p + 4);
        pos = 8;
    } else {
        return false;
    }
    if (n > 65536 || len < pos + 4 * n) return false;
    const uint8_t* desc = p + pos;
    pos += 4 * n;
    if (!run_bitmap || n >= NO_OFFSET_THRESHOLD) {
        if (len < pos + 4 * n) return false;
        pos += 4 * n;  // offset header (containers are read sequentially)
    }
    for (size_t i = 0; i < n; ++i) {
        uint32_t hi = (uint32_t)le16(desc + 4 * i) << 16;
        uint32_t card = (uint32_t)le16(desc + 4 * i + 2) + 1;
        bool is_run = run_bitmap && ((run_bitmap[i >> 3] >> (i & 7)) & 1);
        if (is_run) {
            if (len < pos + 2) return false;
            size_t n_runs = le16(p + pos);
            pos += 2;
            if (len < pos + 4 * n_runs) return false;
            for (size_t r = 0; r < n_runs; ++r) {
                uint32_t start = le16(p + pos), span = le16(p + pos + 2);
                pos += 4;
                if (start + span > 0xffff) return false;
                for (uint32_t v = start; v <= start + span; ++v) out.push_back(hi | v);
            }
        } else if (card > 4096) {
            if (len < pos + 8192) return false;
            for (uint32_t w = 0; w < 1024; ++w) {
                uint64_t x;
                std::memcpy(&x, p + pos + 8 * w, 8);  // little-endian host
                while (x) {
                    out.push_back(hi | (w * 64 + (uint32_t)__builtin_ctzll(x)));
                    x &= x - 1;
                }
            }
            pos += 8192;
        } else {
            if (len < pos + 2 * (size_t)card) return false;
            for (uint32_t k = 0; k < card; ++k) out.push_back(hi | le16(p + pos + 2 * k));
            pos += 2 * (size_t)card;
        }
    }
    return true;
}

static size_t header_size(hb_metric m) { return m == HB_HAMMING ? 8 : 4; }  // NodeHeaderHamming{idx: usize}
static size_t natural_row_bytes(hb_metric m, uint32_t dims) { return m >= HB_HAMMING ? 8 * (((size_t)dims + 63) / 64) : 4 * (size_t)dims; }

// One raw LMDB pair.  Key = [index u16 BE][mode u8][item u32 BE][layer u8] (src/key.rs:54-82);
// modes Metadata=0 Updated=1 Links=2 Item=3 (src/node_id.rs:11-21).
hb_status decode_kv(hb_index* ix, const uint8_t* key, size_t klen, const uint8_t* val, size_t vlen) {
    if (klen != 8) { set_error("key must be 8 bytes, got %zu", klen); return HB_EFORMAT; }
    uint16_t index = (uint16_t)((key[0] << 8) | key[1]);
    if (index != ix->index) return HB_OK;  // another index living in the same database
    uint8_t mode = key[2];
    uint32_t item = be32(key + 3);
    uint8_t layer = key[7];
    switch (mode) {
        case 0: {
            if (item == 0) {  // MetadataCodec — src/metadata.rs:49-73
                if (ix->direct_rows) { set_error("metadata pushed again after items were decoded against it"); return HB_ESTATE; }
                const void* nul = std::memchr(val, 0, vlen);
                if (!nul) { set_error("metadata: distance name not NUL-terminated"); return HB_EFORMAT; }
                size_t nl = (const uint8_t*)nul - val;
                ix->meta_distance.assign((const char*)val, nl);
                const uint8_t* p = val + nl + 1;
                size_t rem = vlen - nl - 1;
                if (rem < 8) { set_error("metadata: truncated"); return HB_EFORMAT; }
                ix->meta_dims = be32(p);
                uint32_t items_size = be32(p + 4);
                p += 8; rem -= 8;
                if (rem < items_size) { set_error("metadata: truncated items bitmap"); return HB_EFORMAT; }
                ix->meta_items.clear();
                if (!roaring_decode(p, items_size, ix->meta_items)) { set_error("metadata: bad roaring bitmap"); return HB_EFORMAT; }
                p += items_size; rem -= items_size;
                ix->meta_eps.clear();
                ix->meta_max_level = 0;
                if (rem > 0) {
                    size_t ne = (rem - 1) / 4;
                    for (size_t i = 0; i < ne; ++i) ix->meta_eps.push_back(le32(p + 4 * i));  // native-endian u32
                    ix->meta_max_level = p[rem - 1];
                }
                ix->have_metadata = true;
            } else if (item == 1) {  // VersionCodec — src/version.rs:48-59
                if (vlen < 12) { set_error("version: truncated"); return HB_EFORMAT; }
                for (int i = 0; i < 3; ++i) ix->version[i] = be32(val + 4 * i);
            }
            return HB_OK;
        }
        case 1:  // an Updated stone: the index needs a build (reader.rs:407-416)
            ix->need_build = true;
            return HB_OK;
        case 2: {  // Links — src/node.rs:162-165
            if (vlen < 1 || val[0] != 1) { set_error("links node: bad tag"); return HB_EFORMAT; }
            std::vector<uint32_t> ids;
            if (!roaring_decode(val + 1, vlen - 1, ids)) { set_error("links node: bad roaring bitmap"); return HB_EFORMAT; }
            ix->kv_links[{item, layer}] = std::move(ids);
            return HB_OK;
        }
        case 3: {  // Item — src/node.rs:154-160: [0][header][vector bytes]
            // Reader::open compares the stored distance name before any item is decoded (reader.rs:400-405): with the
            // metadata already seen (always, in key order) and another distance stored, finalize reports UnmatchingDistance
            if (ix->have_metadata && ix->meta_distance != hb_metric_name(ix->metric)) return HB_OK;
            size_t hs = header_size(ix->metric);
            if (vlen < 1 + hs || val[0] != 0) { set_error("item node: bad tag or truncated"); return HB_EFORMAT; }
            size_t vb = vlen - 1 - hs;
            size_t unit = ix->metric >= HB_HAMMING ? 8 : 4;
            if (vb % unit) { set_error("item node: %zu trailing bytes", vb % unit); return HB_EFORMAT; }  // SizeMismatch
            if (ix->have_metadata) {  // the usual order: no staging copy, the row lands in its slot
                const std::vector<uint32_t>& mi = ix->meta_items;
                auto it = std::lower_bound(mi.begin(), mi.end(), item);
                if (it == mi.end() || *it != item) {
                    // not listed in the metadata: never read by the Reader (a pending `add_item`, which also leaves an Updated
                    // stone => NeedBuild); kept aside for hb_index_build_graph, whose item set is every Item node present
                    // (writer.rs:539-553: (updated | indexed) - deleted)
                    ix->kv_items[item].assign(val + 1, val + vlen);
                    return HB_OK;
                }
                size_t s = (size_t)(it - mi.begin()), rb = natural_row_bytes(ix->metric, ix->meta_dims);
                if (!ix->direct_rows) {
                    ix->host_rows.assign(mi.size() * rb, 0);
                    ix->host_hdr.assign(mi.size(), 0.0f);
                    ix->row_seen.assign(mi.size(), 0);
                    ix->direct_rows = true;
                }
                if (vb < rb) { set_error("item %u: vector shorter than the index dimensions", item); return HB_EFORMAT; }
                if (hs == 4) std::memcpy(&ix->host_hdr[s], val + 1, 4);
                std::memcpy(ix->host_rows.data() + s * rb, val + 1 + hs, rb);
                ix->row_seen[s] = 1;
                return HB_OK;
            }
            ix->kv_items[item].assign(val + 1, val + vlen);
            return HB_OK;
        }
        default:
            set_error("Could not convert %u as a `NodeMode`.", (unsigned)mode);
            return HB_EFORMAT;
    }
}

int64_t slot_of(const hb_index* ix, uint32_t id) {
    const size_t n = ix->ids.size();
    if (n && (uint64_t)ix->ids[n - 1] - ix->ids[0] == n - 1)  // strictly ascending and gap-free: the rank is an offset
        return (id >= ix->ids[0] && id <= ix->ids[n - 1]) ? (int64_t)(id - ix->ids[0]) : -1;
    auto it = std::lower_bound(ix->ids.begin(), ix->ids.end(), id);
    return (it != ix->ids.end() && *it == id) ? (int64_t)(it - ix->ids.begin()) : -1;
}

// Reader::open checks + flatten (reader.rs:387-431)
hb_status build_host_snapshot_from_kv(hb_index* ix) {
    if (!ix->have_metadata) { set_error("Metadata are missing on index %u", (unsigned)ix->index); return HB_EMISSING_METADATA; }
    if (ix->meta_distance != hb_metric_name(ix->metric)) {
        set_error("Internal error: unmatching distance: expected `%s`, received `%s`", ix->meta_distance.c_str(), hb_metric_name(ix->metric));
        return HB_EUNMATCHING_DISTANCE;
    }
    if (ix->need_build) { set_error("The index %u needs to be built before being read", (unsigned)ix->index); return HB_ENEED_BUILD; }
    ix->dims = ix->meta_dims;
    ix->ids = ix->meta_items;  // already ascending
    size_t n = ix->ids.size();
    ix->host_row_bytes = natural_row_bytes(ix->metric, ix->dims);
    size_t hs = header_size(ix->metric);
    if (!ix->direct_rows) {
        ix->host_rows.assign(n * ix->host_row_bytes, 0);
        ix->host_hdr.assign(n, 0.0f);
    }
    for (size_t s = 0; s < n; ++s) {
        if (ix->direct_rows && ix->row_seen[s]) continue;
        auto it = ix->kv_items.find(ix->ids[s]);
        if (it == ix->kv_items.end()) { set_error("item %u listed in metadata has no Item node", ix->ids[s]); return HB_EFORMAT; }
        const std::vector<uint8_t>& v = it->second;
        if (v.size() - hs < ix->host_row_bytes) { set_error("item %u: vector shorter than the index dimensions", ix->ids[s]); return HB_EFORMAT; }
        if (hs == 4) std::memcpy(&ix->host_hdr[s], v.data(), 4);
        std::memcpy(ix->host_rows.data() + s * ix->host_row_bytes, v.data() + hs, ix->host_row_bytes);
    }
    uint32_t n_layers = ix->meta_max_level + 1;
    for (auto& kv : ix->kv_links) n_layers = std::max(n_layers, kv.first.second + 1);
    ix->layers.assign(n_layers, HostLayer());
    std::vector<std::vector<std::pair<uint32_t, const std::vector<uint32_t>*>>> per(n_layers);
    for (auto& kv : ix->kv_links) {
        int64_t s = slot_of(ix, kv.first.first);
        if (s < 0) continue;  // stale links of a deleted item
        per[kv.first.second].push_back({(uint32_t)s, &kv.second});
    }
    for (uint32_t l = 0; l < n_layers; ++l) {
        HostLayer& hl = ix->layers[l];
        hl.off.assign(n + 1, 0);
        for (auto& e : per[l]) hl.off[e.first + 1] = e.second->size();
        for (size_t s = 0; s < n; ++s) hl.off[s + 1] += hl.off[s];
        hl.nbr.resize(hl.off[n]);
        for (auto& e : per[l]) {
            uint64_t o = hl.off[e.first];
            for (uint32_t id : *e.second) {
                int64_t t = slot_of(ix, id);
                if (t < 0) { set_error("link %u -> %u points to an item that is not in the index", ix->ids[e.first], id); return HB_EFORMAT; }
                hl.nbr[o++] = (uint32_t)t;
            }
        }
    }
    ix->eps.clear();
    for (uint32_t ep : ix->meta_eps) {
        int64_t s = slot_of(ix, ep);
        if (s < 0) { set_error("entry point %u is not in the index", ep); return HB_EFORMAT; }
        ix->eps.push_back((uint32_t)s);
    }
    ix->max_level = ix->meta_max_level;
    ix->kv_items.clear();
    ix->kv_links.clear();
    std::vector<uint8_t>().swap(ix->row_seen);
    return HB_OK;
}

// The item set of a (re)build — `Writer::build`, writer.rs:539-553: `(updated | indexed) - deleted`.  `add_item` writes the
// Item node and `del_item` removes it (writer.rs:483-495), so that set is exactly the Item nodes present in the database:
// the rows decoded against the stored metadata (row_seen) plus the ones kept aside in kv_items.  The stored `items`
// bitmap, links, entry points and Updated stones of a previous build are not used: the graph is built from scratch.
hb_status build_host_items_for_build(hb_index* ix, uint32_t dims_opt) {
    uint32_t dims = dims_opt;
    if (ix->have_metadata) {
        if (ix->meta_distance != hb_metric_name(ix->metric)) {
            set_error("Internal error: unmatching distance: expected `%s`, received `%s`", ix->meta_distance.c_str(), hb_metric_name(ix->metric));
            return HB_EUNMATCHING_DISTANCE;
        }
        if (dims_opt && dims_opt != ix->meta_dims) { set_error("hb_index_build_graph: dimensions %u given, the metadata says %u", dims_opt, ix->meta_dims); return HB_EDIM; }
        dims = ix->meta_dims;
    }
    if (!dims) { set_error("hb_index_build_graph: the database has no metadata, pass hb_build_opts.dimensions"); return HB_EMISSING_METADATA; }
    const size_t rb = natural_row_bytes(ix->metric, dims), hs = header_size(ix->metric);
    std::vector<std::pair<uint32_t, int64_t>> src;  // (id, slot in the direct rows or -1 = kv_items), ascending ids
    if (ix->direct_rows)
        for (size_t s = 0; s < ix->meta_items.size(); ++s)
            if (ix->row_seen[s]) src.push_back({ix->meta_items[s], (int64_t)s});
    const size_t n_direct = src.size();
    for (auto& kv : ix->kv_items) src.push_back({kv.first, -1});
    std::inplace_merge(src.begin(), src.begin() + n_direct, src.end());
    const size_t n = src.size();
    std::vector<uint8_t> rows(n * rb, 0);
    std::vector<float> hdr(n, 0.0f);
    std::vector<uint32_t> ids(n);
    for (size_t i = 0; i < n; ++i) {
        ids[i] = src[i].first;
        if (i && ids[i] == ids[i - 1]) { set_error("item %u appears twice", ids[i]); return HB_EFORMAT; }
        if (src[i].second >= 0) {
            std::memcpy(rows.data() + i * rb, ix->host_rows.data() + (size_t)src[i].second * rb, rb);
            hdr[i] = ix->host_hdr[(size_t)src[i].second];
        } else {
            const std::vector<uint8_t>& v = ix->kv_items[ids[i]];
            if (v.size() < hs + rb) { set_error("item %u: vector shorter than the index dimensions", ids[i]); return HB_EFORMAT; }
            if (hs == 4) std::memcpy(&hdr[i], v.data(), 4);
            std::memcpy(rows.data() + i * rb, v.data() + hs, rb);
        }
    }
    ix->dims = dims;
    ix->ids.swap(ids);
    ix->host_row_bytes = rb;
    ix->host_rows.swap(rows);
    ix->host_hdr.swap(hdr);
    ix->layers.clear();
    ix->eps.clear();
    ix->max_level = 0;
    ix->kv_items.clear();
    ix->kv_links.clear();
    ix->direct_rows = false;
    std::vector<uint8_t>().swap(ix->row_seen);
    ix->meta_distance = hb_metric_name(ix->metric);
    ix->meta_dims = dims;
    ix->meta_items = ix->ids;
    ix->meta_eps.clear();
    ix->meta_max_level = 0;
    ix->have_metadata = true;
    ix->need_build = false;  // building is what the caller is about to do
    return HB_OK;
}

// ---- writing a graph back in the reference's encoding --------------------------------------------------------------------
// Roaring portable format without run containers (cookie 12346): what `RoaringBitmap::serialize_into` of roaring 0.10
// emits for bitmaps that were never run-optimised (node.rs:143, metadata.rs:40): per 64K chunk an array container up to
// 4096 values, a 1024 x u64 bitmap container above; descriptive header (key, cardinality - 1), then the offset header.
void roaring_encode(const uint32_t* ids, size_t n, std::vector<uint8_t>& out) {
    auto p16 = [&](uint32_t v) { out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); };
    auto p32 = [&](uint32_t v) { p16(v & 0xffff); p16(v >> 16); };
    std::vector<std::pair<size_t, size_t>> runs;  // [begin, end) per container
    for (size_t i = 0; i < n;) {
        size_t j = i;
        while (j < n && (ids[j] >> 16) == (ids[i] >> 16)) ++j;
        runs.push_back({i, j});
        i = j;
    }
    p32(12346);
    p32((uint32_t)runs.size());
    for (auto& r : runs) { p16(ids[r.first] >> 16); p16((uint32_t)(r.second - r.first - 1)); }
    uint32_t offset = 8 + 8 * (uint32_t)runs.size();
    for (auto& r : runs) {
        p32(offset);
        size_t card = r.second - r.first;
        offset += card > 4096 ? 8192 : 2 * (uint32_t)card;
    }
    for (auto& r : runs) {
        size_t card = r.second - r.first;
        if (card > 4096) {
            size_t at = out.size();
            out.resize(at + 8192, 0);
            for (size_t k = r.first; k < r.second; ++k) { uint32_t v = ids[k] & 0xffff; out[at + (v >> 3)] |= (uint8_t)(1u << (v & 7)); }
        } else {
            for (size_t k = r.first; k < r.second; ++k) p16(ids[k] & 0xffff);
        }
    }
}

// Every pair of the index in LMDB key order (key.rs:54-66, node_id.rs:11-21: Metadata < Updated < Links < Item), in the
// encodings of metadata.rs:22-47, version.rs:33-46 and node.rs:130-149 — what `Writer::build` leaves in the database
// (writer.rs:521-603, hnsw.rs:190-212), so that the CPU `Reader` can open a graph built on the device.
hb_status export_kv(const hb_index* ix, bool with_items, kv_emit_fn fn, void* user) {
    const size_t n = ix->ids.size();
    auto key = [&](uint8_t mode, uint32_t item, uint8_t layer, uint8_t* k) {
        k[0] = (uint8_t)(ix->index >> 8); k[1] = (uint8_t)ix->index; k[2] = mode;
        k[3] = (uint8_t)(item >> 24); k[4] = (uint8_t)(item >> 16); k[5] = (uint8_t)(item >> 8); k[6] = (uint8_t)item; k[7] = layer;
    };
    auto be32p = [](std::vector<uint8_t>& v, uint32_t x) { v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); };
    uint8_t k[8];
    std::vector<uint8_t> val, tmp;
    // metadata
    const char* name = hb_metric_name(ix->metric);
    val.assign(name, name + std::strlen(name));
    val.push_back(0);
    be32p(val, ix->dims);
    tmp.clear();
    roaring_encode(ix->ids.data(), n, tmp);
    be32p(val, (uint32_t)tmp.size());
    val.insert(val.end(), tmp.begin(), tmp.end());
    for (uint32_t s : ix->eps) { uint32_t id = ix->ids[s]; const uint8_t* b = (const uint8_t*)&id; val.insert(val.end(), b, b + 4); }  // ItemIds::raw_bytes: native endian
    val.push_back((uint8_t)ix->max_level);
    key(0, 0, 0, k);
    if (fn(user, k, 8, val.data(), val.size())) return HB_ESTATE;
    // version
    val.clear();
    for (int i = 0; i < 3; ++i) be32p(val, ix->version[i]);
    key(0, 1, 0, k);
    if (fn(user, k, 8, val.data(), val.size())) return HB_ESTATE;
    // links: one node per (item, layer <= level of the item), keys ordered by item then layer.  A snapshot that did not
    // come from the builder does not record levels: the highest layer with a link, or max_level for an entry point.
    std::vector<uint32_t> lvl_of = ix->node_level;
    if (lvl_of.size() != n) {
        lvl_of.assign(n, 0);
        for (uint32_t l = 1; l < ix->layers.size(); ++l)
            for (size_t s = 0; s < n; ++s)
                if (ix->layers[l].off[s + 1] > ix->layers[l].off[s]) lvl_of[s] = l;
        for (uint32_t s : ix->eps) lvl_of[s] = std::max<uint32_t>(lvl_of[s], ix->max_level);
    }
    std::vector<uint32_t> ids;
    for (size_t s = 0; s < n; ++s) {
        for (uint32_t l = 0; l < ix->layers.size(); ++l) {
            const HostLayer& hl = ix->layers[l];
            if (lvl_of[s] < l) continue;
            ids.clear();
            for (uint64_t e = hl.off[s]; e < hl.off[s + 1]; ++e) ids.push_back(ix->ids[hl.nbr[e]]);  // slots ascending => ids ascending
            val.assign(1, 1);  // LINKS_TAG
            roaring_encode(ids.data(), ids.size(), val);
            key(2, ix->ids[s], (uint8_t)l, k);
            if (fn(user, k, 8, val.data(), val.size())) return HB_ESTATE;
        }
    }
    if (with_items) {
        const size_t hs = header_size(ix->metric);
        for (size_t s = 0; s < n; ++s) {
            val.assign(1 + hs, 0);  // NODE_TAG, header (norm; NodeHeaderHamming{idx: usize} is left 0)
            if (hs == 4) std::memcpy(val.data() + 1, &ix->host_hdr[s], 4);
            const uint8_t* row = ix->host_rows.data() + s * ix->host_row_bytes;
            val.insert(val.end(), row, row + ix->host_row_bytes);
            key(3, ix->ids[s], 0, k);
            if (fn(user, k, 8, val.data(), val.size())) return HB_ESTATE;
        }
    }
    return HB_OK;
}

// ---- flat-file snapshot cache ------------------------------------------------------------------------------------
// The decoded host snapshot (what hb_index_finalize uploads) written as one flat little-endian file, so that a restart
// does not walk LMDB and decode a roaring bitmap per node again.  The reference has no counterpart (its Reader reads
// LMDB lazily, reader.rs:951-976); the file is a cache, valid for as long as the caller's key for it (environment path,
// index, LMDB transaction id from hb_lmdb_scan, the index Version of version.rs) is unchanged.
//   header : "HB2SNAP1" | u32 metric | u32 index | u32 dims | u32 max_level | u64 n | u32 n_layers | u32 n_eps |
//            u32 version[3] | u32 row_bytes | u64 nnz[n_layers]
//   body   : ids u32[n] | hdr f32[n] | eps u32[n_eps] (slots) | per layer: off u64[n+1], nbr u32[nnz] (slots) |
//            pad to 8 | rows u8[n * row_bytes]
//   trailer: u64 FNV-1a over every preceding 8-byte word
namespace {
struct Fnv {
    uint64_t h = 0xcbf29ce484222325ull;
    uint8_t carry[8];
    size_t nc = 0;
    void word(uint64_t w) { h = (h ^ w) * 0x100000001b3ull; }
    void feed(const void* p, size_t len) {
        const uint8_t* b = (const uint8_t*)p;
        while (len && nc) { carry[nc++] = *b++; --len; if (nc == 8) { uint64_t w; std::memcpy(&w, carry, 8); word(w); nc = 0; } }
        for (; len >= 8; len -= 8, b += 8) { uint64_t w; std::memcpy(&w, b, 8); word(w); }
        while (len) { carry[nc++] = *b++; --len; }
    }
    uint64_t finish() { if (nc) { std::memset(carry + nc, 0, 8 - nc); uint64_t w; std::memcpy(&w, carry, 8); word(w); nc = 0; } return h; }
};
struct Writer {
    FILE* f; Fnv fnv; bool ok = true;
    void put(const void* p, size_t len) { if (len && ok) { ok = fwrite(p, 1, len, f) == len; fnv.feed(p, len); } }
    template <class T> void val(T v) { put(&v, sizeof(T)); }
};
struct Cursor {
    const uint8_t* p; size_t left; bool ok = true;
    const uint8_t* take(size_t len) { if (len > left) { ok = false; return nullptr; } const uint8_t* r = p; p += len; left -= len; return r; }
    template <class T> T val() { T v{}; const uint8_t* r = take(sizeof(T)); if (r) std::memcpy(&v, r, sizeof(T)); return v; }
    template <class T> bool vec(std::vector<T>& out, size_t count) {
        if (count > left / sizeof(T)) { ok = false; return false; }
        const uint8_t* r = take(count * sizeof(T));
        out.resize(count);
        if (count) std::memcpy(out.data(), r, count * sizeof(T));
        return true;
    }
};
const char kSnapMagic[8] = {'H', 'B', '2', 'S', 'N', 'A', 'P', '1'};
}  // namespace

hb_status snapshot_save(const hb_index* ix, const char* path) {
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("snapshot: cannot create %s", tmp.c_str()); return HB_EINVAL; }
    Writer w{f};
    const uint64_t n = ix->ids.size();
    w.put(kSnapMagic, 8);
    w.val<uint32_t>((uint32_t)ix->metric); w.val<uint32_t>(ix->index); w.val<uint32_t>(ix->dims); w.val<uint32_t>(ix->max_level);
    w.val<uint64_t>(n); w.val<uint32_t>((uint32_t)ix->layers.size()); w.val<uint32_t>((uint32_t)ix->eps.size());
    for (int i = 0; i < 3; ++i) w.val<uint32_t>(ix->version[i]);
    w.val<uint32_t>((uint32_t)ix->host_row_bytes);
    for (const HostLayer& hl : ix->layers) w.val<uint64_t>(hl.nbr.size());
    w.put(ix->ids.data(), n * 4);
    w.put(ix->host_hdr.data(), n * 4);
    w.put(ix->eps.data(), ix->eps.size() * 4);
    size_t body = n * 8 + ix->eps.size() * 4;
    for (const HostLayer& hl : ix->layers) {
        std::vector<uint64_t> zero;
        const std::vector<uint64_t>& off = hl.off.size() == n + 1 ? hl.off : (zero.assign(n + 1, 0), zero);
        w.put(off.data(), (n + 1) * 8);
        w.put(hl.nbr.data(), hl.nbr.size() * 4);
        body += hl.nbr.size() * 4;
    }
    const uint8_t pad[8] = {};
    w.put(pad, (8 - body % 8) % 8);
    w.put(ix->host_rows.data(), n * ix->host_row_bytes);
    uint64_t sum = w.fnv.finish();
    bool ok = w.ok && fwrite(&sum, 1, 8, f) == 8;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); set_error("snapshot: writing %s failed", path); return HB_EINVAL; }
    return HB_OK;
}

hb_status snapshot_load(hb_index* ix, const uint8_t* data, size_t size) {
    if (size < 8 + 8 || std::memcmp(data, kSnapMagic, 8)) { set_error("snapshot: not a hannoy_b200 snapshot file"); return HB_EFORMAT; }
    Fnv fnv;
    fnv.feed(data, size - 8);
    uint64_t want;
    std::memcpy(&want, data + size - 8, 8);
    if (fnv.finish() != want) { set_error("snapshot: checksum mismatch (truncated or corrupt file)"); return HB_EFORMAT; }
    Cursor c{data + 8, size - 16};
    uint32_t metric = c.val<uint32_t>(), index = c.val<uint32_t>(), dims = c.val<uint32_t>(), max_level = c.val<uint32_t>();
    uint64_t n = c.val<uint64_t>();
    uint32_t n_layers = c.val<uint32_t>(), n_eps = c.val<uint32_t>();
    uint32_t ver[3];
    for (int i = 0; i < 3; ++i) ver[i] = c.val<uint32_t>();
    uint32_t row_bytes = c.val<uint32_t>();
    if (!c.ok) { set_error("snapshot: truncated header"); return HB_EFORMAT; }
    if (metric != (uint32_t)ix->metric) {  // the same check Reader::open makes on the stored distance name (reader.rs:400-405)
        set_error("Internal error: unmatching distance: expected `%s`, received `%s`", hb_metric_name((hb_metric)metric), hb_metric_name(ix->metric));
        return HB_EUNMATCHING_DISTANCE;
    }
    if (index != ix->index) { set_error("snapshot: file holds index %u, asked for %u", index, (unsigned)ix->index); return HB_EINVAL; }
    if (n >= 0xffffffffull || n_layers > (uint32_t)MAX_LEVELS || row_bytes != natural_row_bytes(ix->metric, dims) || (n && max_level >= n_layers)) {
        set_error("snapshot: inconsistent header");
        return HB_EFORMAT;
    }
    std::vector<uint64_t> nnz(n_layers);
    for (uint32_t l = 0; l < n_layers; ++l) nnz[l] = c.val<uint64_t>();
    size_t body = n * 8 + (size_t)n_eps * 4;
    c.vec(ix->ids, n);
    c.vec(ix->host_hdr, n);
    c.vec(ix->eps, n_eps);
    ix->layers.assign(n_layers, HostLayer());
    for (uint32_t l = 0; l < n_layers && c.ok; ++l) {
        c.vec(ix->layers[l].off, n + 1);
        c.vec(ix->layers[l].nbr, nnz[l]);
        body += nnz[l] * 4;
    }
    c.take((8 - body % 8) % 8);
    if (c.ok && n * (uint64_t)row_bytes != c.left) c.ok = false;
    if (c.ok) c.vec(ix->host_rows, n * (size_t)row_bytes);
    if (!c.ok) { set_error("snapshot: sizes do not add up"); return HB_EFORMAT; }
    // the arrays are trusted by the kernels: validate what an out-of-range value would break
    for (uint64_t i = 1; i < n; ++i) if (ix->ids[i] <= ix->ids[i - 1]) { set_error("snapshot: ids not ascending"); return HB_EFORMAT; }
    for (uint32_t e : ix->eps) if (e >= n) { set_error("snapshot: entry point out of range"); return HB_EFORMAT; }
    for (uint32_t l = 0; l < n_layers; ++l) {
        const HostLayer& hl = ix->layers[l];
        if (hl.off[0] != 0 || hl.off[n] != hl.nbr.size()) { set_error("snapshot: layer %u offsets do not cover its edges", l); return HB_EFORMAT; }
        for (uint64_t i = 0; i < n; ++i) if (hl.off[i + 1] < hl.off[i]) { set_error("snapshot: layer %u offsets not monotone", l); return HB_EFORMAT; }
        for (uint32_t t : hl.nbr) if (t >= n) { set_error("snapshot: layer %u neighbour out of range", l); return HB_EFORMAT; }
    }
    ix->dims = dims;
    ix->host_row_bytes = row_bytes;
    ix->max_level = max_level;
    for (int i = 0; i < 3; ++i) ix->version[i] = ver[i];
    ix->have_metadata = true;
    ix->meta_distance = hb_metric_name(ix->metric);
    ix->meta_dims = dims;
    ix->meta_items = ix->ids;
    ix->meta_max_level = max_level;
    return HB_OK;
}

// ---- device row layout -------------------------------------------------------------------------------
int kind_for(hb_metric m, uint32_t dims) {
    if (m >= HB_HAMMING) return KIND_BIN;
    if (m == HB_MANHATTAN || dims < 32) return KIND_F32_LANE;
    return KIND_F32_WARP;
}
uint32_t device_row_stride(int kind, uint32_t dims) {
    if (kind == KIND_BIN) return 16 * ((((dims + 63) / 64) + 1) / 2);
    if (kind == KIND_F32_LANE) return 16 * ((dims + 3) / 4);
    uint32_t blocks = dims / 32, chunks = (blocks + 3) / 4, tail = dims % 32;
    return 4 * (chunks * 128 + 4 * ((tail + 3) / 4));
}
void layout_row(int kind, uint32_t dims, const uint8_t* natural, uint8_t* out) {
    uint32_t stride = device_row_stride(kind, dims);
    std::memset(out, 0, stride);
    if (kind == KIND_BIN) { std::memcpy(out, natural, 8 * (((size_t)dims + 63) / 64)); return; }
    if (kind == KIND_F32_LANE) { std::memcpy(out, natural, 4 * (size_t)dims); return; }
    const float* src = (const float*)natural;
    float* dst = (float*)out;
    uint32_t blocks = dims / 32, chunks = (blocks + 3) / 4, main = blocks * 32;
    for (uint32_t e = 0; e < main; ++e) {
        uint32_t blk = e >> 5, j = e & 31;
        float v;
        std::memcpy(&v, natural + 4 * (size_t)e, 4);
        dst[(blk >> 2) * 128 + j * 4 + (blk & 3)] = v;
    }
    (void)src;
    std::memcpy(dst + chunks * 128, natural + 4 * (size_t)main, 4 * (size_t)(dims - main));
}

}  // namespace hb
