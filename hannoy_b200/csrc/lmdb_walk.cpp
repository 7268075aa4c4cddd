// lmdb_walk.cpp — read-only walker over an LMDB data file (LMDB 0.9.x on-disk format, MDB_DATA_VERSION 1), so that
// Reader::open (reference src/reader.rs:387-431) works from an environment directory alone: no liblmdb, no heed, no
// Rust.  The reference reaches the same bytes through `heed::Database::get` / `prefix_iter` on an `RoTxn`
// (src/reader.rs:390-431,951-976; src/item_iter.rs:10-52); here the B+tree of the database is walked once, in key
// order, restricted to the 2-byte big-endian index prefix (src/key.rs:19-23,54-65), and every pair is handed to a
// callback (normally hb_index_push_kv's decoder).
//
// On-disk format restated from LMDB's public description (lmdb.h / mdb.c "MDB_page", "MDB_meta", "MDB_db",
// "MDB_node"); little-endian, 64-bit pgno:
//   page header (16 B): pgno u64 | pad u16 | flags u16 | lower u16, upper u16  (overflow pages: pages u32 instead)
//   flags: BRANCH 0x01, LEAF 0x02, OVERFLOW 0x04, META 0x08, LEAF2 0x20, SUBP 0x40
//   meta page body at +16: magic u32 (0xBEEFC0DE) | version u32 (1) | address u64 | mapsize u64 |
//        dbs[2] x MDB_db (FREE, MAIN) | last_pg u64 | txnid u64;  page size = dbs[FREE].pad
//   MDB_db (48 B): pad u32 | flags u16 | depth u16 | branch_pages u64 | leaf_pages u64 | overflow_pages u64 |
//        entries u64 | root u64 (~0 = empty)
//   node: lo u16 | hi u16 | flags u16 | ksize u16 | key | data.   branch: child pgno = lo | hi<<16 | flags<<32;
//        leaf: data size = lo | hi<<16, flags BIGDATA 0x01 (data = u64 pgno of an overflow run), SUBDATA 0x02
//        (data = MDB_db of a named database), DUPDATA 0x04.
//   node pointers: u16 offsets from the page start, (lower - 16) / 2 of them, sorted by key.
// Named databases are records of the MAIN database whose key is the name.
//
// Consistency without the reader lock table: a snapshot is taken from the meta page with the larger txnid; pages of
// that snapshot can only be recycled by a write transaction that starts after a LATER commit, so the walk is valid iff
// the newest txnid is unchanged when the walk ends — checked; otherwise HB_ESTATE (hb_index_open_lmdb then retries with a fresh snapshot).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <string>

#include "common.h"

namespace hb {

namespace {

constexpr uint32_t MDB_MAGIC = 0xBEEFC0DE;
constexpr uint32_t MDB_DATA_VERSION = 1;
constexpr size_t PAGEHDRSZ = 16, NODESZ = 8, DBSZ = 48;
constexpr uint16_t P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08, P_LEAF2 = 0x20;
constexpr uint16_t F_BIGDATA = 0x01, F_SUBDATA = 0x02, F_DUPDATA = 0x04;
constexpr uint16_t MDB_DUPSORT = 0x04;
constexpr uint64_t P_INVALID = ~(uint64_t)0;

inline uint16_t rd16(const uint8_t* p) { uint16_t v; std::memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }

struct Db {
    uint16_t flags = 0, depth = 0;
    uint64_t entries = 0, root = P_INVALID;
};

struct Meta {
    uint32_t psize = 0;
    uint64_t last_pg = 0, txnid = 0;
    Db main;
};

struct File {
    int fd = -1;
    const uint8_t* base = nullptr;
    size_t size = 0;
    ~File() {
        if (base) munmap((void*)base, size);
        if (fd >= 0) close(fd);
    }
};

Db read_db(const uint8_t* p) {
    Db d;
    d.flags = rd16(p + 4);
    d.depth = rd16(p + 6);
    d.entries = rd64(p + 32);
    d.root = rd64(p + 40);
    return d;
}

// one of the two meta pages; false if it is not a valid meta page
bool read_meta(const File& f, size_t off, Meta* m) {
    if (f.size < off + PAGEHDRSZ + 24 + 2 * DBSZ + 16) return false;
    const uint8_t* pg = f.base + off;
    if (!(rd16(pg + 10) & P_META)) return false;
    const uint8_t* b = pg + PAGEHDRSZ;
    if (rd32(b) != MDB_MAGIC || rd32(b + 4) != MDB_DATA_VERSION) return false;
    m->psize = rd32(b + 24);  // dbs[FREE].pad
    m->main = read_db(b + 24 + DBSZ);
    m->last_pg = rd64(b + 24 + 2 * DBSZ);
    m->txnid = rd64(b + 24 + 2 * DBSZ + 8);
    return true;
}

hb_status pick_meta(const File& f, Meta* out) {
    Meta m0, m1;
    if (!read_meta(f, 0, &m0)) { set_error("lmdb: not an LMDB data file (bad magic/version in meta page 0)"); return HB_EFORMAT; }
    uint32_t ps = m0.psize;
    if (ps < 512 || ps > 65536 || (ps & (ps - 1))) { set_error("lmdb: implausible page size %u", ps); return HB_EFORMAT; }
    bool ok1 = read_meta(f, ps, &m1) && m1.psize == ps;
    *out = (ok1 && m1.txnid > m0.txnid) ? m1 : m0;
    const uint64_t n_file_pages = f.size / ps;
    if (n_file_pages < 2) { set_error("lmdb: data file shorter than its two meta pages"); return HB_EFORMAT; }
    if (out->last_pg > n_file_pages - 1) out->last_pg = n_file_pages - 1;  // untrusted: never beyond the mapping
    return HB_OK;
}

struct Walker {
    const File& f;
    Meta meta;
    lmdb_visit_fn fn;
    void* user;
    const uint8_t* lo;  // inclusive lower key bound (prefix) or null
    size_t lo_len;
    bool stop = false;  // set once a key beyond the prefix range was met (keys are sorted)
    hb_status st = HB_OK;

    const uint8_t* page(uint64_t pgno, uint64_t n_pages = 1) {
        // compared in page units: `(pgno + n_pages) * psize` can wrap for page numbers taken from a damaged file
        const uint64_t n_file_pages = f.size / meta.psize;
        if (n_pages == 0 || pgno > meta.last_pg || pgno >= n_file_pages || n_pages > n_file_pages - pgno || n_pages - 1 > meta.last_pg - pgno) {
            set_error("lmdb: page %llu (+%llu) outside the data file", (unsigned long long)pgno, (unsigned long long)n_pages);
            st = HB_EFORMAT;
            return nullptr;
        }
        return f.base + pgno * (uint64_t)meta.psize;
    }
    static unsigned n_keys(const uint8_t* pg) { return (unsigned)((rd16(pg + 12) - PAGEHDRSZ) >> 1); }
    // node i of a page, bounds-checked; returns null on corruption
    const uint8_t* node(const uint8_t* pg, unsigned i, size_t* room) {
        uint16_t off = rd16(pg + PAGEHDRSZ + 2 * i);
        if (off < PAGEHDRSZ || (size_t)off + NODESZ > meta.psize) { set_error("lmdb: node offset %u outside its page", off); st = HB_EFORMAT; return nullptr; }
        *room = meta.psize - off - NODESZ;
        return pg + off;
    }
    // <0 / 0 / >0: the key's first lo_len bytes against the prefix (shorter keys compare as bytewise prefixes do)
    int cmp_prefix(const uint8_t* k, size_t kl) const {
        if (!lo) return 0;
        size_t n = kl < lo_len ? kl : lo_len;
        int c = std::memcmp(k, lo, n);
        if (c) return c;
        return kl < lo_len ? -1 : 0;
    }

    void walk(uint64_t pgno, unsigned depth) {
        if (stop || st != HB_OK) return;
        if (depth > 64) { set_error("lmdb: tree deeper than 64 levels (cycle?)"); st = HB_EFORMAT; return; }
        const uint8_t* pg = page(pgno);
        if (!pg) return;
        uint16_t flags = rd16(pg + 10), lower = rd16(pg + 12), upper = rd16(pg + 14);
        if (lower < PAGEHDRSZ || lower > upper || upper > meta.psize || (flags & P_LEAF2)) {
            set_error("lmdb: corrupt page %llu (flags %#x lower %u upper %u)", (unsigned long long)pgno, flags, lower, upper);
            st = HB_EFORMAT;
            return;
        }
        unsigned nk = n_keys(pg);
        if (flags & P_BRANCH) {
            for (unsigned i = 0; i < nk && !stop && st == HB_OK; ++i) {
                size_t room;
                const uint8_t* nd = node(pg, i, &room);
                if (!nd) return;
                // child i holds keys in [key_i, key_{i+1}); key_0 is implicit (-inf).  Skip children entirely below the prefix.
                if (i + 1 < nk) {
                    size_t room2;
                    const uint8_t* nx = node(pg, i + 1, &room2);
                    if (!nx) return;
                    uint16_t ks2 = rd16(nx + 6);
                    if (ks2 > room2) { set_error("lmdb: branch key overruns its page"); st = HB_EFORMAT; return; }
                    if (cmp_prefix(nx + NODESZ, ks2) < 0) continue;
                }
                uint16_t ks = rd16(nd + 6);
                if (ks > room) { set_error("lmdb: branch key overruns its page"); st = HB_EFORMAT; return; }
                if (i > 0 && cmp_prefix(nd + NODESZ, ks) > 0) { stop = true; return; }
                uint64_t child = (uint64_t)rd16(nd) | ((uint64_t)rd16(nd + 2) << 16) | ((uint64_t)rd16(nd + 4) << 32);
                walk(child, depth + 1);
            }
        } else if (flags & P_LEAF) {
            for (unsigned i = 0; i < nk && st == HB_OK; ++i) {
                size_t room;
                const uint8_t* nd = node(pg, i, &room);
                if (!nd) return;
                uint16_t nflags = rd16(nd + 4), ks = rd16(nd + 6);
                uint32_t dsz = (uint32_t)rd16(nd) | ((uint32_t)rd16(nd + 2) << 16);
                if (ks > room) { set_error("lmdb: leaf key overruns its page"); st = HB_EFORMAT; return; }
                const uint8_t* key = nd + NODESZ;
                int c = cmp_prefix(key, ks);
                if (c < 0) continue;
                if (c > 0) { stop = true; return; }
                const uint8_t* val;
                if (nflags & F_BIGDATA) {
                    if ((size_t)ks + 8 > room) { set_error("lmdb: overflow reference overruns its page"); st = HB_EFORMAT; return; }
                    uint64_t opg = rd64(key + ks);
                    const uint8_t* op = page(opg);
                    if (!op) return;
                    if (!(rd16(op + 10) & P_OVERFLOW)) { set_error("lmdb: page %llu is not an overflow page", (unsigned long long)opg); st = HB_EFORMAT; return; }
                    uint32_t ovpages = rd32(op + 12);
                    if (!page(opg, ovpages ? ovpages : 1)) return;
                    if ((uint64_t)PAGEHDRSZ + dsz > (uint64_t)ovpages * meta.psize) { set_error("lmdb: value larger than its overflow run"); st = HB_EFORMAT; return; }
                    val = op + PAGEHDRSZ;
                } else {
                    if ((size_t)ks + dsz > room) { set_error("lmdb: leaf value overruns its page"); st = HB_EFORMAT; return; }
                    val = key + ks;
                }
                hb_status s = fn(user, key, ks, val, dsz, nflags);
                if (s != HB_OK) { st = s; return; }
            }
        } else {
            set_error("lmdb: page %llu is neither branch nor leaf (flags %#x)", (unsigned long long)pgno, flags);
            st = HB_EFORMAT;
        }
    }
};

struct FindDb {
    const char* name;
    size_t len;
    bool found = false, is_db = false;
    Db db;
};
hb_status find_db_cb(void* u, const uint8_t* k, size_t kl, const uint8_t* v, size_t vl, unsigned flags) {
    FindDb* fd = (FindDb*)u;
    if (kl != fd->len || std::memcmp(k, fd->name, kl)) return HB_OK;
    fd->found = true;
    if ((flags & F_SUBDATA) && vl >= DBSZ) { fd->is_db = true; fd->db = read_db(v); }
    return HB_OK;
}

hb_status open_file(const char* path, File* f) {
    std::string p(path ? path : "");
    if (p.empty()) { set_error("lmdb: empty path"); return HB_EINVAL; }
    struct stat sb;
    if (stat(p.c_str(), &sb) != 0) { set_error("lmdb: cannot stat %s", p.c_str()); return HB_EINVAL; }
    if (S_ISDIR(sb.st_mode)) {  // an environment directory (heed's EnvOpenOptions::open(dir)); MDB_NOSUBDIR = the file itself
        p += "/data.mdb";
        if (stat(p.c_str(), &sb) != 0) { set_error("lmdb: no data.mdb under %s", path); return HB_EINVAL; }
    }
    f->fd = open(p.c_str(), O_RDONLY | O_CLOEXEC);
    if (f->fd < 0) { set_error("lmdb: cannot open %s", p.c_str()); return HB_EINVAL; }
    f->size = (size_t)sb.st_size;
    if (f->size < 2 * 512) { set_error("lmdb: %s is too small to hold two meta pages", p.c_str()); return HB_EFORMAT; }
    void* m = mmap(nullptr, f->size, PROT_READ, MAP_SHARED, f->fd, 0);
    if (m == MAP_FAILED) { f->base = nullptr; set_error("lmdb: mmap of %s failed", p.c_str()); return HB_ENOMEM; }
    f->base = (const uint8_t*)m;
    madvise(m, f->size, MADV_SEQUENTIAL);
    return HB_OK;
}

}  // namespace

hb_status lmdb_scan(const char* path, const char* db_name, const uint8_t* prefix, size_t prefix_len, lmdb_visit_fn fn, void* user,
                    uint64_t* txnid_out) {
    File f;
    hb_status st = open_file(path, &f);
    if (st != HB_OK) return st;
    Meta meta;
    if ((st = pick_meta(f, &meta)) != HB_OK) return st;
    Db db = meta.main;
    if (db_name && *db_name) {
        FindDb fd{db_name, std::strlen(db_name)};
        if (meta.main.root != P_INVALID) {
            Walker w{f, meta, find_db_cb, &fd, (const uint8_t*)db_name, fd.len};
            w.walk(meta.main.root, 0);
            if (w.st != HB_OK) return w.st;
        }
        if (!fd.found || !fd.is_db) { set_error("lmdb: no database named `%s` in the environment", db_name); return HB_EINVAL; }
        db = fd.db;
    }
    if (db.flags & MDB_DUPSORT) { set_error("lmdb: DUPSORT databases are not hannoy databases"); return HB_EFORMAT; }
    if (db.root != P_INVALID) {
        Walker w{f, meta, fn, user, prefix_len ? prefix : nullptr, prefix_len};
        w.walk(db.root, 0);
        if (w.st != HB_OK) return w.st;
    }
    // a commit that landed while we walked may have let a later writer recycle pages of our snapshot
    Meta again;
    if (pick_meta(f, &again) != HB_OK || again.txnid != meta.txnid) {
        set_error("lmdb: the environment was committed to during the snapshot (txn %llu -> %llu)", (unsigned long long)meta.txnid,
                  (unsigned long long)again.txnid);
        return HB_ESTATE;
    }
    if (txnid_out) *txnid_out = meta.txnid;
    return HB_OK;
}

}  // namespace hb
