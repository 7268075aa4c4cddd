"""Test fixture generator: writes an LMDB 0.9.x data file (MDB_DATA_VERSION 1, little-endian, 64-bit) from sorted
key/value pairs, so that the product's read-only walker (hannoy_b200/csrc/lmdb_walk.cpp) can be exercised without
liblmdb (absent from this image — "LMDB byte-format parity unpinned": no file written by the real library is
available; this writer and the walker are two separate statements of LMDB's published layout).

It follows the layout rules the real library applies when it writes pages:
  * two meta pages (0, 1); the one with the larger txnid is current; dbs[FREE].pad carries the page size;
  * leaf / branch pages: 16-byte header, u16 node pointers growing up from `lower`, nodes packed down from `upper`,
    node sizes rounded up to even; first key of a branch page has size 0 (implicit -inf);
  * a value whose node would exceed nodemax = (((psize - 16) / 2) & ~1) - 2 bytes goes to a run of overflow pages
    (ceil((16 + size) / psize) pages, header on the first only) and the leaf node keeps the u64 page number (F_BIGDATA);
  * named databases are F_SUBDATA records (a 48-byte MDB_db) in the MAIN database, keyed by name.
This is test infrastructure only; nothing in hannoy_b200/ imports it.
"""
import os
import struct

MAGIC, VERSION = 0xBEEFC0DE, 1
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA, F_SUBDATA = 0x01, 0x02
P_INVALID = (1 << 64) - 1
PAGEHDRSZ, NODESZ = 16, 8


def _even(n):
    return (n + 1) & ~1


class Db:
    def __init__(self):
        self.flags = 0
        self.depth = 0
        self.branch_pages = 0
        self.leaf_pages = 0
        self.overflow_pages = 0
        self.entries = 0
        self.root = P_INVALID

    def pack(self, pad=0):
        return struct.pack("<IHHQQQQQ", pad, self.flags, self.depth, self.branch_pages, self.leaf_pages, self.overflow_pages,
                           self.entries, self.root)


class LmdbWriter:
    def __init__(self, psize=4096, fill=1.0, junk_branch_key0=False):
        self.psize = psize
        self.fill = fill                      # fraction of a page the bulk loader uses (real trees are rarely full)
        self.junk_branch_key0 = junk_branch_key0
        self.pages = {}                       # pgno -> bytes (pages 0/1 are written at save time)
        self.next_pg = 2
        self.nodemax = (((psize - PAGEHDRSZ) // 2) & ~1) - 2
        self.named = {}                       # name -> Db
        self.main = Db()

    # -- page allocation --
    def _alloc(self, n=1):
        pg = self.next_pg
        self.next_pg += n
        return pg

    def _finish_page(self, pgno, flags, nodes):
        """nodes = list of packed node byte strings, in key order"""
        page = bytearray(self.psize)
        lower = PAGEHDRSZ + 2 * len(nodes)
        upper = self.psize
        ptrs = []
        for nd in nodes:
            upper -= _even(len(nd))
            page[upper:upper + len(nd)] = nd
            ptrs.append(upper)
        assert lower <= upper, "page overfull"
        struct.pack_into("<QHHHH", page, 0, pgno, 0, flags, lower, upper)
        struct.pack_into(f"<{len(ptrs)}H", page, PAGEHDRSZ, *ptrs)
        self.pages[pgno] = bytes(page)

    def _room(self):
        return int((self.psize - PAGEHDRSZ) * self.fill)

    def _leaf_node(self, db, key, val, node_flags=0):
        if NODESZ + len(key) + len(val) > self.nodemax:
            n_ov = (PAGEHDRSZ - 1 + len(val)) // self.psize + 1
            pg = self._alloc(n_ov)
            run = bytearray(n_ov * self.psize)
            struct.pack_into("<QHHI", run, 0, pg, 0, P_OVERFLOW, n_ov)
            run[PAGEHDRSZ:PAGEHDRSZ + len(val)] = val
            self.pages[pg] = bytes(run)
            db.overflow_pages += n_ov
            data, node_flags = struct.pack("<Q", pg), node_flags | F_BIGDATA
        else:
            data = val
        return struct.pack("<HHHH", len(val) & 0xffff, len(val) >> 16, node_flags, len(key)) + key + data

    @staticmethod
    def _branch_node(key, child):
        return struct.pack("<HHHH", child & 0xffff, (child >> 16) & 0xffff, (child >> 32) & 0xffff, len(key)) + key

    def build_tree(self, pairs, node_flags=None):
        """pairs: list of (key bytes, value bytes) sorted by key (bytewise).  Returns the Db record."""
        db = Db()
        pairs = list(pairs)
        assert all(pairs[i][0] < pairs[i + 1][0] for i in range(len(pairs) - 1)), "keys must be sorted and unique"
        db.entries = len(pairs)
        if not pairs:
            return db
        # leaves
        level = []  # (first key, pgno)
        cur, used, first = [], 0, None
        for i, (k, v) in enumerate(pairs):
            nd = self._leaf_node(db, k, v, node_flags[i] if node_flags else 0)
            need = 2 + _even(len(nd))
            if cur and used + need > self._room():
                pg = self._alloc()
                self._finish_page(pg, P_LEAF, cur)
                level.append((first, pg))
                cur, used = [], 0
            if not cur:
                first = k
            cur.append(nd)
            used += need
        pg = self._alloc()
        self._finish_page(pg, P_LEAF, cur)
        level.append((first, pg))
        db.leaf_pages = len(level)
        db.depth = 1
        # branches
        while len(level) > 1:
            up = []
            cur, used, first = [], 0, None
            for k, child in level:
                if not cur:
                    nk = b"\xde\xad\xbe\xef" if self.junk_branch_key0 else b""
                else:
                    nk = k
                nd = self._branch_node(nk, child)
                need = 2 + _even(len(nd))
                if len(cur) >= 2 and used + need > self._room():
                    pg = self._alloc()
                    self._finish_page(pg, P_BRANCH, cur)
                    up.append((first, pg))
                    cur, used = [], 0
                    nd = self._branch_node(b"\xde\xad\xbe\xef" if self.junk_branch_key0 else b"", child)
                    need = 2 + _even(len(nd))
                if not cur:
                    first = k
                cur.append(nd)
                used += need
            pg = self._alloc()
            self._finish_page(pg, P_BRANCH, cur)
            up.append((first, pg))
            db.branch_pages += len(up)
            db.depth += 1
            level = up
        db.root = level[0][1]
        return db

    def put_unnamed(self, pairs):
        """pairs live directly in the MAIN database (heed: env.create_database(&mut wtxn, None))."""
        assert not self.named
        self.main = self.build_tree(pairs)

    def put_named(self, name, pairs):
        self.named[name] = self.build_tree(pairs)

    def _meta(self, pgno, txnid, main, last_pg):
        page = bytearray(self.psize)
        struct.pack_into("<QHHHH", page, 0, pgno, 0, P_META, 0, 0)
        free = Db()
        body = struct.pack("<IIQQ", MAGIC, VERSION, 0, 1 << 30) + free.pack(pad=self.psize) + main.pack() + struct.pack("<QQ", last_pg, txnid)
        page[PAGEHDRSZ:PAGEHDRSZ + len(body)] = body
        return bytes(page)

    def save(self, path, txnid=7, newest_meta=1, nosubdir=False, stale_main=None):
        """Writes <path>/data.mdb (or `path` itself when nosubdir).  The other meta page carries txnid-1 and
        `stale_main` (default: an empty MAIN database), as after a previous commit."""
        if self.named:
            recs = sorted((n.encode(), d.pack()) for n, d in self.named.items())
            self.main = self.build_tree(recs, node_flags=[F_SUBDATA] * len(recs))
        last_pg = self.next_pg - 1
        metas = [None, None]
        metas[newest_meta] = self._meta(newest_meta, txnid, self.main, last_pg)
        metas[1 - newest_meta] = self._meta(1 - newest_meta, txnid - 1, stale_main or Db(), last_pg)
        if nosubdir:
            fn = path
        else:
            os.makedirs(path, exist_ok=True)
            fn = os.path.join(path, "data.mdb")
        with open(fn, "wb") as f:
            f.write(metas[0])
            f.write(metas[1])
            pg = 2
            while pg < self.next_pg:
                blob = self.pages[pg]
                f.write(blob)
                pg += len(blob) // self.psize
        return fn
