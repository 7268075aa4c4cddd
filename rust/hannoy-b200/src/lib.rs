//! `GpuReader` — a drop-in for `hannoy::Reader` (hannoy src/reader.rs:374-431) whose searches run on a
//! B200 through libhannoy_b200.so.
//!
//! ```ignore
//! let reader = GpuReader::<Cosine>::open(&rtxn, 0, db, /*device*/ 0)?;       // Reader::open + snapshot
//! let nns = reader.nns(10).ef_search(128).by_vector(&rtxn, &q)?.into_nns();   // same call chain as hannoy
//! let all = reader.nns(10).ef_search(128).by_vectors(&queries)?;              // new: one launch per batch
//! ```
//! The LMDB read transaction is consumed ONCE, in `open`: every `(key, value)` pair of the index is
//! handed to `hb_index_push_kv` as raw bytes (the C++ side decodes KeyCodec / NodeCodec / MetadataCodec
//! / Roaring), then `hb_index_finalize` flattens and uploads. Queries ignore the `rtxn` argument,
//! which is kept only for source compatibility.
mod ffi;

use std::ffi::{CStr, CString};
use std::marker::PhantomData;

use hannoy::{Database, Distance, Error, ItemId, Result};
use heed::types::Bytes;
use heed::RoTxn;
use roaring::RoaringBitmap;

pub use hannoy::Searched;

const DEFAULT_EF_SEARCH: usize = 100; // reader.rs:23
const DEFAULT_LINEAR_SCAN_THRESHOLD: usize = 1000; // reader.rs:29
const DEFAULT_LINEAR_SCAN_THRESHOLD_RATIO: f32 = 1.0; // reader.rs:32

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::hb_last_error()).to_string_lossy().into_owned() }
}

fn check(st: ffi::hb_status, dims: (usize, usize)) -> Result<()> {
    match st {
        ffi::HB_OK => Ok(()),
        ffi::HB_EDIM => Err(Error::InvalidVecDimension { expected: dims.0, received: dims.1 }),
        ffi::HB_EMISSING_METADATA => Err(Error::MissingMetadata(0)),
        ffi::HB_ENEED_BUILD => Err(Error::NeedBuild(0)),
        _ => Err(Error::Io(std::io::Error::other(last_error()))),
    }
}

pub struct GpuReader<D: Distance> {
    raw: *mut ffi::hb_index,
    dimensions: usize,
    device: i32,
    _marker: PhantomData<D>,
}

// An hb_index is immutable after finalize and the library serialises its own workspaces: same
// guarantees as `Reader: Send + Sync` over an MVCC snapshot.
unsafe impl<D: Distance> Send for GpuReader<D> {}
unsafe impl<D: Distance> Sync for GpuReader<D> {}

impl<D: Distance> Drop for GpuReader<D> {
    fn drop(&mut self) {
        unsafe { ffi::hb_index_free(self.raw) }
    }
}

impl<D: Distance> GpuReader<D> {
    /// `Reader::open` (reader.rs:387-431): same checks, then the one-off snapshot to `device`.
    pub fn open(rtxn: &RoTxn, index: u16, database: Database<D>, device: i32) -> Result<Self> {
        let name = CString::new(D::name()).unwrap();
        let metric = unsafe { ffi::hb_metric_from_name(name.as_ptr()) };
        assert!(metric >= 0, "distance {:?} has no GPU kernel", D::name());
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::hb_index_begin(metric, index, &mut raw) }, (0, 0))?;
        let mut this = GpuReader { raw, dimensions: 0, device, _marker: PhantomData };
        // every key of this index starts with its big-endian u16 (key.rs:54-66)
        let prefix = index.to_be_bytes();
        let raw_db = database.remap_types::<Bytes, Bytes>();
        for kv in raw_db.prefix_iter(rtxn, &prefix)? {
            let (k, v) = kv?;
            check(unsafe { ffi::hb_index_push_kv(this.raw, k.as_ptr(), k.len(), v.as_ptr(), v.len()) }, (0, 0))?;
        }
        match unsafe { ffi::hb_index_finalize(this.raw, device) } {
            ffi::HB_EUNMATCHING_DISTANCE => {
                return Err(Error::UnmatchingDistance { expected: last_error(), received: D::name() })
            }
            st => check(st, (0, 0))?,
        }
        this.dimensions = unsafe { ffi::hb_index_dimensions(this.raw) } as usize;
        Ok(this) // (a struct update `..this` would move out of a Drop type)
    }

    /// `Reader::open` without heed: the library walks `<path>/data.mdb` itself (`db_name = None` is the unnamed
    /// database), for processes that do not hold the LMDB environment open.
    pub fn open_path(path: &std::path::Path, db_name: Option<&str>, index: u16, device: i32) -> Result<Self> {
        let name = CString::new(D::name()).unwrap();
        let metric = unsafe { ffi::hb_metric_from_name(name.as_ptr()) };
        let cpath = CString::new(path.to_string_lossy().as_bytes()).unwrap();
        let cdb = db_name.map(|n| CString::new(n).unwrap());
        let mut raw = std::ptr::null_mut();
        let st = unsafe {
            ffi::hb_index_open_lmdb(cpath.as_ptr(), cdb.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), metric, index, device, &mut raw)
        };
        match st {
            ffi::HB_EUNMATCHING_DISTANCE => return Err(Error::UnmatchingDistance { expected: last_error(), received: D::name() }),
            st => check(st, (0, 0))?,
        }
        let dimensions = unsafe { ffi::hb_index_dimensions(raw) } as usize;
        Ok(GpuReader { raw, dimensions, device, _marker: PhantomData })
    }

    /// `HannoyBuilder::build` on the GPU for a database whose items were added by `Writer::add_item` but whose graph is not
    /// built (or is to be rebuilt): snapshots the Item pairs, builds on `device`, writes the Metadata and Links pairs back
    /// through `wtxn` (what writer.rs:521-603 / hnsw.rs:190-212 write) and removes the `Updated` stones, then returns the
    /// reader over the new graph.  `M` / `M0` are runtime values here (<= 32).
    #[allow(clippy::too_many_arguments)]
    pub fn build_and_open(
        wtxn: &mut heed::RwTxn, index: u16, database: Database<D>, dimensions: usize, m: u32, m0: u32, ef_construction: u32, alpha: f32,
        seed: u64, device: i32,
    ) -> Result<Self> {
        let name = CString::new(D::name()).unwrap();
        let metric = unsafe { ffi::hb_metric_from_name(name.as_ptr()) };
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::hb_index_begin(metric, index, &mut raw) }, (0, 0))?;
        let this = GpuReader { raw, dimensions, device, _marker: PhantomData };
        let raw_db = database.remap_types::<Bytes, Bytes>();
        let mut stones = Vec::new();
        let mut stale_links: Vec<Vec<u8>> = Vec::new();
        for kv in raw_db.prefix_iter(wtxn, &index.to_be_bytes())? {
            let (k, v) = kv?;
            match k[2] {
                3 => check(unsafe { ffi::hb_index_push_kv(this.raw, k.as_ptr(), k.len(), v.as_ptr(), v.len()) }, (0, 0))?, // Item nodes
                1 => stones.push(k.to_vec()),                                                                          // Updated
                2 => stale_links.push(k.to_vec()),                                                                     // Links of the previous build
                _ => {}
            }
        }
        let opts = ffi::hb_build_opts { m, m0, ef_construction, alpha, seed, batch_max: 0, dimensions: dimensions as u32 };
        check(unsafe { ffi::hb_index_build_graph(this.raw, &opts, device, std::ptr::null_mut()) }, (0, 0))?;
        extern "C" fn collect(user: *mut std::os::raw::c_void, k: *const u8, kl: usize, v: *const u8, vl: usize) -> i32 {
            let out = unsafe { &mut *(user as *mut Vec<(Vec<u8>, Vec<u8>)>) };
            out.push(unsafe { (std::slice::from_raw_parts(k, kl).to_vec(), std::slice::from_raw_parts(v, vl).to_vec()) });
            0
        }
        let mut pairs: Vec<(Vec<u8>, Vec<u8>)> = Vec::new();
        check(unsafe { ffi::hb_index_export_kv(this.raw, 0, collect, &mut pairs as *mut _ as *mut _) }, (0, 0))?;
        // the graph is rebuilt from scratch: links of items that no longer exist (or of layers an item no longer reaches)
        // must not survive (writer.rs:577 `delete_links_from_db`)
        for k in &stale_links {
            raw_db.delete(wtxn, k)?;
        }
        for (k, v) in &pairs {
            raw_db.put(wtxn, k, v)?;
        }
        for k in &stones {
            raw_db.delete(wtxn, k)?;
        }
        check(unsafe { ffi::hb_index_finalize(this.raw, device) }, (0, 0))?;
        Ok(this)
    }

    /// Copies the HBM-resident snapshot to further GPUs; `by_vectors` then partitions every batch into contiguous
    /// slices, one per device (hannoy shares one `Reader` between the threads of a rayon pool, src/parallel.rs:18-38).
    pub fn replicate(&mut self, devices: &[i32]) -> Result<()> {
        check(unsafe { ffi::hb_index_replicate(self.raw, devices.as_ptr(), devices.len() as i32) }, (0, 0))
    }
    pub fn n_devices(&self) -> usize {
        unsafe { ffi::hb_index_n_devices(self.raw) as usize }
    }
    pub fn dimensions(&self) -> usize {
        self.dimensions
    }
    pub fn n_items(&self) -> u64 {
        unsafe { ffi::hb_index_n_items(self.raw) }
    }
    pub fn is_empty(&self) -> bool {
        self.n_items() == 0
    }
    pub fn contains_item(&self, item: ItemId) -> bool {
        unsafe { ffi::hb_index_contains_item(self.raw, item) != 0 }
    }
    pub fn item_ids(&self) -> RoaringBitmap {
        let n = self.n_items();
        let mut ids = vec![0u32; n as usize];
        unsafe { ffi::hb_index_item_ids(self.raw, ids.as_mut_ptr(), n) };
        RoaringBitmap::from_sorted_iter(ids).unwrap()
    }
    pub fn item_vector(&self, item: ItemId) -> Option<Vec<f32>> {
        let mut v = vec![0f32; self.dimensions];
        (unsafe { ffi::hb_index_item_vector(self.raw, item, v.as_mut_ptr()) } == ffi::HB_OK).then_some(v)
    }

    /// `Reader::nns` (reader.rs:611-620)
    pub fn nns(&self, count: usize) -> GpuQueryBuilder<'_, D> {
        GpuQueryBuilder {
            reader: self,
            candidates: None,
            count,
            ef: DEFAULT_EF_SEARCH,
            linear_below: DEFAULT_LINEAR_SCAN_THRESHOLD,
            linear_below_ratio: DEFAULT_LINEAR_SCAN_THRESHOLD_RATIO,
            cancel: std::ptr::null(),
        }
    }
}

/// `QueryBuilder` (reader.rs:60-261) — same fields, same setters.
pub struct GpuQueryBuilder<'a, D: Distance> {
    reader: &'a GpuReader<D>,
    candidates: Option<&'a RoaringBitmap>,
    count: usize,
    ef: usize,
    linear_below: usize,
    linear_below_ratio: f32,
    cancel: *const ffi::hb_cancel_token,
}

/// Runs `search` while a watcher thread evaluates `cancel_fn` on the host and trips a device-side token the first time
/// it returns true: the kernels poll that token where the reference calls `cancel_fn` (reader.rs:330,684).
fn with_watcher<T>(device: i32, cancel_fn: impl Fn() -> bool + Sync, search: impl FnOnce(*const ffi::hb_cancel_token) -> Result<T>) -> Result<T> {
    use std::sync::atomic::{AtomicBool, Ordering};
    let mut tok = std::ptr::null_mut();
    check(unsafe { ffi::hb_cancel_token_create(device, &mut tok) }, (0, 0))?;
    struct SendPtr(*mut ffi::hb_cancel_token);
    unsafe impl Send for SendPtr {}
    unsafe impl Sync for SendPtr {}
    let tokp = SendPtr(tok);
    let done = AtomicBool::new(false);
    let res = std::thread::scope(|s| {
        s.spawn(|| {
            while !done.load(Ordering::Acquire) {
                if cancel_fn() {
                    unsafe { ffi::hb_cancel_token_cancel(tokp.0) };
                    break;
                }
                std::thread::sleep(std::time::Duration::from_micros(50));
            }
        });
        let r = search(tok as *const _);
        done.store(true, Ordering::Release);
        r
    });
    unsafe { ffi::hb_cancel_token_free(tok) };
    res
}

impl<'a, D: Distance> GpuQueryBuilder<'a, D> {
    pub fn ef_search(&mut self, ef: usize) -> &mut Self {
        self.ef = ef.max(self.count); // reader.rs:217-220
        self
    }
    pub fn candidates(&mut self, candidates: &'a RoaringBitmap) -> &mut Self {
        self.candidates = Some(candidates);
        self
    }
    pub fn linear_below(&mut self, threshold: usize) -> &mut Self {
        self.linear_below = threshold;
        self
    }
    pub fn linear_below_ratio(&mut self, ratio: f32) -> &mut Self {
        self.linear_below_ratio = ratio;
        self
    }

    fn opts(&self, ids: &'a [u32]) -> ffi::hb_query_opts {
        ffi::hb_query_opts {
            candidates: ids.as_ptr(),
            n_candidates: ids.len() as u64,
            has_candidates: self.candidates.is_some() as i32,
            linear_below: self.linear_below.min(u32::MAX as usize) as u32,
            linear_below_ratio: self.linear_below_ratio,
            cancel: self.cancel,
            cancel_after_polls: 0,
        }
    }

    /// `QueryBuilder::by_vector_with_cancellation` (reader.rs:167-188)
    pub fn by_vector_with_cancellation(&self, rtxn: &RoTxn, vector: &'a [f32], cancel_fn: impl Fn() -> bool + Sync) -> Result<Searched> {
        with_watcher(self.reader.device, cancel_fn, |tok| {
            let qb = GpuQueryBuilder { cancel: tok, ..*self };
            qb.by_vector(rtxn, vector)
        })
    }

    /// `QueryBuilder::by_item_with_cancellation` (reader.rs:108-118)
    pub fn by_item_with_cancellation(&self, rtxn: &RoTxn, item: ItemId, cancel_fn: impl Fn() -> bool + Sync) -> Result<Option<Searched>> {
        with_watcher(self.reader.device, cancel_fn, |tok| {
            let qb = GpuQueryBuilder { cancel: tok, ..*self };
            qb.by_item(rtxn, item)
        })
    }

    fn unpack(&self, nq: usize, ids: Vec<u32>, dist: Vec<f32>, lens: Vec<u32>) -> Vec<Option<Searched>> {
        (0..nq)
            .map(|i| {
                (lens[i] != u32::MAX).then(|| {
                    let n = (lens[i] & !ffi::HB_LEN_CANCELLED) as usize;
                    let row = i * self.count;
                    let nns = (0..n).map(|j| (ids[row + j], dist[row + j])).collect();
                    Searched { nns, did_cancel: lens[i] & ffi::HB_LEN_CANCELLED != 0 }
                })
            })
            .collect()
    }

    /// New: every row of `vectors` (nq x dimensions, row-major) searched in one launch.
    pub fn by_vectors(&self, vectors: &[f32]) -> Result<Vec<Searched>> {
        let d = self.reader.dimensions;
        if d == 0 || vectors.len() % d != 0 {
            return Err(Error::InvalidVecDimension { expected: d, received: vectors.len() });
        }
        let nq = vectors.len() / d;
        let cand: Vec<u32> = self.candidates.map(|c| c.iter().collect()).unwrap_or_default();
        let opts = self.opts(&cand);
        let (mut ids, mut dist, mut lens) = (vec![0u32; nq * self.count], vec![0f32; nq * self.count], vec![0u32; nq]);
        let st = unsafe {
            ffi::hb_search_by_vector(
                self.reader.raw, vectors.as_ptr(), nq as u64, d as u32, self.count as u32, self.ef as u32, &opts,
                ids.as_mut_ptr(), dist.as_mut_ptr(), lens.as_mut_ptr(), std::ptr::null_mut(),
            )
        };
        check(st, (d, d))?;
        Ok(self.unpack(nq, ids, dist, lens).into_iter().map(|s| s.unwrap_or(Searched { nns: vec![], did_cancel: false })).collect())
    }

    /// `QueryBuilder::by_vector` (reader.rs:132-148). `_rtxn` is unused (snapshot already taken).
    pub fn by_vector(&self, _rtxn: &RoTxn, vector: &'a [f32]) -> Result<Searched> {
        if vector.len() != self.reader.dimensions {
            return Err(Error::InvalidVecDimension { expected: self.reader.dimensions, received: vector.len() });
        }
        Ok(self.by_vectors(vector)?.pop().unwrap())
    }

    /// New: batched `by_item`; `None` where the item does not exist (reader.rs:826).
    pub fn by_items(&self, items: &[ItemId]) -> Result<Vec<Option<Searched>>> {
        let nq = items.len();
        let cand: Vec<u32> = self.candidates.map(|c| c.iter().collect()).unwrap_or_default();
        let opts = self.opts(&cand);
        let (mut ids, mut dist, mut lens) = (vec![0u32; nq * self.count], vec![0f32; nq * self.count], vec![0u32; nq]);
        let st = unsafe {
            ffi::hb_search_by_item(
                self.reader.raw, items.as_ptr(), nq as u64, self.count as u32, self.ef as u32, &opts,
                ids.as_mut_ptr(), dist.as_mut_ptr(), lens.as_mut_ptr(), std::ptr::null_mut(),
            )
        };
        check(st, (0, 0))?;
        Ok(self.unpack(nq, ids, dist, lens))
    }

    /// `QueryBuilder::by_item` (reader.rs:81-89).
    pub fn by_item(&self, _rtxn: &RoTxn, item: ItemId) -> Result<Option<Searched>> {
        Ok(self.by_items(&[item])?.pop().unwrap())
    }
}
