//! Raw C-ABI of libhannoy_b200.so — one `extern` per declaration in include/hannoy_b200.h.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct hb_index {
    _private: [u8; 0],
}

pub type hb_status = c_int;
pub const HB_OK: hb_status = 0;
pub const HB_EDIM: hb_status = 2;
pub const HB_EMISSING_METADATA: hb_status = 7;
pub const HB_EUNMATCHING_DISTANCE: hb_status = 8;
pub const HB_ENEED_BUILD: hb_status = 9;

pub type hb_metric = c_int; // 0 euclidean .. 6 binary quantized manhattan, see hb_metric_from_name

#[repr(C)]
pub struct hb_cancel_token {
    _private: [u8; 0],
}

#[repr(C)]
pub struct hb_query_opts {
    pub candidates: *const u32,
    pub n_candidates: u64,
    pub has_candidates: c_int,
    pub linear_below: u32,
    pub linear_below_ratio: f32,
    pub cancel: *const hb_cancel_token,
    pub cancel_after_polls: u64,
}
pub const HB_LEN_CANCELLED: u32 = 0x8000_0000;

#[repr(C)]
pub struct hb_build_opts {
    pub m: u32,
    pub m0: u32,
    pub ef_construction: u32,
    pub alpha: f32,
    pub seed: u64,
    pub batch_max: u32,
    pub dimensions: u32,
}
pub type hb_kv_visit = extern "C" fn(user: *mut c_void, key: *const u8, klen: usize, val: *const u8, vlen: usize) -> c_int;

extern "C" {
    pub fn hb_metric_from_name(name: *const c_char) -> c_int;
    pub fn hb_index_begin(m: hb_metric, index: u16, out: *mut *mut hb_index) -> hb_status;
    pub fn hb_index_push_kv(ix: *mut hb_index, key: *const u8, klen: usize, val: *const u8, vlen: usize) -> hb_status;
    pub fn hb_index_finalize(ix: *mut hb_index, device: c_int) -> hb_status;
    pub fn hb_index_open_lmdb(
        path: *const c_char, db_name: *const c_char, m: hb_metric, index: u16, device: c_int, out: *mut *mut hb_index,
    ) -> hb_status;
    pub fn hb_index_build_graph(ix: *mut hb_index, opts: *const hb_build_opts, device: c_int, stats_out: *mut u64) -> hb_status;
    pub fn hb_index_export_kv(ix: *const hb_index, with_items: c_int, f: hb_kv_visit, user: *mut c_void) -> hb_status;
    pub fn hb_cancel_token_create(device: c_int, out: *mut *mut hb_cancel_token) -> hb_status;
    pub fn hb_cancel_token_cancel(t: *mut hb_cancel_token) -> hb_status;
    pub fn hb_cancel_token_free(t: *mut hb_cancel_token);
    pub fn hb_index_free(ix: *mut hb_index);
    pub fn hb_index_dimensions(ix: *const hb_index) -> u32;
    pub fn hb_index_n_items(ix: *const hb_index) -> u64;
    pub fn hb_index_n_entry_points(ix: *const hb_index) -> u32;
    pub fn hb_index_max_level(ix: *const hb_index) -> u32;
    pub fn hb_index_version(ix: *const hb_index, major: *mut u32, minor: *mut u32, patch: *mut u32) -> hb_status;
    pub fn hb_index_item_ids(ix: *const hb_index, out: *mut u32, cap: u64) -> u64;
    pub fn hb_index_contains_item(ix: *const hb_index, item: u32) -> c_int;
    pub fn hb_index_item_vector(ix: *const hb_index, item: u32, out: *mut f32) -> hb_status;
    pub fn hb_search_by_vector(
        ix: *const hb_index, q: *const f32, nq: u64, dims: u32, count: u32, ef: u32, opts: *const hb_query_opts,
        out_ids: *mut u32, out_dist: *mut f32, out_len: *mut u32, out_counters: *mut u64,
    ) -> hb_status;
    pub fn hb_search_by_item(
        ix: *const hb_index, items: *const u32, nq: u64, count: u32, ef: u32, opts: *const hb_query_opts,
        out_ids: *mut u32, out_dist: *mut f32, out_len: *mut u32, out_counters: *mut u64,
    ) -> hb_status;
    pub fn hb_exact_knn(ix: *const hb_index, q: *const f32, nq: u64, dims: u32, k: u32, out_ids: *mut u32, out_dist: *mut f32) -> hb_status;
    pub fn hb_last_error() -> *const c_char;
    pub fn hb_index_replicate(ix: *mut hb_index, devices: *const c_int, n_dev: c_int) -> hb_status;
    pub fn hb_index_n_devices(ix: *const hb_index) -> c_int;
    #[allow(dead_code)]
    pub fn hb_search_by_vector_device(
        ix: *const hb_index, d_q: *const f32, nq: u64, count: u32, ef: u32, d_out_ids: *mut u32, d_out_dist: *mut f32,
        d_out_len: *mut u32, d_out_counters: *mut u64, stream: *mut c_void,
    ) -> hb_status;
}
