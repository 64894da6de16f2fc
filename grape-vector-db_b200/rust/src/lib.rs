//! gvdb-sys: Rust binding of the C ABI in include/gvdb.h, plus `GpuVectorIndex`, an
//! `impl VectorIndex` (reference src/index.rs:35-62) that can be dropped in at the wiring point
//! src/lib.rs:256-261 in place of `HnswVectorIndex::new()` / `FaissVectorIndex::new(..)`.
//!
//! Not compiled in this repository (the image has no Rust toolchain); the C++ mirror in
//! ../host/gvdb_host.hpp has the same structure and IS tested on a B200.
#![allow(non_camel_case_types)]
use std::collections::HashMap;
use std::ffi::{c_char, c_void, CStr};

#[repr(C)]
pub struct gvdb_index { _private: [u8; 0] }
#[repr(C)]
pub struct gvdb_sparse { _private: [u8; 0] }

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct gvdb_config {
    pub struct_size: u32,
    pub dim: u32,
    pub threshold: f32,
    pub rescore_ratio: f32,
    pub device: i32,
    pub flags: u32,
    pub capacity_rows: u64,
    pub row_base: u64,
    /// GVDB_FLAG_ROW_WINDOW (flags bit 0): f32 rows kept for local rows [window_first, window_first + window_count)
    pub window_first: u64,
    pub window_count: u64,
}
pub const GVDB_FLAG_ROW_WINDOW: u32 = 1;

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct gvdb_stats {
    pub vector_count: u64,
    pub rows: u64,
    pub dimension: u64,
    pub memory_usage: u64,
    pub hbm_bytes: u64,
    pub code_bytes_per_row: u64,
}

pub const GVDB_NO_ID: u64 = u64::MAX;

extern "C" {
    pub fn gvdb_abi_version() -> u32;
    pub fn gvdb_last_error() -> *const c_char;
    pub fn gvdb_create(cfg: *const gvdb_config, out: *mut *mut gvdb_index) -> i32;
    pub fn gvdb_destroy(h: *mut gvdb_index);
    pub fn gvdb_add(h: *mut gvdb_index, rows: *const f32, n: u64, first_row_out: *mut u64) -> i32;
    pub fn gvdb_add_device(h: *mut gvdb_index, stream: *mut c_void, rows_dev: *const f32, n: u64, first_row_out: *mut u64) -> i32;
    pub fn gvdb_reserve(h: *mut gvdb_index, capacity_rows: u64) -> i32;
    pub fn gvdb_remove(h: *mut gvdb_index, local_row: u64, was_live_out: *mut i32) -> i32;
    pub fn gvdb_clear(h: *mut gvdb_index) -> i32;
    pub fn gvdb_len(h: *const gvdb_index) -> u64;
    pub fn gvdb_get_stats(h: *const gvdb_index, out: *mut gvdb_stats) -> i32;
    pub fn gvdb_quantize(h: *mut gvdb_index, x: *const f32, n: u64, codes_out: *mut u8) -> i32;
    pub fn gvdb_get_codes(h: *mut gvdb_index, first: u64, n: u64, codes_out: *mut u8) -> i32;
    pub fn gvdb_hamming(h: *mut gvdb_index, q_codes: *const u8, nq: u32, dist_out: *mut u32) -> i32;
    pub fn gvdb_rescore_count(n: u64, ratio: f32) -> u64;
    pub fn gvdb_search_batch(h: *mut gvdb_index, queries: *const f32, nq: u32, k: u32, rescore_count: u32,
                             ids_out: *mut u64, scores_out: *mut f32, cand_ids_out: *mut u64, cand_ham_out: *mut u32) -> i32;
    pub fn gvdb_flat_search_batch(h: *mut gvdb_index, queries: *const f32, nq: u32, k: u32,
                                  ids_out: *mut u64, dist_out: *mut f32) -> i32;
    pub fn gvdb_similarity_search_batch(h: *mut gvdb_index, queries: *const f32, nq: u32, k: u32, threshold: f32,
                                        use_threshold: i32, ids_out: *mut u64, sims_out: *mut f32) -> i32;
    pub fn gvdb_save(h: *mut gvdb_index, path: *const std::os::raw::c_char) -> i32;
    pub fn gvdb_load(path: *const std::os::raw::c_char, device: i32, out: *mut *mut gvdb_index) -> i32;
    pub fn gvdb_export_rows_ipc(h: *mut gvdb_index, handle_out: *mut u8) -> i32;
    pub fn gvdb_attach_peer_rows_ipc(h: *mut gvdb_index, n_owners: u32, rows_per_owner: u64, my_owner: u32,
                                     handles: *const u8) -> i32;
    pub fn gvdb_shard_record_bytes(nq: u32, rescore_count: u32) -> u64;
    pub fn gvdb_search_shard_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32, nq: u32,
                                    rescore_count: u32, records_dev: *mut c_void) -> i32;
    pub fn gvdb_search_shard_sliced_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32, nq: u32,
                                           rescore_count: u32, n_slices: u32, records_dev: *mut c_void) -> i32;
    // ratio mode across row shards (rescore_count > 2048): histograms -> gather -> shard search -> gather -> merge
    pub fn gvdb_shard_hist_bins(h: *const gvdb_index) -> u32;
    pub fn gvdb_shard_hist_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32, nq: u32,
                                  hist_out_dev: *mut u32) -> i32;
    pub fn gvdb_search_shard_ratio_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32, nq: u32,
                                          rescore_count: u64, k: u32, hists_all_dev: *const u32, n_shards: u32,
                                          my_shard: u32, records_dev: *mut c_void) -> i32;
    pub fn gvdb_merge_shards_ratio_device(h: *mut gvdb_index, stream: *mut c_void, n_shards: u32,
                                          records_dev: *const c_void, nq: u32, k: u32, ids_out_dev: *mut u64,
                                          scores_out_dev: *mut f32) -> i32;
    pub fn gvdb_merge_shards_device(h: *mut gvdb_index, stream: *mut c_void, n_shards: u32, records_dev: *const c_void,
                                    nq: u32, rescore_count: u32, k: u32, ids_out_dev: *mut u64, scores_out_dev: *mut f32) -> i32;
    // filtered searches: allow_bits has ceil(rows / 32) words, bit (r % 32) of word r / 32 = row r may be returned
    pub fn gvdb_search_batch_filtered(h: *mut gvdb_index, queries: *const f32, allow_bits: *const u32, nq: u32, k: u32,
                                      rescore_count: u32, ids_out: *mut u64, scores_out: *mut f32) -> i32;
    pub fn gvdb_search_batch_filtered_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32,
                                             allow_bits_dev: *const u32, nq: u32, k: u32, rescore_count: u32,
                                             ids_out_dev: *mut u64, scores_out_dev: *mut f32) -> i32;
    pub fn gvdb_flat_search_batch_filtered(h: *mut gvdb_index, queries: *const f32, allow_bits: *const u32, nq: u32, k: u32,
                                           ids_out: *mut u64, dist_out: *mut f32) -> i32;
    // peer exchange: codes replicated, rows sharded, queries partitioned; no collective library in the data path
    pub fn gvdb_exchange_create(h: *mut gvdb_index, world: u32, rank: u32, rows_per_owner: u64, nq_max: u32, rescore_max: u32) -> i32;
    pub fn gvdb_exchange_export_ipc(h: *mut gvdb_index, handle_out: *mut u8) -> i32;
    pub fn gvdb_exchange_attach_ipc(h: *mut gvdb_index, handles: *const u8) -> i32;
    pub fn gvdb_exchange_mailbox_ptr(h: *const gvdb_index) -> *mut c_void;
    pub fn gvdb_exchange_attach_ptr(h: *mut gvdb_index, mailboxes: *const *mut c_void) -> i32;
    pub fn gvdb_search_exchange_device(h: *mut gvdb_index, stream: *mut c_void, queries_dev: *const f32, nq: u32, k: u32,
                                       rescore_count: u32, ids_out_dev: *mut u64, scores_out_dev: *mut f32) -> i32;
    pub fn gvdb_exchange_status(h: *mut gvdb_index, timed_out_kinds: *mut u32) -> i32;
    // sparse side of the hybrid search (SparseIndex::search_bm25, reference src/sparse.rs:153-222)
    pub fn gvdb_sparse_create(device: i32, k1: f32, b: f32, out: *mut *mut gvdb_sparse) -> i32;
    pub fn gvdb_sparse_destroy(s: *mut gvdb_sparse);
    pub fn gvdb_sparse_build(s: *mut gvdb_sparse, n_docs: u64, n_terms: u32, post_off: *const u64, post_doc: *const u32,
                             post_tf: *const f32, doc_len: *const f32) -> i32;
    pub fn gvdb_sparse_average_document_length(s: *const gvdb_sparse) -> f32;
    pub fn gvdb_sparse_search_bm25_batch(s: *mut gvdb_sparse, nq: u32, q_off: *const u64, q_terms: *const u32,
                                         q_tfs: *const f32, limit: u32, doc_out: *mut u64, score_out: *mut f32) -> i32;
    pub fn gvdb_sparse_search_bm25_batch_device(s: *mut gvdb_sparse, stream: *mut c_void, nq: u32, q_off: *const u64,
                                                q_terms: *const u32, q_tfs: *const f32, limit: u32,
                                                doc_out_dev: *mut u64, score_out_dev: *mut f32) -> i32;
    pub fn gvdb_sparse_launches(s: *const gvdb_sparse) -> u64;
    // HybridSearchEngine::rrf_fusion (reference src/hybrid.rs:422-488) for a batch of requests
    pub fn gvdb_rrf_fusion_batch(device: i32, dense: *const u64, n_dense: u32, sparse: *const u64, n_sparse: u32,
                                 text: *const u64, n_text: u32, nq: u32, k: f32, limit: u32,
                                 ids_out: *mut u64, scores_out: *mut f32) -> i32;
    pub fn gvdb_rrf_fusion_batch_device(device: i32, stream: *mut c_void, dense_dev: *const u64, n_dense: u32,
                                        sparse_dev: *const u64, n_sparse: u32, text_dev: *const u64, n_text: u32,
                                        nq: u32, k: f32, limit: u32, ids_out_dev: *mut u64, scores_out_dev: *mut f32) -> i32;
    // HybridSearchEngine::linear_fusion / normalized_fusion (reference src/hybrid.rs:491-616) for a batch of requests
    pub fn gvdb_weighted_fusion_batch(device: i32, dense: *const u64, dense_scores: *const f32, n_dense: u32,
                                      sparse: *const u64, sparse_scores: *const f32, n_sparse: u32,
                                      text: *const u64, text_scores: *const f32, n_text: u32, nq: u32,
                                      dense_weight: f32, sparse_weight: f32, text_weight: f32, normalize: i32,
                                      limit: u32, ids_out: *mut u64, scores_out: *mut f32) -> i32;
    pub fn gvdb_weighted_fusion_batch_device(device: i32, stream: *mut c_void, dense_dev: *const u64,
                                             dense_scores_dev: *const f32, n_dense: u32, sparse_dev: *const u64,
                                             sparse_scores_dev: *const f32, n_sparse: u32, text_dev: *const u64,
                                             text_scores_dev: *const f32, n_text: u32, nq: u32, dense_weight: f32,
                                             sparse_weight: f32, text_weight: f32, normalize: i32, limit: u32,
                                             ids_out_dev: *mut u64, scores_out_dev: *mut f32) -> i32;
    pub fn gvdb_measure_fp4_mma_rate(device: i32, tmacs_per_s_out: *mut f64, clk_per_mma_out: *mut f64) -> i32;
}

/// Owned handle; `Send + Sync` because the C ABI's search entry points are re-entrant and the
/// mutating ones are only reachable through `&mut self` (the reference's write guard).
pub struct GpuIndex { h: *mut gvdb_index, dim: usize }
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {}

impl Drop for GpuIndex {
    fn drop(&mut self) { unsafe { gvdb_destroy(self.h) } }
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(gvdb_last_error()).to_string_lossy().into_owned() }
}

#[cfg(feature = "vector-index")]
mod trait_impl {
    use super::*;
    use grape_vector_db::index::{IndexStats, VectorIndex};
    use grape_vector_db::types::VectorDbError;

    fn map_err(st: i32) -> VectorDbError {
        match st {
            1 => VectorDbError::IndexNotBuilt,
            // the C ABI takes flat buffers and cannot see a wrong length: both numbers are checked (and the
            // variant raised) on this side before the call; a status 2 from the library carries none
            2 => VectorDbError::DimensionMismatch { expected: 0, actual: 0 },
            3 => VectorDbError::InvalidVectorDimension,
            4 => VectorDbError::QuantizationError(last_error()),
            6 => VectorDbError::ConfigError(last_error()),
            7 => VectorDbError::NotImplemented(last_error()),
            _ => VectorDbError::IndexError(last_error()),
        }
    }

    /// The GPU index behind `Arc<RwLock<dyn VectorIndex>>` (src/lib.rs:238).
    pub struct GpuVectorIndex {
        inner: Option<GpuIndex>,
        id_to_index: HashMap<String, u64>,   // same two maps as FaissVectorIndex (src/index.rs:333-334)
        index_to_id: HashMap<u64, String>,
        exact: bool,
        oversample: u32,
        threshold: f32,
        device: i32,
    }

    impl GpuVectorIndex {
        pub fn new(exact: bool, oversample: u32, device: i32) -> Self {
            Self { inner: None, id_to_index: HashMap::new(), index_to_id: HashMap::new(), exact, oversample, threshold: 0.0, device }
        }
    }

    impl VectorIndex for GpuVectorIndex {
        fn add_vector(&mut self, id: String, vector: Vec<f32>) -> Result<(), VectorDbError> {
            self.add_vectors(vec![(id, vector)])
        }
        fn add_vectors(&mut self, vectors: Vec<(String, Vec<f32>)>) -> Result<(), VectorDbError> {
            if vectors.is_empty() { return Ok(()); }
            if self.inner.is_none() {
                let dim = vectors[0].1.len();
                let cfg = gvdb_config { struct_size: std::mem::size_of::<gvdb_config>() as u32, dim: dim as u32,
                                        threshold: self.threshold, rescore_ratio: 0.1, device: self.device, ..Default::default() };
                let mut h = std::ptr::null_mut();
                let st = unsafe { gvdb_create(&cfg, &mut h) };
                if st != 0 { return Err(map_err(st)); }
                self.inner = Some(GpuIndex { h, dim });
            }
            let ix = self.inner.as_ref().unwrap();
            let mut flat = Vec::with_capacity(vectors.len() * ix.dim);
            for (_, v) in &vectors {
                if v.len() != ix.dim {   // src/index.rs:590-594
                    return Err(VectorDbError::DimensionMismatch { expected: ix.dim, actual: v.len() });
                }
                flat.extend_from_slice(v);
            }
            let mut first = 0u64;
            let st = unsafe { gvdb_add(ix.h, flat.as_ptr(), vectors.len() as u64, &mut first) };
            if st != 0 { return Err(map_err(st)); }
            for (i, (id, _)) in vectors.into_iter().enumerate() {
                if let Some(old) = self.id_to_index.insert(id.clone(), first + i as u64) {
                    let mut was = 0i32;
                    unsafe { gvdb_remove(ix.h, old, &mut was) };
                    self.index_to_id.remove(&old);
                }
                self.index_to_id.insert(first + i as u64, id);
            }
            Ok(())
        }
        fn search(&self, query: &[f32], k: usize) -> Result<Vec<(String, f32)>, VectorDbError> {
            let ix = self.inner.as_ref().ok_or(VectorDbError::IndexNotBuilt)?;   // src/index.rs:621-623
            if query.len() != ix.dim {
                return Err(VectorDbError::DimensionMismatch { expected: ix.dim, actual: query.len() });
            }
            if k == 0 { return Ok(Vec::new()); }
            let mut ids = vec![GVDB_NO_ID; k];
            let mut val = vec![0f32; k];
            let st = unsafe {
                if self.exact {
                    gvdb_flat_search_batch(ix.h, query.as_ptr(), 1, k as u32, ids.as_mut_ptr(), val.as_mut_ptr())
                } else {
                    gvdb_search_batch(ix.h, query.as_ptr(), 1, k as u32, (k as u32) * self.oversample,
                                      ids.as_mut_ptr(), val.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut())
                }
            };
            if st != 0 { return Err(map_err(st)); }
            // score convention of the exact reference path: distance = 1 - cos, ascending (src/index.rs:630-637)
            Ok(ids.iter().zip(val.iter()).take_while(|(i, _)| **i != GVDB_NO_ID)
                  .map(|(i, v)| (self.index_to_id[i].clone(), if self.exact { *v } else { 1.0 - *v })).collect())
        }
        fn remove_vector(&mut self, id: &str) -> Result<bool, VectorDbError> {          // src/index.rs:642-650
            if let Some(row) = self.id_to_index.remove(id) {
                self.index_to_id.remove(&row);
                let mut was = 0i32;
                let st = unsafe { gvdb_remove(self.inner.as_ref().unwrap().h, row, &mut was) };
                if st != 0 { return Err(map_err(st)); }
                Ok(true)
            } else { Ok(false) }
        }
        fn len(&self) -> usize { self.id_to_index.len() }
        fn is_empty(&self) -> bool { self.id_to_index.is_empty() }
        fn optimize(&mut self) -> Result<(), VectorDbError> { Ok(()) }
        fn clear(&mut self) { self.inner = None; self.id_to_index.clear(); self.index_to_id.clear(); }
        fn get_stats(&self) -> IndexStats {
            let mut s = gvdb_stats::default();
            if let Some(ix) = &self.inner { unsafe { gvdb_get_stats(ix.h, &mut s) }; }
            IndexStats { vector_count: self.len(), dimension: s.dimension as usize,
                         index_type: if self.exact { "GpuFlat".into() } else { "GpuBinaryTwoStage".into() },
                         memory_usage: s.memory_usage as usize }
        }
    }
}
#[cfg(feature = "vector-index")]
pub use trait_impl::GpuVectorIndex;

/// `BinaryQuantizer` with the reference's signatures (src/quantization.rs:67-241), arithmetic on the GPU.
/// In the reference's crate this replaces the body of `impl BinaryQuantizer`; `BinaryVector`,
/// `BinaryQuantizationConfig`, `CacheStats` and the error type stay the reference's own.
#[cfg(feature = "vector-index")]
mod quantizer_shim {
    use super::*;
    use grape_vector_db::quantization::{BinaryQuantizationConfig, BinaryVector, CacheStats};
    use grape_vector_db::types::VectorDbError;
    use std::sync::Mutex;

    /// The candidate set the last `multi_stage_search` uploaded: the reference's signature hands the
    /// whole candidate set in on every call, so the index built from it is kept and reused while the
    /// caller keeps passing the same slice (same address, length and dimension, same first/last row).
    struct Resident { ptr: usize, n: usize, dim: usize, head: Vec<f32>, tail: Vec<f32>, index: GpuIndex }

    pub struct BinaryQuantizer {
        config: BinaryQuantizationConfig,
        device: i32,
        /// one small index per code width, for `quantize*` / `hamming_distance` (created on first use)
        scratch: Mutex<HashMap<usize, GpuIndex>>,
        resident: Mutex<Option<Resident>>,
    }

    /// status codes of include/gvdb.h -> the reference's error variants (src/types.rs:859-920)
    fn err(st: i32, expected_dim: usize, actual_dim: usize) -> VectorDbError {
        match st {
            1 => VectorDbError::IndexNotBuilt,
            2 => VectorDbError::DimensionMismatch { expected: expected_dim, actual: actual_dim },
            3 => VectorDbError::InvalidVectorDimension,
            4 => VectorDbError::QuantizationError(last_error()),
            6 => VectorDbError::ConfigError(last_error()),
            7 => VectorDbError::NotImplemented(last_error()),
            _ => VectorDbError::IndexError(last_error()),
        }
    }

    fn create(dim: usize, cfg: &BinaryQuantizationConfig, device: i32, capacity: u64) -> Result<GpuIndex, VectorDbError> {
        let c = gvdb_config { struct_size: std::mem::size_of::<gvdb_config>() as u32, dim: dim as u32, threshold: cfg.threshold,
                              rescore_ratio: cfg.rescore_ratio, device, capacity_rows: capacity, ..Default::default() };
        let mut h = std::ptr::null_mut();
        let st = unsafe { gvdb_create(&c, &mut h) };
        if st != 0 { return Err(err(st, dim, dim)); }
        Ok(GpuIndex { h, dim })
    }

    impl BinaryQuantizer {
        pub fn new(config: BinaryQuantizationConfig) -> Self {
            Self { config, device: 0, scratch: Mutex::new(HashMap::new()), resident: Mutex::new(None) }
        }

        fn with_scratch<T>(&self, dim: usize, f: impl FnOnce(&GpuIndex) -> Result<T, VectorDbError>) -> Result<T, VectorDbError> {
            let mut map = self.scratch.lock().unwrap();
            if !map.contains_key(&dim) {
                let ix = create(dim, &self.config, self.device, 0)?;
                map.insert(dim, ix);
            }
            f(&map[&dim])
        }

        /// :86-122 — bit = value > threshold, Msb0 bytes; the cache of the reference is not needed
        pub fn quantize(&mut self, vector: &[f32]) -> Result<BinaryVector, VectorDbError> {
            Ok(self.quantize_batch(&[vector.to_vec()])?.pop().unwrap())
        }

        /// :125-127 — one GPU call for a batch of equal dimension
        pub fn quantize_batch(&mut self, vectors: &[Vec<f32>]) -> Result<Vec<BinaryVector>, VectorDbError> {
            let Some(first) = vectors.first() else { return Ok(Vec::new()) };
            let dim = first.len();
            if dim == 0 || vectors.iter().any(|v| v.len() != dim) { return Err(VectorDbError::InvalidVectorDimension); }
            let flat: Vec<f32> = vectors.iter().flatten().copied().collect();
            let nb = (dim + 7) / 8;
            let mut codes = vec![0u8; vectors.len() * nb];
            self.with_scratch(dim, |ix| {
                let st = unsafe { gvdb_quantize(ix.h, flat.as_ptr(), vectors.len() as u64, codes.as_mut_ptr()) };
                if st != 0 { Err(err(st, dim, dim)) } else { Ok(()) }
            })?;
            // BinaryVector::from_bytes(bytes, dimension)  (:59-62) makes a BitVec of 8 * bytes bits, while the
            // reference's quantize pushes exactly `dimension` bits (:97-101) and its own test asserts
            // `data.len() == 5` for a 5-d vector (:369): the pad bits (zero) are cut off again
            Ok(codes.chunks(nb).map(|c| {
                let mut v = BinaryVector::from_bytes(c.to_vec(), dim);
                v.data.truncate(dim);
                v
            }).collect())
        }

        /// :130-141 — popcount(a XOR b) over the code bytes as f32; InvalidVectorDimension if the dimensions differ
        pub fn hamming_distance(&self, a: &BinaryVector, b: &BinaryVector) -> Result<f32, VectorDbError> {
            if a.dimension != b.dimension { return Err(VectorDbError::InvalidVectorDimension); }
            // two codes: the host does it (the GPU path is gvdb_hamming / the scan, used by the searches)
            let (ab, bb) = (a.to_bytes(), b.to_bytes());
            Ok(ab.iter().zip(bb.iter()).map(|(x, y)| (x ^ y).count_ones()).sum::<u32>() as f32)
        }

        /// :144-148 — 1 - distance / dimension
        pub fn similarity(&self, a: &BinaryVector, b: &BinaryVector) -> Result<f32, VectorDbError> {
            let distance = self.hamming_distance(a, b)?;
            let max_distance = a.dimension as f32;
            Ok(1.0 - (distance / max_distance))
        }

        /// :151-193 — stage 1 over every candidate, R = (n as f32 * rescore_ratio) as usize, exact rescoring of
        /// the R kept, both stable sorts; returns all R (index, cosine) pairs like the reference
        pub fn multi_stage_search(&self, _query_binary: &BinaryVector, candidates_binary: &[BinaryVector],
                                  original_query: &[f32], original_candidates: &[Vec<f32>])
                                  -> Result<Vec<(usize, f32)>, VectorDbError> {
            if candidates_binary.len() != original_candidates.len() {
                return Err(VectorDbError::QuantizationError("Mismatch between binary and original candidate counts".into()));
            }
            let n = original_candidates.len();
            let r = unsafe { gvdb_rescore_count(n as u64, self.config.rescore_ratio) } as usize;
            if n == 0 || r == 0 { return Ok(Vec::new()); }
            let dim = original_query.len();
            if original_candidates.iter().any(|v| v.len() != dim) { return Err(VectorDbError::InvalidVectorDimension); }
            let mut slot = self.resident.lock().unwrap();
            let same = slot.as_ref().map_or(false, |c| {
                c.ptr == original_candidates.as_ptr() as usize && c.n == n && c.dim == dim
                    && c.head == original_candidates[0] && c.tail == original_candidates[n - 1]
            });
            if !same {
                let ix = create(dim, &self.config, self.device, n as u64)?;
                let flat: Vec<f32> = original_candidates.iter().flatten().copied().collect();
                let mut first = 0u64;
                let st = unsafe { gvdb_add(ix.h, flat.as_ptr(), n as u64, &mut first) };
                if st != 0 { return Err(err(st, dim, dim)); }
                *slot = Some(Resident { ptr: original_candidates.as_ptr() as usize, n, dim,
                                        head: original_candidates[0].clone(), tail: original_candidates[n - 1].clone(), index: ix });
            }
            let ix = &slot.as_ref().unwrap().index;
            let (mut ids, mut sc) = (vec![GVDB_NO_ID; r], vec![0f32; r]);
            let st = unsafe { gvdb_search_batch(ix.h, original_query.as_ptr(), 1, r as u32, r as u32, ids.as_mut_ptr(),
                                                sc.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) };
            if st != 0 { return Err(err(st, dim, dim)); }
            Ok(ids.iter().zip(sc.iter()).take_while(|(i, _)| **i != GVDB_NO_ID).map(|(i, s)| (*i as usize, *s)).collect())
        }

        /// :219-223 — nothing is cached on this side; the resident candidate index is dropped
        pub fn clear_cache(&mut self) { *self.resident.lock().unwrap() = None; }

        /// :226-240 — the quantization cache of the reference does not exist here
        pub fn get_cache_stats(&self) -> CacheStats { CacheStats { enabled: false, size: 0, capacity: 0 } }
    }
}
#[cfg(feature = "vector-index")]
pub use quantizer_shim::BinaryQuantizer;
