//! The reference's own quantizer tests (src/quantization.rs:359-400), run against the GPU-backed
//! `BinaryQuantizer` of this crate, with the values those tests only describe in comments pinned
//! (bits 10101 = 0xA8; 0xA0 / 0xC0, Hamming 2, similarity 0.5; 0xE0).
//! Needs a B200 and `--features vector-index`; NOT compiled in this repository's image (no cargo).
#![cfg(feature = "vector-index")]
use grape_vector_db::quantization::{BinaryQuantizationConfig, BinaryVectorStore};
use gvdb_sys::BinaryQuantizer;

#[test]
fn test_binary_quantization() {
    let mut quantizer = BinaryQuantizer::new(BinaryQuantizationConfig::default());
    let vector = vec![0.5, -0.3, 0.8, -0.1, 0.2];
    let binary_vec = quantizer.quantize(&vector).unwrap();
    assert_eq!(binary_vec.dimension, 5);
    assert_eq!(binary_vec.data.len(), 5);               // the reference's own assertion (:369)
    assert_eq!(binary_vec.to_bytes(), vec![0xA8u8]);    // [1, 0, 1, 0, 1], Msb0, pad bits zero
}

#[test]
fn test_hamming_distance() {
    let mut quantizer = BinaryQuantizer::new(BinaryQuantizationConfig::default());
    let bin1 = quantizer.quantize(&[1.0, -1.0, 1.0, -1.0]).unwrap();
    let bin2 = quantizer.quantize(&[1.0, 1.0, -1.0, -1.0]).unwrap();
    assert_eq!(bin1.to_bytes(), vec![0xA0u8]);
    assert_eq!(bin2.to_bytes(), vec![0xC0u8]);
    let distance = quantizer.hamming_distance(&bin1, &bin2).unwrap();
    assert!(distance > 0.0);                            // the reference's own assertion (:385)
    assert_eq!(distance, 2.0);
    assert_eq!(quantizer.similarity(&bin1, &bin2).unwrap(), 0.5);
}

#[test]
fn test_binary_vector_store() {
    let mut store = BinaryVectorStore::new(BinaryQuantizationConfig::default());
    let mut quantizer = BinaryQuantizer::new(BinaryQuantizationConfig::default());
    let binary_vec = quantizer.quantize(&[0.1, 0.2, 0.3]).unwrap();
    assert_eq!(binary_vec.to_bytes(), vec![0xE0u8]);
    store.add_vector(binary_vec, "test_id".to_string()).unwrap();
    assert_eq!(store.len(), 1);
    assert!(!store.is_empty());
}

#[test]
fn strict_threshold_negative_zero_and_nan_quantize_to_zero() {
    let mut quantizer = BinaryQuantizer::new(BinaryQuantizationConfig::default());
    let v = quantizer.quantize(&[0.0, -0.0, f32::NAN, 1e-45, f32::INFINITY, f32::NEG_INFINITY, 1.0, -1.0]).unwrap();
    assert_eq!(v.to_bytes(), vec![0b0001_1010u8]);      // value > threshold, strictly (:99)
}

#[test]
fn multi_stage_search_orders_like_the_reference() {
    // 10 candidates, ratio 0.5 -> R = 5 (f32 product, truncation, :178-179); stage 1 ties break by index
    let mut cfg = BinaryQuantizationConfig::default();
    cfg.rescore_ratio = 0.5;
    let mut quantizer = BinaryQuantizer::new(cfg);
    let cands: Vec<Vec<f32>> = (0..10).map(|i| {
        let s = if i % 2 == 0 { 1.0 } else { -1.0 };
        vec![s * (1.0 + i as f32), 1.0, -1.0, s]
    }).collect();
    let query = vec![1.0f32, 1.0, -1.0, 1.0];
    let qb = quantizer.quantize(&query).unwrap();
    let cb = quantizer.quantize_batch(&cands).unwrap();
    let out = quantizer.multi_stage_search(&qb, &cb, &query, &cands).unwrap();
    assert_eq!(out.len(), 5);
    // the even candidates have Hamming distance 0 to the query, the odd ones 2: stage 1 keeps 0, 2, 4, 6, 8
    let mut ids: Vec<usize> = out.iter().map(|p| p.0).collect();
    ids.sort();
    assert_eq!(ids, vec![0, 2, 4, 6, 8]);
    assert!(out.windows(2).all(|w| w[0].1 >= w[1].1));   // cosine descending (:190)
    assert_eq!(out[0].0, 0);                             // [1, 1, -1, 1] is the query itself: cosine 1
}
