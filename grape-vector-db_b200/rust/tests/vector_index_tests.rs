//! `impl VectorIndex for GpuVectorIndex` (reference trait: src/index.rs:35-62) behind the reference's own
//! error variants and score convention — the Rust twin of tests/cpp/test_host_mirror.cpp::test_vector_index_trait,
//! which IS compiled and run on a B200 in this repository.  Needs a B200 and `--features vector-index`;
//! NOT compiled in this repository's image (no cargo).
#![cfg(feature = "vector-index")]
use grape_vector_db::index::VectorIndex;
use grape_vector_db::types::VectorDbError;
use gvdb_sys::GpuVectorIndex;

fn exercise(mut index: GpuVectorIndex) {
    // search before any data: IndexNotBuilt (src/index.rs:621-623)
    assert!(matches!(index.search(&[1.0, 0.0, 0.0], 1), Err(VectorDbError::IndexNotBuilt)));
    index.add_vector("doc1".to_string(), vec![1.0, 0.0, 0.0]).unwrap();
    index.add_vectors(vec![("doc2".to_string(), vec![0.9, 0.1, 0.0]),
                           ("doc3".to_string(), vec![-1.0, 0.0, 0.0]),
                           ("doc4".to_string(), vec![0.0, 1.0, 0.0])]).unwrap();
    assert_eq!(index.len(), 4);
    assert!(!index.is_empty());
    // a vector of another length: DimensionMismatch with both numbers (src/index.rs:590-594)
    match index.add_vector("bad".to_string(), vec![1.0]) {
        Err(VectorDbError::DimensionMismatch { expected, actual }) => assert_eq!((expected, actual), (3, 1)),
        other => panic!("expected DimensionMismatch, got {:?}", other.map(|_| ())),
    }
    assert!(matches!(index.search(&[1.0], 1), Err(VectorDbError::DimensionMismatch { expected: 3, actual: 1 })));
    // best first, distance = 1 - cos ascending (src/index.rs:630-637,699)
    let res = index.search(&[1.0, 0.0, 0.0], 2).unwrap();
    assert_eq!(res.iter().map(|r| r.0.as_str()).collect::<Vec<_>>(), vec!["doc1", "doc2"]);
    assert_eq!(res[0].1, 0.0);
    // removal is a tombstone (src/index.rs:642-650): the id is gone, a second removal reports false
    assert!(index.remove_vector("doc1").unwrap());
    assert!(!index.remove_vector("doc1").unwrap());
    let res = index.search(&[1.0, 0.0, 0.0], 1).unwrap();
    assert_eq!(res.len(), 1);
    assert_eq!(res[0].0, "doc2");
    // re-adding a live id replaces its vector: one entry per id
    index.add_vector("doc2".to_string(), vec![0.0, 0.0, 1.0]).unwrap();
    assert_eq!(index.len(), 3);
    let res = index.search(&[0.0, 0.0, 1.0], 1).unwrap();
    assert_eq!(res[0].0, "doc2");
    let stats = index.get_stats();
    assert_eq!((stats.vector_count, stats.dimension), (3, 3));
    // more answers asked for than live rows: only the live rows come back
    assert_eq!(index.search(&[1.0, 0.0, 0.0], 10).unwrap().len(), 3);
    index.clear();
    assert!(index.is_empty());
    assert!(matches!(index.search(&[1.0, 0.0, 0.0], 1), Err(VectorDbError::IndexNotBuilt)));
}

#[test]
fn exact_flat_search_behind_the_trait() {
    exercise(GpuVectorIndex::new(true, 4, 0));
}

#[test]
fn two_stage_search_behind_the_trait() {
    exercise(GpuVectorIndex::new(false, 4, 0));
}

#[test]
fn usable_as_a_trait_object_across_threads() {
    // VectorIndex: Send + Sync (src/index.rs:35); VectorDatabase holds Arc<RwLock<dyn VectorIndex>> (src/lib.rs:238)
    let mut index = GpuVectorIndex::new(true, 4, 0);
    index.add_vectors((0..64).map(|i| (format!("d{i}"), vec![i as f32, 1.0, 64.0 - i as f32])).collect()).unwrap();
    let shared: std::sync::Arc<std::sync::RwLock<dyn VectorIndex>> = std::sync::Arc::new(std::sync::RwLock::new(index));
    let handles: Vec<_> = (0..4).map(|t| {
        let ix = shared.clone();
        std::thread::spawn(move || ix.read().unwrap().search(&[t as f32 * 16.0, 1.0, 64.0 - t as f32 * 16.0], 1).unwrap()[0].0.clone())
    }).collect();
    let got: Vec<String> = handles.into_iter().map(|h| h.join().unwrap()).collect();
    assert_eq!(got, vec!["d0", "d16", "d32", "d48"]);
}
