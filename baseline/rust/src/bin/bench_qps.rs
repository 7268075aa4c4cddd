//! The CPU baseline of bench.py on the real crate: batched `reader.nns(k).ef_search(ef).by_vector(..)` over all host cores.
//!
//! One `Reader` (it is `Send + Sync`), one read transaction per rayon thread — the pattern hannoy itself uses for its
//! parallel build (src/parallel.rs:18-38).  Prints one JSON line: queries/s, threads, mean results per query.
//!
//!   bench_qps --vectors x.f32 --queries q.f32 --dims 768 [--k 10] [--ef 128] [--m 16] [--efc 100] [--reps 5] [--db DIR]
//!
//! `x.f32` / `q.f32` are raw row-major f32 files; `python tools/dump_workload.py c3 x.f32 q.f32` writes the vectors
//! bench.py generates, so the two arms see the same data.  Metric: Cosine (config 3) — change the `D` alias for others.
use std::path::PathBuf;
use std::time::Instant;

use hannoy::distances::Cosine;
use hannoy::{Database, Reader, Writer};
use hannoy_b200_baseline::{open_env, read_f32_matrix, Args};
use rand::rngs::StdRng;
use rand::SeedableRng;
use rayon::prelude::*;

type D = Cosine;
const M: usize = 16;
const M0: usize = 32;

fn main() {
    let args = Args::parse();
    let dims: usize = args.num("--dims", 768);
    let k: usize = args.num("--k", 10);
    let ef: usize = args.num("--ef", 128);
    let efc: usize = args.num("--efc", 100);
    let reps: usize = args.num("--reps", 5);
    let vectors = read_f32_matrix(&PathBuf::from(args.get("--vectors").expect("--vectors")), dims).unwrap();
    let queries = read_f32_matrix(&PathBuf::from(args.get("--queries").expect("--queries")), dims).unwrap();
    let dir = args.get("--db").map(PathBuf::from).unwrap_or_else(|| tempfile::tempdir().unwrap().into_path());
    let env = open_env(&dir, 64);

    let mut wtxn = env.write_txn().unwrap();
    let db: Database<D> = env.create_database(&mut wtxn, None).unwrap();
    let t = Instant::now();
    if db.is_empty(&wtxn).unwrap() {
        let writer = Writer::<D>::new(db, 0, dims);
        for (id, v) in vectors.iter().enumerate() {
            writer.add_item(&mut wtxn, id as u32, v).unwrap();
        }
        let mut rng = StdRng::seed_from_u64(42);
        writer.builder(&mut rng).ef_construction(efc).build::<M, M0>(&mut wtxn).unwrap();
        eprintln!("[bench_qps] built {} items in {:.1}s", vectors.len(), t.elapsed().as_secs_f64());
    }
    wtxn.commit().unwrap();

    let rtxn0 = env.read_txn().unwrap();
    let reader = Reader::<D>::open(&rtxn0, 0, db).unwrap();
    let threads = rayon::current_num_threads();
    // warm-up + timed repetitions; every rayon task opens (and drops) its own read transaction: LMDB readers are cheap
    let run = || -> usize {
        queries
            .par_iter()
            .map_init(
                || env.read_txn().unwrap(),
                |rtxn, q| reader.nns(k).ef_search(ef).by_vector(rtxn, q).unwrap().into_nns().len(),
            )
            .sum()
    };
    let _ = run();
    let mut best = f64::MAX;
    let mut found = 0usize;
    for _ in 0..reps {
        let t = Instant::now();
        found = run();
        best = best.min(t.elapsed().as_secs_f64());
    }
    println!(
        "{{\"impl\": \"hannoy {}\", \"metric\": \"QPS (batched, rayon par_iter)\", \"value\": {:.1}, \"unit\": \"queries/s\", \"threads\": {}, \"queries\": {}, \"k\": {}, \"ef_search\": {}, \"mean_results\": {:.2}}}",
        "0.1.3", queries.len() as f64 / best, threads, queries.len(), k, ef, found as f64 / queries.len() as f64
    );
}
