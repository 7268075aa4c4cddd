//! Reference golden: a seeded index built by the REAL `Writer`, searched by the REAL `Reader`, dumped for
//! tests/test_reference_golden.py.  Output directory (default ../../tests/golden/ref_reader):
//!
//!   meta.json     {"metric", "dims", "n", "nq", "k", "ef", "index", "arch", "hannoy"}
//!   kv.bin        every LMDB pair of the index in key order: [u32 klen][key][u32 vlen][value] (little-endian lengths)
//!   queries.f32   nq x dims raw f32
//!   results.bin   per query: [u32 n][n x u32 ids][n x f32 distances]   (by_vector, then the same block for by_item on the
//!                 items listed in items.u32; n = u32::MAX encodes `None`)
//!   items.u32     the by_item queries
//!
//! The Python test feeds kv.bin to the oracle and to the CUDA engine (Reader.open over raw pairs) and compares answers:
//! ids exactly, distances bit for bit when `arch` reports x86_64 with AVX (the summation order the oracle restates).
use std::fs::{self, File};
use std::io::{BufWriter, Write};
use std::path::PathBuf;

use hannoy::distances::{BinaryQuantizedCosine, Cosine, Euclidean, Hamming, Manhattan};
use hannoy::{Database, Distance, Reader, Writer};
use hannoy_b200_baseline::{open_env, Args};
use heed::types::Bytes;
use rand::rngs::StdRng;
use rand::{Rng, SeedableRng};

const M: usize = 16;
const M0: usize = 32;

fn dump<Dist: Distance>(out: &PathBuf, name: &str, n: usize, dims: usize, nq: usize, k: usize, ef: usize, seed: u64) {
    let dir = out.join(name);
    fs::create_dir_all(&dir).unwrap();
    let env = open_env(&dir.join("lmdb"), 4);
    let mut rng = StdRng::seed_from_u64(seed);
    // clustered vectors: 32 centres + noise (i.i.d. uniform data is adversarial for any graph index)
    let centres: Vec<Vec<f32>> = (0..32).map(|_| (0..dims).map(|_| rng.gen_range(-1.0f32..1.0)).collect()).collect();
    let mut gen = |rng: &mut StdRng| -> Vec<f32> {
        let c = &centres[rng.gen_range(0..centres.len())];
        c.iter().map(|v| v + 0.3 * rng.gen_range(-1.0f32..1.0)).collect()
    };
    let index: u16 = 7;
    let mut wtxn = env.write_txn().unwrap();
    let db: Database<Dist> = env.create_database(&mut wtxn, None).unwrap();
    let writer = Writer::<Dist>::new(db, index, dims);
    for id in 0..n {
        let v = gen(&mut rng);
        writer.add_item(&mut wtxn, (id * 3 + 1) as u32, &v).unwrap(); // sparse ids
    }
    let mut brng = StdRng::seed_from_u64(seed + 1);
    writer.builder(&mut brng).ef_construction(100).build::<M, M0>(&mut wtxn).unwrap();
    wtxn.commit().unwrap();

    let rtxn = env.read_txn().unwrap();
    // raw pairs, exactly as a heed cursor yields them
    let mut kv = BufWriter::new(File::create(dir.join("kv.bin")).unwrap());
    let raw = db.remap_types::<Bytes, Bytes>();
    for pair in raw.prefix_iter(&rtxn, &index.to_be_bytes()).unwrap() {
        let (key, val) = pair.unwrap();
        kv.write_all(&(key.len() as u32).to_le_bytes()).unwrap();
        kv.write_all(key).unwrap();
        kv.write_all(&(val.len() as u32).to_le_bytes()).unwrap();
        kv.write_all(val).unwrap();
    }
    kv.flush().unwrap();

    let reader = Reader::<Dist>::open(&rtxn, index, db).unwrap();
    let queries: Vec<Vec<f32>> = (0..nq).map(|_| gen(&mut rng)).collect();
    let mut qf = BufWriter::new(File::create(dir.join("queries.f32")).unwrap());
    for q in &queries {
        for v in q {
            qf.write_all(&v.to_le_bytes()).unwrap();
        }
    }
    let items: Vec<u32> = (0..64u32).map(|i| if i % 16 == 15 { 2 } else { (i * 37 % n as u32) * 3 + 1 }).collect(); // id 2 is absent
    let mut itf = BufWriter::new(File::create(dir.join("items.u32")).unwrap());
    for i in &items {
        itf.write_all(&i.to_le_bytes()).unwrap();
    }
    let mut rf = BufWriter::new(File::create(dir.join("results.bin")).unwrap());
    let mut put = |nns: Option<Vec<(u32, f32)>>| match nns {
        None => rf.write_all(&u32::MAX.to_le_bytes()).unwrap(),
        Some(v) => {
            rf.write_all(&(v.len() as u32).to_le_bytes()).unwrap();
            for (id, _) in &v {
                rf.write_all(&id.to_le_bytes()).unwrap();
            }
            for (_, d) in &v {
                rf.write_all(&d.to_le_bytes()).unwrap();
            }
        }
    };
    for q in &queries {
        put(Some(reader.nns(k).ef_search(ef).by_vector(&rtxn, q).unwrap().into_nns()));
    }
    for &i in &items {
        put(reader.nns(k).ef_search(ef).by_item(&rtxn, i).unwrap().map(|s| s.into_nns()));
    }
    let arch = format!(
        "{}{}",
        std::env::consts::ARCH,
        if cfg!(target_arch = "x86_64") && std::is_x86_feature_detected!("avx") && std::is_x86_feature_detected!("fma") { "+avx+fma" } else { "" }
    );
    fs::write(
        dir.join("meta.json"),
        format!(
            "{{\"metric\": \"{}\", \"dims\": {}, \"n\": {}, \"nq\": {}, \"k\": {}, \"ef\": {}, \"index\": {}, \"arch\": \"{}\", \"hannoy\": \"0.1.3\"}}\n",
            Dist::name(), dims, n, nq, k, ef, index, arch
        ),
    )
    .unwrap();
    eprintln!("[gen_golden] {name}: {n} x {dims} {}, {nq} queries -> {}", Dist::name(), dir.display());
}

fn main() {
    let args = Args::parse();
    let out = PathBuf::from(args.get("--out").unwrap_or("../../tests/golden/ref_reader"));
    let n: usize = args.num("--n", 10_000);
    let nq: usize = args.num("--nq", 1_000);
    dump::<Euclidean>(&out, "euclidean_128", n, 128, nq, 10, 64, 1);
    dump::<Cosine>(&out, "cosine_100", n, 100, nq, 10, 64, 2);
    dump::<Manhattan>(&out, "manhattan_24", n / 2, 24, nq / 2, 10, 64, 3);
    dump::<Hamming>(&out, "hamming_256", n / 2, 256, nq / 2, 10, 64, 4);
    dump::<BinaryQuantizedCosine>(&out, "bq_cosine_1024", n / 2, 1024, nq / 2, 100, 200, 5);
}
