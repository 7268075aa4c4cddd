//! Shared helpers of the two binaries: raw f32 matrix files, LMDB set-up, argument parsing.
use std::fs::File;
use std::io::{BufReader, Read};
use std::path::Path;

use heed::{Env, EnvOpenOptions, WithoutTls};

/// A row-major `rows x dims` f32 matrix stored as raw little-endian bytes (what `numpy.ndarray.tofile` writes).
pub fn read_f32_matrix(path: &Path, dims: usize) -> std::io::Result<Vec<Vec<f32>>> {
    let mut bytes = Vec::new();
    BufReader::new(File::open(path)?).read_to_end(&mut bytes)?;
    assert!(bytes.len() % (4 * dims) == 0, "{}: not a multiple of {} floats", path.display(), dims);
    Ok(bytes
        .chunks_exact(4 * dims)
        .map(|row| row.chunks_exact(4).map(|b| f32::from_le_bytes([b[0], b[1], b[2], b[3]])).collect())
        .collect())
}

/// An LMDB environment without thread-local storage for readers, so that read transactions can be handed to the threads
/// of a rayon pool (hannoy does the same for its own parallel build, src/parallel.rs:18-38).
pub fn open_env(dir: &Path, map_size_gib: usize) -> Env<WithoutTls> {
    std::fs::create_dir_all(dir).unwrap();
    unsafe { EnvOpenOptions::new().read_txn_without_tls().map_size(map_size_gib << 30).max_dbs(4).open(dir) }.unwrap()
}

/// `--name value` pairs.
pub struct Args(Vec<String>);
impl Args {
    pub fn parse() -> Self {
        Args(std::env::args().skip(1).collect())
    }
    pub fn get(&self, name: &str) -> Option<&str> {
        self.0.iter().position(|a| a == name).and_then(|i| self.0.get(i + 1)).map(|s| s.as_str())
    }
    pub fn num<T: std::str::FromStr>(&self, name: &str, default: T) -> T {
        self.get(name).and_then(|v| v.parse().ok()).unwrap_or(default)
    }
}
