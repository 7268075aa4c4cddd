"""The C++ host mirror (include/hannoy_b200.hpp) over the C-ABI: compiled with g++ against libhannoy_b200.so.
CPU part: Reader::open error paths (reader.rs:390-416).  GPU part: open from raw KV pairs, batched search,
by_item, dimension check — results compared bit for bit with the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import make_db, make_vectors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hannoy_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "host_mirror_test")


@pytest.fixture(scope="module")
def exe():
    from hannoy_b200 import build as hb_build
    hb_build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(os.path.join(ROOT, "include", "hannoy_b200.hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR,
                               "-lhannoy_b200", f"-Wl,-rpath,{LIBDIR}", "-o", EXE])
    return EXE


def test_cpp_mirror_open_errors(exe):
    out = subprocess.run([exe, "cpu"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "OK cpu" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("metric,mid", [("euclidean", 0), ("cosine", 1), ("hamming", 3), ("binary quantized cosine", 4)])
def test_cpp_mirror_search_matches_oracle(exe, tmp_path, metric, mid):
    n, dims, nq, k, ef, index = 600, 72, 40, 7, 33, 5
    db, x = make_db(metric, n, dims, seed=mid + 3)
    q = make_vectors(nq, dims, seed=99)
    ids, dist, lens, _ = db.search_by_vector(q, k, ef=ef)
    kv = list(db.export_kv(index))
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<III", mid, index, len(kv)))
        for key, val in kv:
            f.write(struct.pack("<I", len(key)) + bytes(key))
            f.write(struct.pack("<I", len(val)) + bytes(val))
        f.write(struct.pack("<IIII", dims, nq, k, ef))
        f.write(np.ascontiguousarray(q, np.float32).tobytes())
        f.write(np.ascontiguousarray(lens, np.uint32).tobytes())
        f.write(np.ascontiguousarray(ids, np.uint32).tobytes())
        f.write(np.ascontiguousarray(dist, np.float32).tobytes())
    out = subprocess.run([exe, "gpu", str(path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
