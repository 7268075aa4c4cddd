"""Builds libhannoy_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

`python -m hannoy_b200.build [--force] [-v] [--variant NAME]`; variants are dev builds written to
libhannoy_b200_NAME.so and loaded instead of the product library when HB_LIB_VARIANT=NAME: `phases` (-DHB_PHASES:
per-phase cycle counters in the search kernel), `trace` (-DHB_TRACE: event trace of one warp), `rg2` (rows reduced
two at a time instead of four).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["capi.cu", "search.cu", "exact.cu", "exact_tc.cu", "build.cu", "snapshot.cpp", "lmdb_walk.cpp"]
HEADERS = ["common.h", "dist.cuh", "sorted.cuh", "ring.cuh", "stage.cuh", os.path.join("..", "..", "include", "hannoy_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exactness: no implicit FMA contraction, IEEE div/sqrt, no flush-to-zero
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    # (no -split-compile: with it ptxas produced kernels of 8 296 or 9 144 instructions from the SAME source from one run to the
    # next — and the larger one is ~10 % slower; a single-threaded compile of search.cu takes 26 s and is reproducible)
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]
VARIANTS = {"": [], "phases": ["-DHB_PHASES"], "trace": ["-DHB_TRACE"], "rg2": ["-DHB_ROW_GROUP=2"], "rg8": ["-DHB_ROW_GROUP=8"],
            "late": ["-DHB_EARLY_ROWS=0"], "adjtop": ["-DHB_ADJ_PREFETCH_ALL=0"]}
for _b in (4, 5, 6, 7, 8):   # dev: occupancy the binary / f32 ring kernels are compiled for, heap-merge block size
    VARIANTS[f"bin{_b}"] = [f"-DHB_MIN_BLOCKS_BIN={_b}"]
    VARIANTS[f"g2bin{_b}"] = [f"-DHB_MIN_BLOCKS_BIN={_b}", "-DHB_MERGE_BLOCK=2"]
    VARIANTS[f"f32b{_b}"] = [f"-DHB_MIN_BLOCKS_F32={_b}"]
    VARIANTS[f"g2f32b{_b}"] = [f"-DHB_MIN_BLOCKS_F32={_b}", "-DHB_MERGE_BLOCK=2"]
VARIANTS["nospec"] = ["-DHB_SPEC_VIS=0"]
VARIANTS["nospecbin"] = ["-DHB_SPEC_VIS_BIN=0"]
VARIANTS["nodedupe"] = ["-DHB_SPEC_DEDUPE=0"]
VARIANTS["dedupef32"] = ["-DHB_SPEC_DEDUPE_F32=1"]
VARIANTS["keep"] = ["-DHB_UPPER_KEEP=1"]
VARIANTS["opaque"] = ["-DHB_OPAQUE_ADDR=1"]
VARIANTS["nolean"] = ["-DHB_LEAN=0"]
VARIANTS["visatom"] = ["-DHB_VIS_LOAD=0"]
VARIANTS["share"] = ["-DHB_TEAM_SHARE=1"]
VARIANTS["f32s5"] = ["-DHB_MIN_BLOCKS_F32_SHORT=5"]
for _g in (1, 2, 3, 4, 8):
    VARIANTS[f"g{_g}"] = [f"-DHB_MERGE_BLOCK={_g}"]


def out_path(variant=""):
    return os.path.join(HERE, f"libhannoy_b200{'_' + variant if variant else ''}.so")


def needs_build(variant=""):
    out = out_path(variant)
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, variant=""):
    out = out_path(variant)
    if not force and not needs_build(variant):
        return out
    objdir = os.path.join(HERE, "build", variant or "product")
    os.makedirs(objdir, exist_ok=True)
    flags = FLAGS + VARIANTS[variant] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        subprocess.check_call([NVCC] + flags + ["-c", os.path.join(CSRC, src), "-o", obj])
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"] + objs + ["-o", out])
    return out


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant))
