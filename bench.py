#!/usr/bin/env python
"""bench.py — batched HNSW search throughput (BASELINE.json: QPS at recall@10 >= 0.95, batched).

A "step" is one pass of the hot path over ONE batch of synthetic queries against an HBM-resident index.  Default
workload = BASELINE configs[2]: 1M x 768 f32 Cosine, one 10k-query batch, top-10, smallest ef_search in
{32,64,128,256} with recall@10 >= 0.95 (measured against the exact k-NN kernel).  Multi-GPU (`--gpus N`, one process
per GPU under torchrun): the graph is replicated and the SAME 10k batch is partitioned into contiguous nq/N slices, one
per rank, no collective in the data path — STRONG scaling: `value` = 10 000 queries / max-over-ranks step time.  The
weak-scaling figure (every rank searches a whole 10k batch) is reported beside it as `weak_scaling`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4s|c5s|c5] [--impl reference]

At N = 1 the line also carries `other_workloads` (configs 1, 2 and a 1M-item config 4: value, roofline, cpu_baseline,
parity; graphs built on the device) and `latency` (one query at a time, next to the CPU's); at N > 1 `sharded` holds
the id-sharded configuration (config 5 shaped: 6.25M x 768 items per shard, 100k queries, top-k exchange fused into
the search kernel vs NCCL all-gather).  HB_BENCH_EXTRAS=0 switches the extras off.

`--impl reference` times the reference's CPU algorithm (the oracle port under oracle/, all host threads) on the same
workload and the whole batch; the reference itself is Rust and cannot be built in this image.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: metric, n, dims, nq, k, ef candidates, generator
    "c1": dict(metric="euclidean", n=10_000, dims=128, nq=1_000, k=10, efs=[64], gen="uniform", seed=1,
               desc="10k x 128 f32 Euclidean, M=16/M0=32, efC=100, 1k queries top-10 ef_search=64"),
    "c2": dict(metric="euclidean", n=1_000_000, dims=128, nq=10_000, k=10, efs=[32, 64, 128, 256], gen="siftlike", seed=3,
               desc="SIFT-shaped 1M x 128 f32 Euclidean, 10k-query batch, top-10"),
    "c3": dict(metric="cosine", n=1_000_000, dims=768, nq=10_000, k=10, efs=[32, 64, 128, 256], gen="lowrank", seed=5,
               desc="1M x 768 f32 Cosine (text-embedding shaped), 10k-query batch, top-10"),
    "c4s": dict(metric="binary quantized cosine", n=1_000_000, dims=1024, nq=100_000, k=100, efs=[100, 200, 400], gen="lowrank", seed=7,
                desc="1M x 1024 BinaryQuantizedCosine codes (scaled-down config 4), 100k-query batch, top-100"),
    # config 4 at full size (tools/c4_full.py, tools/dev_sweep.py; not part of the default run: the 10M-item device build alone takes 22 s)
    "c4": dict(metric="binary quantized cosine", n=10_000_000, dims=1024, nq=100_000, k=100, efs=[100, 200, 400, 800], gen="lowrank", seed=7,
               desc="10M x 1024 BinaryQuantizedCosine codes (config 4), 100k-query batch, top-100"),
    # config 5, scaled: index sharded by item id (id % n_gpus), every GPU searches ALL queries on its shard, per-shard
    # top-k exchanged over NVLink and merged.  n = items PER SHARD.
    "c5s": dict(metric="cosine", n=250_000, dims=768, nq=20_000, k=10, efs=[128], gen="lowrank", seed=9, sharded=True,
                desc="id-sharded 768-d f32 Cosine, 250k items per shard/GPU, 20k-query batch searched by every shard, top-10, top-k exchange + merge"),
    # config 5 at its per-GPU shape: 50M x 768 over 8 shards = 6.25M items per shard (device-built graphs), 100k queries
    "c5": dict(metric="cosine", n=6_250_000, dims=768, nq=100_000, k=10, efs=[128], gen="lowrank", seed=9, sharded=True, device_build=True,
               desc="id-sharded 768-d f32 Cosine, 6.25M items per shard/GPU (config 5: 50M over 8 shards), 100k-query batch searched by every shard, top-10, top-k exchange + merge"),
}
M, M0, EFC, ALPHA = 16, 32, 100, 1.0
RECALL_TARGET = 0.95


def gen_vectors(gen, n, dims, seed, device):
    """Seeded synthetic vectors, generated with torch on `device`, returned as a float32 torch tensor."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    gb = torch.Generator(device=device)
    gb.manual_seed(1234)  # basis / cluster centres shared by base vectors and queries
    if gen == "uniform":
        return (torch.rand((n, dims), generator=g, device=device) * 2 - 1).float()
    if gen == "siftlike":  # SIFT-shaped: non-negative integers in [0, 255], ||x|| ~ 512, ~25 % zeros, low intrinsic dimension
        # (64 clusters in a 32-d latent space pushed through a random linear map, + small noise, clipped and rounded) --
        # i.i.d. noise around 64 centres in 128-d is adversarial for any graph index (recall@10 0.75 at ef = 256)
        r, nc = 32, 64
        A = torch.randn((r, dims), generator=gb, device=device) / (r ** 0.5)
        C = torch.randn((nc, r), generator=gb, device=device) * 1.5
        idx = torch.randint(0, nc, (n,), generator=g, device=device)
        z = C[idx] + torch.randn((n, r), generator=g, device=device)
        x = 27.0 + 21.5 * (z @ A) + 5.7 * torch.randn((n, dims), generator=g, device=device)
        return x.clamp_(0, 255).round_().float()
    if gen == "lowrank":  # embedding-shaped: 256 clusters in a 32-d latent space + small isotropic noise, unit norm
        r, nc = 32, 256
        A = torch.randn((r, dims), generator=gb, device=device) / (r ** 0.5)
        C = torch.randn((nc, r), generator=gb, device=device)
        out = torch.empty((n, dims), device=device, dtype=torch.float32)
        step = 200_000
        for s in range(0, n, step):
            m = min(step, n - s)
            idx = torch.randint(0, nc, (m,), generator=g, device=device)
            z = C[idx] + torch.randn((m, r), generator=g, device=device)
            x = z @ A + 0.02 * torch.randn((m, dims), generator=g, device=device)
            out[s:s + m] = x / x.norm(dim=1, keepdim=True)
        return out
    raise ValueError(gen)


def cache_dir(key):
    d = os.path.join(os.environ.get("HB_BENCH_CACHE", "/tmp/hannoy_b200_cache"), key)
    os.makedirs(d, exist_ok=True)
    return d


def build_or_load_graph(w, x_host, device_type, threads, log):
    """Graph built by the restated reference builder (oracle/), cached on local disk so the two bench arms
    and all ranks of one box share it.  Returns the OracleDb."""
    from oracle.oracle import OracleDb
    key = hashlib.sha1(json.dumps([w["metric"], w["n"], w["dims"], w["gen"], w["seed"], M, M0, EFC, ALPHA, device_type]).encode()).hexdigest()[:16]
    d = cache_dir(key)
    db = OracleDb(w["metric"], w["dims"])
    ids = np.arange(w["n"], dtype=np.uint32)
    db.add_items(ids, x_host)
    done = os.path.join(d, "done")
    if os.path.exists(done):
        t = time.time()
        meta = json.load(open(os.path.join(d, "meta.json")))
        from oracle import oracle as O
        for l in range(meta["n_layers"]):
            off = np.load(os.path.join(d, f"off{l}.npy"))
            nbr = np.load(os.path.join(d, f"nbr{l}.npy"))
            O.lib().orc_db_set_csr(db.h, l, O._p(off), O._p(nbr), len(nbr))
        db.set_entry_points(np.array(meta["eps"], np.uint32), meta["max_level"])
        log(f"graph loaded from cache in {time.time() - t:.1f}s")
    else:
        t = time.time()
        db.build(M=M, M0=M0, ef_construction=EFC, alpha=ALPHA, seed=42, n_threads=threads)
        log(f"graph built by the oracle builder in {time.time() - t:.1f}s on {threads} threads")
        layers = db.layers()
        for l, (off, nbr) in enumerate(layers):
            np.save(os.path.join(d, f"off{l}.npy"), off)
            np.save(os.path.join(d, f"nbr{l}.npy"), nbr)
        json.dump(dict(n_layers=len(layers), eps=[int(e) for e in db.entry_points], max_level=int(db.max_level)),
                  open(os.path.join(d, "meta.json"), "w"))
        open(done, "w").write("ok")
    return db


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", 1e30)
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.12] or [r for (_, r) in self.rows]
        for r in rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def recall_at_k(ids, lens, gt, k):
    hit = 0
    for i in range(len(ids)):
        hit += len(set(ids[i, :lens[i]].tolist()) & set(gt[i, :k].tolist()))
    return hit / (len(ids) * k)


def cpu_ground_truth(w, db, x_host, q, k, threads):
    """Exact top-k ids on the host for the reference arm's recall rule (no GPU code involved): blocked sgemm for
    the f32 metrics, the oracle's popcount scan for the binary ones."""
    if w["metric"] not in ("euclidean", "cosine"):
        return db.exact_knn(q, k, n_threads=threads)[0]
    import torch
    torch.set_num_threads(threads)
    tq = torch.from_numpy(np.ascontiguousarray(q))
    best_s = torch.full((len(q), k), -float("inf"))
    best_i = torch.zeros((len(q), k), dtype=torch.int64)
    chunk = 50_000
    for s in range(0, len(x_host), chunk):
        xb = torch.from_numpy(x_host[s:s + chunk])
        dots = tq @ xb.T
        if w["metric"] == "cosine":
            score = dots / xb.norm(dim=1).clamp_min(1e-30)       # the query norm is constant per row: same ranking
        else:
            score = 2 * dots - xb.pow(2).sum(1)                   # -(|x|^2 - 2 q.x): same ranking as squared L2
        ts, ti = score.topk(min(k, score.shape[1]), dim=1)
        cat_s, cat_i = torch.cat([best_s, ts], 1), torch.cat([best_i, ti + s], 1)
        best_s, sel = cat_s.topk(k, dim=1)
        best_i = cat_i.gather(1, sel)
    return best_i.numpy().astype(np.uint32)


N_GT = 2000  # queries the recall rule is evaluated on (both arms, same queries)


def algorithmic_bytes(ctr, w):
    """sum over layers of n_dist_evals*(row_bytes+hdr_bytes) + n_expansions*8 + 4*sum(deg)."""
    binary = w["metric"] not in ("euclidean", "cosine", "manhattan")
    row = 8 * ((w["dims"] + 63) // 64) if binary else 4 * w["dims"]
    hdr = 4 if w["metric"] == "cosine" else 0   # BQ-Cosine headers are all sqrt(padded length): the walk does not load them
    c = ctr.sum(0).astype(np.float64)
    vec = (c[0] + c[1]) * (row + hdr)
    adj = (c[2] + c[3]) * 8 + 4 * (c[4] + c[5])
    return vec + adj, vec


_REAL_STDOUT = None
_EMIT_LOCK = threading.Lock()
_EMITTED = False


def emit(line):
    """Write THE json line (once: the deadline guard and the main thread may both get here)."""
    global _EMITTED
    with _EMIT_LOCK:
        if _EMITTED:
            return
        _EMITTED = True
        data = (json.dumps(line) + "\n").encode()
        if _REAL_STDOUT is None:
            sys.stdout.write(data.decode())
            sys.stdout.flush()
        else:
            os.write(_REAL_STDOUT, data)


def arm_deadline_guard(t_start, rank, get_line, what):
    """The extras (other workloads, the config-5 sub-record) run AFTER the headline was measured but before the one JSON
    line is printed.  The driver kills a run at its per-N limit, which would lose the headline: HB_BENCH_DEADLINE seconds
    after the start, rank 0 prints the line as it stands (the unfinished extra marked skipped) and every rank leaves."""
    deadline = float(os.environ.get("HB_BENCH_DEADLINE", "780"))

    def fire():
        if rank == 0:
            line = get_line()
            if line is not None:
                line.setdefault(what, {"skipped": f"not finished {deadline:.0f}s after the start of the run (deadline guard)"})
                emit(line)
        else:
            time.sleep(3.0)   # rank 0 prints first
        os._exit(0)

    t = threading.Timer(max(1.0, deadline - (time.time() - t_start)), fire)
    t.daemon = True
    t.start()
    return t


def kernel_source_hash():
    """Identifies the kernel an ncu traffic figure was captured on (profiles/traffic.json): hash of the search kernel's sources."""
    h = hashlib.sha1()
    for f in ("search.cu", "ring.cuh", "dist.cuh", "sorted.cuh", "common.h"):
        h.update(open(os.path.join(ROOT, "hannoy_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def oracle_from_reader(rd, metric, dims, rows, ids):
    """The CPU oracle over the graph a Reader holds (device-built graphs: hb_index_layer_csr -> orc_db_set_csr)."""
    from oracle.oracle import OracleDb
    from oracle import oracle as O
    db = OracleDb(metric, dims)
    if rows.dtype == np.uint64:
        db.add_rows(ids, rows)
    else:
        db.add_items(ids, rows)
    for l, (off, nbr) in enumerate(rd.layers()):
        O.lib().orc_db_set_csr(db.h, l, O._p(off), O._p(nbr), len(nbr))
    db.set_entry_points(rd.entry_points(), rd.max_level())
    return db


def quantize_bq(x):
    """binary_quantized.rs:80-91: bit = sign bit clear, LSB-first in little-endian u64 words.  x: [m, dims] f32 on the GPU."""
    import torch
    m, dims = x.shape
    assert dims % 64 == 0
    bits = (~torch.signbit(x)).view(m, dims // 8, 8).to(torch.uint8)
    wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=x.device)
    return (bits * wts).sum(dim=2, dtype=torch.uint8).cpu().numpy().view(np.uint64).reshape(m, dims // 64)


def time_device_steps(rd, dq, nq, k, ef_raw, steps, warmup, dev, barrier=None):
    """Device-resident search of `nq` queries, `steps` launches back to back on the current stream, CUDA events.
    -> (ms per step on this rank, traversal counters [nq, 8] of one launch, kernel launches in the timed region)"""
    import torch
    from hannoy_b200 import _lib
    d_ids = torch.empty((max(nq, 1), k), dtype=torch.int32, device=dev)
    d_dist = torch.empty((max(nq, 1), k), dtype=torch.float32, device=dev)
    d_len = torch.empty((max(nq, 1),), dtype=torch.int32, device=dev)
    d_ctr = torch.zeros((max(nq, 1), 8), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step(ctr=False):
        if nq:
            rd.search_device(dq.data_ptr(), nq, k, ef_raw, d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr(),
                             d_ctr.data_ptr() if ctr else None, stream.cuda_stream)

    step(ctr=True)
    torch.cuda.synchronize()
    ctr = d_ctr.cpu().numpy().astype(np.uint64)[:nq]
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    l0 = _lib.lib().hb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    launches = _lib.lib().hb_launch_count() - l0
    if barrier:
        barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, ctr, launches


def pick_ef(rd, q_host, gt, k, efs):
    sweep, ef_pick = {}, None
    for ef in efs:
        ids, dd, lens = rd.nns(k).ef_search(ef).by_vectors_raw(q_host[:len(gt)])
        sweep[ef] = round(recall_at_k(ids, lens, gt, k), 4)
        if sweep[ef] >= RECALL_TARGET:
            ef_pick = ef
            break
    return (ef_pick if ef_pick is not None else max(sweep)), sweep


def parity_check(rd, db, q_host, k, ef_pick, threads, n_par=64):
    want = db.search_by_vector(q_host[:n_par], k, ef=max(ef_pick, k), n_threads=threads, counters=True)
    got = rd.nns(k).ef_search(ef_pick).by_vectors_raw(q_host[:n_par], counters=True)
    return bool(np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0]) and
                np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)) and np.array_equal(got[3][:, :6], want[3][:, :6]))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def gather_ceiling(row_bytes):
    """What random gathers of rows of this length reach on a B200 when nothing else is done: the best figure of the committed
    microbenchmarks (tools/gather_bench.cu, tools/gather4_bench.cu) for the nearest measured row length.  The copy peak is the
    roofline of long rows only: the memory system serves a bounded number of random rows per second whatever their length."""
    best, src = {}, {}
    for f in ("r01_gather_microbench.json", "r02_gather4_microbench.json"):
        try:
            for r in json.load(open(os.path.join(ROOT, "profiles", f)))["results"]:
                if not r.get("err") and r["gbs"] > best.get(r["row_bytes"], 0):
                    best[r["row_bytes"]], src[r["row_bytes"]] = r["gbs"], f"profiles/{f}: {r['kind']}, {r['warps_per_sm']} warps/SM"
        except Exception:
            pass
    if not best:
        return None
    rb = min(best, key=lambda b: abs(np.log(b / row_bytes)))
    return {"gbs": best[rb], "measured_row_bytes": rb, "source": src[rb]}


def roofline_of(alg_bytes, vec_bytes, ms, peaks, traffic=None, traffic_src=None, row_bytes=None):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ms / 1e3) / 1e9
    r = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
         "traffic": traffic, "traffic_source": traffic_src, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
         "algorithmic_bytes_per_step": int(alg_bytes), "gathered_vector_bytes_per_step": int(vec_bytes),
         "kernel": "hnsw_search_kernel", "frac_of_nominal_8TBs": round(achieved / 8000.0, 4)}
    g = gather_ceiling(row_bytes) if row_bytes else None
    if g:
        g["frac"] = round(achieved / g["gbs"], 4)
        r["random_gather_ceiling"] = g
    return r


def cpu_baseline_of(db, q_host, k, ef_raw, threads, n_cpu, ef_pick):
    n_cpu = min(len(q_host), n_cpu)
    t0 = time.perf_counter()
    db.search_by_vector(q_host[:n_cpu], k, ef=ef_raw, n_threads=threads)
    t_cpu = time.perf_counter() - t0
    v = n_cpu / t_cpu
    return {"value": round(v, 1), "unit": "queries/s", "cores": threads, "per_core": round(v / threads, 1), "kind": "port",
            "sample": f"first {n_cpu} queries of the batch, ef_search={ef_pick}, oracle port on {threads} threads"}


def run_other_workload(name, dev, local_rank, threads, log, steps=5):
    """One of the other single-GPU configurations as a sub-record of the line: graph built on the device, ef by the recall
    rule, device-resident QPS, roofline fraction, parity against the oracle on the same graph, CPU baseline on a sample."""
    import torch
    import hannoy_b200 as hb
    w = dict(WORKLOADS[name])
    t0 = time.time()
    binary = "binary" in w["metric"] or w["metric"] == "hamming"
    ids = np.arange(w["n"], dtype=np.uint32)
    if binary:
        rows = np.empty((w["n"], w["dims"] // 64), np.uint64)
        chunk = 250_000
        for s in range(0, w["n"], chunk):
            m = min(chunk, w["n"] - s)
            rows[s:s + m] = quantize_bq(gen_vectors(w["gen"], m, w["dims"], w["seed"] * 1000 + s // chunk, dev))
        hdr = np.full(w["n"], np.float32(np.sqrt(np.float32(w["dims"]))), np.float32)
    else:
        x = gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev)
        rows = x.cpu().numpy()
        assert w["metric"] != "cosine", "cosine headers come from the oracle's encoder (see run_sharded)"
        hdr = None
        del x
    q = gen_vectors(w["gen"], w["nq"], w["dims"], w["seed"] + 1, dev)
    q_host = q.cpu().numpy()
    st = {}
    rd = hb.Reader.build(w["metric"], w["dims"], ids, rows, hdr, M=M, M0=M0, ef_construction=EFC, alpha=ALPHA, seed=42, device=local_rank, stats=st)
    k, nq = w["k"], w["nq"]
    n_gt = min(nq, 1000)
    gt, _ = hb.exact_knn(rd, q_host[:n_gt], k)
    ef_pick, sweep = pick_ef(rd, q_host, gt, k, w["efs"])
    ef_raw = max(ef_pick, k)
    db = oracle_from_reader(rd, w["metric"], w["dims"], rows, ids)
    parity_ok = parity_check(rd, db, q_host, k, ef_pick, threads)
    ms, ctr, _ = time_device_steps(rd, q.contiguous(), nq, k, ef_raw, steps, 3, dev)
    alg, vec = algorithmic_bytes(ctr, w)
    rec = {"workload": w["desc"], "graph": f"built on the device (hb_index_build_graph) in {st.get('build_call_ms', 0) / 1e3:.1f}s", "value": round(nq / ms * 1e3, 1), "unit": "queries/s",
           "ms_per_step": round(ms, 4), "steps": steps, "ef_search": ef_pick, "recall_at_k": sweep[ef_pick], "recall_sweep": sweep,
           "parity_vs_oracle": "bit-exact (ids, distance bits, traversal counters; 64 queries)" if parity_ok else "MISMATCH",
           "roofline": roofline_of(alg, vec, ms, load_peaks(), row_bytes=rows.shape[1] * rows.itemsize) if w["n"] * w["dims"] > (1 << 26) else "index fits L2: HBM fraction not meaningful",
           "cpu_baseline": cpu_baseline_of(db, q_host, k, ef_raw, threads, 1000 if k <= 10 else 300, ef_pick)}
    log(f"{name}: {rec['value']:.0f} QPS at ef={ef_pick} recall {sweep[ef_pick]} parity {parity_ok} ({time.time() - t0:.1f}s)")
    rd.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("HB_BENCH_WORKLOAD", "c3"))
    ap.add_argument("--n-items", type=int, default=0, help="override the item count (debug; reported in config)")
    ap.add_argument("--nq", type=int, default=0, help="override the batch size (debug; reported in config)")
    ap.add_argument("--ef", type=int, default=0, help="force ef_search instead of the recall sweep")
    ap.add_argument("--extras", type=int, default=int(os.environ.get("HB_BENCH_EXTRAS", "1")), help="0: only the headline workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    t_start = time.time()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    w = dict(WORKLOADS[args.workload])
    if args.n_items:
        w["n"] = args.n_items
    if args.nq:
        w["nq"] = args.nq
    threads = len(os.sched_getaffinity(0))

    def log(msg):
        print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    # stdout carries exactly one JSON line (rank 0): anything a library prints there (NCCL's version banner when the box
    # sets NCCL_DEBUG, torch warnings) is sent to stderr instead; emit() writes the line to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    import torch
    have_cuda = torch.cuda.is_available()
    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 only
    if args.impl != "reference" and not have_cuda:
        raise SystemExit("bench.py: no CUDA device — the hannoy_b200 product path has no CPU fallback")
    dev = torch.device("cuda", local_rank) if have_cuda else torch.device("cpu")
    if have_cuda:
        torch.cuda.set_device(dev)
    use_dist = world > 1 and args.impl != "reference"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    if w.get("sharded"):
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "the sharded workload has no single-process CPU arm; use the default workload"})
            return 0
        line = run_sharded(args, w, rank, world, local_rank, dev, use_dist, threads, log)
        if rank == 0:
            emit(line)
        if torch.distributed.is_initialized():
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return 0

    # ---- synthetic data (seeded; every rank generates the same items and the same batch) ----
    t0 = time.time()
    x = gen_vectors(w["gen"], w["n"], w["dims"], w["seed"], dev)
    q = gen_vectors(w["gen"], w["nq"], w["dims"], w["seed"] + 1, dev)
    x_host = x.cpu().numpy()
    q_host = q.cpu().numpy()
    del x
    log(f"data generated in {time.time() - t0:.1f}s")

    # ---- graph (setup, untimed): rank 0 builds or loads, the others wait and load ----
    if use_dist and rank != 0:
        dist.barrier()
    db = build_or_load_graph(w, x_host, dev.type, threads, log)
    if use_dist and rank == 0:
        dist.barrier()

    n_ranks = args.gpus if args.impl == "reference" else world
    config = {"workload": w["desc"], "metric": w["metric"], "n_items": w["n"], "dims": w["dims"], "batch_queries": w["nq"],
              "global_batch_queries": w["nq"], "queries_per_gpu": -(-w["nq"] // max(n_ranks, 1)), "k": w["k"], "M": M, "M0": M0, "ef_construction": EFC,
              "ef_rule": f"smallest ef_search in {w['efs']} with recall@{w['k']} >= {RECALL_TARGET} on the first {min(w['nq'], N_GT)} queries (exact k-NN ground truth)",
              "graph": "replicated per GPU; ONE batch partitioned into contiguous slices, one per GPU (no collective)",
              "cache": "inputs larger than L2 (index rows >> 126 MB); no explicit flush" if w["n"] * w["dims"] * 4 > (1 << 28) else "index fits L2: HBM fraction not meaningful",
              "builder": "oracle restatement of hannoy Writer (reference Writer is Rust, not buildable here)"}

    if args.impl == "reference":
        return run_reference(args, w, db, x_host, q_host, threads, config, log)

    import hannoy_b200 as hb
    from hannoy_b200.sharded import partition_queries
    t0 = time.time()
    rd = hb.Reader.from_arrays(w["metric"], w["dims"], db.ids(), db.rows(), db.headers(), db.layers(), db.entry_points,
                               db.max_level, device=local_rank)
    log(f"snapshot uploaded in {time.time() - t0:.1f}s")
    k = w["k"]
    nq = w["nq"]
    barrier = dist.barrier if use_dist else None

    # ---- pick ef on rank 0 (smallest with recall@k >= target against the exact k-NN kernel), broadcast it ----
    sweep, recall, ef_pick, parity_ok = {}, None, 0, True
    if rank == 0:
        t0 = time.time()
        n_gt = min(nq, N_GT)
        gt, _ = hb.exact_knn(rd, q_host[:n_gt], k)
        log(f"exact k-NN ground truth for {n_gt} queries in {time.time() - t0:.1f}s")
        ef_pick, sweep = pick_ef(rd, q_host, gt, k, [args.ef] if args.ef else w["efs"])
        recall = sweep[ef_pick]
        log(f"recall sweep {sweep} -> ef_search={ef_pick}")
        # parity spot check against the oracle on this very graph (small sample, untimed)
        parity_ok = parity_check(rd, db, q_host, k, ef_pick, threads)
        log(f"parity vs oracle on 64 queries: {'bit-exact' if parity_ok else 'MISMATCH'}")
    if use_dist:
        t = torch.tensor([ef_pick], device=dev, dtype=torch.int64)
        dist.broadcast(t, 0)
        ef_pick = int(t.item())
    ef_raw = max(ef_pick, k)

    # ---- device-resident timing: this rank's contiguous slice of the ONE batch (strong scaling) ----
    a, b = partition_queries(nq, world, rank)
    dq_all = q.contiguous()
    dq = dq_all[a:b].contiguous()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w0 = time.time()
    ms_mine, ctr, launches = time_device_steps(rd, dq, b - a, k, ef_raw, args.steps, args.warmup, dev, barrier)
    sampler.window(t_w0, time.time())
    clocks = sampler.stop()
    ms_per_step, ms_weak = ms_mine, None
    if use_dist:
        t = torch.tensor([ms_mine], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
        # weak scaling beside it: every rank searches a whole batch
        ms_w, _, _ = time_device_steps(rd, dq_all, nq, k, ef_raw, max(3, args.steps // 4), 3, dev, barrier)
        t = torch.tensor([ms_w], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_weak = float(t.item())
    qps = nq / (ms_per_step / 1e3)
    alg_bytes, vec_bytes = algorithmic_bytes(ctr, w)   # of THIS rank's launch (rank 0 reports its own kernel)

    # ---- end to end through the host API (host buffers, H2D + D2H inside the timed region) ----
    qb = rd.nns(k).ef_search(ef_pick)
    q_pinned = torch.from_numpy(q_host[a:b].copy()).pin_memory().numpy()
    for _ in range(2):
        qb.by_vectors_raw(q_pinned)
    if use_dist:
        dist.barrier()
    e2e_steps = max(4, min(args.steps, 12)) // 2 * 2

    def e2e_run(in_flight):
        """`e2e_steps` batches through the host API with `in_flight` host threads submitting concurrently (the C-ABI
        is thread-safe: every call takes its own workspace and stream, so one batch's copies and kernel tail overlap
        the next batch's).  Returns seconds per batch."""
        per = e2e_steps // in_flight
        outs = [None] * in_flight

        def worker(i):
            qb_i = rd.nns(k).ef_search(ef_pick)
            for _ in range(per):
                outs[i] = qb_i.by_vectors_raw(q_pinned)

        ths = [threading.Thread(target=worker, args=(i,)) for i in range(in_flight)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return (time.perf_counter() - t0) / (per * in_flight), outs[0]

    e2e_run(2)  # warm the second workspace
    if use_dist:
        dist.barrier()
    t_serial, out = e2e_run(1)
    if use_dist:
        dist.barrier()
    t_e2e, out = e2e_run(2)
    if use_dist:
        t = torch.tensor([t_e2e, t_serial], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e, t_serial = float(t[0].item()), float(t[1].item())
    e2e_qps = nq / t_e2e
    h2d = q_pinned.nbytes
    d2h = out[0].nbytes + out[1].nbytes + out[2].nbytes

    line = None
    if rank == 0:
        peaks = load_peaks()
        traffic, traffic_src = None, None
        try:  # measured DRAM bytes per launch from the committed ncu capture of this workload — only while it describes this kernel
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
            if t and t["ef_search"] == ef_pick and not args.n_items and not args.nq and world == 1:
                if t.get("kernel_source_hash") == kernel_source_hash():
                    traffic, traffic_src = int(t["dram_bytes_per_launch"]), t["source"]
                else:
                    traffic_src = f"stale: {t['source']} was captured on another version of the kernel"
        except Exception:
            pass
        # one query at a time: Reader::by_vector latency, next to the CPU's (one thread, as one rayon worker sees it)
        lat = None
        try:
            n_lat = 200
            qb1 = rd.nns(k).ef_search(ef_pick)
            for i in range(10):
                qb1.by_vectors_raw(q_host[i:i + 1])
            t0 = time.perf_counter()
            for i in range(n_lat):
                qb1.by_vectors_raw(q_host[i:i + 1])
            t_gpu1 = (time.perf_counter() - t0) / n_lat
            t0 = time.perf_counter()
            db.search_by_vector(q_host[:n_lat], k, ef=ef_raw, n_threads=1)
            t_cpu1 = (time.perf_counter() - t0) / n_lat
            lat = {"single_query_ms": round(t_gpu1 * 1e3, 4), "cpu_single_query_ms_one_thread": round(t_cpu1 * 1e3, 4),
                   "note": f"hb_search_by_vector with nq = 1 (host buffers, copies included; 4 warps share the rows of the one query), mean of {n_lat} calls"}
        except Exception as e:  # noqa: BLE001
            lat = {"error": repr(e)}
        line = {
            "metric": "QPS at recall@10>=0.95 (batched)", "value": round(qps, 1), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64 popcount" if "binary" in w["metric"] or w["metric"] == "hamming" else "f32",
            "data": "synthetic", "config": config,
            "quality": {"ef_search": ef_pick, "recall_at_k": recall, "recall_sweep": sweep,
                        "parity_vs_oracle": "bit-exact (ids, distance bits, traversal counters; 64 queries)" if parity_ok else "MISMATCH"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_qps, 1), "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "batches_in_flight": 2, "one_batch_at_a_time": round(nq / t_serial, 1),
                    "note": "hb_search_by_vector on pinned host buffers (per rank: its slice of the batch); 2 host threads submit whole batches concurrently"},
            "gpu_launches": int(launches),
            "roofline": roofline_of(alg_bytes, vec_bytes, ms_mine, peaks, traffic, traffic_src, row_bytes=4 * w["dims"]),
            "cpu_baseline": cpu_baseline_of(db, q_host, k, ef_raw, threads, 2000, ef_pick),
            "latency": lat,
        }
        if world > 1:
            line["weak_scaling"] = {"value": round(nq * world / (ms_weak / 1e3), 1), "unit": "queries/s", "ms_per_step": round(ms_weak, 4),
                                    "note": "every rank searches a whole batch of its own (what round 1 reported as the headline)"}
            line["roofline"]["note"] = f"rank 0's launch: {b - a} queries of the partitioned batch"
            line["limiter"] = ("one query is one dependent walk: a slice of nq/N queries cannot finish faster than its slowest walk; below ~1 776 queries per GPU "
                               "(the resident warps) idle warps help gather rows, which shortens the walk but does not parallelise it")
    # ---- extras (untimed for the headline): the other single-GPU configurations, or the id-sharded configuration ----
    guard = None
    if args.extras and args.workload == "c3" and not args.n_items and not args.nq:
        rd.close()
        del rd
        guard = arm_deadline_guard(t_start, rank, lambda: line, "other_workloads" if world == 1 else "sharded")
        if world == 1 and rank == 0:
            others = line["other_workloads"] = {}   # filled in place: the deadline guard prints what is there
            for name in ("c1", "c2", "c4s"):
                if time.time() - t_start > 420:
                    others[name] = {"skipped": "time budget of the default run"}
                    continue
                try:
                    others[name] = run_other_workload(name, dev, local_rank, threads, log)
                except Exception as e:  # noqa: BLE001
                    others[name] = {"error": repr(e)}
        elif world > 1 and int(os.environ.get("HB_BENCH_SHARDED", "1")):
            del db, x_host
            try:
                wl = dict(WORKLOADS["c5"])
                # what is left of the run's time decides the shard size: generation + encoding + device build cost about
                # 30 s + 25 us per item per shard (measured: 6.25M items in 180 s on a 2-GPU box)
                left = float(os.environ.get("HB_BENCH_DEADLINE", "780")) - (time.time() - t_start) - 90.0
                n_time = int(max(0.0, left - 30.0) / 25e-6)
                if n_time < wl["n"]:
                    wl["n"] = max(250_000, n_time // 250_000 * 250_000)
                    wl["time_note"] = f"shards cut to {wl['n']} items to fit the time left in the run ({left:.0f}s)"
                    log(wl["time_note"])
                sh = run_sharded(args, wl, rank, world, local_rank, dev, use_dist, threads, log, steps=3)
            except Exception as e:  # noqa: BLE001
                sh = {"error": repr(e)}
            if rank == 0:
                line["sharded"] = sh
    if guard is not None:
        guard.cancel()
    if rank == 0:
        emit(line)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_sharded(args, w, rank, world, local_rank, dev, use_dist, threads, log, steps=None):
    """Config-5-shaped run: shard `rank` holds the items with id % world == rank (its own graph = one hannoy index per
    shard, src/key.rs:19-23).  A step = every rank searches the whole query batch on its shard, the per-shard top-k
    lists are exchanged and merged on every rank.  Two exchanges are timed: NCCL all-gather, and the exchange fused
    into the search kernel's epilogue (peer-memory stores over NVLink).  value = queries / step time.  Returns the
    record (rank 0) — a bench line when this is the workload, a sub-record of the c3 line otherwise."""
    import torch
    import torch.distributed as dist
    import hannoy_b200 as hb
    from hannoy_b200.sharded import ShardedSearcher
    from oracle.oracle import OracleDb
    steps = steps or args.steps
    if not dist.is_initialized():
        dist.init_process_group("nccl" if dev.type == "cuda" else "gloo", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                                **({"device_id": dev} if dev.type == "cuda" else {}))
    n, dims, nq, k, ef = w["n"], w["dims"], w["nq"], w["k"], w["efs"][0]
    # host memory: every rank holds its shard three times on the host while it is set up (the generated array, the snapshot
    # inside the library, the oracle's copy for the parity check): shrink the shards rather than take the box down
    n_asked, ram_note = n, None
    try:
        import psutil
        avail = psutil.virtual_memory().available
        n_fit = int(avail * 0.6 / max(world, 1) / (dims * 4 * 3.2))
        if n_fit < n:
            n = max(100_000, n_fit // 100_000 * 100_000)
            ram_note = f"shards cut from {n_asked} to {n} items: the host has {avail / 2**30:.0f} GiB free for {world} ranks"
            log(ram_note)
    except Exception:
        pass
    if use_dist or dist.is_initialized():   # every rank must use the same shard size
        t = torch.tensor([n], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t.item())
    t0 = time.time()
    x = gen_vectors(w["gen"], n, dims, w["seed"] + 100 * rank, dev).cpu().numpy()
    q = gen_vectors(w["gen"], nq, dims, w["seed"] + 1, dev)
    q_host = q.cpu().numpy()
    ids = (np.arange(n, dtype=np.uint64) * world + rank).astype(np.uint32)
    log(f"shard {rank}: {n} items generated in {time.time() - t0:.1f}s")
    t0 = time.time()
    st = {}
    if w.get("device_build"):
        # the item header (cosine.rs:36-38: sqrt(dot(v, v)) in the reference's summation order) is what `Writer::add_item`
        # stores with the vector: taken from the oracle's encoder, which also serves as the parity checker below
        db = OracleDb(w["metric"], dims)
        db.add_items(ids, x)
        rd = hb.Reader.build(w["metric"], dims, ids, x, db.headers(), M=M, M0=M0, ef_construction=EFC, alpha=ALPHA, seed=42 + rank, index=rank,
                             device=local_rank, stats=st)
        del x
        from oracle import oracle as O
        for l, (off, nbr) in enumerate(rd.layers()):
            O.lib().orc_db_set_csr(db.h, l, O._p(off), O._p(nbr), len(nbr))
        db.set_entry_points(rd.entry_points(), rd.max_level())
        how = f"built on the device in {time.time() - t0:.1f}s"
    else:
        db = OracleDb(w["metric"], dims)
        db.add_items(ids, x)
        db.build(M=M, M0=M0, ef_construction=EFC, alpha=ALPHA, seed=42 + rank, n_threads=max(1, threads // world))
        rd = hb.Reader.from_arrays(w["metric"], dims, db.ids(), db.rows(), db.headers(), db.layers(), db.entry_points, db.max_level,
                                   index=rank, device=local_rank)
        how = f"built by the oracle builder in {time.time() - t0:.1f}s"
    log(f"shard {rank}: graph of {n} items {how}")
    ss = ShardedSearcher(reader=rd, device=local_rank).connect_fused(nq_cap=nq, k_cap=k)
    stream = torch.cuda.current_stream()
    warm = 3

    def timed(fn):
        for _ in range(warm):
            out = fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    from hannoy_b200 import _lib
    ms_nccl, out_nccl = timed(lambda: ss.search_device(q, k, ef))
    l0 = _lib.lib().hb_launch_count()
    ms_fused, out_fused = timed(lambda: ss.search_device_fused(q, k, ef))
    launches = (_lib.lib().hb_launch_count() - l0) * steps // (steps + warm)
    # the search alone (no exchange, no merge): what the exchange costs on top
    d_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    d_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_len = torch.empty((nq,), dtype=torch.int32, device=dev)
    ms_search, _ = timed(lambda: rd.search_device(q.data_ptr(), nq, k, max(ef, k), d_ids.data_ptr(), d_dist.data_ptr(), d_len.data_ptr(), None, stream.cuda_stream))
    same = all(bool(torch.equal(a_, b_)) for a_, b_ in zip(out_nccl, out_fused))
    # parity: the reference reader on each shard index, merged by (distance bits, id), on a sample of the batch
    n_par = 64
    mine = db.search_by_vector(q_host[:n_par], k, ef=max(ef, k), n_threads=max(1, threads // world))
    parts = [None] * world
    dist.all_gather_object(parts, (mine[0], mine[1], mine[2]))
    parity_ok = True
    gi, gd = out_fused[0][:n_par].cpu().numpy().view(np.uint32), out_fused[1][:n_par].cpu().numpy().view(np.uint32)
    for i in range(n_par):
        keys = sorted((int(p[1][i, j:j + 1].view(np.uint32)[0]), int(p[0][i, j])) for p in parts for j in range(int(p[2][i])))[:k]
        parity_ok &= gi[i, :len(keys)].tolist() == [kk[1] for kk in keys] and gd[i, :len(keys)].tolist() == [kk[0] for kk in keys]
    line = None
    if rank == 0:
        line = {"metric": "QPS (batched, id-sharded index)", "value": round(nq / (ms_fused / 1e3), 1), "unit": "queries/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": round(ms_fused, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "metric": w["metric"], "items_per_shard": n, "items_per_shard_asked": WORKLOADS["c5"]["n"] if w.get("device_build") else n_asked, "host_ram_note": ram_note, "time_note": w.get("time_note"),
                           "n_shards": world, "total_items": n * world, "dims": dims,
                           "batch_queries": nq, "k": k, "ef_search": ef, "graphs": how,
                           "exchange": "fused into the search kernel epilogue (peer-memory stores over NVLink) + merge kernel",
                           "cache": "inputs larger than L2"},
                "exchange": {"fused_ms_per_step": round(ms_fused, 4), "nccl_all_gather_ms_per_step": round(ms_nccl, 4), "search_only_ms_per_step": round(ms_search, 4),
                             "nccl_all_gather_qps": round(nq / (ms_nccl / 1e3), 1), "bytes_gathered_per_rank": int(nq * k * 8 * world),
                             "fused_equals_nccl": same,
                             "limiter": "the per-shard walk; the exchange is nq*k*8 B per shard (8 MB at 100k x top-10) and costs the difference to search_only"},
                "parity_vs_oracle": "bit-exact (per-shard oracle reader, merged by (distance bits, id); 64 queries)" if parity_ok else "MISMATCH",
                "gpu_launches": int(launches)}
    dist.barrier()
    ss.close_fused()
    rd.close()
    return line


def run_reference(args, w, db, x_host, q_host, threads, config, log):
    """Reference arm: the reference's CPU search (oracle port; the Rust crate cannot be built here), whole batch per step."""
    k = w["k"]
    ef = args.ef or int(os.environ.get("HB_REF_EF", 0))
    sweep = {}
    if not ef:
        # same ef rule as our arm, on the same N_GT queries, with a host-side exact ground truth
        n_gt = min(w["nq"], N_GT)
        t0 = time.time()
        gt = cpu_ground_truth(w, db, x_host, q_host[:n_gt], k, threads)
        log(f"host exact k-NN ground truth for {n_gt} queries in {time.time() - t0:.1f}s")
        for ef in w["efs"]:
            ids, dd, lens, _ = db.search_by_vector(q_host[:n_gt], k, ef=max(ef, k), n_threads=threads)
            sweep[ef] = round(recall_at_k(ids, lens, gt, k), 4)
            if sweep[ef] >= RECALL_TARGET:
                break
        log(f"recall sweep {sweep} -> ef_search={ef}")
    ef_raw = max(ef, k)
    n_s = w["nq"]   # the whole batch, like the GPU arm
    for _ in range(max(args.warmup, 1)):
        db.search_by_vector(q_host[:n_s], k, ef=ef_raw, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        db.search_by_vector(q_host[:n_s], k, ef=ef_raw, n_threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    qps = n_s / dt
    line = {
        "impl": "reference", "metric": "QPS at recall@10>=0.95 (batched)", "value": round(qps, 1), "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "quality": {"ef_search": ef, "recall_sweep": sweep},
        "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": threads, "per_core": round(qps / threads, 1), "kind": "port",
                         "sample": f"the whole batch ({n_s} queries) per step, oracle port on {threads} threads"},
        "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
