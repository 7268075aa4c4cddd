"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two layouts:
  * replicas  — the graph is replicated on every GPU, the query batch is partitioned into contiguous
                slices; there is no collective in the data path (`partition_queries`).
  * id-shards — item `id` lives on shard `id % n_shards` (one hannoy index per shard, src/key.rs:19-23);
                every rank searches ALL queries on its shard, the per-shard top-k lists are all-gathered
                and merged by (distance bits, id) (`ShardedSearcher`).  Result == the reference reader run
                on each of the shard indexes, merged — not a single graph over all items.
"""
import numpy as np


def partition_queries(nq, world, rank):
    """Contiguous slice [start, stop) of a batch of nq queries owned by `rank`."""
    base, rem = divmod(nq, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_of(item_ids, n_shards):
    return np.asarray(item_ids, dtype=np.uint64) % n_shards


def pad_topk(ids, dist, lens, k):
    """Pad per-query result rows to k entries with id = UINT32_MAX, dist = +inf (the merge kernel's sentinel)."""
    ids = np.array(ids, dtype=np.uint32, copy=True)
    dist = np.array(dist, dtype=np.float32, copy=True)
    for i, n in enumerate(lens):
        n = 0 if n == 0xFFFFFFFF else int(n)
        ids[i, n:] = 0xFFFFFFFF
        dist[i, n:] = np.inf
    return ids, dist


class ShardedSearcher:
    """Search an id-sharded index: local search on this rank's shard, all-gather, merge.

    `local_search(q, count, ef) -> (ids[nq,count], dist[nq,count], lens[nq])` and
    `merge(ids[world,nq,count], dist[world,nq,count]) -> (ids[nq,count], dist[nq,count], lens[nq])`
    default to the CUDA engine (Reader + hb_merge_topk_device); tests inject CPU stand-ins to exercise the
    gather layout under the gloo backend.
    """

    def __init__(self, reader=None, group=None, local_search=None, merge=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.reader = reader
        self.device = device
        self.local_search = local_search or self._cuda_local_search
        self.merge = merge or self._cuda_merge

    def _cuda_local_search(self, q, count, ef):
        return self.reader.nns(count).ef_search(ef).by_vectors_raw(q)

    def _cuda_merge(self, ids, dist):
        import torch
        from .reader import merge_topk_device
        world, nq, k = ids.shape
        dev = torch.device("cuda", self.device if self.device is not None else torch.cuda.current_device())
        d_ids = torch.from_numpy(ids.view(np.int32)).to(dev)
        d_dist = torch.from_numpy(dist).to(dev)
        o_ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        o_dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
        o_len = torch.empty((nq,), dtype=torch.int32, device=dev)
        merge_topk_device(dev.index, d_ids.data_ptr(), d_dist.data_ptr(), world, nq, k, o_ids.data_ptr(), o_dist.data_ptr(),
                          o_len.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return (o_ids.cpu().numpy().view(np.uint32), o_dist.cpu().numpy(), o_len.cpu().numpy().view(np.uint32))

    def search_device(self, d_q, count, ef):
        """Device-resident path (NCCL): `d_q` is a CUDA float32 tensor [nq, dims] holding ALL queries on this rank.
        Local search on this rank's shard -> one all-gather of the padded per-shard top-k over NVLink -> k-way merge
        kernel; everything stays in HBM and is ordered on the current stream.  Returns CUDA tensors
        (ids int32 [nq, count] (bit pattern of the u32 ids), dist float32 [nq, count], len int32 [nq])."""
        import torch
        nq = d_q.shape[0]
        dev = d_q.device
        stream = torch.cuda.current_stream(dev)
        ids = torch.full((nq, count), -1, dtype=torch.int32, device=dev)          # 0xFFFFFFFF = the merge sentinel
        dist = torch.full((nq, count), float("inf"), dtype=torch.float32, device=dev)
        lens = torch.empty((nq,), dtype=torch.int32, device=dev)
        self.reader.search_device(d_q.data_ptr(), nq, count, max(ef, count), ids.data_ptr(), dist.data_ptr(), lens.data_ptr(),
                                  None, stream.cuda_stream)
        # the kernel zeroes the slots past out_len (its output is a function of its inputs): re-pad them with the merge
        # sentinel, or a shard that finds fewer than `count` hits would contribute bogus (id 0, distance 0) entries
        pad = torch.arange(count, device=dev, dtype=torch.int32)[None, :] >= (lens & 0x7fffffff)[:, None]
        ids.masked_fill_(pad, -1)
        dist.masked_fill_(pad, float("inf"))
        g_ids = torch.empty((self.world * nq, count), dtype=torch.int32, device=dev)   # == [world][nq][count]
        g_dist = torch.empty((self.world * nq, count), dtype=torch.float32, device=dev)
        self.dist.all_gather_into_tensor(g_ids, ids, group=self.group)
        self.dist.all_gather_into_tensor(g_dist, dist, group=self.group)
        from .reader import merge_topk_device
        o_ids = torch.empty((nq, count), dtype=torch.int32, device=dev)
        o_dist = torch.empty((nq, count), dtype=torch.float32, device=dev)
        o_len = torch.empty((nq,), dtype=torch.int32, device=dev)
        merge_topk_device(dev.index, g_ids.data_ptr(), g_dist.data_ptr(), self.world, nq, count, o_ids.data_ptr(),
                          o_dist.data_ptr(), o_len.data_ptr(), stream.cuda_stream)
        return o_ids, o_dist, o_len

    # ---- all-gather fused into the search kernel (peer memory over NVLink, no NCCL in the data path) ----
    def connect_fused(self, nq_cap, k_cap):
        """Create this rank's exchange buffer, swap the CUDA-IPC handles (one small host-side all-gather, set-up only)
        and map every peer's buffer.  Afterwards `search_device_fused` needs no collective library call."""
        import ctypes as C
        import torch
        from . import _lib as L
        from .reader import _check
        dev = self.device if self.device is not None else torch.cuda.current_device()
        g = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _check(L.lib().hb_shard_group_create(dev, self.world, self.rank, nq_cap, k_cap, C.byref(g), C.cast(handle, C.c_void_p)))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=torch.device("cuda", dev) if self.dist.get_backend(self.group) == "nccl" else "cpu")
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=mine.device)
        self.dist.all_gather_into_tensor(allh, mine, group=self.group)
        raw = bytes(allh.cpu().numpy().tobytes())
        _check(L.lib().hb_shard_group_connect(g, raw))
        self._group_handle = g
        self.dist.barrier(group=self.group)
        return self

    def search_device_fused(self, d_q, count, ef):
        """Collective.  Like search_device, but the per-shard top-k lists travel inside the search kernel's epilogue
        (stores into every peer's gather buffer) instead of through an all-gather."""
        import torch
        from . import _lib as L
        from .reader import _check
        nq = d_q.shape[0]
        dev = d_q.device
        stream = torch.cuda.current_stream(dev)
        o_ids = torch.empty((nq, count), dtype=torch.int32, device=dev)
        o_dist = torch.empty((nq, count), dtype=torch.float32, device=dev)
        o_len = torch.empty((nq,), dtype=torch.int32, device=dev)
        _check(L.lib().hb_search_sharded_device(self.reader._h, self._group_handle, d_q.data_ptr(), nq, count, max(ef, count),
                                                o_ids.data_ptr(), o_dist.data_ptr(), o_len.data_ptr(), stream.cuda_stream))
        return o_ids, o_dist, o_len

    def close_fused(self):
        from . import _lib as L
        if getattr(self, "_group_handle", None):
            L.lib().hb_shard_group_free(self._group_handle)
            self._group_handle = None

    def search(self, q, count, ef):
        import torch
        if self.reader is not None and self.dist.get_backend(self.group) == "nccl" and self.local_search == self._cuda_local_search:
            dev = torch.device("cuda", self.device if self.device is not None else torch.cuda.current_device())
            d_q = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(dev)
            o_ids, o_dist, o_len = self.search_device(d_q, count, ef)
            torch.cuda.synchronize(dev)
            return (o_ids.cpu().numpy().view(np.uint32), o_dist.cpu().numpy(), o_len.cpu().numpy().view(np.uint32))
        ids, dist, lens = self.local_search(q, count, ef)[:3]
        ids, dist = pad_topk(ids, dist, lens, count)
        backend = self.dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t_ids = torch.from_numpy(ids.view(np.int32)).to(dev)
        t_dist = torch.from_numpy(dist).to(dev)
        nq, k = t_ids.shape
        g_ids = torch.empty((self.world * nq, k), dtype=t_ids.dtype, device=dev)     # concatenated == [world][nq][k]
        g_dist = torch.empty((self.world * nq, k), dtype=t_dist.dtype, device=dev)
        self.dist.all_gather_into_tensor(g_ids, t_ids, group=self.group)    # one all-gather of nq*k*(4+4) bytes per rank
        self.dist.all_gather_into_tensor(g_dist, t_dist, group=self.group)
        return self.merge(g_ids.cpu().numpy().view(np.uint32).reshape(self.world, nq, k),
                          g_dist.cpu().numpy().reshape(self.world, nq, k))
