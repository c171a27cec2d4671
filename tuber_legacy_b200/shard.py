"""One scene over several GPUs: the per-frame exchange around the C ABI's
tb_shard_* calls. One process per GPU; `torch.distributed` is only the transport for two small
collectives per frame (3 words of key statistics, one short-key histogram per rank). The instance
payload itself never goes through a collective: every rank's emit kernel stores its instances at
their global draw-order positions straight into the presenting rank's buffers (peer memory over
NVLink, mapped with CUDA IPC).

The driver is written against a small backend interface so the exchange logic can be exercised
with the gloo backend on CPU (tests/test_shard_gloo.py) without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


class ShardBackend:
    """What the driver needs from one rank's device context (capi.Context provides it on a GPU)."""

    device = "cpu"

    def key_stats(self):  # -> [or, or_of_not, count]
        raise NotImplementedError

    def plan(self, global_stats):  # -> nbins
        raise NotImplementedError

    def key_histogram(self, hist):  # fills the int32 tensor `hist` (nbins)
        raise NotImplementedError

    def place(self, all_hists):  # int32 tensor (world, nbins) -> global instance count
        raise NotImplementedError

    def frame(self):
        raise NotImplementedError


class CudaShard(ShardBackend):
    """A capi.Context on one GPU."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device

    def key_stats(self):
        return self.ctx.shard_key_stats()

    def plan(self, global_stats):
        return self.ctx.shard_plan(global_stats)

    def key_histogram(self, hist):
        self.ctx.shard_key_histogram(hist.data_ptr())

    def place(self, all_hists):
        return self.ctx.shard_place(all_hists.data_ptr())

    def frame(self):
        return self.ctx.frame()


def reduce_stats(per_rank):
    """OR the two mask words, sum the counts. (NCCL has no bitwise-OR reduction, so the three
    words are all-gathered and folded here.)"""
    g_or, g_nand, g_cnt = 0, 0, 0
    for st in per_rank:
        g_or |= int(st[0])
        g_nand |= int(st[1])
        g_cnt += int(st[2])
    return [g_or, g_nand, g_cnt]


def _to_i64(v):
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


class ShardedFrame:
    """Runs the multi-GPU frame protocol of include/tuber_b200.h on this rank."""

    def __init__(self, backend, rank, world, group=None):
        self.backend = backend
        self.rank = rank
        self.world = world
        self.group = group
        self._hist = None
        self._all = None

    def step(self):
        b = self.backend
        st = b.key_stats()
        mine = torch.tensor([_to_i64(v) for v in st], dtype=torch.int64, device=b.device)
        if self.world > 1:
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=self.group)
            per_rank = [[int(x) & 0xFFFFFFFFFFFFFFFF for x in t.tolist()] for t in gathered]
        else:
            per_rank = [st]
        g = reduce_stats(per_rank)
        nbins = b.plan(g)
        if self._hist is None or self._hist.numel() != nbins:
            self._hist = torch.zeros(nbins, dtype=torch.int32, device=b.device)
            self._all = torch.zeros((self.world, nbins), dtype=torch.int32, device=b.device)
        b.key_histogram(self._hist)
        if self.world > 1:
            dist.all_gather_into_tensor(self._all.view(-1), self._hist, group=self.group)
        else:
            self._all[0].copy_(self._hist)
        total = b.place(self._all)
        out = b.frame()
        if self.world > 1:
            dist.barrier(group=self.group)  # every rank's stores have landed in the presenter's buffers
        return out, total


# ---- scene graphs across shards --------------------------------------------------------------------

NO_PARENT = 0xFFFFFFFF


def ancestor_closure(parents, first, count):
    """Which entities of other shards a shard must also hold so that every Parent link of its own
    entities resolves locally (the chain's world matrices are then recomputed on this rank: no
    exchange of parent matrices at all).

    parents: global uint32 array (NO_PARENT for roots). Returns (ghost_ids, local_parents):
      ghost_ids      ascending global ids of the ancestors outside [first, first+count): the shard
                     stores them after its own entities, WITHOUT a Sprite (they are never drawn);
      local_parents  the Parent column to upload for the count + len(ghost_ids) local entities, as
                     the ids the C ABI expects after tb_shard_set(rank, nranks, first): first + local index.
    """
    parents = np.asarray(parents, dtype=np.uint32)
    own = np.arange(first, first + count, dtype=np.int64)
    in_own = lambda ids: (ids >= first) & (ids < first + count)  # noqa: E731
    ghosts = np.zeros(0, dtype=np.int64)
    frontier = own
    while True:
        p = parents[frontier].astype(np.int64)
        p = p[p != NO_PARENT]
        p = np.unique(p[~in_own(p)])
        new = np.setdiff1d(p, ghosts, assume_unique=True)
        if new.size == 0:
            break
        ghosts = np.union1d(ghosts, new)
        frontier = new
    local_ids = np.concatenate([own, ghosts])
    p = parents[local_ids].astype(np.int64)
    has = p != NO_PARENT
    local_index = np.where(in_own(p), p - first, count + np.searchsorted(ghosts, p))
    local_parents = np.where(has, first + local_index, NO_PARENT).astype(np.uint32)
    return ghosts.astype(np.uint32), local_parents
