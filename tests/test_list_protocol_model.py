"""Model of the completion flags of steady list frames (csrc/emit.cuh list_signal_* / list_wait_kernel), in
numpy: whatever the interleaving of the ranks' launches, the presenting rank never expands a list position before its owner
has written it, and after the last chunk it has expanded every position exactly once. The CUDA path itself is checked on
hardware (tests/test_gpu_sharded_lib.py); this pins the arithmetic the flags carry."""
import numpy as np
import pytest

TILE = 64


def global_placement(keys_per_rank, nkeys):
    """Histogram placement: position of (key, rank, i-th of that key on that rank) in the global draw order."""
    hist = np.array([np.bincount(k, minlength=nkeys) for k in keys_per_rank])      # [rank][key]
    key_start = np.concatenate([[0], np.cumsum(hist.sum(axis=0))[:-1]])            # first position of every key
    rank_start = key_start[None, :] + np.concatenate([np.zeros((1, nkeys), int), np.cumsum(hist, axis=0)[:-1]])
    return hist, rank_start


@pytest.mark.parametrize("seed", range(12))
def test_sorted_plans_one_flag_per_chunk(seed):
    rng = np.random.default_rng(seed)
    nranks, nkeys, nchunks = int(rng.integers(1, 9)), int(rng.integers(1, 300)), 16
    counts = rng.integers(0, 5000, nranks)
    if seed % 4 == 0:
        counts[rng.integers(0, nranks)] = 0  # an empty shard still reports
    keys = [np.sort(rng.integers(0, nkeys, c)) for c in counts]  # a shard's draw order: ascending key (ties by entity id)
    hist, rank_start = global_placement(keys, nkeys)
    total = int(counts.sum())
    # remap[key] = global start of (key, rank) - local start of key
    remap = [rank_start[r] - np.concatenate([[0], np.cumsum(hist[r])[:-1]]) for r in range(nranks)]
    written = np.full(total, -1)  # chunk in which the owner stores the position
    flags = np.zeros((nranks, nchunks), np.int64)
    for r in range(nranks):
        c = int(counts[r])
        ntiles = -(-c // TILE)
        per = max(-(-ntiles // nchunks), 1)
        for k in range(nchunks):
            t0, t1 = min(ntiles, k * per), (ntiles if k + 1 == nchunks else min(ntiles, (k + 1) * per))
            lo, hi = t0 * TILE, min(t1 * TILE, c)
            if hi > lo:
                pos = np.arange(lo, hi) + remap[r][keys[r][lo:hi]]
                assert np.all(written[pos] == -1)
                written[pos] = k
            nxt = min(t1 * TILE, c)
            flags[r, k] = total if nxt >= c else nxt + remap[r][keys[r][nxt]]  # list_signal_kernel
    assert np.all(written >= 0), "every position has exactly one owner"
    prev = 0
    expanded = np.zeros(total, int)
    for k in range(nchunks):  # list_wait_kernel, sorted plans: [previous end, lowest position some rank has not passed)
        hi = max(int(flags[:, k].min()), prev)
        assert np.all(written[prev:hi] <= k), "expanded before its owner stored it"
        expanded[prev:hi] += 1
        prev = hi
    assert prev == total and np.all(expanded == 1)


@pytest.mark.parametrize("seed", range(12))
def test_one_pass_plans_one_flag_per_bucket(seed):
    """emit_small walks the shard in ENTITY order: per bucket, positions grow with the tile; the cached per-tile offsets
    tile_off[b][t] (where tile t's bucket-b elements start, shard-local) are what list_signal_buckets_kernel reports."""
    rng = np.random.default_rng(100 + seed)
    nranks, nb, nchunks = int(rng.integers(1, 9)), int(rng.choice([1, 2, 4, 8])), 16
    counts = rng.integers(0, 6000, nranks)
    if seed % 3 == 0:
        counts[rng.integers(0, nranks)] = 0
    buckets = [rng.integers(0, nb, c) for c in counts]  # bucket of every entity, entity order
    hist, rank_start = global_placement(buckets, nb)
    total = int(counts.sum())
    written = np.full(total, -1)
    flags = np.zeros((nranks, nchunks, nb), np.int64)
    for r in range(nranks):
        c = int(counts[r])
        ntiles = -(-c // TILE)
        local_start = np.concatenate([[0], np.cumsum(hist[r])[:-1]])
        remap = rank_start[r] - local_start
        # tile_off[b][t]: shard-local position of the first bucket-b element of tile t
        per_tile = np.array([np.bincount(buckets[r][t * TILE:(t + 1) * TILE], minlength=nb) for t in range(ntiles)]).reshape(ntiles, nb)
        tile_off = local_start[None, :] + np.concatenate([np.zeros((1, nb), int), np.cumsum(per_tile, axis=0)[:-1]]) if ntiles else np.zeros((0, nb), int)
        per = max(-(-ntiles // nchunks), 1)
        for k in range(nchunks):
            t0, t1 = min(ntiles, k * per), (ntiles if k + 1 == nchunks else min(ntiles, (k + 1) * per))
            for t in range(t0, t1):
                b = buckets[r][t * TILE:(t + 1) * TILE]
                rank_in_tile = np.array([np.sum(b[:i] == b[i]) for i in range(len(b))])
                pos = tile_off[t][b] + rank_in_tile + remap[b]
                assert np.all(written[pos] == -1)
                written[pos] = k
            for b in range(nb):  # list_signal_buckets_kernel
                if ntiles == 0:
                    flags[r, k, b] = 0
                else:
                    flags[r, k, b] = remap[b] + (tile_off[t1][b] if t1 < ntiles else tile_off[0][b] + hist[r][b])
    assert np.all(written >= 0)
    done = rank_start.copy()  # list_starts_kernel: where every (rank, bucket) stream starts
    expanded = np.zeros(total, int)
    for k in range(nchunks):  # list_wait_kernel, ranks x buckets streams
        for r in range(nranks):
            for b in range(nb):
                lo, hi = int(done[r, b]), max(int(flags[r, k, b]), int(done[r, b]))
                assert np.all(written[lo:hi] <= k), "expanded before its owner stored it"
                expanded[lo:hi] += 1
                done[r, b] = hi
    assert np.all(expanded == 1)
