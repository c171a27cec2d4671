"""Seeded synthetic scenes for the BASELINE.json configs.

Values come from a counter-based generator (SplitMix64 of (seed, field, entity)), so any rank
or the CPU oracle can generate exactly the same entity range without transferring it.
"""
import numpy as np

from .capi import SPRITE_DTYPE, TB_NO_PARENT, TRANSFORM_DTYPE, VELOCITY_DTYPE

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(seed, field, idx):
    """uint64 hash of (seed, field, idx) for an array of entity indices."""
    with np.errstate(over="ignore"):
        x = idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        x = x + np.uint64((seed * 0xD1B54A32D192ED03 + field * 0x8CB92BA72F3D8DD7) & 0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def _uniform(seed, field, idx, lo, hi):
    """float32 uniform in [lo, hi) with 24 random mantissa bits."""
    u = (_splitmix(seed, field, idx) >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))
    return (lo + (hi - lo) * u).astype(np.float32)


def _randint(seed, field, idx, n):
    return (_splitmix(seed, field, idx) % np.uint64(n)).astype(np.uint32)


class Scene:
    """Component arrays of entities [first, first + n) of one synthetic scene."""

    def __init__(self, name, n_total, first, n):
        self.name = name
        self.n_total = n_total
        self.first = first
        self.n = n
        self.transforms = np.zeros(n, TRANSFORM_DTYPE)
        self.sprites = np.zeros(n, SPRITE_DTYPE)
        self.velocities = None
        self.parents = None
        self.dt = None

    def bytes_in(self):
        b = self.transforms.nbytes + self.sprites.nbytes
        if self.velocities is not None:
            b += self.velocities.nbytes
        if self.parents is not None:
            b += self.parents.nbytes
        return b


def _flat(name, seed, n_total, first, n, textures, layers, z_mode, center=True):
    sc = Scene(name, n_total, first, n)
    idx = np.arange(first, first + n, dtype=np.uint64)
    t = sc.transforms
    t["translation"][:, 0] = _uniform(seed, 0, idx, -1000.0, 1000.0)
    t["translation"][:, 1] = _uniform(seed, 1, idx, -1000.0, 1000.0)
    if z_mode == "zero":
        t["translation"][:, 2] = 0.0
    elif z_mode == "steps":
        t["translation"][:, 2] = _randint(seed, 2, idx, 16).astype(np.float32)
    elif z_mode == "random":
        t["translation"][:, 2] = _uniform(seed, 2, idx, -1.0, 1.0)
    else:
        raise ValueError(z_mode)
    t["angle"][:, 2] = _uniform(seed, 3, idx, 0.0, 2.0 * np.pi)
    t["scale"][:, 0] = _uniform(seed, 4, idx, 0.5, 2.0)
    t["scale"][:, 1] = _uniform(seed, 5, idx, 0.5, 2.0)
    t["scale"][:, 2] = 1.0
    s = sc.sprites
    s["size"][:, 0] = _uniform(seed, 6, idx, 8.0, 64.0)
    s["size"][:, 1] = _uniform(seed, 7, idx, 8.0, 64.0)
    s["texture"] = _randint(seed, 8, idx, textures) if textures > 1 else 0
    s["layer"] = _randint(seed, 9, idx, layers) if layers > 1 else 0
    if center:
        t["rotation_center"][:, 0] = s["size"][:, 0] * np.float32(0.5)
        t["rotation_center"][:, 1] = s["size"][:, 1] * np.float32(0.5)
    return sc


def config1(n=10_000, first=0, count=None, seed=1):
    """10k flat sprites, 1 texture, 1 layer, no rotation centre (BASELINE configs[0])."""
    return _flat("10k-flat", seed, n, first, n if count is None else count, 1, 1, "zero", center=False)


def config2(n=1 << 20, first=0, count=None, seed=2, z_mode="steps"):
    """1M flat sprites, 64 textures, 16 layers; z in {0..15} ("steps") or U(-1,1) ("random")."""
    return _flat("1M-flat", seed, n, first, n if count is None else count, 64, 16, z_mode)


def hierarchy_level_sizes(n, depth=8):
    """Integer level sizes of a `depth`-level tree with a constant fan-out that sum to n."""
    lo, hi = 1.0, float(n)
    for _ in range(200):
        f = 0.5 * (lo + hi)
        tot = sum(f ** k for k in range(depth))
        if tot > n:
            hi = f
        else:
            lo = f
    f = lo
    sizes = [max(1, int(f ** k)) for k in range(depth)]
    sizes[-1] += n - sum(sizes)
    assert sizes[-1] > 0 and sum(sizes) == n
    return sizes


def config3(n=1 << 22, first=0, count=None, seed=3, depth=8, random_parents=False, z_mode="steps"):
    """4M-entity scene graph of depth 8, BFS-ordered ids; parent of the k-th node of level L is
    node floor(k / fan) of level L-1 (siblings contiguous), or a random node of level L-1."""
    cnt = n if count is None else count
    sc = _flat("4M-hierarchy", seed, n, first, cnt, 64, 16, z_mode)
    idx = np.arange(first, first + cnt, dtype=np.uint64)
    sc.transforms["scale"][:, 0] = _uniform(seed, 4, idx, 0.9, 1.1)
    sc.transforms["scale"][:, 1] = _uniform(seed, 5, idx, 0.9, 1.1)
    sc.transforms["translation"][:, 0] = _uniform(seed, 0, idx, -100.0, 100.0)
    sc.transforms["translation"][:, 1] = _uniform(seed, 1, idx, -100.0, 100.0)
    sizes = hierarchy_level_sizes(n, depth)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    parents = np.full(cnt, TB_NO_PARENT, dtype=np.uint32)
    ids = idx.astype(np.int64)
    level = np.searchsorted(starts, ids, side="right") - 1
    for L in range(1, depth):
        sel = level == L
        if not sel.any():
            continue
        k = ids[sel] - starts[L]
        if random_parents:
            pk = (_splitmix(seed, 10, idx[sel]) % np.uint64(sizes[L - 1])).astype(np.int64)
        else:
            pk = (k * sizes[L - 1]) // sizes[L]
        parents[sel] = (starts[L - 1] + pk).astype(np.uint32)
    sc.parents = parents
    sc.level_sizes = sizes
    return sc


def config4(n=1 << 24, first=0, count=None, seed=4):
    """16M particle-style entities: Transform + Velocity + Sprite, 4 textures, 1 layer, dt = 0.01."""
    cnt = n if count is None else count
    sc = _flat("16M-particles", seed, n, first, cnt, 4, 1, "zero")
    idx = np.arange(first, first + cnt, dtype=np.uint64)
    v = np.zeros(cnt, VELOCITY_DTYPE)
    v["v"][:, 0] = _uniform(seed, 11, idx, -100.0, 100.0)
    v["v"][:, 1] = _uniform(seed, 12, idx, -100.0, 100.0)
    sc.velocities = v
    sc.dt = 0.01
    sc.name = "16M-particles"
    return sc


def config5(n=1 << 26, first=0, count=None, seed=5, z_mode="steps"):
    """64M flat sprites as config 2, to be partitioned by contiguous id range over the ranks."""
    sc = _flat("64M-sharded", seed, n, first, n if count is None else count, 64, 16, z_mode)
    return sc


def shard_range(n_total, rank, nranks):
    """Contiguous id range [first, first+count) of `rank`."""
    per = n_total // nranks
    first = rank * per
    count = per if rank + 1 < nranks else n_total - first
    return first, count
