"""BASELINE.json's full-size configs on the GPU, checked through size-independent properties (the
CPU oracle would need minutes per frame at these sizes): the order is a permutation of the
renderables, keys recomputed from the emitted instances are sorted with ties in entity order, the
batch table tiles [0, M) at exactly the (layer, texture) run heads, frames are idempotent, every
internal path emits the same bytes, and a random sample of entities matches the oracle's compose."""
import zlib

import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu


def checksum(g):
    return (zlib.crc32(g.order.tobytes()), zlib.crc32(g.instances.tobytes()), zlib.crc32(g.vertices.tobytes()),
            zlib.crc32(g.batches.tobytes()))


def sample_matches_oracle(sc, g, k=4096, seed=0):
    """flat scenes: instance of sorted position p is the oracle's compose of entity order[p]"""
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, len(g.order), k)
    ent = g.order[pos]
    want = oracle.as_matrix4(sc.transforms[ent].view(np.float32).reshape(-1, 12))
    got = g.instances["model"][pos].reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)  # column-major -> row-major
    util.assert_floats_match(got, want, "sampled instance matrix")
    np.testing.assert_array_equal(g.instances["texture"][pos], sc.sprites["texture"][ent])
    np.testing.assert_array_equal(g.instances["layer"][pos], sc.sprites["layer"][ent])


def test_config2_1M_flat_full_size():
    sc = scenes.config2()
    ctx = util.make_ctx(sc)
    g = util.GpuFrame(ctx)
    util.check_frame_properties(ctx, g, sc.n)
    sample_matches_oracle(sc, g)
    assert checksum(g) == checksum(util.GpuFrame(ctx))  # idempotent
    ctx.close()
    # narrower digits / interleaved rows / unstaged stores + generic batches: same bytes
    for opts in ({"digit_bits": 5}, {"interleaved_rows": 1}, {"staged_emit": 0, "lt_limit": 0}):
        c2 = util.make_ctx(sc, **opts)
        assert checksum(util.GpuFrame(c2)) == checksum(g), opts
        c2.close()


def test_config2_1M_random_z_full_size():
    sc = scenes.config2(z_mode="random")
    ctx = util.make_ctx(sc)
    g = util.GpuFrame(ctx)
    assert g.info.key_bits > 32
    util.check_frame_properties(ctx, g, sc.n)
    sample_matches_oracle(sc, g)
    ctx.close()


def test_config3_4M_hierarchy_full_size():
    sc = scenes.config3()
    ctx = util.make_ctx(sc)
    g = util.GpuFrame(ctx)
    assert g.info.hierarchy_levels == 8
    util.check_frame_properties(ctx, g, sc.n)
    # world matrices of a sample of chains, recomputed on the CPU with the oracle's multiply
    w = ctx.download_world_matrices(0, sc.n)
    L = oracle.lib()
    import ctypes as C
    rng = np.random.default_rng(1)
    local = None
    for e in rng.integers(0, sc.n, 300):
        chain = [int(e)]
        while sc.parents[chain[-1]] != capi.TB_NO_PARENT:
            chain.append(int(sc.parents[chain[-1]]))
        local = oracle.as_matrix4(sc.transforms[chain].view(np.float32).reshape(-1, 12))
        m = local[-1].copy()
        for k in range(len(chain) - 2, -1, -1):
            out = np.zeros(16, np.float32)
            L.orc_mat4_mul_f32(C.c_void_p(m.ctypes.data), C.c_void_p(local[k].ctypes.data), C.c_void_p(out.ctypes.data))
            m = out
        util.assert_floats_match(w[e], m, "world matrix of a depth-%d chain" % len(chain))
    ctx.close()


def test_config4_16M_particles_full_size():
    sc = scenes.config4()
    ctx = util.make_ctx(sc)
    t0 = sc.transforms["translation"].copy()
    for _ in range(2):
        ctx.system_integrate(sc.dt)
    g = util.GpuFrame(ctx)
    assert g.info.path == capi.TB_PATH_BUCKET and g.info.num_batches == 4
    util.check_frame_properties(ctx, g, sc.n)
    # velocity system: translation advanced twice, each product and sum individually rounded
    dt = np.float32(sc.dt)
    want = t0.copy()
    for _ in range(2):
        want = want + sc.velocities["v"] * dt
    got = ctx.download_transforms(0, sc.n)["translation"]
    np.testing.assert_array_equal(got.view(np.uint32), want.astype(np.float32).view(np.uint32))
    sc.transforms["translation"][:] = got
    sample_matches_oracle(sc, g)
    # the CTA-tile emit, with nothing cached, gives the same bytes
    c2 = util.make_ctx(sc, small_emit=0, static_buckets=0)
    assert checksum(util.GpuFrame(c2)) == checksum(g)
    c2.close()
    ctx.close()


def test_config5_64M_flat_full_size_one_gpu():
    """BASELINE configs[4] on one GPU: 2^26 entities, 10 + 4 key bits (16 depths by rank), 2 radix passes. Checked without
    downloading the 10 GB of payload: the order is a permutation, the batch table equals the host's
    (layer, texture) histogram, and windows of the output are sorted and match the oracle's compose."""
    n = 1 << 26
    sc = scenes.config5(n=n)
    ctx = util.make_ctx(sc)
    info = ctx.frame()
    assert info.num_instances == n and info.sort_passes == 2 and info.key_bits == 14
    order = ctx.download_order(n)
    seen = np.zeros(n, np.uint8)
    seen[order] = 1
    assert int(seen.sum()) == n, "order is not a permutation"
    del seen
    # batch table == histogram of (layer, texture) in key order
    lt = (sc.sprites["layer"].astype(np.int64) << 16) | sc.sprites["texture"].astype(np.int64)
    vals, counts = np.unique(lt, return_counts=True)
    b = ctx.download_batches(info.num_batches)
    np.testing.assert_array_equal((b["layer"].astype(np.int64) << 16) | b["texture"], vals)
    np.testing.assert_array_equal(b["count"], counts)
    np.testing.assert_array_equal(b["first_instance"], np.concatenate([[0], np.cumsum(counts)[:-1]]))
    # windows of the sorted output (start, middle across a batch boundary, end)
    for first in (0, int(b["first_instance"][len(b) // 2]) - 5000, n - 10000):
        m = 10000
        inst = ctx.download_instances(m, first=first)
        ent = order[first:first + m]
        z = inst["model"][:, 14] + np.float32(0.0)
        u = z.view(np.uint32).astype(np.uint64)
        u = np.where(u >> np.uint64(31), u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
        key = (inst["layer"].astype(np.uint64) << np.uint64(48)) | (inst["texture"].astype(np.uint64) << np.uint64(32)) | u
        assert np.all(key[:-1] <= key[1:])
        tie = key[:-1] == key[1:]
        assert np.all(ent[:-1][tie] < ent[1:][tie])
        want = oracle.as_matrix4(sc.transforms[ent].view(np.float32).reshape(-1, 12))
        got = inst["model"].reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1, 16)
        util.assert_floats_match(got, want, "instance matrix")
        np.testing.assert_array_equal(inst["texture"], sc.sprites["texture"][ent])
        vtx = ctx.download_vertices(m, first=first)
        assert np.array_equal(vtx["x"][:, 0], inst["model"][:, 12]) and np.array_equal(vtx["y"][:, 0], inst["model"][:, 13])
    ctx.close()


def test_16M_sorted_steady_frames_full_size():
    """2^24 sorted sprites: the full frame, the frame that keeps the sort (key pass + ordered emit, forced by an
    upload of unchanged values), the single-kernel frame and the uncached path all emit the same bytes; after a
    tick that only moves x / y the order is kept and the sampled instances match the oracle's compose."""
    n = 1 << 24
    sc = scenes.config2(n=n, seed=6)
    ctx = util.make_ctx(sc)
    g0 = util.GpuFrame(ctx)
    assert g0.info.sort_passes == 2 and g0.info.num_instances == n
    util.check_frame_properties(ctx, g0, n)
    sample_matches_oracle(sc, g0)
    c0 = checksum(g0)
    del g0
    ctx.upload_transforms(0, sc.transforms[:1000])  # same values: plan and sort survive, the key pass runs
    g1 = util.GpuFrame(ctx)
    assert g1.info.replanned == 0 and g1.info.kernel_launches > 1 and checksum(g1) == c0
    del g1
    g2 = util.GpuFrame(ctx)  # nothing touched: one kernel
    assert g2.info.kernel_launches == 1 and checksum(g2) == c0
    del g2
    for _ in range(3):  # still nothing touched: the rows get copied into draw order, then frames stream from the copy
        g2 = util.GpuFrame(ctx)
        assert checksum(g2) == c0
        del g2
    c2 = util.make_ctx(sc, sorted_cache=0, static_buckets=0, cuda_graphs=0)
    util.GpuFrame(c2)
    assert checksum(util.GpuFrame(c2)) == c0
    c2.close()
    # x / y motion through the velocity system: the draw order stays, the matrices move
    vel = np.zeros(n, capi.VELOCITY_DTYPE)
    vel["v"][:, 0] = 3.0
    vel["v"][:, 1] = -2.0
    ctx.upload_velocities(0, vel)
    ctx.system_integrate(0.5)
    g3 = util.GpuFrame(ctx)
    assert g3.info.replanned == 0 and zlib.crc32(g3.order.tobytes()) == c0[0]
    sc.transforms["translation"][:, 0] += np.float32(1.5)
    sc.transforms["translation"][:, 1] += np.float32(-1.0)
    sample_matches_oracle(sc, g3)
    t = ctx.download_transforms(0, 4096)
    np.testing.assert_array_equal(t.view(np.uint32), sc.transforms[:4096].view(np.uint32))
    ctx.close()
