"""The headless presenter (tb_present_headless): the frame's own draw list, rasterised on the device in
draw order through Matrix4::new_orthographic (matrix.rs:41-58), against a CPU rasterisation of the oracle's list —
every pixel's top-most draw position must agree."""
import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["sorted", "particles", "hierarchy"])
def test_id_buffer_matches_cpu_rasterisation_of_the_oracle_list(kind):
    sc = {"sorted": lambda: scenes.config2(n=6_000, seed=81), "particles": lambda: scenes.config4(n=6_000, seed=82),
          "hierarchy": lambda: scenes.config3(n=6_000, seed=83)}[kind]()
    ctx = util.make_ctx(sc)
    ref = util.make_oracle(sc)
    # the camera sees the middle of the scene: some quads are outside (culled), some straddle the border
    vp = capi.orthographic(-600, 600, -450, 450, -100, 100)
    ctx.set_view_projection(vp)
    g = util.GpuFrame(ctx)
    r = ref.frame(view_proj=vp)
    util.assert_frame_parity(g, r)
    W, H = 320, 240
    ids, rgba, drawn, culled = ctx.present_headless(W, H)
    want = oracle.rasterize(r.vertices, W, H)
    assert np.array_equal(ids, want), f"{int((ids != want).sum())} pixels differ"
    assert drawn + culled == g.info.num_instances and drawn > 0
    if kind != "hierarchy":
        assert culled > 0
    covered = ids > 0
    assert covered.any() and (rgba[covered] >> 24 == 0xFF).all() and (rgba[~covered] == 0).all()
    # the colour is a function of the (texture, layer) of the instance on top
    top = ids[covered] - 1
    tl = (g.instances["texture"][top].astype(np.uint64) << 32) | g.instances["layer"][top]
    assert len(np.unique(rgba[covered])) <= len(np.unique(tl))
    ctx.close()


def test_painters_rule_follows_the_draw_order_not_the_entity_order():
    # two overlapping sprites: the one on the higher layer is drawn later and wins the overlap, whatever its id
    sc = scenes.config1(n=2)
    sc.transforms["translation"][:] = [[0, 0, 0], [10, 10, 0]]
    sc.transforms["angle"][:] = 0
    sc.transforms["scale"][:] = 1
    sc.sprites["size"][:] = 40
    sc.sprites["layer"][:] = [5, 1]
    ctx = util.make_ctx(sc)
    ctx.set_view_projection(capi.orthographic(-50, 100, -50, 100, -1, 1))
    g = util.GpuFrame(ctx)
    assert list(g.order) == [1, 0]
    ids, _, drawn, culled = ctx.present_headless(150, 150)
    assert (drawn, culled) == (2, 0)
    # pixel at world (20, 20) is covered by both: entity 0 (layer 5) is draw position 1 -> id 2
    assert ids[70, 70] == 2
    assert ids[50 + 45, 50 + 45] == 1  # only entity 1 reaches (45, 45)
    ctx.close()


def test_presenter_needs_a_frame_and_reports_empty_scenes():
    with capi.Context(16) as ctx:
        with pytest.raises(capi.TuberError):
            ctx.present_headless(8, 8)
        ctx.set_entity_count(4)
        ctx.frame()
        ids, rgba, drawn, culled = ctx.present_headless(8, 8)
        assert not ids.any() and not rgba.any() and (drawn, culled) == (0, 0)
