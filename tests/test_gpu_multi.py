"""Multi-device paths on a box with >= 2 GPUs (skipped otherwise): the single-process mode of
c2rt_render (c2rt_init(N)) must give the same frame, bit for bit, as one device — both when every
device copies its interleaved bands to the host itself (default) and when peers store their bands
straight into device 0's frame through peer-mapped pointers (C2RT_GATHER=root)."""
import os

import numpy as np
import pytest

import chess2rt_b200 as c2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "scenes")


@pytest.mark.parametrize("gather", ["direct", "root"])
@pytest.mark.parametrize("scene,size", [("lecture5.sdl", (333, 217)), ("chessboard.sdl", (640, 360)), ("lecture4.sdl", (64, 1100))])
def test_single_process_multi_device_frame_is_bit_identical(scene, size, gather, monkeypatch):
    monkeypatch.setenv("C2RT_GATHER", gather)
    n = c2.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    try:
        c2.init(1, [0])
        one = c2.HostScene(os.path.join(SC, scene))
        one.set_frame_size(*size)
        ref, ref_a, st1 = one.render(argb=True, count_rays=True)
        one.close()
        for k in sorted({2, min(n, 4), n}):
            c2.init(k)
            many = c2.HostScene(os.path.join(SC, scene))
            many.set_frame_size(*size)
            rgb, argb, st = many.render(argb=True, count_rays=True)
            many.close()
            assert st.n_gpus == k and st.launches >= k
            np.testing.assert_array_equal(rgb, ref)
            np.testing.assert_array_equal(argb, ref_a)
            assert (st.primary_rays, st.shadow_rays) == (st1.primary_rays, st1.shadow_rays)
    finally:
        c2.init(1, [0])
