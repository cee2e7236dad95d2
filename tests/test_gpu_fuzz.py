"""Parity fuzzing: random scenes over the whole supported surface (tests/scene_fuzz.py), CUDA path vs the
CPU oracle under the north-star bar.  Seeds are fixed, so a failure is reproducible with
`python tests/scene_fuzz.py <seed> > /tmp/s.sdl`."""
import os

import numpy as np
import pytest

import chess2rt_b200 as c2
from oracle_binding import OracleScene, pack_rgb32, parity_report
from scene_fuzz import generate, generate_planes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _ctx():
    c2.init(1, [0])
    yield
    c2.shutdown()


@pytest.mark.parametrize("seed", list(range(48)))
def test_random_scene_matches_oracle(seed, tmp_path):
    path = tmp_path / f"fuzz{seed}.sdl"
    path.write_text(generate(seed))
    g, o = c2.HostScene(path), OracleScene(path)
    rgb, argb, st = g.render(argb=True, seed=seed, count_rays=True)
    ref, ost = o.render(threads=0, seed=seed)
    rep = parity_report(rgb, ref, argb)
    # The bar of BASELINE.json: float RGB within 1e-3; <= 0.1 % of 8-bit pixels off by more than 1 LSB.  Random scenes
    # contain exactly coincident surfaces (a CSG child flush with its sibling, pieces cut by the floor plane), where a
    # 1-ulp difference legitimately flips a hit: allow the 0.1 % there, but never a systematic difference.
    assert rep["frac_over_1e-3"] <= 1e-3, (seed, rep)
    assert rep["frac_over_1lsb"] <= 1e-3, (seed, rep)
    np.testing.assert_array_equal(argb, pack_rgb32(rgb))
    assert st.primary_rays == ost.primary_rays
    assert abs(int(st.shadow_rays) - int(ost.shadow_rays)) <= max(2, int(1e-3 * ost.shadow_rays)), (st.shadow_rays, ost.shadow_rays)
    assert ost.csg_max_crossings <= 8


@pytest.mark.parametrize("seed", list(range(32)))
def test_random_plane_scene_matches_oracle(seed, tmp_path):
    """The MODE_SOLO and plane-only kernel classes (un-normalised camera rays, sign-test shadowing among planes)."""
    path = tmp_path / f"planes{seed}.sdl"
    path.write_text(generate_planes(seed))
    g, o = c2.HostScene(path), OracleScene(path)
    rgb, argb, st = g.render(argb=True, seed=seed, count_rays=True)
    ref, ost = o.render(threads=0, seed=seed)
    rep = parity_report(rgb, ref, argb)
    assert rep["frac_over_1e-3"] <= 1e-3, (seed, rep)
    assert rep["frac_over_1lsb"] <= 1e-3, (seed, rep)
    np.testing.assert_array_equal(argb, pack_rgb32(rgb))
    assert (st.primary_rays, st.shadow_rays) == (ost.primary_rays, ost.shadow_rays)
