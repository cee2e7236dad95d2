"""Pins the CPU oracle: analytic known answers, the reference's own BMP KATs, the quirks of
SURVEY.md F9, and the committed golden fixtures.  No GPU.  (The reference has no test on the render
path — SURVEY.md §4 — so for that path parity is UNPINNED against the reference itself; see DESIGN.md.)"""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from bmp_kats import KAT1, KAT1_PIXELS, KAT1_SIZE, KAT2, KAT2_PIXELS, KAT2_SIZE
from oracle_binding import OracleScene, oracle_lib, pack_rgb32, srgb_lut

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "scenes")
GOLD = os.path.join(ROOT, "tests", "golden")


def kat_intersect(kind, params, o, d, max_dist=1e99):
    lib = oracle_lib()
    P = (C.c_double * 4)(*params)
    O = (C.c_double * 3)(*o)
    n = math.sqrt(sum(x * x for x in d))
    D = (C.c_double * 3)(*[x / n for x in d])
    out = (C.c_double * 10)()
    assert lib.orc_kat_intersect(kind, P, O, D, max_dist, out) == 0
    o_ = list(out)
    return bool(o_[0]), o_[1], o_[2:5], o_[5:8], o_[8], o_[9]


def test_camera_looks_down_and_corners():
    # SURVEY.md §8c behavioural pin of the recalled gfm conventions: pitch -30 looks DOWN
    s = OracleScene(os.path.join(SC, "lecture4.sdl"))
    v = s.camera_vectors()
    pos, ul, ur, dl, right, up, front = v
    np.testing.assert_allclose(front, [0, -0.5, math.sqrt(3) / 2], atol=1e-15)
    np.testing.assert_allclose(right, [1, 0, 0], atol=1e-15)
    np.testing.assert_allclose(up, [0, math.sqrt(3) / 2, 0.5], atol=1e-15)
    # corner scaling: |(x, y)| = tan(fov/2) = 1 with x/y = -aspect (camera.d:84-100)
    aspect = 640 / 480
    y = 1 / math.hypot(aspect, 1)
    x = -aspect * y
    c, s_ = math.cos(math.radians(-30)), math.sin(math.radians(-30))
    def rx(p):  # row vector x rotateX(a): c[1][1]=cos c[1][2]=-sin c[2][1]=sin c[2][2]=cos
        return np.array([p[0], p[1] * c + p[2] * s_, -p[1] * s_ + p[2] * c])
    np.testing.assert_allclose(ul - pos, rx([x, y, 1]), atol=1e-14)
    np.testing.assert_allclose(ur - pos, rx([-x, y, 1]), atol=1e-14)
    np.testing.assert_allclose(dl - pos, rx([x, -y, 1]), atol=1e-14)


def test_plane_closed_form():
    hit, t, p, n, u, v = kat_intersect(0, [2, 0, 0, 0], [0, 165, 0], [0.1, -1, 0.5])
    assert hit
    d = np.array([0.1, -1, 0.5]) / np.linalg.norm([0.1, -1, 0.5])
    t_exp = (165 - 2) / -d[1]
    assert abs(t - t_exp) < 1e-12 * t_exp
    np.testing.assert_allclose(p, np.array([0, 165, 0]) + d * t_exp, rtol=1e-14)
    assert n == [0, 1, 0] and u == p[0] and v == p[2]
    # pointing up from above: rejected by sign before any arithmetic (geometry.d:35-36)
    assert not kat_intersect(0, [2, 0, 0, 0], [0, 165, 0], [0, 1, 0])[0]
    # farther than the current best: rejected, dist untouched
    hit, t, *_ = kat_intersect(0, [2, 0, 0, 0], [0, 165, 0], [0, -1, 0], max_dist=10)
    assert not hit and t == 10
    # below the plane looking up hits it from underneath
    assert kat_intersect(0, [2, 0, 0, 0], [0, -5, 0], [0, 1, 0])[0]


def test_sphere_closed_form_and_uv():
    hit, t, p, n, u, v = kat_intersect(1, [0, 0, 10, 2], [0, 0, 0], [0, 0, 1])
    assert hit and abs(t - 8) < 1e-14
    np.testing.assert_allclose(p, [0, 0, 8], atol=1e-14)
    np.testing.assert_allclose(n, [0, 0, -1], atol=1e-15)
    # u = (pi + atan2(dz, dx)) / 2pi with dz=-2, dx=0 -> atan2 = -pi/2 -> 0.25 ; v = 1 - (pi/2 + asin(0))/pi = 0.5
    assert abs(u - 0.25) < 1e-15 and abs(v - 0.5) < 1e-15
    # from inside: the far root
    hit, t, *_ = kat_intersect(1, [0, 0, 0, 5], [0, 0, 0], [1, 0, 0])
    assert hit and abs(t - 5) < 1e-14
    # behind the origin / missing
    assert not kat_intersect(1, [0, 0, -10, 2], [0, 0, 0], [0, 0, 1])[0]
    assert not kat_intersect(1, [5, 0, 10, 2], [0, 0, 0], [0, 0, 1])[0]


def test_cube_faces_and_uv_quirk():
    # +Y face hit from above: pass 1, u = p.x - c.x, v = p.z - c.z
    hit, t, p, n, u, v = kat_intersect(2, [0, 0, 0, 2], [0.25, 5, -0.5], [0, -1, 0])
    assert hit and abs(t - 4) < 1e-14 and n == [0, 1, 0] and abs(u - 0.25) < 1e-15 and abs(v + 0.5) < 1e-15
    # -X face hit from the left: found in the swapped frame; n and p are un-projected, u/v are NOT
    # (geometry.d:178-183): u = p.y - c.y, v = p.z - c.z
    hit, t, p, n, u, v = kat_intersect(2, [0, 0, 0, 2], [-5, 0.3, 0.6], [1, 0, 0])
    assert hit and abs(t - 4) < 1e-14 and n == [-1, 0, 0]
    np.testing.assert_allclose(p, [-1, 0.3, 0.6], atol=1e-14)
    assert abs(u - 0.3) < 1e-15 and abs(v - 0.6) < 1e-15
    # -Z face: u = p.x - c.x, v = p.y - c.y
    hit, t, p, n, u, v = kat_intersect(2, [0, 0, 0, 2], [0.1, -0.2, -7], [0, 0, 1])
    assert hit and abs(t - 6) < 1e-14 and n == [0, 0, -1] and abs(u - 0.1) < 1e-15 and abs(v + 0.2) < 1e-15
    # from inside: exit face
    hit, t, p, n, u, v = kat_intersect(2, [0, 0, 0, 2], [0, 0, 0], [0, 0, 1])
    assert hit and abs(t - 1) < 1e-14 and n == [0, 0, 1]
    assert not kat_intersect(2, [0, 0, 0, 2], [3, 3, -5], [0, 0, 1])[0]


def test_checker_parity_and_c_modulo():
    lib = oracle_lib()
    c1 = (C.c_float * 3)(0, 0, 0)
    c2 = (C.c_float * 3)(0, 0.5, 1.0)
    out = (C.c_float * 3)()

    def col(u, v, size=5.0):
        lib.orc_kat_checker(size, u, v, c1, c2, out)
        return list(out)

    assert col(1, 1) == [0, 0, 0]            # cells (0,0) -> even -> color1
    assert col(6, 1) == [0, 0.5, 1.0]        # (1,0) -> odd -> color2
    assert col(6, 6) == [0, 0, 0]            # (1,1)
    assert col(-1, 1) == [0, 0.5, 1.0]       # (-1,0): -1 % 2 == -1 in C/D -> non-zero -> color2
    assert col(-1, -1) == [0, 0, 0]          # (-1,-1) -> -2 % 2 == 0
    assert col(150, -230, 100) == [0, 0, 0]  # the example in texture.d:45-46: (1,-3) -> -2


def test_srgb_table_quirks():
    lut = srgb_lut()
    assert lut.shape == (4097,) and lut[0] == 0 and lut[4096] == 255
    assert np.all(np.diff(lut.astype(int)) >= 0)
    # linear segment uses 12.02 (not 12.92) and floor (color.d:200-201,216-219)
    i = 10  # x = 10/4096 = 0.00244 <= 0.0031308
    assert lut[i] == math.floor(np.float32(np.float32(i / 4096) * np.float32(12.02)) * np.float32(255.0))
    x = np.float32(2048 / 4096)
    assert lut[2048] == math.floor(np.float32(1.055 * float(x) ** (1 / 2.4) - 0.055) * np.float32(255.0))
    np.testing.assert_array_equal(lut, np.load(os.path.join(GOLD, "srgb_lut.npy")))
    # packing: r<<16 | g<<8 | b, clamped
    px = np.array([[[2.0, 0.5, -1.0]]], np.float32)
    assert pack_rgb32(px)[0, 0] == (255 << 16) | (int(lut[2048]) << 8) | 0


def d_shell_sort(keys):
    """pure-Python emulation of util/array.d:95-111 (ref loop index, strict >, gap sequence)"""
    arr = list(range(len(keys)))
    inc = len(arr) // 2
    while inc:
        key = 0
        while key < len(arr):
            i = key
            elem = arr[i]
            while i >= inc and keys[arr[i - inc]] > keys[elem]:
                arr[i] = arr[i - inc]
                i -= inc
            arr[i] = elem
            key = i + 1
        inc = 1 if inc == 2 else int(inc * 5.0 / 11)
    return arr


def test_shell_sort_matches_reference_algorithm():
    lib = oracle_lib()
    rng = np.random.default_rng(5)
    for n in range(0, 9):
        for _ in range(40):
            keys = rng.integers(0, 4, n).astype(np.float64)  # many ties
            perm = (C.c_int * max(n, 1))()
            lib.orc_kat_shell_sort(keys.ctypes.data_as(C.POINTER(C.c_double)), n, perm)
            got = list(perm)[:n]
            assert got == d_shell_sort(list(keys))
            assert all(keys[got[i]] <= keys[got[i + 1]] for i in range(n - 1))


@pytest.mark.parametrize("blob,size,pixels", [(KAT1, KAT1_SIZE, KAT1_PIXELS), (KAT2, KAT2_SIZE, KAT2_PIXELS)])
def test_bmp_reference_kats(blob, size, pixels):
    lib = oracle_lib()
    w, h = C.c_uint32(), C.c_uint32()
    out = np.zeros(16, np.uint32)
    buf = np.frombuffer(blob, np.uint8)
    assert lib.orc_kat_decode_bmp(buf.ctypes.data, len(blob), C.byref(w), C.byref(h), out.ctypes.data, out.size) == 0
    assert (w.value, h.value) == size
    assert list(out[: len(pixels)]) == pixels


def test_bundled_bitmaps_decode_like_pil():
    from PIL import Image
    s = OracleScene(os.path.join(SC, "lecture5.sdl"))
    for idx, name, gamma in [(0, "floor.bmp", True), (1, "world.bmp", True)]:
        tex = s.texture_texels(idx)
        im = np.asarray(Image.open(os.path.join(SC, name)).convert("RGB"), np.float32) * np.float32(1 / 255.0)
        assert tex.shape == im.shape
        x = im.astype(np.float64)
        lin = np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)
        lin = np.where(im == 0, 0.0, np.where(im == 1, 1.0, lin))
        np.testing.assert_allclose(tex, lin, atol=2e-7)


def test_json_and_sdl_lecture4_agree():
    a = OracleScene(os.path.join(SC, "lecture4.sdl"))
    a.override(aa=0, prepass=0)
    b = OracleScene(os.path.join(SC, "lecture4.json"))
    assert b.info()["aa"] == 0
    for s in (a, b):
        s.set_frame_size(96, 72)
    ia, sa = a.render(threads=2)
    ib, sb = b.render(threads=2)
    np.testing.assert_array_equal(ia, ib)
    assert (sa.primary_rays, sa.shadow_rays) == (sb.primary_rays, sb.shadow_rays) == (96 * 72, sb.shadow_rays)


def test_prepass_has_no_pixel_effect_and_threads_do_not_matter():
    a = OracleScene(os.path.join(SC, "lecture5.sdl"))
    a.set_frame_size(100, 75)
    i1, s1 = a.render(threads=1)
    a.override(prepass=1)
    i2, s2 = a.render(threads=3)
    np.testing.assert_array_equal(i1, i2)
    assert s2.prepass_rays > 0 and s1.prepass_rays == 0
    assert s1.primary_rays == 100 * 75 * 5 == s2.primary_rays


def test_rotate_is_applied_as_scale_quirk(tmp_path):
    base = open(os.path.join(ROOT, "tests", "scenes", "quirks.sdl")).read()
    # replace every `rotate a b c` by `scale a b c` after the node's own scale: identical image
    p1 = tmp_path / "a.sdl"
    p1.write_text(base.replace('"../../scenes/', '"' + SC + "/"))
    # build the same transform by hand: scale s then "rotate" r == scale s then scale r
    manual = base
    manual = manual.replace("rotate 1.5 0.5 1.5;", "scale 1.5 0.5 1.5;")
    manual = manual.replace("scale 1.2 0.9 1.2; rotate 1 1.1 1;", "scale 1.2 0.99 1.2;")
    p2 = tmp_path / "b.sdl"
    p2.write_text(manual.replace('"../../scenes/', '"' + SC + "/"))
    a, b = OracleScene(p1), OracleScene(p2)
    for s in (a, b):
        s.set_frame_size(80, 50)
    ia, _ = a.render()
    ib, _ = b.render()
    # 0.9*1.1 differs from 0.99 by one ulp -> allow rounding-level differences only
    assert np.abs(ia - ib).max() < 1e-5


def test_samples_at_pixel_corner():
    # renderer.d:306: the sample is taken at integer (x, y), no +0.5: pixel (0,0) looks along upLeft
    s = OracleScene(os.path.join(SC, "lecture4.json"))
    pos, ul = s.camera_vectors()[:2]
    rgb, hit = s.render_pixel(0, 479)
    d = ul - pos
    # bottom-left pixel hits the floor; reconstruct the hit from the corner formula
    v = s.camera_vectors()
    target = v[1] + (v[2] - v[1]) * (0 / 640) + (v[3] - v[1]) * (479 / 480)
    d = (target - pos) / np.linalg.norm(target - pos)
    t = (165 - 2) / -d[1]
    np.testing.assert_allclose(hit[2:5], pos + d * t, rtol=1e-13)
    assert hit[0] == 0


GI_SCENE = """Scene {{
  GlobalSettings {{ frameWidth 96; frameHeight 64; GIEnabled true; pathsPerPixel {paths}; maxTraceDepth 3; ambientLightColor 0.2 0.2 0.2 }}
  Camera {{ pos 0 60 -120; pitch -20; fov 70{cam} }}
  Lights {{ PointLight "l" {{ pos -40 120 -60; color 1 1 1; power 40000 }} PointLight "m" {{ pos 60 90 -20; color 0.4 0.5 1; power 9000 }} }}
  Geometries {{ Plane "floor" {{ y 0 }} Sphere "ball" {{ center 0 25 0; R 25 }} Cube "box" {{ center 50 15 10; side 30 }} }}
  Textures {{ Checker "chk" {{ color1 0.1 0.1 0.1; color2 0.9 0.9 0.9; size 10 }} }}
  Shaders {{ Lambert "f" {{ texture "chk" }} {ball_shader} "b" {{ color 0.9 0.2 0.2 }} }}
  Nodes {{ Node "n1" {{ geometry "floor"; shader "f" }} Node "n2" {{ geometry "ball"; shader "b" }} Node "n3" {{ geometry "box"; shader "f"; scale 1 1.5 1 }} }}
}}
"""


def test_gi_literal_walk_is_black_and_halts_on_phong(tmp_path):
    """renderSampleGI / pathtrace_impl (renderer.d:289-301,378-463) restated literally: with PointLights (solidAngle 0,
    light.d:72-75) every path sums to exactly black whatever the random walk; a Phong surface makes the reference halt."""
    p = tmp_path / "gi.sdl"
    p.write_text(GI_SCENE.format(paths=6, cam="", ball_shader="Lambert"))
    o = OracleScene(str(p))
    for seed, mode in ((0, 1), (7, 1), (0, 0)):   # pinned generator twice, libc rand once
        rgb, st = o.render(seed=seed, rng_mode=mode)
        assert np.all(rgb == 0.0)
        assert st.primary_rays == 96 * 64 * 5 * 6
        # every surface hit sends one shadow ray and spawns one continuation ray, up to maxTraceDepth
        assert st.shadow_rays == st.gi_bounce_rays and 0 < st.gi_bounce_rays <= st.primary_rays * 4
    p.write_text(GI_SCENE.format(paths=0, cam="", ball_shader="Lambert"))
    rgb, _ = OracleScene(str(p)).render()
    assert np.all(np.isnan(rgb))      # mean of zero paths: 0 / 0 (renderer.d:300)
    p.write_text(GI_SCENE.format(paths=2, cam="", ball_shader="Phong"))
    with pytest.raises(Exception, match="assert"):
        OracleScene(str(p)).render()
    # the DOF branch is tested first (renderer.d:256-259): GIEnabled is ignored under DOF
    p.write_text(GI_SCENE.format(paths=2, cam="; dof true; numSamples 2; focalPlaneDist 120; fNumber 8", ball_shader="Phong"))
    rgb, _ = OracleScene(str(p)).render()
    assert rgb.max() > 0.1


def test_patch_cone_keeps_every_node_a_camera_ray_hits():
    """The geometry behind the kernel's per-warp camera-ray node mask (render_kernel.cu patch_cone / cone_reaches_node), replayed in
    numpy with the same formulae: every camera ray of a warp's 8x4 pixel patch lies in the circular cone around the normalised sum of
    the patch's corner directions, so a node whose bounding sphere misses that cone cannot be the oracle's hit for any pixel of it."""
    import random
    W, H = 960, 540
    o = OracleScene(os.path.join(ROOT, "scenes", "chessboard.sdl"))
    o.set_frame_size(W, H)
    cv = o.camera_vectors()
    pos, ul, ur, dl = cv[0], cv[1], cv[2], cv[3]
    du, dv, ulrel = ur - ul, dl - ul, ul - pos
    centres = [None] + [np.array([(f - 3.5) * 20.0, 12.0, (rank - 4.5) * 20.0]) for rank in (1, 2, 7, 8) for f in range(8)]
    R = 24.0   # every piece of chess2rt_b200/chessboard.py fits in this sphere around (x, 12, z)

    def reaches(x0, y0, c, r):
        u = []
        for k in range(4):
            d = ulrel + du * ((x0 + (8 if k & 1 else -0.01)) / W) + dv * ((y0 + (4 if k & 2 else -0.01)) / H)
            u.append(d / np.linalg.norm(d))
        a = sum(u)
        a /= np.linalg.norm(a)
        cosphi = min(float(a @ uk) for uk in u)
        if not cosphi > 0.5:
            return True
        sinphi = np.sqrt(max(0.0, 1 - cosphi * cosphi))
        v = c - pos
        d2 = float(v @ v)
        d = np.sqrt(d2)
        rr = r * (1 + 1e-6) + 1e-6 * d
        if not d > rr:
            return True
        xa = float(v @ a)
        ya = np.sqrt(max(0.0, d2 - xa * xa))
        if xa * cosphi + ya * sinphi >= 0:
            return not (ya * cosphi - xa * sinphi > rr)
        return False

    rnd = random.Random(1)
    piece_hits, kept = 0, []
    for _ in range(1500):
        x, y = rnd.randrange(W), rnd.randrange(H)
        _, hit = o.render_pixel(x, y)
        node = int(hit[0])
        mask = [i for i in range(1, 33) if reaches(x // 8 * 8, y // 4 * 4, centres[i], R)]
        kept.append(len(mask))
        if node >= 1:
            piece_hits += 1
            assert node in mask, (x, y, node, mask)
    assert piece_hits > 200            # the sample does look at the pieces
    assert np.mean(kept) < 4           # and the cone does cull: 32 pieces, a handful per patch


def test_shadow_capsule_keeps_every_node_a_shadow_ray_can_touch():
    """The geometry behind the per-warp shadow-ray node mask (render_kernel.cu shadow_mask), replayed in float32 numpy with the
    same formulae: 32 shadow origins, their bounding box, the capsule (box centre -> light, radius = half diagonal) against node
    spheres.  Whenever ANY of the 32 exact FP64 segments origin -> light comes within a sphere's radius, the capsule test must
    keep that sphere (it may keep more); and it must cull most far-away spheres."""
    rnd = np.random.default_rng(5)
    f32 = np.float32
    kept_n, total_n, must = 0, 0, 0
    for trial in range(400):
        light = rnd.uniform([-300, 100, -300], [300, 400, 300])
        base = rnd.uniform([-150, 0, -150], [150, 40, 150])
        spread = 10 ** rnd.uniform(-2, 1.3)
        origins = base + rnd.uniform(-spread, spread, size=(32, 3))
        centres = rnd.uniform([-200, 0, -200], [200, 60, 200], size=(40, 3))
        radii = rnd.uniform(2, 30, size=40)
        # kernel side, FP32
        o32 = origins.astype(f32)
        mn, mx = o32.min(axis=0), o32.max(axis=0)
        c = f32(0.5) * (mn + mx)
        hdiag = mx - c
        rho = np.sqrt(np.dot(hdiag, hdiag), dtype=f32) * f32(1.000002)   # sqrt_up
        d = light.astype(f32) - c
        dd = np.dot(d, d)
        inv_dd = f32(1.0) / dd if dd > 0 else f32(0)
        mag = np.abs(c).sum(dtype=f32) + np.abs(d).sum(dtype=f32)          # L1 norms: cheaper, larger margins
        for k in range(40):
            b = centres[k].astype(f32)
            r = f32(radii[k])
            a = b - c
            t = np.clip(np.dot(a, d) * inv_dd, f32(0), f32(1))
            q = a - t * d
            reachr = r + rho + f32(8e-6) * (mag + np.abs(b).sum(dtype=f32) + rho)
            keep = not (np.dot(q, q) > reachr * reachr)
            # exact side, FP64: does any segment come within the sphere?
            touches = False
            for o in origins:
                seg = light - o
                tt = np.clip(np.dot(centres[k] - o, seg) / np.dot(seg, seg), 0.0, 1.0)
                if np.linalg.norm(centres[k] - (o + tt * seg)) <= radii[k]:
                    touches = True
                    break
            if touches:
                must += 1
                assert keep, (trial, k)
            kept_n += keep
            total_n += 1
    assert must > 100                   # the sample does contain occluders
    assert kept_n < 0.5 * total_n       # and the capsule does cull


def test_cubemap_environment_extension_face_and_uv_convention():
    """The cubemap environment is an EXTENSION (the reference's Environment is a black stub: environment.d:5-15).
    This pins what the oracle defines: face = largest |component| (x wins ties over y over z), the other two components divided
    by it give face coordinates in [-1, 1] -> texel coordinates in [0, size - 1] -> the ordinary bilinear fetch (bitmap.d:48-63),
    after the same load-time sRGB decode as BitmapTexture (texture.d:137-141)."""
    from PIL import Image
    o = OracleScene(os.path.join(ROOT, "tests", "scenes", "sky.sdl"))
    assert not OracleScene(os.path.join(SC, "lecture4.sdl")).environment((0, 1, 0))[0]       # the reference's black environment
    np.testing.assert_array_equal(OracleScene(os.path.join(SC, "lecture4.sdl")).environment((0.3, 0.5, 1))[1], 0)

    def srgb_decode(v8):   # bitmap.d:116-126 on x = v / 255 (color.d:60-66)
        x = np.float32(v8) / np.float32(255.0)
        if x == 0 or x == 1:
            return np.float32(x)
        if x <= np.float32(0.04045):
            return np.float32(x / np.float32(12.92))
        return np.float32(((np.float64(x) + np.float64(np.float32(0.055))) / np.float64(np.float32(1.055))) ** np.float64(np.float32(2.4)))

    faces = {n: np.asarray(Image.open(os.path.join(SC, "skybox", n + ".bmp")).convert("RGB")) for n in ("posx", "negx", "posy", "negy", "posz", "negz")}

    def expect(face, sx, sy):
        img = faces[face]
        h, w = img.shape[:2]
        x, y = np.float32((sx + 1) * 0.5 * (w - 1)), np.float32((sy + 1) * 0.5 * (h - 1))
        tx, ty = int(x), int(y)
        p, q = np.float32(x - np.float32(tx)), np.float32(y - np.float32(ty))
        out = np.zeros(3, np.float64)
        for (xx, yy, wgt) in ((tx, ty, (1 - p) * (1 - q)), ((tx + 1) % w, ty, p * (1 - q)), (tx, (ty + 1) % h, (1 - p) * q), ((tx + 1) % w, (ty + 1) % h, p * q)):
            out += np.array([srgb_decode(v) for v in img[yy, xx]], np.float64) * np.float64(wgt)
        return out

    cases = [((2, 0.5, -1), "posx", 0.5, -0.25), ((-4, 1, 2), "negx", 0.5, -0.25), ((0.3, 5, -1), "posy", 0.06, -0.2),
             ((0.3, -5, -1), "negy", 0.06, 0.2), ((1, -0.5, 4), "posz", 0.25, 0.125), ((1, -0.5, -4), "negz", -0.25, 0.125),
             ((1, 1, 0.2), "posx", -0.2, -1.0),      # tie |x| == |y|: x wins
             ((0.1, -3, 3), "negy", 0.1 / 3, -1.0)]  # tie |y| == |z|: y wins
    for d, face, sx, sy in cases:
        is_cube, rgb = o.environment(d)
        assert is_cube
        np.testing.assert_allclose(rgb, expect(face, sx, sy), rtol=2e-6, atol=2e-7, err_msg=str((d, face)))
    # scale invariance: only the direction matters
    np.testing.assert_array_equal(o.environment((0.2, 0.4, 1.0))[1], o.environment((0.4, 0.8, 2.0))[1])
    assert not o.environment((0, 0, 0))[1].any()


def test_golden_fixtures_reproduce():
    meta = json.load(open(os.path.join(GOLD, "golden.json")))
    for name, m in meta.items():
        s = OracleScene(os.path.join(ROOT, m["scene"]))
        s.set_frame_size(*m["size"])
        s.override(**m["override"])
        img, st = s.render(threads=2, seed=m["seed"])
        np.testing.assert_array_equal(img, np.load(os.path.join(GOLD, name + ".npy")), err_msg=name)
        assert st.primary_rays == m["primary_rays"] and st.shadow_rays == m["shadow_rays"]


def test_flop_counting_build():
    s = OracleScene(os.path.join(SC, "lecture4-proc-texture.sdl"), count_flops=True)
    s.set_frame_size(64, 48)
    img, st = s.render(threads=2)
    px = 64 * 48
    assert st.primary_rays == 5 * px
    per_px = st.flops / px
    assert 1200 < per_px < 2600, per_px  # SURVEY.md §8d hand estimate: ~1.8 kFLOP/px
    plain = OracleScene(os.path.join(SC, "lecture4-proc-texture.sdl"))
    plain.set_frame_size(64, 48)
    ref, st2 = plain.render(threads=2)
    np.testing.assert_array_equal(img, ref)   # counting does not change the arithmetic
    assert st2.flops == 0


def test_dof_pinned_rng_is_deterministic_and_matches_libc_statistics():
    s = OracleScene(os.path.join(SC, "zaphod.sdl"))
    s.set_frame_size(60, 40)
    s.override(num_samples=40)
    a, _ = s.render(threads=3, rng_mode=OracleScene.RNG_PINNED, seed=3)
    b, _ = s.render(threads=1, rng_mode=OracleScene.RNG_PINNED, seed=3)
    np.testing.assert_array_equal(a, b)
    c, _ = s.render(threads=1, rng_mode=OracleScene.RNG_PINNED, seed=4)
    assert np.abs(a - c).max() > 0
    d, _ = s.render(threads=1, rng_mode=OracleScene.RNG_LIBC)
    # same estimator, different random streams: image means agree closely
    assert abs(a.mean() - d.mean()) < 0.01 * a.mean()
    assert np.abs(a - d).mean() < 0.05 * a.mean()
