"""Known answers DERIVED BY HAND from the reference's D formulas — not produced by oracle/ — so that a shared misreading of the
reference by the oracle and the CUDA path cannot hide (all other render-path evidence compares against the oracle).
Every expected number below follows from the cited lines with pencil-and-paper arithmetic on a 4x4 frame, a camera at (0, 5, 0)
looking down +z with no rotation and fov 90:

  camera.d:84-100  aspect = 4/4 = 1, corner = (-1, 1, 1), lenXY = sqrt(2), wantedLength = tan(45 deg) = 1, scaling = 1/sqrt(2)
                   -> upLeft = (-s, s, 1), upRight = (s, s, 1), downLeft = (-s, -s, 1) with s = 0.70710678..., + pos
  camera.d:139-146 pixel (x, y): target = upLeft + (upRight - upLeft) x/4 + (downLeft - upLeft) y/4; dir = normalize(target - pos)
     pixel (2, 2): target - pos = (0, 0, 1)                      -> dir = (0, 0, 1)
     pixel (2, 3): target - pos = (0, s - 2 s 3/4, 1) = (0, -s/2, 1), length = sqrt(1 + 1/8) = 3/(2 sqrt 2) -> dir = (0, -1/3, 2 sqrt(2)/3)
The CPU test checks the oracle, the GPU test checks libc2rt.so, both against the same literals."""
import math
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEAD = """Scene {{
  GlobalSettings {{ frameWidth 4; frameHeight 4; ambientLightColor 0.1 0.1 0.1; AAEnabled false; prepassEnabled false }}
  Camera {{ pos 0 5 0; yaw 0; pitch 0; roll 0; fov 90 }}
  Lights {{ PointLight "l" {{ pos {light}; color 1 1 1; power {power} }} }}
  Geometries {{ {geoms} }}
  Shaders {{ Lambert "s" {{ color 1 0.5 0.25 }} }}
  Nodes {{ Node "n" {{ geometry "{geom}"; shader "s" }} }}
}}
"""

S2 = math.sqrt(2.0)

# name -> (scene pieces, pixel, expected dist, p, normal, (u, v), rgb) with the derivation
CASES = {
    # geometry.d:30-59 Plane y = 0 from (0, 5, 0) along (0, -1/3, 2 sqrt2/3): t = (5 - 0) / (1/3) = 15, p = (0, 0, 10 sqrt 2), n = (0, 1, 0), (u, v) = (p.x, p.z).
    # shader.d:67-105 Lambert: light straight above p at height 10: |L - p|^2 = 100, cos = 1 -> L = ambient 0.1 + 50/100 = 0.6; colour = diffuse * 0.6
    "plane": (dict(light="0 10 %.17g" % (10 * S2), power=50, geoms='Plane "g" { y 0 }', geom="g"), (2, 3),
              15.0, (0.0, 0.0, 10 * S2), (0.0, 1.0, 0.0), (0.0, 10 * S2), (0.6, 0.3, 0.15)),
    # geometry.d:92-125 Sphere centre (0, 5, 20), R = 5 along (0, 0, 1): H = (0, 0, -20), A = 1, B = -40, C = 375, D = 100 -> t = (40 - 10)/2 = 15;
    # p = (0, 5, 15), n = (0, 0, -1); u = (pi + atan2(-5, 0)) / 2pi = 1/4, v = 1 - (pi/2 + asin(0)) / pi = 1/2.
    # light at the camera: |L - p|^2 = 225, cos = 1 -> 0.1 + 112.5/225 = 0.6
    "sphere": (dict(light="0 5 0", power=112.5, geoms='Sphere "g" { center 0 5 20; R 5 }', geom="g"), (2, 2),
               15.0, (0.0, 5.0, 15.0), (0.0, 0.0, -1.0), (0.25, 0.5), (0.6, 0.3, 0.15)),
    # geometry.d:172-235 Cube centre (0, 5, 20), side 10: the Z pass (swap of y and z, :186-190) hits the face z = 15 at t = 15; n = (0, 0, -1);
    # u, v are left in the permuted frame (quirk, :224-230): u = p'.x - c'.x = 0, v = p'.z - c'.z = p.y - c.y = 0
    "cube": (dict(light="0 5 0", power=112.5, geoms='Cube "g" { center 0 5 20; side 10 }', geom="g"), (2, 2),
             15.0, (0.0, 5.0, 15.0), (0.0, 0.0, -1.0), (0.0, 0.0), (0.6, 0.3, 0.15)),
    # geometry.d:271-332,382-397 CsgDiff(cube above, sphere centre (0, 5, 15) R = 3).  Crossings along (0, 0, 1): sphere in at 12, cube in at 15, sphere
    # out at 18, cube out at 25.  findAllIntersections restarts 1e-6 past each crossing and never adds the offset back (:283-286), so a child's 2nd
    # crossing is recorded 1e-6 short: 18 - 1e-6 (and 25 - 1e-6).  Walk with inL = inR = false: 12 -> inR; 15 -> inL, L && !R false; 18 - 1e-6 -> !inR:
    # L && !R true -> hit: dist = 18 - 1e-6, p = the sphere's exit point (0, 5, 18), the sphere's outward normal (0, 0, 1) FLIPPED by :394-395
    # (right.isInside differs 1e-6 before / after p) -> (0, 0, -1).  Sphere uv at Delta = (0, 0, 3): u = (pi + pi/2) / 2pi = 3/4, v = 1/2.
    # The shadow ray runs back through the carved hole (never inside the solid): lit; |L - p|^2 = 324 (to 1e-7), cos = 1 -> 0.1 + 162/324 = 0.6
    "csg_diff": (dict(light="0 5 0", power=162, geoms='Cube "c" { center 0 5 20; side 10 }; Sphere "h" { center 0 5 15; R 3 }; CsgDiff "g" { left "c"; right "h" }', geom="g"),
                 (2, 2), 18.0 - 1e-6, (0.0, 5.0, 18.0), (0.0, 0.0, -1.0), (0.75, 0.5), (0.6, 0.3, 0.15)),
}


def write_scene(tmp_path, name):
    p = tmp_path / (name + ".sdl")
    p.write_text(HEAD.format(**CASES[name][0]))
    return str(p)


def check(name, rgb, node, dist, p, n, uv, uv_atol=1e-11):
    _, _, e_dist, e_p, e_n, e_uv, e_rgb = CASES[name]
    assert node == 0, name
    np.testing.assert_allclose(dist, e_dist, rtol=1e-12, err_msg=name)
    np.testing.assert_allclose(p, e_p, rtol=0, atol=1e-11, err_msg=name)
    np.testing.assert_allclose(n, e_n, rtol=0, atol=1e-12, err_msg=name)
    np.testing.assert_allclose(uv, e_uv, rtol=0, atol=uv_atol, err_msg=name)
    np.testing.assert_allclose(rgb, e_rgb, rtol=0, atol=2e-6, err_msg=name)   # FP32 colour arithmetic (color.d:27-35)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_hand_derived_hits(name, tmp_path):
    from oracle_binding import OracleScene
    o = OracleScene(write_scene(tmp_path, name))
    x, y = CASES[name][1]
    rgb, hit = o.render_pixel(x, y)
    check(name, rgb, int(hit[0]), hit[1], hit[2:5], hit[5:8], hit[8:10])
    # and the full frame agrees with the pixel pick (renderer.d:223-228 vs :46-57)
    frame, _ = o.render()
    np.testing.assert_array_equal(frame[y, x], rgb)


def test_oracle_camera_vectors_match_hand_derivation(tmp_path):
    from oracle_binding import OracleScene
    s = 1 / S2
    v = OracleScene(write_scene(tmp_path, "plane")).camera_vectors()   # pos, upLeft, upRight, downLeft, right, up, front
    np.testing.assert_allclose(v, [[0, 5, 0], [-s, 5 + s, 1], [s, 5 + s, 1], [-s, 5 - s, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_hand_derived_hits(name, tmp_path):
    import chess2rt_b200 as c2
    c2.init(1, [0])
    g = c2.HostScene(write_scene(tmp_path, name))
    x, y = CASES[name][1]
    rgb, hit = g.render_pixel(x, y)
    # (sphere uv is the one FP32 quantity of the hit record on the GPU: atan2f on the FP64 difference vector)
    check(name, rgb, hit.node, hit.dist, list(hit.p), list(hit.normal), [hit.u, hit.v], uv_atol=1e-6)
    frame, _, _ = g.render()
    np.testing.assert_allclose(frame[y, x], CASES[name][6], rtol=0, atol=2e-6)


# ---------------------------------------------------------------- gfm conventions, pinned without recalling gfm
# gfm:math is not under /root/reference: host/rt.cpp and oracle/orc_math.hpp both RESTATE rotateX/Y/Z.  These
# checks do not depend on that recollection: they follow from the reference's own sources —
#   imported_types.d:13-20  mul(v, M) is row-vector x matrix, M.c[row][col]
#   camera.d:102-112        rotation = rotateZ(roll) * rotateX(pitch) * rotateY(yaw); rightDir / upDir / frontDir = e_i * rotation
#   raytracer_demo.d:276-301 the key bindings: RIGHT+SHIFT applies dYaw = -4 and must turn the view RIGHT, RIGHT+CTRL applies
#                            dRoll = +4 ("roll right"), UP+SHIFT applies dPitch = +4 (look up); :316 mouse right -> yaw decreases
# and from rotations being rotations (orthonormal, det +1, angles add).
CAM = """Scene {{ GlobalSettings {{ frameWidth 8; frameHeight 6 }}
  Camera {{ pos 1 2 3; yaw {yaw}; pitch {pitch}; roll {roll}; fov 70 }} }}
"""


def camera_vectors(kind, tmp_path, yaw=0.0, pitch=0.0, roll=0.0):
    p = tmp_path / "cam.sdl"
    p.write_text(CAM.format(yaw=yaw, pitch=pitch, roll=roll))
    if kind == "oracle":
        from oracle_binding import OracleScene
        return OracleScene(str(p)).camera_vectors()
    import chess2rt_b200 as c2
    cam, _ = c2.HostScene(str(p)).frame_blocks()
    return np.array([list(cam.pos), list(cam.up_left), list(cam.up_right), list(cam.down_left), list(cam.right_dir), list(cam.up_dir), list(cam.front_dir)])


@pytest.mark.parametrize("kind", ["oracle", "host"])
def test_camera_rotation_conventions_follow_the_key_bindings(kind, tmp_path):
    d = math.radians(4.0)
    right, up, front = camera_vectors(kind, tmp_path)[4:7]
    np.testing.assert_allclose([right, up, front], np.eye(3), atol=1e-15)
    # RIGHT+SHIFT: yaw -4 turns the view to the right (+x): front = (sin 4, 0, cos 4), right = (cos 4, 0, -sin 4)
    right, up, front = camera_vectors(kind, tmp_path, yaw=-4.0)[4:7]
    np.testing.assert_allclose(front, [math.sin(d), 0, math.cos(d)], atol=1e-15)
    np.testing.assert_allclose(right, [math.cos(d), 0, -math.sin(d)], atol=1e-15)
    # UP+SHIFT: pitch +4 looks up: front = (0, sin 4, cos 4); lecture4's pitch -30 looks down at the floor
    front = camera_vectors(kind, tmp_path, pitch=4.0)[6]
    np.testing.assert_allclose(front, [0, math.sin(d), math.cos(d)], atol=1e-15)
    np.testing.assert_allclose(camera_vectors(kind, tmp_path, pitch=-30.0)[6], [0, -0.5, math.sqrt(3) / 2], atol=1e-15)
    # RIGHT+CTRL: roll +4 rolls to the right: the up vector leans towards +x
    right, up, front = camera_vectors(kind, tmp_path, roll=4.0)[4:7]
    np.testing.assert_allclose(up, [math.sin(d), math.cos(d), 0], atol=1e-15)
    np.testing.assert_allclose(front, [0, 0, 1], atol=1e-15)
    # order of camera.d:102-104 on a row vector: roll first, then pitch, then yaw.  yaw 90 (left), pitch -30 (down):
    # (0,0,1) -rotX(-30)-> (0, -1/2, sqrt3/2) -rotY(90)-> (-sqrt3/2, -1/2, 0)
    np.testing.assert_allclose(camera_vectors(kind, tmp_path, yaw=90.0, pitch=-30.0)[6], [-math.sqrt(3) / 2, -0.5, 0], atol=1e-15)


@pytest.mark.parametrize("kind", ["oracle", "host"])
def test_camera_basis_is_a_rotation_and_corners_are_consistent(kind, tmp_path):
    rnd = np.random.default_rng(3)
    for _ in range(20):
        yaw, pitch, roll = rnd.uniform(-180, 180), rnd.uniform(-90, 90), rnd.uniform(-180, 180)
        v = camera_vectors(kind, tmp_path, yaw, pitch, roll)
        pos, ul, ur, dl, right, up, front = v
        R = np.array([right, up, front])
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-14)        # orthonormal
        np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-14)     # a proper rotation (no reflection)
        # corners (camera.d:84-100,106-116): (+-x, +-y, 1) * rotation + pos with x = aspect * k, y = k, k = tan(fov/2) / |(-aspect, 1)|
        aspect, k = 8 / 6, math.tan(math.radians(35.0)) / math.hypot(8 / 6, 1.0)
        np.testing.assert_allclose(ul - pos, -aspect * k * right + k * up + front, atol=1e-14)
        np.testing.assert_allclose(ur - pos, aspect * k * right + k * up + front, atol=1e-14)
        np.testing.assert_allclose(dl - pos, -aspect * k * right - k * up + front, atol=1e-14)
    # angles add: yaw a then yaw b equals yaw a + b (rotateY is a one-parameter group)
    np.testing.assert_allclose(camera_vectors(kind, tmp_path, yaw=25.0)[4:7] @ camera_vectors(kind, tmp_path, yaw=40.0)[4:7],
                               camera_vectors(kind, tmp_path, yaw=65.0)[4:7], atol=1e-14)


def test_node_transform_inverse_and_transpose_are_consistent():
    """transform.d:32-41: scale() keeps inverseTransform = transform.inverse(), transposedInverse = inverseTransform.transposed():
    M * Minv = I and MinvT = Minv^T for every node of the bundled scenes (quirks.sdl holds scaled and 'rotated' nodes)."""
    import chess2rt_b200 as c2
    for path in ("tests/scenes/quirks.sdl", "scenes/zaphod.sdl", "scenes/lecture5.sdl"):
        scene = c2.HostScene(os.path.join(ROOT, path))   # (keep it alive: desc() borrows its arrays)
        d = scene.desc().contents
        for i in range(d.n_nodes):
            M = np.array([d.node_transform[9 * i + k] for k in range(9)]).reshape(3, 3)
            Mi = np.array([d.node_inverse[9 * i + k] for k in range(9)]).reshape(3, 3)
            MiT = np.array([d.node_inverse_t[9 * i + k] for k in range(9)]).reshape(3, 3)
            np.testing.assert_allclose(M @ Mi, np.eye(3), atol=1e-15)
            np.testing.assert_array_equal(MiT, Mi.T)


# ---------------------------------------------------------------- shading and textures, derived by hand
# Same 4x4 frame and camera.  Pixel (3, 3): target - pos = upLeft + (upRight - upLeft) 3/4 + (downLeft - upLeft) 3/4 - pos = (s/2, -s/2, 1),
# s = 1/sqrt(2); on the plane y = 0 the ray from (0, 5, 0) arrives at t with 5 = t (s/2)/len, i.e. p = (5, 0, 5 / (s/2)) = (5, 0, 10 sqrt 2)
# whatever the length: (u, v) = (p.x, p.z) = (5, 10 sqrt 2) (geometry.d:49-55).  The light sits 10 above p: |L - p|^2 = 100, lightDir = N =
# (0, 1, 0), cosTheta = 1, baseLight = color * power / 100 = 0.5 (shader.d:67-105,197-250; light.d:11-14), so lightContrib = ambient 0.1 + 0.5 = 0.6.
V2 = 10 * S2
SHADE_HEAD = """Scene {{
  GlobalSettings {{ frameWidth 4; frameHeight 4; ambientLightColor 0.1 0.1 0.1; AAEnabled false; prepassEnabled false }}
  Camera {{ pos 0 5 0; yaw 0; pitch 0; roll 0; fov 90 }}
  Lights {{ PointLight "l" {{ pos 5 10 %.17g; color 1 1 1; power 50 }} }}
  Geometries {{ Plane "g" {{ y 0 }} }}
  Textures {{ {textures} }}
  Shaders {{ {shader} }}
  Nodes {{ Node "n" {{ geometry "g"; shader "s" }} }}
}}
""" % V2

_Q = float(np.float32(np.float32(V2 * 0.1 - 1.0) * np.float32(2)))   # bitmap case: ty = float(v') * height, v' = frac(0.1 * 10 sqrt 2)

SHADE_CASES = {
    # shader.d:197-250 Phong: R = reflect(-lightDir, N) = (0, -1, 0) - 2 (-1) (0, 1, 0) = (0, 1, 0) (imported_types.d:62-67);
    # -ray.dir = -(s/2, -s/2, 1) / sqrt(1.25): cosGamma = (s/2) / sqrt(1.25) = sqrt(0.1); specular = baseLight * cosGamma^2 * strength
    # = 0.5 * 0.1 * 0.5 = 0.025; colour = diffuse * 0.6 + 0.025
    "phong": (dict(textures="", shader='Phong "s" { color 1 0.5 0.25; exponent 2; strength 0.5 }'),
              (0.6 + 0.025, 0.3 + 0.025, 0.15 + 0.025)),
    # texture.d:36-54 Checker, size 5: x = floor(5 / 5) = 1, y = floor(14.142 / 5) = 2, (1 + 2) % 2 = 1 -> color2; size 4: x = 1, y = 3 -> 0 -> color1
    "checker_color2": (dict(textures='Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 5 }', shader='Lambert "s" { color 1 1 1; texture "t" }'),
                       (0.6 * 1.0, 0.6 * 0.9, 0.6 * 0.8)),
    "checker_color1": (dict(textures='Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 4 }', shader='Lambert "s" { color 1 1 1; texture "t" }'),
                       (0.6 * 0.2, 0.6 * 0.4, 0.6 * 0.6)),
    # texture.d:77-86 Procedure2: sum_i colorU[i] sin(u freqU[i]) + colorV[i] sin(v freqV[i]) at (u, v) = (5, 10 sqrt 2)
    "procedure2": (dict(textures='Procedure2 "t" { freqU 0.1 0.2 0.3; freqV 0.1 0.2 0.3; '
                                 'colorU { color 1 0 0; color 0 1 0; color 0 0 1 }; colorV { color 0.5 0.5 0.5; color 0.25 0 0; color 0 0 0.125 } }',
                        shader='Lambert "s" { color 1 1 1; texture "t" }'),
                   tuple(0.6 * c for c in (
                       math.sin(0.5) + 0.5 * math.sin(0.1 * V2) + 0.25 * math.sin(0.2 * V2),
                       math.sin(1.0) + 0.5 * math.sin(0.1 * V2),
                       math.sin(1.5) + 0.5 * math.sin(0.1 * V2) + 0.125 * math.sin(0.3 * V2)))),
    # texture.d:116-126 + bitmap.d:48-63 BitmapTexture, scaling 0.1, assumedGamma 1 (no remap, texture.d:137-141), on the 2x2 bitmap written by
    # _write_bmp (top row red, green; bottom row blue, white): u' = frac(0.5) = 0.5 -> tx = 1.0, tx_next = (1 + 1) % 2 = 0, p = 0;
    # v' = frac(1.41421...) -> ty = float(v') * 2 = 0.828427..., ty = 0, ty_next = 1, q = 0.828427...;
    # colour = data[1, 0] (1 - q) + data[1, 1] q = green (1 - q) + white q = (q, 1, q)
    "bitmap": (dict(textures='BitmapTexture "t" { file "t.bmp"; scaling 0.1; assumedGamma 1 }', shader='Lambert "s" { color 1 1 1; texture "t" }'),
               (0.6 * _Q, 0.6, 0.6 * _Q)),
}


def _write_bmp(path):
    """2x2, 24 bpp, BITMAPINFOHEADER; rows bottom-up (imageio/bmp.d:130-150), pixels B, G, R, rows padded to 4 bytes."""
    import struct
    rows = bytes([255, 0, 0, 255, 255, 255, 0, 0]) + bytes([0, 0, 255, 0, 255, 0, 0, 0])   # bottom: blue, white; top: red, green
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, 2, 2, 1, 24, 0, len(rows), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + rows)


def write_shade_scene(tmp_path, name):
    _write_bmp(str(tmp_path / "t.bmp"))
    p = tmp_path / (name + ".sdl")
    p.write_text(SHADE_HEAD.format(**SHADE_CASES[name][0]))
    return str(p)


@pytest.mark.parametrize("name", sorted(SHADE_CASES))
def test_oracle_reproduces_hand_derived_shading(name, tmp_path):
    from oracle_binding import OracleScene
    o = OracleScene(write_shade_scene(tmp_path, name))
    rgb, hit = o.render_pixel(3, 3)
    assert int(hit[0]) == 0
    np.testing.assert_allclose(hit[2:5], (5.0, 0.0, V2), rtol=0, atol=1e-11)
    np.testing.assert_allclose(hit[8:10], (5.0, V2), rtol=0, atol=1e-11)
    np.testing.assert_allclose(rgb, SHADE_CASES[name][1], rtol=0, atol=2e-6, err_msg=name)
    frame, _ = o.render()
    np.testing.assert_allclose(frame[3, 3], SHADE_CASES[name][1], rtol=0, atol=2e-6, err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SHADE_CASES))
def test_gpu_reproduces_hand_derived_shading(name, tmp_path):
    import chess2rt_b200 as c2
    c2.init(1, [0])
    g = c2.HostScene(write_shade_scene(tmp_path, name))
    rgb, hit = g.render_pixel(3, 3)
    assert hit.node == 0
    np.testing.assert_allclose(list(hit.p), (5.0, 0.0, V2), rtol=0, atol=1e-11)
    # (Procedure2: six SFU sines of a phase kept to 2^-23 revolution, each good to ~1e-6: DESIGN.md section 4.2)
    atol = 1e-5 if name == "procedure2" else 3e-6
    np.testing.assert_allclose(rgb, SHADE_CASES[name][1], rtol=0, atol=atol, err_msg=name)
    frame, _, _ = g.render()   # one plane, one light: the MODE_SOLO kernels (Procedure2 through sin_phase, the bitmap through its palette quads)
    np.testing.assert_allclose(frame[3, 3], SHADE_CASES[name][1], rtol=0, atol=atol, err_msg=name)


# ---------------------------------------------------------------- a whole (tiny) frame, restated independently in Python
# A third implementation, written from the D sources alone and sharing nothing with oracle/ or the CUDA path: camera.d:84-117,139-146
# (corners, screen ray), renderer.d:223-251 (corner sample + the four AA taps, accum / 5), geometry.d:30-59 (plane), texture.d:36-54
# (checker), shader.d:67-105 (Lambert, faceforward, shadow-ray origin p + N 1e-6), scene.d:62-78 (visibility), light.d:11-14
# (color * power), color.d:122-132 (FP32 colour arithmetic).  4x4 frame, fov 90, camera (0, 5, 0) looking down +z: rows 0 and 1 miss,
# row 2 sits ON the horizon (its corner sample misses, some of its AA taps hit the plane far away), row 3 hits.
FRAME_SCENE = """Scene {
  GlobalSettings { frameWidth 4; frameHeight 4; ambientLightColor 0.1 0.2 0.3; AAEnabled true; prepassEnabled false }
  Camera { pos 0 5 0; yaw 0; pitch 0; roll 0; fov 90 }
  Lights { PointLight "l" { pos 1 10 12; color 1 0.9 0.8; power 300 } }
  Geometries { Plane "g" { y 0 } }
  Textures { Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 3 } }
  Shaders { Lambert "s" { color 1 1 1; texture "t" } }
  Nodes { Node "n" { geometry "g"; shader "s" } }
}
"""


def python_frame():
    f32 = np.float32
    W = H = 4
    pos = np.array([0.0, 5.0, 0.0])
    light = np.array([1.0, 10.0, 12.0])
    light_color = np.array([f32(1) * f32(300), f32(0.9) * f32(300), f32(0.8) * f32(300)], dtype=f32)   # light.d:11-14
    ambient = np.array([0.1, 0.2, 0.3], dtype=f32)
    c1, c2 = np.array([0.2, 0.4, 0.6], dtype=f32), np.array([1, 0.9, 0.8], dtype=f32)
    # camera.d:84-100
    x, y = -(W / H), 1.0
    len_xy = math.hypot(x, y)
    scaling = math.tan(math.radians(90.0 / 2)) / len_xy
    x, y = x * scaling, y * scaling
    up_left, up_right, down_left = np.array([x, y, 1.0]) + pos, np.array([-x, y, 1.0]) + pos, np.array([x, -y, 1.0]) + pos

    def sample(sx, sy):
        target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)    # camera.d:139-146
        d = target - pos
        d = d / math.sqrt(d @ d)
        if d[1] > -1e-9:                                                                           # geometry.d:35 (origin above the plane)
            return np.zeros(3, f32)                                                                # environment.d:7-10
        t = (pos[1] - 0.0) / -d[1]
        p = pos + d * t
        n = np.array([0.0, 1.0, 0.0])                                                              # faceforward: d . n < 0
        cx, cz = int(math.floor(p[0] / 3.0)), int(math.floor(p[2] / 3.0))                          # texture.d:47-48
        white = int(math.fmod(cx + cz, 2))                                                         # :50, D's % truncates: -1 is true as well
        diffuse = c2 if white else c1
        contrib = ambient.copy()
        frm = p + n * 1e-6
        # scene.d:62-78: the only node is the plane; from is above it and the light higher still: dir.y > -1e-9 -> no hit -> visible
        assert frm[1] > 0 and light[1] > frm[1]
        ld = light - p
        dist2 = ld @ ld
        cos_theta = (ld / math.sqrt(dist2)) @ n
        if cos_theta > 0:
            contrib = contrib + (light_color / f32(dist2)) * f32(cos_theta)                        # shader.d:97, FP32 Color ops
        return (diffuse * contrib).astype(f32)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = sample(px, py)
            for kx, ky in ((0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6)):                        # renderer.d:235-247
                acc = acc + sample(px + kx, py + ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_frame(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "frame.sdl"
    p.write_text(FRAME_SCENE)
    want = python_frame()
    assert (want[:2] == 0).all() and (want[2] > 0).any() and (want[2] < want[3].max()).all() and (want[3] > 0).all()
    got, _ = OracleScene(str(p)).render()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_gpu_matches_the_independent_python_frame(tmp_path):
    import chess2rt_b200 as c2
    c2.init(1, [0])
    p = tmp_path / "frame.sdl"
    p.write_text(FRAME_SCENE)
    got, _, _ = c2.HostScene(str(p)).render()
    np.testing.assert_allclose(got, python_frame(), rtol=0, atol=3e-6)


# ---------------------------------------------------------------- the same, with spheres, shadows, two lights and Phong
# The independent Python restatement extended by geometry.d:92-125 (sphere: the quadratic as written, closer root first), node.d:23-49
# (translated node: origin minus offset, identity matrix), renderer.d:325-376 (closest hit over the nodes in scene order),
# scene.d:62-78 (shadow ray: any node hit closer than the light), shader.d:197-250 (Phong) and imported_types.d:62-73
# (reflect, faceforward).  8x6 anti-aliased frame of a checkered floor with three spheres sharing one geometry — one as it is, one a
# translated node, one scaled: node.d:84-93 applies `scale`, then `rotate` AS A SECOND SCALE (the reference's quirk), so the matrix is
# diag(1.5 * 1, 0.6 * 2, 1.2 * 0.5) about the object-space origin (transform.d:24-86, imported_types.d:13-20: row vector x matrix) —
# lit by two lights: shadows on the floor and on a sphere, highlights, silhouettes, horizon.
SPHERES_SCENE = """Scene {
  GlobalSettings { frameWidth 8; frameHeight 6; ambientLightColor 0.05 0.1 0.15; AAEnabled true; prepassEnabled false }
  Camera { pos 0 6 -4; yaw 0; pitch 0; roll 0; fov 80 }
  Lights {
    PointLight "l0" { pos -6 14 4; color 1 0.9 0.8; power 500 }
    PointLight "l1" { pos 8 9 10; color 0.3 0.5 1; power 400 }
  }
  Geometries { Plane "floor" { y 0 }; Sphere "ball" { center 0 3 12; R 3 } }
  Textures { Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 2.5 } }
  Shaders { Lambert "fs" { color 1 1 1; texture "t" }; Phong "bs" { color 0.8 0.3 0.2; exponent 12; strength 0.7 }; Lambert "cs" { color 0.1 0.9 0.4 } }
  Nodes {
    Node "n0" { geometry "floor"; shader "fs" }
    Node "n1" { geometry "ball"; shader "bs" }
    Node "n2" { geometry "ball"; shader "cs"; translate 5 1 -3 }
    Node "n3" { geometry "ball"; shader "cs"; scale 1.5 0.6 1.2; rotate 1 2 0.5; translate -5 -1 3 }
  }
}
"""


def python_spheres_frame():
    f32 = np.float32
    W, H = 8, 6
    pos = np.array([0.0, 6.0, -4.0])
    lights = [(np.array([-6.0, 14.0, 4.0]), np.array([f32(1) * f32(500), f32(0.9) * f32(500), f32(0.8) * f32(500)], f32)),
              (np.array([8.0, 9.0, 10.0]), np.array([f32(0.3) * f32(400), f32(0.5) * f32(400), f32(1) * f32(400)], f32))]
    ambient = np.array([0.05, 0.1, 0.15], f32)
    c1, c2 = np.array([0.2, 0.4, 0.6], f32), np.array([1, 0.9, 0.8], f32)
    centre, R = np.array([0.0, 3.0, 12.0]), 3.0
    # nodes in scene order: (kind, offset, shader, diagonal of the transform)
    one = np.ones(3)
    nodes = [("plane", np.zeros(3), "fs", one), ("sphere", np.zeros(3), "bs", one), ("sphere", np.array([5.0, 1.0, -3.0]), "cs", one),
             ("sphere", np.array([-5.0, -1.0, 3.0]), "cs", np.array([1.5 * 1.0, 0.6 * 2.0, 1.2 * 0.5]))]
    x, y = -(W / H), 1.0
    scaling = math.tan(math.radians(80.0 / 2)) / math.hypot(x, y)
    x, y = x * scaling, y * scaling
    up_left, up_right, down_left = np.array([x, y, 1.0]) + pos, np.array([-x, y, 1.0]) + pos, np.array([x, -y, 1.0]) + pos

    def hit_node(kind, off, o, d, best, scale=one):
        """-> (dist, p, normal) or None; (o, d) in world space, d unit; node.d:23-49"""
        if scale is not one:                                          # a diagonal matrix: undoPoint / undoDirection divide, point multiplies,
            o2, d2 = (o - off) / scale, d / scale                     # normal goes through the transposed inverse (divide again)
            ln = math.sqrt(d2 @ d2)
            r = hit_node(kind, np.zeros(3), o2, d2 / ln, best * ln)
            if r is None:
                return None
            n = r[2] / scale
            return r[0] / ln, r[1] * scale + off, n / math.sqrt(n @ n)
        o = o - off
        if kind == "plane":                                           # geometry.d:30-59, y = 0
            if (o[1] > 0 and d[1] > -1e-9) or (o[1] < 0 and d[1] < 1e-9):
                return None
            t = o[1] / -d[1]
            if t > best:
                return None
            return t, o + d * t + off, np.array([0.0, 1.0, 0.0])
        h = o - centre                                                # geometry.d:92-125
        a, b, c = d @ d, 2 * (h @ d), h @ h - R * R
        dscr = b * b - 4 * a * c
        if dscr < 0:
            return None
        x1, x2 = (-b + math.sqrt(dscr)) / (2 * a), (-b - math.sqrt(dscr)) / (2 * a)
        sol = x2 if x2 >= 0 else x1
        if sol < 0 or sol > best:
            return None
        p = o + d * sol
        n = p - centre
        return sol, p + off, n / math.sqrt(n @ n)

    def visible(frm, to):                                             # scene.d:62-78
        d = to - frm
        dist = math.sqrt(d @ d)
        d = d / dist
        return not any(hit_node(k, off, frm, d, dist, sc) for k, off, _, sc in nodes)

    def sample(sx, sy):
        target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)
        d = target - pos
        d = d / math.sqrt(d @ d)
        best, rec = 1e99, None
        for k, off, sh, sc in nodes:                                  # renderer.d:336-338
            r = hit_node(k, off, pos, d, best, sc)
            if r:
                best, rec = r[0], (r[1], r[2], sh)
        if rec is None:
            return np.zeros(3, f32)
        p, n, sh = rec
        if not d @ n < 0:                                             # faceforward
            n = -n
        if sh == "fs":
            white = int(math.fmod(int(math.floor(p[0] / 2.5)) + int(math.floor(p[2] / 2.5)), 2))
            diffuse = c2 if white else c1
        else:
            diffuse = np.array([0.8, 0.3, 0.2], f32) if sh == "bs" else np.array([0.1, 0.9, 0.4], f32)
        contrib, spec = ambient.copy(), np.zeros(3, f32)
        for lp, lc in lights:
            if not visible(p + n * 1e-6, lp):
                continue
            ld = lp - p
            dist2 = ld @ ld
            ld = ld / math.sqrt(dist2)
            cos_theta = ld @ n
            base = lc / f32(dist2)
            if cos_theta > 0:
                contrib = contrib + base * f32(cos_theta)
            if sh == "bs":                                            # shader.d:235-241
                r = -ld - 2 * (-ld @ n) * n
                r = r / math.sqrt(r @ r)
                cos_gamma = r @ -d
                if cos_gamma > 0:
                    spec = spec + base * f32(cos_gamma ** 12) * f32(0.7)
        return (diffuse * contrib + spec).astype(f32)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = sample(px, py)
            for kx, ky in ((0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6)):
                acc = acc + sample(px + kx, py + ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_frame_with_spheres_and_shadows(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "spheres.sdl"
    p.write_text(SPHERES_SCENE)
    want = python_spheres_frame()
    got, st = OracleScene(str(p)).render()
    assert st.shadow_rays < 2 * st.primary_rays and st.shadow_rays > st.primary_rays // 2   # hits, two lights each
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)


# ---------------------------------------------------------------- the same, with cubes and CSG
# The independent Python restatement extended by geometry.d:172-235 (cube: the Y, X, Z side passes through project / unproject,
# imported_types.d:44-60) and geometry.d:271-402 (CSG: findAllIntersections with its 1e-6 restarts whose offset is never added back,
# the merged list shell-sorted by util/array.d:95-111 — `foreach (ref i, ...)` with the index moved inside the loop — the walk that
# flips inL / inR by `current.g is left`, boolOp per class, CsgDiff's normal flip), written from the D text alone.  10x8 frame:
# a floor, lecture5's cube-minus-sphere, a translated sphere-and-cube intersection, a union of two spheres; one light.
CSG_SCENE = """Scene {
  GlobalSettings { frameWidth 10; frameHeight 8; ambientLightColor 0.1 0.1 0.1; AAEnabled true; prepassEnabled false }
  Camera { pos 0 9 -10; yaw 0; pitch 0; roll 0; fov 85 }
  Lights { PointLight "l" { pos -7 20 -2; color 1 1 1; power 700 } }
  Geometries {
    Plane "floor" { y 0 }
    Cube "c" { center -4 3 8; side 6 }
    Sphere "s" { center -4 3 8; R 3.9 }
    CsgDiff "diff" { left "c"; right "s" }
    Sphere "s2" { center 0 0 0; R 3 }
    Cube "c2" { center 0.5 0.5 0; side 4.4 }
    CsgInter "inter" { left "s2"; right "c2" }
    Sphere "s3" { center 5 2.5 14; R 2.5 }
    Sphere "s4" { center 7 4 13; R 2 }
    CsgUnion "union" { left "s3"; right "s4" }
  }
  Shaders { Lambert "f" { color 0.7 0.7 0.7 }; Lambert "a" { color 0.9 0.6 0.1 }; Lambert "b" { color 0.2 0.6 0.9 }; Lambert "u" { color 0.4 0.9 0.3 } }
  Nodes {
    Node "n0" { geometry "floor"; shader "f" }
    Node "n1" { geometry "diff"; shader "a" }
    Node "n2" { geometry "inter"; shader "b"; translate 4 3 5 }
    Node "n3" { geometry "union"; shader "u" }
  }
}
"""


class _Hit:
    def __init__(self, dist, p, n, g):
        self.dist, self.p, self.n, self.g = dist, p, n, g


class _Plane:
    def inside(self, p):
        return False

    def isect(self, o, d, best):
        if (o[1] > 0 and d[1] > -1e-9) or (o[1] < 0 and d[1] < 1e-9):
            return None
        t = o[1] / -d[1]
        if t > best:
            return None
        return _Hit(t, o + d * t, np.array([0.0, 1.0, 0.0]), self)


class _Sphere:
    def __init__(self, c, r):
        self.c, self.r = np.array(c, float), r

    def inside(self, p):
        v = self.c - p
        return v @ v < self.r * self.r

    def isect(self, o, d, best):
        h = o - self.c
        a, b, c = d @ d, 2 * (h @ d), h @ h - self.r * self.r
        dscr = b * b - 4 * a * c
        if dscr < 0:
            return None
        x1, x2 = (-b + math.sqrt(dscr)) / (2 * a), (-b - math.sqrt(dscr)) / (2 * a)
        sol = x2 if x2 >= 0 else x1
        if sol < 0 or sol > best:
            return None
        p = o + d * sol
        n = p - self.c
        return _Hit(sol, p, n / math.sqrt(n @ n), self)


class _Cube:
    def __init__(self, c, side):
        self.c, self.side = np.array(c, float), side

    def inside(self, p):
        return bool((np.abs(p - self.c) <= self.side * 0.5).all())

    def _side(self, o, d, c, best):
        """intersectCubeSide on (o, d, c) in the pass's frame -> (dist, p, normal) in that frame or None"""
        if abs(d[1]) < 1e-9:
            return None
        half, found = self.side * 0.5, None
        for s in (-1, 1):
            mult = (o[1] - (c[1] + s * half)) / -d[1]
            if mult < 0 or mult > best:
                continue
            p = o + d * mult
            if p[0] < c[0] - half or p[0] > c[0] + half or p[2] < c[2] - half or p[2] > c[2] + half:
                continue
            best, found = mult, (mult, p, np.array([0.0, float(s), 0.0]))
        return found

    def isect(self, o, d, best):
        res = None
        for perm in ((0, 1, 2), (1, 0, 2), (0, 2, 1)):       # Y sides; X sides (swap x, y); Z sides (swap y, z): each permutation is its own inverse
            idx = list(perm)
            r = self._side(o[idx], d[idx], self.c[idx], best)
            if r:
                best = r[0]
                res = _Hit(r[0], r[1][idx], r[2][idx], self)
        return res


class _Csg:
    def __init__(self, op, left, right):
        self.op, self.left, self.right = op, left, right

    def bool_op(self, l, r):
        return (l or r) if self.op == "union" else (l and r) if self.op == "inter" else (l and not r)

    def inside(self, p):
        return self.bool_op(self.left.inside(p), self.right.inside(p))

    @staticmethod
    def _all(geom, o, d):
        out, cur = [], 0.0
        while True:
            h = geom.isect(o, d, 1e99)
            if h is None:
                return out
            h.dist += cur
            cur = h.dist
            o = h.p + d * 1e-6
            out.append(h)

    def isect(self, o, d, best):
        ld, rd = self._all(self.left, o, d), self._all(self.right, o, d)
        arr = ld + rd
        n, inc = len(arr), len(arr) // 2
        while inc:                                              # util/array.d:95-111
            i = 0
            while i < n:
                elem = arr[i]
                while i >= inc and arr[i - inc].dist > elem.dist:
                    arr[i] = arr[i - inc]
                    i -= inc
                arr[i] = elem
                i += 1
            inc = 1 if inc == 2 else int(inc * 5.0 / 11)
        in_l, in_r = len(ld) % 2 == 1, len(rd) % 2 == 1
        for cur in arr:
            if cur.g is self.left:
                in_l = not in_l
            else:
                in_r = not in_r
            if self.bool_op(in_l, in_r):
                if cur.dist > best:
                    return None
                hit = _Hit(cur.dist, cur.p, cur.n, cur.g)
                if self.op == "diff" and self.right.inside(hit.p - d * 1e-6) != self.right.inside(hit.p + d * 1e-6):
                    hit.n = -hit.n
                return hit
        return None


def python_csg_frame():
    f32 = np.float32
    W, H = 10, 8
    pos, light = np.array([0.0, 9.0, -10.0]), np.array([-7.0, 20.0, -2.0])
    lc = np.array([f32(1) * f32(700)] * 3, f32)
    ambient = np.array([0.1, 0.1, 0.1], f32)
    diff = _Csg("diff", _Cube((-4, 3, 8), 6), _Sphere((-4, 3, 8), 3.9))
    inter = _Csg("inter", _Sphere((0, 0, 0), 3), _Cube((0.5, 0.5, 0), 4.4))
    union = _Csg("union", _Sphere((5, 2.5, 14), 2.5), _Sphere((7, 4, 13), 2))
    nodes = [(_Plane(), np.zeros(3), (0.7, 0.7, 0.7)), (diff, np.zeros(3), (0.9, 0.6, 0.1)),
             (inter, np.array([4.0, 3.0, 5.0]), (0.2, 0.6, 0.9)), (union, np.zeros(3), (0.4, 0.9, 0.3))]
    x, y = -(W / H), 1.0
    scaling = math.tan(math.radians(85.0 / 2)) / math.hypot(x, y)
    x, y = x * scaling, y * scaling
    up_left, up_right, down_left = np.array([x, y, 1.0]) + pos, np.array([-x, y, 1.0]) + pos, np.array([x, -y, 1.0]) + pos

    def node_hit(node, o, d, best):            # node.d:23-49, identity matrix: the ray moves by -offset, the point comes back by +offset
        g, off, _ = node
        h = g.isect(o - off, d, best)
        if h:
            h.p = h.p + off
        return h

    def visible(frm, to):
        d = to - frm
        dist = math.sqrt(d @ d)
        d = d / dist
        return not any(node_hit(nd, frm, d, dist) for nd in nodes)

    def sample(sx, sy):
        target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)
        d = target - pos
        d = d / math.sqrt(d @ d)
        best, rec = 1e99, None
        for nd in nodes:
            h = node_hit(nd, pos, d, best)
            if h:
                best, rec = h.dist, (h, nd)
        if rec is None:
            return np.zeros(3, f32)
        h, nd = rec
        n = h.n if d @ h.n < 0 else -h.n
        contrib = ambient.copy()
        if visible(h.p + n * 1e-6, light):
            ld = light - h.p
            dist2 = ld @ ld
            cos_theta = (ld / math.sqrt(dist2)) @ n
            if cos_theta > 0:
                contrib = contrib + (lc / f32(dist2)) * f32(cos_theta)
        return (np.array(nd[2], f32) * contrib).astype(f32)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = sample(px, py)
            for kx, ky in ((0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6)):
                acc = acc + sample(px + kx, py + ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_frame_with_cubes_and_csg(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "csg.sdl"
    p.write_text(CSG_SCENE)
    want = python_csg_frame()
    got, st = OracleScene(str(p)).render()
    # every CSG node is on screen: its solid colour dominates some pixel (orange, blue, green: largest channel r, b, g well above the floor's grey)
    assert (want[..., 0] > 1.3 * want[..., 2]).any() and (want[..., 2] > 1.3 * want[..., 0]).any() and (want[..., 1] > 1.3 * want[..., 0]).any()
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)


# ---------------------------------------------------------------- the same, with depth of field
# renderer.d:270-287 (renderSampleDof: per sample the pixel jitter x + uniform * dx, y + uniform * dy, then the ray) and camera.d:154-173,
# 258-269 (focal point T = orig + dir * focalPlaneDist / (dir . front); unitDiscSample: angle = uniform * 2 pi, rad = sqrt(uniform),
# (sin, cos) * rad * discMultiplier along right / up; dir = normalize(T - orig)), restated in Python.  The reference draws from libc rand();
# oracle and library draw from the pinned generator c2rt_rng_u31 keyed by (seed, pixel, AA tap, sample, draw), uniform = value / RAND_MAX
#  — called here through the library's host function, in the order the D code consumes its draws.
DOF_SCENE = """Scene {
  GlobalSettings { frameWidth 4; frameHeight 4; ambientLightColor 0.1 0.2 0.3; AAEnabled true; prepassEnabled false }
  Camera { pos 0 5 0; yaw 0; pitch -30; roll 0; fov 90; dof true; numSamples 3; focalPlaneDist 12; fNumber 8 }
  Lights { PointLight "l" { pos 1 10 12; color 1 0.9 0.8; power 300 } }
  Geometries { Plane "g" { y 0 } }
  Textures { Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 3 } }
  Shaders { Lambert "s" { color 1 1 1; texture "t" } }
  Nodes { Node "n" { geometry "g"; shader "s" } }
}
"""


def python_dof_frame(seed, cam):
    """`cam`: pos, upLeft, upRight, downLeft, right, up, front (the camera basis is pinned separately above: rotation conventions)"""
    import chess2rt_b200 as c2
    f32 = np.float32
    W = H = 4
    pos, up_left, up_right, down_left, right, up, front = [np.array(v, float) for v in cam]
    light = np.array([1.0, 10.0, 12.0])
    light_color = np.array([f32(1) * f32(300), f32(0.9) * f32(300), f32(0.8) * f32(300)], dtype=f32)
    ambient = np.array([0.1, 0.2, 0.3], dtype=f32)
    c1, c2c = np.array([0.2, 0.4, 0.6], dtype=f32), np.array([1, 0.9, 0.8], dtype=f32)
    focal, disc, n_samples = 12.0, 10.0 / 8.0, 3

    def shade(o, d):
        if (o[1] > 0 and d[1] > -1e-9) or (o[1] < 0 and d[1] < 1e-9):
            return np.zeros(3, f32)
        t = o[1] / -d[1]
        p = o + d * t
        n = np.array([0.0, 1.0, 0.0]) if d[1] < 0 else np.array([0.0, -1.0, 0.0])
        white = int(math.fmod(int(math.floor(p[0] / 3.0)) + int(math.floor(p[2] / 3.0)), 2))
        contrib = ambient.copy()
        frm = p + n * 1e-6
        sd = light - frm
        sd = sd / math.sqrt(sd @ sd)
        if not ((frm[1] > 0 and sd[1] > -1e-9) or (frm[1] < 0 and sd[1] < 1e-9)):   # the plane between the point and the light
            return ((c2c if white else c1) * contrib).astype(f32)
        ld = light - p
        dist2 = ld @ ld
        cos_theta = (ld / math.sqrt(dist2)) @ n
        if cos_theta > 0:
            contrib = contrib + (light_color / f32(dist2)) * f32(cos_theta)
        return ((c2c if white else c1) * contrib).astype(f32)

    def sample(px, py, tap, kx, ky):
        avg = np.zeros(3, f32)
        for i in range(n_samples):
            u = [c2.rng_u31(seed, px, py, tap, i, k) / 2147483647.0 for k in range(4)]
            sx, sy = px + kx + u[0] * 1, py + ky + u[1] * 1
            target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)
            d = target - pos
            d = d / math.sqrt(d @ d)
            T = pos + d * (focal / (d @ front))
            angle, rad = u[2] * 2 * math.pi, math.sqrt(u[3])
            o = pos + (math.sin(angle) * rad * disc) * right + (math.cos(angle) * rad * disc) * up
            d = T - o
            avg = avg + shade(o, d / math.sqrt(d @ d))
        return avg / f32(n_samples)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = np.zeros(3, f32)
            for tap, (kx, ky) in enumerate(((0.0, 0.0), (0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6))):
                acc = acc + sample(px, py, tap, kx, ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_frame_with_depth_of_field(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "dof.sdl"
    p.write_text(DOF_SCENE)
    o = OracleScene(str(p))
    got, st = o.render(seed=77)
    assert st.primary_rays == 4 * 4 * 5 * 3
    want = python_dof_frame(77, o.camera_vectors())
    assert (want > 0).any()
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)


# ---------------------------------------------------------------- the same, with a bitmap-textured sphere (lecture5's globe)
# geometry.d:114-120 (sphere uv: u = (pi + atan2(dz, dx)) / 2 pi, v = 1 - (pi/2 + asin(dy / R)) / pi), texture.d:116-141 (BitmapTexture:
# scaling, wrap to [0, 1), float texel coordinates, assumedGamma 2.2 -> sRGB decompression at load time), bitmap.d:48-63,105-126 (bilinear
# fetch with wrap; the decompression formula), color.d:60-66 (byte / 255 in FP32), restated in Python; the 8-bpp world.bmp is decoded by
# PIL, an independent decoder.
GLOBE_SCENE = """Scene {
  GlobalSettings { frameWidth 8; frameHeight 6; ambientLightColor 0.2 0.2 0.2; AAEnabled true; prepassEnabled false }
  Camera { pos 0 3 -9; yaw 0; pitch 0; roll 0; fov 60 }
  Lights { PointLight "l" { pos -8 12 -10; color 1 1 1; power 400 } }
  Geometries { Sphere "globe" { center 0.5 3 2; R 4 } }
  Textures { BitmapTexture "world" { file "%s" } }
  Shaders { Lambert "s" { color 1 1 1; texture "world" } }
  Nodes { Node "n" { geometry "globe"; shader "s" } }
}
""" % os.path.join(ROOT, "scenes", "world.bmp")


def python_globe_frame():
    from PIL import Image
    f32 = np.float32
    W, H = 8, 6
    tex = np.asarray(Image.open(os.path.join(ROOT, "scenes", "world.bmp")).convert("RGB")).astype(f32) * f32(1.0 / 255.0)   # color.d:60-66
    lin = np.where(tex <= f32(0.04045), tex / f32(12.92), np.power((tex + f32(0.055)) / f32(1.055), f32(2.4))).astype(f32)   # bitmap.d:115-126
    lin[tex == 0] = 0
    lin[tex == 1] = 1
    th, tw = lin.shape[:2]
    pos, light = np.array([0.0, 3.0, -9.0]), np.array([-8.0, 12.0, -10.0])
    lc = np.array([f32(400)] * 3, f32)
    ambient = np.array([0.2, 0.2, 0.2], f32)
    centre, R = np.array([0.5, 3.0, 2.0]), 4.0
    x, y = -(W / H), 1.0
    scaling = math.tan(math.radians(60.0 / 2)) / math.hypot(x, y)
    x, y = x * scaling, y * scaling
    up_left, up_right, down_left = np.array([x, y, 1.0]) + pos, np.array([-x, y, 1.0]) + pos, np.array([x, -y, 1.0]) + pos

    def texel(u, v):
        u, v = u - math.floor(u), v - math.floor(v)                      # scaling 1 (texture.d:118-122)
        fx, fy = f32(u) * f32(tw), f32(v) * f32(th)
        if int(fx) >= tw or int(fy) >= th:
            return np.array([1, 0, 0], f32)                              # NamedColors.red (bitmap.d:50-51)
        tx, ty = int(math.floor(fx)), int(math.floor(fy))
        txn, tyn = (tx + 1) % tw, (ty + 1) % th
        p, q = fx - f32(tx), fy - f32(ty)
        one = f32(1)
        return (lin[ty, tx] * ((one - p) * (one - q)) + lin[ty, txn] * (p * (one - q)) + lin[tyn, tx] * ((one - p) * q) + lin[tyn, txn] * (p * q)).astype(f32)

    def sample(sx, sy):
        target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)
        d = target - pos
        d = d / math.sqrt(d @ d)
        h = pos - centre
        a, b, c = d @ d, 2 * (h @ d), h @ h - R * R
        dscr = b * b - 4 * a * c
        if dscr < 0:
            return np.zeros(3, f32)
        sol = (-b - math.sqrt(dscr)) / (2 * a)
        if sol < 0:
            sol = (-b + math.sqrt(dscr)) / (2 * a)
        if sol < 0:
            return np.zeros(3, f32)
        p = pos + d * sol
        n = (p - centre) / math.sqrt((p - centre) @ (p - centre))
        angle = math.atan2(p[2] - centre[2], p[0] - centre[0])
        u = (math.pi + angle) / (2 * math.pi)
        v = 1.0 - (math.pi / 2 + math.asin((p[1] - centre[1]) / R)) / math.pi
        if not d @ n < 0:
            n = -n
        contrib = ambient.copy()
        frm = p + n * 1e-6                                               # visible unless the sphere itself is in the way
        sd = light - frm
        sdist = math.sqrt(sd @ sd)
        sd = sd / sdist
        hh = frm - centre
        bb, cc = 2 * (hh @ sd), hh @ hh - R * R
        ds = bb * bb - 4 * cc
        blocked = False
        if ds >= 0:
            s2, s1 = (-bb - math.sqrt(ds)) / 2, (-bb + math.sqrt(ds)) / 2
            s = s2 if s2 >= 0 else s1
            blocked = 0 <= s <= sdist
        if not blocked:
            ld = light - p
            dist2 = ld @ ld
            cos_theta = (ld / math.sqrt(dist2)) @ n
            if cos_theta > 0:
                contrib = contrib + (lc / f32(dist2)) * f32(cos_theta)
        return (texel(u, v) * contrib).astype(f32)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = sample(px, py)
            for kx, ky in ((0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6)):
                acc = acc + sample(px + kx, py + ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_frame_of_a_textured_globe(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "globe.sdl"
    p.write_text(GLOBE_SCENE)
    want = python_globe_frame()
    got, _ = OracleScene(str(p)).render()
    assert (want > 0.05).mean() > 0.3 and want.std() > 0.02          # the globe fills a good part of the frame, with texture detail
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-6)


# ---------------------------------------------------------------- the same, with the stereo anaglyph
# renderer.d:303-312 (renderSampleDefault with stereoSeparation != 0: one ray per eye, the origin moved by -/+ stereoSeparation along
# rightDir, camera.d:148-152), color.d:10-15 (combineStereo: both eyes desaturated to 0.25, left -> red, right -> green and blue) and
# color.d:76-82,139-142 (adjustSaturation around the mean of the three channels), restated in Python.
STEREO_SCENE = """Scene {
  GlobalSettings { frameWidth 4; frameHeight 4; ambientLightColor 0.1 0.2 0.3; AAEnabled true; prepassEnabled false }
  Camera { pos 0 5 0; yaw 0; pitch -30; roll 0; fov 90; stereoSeparation 1.5 }
  Lights { PointLight "l" { pos 1 10 12; color 1 0.9 0.8; power 300 } }
  Geometries { Plane "g" { y 0 } }
  Textures { Checker "t" { color1 0.2 0.4 0.6; color2 1 0.9 0.8; size 3 } }
  Shaders { Lambert "s" { color 1 1 1; texture "t" } }
  Nodes { Node "n" { geometry "g"; shader "s" } }
}
"""


def python_stereo_frame(cam):
    f32 = np.float32
    W = H = 4
    pos, up_left, up_right, down_left, right, up, front = [np.array(v, float) for v in cam]
    light = np.array([1.0, 10.0, 12.0])
    light_color = np.array([f32(1) * f32(300), f32(0.9) * f32(300), f32(0.8) * f32(300)], dtype=f32)
    ambient = np.array([0.1, 0.2, 0.3], dtype=f32)
    c1, c2c = np.array([0.2, 0.4, 0.6], dtype=f32), np.array([1, 0.9, 0.8], dtype=f32)

    def shade(o, d):
        if d[1] > -1e-9:                          # both eyes are above the plane
            return np.zeros(3, f32)
        p = o + d * (o[1] / -d[1])
        white = int(math.fmod(int(math.floor(p[0] / 3.0)) + int(math.floor(p[2] / 3.0)), 2))
        ld = light - p
        dist2 = ld @ ld
        contrib = ambient + (light_color / f32(dist2)) * f32(ld[1] / math.sqrt(dist2))     # light above the plane: visible, cos > 0
        return ((c2c if white else c1) * contrib).astype(f32)

    def desat(c):
        mid = (c[0] + c[1] + c[2]) / f32(3)
        return (c * f32(0.25) + mid * (f32(1) - f32(0.25))).astype(f32)

    def sample(sx, sy):
        target = up_left + (up_right - up_left) * (sx / W) + (down_left - up_left) * (sy / H)
        d = target - pos
        d = d / math.sqrt(d @ d)
        left, rgt = desat(shade(pos - right * 1.5, d)), desat(shade(pos + right * 1.5, d))
        return np.array([left[0], rgt[1], rgt[2]], f32)          # left * (1, 0, 0) + right * (0, 1, 1)

    img = np.zeros((H, W, 3), f32)
    for py in range(H):
        for px in range(W):
            acc = sample(px, py)
            for kx, ky in ((0.3, 0.3), (0.6, 0.0), (0.0, 0.6), (0.6, 0.6)):
                acc = acc + sample(px + kx, py + ky)
            img[py, px] = acc / f32(5)
    return img


def test_oracle_matches_the_independent_python_stereo_frame(tmp_path):
    from oracle_binding import OracleScene
    p = tmp_path / "stereo.sdl"
    p.write_text(STEREO_SCENE)
    o = OracleScene(str(p))
    got, st = o.render()
    assert st.primary_rays == 4 * 4 * 5 * 2
    want = python_stereo_frame(o.camera_vectors())
    assert (np.abs(want[..., 0] - want[..., 1]) > 1e-3).any()    # the two eyes do see different things
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6)
