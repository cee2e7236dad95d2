"""Random scene generator for the parity fuzz tests: emits `.sdl` text over the whole surface the
render path supports (planes, spheres, cubes, all CSG ops incl. nesting, identity / translated /
scaled / "rotated" nodes, Lambert / Phong, checker / Procedure2 / bitmap textures, several lights
incl. a dark one, AA on/off, DOF, stereo).  Deterministic in `seed`."""
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SC = os.path.join(ROOT, "scenes")


def _v(r, lo, hi, n=3):
    return " ".join(f"{r.uniform(lo, hi):.6g}" for _ in range(n))


def generate(seed, width=96, height=64):
    r = random.Random(seed)
    out = []
    w = out.append
    w(f'Scene {{\n  Name "fuzz{seed}"\n')
    w("  GlobalSettings {\n")
    w(f"    frameWidth {width}\n    frameHeight {height}\n")
    w(f"    ambientLightColor {_v(r, 0, 0.3)}\n")
    w(f"    AAEnabled {'true' if r.random() < 0.6 else 'false'}\n    prepassEnabled false\n  }}\n")
    w("  Camera {\n")
    w(f"    pos {r.uniform(-40, 40):.6g} {r.uniform(60, 160):.6g} {r.uniform(-220, -120):.6g}\n")
    w(f"    yaw {r.uniform(-15, 15):.6g}\n    pitch {r.uniform(-45, -15):.6g}\n    roll {r.uniform(-8, 8):.6g}\n")
    w(f"    fov {r.uniform(50, 100):.6g}\n")
    if r.random() < 0.2:
        w(f"    dof true\n    numSamples {r.randint(2, 4)}\n    focalPlaneDist {r.uniform(120, 260):.6g}\n    fNumber {r.uniform(2, 16):.6g}\n")
    if r.random() < 0.15:
        w(f"    stereoSeparation {r.uniform(1, 8):.6g}\n")
    w("  }\n  Lights {\n")
    for i in range(r.randint(1, 3)):
        power = 0 if (i > 0 and r.random() < 0.2) else r.uniform(20000, 90000)
        w(f'    PointLight "l{i}" {{ pos {r.uniform(-200, 200):.6g} {r.uniform(120, 320):.6g} {r.uniform(-250, 150):.6g}; '
          f"color {_v(r, 0.3, 1)}; power {power:.6g} }}\n")
    w("  }\n  Geometries {\n")
    w(f'    Plane "floor" {{ y {r.uniform(-3, 3):.6g} }}\n')
    prims = []
    for i in range(r.randint(3, 7)):
        if r.random() < 0.5:
            w(f'    Sphere "p{i}" {{ center {_v(r, -12, 12)}; R {r.uniform(8, 28):.6g} }}\n')
        else:
            w(f'    Cube "p{i}" {{ center {_v(r, -12, 12)}; side {r.uniform(14, 46):.6g} }}\n')
        prims.append(f"p{i}")
    if r.random() < 0.3:
        w(f'    Plane "cutter" {{ y {r.uniform(-8, 8):.6g} }}\n')
        prims.append("cutter")
    geoms = list(prims)
    csgs = []
    for i in range(r.randint(1, 4)):
        op = r.choice(["CsgUnion", "CsgInter", "CsgDiff"])
        a, b = r.sample(prims, 2)
        w(f'    {op} "c{i}" {{ left "{a}"; right "{b}" }}\n')
        csgs.append(f"c{i}")
    nested = []
    if r.random() < 0.5 and csgs:
        for i in range(r.randint(1, 2)):
            op = r.choice(["CsgUnion", "CsgInter", "CsgDiff"])
            a = r.choice(csgs)
            b = r.choice(prims + csgs)
            if r.random() < 0.5:
                a, b = b, a
            w(f'    {op} "n{i}" {{ left "{a}"; right "{b}" }}\n')
            nested.append(f"n{i}")
    w("  }\n  Textures {\n")
    w(f'    Checker "chk" {{ color1 {_v(r, 0, 1)}; color2 {_v(r, 0, 1)}; size {r.uniform(3, 25):.6g} }}\n')
    w(f'    Procedure2 "proc" {{\n      freqU {_v(r, 0.01, 0.4)}\n      freqV {_v(r, 0.01, 0.4)}\n')
    w("      colorU { " + "; ".join(f"color {_v(r, 0, 0.8)}" for _ in range(3)) + " }\n")
    w("      colorV { " + "; ".join(f"color {_v(r, 0, 0.8)}" for _ in range(3)) + " }\n    }\n")
    w(f'    BitmapTexture "bmp1" {{ file "{SC}/floor.bmp"; scaling {r.uniform(0.004, 0.05):.6g} }}\n')
    w(f'    BitmapTexture "bmp2" {{ file "{SC}/world.bmp"; assumedGamma {r.choice([1, 2.2, 1.8])}; scaling {r.uniform(0.5, 2):.6g} }}\n')
    w("  }\n  Shaders {\n")
    texs = ["chk", "proc", "bmp1", "bmp2", None, None]
    shaders = []
    for i in range(6):
        t = r.choice(texs)
        tex = f'texture "{t}"; ' if t else ""
        if r.random() < 0.5:
            w(f'    Lambert "s{i}" {{ {tex}color {_v(r, 0.1, 1)} }}\n')
        else:
            e = r.choice([r.uniform(2, 120), r.uniform(2, 120), r.uniform(300, 1500), 5000.0])
            w(f'    Phong "s{i}" {{ {tex}color {_v(r, 0.1, 1)}; exponent {e:.6g}; strength {r.uniform(0.2, 2):.6g} }}\n')
        shaders.append(f"s{i}")
    w("  }\n  Nodes {\n")
    w(f'    Node "floor" {{ geometry "floor"; shader "{r.choice(shaders)}" }}\n')
    pool = prims[: max(1, len(prims) - 1)] + csgs + csgs + nested + nested
    for i in range(r.randint(3, 8)):
        g = r.choice(pool)
        if g == "cutter":
            g = prims[0]
        line = f'    Node "n{i}" {{ geometry "{g}"; shader "{r.choice(shaders)}"; '
        if r.random() < 0.4:
            line += f"scale {_v(r, 0.5, 1.8)}; "
        if r.random() < 0.2:
            line += f"rotate {_v(r, 0.7, 1.4)}; "
        line += f"translate {r.uniform(-140, 140):.6g} {r.uniform(5, 45):.6g} {r.uniform(-60, 160):.6g} }}\n"
        w(line)
    w("  }\n}\n")
    return "".join(out)


def generate_planes(seed, width=96, height=64):
    """Plane-only scenes (the MODE_SOLO and plane-only kernel classes): 1-3 horizontal planes, some scaled / translated,
    camera and lights above, below or between them — so planes shadow each other and the closest-hit order matters —
    every texture / shader kind, DOF and stereo now and then.  One plane + one light lands on the MODE_SOLO kernels."""
    r = random.Random(10_000 + seed)
    n_planes = r.choice([1, 1, 2, 3])
    n_lights = 1 if (n_planes == 1 and r.random() < 0.7) else r.randint(1, 3)
    ys = sorted(r.uniform(-40, 40) for _ in range(n_planes))
    out = []
    w = out.append
    w(f'Scene {{\n  Name "planes{seed}"\n  GlobalSettings {{\n    frameWidth {width}\n    frameHeight {height}\n')
    w(f"    ambientLightColor {_v(r, 0, 0.3)}\n    AAEnabled {'true' if r.random() < 0.6 else 'false'}\n    prepassEnabled false\n  }}\n")
    where = r.choice(["above", "above", "below", "between"])
    cam_y = {"above": ys[-1] + r.uniform(5, 120), "below": ys[0] - r.uniform(5, 120), "between": (ys[0] + ys[-1]) / 2 + 0.37}[where]
    pitch = {"above": -r.uniform(8, 60), "below": r.uniform(8, 60), "between": r.uniform(-60, 60)}[where]   # look at a plane
    w(f"  Camera {{\n    pos {r.uniform(-40, 40):.6g} {cam_y:.6g} {r.uniform(-100, 100):.6g}\n")
    w(f"    yaw {r.uniform(-180, 180):.6g}\n    pitch {pitch:.6g}\n    roll {r.uniform(-20, 20):.6g}\n    fov {r.uniform(40, 110):.6g}\n")
    if r.random() < 0.3:
        w(f"    dof true\n    numSamples {r.randint(2, 4)}\n    focalPlaneDist {r.uniform(20, 200):.6g}\n    fNumber {r.uniform(1, 16):.6g}\n")
    if r.random() < 0.15:
        w(f"    stereoSeparation {r.uniform(0.5, 6):.6g}\n")
    w("  }\n  Lights {\n")
    for i in range(n_lights):
        power = 0 if (i > 0 and r.random() < 0.2) else r.uniform(5000, 90000)
        ly = r.choice([cam_y + r.uniform(-3, 40), ys[-1] + r.uniform(1, 200), ys[0] - r.uniform(1, 200), (ys[0] + ys[-1]) / 2 - 0.21])
        w(f'    PointLight "l{i}" {{ pos {r.uniform(-150, 150):.6g} {ly:.6g} {r.uniform(-150, 150):.6g}; color {_v(r, 0.3, 1)}; power {power:.6g} }}\n')
    w("  }\n  Geometries {\n")
    for i, y in enumerate(ys):
        w(f'    Plane "pl{i}" {{ y {y:.6g} }}\n')
    w("  }\n  Textures {\n")
    w(f'    Checker "chk" {{ color1 {_v(r, 0, 1)}; color2 {_v(r, 0, 1)}; size {r.uniform(0.5, 25):.6g} }}\n')
    w(f'    Procedure2 "proc" {{\n      freqU {_v(r, 0.01, 2)}\n      freqV {_v(r, 0.01, 2)}\n')
    w("      colorU { " + "; ".join(f"color {_v(r, 0, 0.8)}" for _ in range(3)) + " }\n")
    w("      colorV { " + "; ".join(f"color {_v(r, 0, 0.8)}" for _ in range(3)) + " }\n    }\n")
    w(f'    BitmapTexture "bmp1" {{ file "{SC}/floor.bmp"; scaling {r.uniform(0.004, 0.05):.6g} }}\n')
    w(f'    BitmapTexture "bmp2" {{ file "{SC}/texture/zaphod.bmp"; assumedGamma {r.choice([1, 2.2, 1.8])}; scaling {r.uniform(0.01, 0.5):.6g} }}\n')
    w("  }\n  Shaders {\n")
    shaders = []
    for i in range(n_planes + 2):
        t = r.choice(["chk", "proc", "bmp1", "bmp2", None])
        tex = f'texture "{t}"; ' if t else ""
        if r.random() < 0.5:
            w(f'    Lambert "s{i}" {{ {tex}color {_v(r, 0.1, 1)} }}\n')
        else:
            e = r.choice([r.uniform(2, 120), r.uniform(300, 1500), 5000.0])
            w(f'    Phong "s{i}" {{ {tex}color {_v(r, 0.1, 1)}; exponent {e:.6g}; strength {r.uniform(0.2, 2):.6g} }}\n')
        shaders.append(f"s{i}")
    w("  }\n  Nodes {\n")
    for i in range(n_planes):
        line = f'    Node "n{i}" {{ geometry "pl{i}"; shader "{r.choice(shaders)}"; '
        if r.random() < 0.4:
            line += f"scale {_v(r, 0.3, 4)}; "
        if r.random() < 0.3:
            line += f"translate {r.uniform(-30, 30):.6g} {r.uniform(-3, 3):.6g} {r.uniform(-30, 30):.6g}; "
        w(line + "}\n")
    w("  }\n}\n")
    return "".join(out)


if __name__ == "__main__":
    import sys
    print(generate(int(sys.argv[1]) if len(sys.argv) > 1 else 0))
