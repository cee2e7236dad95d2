"""Generates the committed golden fixtures from the CPU oracle (oracle/liborc.so).

The reference has no golden vector for the render path and cannot be executed in
this image (no D toolchain), so these fixtures pin the ORACLE's output, not the reference's: they make
later edits of the oracle or of the CUDA path visible.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import OracleScene, srgb_lut  # noqa: E402

CASES = [  # name, scene path, (w, h), overrides, seed
    ("lecture4", "scenes/lecture4.sdl", (128, 96), {}, 0),
    ("lecture4_json", "scenes/lecture4.json", (128, 96), {}, 0),
    ("lecture4_proc", "scenes/lecture4-proc-texture.sdl", (128, 96), {}, 0),
    ("lecture5", "scenes/lecture5.sdl", (128, 96), {}, 0),
    ("zaphod_nodof", "scenes/zaphod.sdl", (129, 86), {"dof": 0}, 0),
    ("zaphod_dof", "scenes/zaphod.sdl", (129, 86), {"num_samples": 5}, 12345),
    ("chessboard", "scenes/chessboard.sdl", (128, 72), {}, 0),
    ("quirks", "tests/scenes/quirks.sdl", (160, 100), {}, 0),
    ("nested", "tests/scenes/nested.sdl", (160, 100), {}, 0),
    ("stereo", "tests/scenes/stereo.sdl", (128, 96), {}, 0),
    ("stereo_dof", "tests/scenes/stereo_dof.sdl", (129, 86), {}, 77),
    ("sky", "tests/scenes/sky.sdl", (160, 100), {}, 0),              # cubemap-environment EXTENSION (no reference counterpart)
    ("sky_plane", "tests/scenes/sky_plane.sdl", (161, 101), {}, 0),
]


def main():
    meta = {}
    for name, path, (w, h), over, seed in CASES:
        s = OracleScene(os.path.join(ROOT, path))
        s.set_frame_size(w, h)
        s.override(**over)
        img, st = s.render(threads=1, seed=seed)
        np.save(os.path.join(HERE, name + ".npy"), img)
        meta[name] = {"scene": path, "size": [w, h], "override": over, "seed": seed,
                      "primary_rays": st.primary_rays, "shadow_rays": st.shadow_rays}
        print(name, img.shape, float(img.mean()))
    np.save(os.path.join(HERE, "srgb_lut.npy"), srgb_lut())
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
