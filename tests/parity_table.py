"""Prints the parity figures DESIGN.md quotes (GPU vs oracle, bundled scenes at their file resolution).
Run on a GPU box: python tests/parity_table.py > gpurun_out/parity_table.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess2rt_b200 as c2  # noqa: E402
from oracle_binding import OracleScene, parity_report  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("lecture4.sdl", {}), ("lecture4-proc-texture.sdl", {}), ("lecture5.sdl", {}), ("zaphod.sdl", {"num_samples": 6}),
         ("chessboard.sdl", {}), ("../tests/scenes/quirks.sdl", {}), ("../tests/scenes/nested.sdl", {}),
         ("../tests/scenes/stereo_dof.sdl", {})]


def main():
    c2.init(1, [0])
    out = {}
    for name, over in CASES:
        path = os.path.join(ROOT, "scenes", name)
        g, o = c2.HostScene(path), OracleScene(path)
        if name == "chessboard.sdl":
            g.set_frame_size(960, 540)
            o.set_frame_size(960, 540)
        g.override(**over)
        o.override(**over)
        rgb, argb, st = g.render(argb=True, seed=11, count_rays=True)
        ref, ost = o.render(seed=11)
        rep = parity_report(rgb, ref, argb)
        rep["rays_equal"] = (st.primary_rays, st.shadow_rays) == (ost.primary_rays, ost.shadow_rays)
        out[os.path.basename(name)] = rep
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
