"""One-off extended parity fuzz on a GPU box (not collected by pytest): many more seeds than tests/test_gpu_fuzz.py, with the same
per-pixel classification of every outlier, plus the global-memory scene form on every 4th seed.
usage: python tests/fuzz_extended.py <first_seed> <last_seed>   -> one summary line per generator, details of any failure"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chess2rt_b200 as c2  # noqa: E402
from oracle_binding import OracleScene, pack_rgb32  # noqa: E402
from scene_fuzz import generate, generate_planes  # noqa: E402
from test_gpu_fuzz import assert_every_outlier_is_classified  # noqa: E402


def main():
    a, b = int(sys.argv[1]), int(sys.argv[2])
    c2.init(1, [0])
    tmp = tempfile.mkdtemp()
    for gen_name, gen in (("fuzz", generate), ("planes", generate_planes)):
        classes, outliers, pixels, failures, ray_mismatch = {}, 0, 0, [], 0
        for seed in range(a, b):
            path = os.path.join(tmp, f"{gen_name}{seed}.sdl")
            open(path, "w").write(gen(seed))
            if seed % 4 == 0:
                os.environ["C2RT_FORCE_GLOBAL"] = "1"
            else:
                os.environ.pop("C2RT_FORCE_GLOBAL", None)
            g, o = c2.HostScene(path), OracleScene(path)
            rgb, argb, st = g.render(argb=True, seed=seed, count_rays=True)
            ref, ost = o.render(threads=0, seed=seed)
            try:
                cl = assert_every_outlier_is_classified(o, rgb, ref, seed, gen_name)
                assert np.array_equal(argb, pack_rgb32(rgb))
                assert st.primary_rays == ost.primary_rays
            except AssertionError as e:
                failures.append((seed, str(e)[:300]))
                continue
            for k, v in cl.items():
                classes[k] = classes.get(k, 0) + v
            outliers += sum(cl.values())
            pixels += rgb.shape[0] * rgb.shape[1]
            ray_mismatch += int(st.shadow_rays != ost.shadow_rays)
            g.close()
        os.environ.pop("C2RT_FORCE_GLOBAL", None)
        print(f"{gen_name}: seeds {a}..{b - 1}: {b - a - len(failures)} scenes green, {pixels} pixels, {outliers} pixels over 1e-3 "
              f"(all classified: {classes}), scenes with a different shadow-ray count: {ray_mismatch}, FAILURES: {failures}", flush=True)


if __name__ == "__main__":
    main()
