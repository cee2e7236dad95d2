import sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import chess2rt_b200 as c2
from oracle_binding import OracleScene, parity_report
c2.init(1)
print('fma peak fp32', c2.measure_fma_peak(False), 'fp64', c2.measure_fma_peak(True), flush=True)
for name, kw in [('lecture4.sdl', {}), ('lecture4.json', {}), ('lecture4-proc-texture.sdl', {}), ('lecture5.sdl', {}), ('zaphod.sdl', {'dof': 0}), ('zaphod.sdl', {'num_samples': 4})]:
    p = '/root/repo/scenes/' + name
    hs = c2.HostScene(p); hs.override(**kw)
    t0 = time.time(); rgb, argb, st = hs.render(argb=True, seed=7, count_rays=True); t1 = time.time()
    rgb2, _, st2 = hs.render(argb=False, seed=7)
    os_ = OracleScene(p); os_.override(**kw)
    ref, ost = os_.render(seed=7)
    rep = parity_report(rgb, ref, argb)
    print(name, kw, 'kernel_ms', round(st.kernel_ms, 3), round(st2.kernel_ms, 3), 'total', round(st2.total_ms, 3), 'rays', st.primary_rays, st.shadow_rays, 'oracle', ost.primary_rays, ost.shadow_rays, round(ost.seconds, 3), json.dumps(rep), flush=True)
