"""Tuning aid: kernel time of ONE rank's share of a frame on one GPU (c2rt_render_device with a band, no gather, no flags),
for n_ranks in 1, 2, 4, 8 — what a perfectly overlapped gather could reach.  usage: python profiles/band_time.py <workload>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import chess2rt_b200 as c2  # noqa: E402
from chess2rt_b200 import api  # noqa: E402
from bench import RNG_SEED, WORKLOADS  # noqa: E402

name = sys.argv[1]
path, W, H, over = WORKLOADS[name]
c2.init(1, [0])
scene = c2.HostScene(os.path.join(ROOT, path))
scene.set_frame_size(W, H)
handle = scene.device_scene()
cam, st = scene.frame_blocks(seed=RNG_SEED)
frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for n in (1, 2, 4, 8):
    res = []
    for rank in range(n):
        band = api.Band(rank, n, 8, 0, None, 0, 0)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)]
        for e0, e1 in ev:
            flush.fill_(1)
            e0.record()
            c2.render_device(handle, cam, st, frame.data_ptr(), None, band, stream)
            e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev[2:])
        res.append(sum(ms) / len(ms))
    print(name, "n_ranks", n, "kernel ms per rank:", " ".join(f"{m:.4f}" for m in res), "| ideal", f"{res and (res[0] if n == 1 else 0):.4f}")
