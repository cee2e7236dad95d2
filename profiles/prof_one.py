"""Tuning aid: renders `frames` frames of one bench workload on cuda:0 through c2rt_render_device (device-resident output,
nothing else launched), prints the mean kernel time, and gives ncu a clean launch list:
  ncu --set full --import-source on --clock-control none -k regex:render_frame --launch-skip 2 --launch-count 1 \
      -o gpurun_out/prof python profiles/prof_one.py c4 3
usage: python profiles/prof_one.py <workload | path/to/scene.sdl@WxH> [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import chess2rt_b200 as c2  # noqa: E402
from bench import RNG_SEED, WORKLOADS  # noqa: E402


def main():
    name = sys.argv[1]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if "@" in name:   # an arbitrary scene file at a given size
        f, size = name.split("@")
        path, (W, H), over = f, tuple(int(v) for v in size.split("x")), {}
    else:
        path, W, H, over = WORKLOADS[name]
    c2.init(1, [0])
    scene = c2.HostScene(os.path.join(ROOT, path))
    scene.set_frame_size(W, H)
    scene.override(**over)
    handle = scene.device_scene()
    cam, st = scene.frame_blocks(seed=RNG_SEED)
    frame = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
    for e0, e1 in ev:
        e0.record()
        c2.render_device(handle, cam, st, frame.data_ptr(), None, None, stream)
        e1.record()
    torch.cuda.synchronize()
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    print(name, "kernel ms per frame:", " ".join(f"{m:.4f}" for m in ms), "checksum", float(frame.double().sum()))
    scene.close()
    c2.shutdown()


if __name__ == "__main__":
    main()
