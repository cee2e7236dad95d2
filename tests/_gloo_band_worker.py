"""Worker of test_two_rank_band_gather_over_gloo (launched by torch.distributed.run, backend gloo)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chess2rt_b200 import bands  # noqa: E402
from oracle_binding import OracleScene  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, n = dist.get_rank(), dist.get_world_size()
    W, H, B = 64, 52, 8   # H is not a multiple of the band height: the last band is ragged
    s = OracleScene(os.path.join(ROOT, "scenes", "lecture5.sdl"))
    s.set_frame_size(W, H)
    rows = bands.owned_rows(H, rank, n, B)
    pad = bands.rows_padded(H, n, B)
    mine = np.zeros((pad, W, 3), np.float32)
    # render only this rank's bands
    k = 0
    for y0 in range(rank * B, H, n * B):
        y1 = min(H, y0 + B)
        band, _ = s.render_rows(y0, y1, threads=1)
        mine[k:k + (y1 - y0)] = band
        k += y1 - y0
    assert k == rows.size
    t = torch.from_numpy(mine)
    gathered = [torch.empty_like(t) for _ in range(n)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank == 0:
        frame = bands.scatter_rows(torch.stack(gathered).numpy(), H, n, B)
        full, _ = s.render(threads=1)
        assert np.array_equal(frame, full)
        open(sys.argv[1], "w").write("ok\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
