"""One VoxelGrid workload a few times through one pipeline (for ncu):  python tools/vg_one.py [path] [tiles_x tiles_y]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pcgol_b200 as pg  # noqa: E402
from pcgol_b200 import _lib, synth  # noqa: E402

path = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xyz = synth.tiled_map(int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else synth.lidar_scan(2, n_az=15625)
n = len(xyz)
d_in = torch.from_numpy(xyz.view(np.uint8).reshape(-1).copy()).cuda()
d_out = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
vg = pg.VoxelGrid((0.05, 0.05, 0.05), (128, 128, 128))
_lib.set_vg_path(path)
for _ in range(3):
    m = vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr())
torch.cuda.synchronize()
print(n, m)
