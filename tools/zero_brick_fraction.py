"""How many 128-byte bricks (8 x 4 x 4 voxels) of the config-2 field are exactly zero?  (a 1-bit-per-brick
"all zero" mask would let the update skip the gather for them.)  python tools/zero_brick_fraction.py [C2|C1|C5s]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, synth  # noqa: E402


class _Dev:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


name = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg = synth.CONFIGS[name]
sm = synth.make_map(cfg["extent"], cfg["resol"])
g = Grid3d(0)
g.computePointCloud(sm.points, cfg["resol"])
g.computeGrid(sm.points, 0.05, sub=2)
sx, sy, sz = g.info()["size"]
prob = torch.as_tensor(_Dev(g.device_ptr(), sx * sy * sz), device="cuda:0").view(sz, sy, sx)
z4, y4, x8 = sz // 4 * 4, sy // 4 * 4, sx // 8 * 8
zero_voxels = 0
zero_bricks = 0
bricks = 0
for z0 in range(0, z4, 40):  # slabs: keeps the temporaries small
    z1 = min(z0 + 40, z4)
    p = prob[z0:z1, :y4, :x8]
    zero_voxels += int((p == 0).sum().item())
    b = p.reshape((z1 - z0) // 4, 4, y4 // 4, 4, x8 // 8, 8).amax(dim=(1, 3, 5))
    zero_bricks += int((b == 0).sum().item())
    bricks += b.numel()
print(json.dumps({"map": name, "size": [sx, sy, sz], "zero_voxel_fraction": zero_voxels / (z4 * y4 * x8),
                  "all_zero_brick_fraction": zero_bricks / bricks, "bricks": bricks}))
