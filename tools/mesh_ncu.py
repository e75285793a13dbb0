#!/usr/bin/env python
"""One extraction + one OBJ write on a synthetic 512^3 lattice (sphere with bumps), for an ncu launch list of the mesh kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnb_loader
from common import MID, product_config
import torch

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
pkg = rnb_loader.load_package()
t = pkg.Testbed(product_config(pkg, MID)); t.init_params()
ax = torch.arange(res, device="cuda", dtype=torch.float32) / res
z, y, x = torch.meshgrid(ax, ax, ax, indexing="ij")
d = (torch.sqrt((x - 0.5) ** 2 + (y - 0.47) ** 2 + (z - 0.52) ** 2) - 0.4 + 0.02 * torch.sin(19 * x) * torch.cos(23 * y) * torch.sin(17 * z)).contiguous()
del x, y, z
for _ in range(2):
    info = t.marching_cubes_from_density(d.data_ptr(), (res,) * 3, with_colors=True, use_ema=False)
nb = t.save_mesh("/tmp/ncu_mesh.obj", 0.5, (0.5, 0.5, 0.5), 1.0, (0, 0, 0), True)
print(info, nb)
os.remove("/tmp/ncu_mesh.obj")
