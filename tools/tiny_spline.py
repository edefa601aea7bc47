import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from waveflow_b200.splines.tables import SplineTables
from waveflow_b200.splines.factories import spline_apply
dev = torch.device("cuda:0")
tabs = SplineTables.get("I", 6, 23)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 300 + 7
c = torch.rand(M, 29, device=dev); x = torch.rand(M, device=dev)
v, g = spline_apply(tabs, c, x, 0, 2)
torch.cuda.synchronize()
vd, gd = spline_apply(tabs, c, x, 0, 2, force_dense=True)
torch.cuda.synchronize()
print("ok", torch.equal(v, vd), torch.equal(g, gd))
