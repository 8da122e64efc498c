#!/usr/bin/env python3
"""Where the time of analyses.spot_diagram goes (run on the GPU box):
fixed per-call overhead (tiny grid), pinned D2H bandwidth, and the full call at BASELINE size
for several `pieces`."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_model
from rayoptics_b200 import table as T, engine as E, analyses as A

def timeit(f, n=30, warm=5):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0)/n*1e3

opm = load_model('dblgauss')
tab = T.SurfaceTable.from_model(opm.seq_model, device=0)
n = 3*3*512*512
pinned = {'abr': torch.empty((2, n), dtype=torch.float64).pin_memory()}
print('tiny grid (8x8), per call ms      :', timeit(lambda: A.spot_diagram(opm, 8, table=tab)))
dev = torch.empty((2, n), dtype=torch.float64, device='cuda')
def d2h():
    pinned['abr'][0].copy_(dev[0], non_blocking=True); pinned['abr'][1].copy_(dev[1], non_blocking=True)
    torch.cuda.current_stream().synchronize()
ms = timeit(d2h)
print('D2H 37.7 MB pinned ms / GB/s      :', ms, n*16/ms/1e6)
grid = E.grid_for_model(opm, tab, 512)
res = E.BundleResult(grid.n_rays, tab.n_ifc, torch.device('cuda', 0), ('abr',), nan_status=True)
print('trace only (abr, nan) ms          :', timeit(lambda: E.trace_grid(tab, grid, res=res)))
for pieces in (1, 2, 4, 8, 16):
    print(f'spot_diagram 512, pieces={pieces:2d} ms  :', timeit(lambda: A.spot_diagram(opm, 512, table=tab, pinned=pinned, pieces=pieces)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    A.spot_diagram(opm, 512, table=tab, pinned=pinned)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
