"""Tuning sweep of the warp-per-tile bit-walk aggregation kernel (20 < N <= 96): tile bytes x CTAs per SM, both launch regimes."""
import os, subprocess, sys
code = r'''
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import sweep
from bench import load_peaks
PEAK = load_peaks()["hbm_gbs"]
import torch
for N, B in ((32, 8192), (64, 8192), (32, 32768), (40, 8192)):
    by, dep, ind = sweep.agg_point(B, N)
    print(f"  N={N:3d} B={B:5d}: dep {dep:6.2f} us {by/dep/1e3/PEAK:.3f} | ind {ind:6.2f} us {by/ind/1e3/PEAK:.3f}", flush=True)
'''
for tb in (2048, 3072):
    for cap in (1, 2, 4, 8):
        env = dict(os.environ, V2V_AGG_TILE_BYTES=str(tb), V2V_AGG_CTAS=str(cap))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"tile_bytes={tb} ctas_cap={cap}\n" + (r.stdout if r.returncode == 0 else r.stderr[-400:]), flush=True)
