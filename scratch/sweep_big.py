import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scripts')
import sweep
for N in (32, 48, 64, 128, 256):
    for B in (1024, 8192):
        nb, d, i = sweep.agg_point(B, N)
        print(N, B, f"{nb/1e6:8.1f} MB dep {d:8.2f} us {nb/d/1e3/sweep.PEAK:.3f}  ind {i:8.2f} us {nb/i/1e3/sweep.PEAK:.3f}", flush=True)
