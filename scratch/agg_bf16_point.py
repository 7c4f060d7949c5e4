"""bf16-storage aggregation point (B x N=20 x F=16), both launch regimes, walk vs dense form (V2V_AGG_WALK)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, ROOT)
import torch
import sweep
for B in (8192, 32768):
    nbytes, us_dep, us_ind = sweep.agg_point(B, 20, torch.bfloat16)
    print(f"V2V_AGG_WALK={os.environ.get('V2V_AGG_WALK', 'default')} bf16 B={B}: dep {us_dep:.2f} us ({nbytes / us_dep / 1e3 / sweep.PEAK:.3f}), ind {us_ind:.2f} us ({nbytes / us_ind / 1e3 / sweep.PEAK:.3f})")
nbytes, us_dep, us_ind = sweep.agg_point(8192, 20, torch.float32)
print(f"fp32 B=8192: dep {us_dep:.2f} us ({nbytes / us_dep / 1e3 / sweep.PEAK:.3f}), ind {us_ind:.2f} us ({nbytes / us_ind / 1e3 / sweep.PEAK:.3f})")
