import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle_lib as ol
import test_gpu_push as t
from b200_helpers import gpu_push
for name in t.CASES:
    for vth in (0.05, 0.7):
        og, flds, prts, off = t._setup(name, vth)
        for path in ("tiled_warp",):
            opts, sort_first = t.PATHS[path]
            res = {}
            for fma in (0, 1):
                f_ref, p_ref = flds.copy(), prts.copy()
                if sort_first: ol.sort(og, p_ref, off)
                ol.push_mprts(og, f_ref, p_ref, off)
                f_gpu, p_gpu = flds.copy(), prts.copy()
                gpu_push(dict(opts, fma=fma), sort_first)(og, f_gpu, p_gpu, off)
                jr, jg = f_ref[:, :3], f_gpu[:, :3]
                res[fma] = np.abs(jg - jr).max() / np.abs(jr).max()
            print("%-22s vth %.2f  J rel err exact %.2e  fma %.2e" % (name, vth, res[0], res[1]))
