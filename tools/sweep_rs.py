import os, sys, itertools, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tools"))
import numpy as np, torch
import b200arrays as b
from b200arrays import dtypes as dt
from microbench import rand_array, timeit
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for e, name in ((dt.F32, "F32"), (dt.BF16, "BF16")):
    A = rand_array(e, (16384, 16384)); R = b.CuArray.undef(e, (16384, 1))
    t = timeit(lambda: b.mapreducedim_(b.identity, b.add_sum, R, A, init=0), flush=flush, reps=10)
    print(name, "%%.4f" %% t[1], "ms med", "%%.0f GB/s" %% (A.nbytes / t[1] / 1e6), end="   ")
print()
''' % (ROOT, ROOT)
for bx, ty, s in [(64, 4, 19), (128, 2, 37), (32, 8, 10), (64, 4, 37), (128, 2, 74), (256, 1, 74), (256, 1, 148), (128, 2, 19), (64, 4, 10)]:
    env = dict(os.environ, B200_RS_BX=str(bx), B200_RS_TY=str(ty), B200_RS_S=str(s))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"BX={bx} TY={ty} S={s}:", out.stdout.strip() or out.stderr[-300:], flush=True)
