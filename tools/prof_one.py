"""Run ONE library call a few times (for ncu): python tools/prof_one.py <case> [log2n]"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200arrays as b
from b200arrays import dtypes as dt
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from microbench import rand_array

case = sys.argv[1]
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 28
n = 1 << lg
torch.cuda.set_device(0)
if case == "cumsum_f32":
    V = rand_array(dt.F32, (n,)); O = b.CuArray.undef(dt.F32, (n,))
    fn = lambda: b.accumulate_(b.add, O, V)
elif case == "cumsum_i64":
    V = rand_array(dt.I64, (n,)); O = b.CuArray.undef(dt.I64, (n,))
    fn = lambda: b.accumulate_(b.add, O, V)
elif case == "sum_f32":
    V = rand_array(dt.F32, (n,)); O = b.CuArray.undef(dt.F32, (1,))
    fn = lambda: b.mapreducedim_(b.identity, b.add_sum, O, V, init=0)
elif case == "sort_u32":
    V = rand_array(dt.U32, (n,)); K = V.copy()
    def fn():
        K.buf.copy_(V.buf); b.sort_(K)
elif case == "sortperm_f32":
    V = rand_array(dt.F32, (n,))
    fn = lambda: b.sortperm(V)
for _ in range(3):
    fn()
torch.cuda.synchronize()
print("done", case, n)
