import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import CorrBlock, raft_corr, _lib
B=4
f1 = torch.randn(B, 256, 48, 160, device="cuda"); f2 = torch.randn(B, 256, 48, 160, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/n*1e3
print("empty 943MB", t(lambda: torch.empty((B*7680,1,48,160), device="cuda")))
keep=[None]
def build():
    keep[0]=raft_corr.allpairs_pyramid(f1,f2,4,"tf32")
print("allpairs_pyramid keep-prev", t(build))
def build2():
    keep[0]=None
    keep[0]=raft_corr.allpairs_pyramid(f1,f2,4,"tf32")
print("allpairs_pyramid drop-prev", t(build2))
def build3():
    keep[0]=None
    keep[0]=CorrBlock(f1,f2,4,4)
print("CorrBlock drop-prev", t(build3))
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); build2(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
