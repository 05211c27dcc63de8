import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import raft_corr
B=4; C=int(os.environ.get("CH","256"))
f1 = torch.randn(B, C, 48, 160, device="cuda"); f2 = torch.randn(B, C, 48, 160, device="cuda")
keep=[None]
def build():
    keep[0]=None
    keep[0]=raft_corr.allpairs_pyramid(f1,f2,int(os.environ.get("LEVELS","4")),os.environ.get("PREC","tf32"),blocked=os.environ.get("BLOCKED","1")=="1")
for _ in range(3): build()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): build()
e1.record(); torch.cuda.synchronize()
print("blocked", os.environ.get("BLOCKED","1"), "C", C, "debug", os.environ.get("B200CORR_DEBUG","0"), "levels", os.environ.get("LEVELS","4"), "build ms", round(e0.elapsed_time(e1)/10,4))
