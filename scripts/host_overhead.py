import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import backend
B=8
a = torch.randn(B, 256, 48, 160, device="cuda"); b = torch.randn(B, 256, 48, 160, device="cuda")
g = torch.randn(B, 21, 21, 48, 160, device="cuda")
q = (1, 1, 21, 21, 0, 0, 1, 1, 2, 2, 1, 1)
def ev(): return torch.cuda.Event(enable_timing=True)
for name, fn in [("fwd only", lambda: backend.forward(a,b,*q)), ("bwd only", lambda: backend.backward(a,b,g,*q)),
                 ("fwd+bwd", lambda: (backend.forward(a,b,*q), backend.backward(a,b,g,*q)))]:
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0,e1=ev(),ev(); t0=time.perf_counter(); e0.record()
    for _ in range(100): fn()
    e1.record(); th=time.perf_counter()-t0; torch.cuda.synchronize()
    print(name, "gpu ms/iter", round(e0.elapsed_time(e1)/100,4), "host enqueue ms/iter", round(th*10,4))
