import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from understanding_flow_robustness_b200 import attack
from understanding_flow_robustness_b200.harness import FlowNetCHarness
dev = torch.device("cuda")
def run(tag, cl, bench, nb=16):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    net = FlowNetCHarness().to(dev).eval()
    if cl: net = net.to(memory_format=torch.channels_last)
    for q in net.parameters(): q.requires_grad_(False)
    i1 = torch.rand(nb, 3, 384, 1280, device=dev); i2 = torch.rand(nb, 3, 384, 1280, device=dev)
    if cl: i1 = i1.contiguous(memory_format=torch.channels_last); i2 = i2.contiguous(memory_format=torch.channels_last)
    patch = torch.rand(1, 3, 100, 100, device=dev); mask = attack.circle_mask(100, dev); cfg = attack.PatchAttackConfig()
    def it(pt): return attack.patch_attack_iteration(net, i1, i2, pt, mask, patch.clone(), cfg, nb)[0]
    p = it(patch); p = it(p); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): p = it(p)
    torch.cuda.synchronize()
    print(tag, "ms/iter", round((time.perf_counter()-t0)/3*1e3,1), "pairs/s", round(nb*3/(time.perf_counter()-t0),1), flush=True)
run("baseline", False, False)
run("cudnn.benchmark", False, True)
run("channels_last+benchmark", True, True)
