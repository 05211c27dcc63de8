import torch
x = torch.empty(1253376000 // 4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
tf = t(lambda: x.fill_(1.0))
print("fill 1.25GB ms", round(tf,4), "GB/s", round(1.2534/tf*1e3))
tz = t(lambda: torch.cuda.memset if False else x.zero_())
print("zero 1.25GB ms", round(tz,4), "GB/s", round(1.2534/tz*1e3))
tc = t(lambda: y.copy_(x))
print("copy 1.25GB ms", round(tc,4), "GB/s (r+w)", round(2*1.2534/tc*1e3))
s = t(lambda: x.sum())
print("read 1.25GB ms", round(s,4), "GB/s", round(1.2534/s*1e3))
