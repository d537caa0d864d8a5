import torch, time
dev="cuda:0"
n=16384*60
d=torch.randn(n, device=dev); h=torch.empty(n).pin_memory()
a_h=torch.randn(16384*8).pin_memory(); a_d=torch.empty(16384*8, device=dev)
def t(fn, it=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/it*1e6
def one():
    h.copy_(d, non_blocking=True); torch.cuda.current_stream().synchronize()
print("D2H 3.9MB single + sync: %.1f us" % t(one))
streams=[torch.cuda.Stream() for _ in range(4)]
def split(k):
    def f():
        ev=torch.cuda.Event(); ev.record()
        c=n//k
        for i in range(k):
            s=streams[i]; s.wait_event(ev)
            with torch.cuda.stream(s):
                h[i*c:(i+1)*c].copy_(d[i*c:(i+1)*c], non_blocking=True)
        for i in range(k): streams[i].synchronize()
    return f
for k in (2,4): print("D2H split %d: %.1f us" % (k, t(split(k))))
def h2d():
    a_d.copy_(a_h, non_blocking=True); torch.cuda.current_stream().synchronize()
print("H2D 0.5MB + sync: %.1f us" % t(h2d))
small=[torch.empty(16384, device=dev) for _ in range(3)]; small_h=[torch.empty(16384).pin_memory() for _ in range(3)]
def three():
    for s_,h_ in zip(small,small_h): h_.copy_(s_, non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("3 small D2H (64KB each) + sync: %.1f us" % t(three))
def nothing():
    torch.cuda.current_stream().synchronize()
print("sync only: %.1f us" % t(nothing))
big=torch.randn(64*1024*1024//4, device=dev); bh=torch.empty(64*1024*1024//4).pin_memory()
def bigc():
    bh.copy_(big, non_blocking=True); torch.cuda.current_stream().synchronize()
print("D2H 64MB: %.1f us -> %.1f GB/s" % (t(bigc,20), 64*1.048576/ (t(bigc,20)) *1e3))
