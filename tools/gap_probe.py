import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from dmip.models.diffusion import CDE
torch.manual_seed(0)
m = CDE(100, 27, [512,512,512]); m.sde.eval()
y = torch.randn(27).cuda()
N, S = 1<<20, 200
kw = dict(num_samples=N, num_steps=S, seed=1234, return_tensor=True)
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
for _ in range(2): m(y, **kw)
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
for use_flush in (False, True):
    evs = []
    t0 = time.perf_counter()
    for i in range(3):
        a, b, c = E(), E(), E()
        a.record()
        if use_flush: flush.zero_()
        b.record()
        out = m(y, **kw)
        c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("flush", use_flush, "wall %.1f ms" % (wall*1e3), [(round(a.elapsed_time(b),2), round(b.elapsed_time(c),2)) for a,b,c in evs],
          "gaps", [round(evs[i][2].elapsed_time(evs[i+1][0]),2) for i in range(2)])
