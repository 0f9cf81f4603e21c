"""Does a plain SM streaming kernel (torch elementwise add) also slow down when the host synchronises before every launch?"""
import time, torch
a = torch.empty(1 << 28, device="cuda", dtype=torch.float32); b = torch.empty_like(a)
def run(name, fn, sync, n=20, gap=0.0):
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); evs.append((e0, e1))
        if sync: torch.cuda.synchronize(); time.sleep(gap)
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in evs)
    print(f"{name:40s} median {ms[n//2]:.3f} ms  min {ms[0]:.3f}  max {ms[-1]:.3f}  -> {2*(1<<30)/ms[n//2]/1e6:.0f} GB/s median")
add = lambda: torch.add(a, 1.0, out=b)
cp = lambda: b.copy_(a)
for _ in range(3): add(); cp()
run("memcpy d2d 1GiB, no sync", cp, False)
run("memcpy d2d 1GiB, sync each", cp, True)
run("elementwise add 1GiB, no sync", add, False)
run("elementwise add 1GiB, sync each", add, True)
run("elementwise add 1GiB, sync + 5ms idle", add, True, gap=0.005)
run("elementwise add 1GiB, no sync", add, False)
