import torch
dev='cuda'
flush=torch.empty(256*1024*1024//4, device=dev)
def timeit(fn, n=20):
    ts=[]
    for _ in range(n):
        flush.fill_(1.0); torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    ts.sort(); return ts[len(ts)//2]
for mb in (14.4, 28.8, 57.6, 115.2, 460.8, 1843.2):
    n=int(mb*1e6/2)
    x=torch.randn(n, device=dev).bfloat16(); y=torch.empty_like(x)
    t_sum=timeit(lambda: x.float().sum()) if False else None
    t_copy=timeit(lambda: y.copy_(x))
    t_read=timeit(lambda: torch.max(x))
    print(f"{mb:8.1f} MB: copy {t_copy:8.1f} us -> {2*mb/t_copy:6.2f} TB/s (r+w);  max-reduce {t_read:8.1f} us -> {mb/t_read:6.2f} TB/s read")
