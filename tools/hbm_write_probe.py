import torch
x = torch.empty(1536*1024*1024//4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(f, n=10):
    f(); torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(n):
        a.record(); f(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
ms=t(lambda: x.fill_(1.0)); print(f"fill 1.5 GiB: {ms:.3f} ms = {x.numel()*4/ms/1e9:.2f} TB/s")
ms=t(lambda: y.copy_(x)); print(f"copy 1.5 GiB: {ms:.3f} ms = {2*x.numel()*4/ms/1e9:.2f} TB/s (r+w)")
ms=t(lambda: x.sum()); print(f"sum 1.5 GiB: {ms:.3f} ms = {x.numel()*4/ms/1e9:.2f} TB/s read")
