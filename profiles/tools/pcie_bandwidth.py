"""Raw pinned-memory H2D / D2H bandwidth of the box (contiguous torch copies): the ceiling bench.py's e2e leg runs against."""
import torch, time
dev=torch.device('cuda',0)
for mb in (32, 256, 1024):
    n=mb*1024*1024
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
    for _ in range(2): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    reps=max(4, 4096//mb)
    e0.record()
    for _ in range(reps): d.copy_(h,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('H2D',mb,'MB', round(n*reps/e0.elapsed_time(e1)/1e6,1),'GB/s')
    e0.record()
    for _ in range(reps): h.copy_(d,non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print('D2H',mb,'MB', round(n*reps/e0.elapsed_time(e1)/1e6,1),'GB/s')
