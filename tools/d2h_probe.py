import torch, time
dev=torch.device("cuda")
n=3*1024*1024
src=[torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(8)]
dst=[torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(8)]
for ns in (1,2,4):
    streams=[torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    for rep in range(2):
        t0=time.perf_counter()
        for i in range(400):
            with torch.cuda.stream(streams[i%ns]):
                dst[i%8].copy_(src[i%8], non_blocking=True)
        torch.cuda.synchronize()
        dt=time.perf_counter()-t0
    print(ns, "streams: %.1f GB/s, %.1f us per 3 MB copy" % (400*n/dt/1e9, dt/400*1e6))
