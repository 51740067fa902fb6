import torch, time
d = torch.device("cuda")
for mb in (256, 2048):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, pin_memory=True)
    g = torch.empty_like(h, device=d)
    for name, src, dst in (("h2d", h, g), ("d2h", g, h)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        print(name, mb, "MB:", 3 * mb / 1024 / (time.perf_counter() - t), "GB/s")
