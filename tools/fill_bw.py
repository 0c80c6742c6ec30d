import torch, time
x = torch.empty(int(13.5e9)//8*2, dtype=torch.float64, device="cuda")  # 27 GB
for name, fn in (("zero_", lambda: x.zero_()), ("fill_", lambda: x.fill_(1.5)), ("memset", lambda: torch.cuda.memset(x.data_ptr(), 0, x.numel()*8) if hasattr(torch.cuda, "memset") else x.zero_())):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, ms, "ms", x.numel() * 8 / ms / 1e6, "GB/s")
