import torch, time
x = torch.empty(2_000_000_000, dtype=torch.float32, device="cuda")
for fn, name in ((lambda: x.zero_(), "memset 8 GB"), (lambda: x.fill_(1.5), "fill 8 GB")):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1), "ms", 8e9 / e0.elapsed_time(e1) / 1e6, "GB/s")
