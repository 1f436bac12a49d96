import torch, time
n = 402_665_184
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("H2D", n / e0.elapsed_time(e1) / 1e6, "GB/s", e0.elapsed_time(e1), "ms")
o = torch.empty(116_404_224, dtype=torch.uint8).pin_memory(); dd = torch.empty(116_404_224, dtype=torch.uint8, device="cuda")
o.copy_(dd); torch.cuda.synchronize()
e0.record(); o.copy_(dd, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H", 116_404_224 / e0.elapsed_time(e1) / 1e6, "GB/s", e0.elapsed_time(e1), "ms")
