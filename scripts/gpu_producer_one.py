"""GPU-box helper for ncu: a few launches of the fused theta / A producer (B = 1024, L = 256, D = 1024)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import producer
d = torch.device("cuda:0")
B, L, D = (int(x) for x in (sys.argv[1:4] or (1024, 256, 1024)))
g = torch.Generator(device=d).manual_seed(0)
zs = [torch.randn(B, L, D, generator=g, device=d) * (1.1 / D ** 0.25) for _ in range(4)]
for _ in range(3):
    th, a = producer.theta_a(*zs)
torch.cuda.synchronize()
print("ok", float(th.sum()))
