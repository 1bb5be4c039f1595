"""GPU-box helper: the fused theta / A producer against torch (fp32 einsum + softplus / logsigmoid; and with TF32)."""
import sys, os, json
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepblast_b200 import producer
from gpu_sq_perf import timeit
d = torch.device("cuda:0")
for B, L, D in ((1024, 256, 1024), (32, 512, 1024), (256, 512, 1024)):
    g = torch.Generator(device=d).manual_seed(0)
    zs = [torch.randn(B, L, D, generator=g, device=d) * (2.0 / D ** 0.25) for _ in range(4)]
    t_ours = timeit(lambda: producer.theta_a(*zs), it=5, warm=2)
    def ref():
        th = F.softplus(torch.einsum('bid,bjd->bij', zs[0], zs[1]))
        a = F.logsigmoid(torch.einsum('bid,bjd->bij', zs[2], zs[3]))
        return th, a
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = timeit(ref, it=3, warm=1)
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(ref, it=3, warm=1)
    th_t, _ = ref()
    torch.backends.cuda.matmul.allow_tf32 = False
    th_o, _ = producer.theta_a(*zs)
    th_r, _ = ref()
    flop = 2 * 2.0 * B * L * L * D
    print(json.dumps({"B": B, "L": L, "D": D, "ours_ms": t_ours, "torch_fp32_ms": t_fp32, "torch_tf32_ms": t_tf32,
                      "ours_TFLOPs_bf16_issued": 3 * flop / t_ours / 1e9, "ours_TFLOPs_useful": flop / t_ours / 1e9,
                      "max_err_ours_vs_fp32": float((th_o - th_r).abs().max()),
                      "max_err_tf32_vs_fp32": float((th_t - th_r).abs().max())}), flush=True)
