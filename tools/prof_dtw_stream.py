"""ncu target: a few launches of the streaming DTW kernel on BASELINE configs[3]'s shape."""
import sys
import torch
sys.path.insert(0, ".")
import rustpotter_b200 as rp
P = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
g = torch.Generator(device="cuda").manual_seed(1234)
scale = torch.tensor([8, 4, 3, 2, 2, 1.5] + [1.0] * 10, device="cuda")
a = torch.randn((P, 120, 16), device="cuda", generator=g) * scale
w = torch.randn((P, 100, 16), device="cuda", generator=g) * scale
out = torch.empty(P, device="cuda")
for _ in range(3):
    rp.dtw_scores(a, w, band=5, out=out)
torch.cuda.synchronize()
print(float(out.mean()))
