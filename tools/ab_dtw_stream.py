"""A/B timing of the streaming DTW kernel variants on BASELINE configs[3] (1M pairs 120x16 vs 100x16, band 5).
CUDA events on the launching stream, warm-up first; prints one JSON line per variant. Not product code."""
import json
import sys

import torch

sys.path.insert(0, ".")
import rustpotter_b200 as rp  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
g = torch.Generator(device="cuda").manual_seed(1234)
scale = torch.tensor([8, 4, 3, 2, 2, 1.5] + [1.0] * 10, device="cuda")
a = torch.randn((P, 120, 16), device="cuda", generator=g) * scale
w = torch.randn((P, 100, 16), device="cuda", generator=g) * scale
out = torch.empty(P, device="cuda")
rp.set_dtw_variant(1)
sub = 20_000
ref = rp.dtw_scores(a[:sub].contiguous(), w[:sub].contiguous(), band=5)
for v in variants:
    rp.set_dtw_variant(v)
    for _ in range(2):
        rp.dtw_scores(a, w, band=5, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    for i in range(5):
        ev[i].record()
        rp.dtw_scores(a, w, band=5, out=out)
    ev[5].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    rel = ((out[:sub] - ref).abs() / ref.abs().clamp_min(1e-12)).max().item()
    best = min(ms)
    print(json.dumps({"variant": v, "pairs": P, "ms_best": round(best, 3), "ms_all": [round(x, 3) for x in ms],
                      "GBps": round(P * 14084 / best / 1e6, 1), "max_rel_vs_generic": rel}))
rp.set_dtw_variant(0)
