"""ncu target: the reference's 30 ms cadence on the batched handle (K1 -> K2c -> K3 per call).
Usage: python tools/prof_cadence.py [streams] [config 2|5] [calls]"""
import sys

import torch

sys.path.insert(0, ".")
import rustpotter_b200 as rp  # noqa: E402
from bench import CONFIGS, load_wakewords, synth_streams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 130
c = CONFIGS[cfg]
rpws, utts = load_wakewords(c["n_wakewords"])
dev = torch.device("cuda", 0)
audio = synth_streams(torch, 0, n, dev, utts)[:, : 480 * calls].contiguous()
bt = rp.RustpotterBatch(n, rp.default_config(score_mode=c["score_mode"]))
for i, r in enumerate(rpws):
    bt.add_wakeword_from_buffer(f"w{i}", r)
bt.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
nd = 0
dtw = []
for k in range(calls):
    nd += bt.process_count(audio[:, 480 * k: 480 * (k + 1)].contiguous())
    if k >= 40:
        dtw.append(bt.last_timings()["dtw_ms"])
torch.cuda.synchronize()
dtw.sort()
print("detections", nd, "windows", bt.windows_scored(), "dtw_ms median", round(dtw[len(dtw) // 2], 4) if dtw else None, "stage", bt.last_timings())
