"""ncu target: a few steps of the batched pipeline (K1 MFCC -> K2p window scorer -> K3 judge) on BASELINE configs[1]'s
shape with the audio resident in HBM. Usage: python tools/prof_pipeline.py [streams] [config 2|5] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
import rustpotter_b200 as rp  # noqa: E402
from bench import CONFIGS, load_wakewords, synth_streams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = CONFIGS[cfg]
rpws, utts = load_wakewords(c["n_wakewords"])
dev = torch.device("cuda", 0)
audio = synth_streams(torch, 0, n, dev, utts)
bt = rp.RustpotterBatch(n, rp.default_config(score_mode=c["score_mode"]))
for i, r in enumerate(rpws):
    bt.add_wakeword_from_buffer(f"w{i}", r)
bt.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(steps):
    bt.reset()
    nd = bt.process_count(audio)
torch.cuda.synchronize()
print("detections", nd, "windows", bt.windows_scored(), "stage", bt.last_timings(), "gate", bt.last_gate_stats())
