"""Observed error of the MFCC kernel (K1) against the oracle on synthetic and fixture audio: max |delta coefficient|, the
coefficient scale, and the error relative to the frame's largest coefficient. Run on the GPU box; the numbers back the
tolerances in tests/test_gpu_parity.py and DESIGN.md. Usage: python tools/measure_k1_error.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rustpotter_b200 as rp  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.helpers import golden, read_wav_i16, synth_audio  # noqa: E402

out = {}
for variant in (0, 1):
    rp.set_mfcc_variant(variant)
    for d in (5, 13, 16, 31):
        audio = synth_audio(8, 160 * 400, seed=100 + d)
        audio[3] *= np.float32(0.01)                  # a quiet stream
        audio[5, 160 * 100:160 * 140] *= np.float32(1e-3)
        got = rp.mfcc_frames(torch.from_numpy(audio).cuda(), d).cpu().numpy()
        worst_abs, worst_rel, scale = 0.0, 0.0, 0.0
        for b in range(audio.shape[0]):
            want = O.mfcc_stream(audio[b], d)
            err = np.abs(got[b] - want)
            worst_abs = max(worst_abs, float(err.max()))
            worst_rel = max(worst_rel, float((err.max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-6)).max()))
            scale = max(scale, float(np.abs(want).max()))
        out[f"variant{variant}_d{d}"] = {"max_abs_err": worst_abs, "max_err_rel_to_frame_max": worst_rel, "max_abs_coefficient": scale}
    s = read_wav_i16(golden("oye_casa_g_1.wav")).astype(np.float32) / np.float32(32767.0)
    s = s[: len(s) // 480 * 480]
    got = rp.mfcc_frames(torch.from_numpy(s[None]).cuda(), 5).cpu().numpy()[0]
    want = O.mfcc_stream(s, 5)
    out[f"variant{variant}_fixture_oye_casa_g_1_d5"] = {"max_abs_err": float(np.abs(got - want).max()), "max_abs_coefficient": float(np.abs(want).max())}
rp.set_mfcc_variant(0)
print(json.dumps(out, indent=1))
