#!/usr/bin/env python
"""Generates tests/golden/bench_templates.npz: the template matrices of the benchmark's four synthetic WakewordRefs
(BASELINE configs[1] uses the first, configs[4] all four), i.e. the ORACLE's MFCC + CMN (reference
src/mfcc/extractor.rs:60-163, src/mfcc/normalizer.rs:3-31) of the deterministic numpy utterances of
tests/helpers.wakeword_utterances. Both arms of bench.py (CUDA path and `--impl reference`) load these vectors, so they
score identical templates; the GPU arm never touches the oracle. tests/test_oracle_golden.py pins the file to the
oracle. Run from the repo root: python tools/make_bench_templates.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.helpers import CONFIG5_LENGTHS, CONFIG5_SEEDS, wakeword_utterances  # noqa: E402


def main():
    out = {}
    for w, (lengths, seed) in enumerate(zip(CONFIG5_LENGTHS, CONFIG5_SEEDS)):
        for i, u in enumerate(wakeword_utterances(lengths, seed)):
            m = O.normalize(O.mfcc_stream(u, 16))
            assert m.shape == (lengths[i], 16), m.shape
            out[f"w{w}_t{i}"] = m.astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "bench_templates.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "matrices")


if __name__ == "__main__":
    main()
