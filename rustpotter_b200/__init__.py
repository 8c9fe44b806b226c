"""rustpotter_b200 — B200-native MFCC + WakewordRef/DTW scoring path behind rustpotter's API.

The product is the C-ABI library `librustpotter_b200.so` (include/rustpotter_b200.h); this package
is its ctypes binding plus the in-tree build script. There is no CPU implementation here.
"""
from .api import (Config, Rustpotter, RustpotterBatch, RustpotterError, default_config, device_count, dtw_scores,  # noqa: F401
                  host_replay, lib, mfcc_frames, resample_to_16k, set_avg_gate, set_dtw_variant, set_mfcc_variant, wakeword_inspect, build_wakeword, wakeword_from_features, wakeword_template)
