"""ctypes binding of librustpotter_b200.so (include/rustpotter_b200.h).

`Rustpotter` mirrors the reference's `Rustpotter` struct (src/detector.rs:34-302) method for method;
`RustpotterBatch` is the batched front-end; `mfcc_frames` / `dtw_scores` are the raw kernels.
The library is loaded from the in-tree build; importing fails loudly if it has not been built, and
creating a detector fails loudly without a CUDA device — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RP_LIB_PATH: another build of the same library (same-box A/B measurements of kernel variants; debug only)
LIB_PATH = os.environ.get("RP_LIB_PATH") or os.path.join(_HERE, "librustpotter_b200.so")
NAME_MAX = 128

SCORE_MODES = {"average": 0, "max": 1, "median": 2, "p25": 3, "p50": 4, "p75": 5, "p80": 6, "p90": 7, "p95": 8}
SAMPLE_FORMATS = {"i8": 0, "i16": 1, "i32": 2, "f32": 3}
VAD_MODES = {None: -1, "easy": 0, "medium": 1, "hard": 2}
ENDIANNESS = {"little": 0, "big": 1, "native": 2}


class RustpotterError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class Config(C.Structure):
    """rp_config — RustpotterConfig flattened (src/config.rs)."""

    _fields_ = [
        ("sample_rate", C.c_uint32), ("sample_format", C.c_uint32), ("channels", C.c_uint32), ("endianness", C.c_uint32),
        ("avg_threshold", C.c_float), ("threshold", C.c_float), ("min_scores", C.c_uint64), ("eager", C.c_uint32),
        ("score_ref", C.c_float), ("band_size", C.c_uint32), ("score_mode", C.c_uint32), ("vad_mode", C.c_int32),
        ("gain_normalizer_enabled", C.c_uint32), ("gain_ref_set", C.c_uint32), ("gain_ref", C.c_float),
        ("min_gain", C.c_float), ("max_gain", C.c_float), ("band_pass_enabled", C.c_uint32), ("low_cutoff", C.c_float),
        ("high_cutoff", C.c_float),
    ]


class CDetection(C.Structure):
    _fields_ = [
        ("name", C.c_char * NAME_MAX), ("avg_score", C.c_float), ("score", C.c_float), ("counter", C.c_uint64),
        ("gain", C.c_float), ("n_scores", C.c_uint32), ("score_names", C.POINTER(C.c_char_p)),
        ("score_values", C.POINTER(C.c_float)),
    ]

    def to_dict(self) -> dict:
        return {
            "name": self.name.decode(),
            "avg_score": np.float32(self.avg_score),
            "score": np.float32(self.score),
            "counter": int(self.counter),
            "gain": np.float32(self.gain),
            "scores": {self.score_names[i].decode(): np.float32(self.score_values[i]) for i in range(self.n_scores)},
        }


class CBatchDetection(C.Structure):
    _fields_ = [("stream", C.c_int64), ("chunk", C.c_int64), ("det", CDetection)]


class WakewordInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * NAME_MAX), ("mfcc_size", C.c_int32), ("n_templates", C.c_int32), ("avg_frames", C.c_int32),
        ("max_frames", C.c_int32), ("has_threshold", C.c_int32), ("has_avg_threshold", C.c_int32), ("threshold", C.c_float),
        ("avg_threshold", C.c_float), ("rms_level", C.c_float), ("is_v2", C.c_int32),
    ]


_lib = None

# every symbol include/rustpotter_b200.h declares
EXPORTED = [
    "rp_version", "rp_device_count", "rp_config_default", "rp_last_error", "rp_create", "rp_destroy",
    "rp_add_wakeword_from_buffer", "rp_add_wakeword_from_file", "rp_remove_wakeword", "rp_remove_wakewords",
    "rp_get_samples_per_frame", "rp_get_bytes_per_frame", "rp_get_partial_detection", "rp_get_rms_level", "rp_get_gain",
    "rp_get_rms_level_ref", "rp_process_bytes", "rp_process_samples_i8", "rp_process_samples_i16",
    "rp_process_samples_i32", "rp_process_samples_f32", "rp_update_config", "rp_update_detector_config",
    "rp_update_filters_config", "rp_reset", "rp_windows_scored", "rp_batch_create", "rp_batch_destroy",
    "rp_batch_add_wakeword_from_buffer", "rp_batch_add_wakeword_from_file", "rp_batch_remove_wakewords",
    "rp_batch_set_cuda_stream", "rp_batch_process", "rp_batch_update_config", "rp_batch_reset",
    "rp_batch_windows_scored", "rp_batch_n_streams", "rp_batch_max_mfcc_frames", "rp_batch_last_timings",
    "rp_batch_last_launches", "rp_batch_copy_last_scores", "rp_batch_create_multi", "rp_batch_n_devices",
    "rp_batch_remove_wakeword", "rp_batch_process_samples", "rp_batch_process_bytes", "rp_batch_last_gate_stats", "rp_set_avg_gate", "rp_batch_samples_per_frame", "rp_resample_to_16k", "rp_mfcc_frames", "rp_dtw_scores", "rp_set_dtw_variant", "rp_set_mfcc_variant", "rp_wakeword_inspect",
    "rp_wakeword_template", "rp_host_replay", "rp_wakeword_build", "rp_wakeword_from_features", "rp_debug_stream4_schedule", "rp_debug_stream4_ctl",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m rustpotter_b200.build` (or __graft_entry__.build()) first; "
                          "there is no fallback implementation")
    L = C.CDLL(LIB_PATH)
    vp, f32p, u8p = C.c_void_p, C.POINTER(C.c_float), C.c_char_p
    cfgp, detp = C.POINTER(Config), C.POINTER(CDetection)
    L.rp_version.restype = C.c_char_p
    L.rp_config_default.argtypes = [cfgp]
    L.rp_last_error.restype = C.c_char_p
    L.rp_last_error.argtypes = [vp]
    L.rp_create.argtypes = [cfgp, C.c_int, C.POINTER(vp)]
    L.rp_destroy.argtypes = [vp]
    L.rp_add_wakeword_from_buffer.argtypes = [vp, C.c_char_p, u8p, C.c_size_t]
    L.rp_add_wakeword_from_file.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.rp_remove_wakeword.argtypes = [vp, C.c_char_p]
    L.rp_remove_wakewords.argtypes = [vp]
    for n in ("rp_get_samples_per_frame", "rp_get_bytes_per_frame"):
        getattr(L, n).restype = C.c_size_t
        getattr(L, n).argtypes = [vp]
    L.rp_get_partial_detection.argtypes = [vp, detp]
    for n in ("rp_get_rms_level", "rp_get_gain", "rp_get_rms_level_ref"):
        getattr(L, n).restype = C.c_float
        getattr(L, n).argtypes = [vp]
    L.rp_process_bytes.argtypes = [vp, u8p, C.c_size_t, detp]
    for n, ct in (("i8", C.c_int8), ("i16", C.c_int16), ("i32", C.c_int32), ("f32", C.c_float)):
        getattr(L, f"rp_process_samples_{n}").argtypes = [vp, C.POINTER(ct), C.c_size_t, detp]
    for n in ("rp_update_config", "rp_update_detector_config", "rp_update_filters_config"):
        getattr(L, n).argtypes = [vp, cfgp]
    L.rp_reset.argtypes = [vp]
    L.rp_windows_scored.restype = C.c_uint64
    L.rp_windows_scored.argtypes = [vp]
    L.rp_batch_create.argtypes = [cfgp, C.c_int64, C.c_int, C.POINTER(vp)]
    L.rp_batch_create_multi.argtypes = [cfgp, C.c_int64, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.rp_batch_n_devices.argtypes = [vp]
    L.rp_batch_destroy.argtypes = [vp]
    L.rp_batch_remove_wakeword.argtypes = [vp, C.c_char_p]
    L.rp_batch_process_samples.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int, C.POINTER(C.POINTER(CBatchDetection)), C.POINTER(C.c_int64)]
    L.rp_batch_process_bytes.argtypes = [vp, vp, C.c_int64, C.c_int, C.POINTER(C.POINTER(CBatchDetection)), C.POINTER(C.c_int64)]
    L.rp_batch_last_gate_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.rp_set_avg_gate.argtypes = [C.c_int]
    L.rp_batch_samples_per_frame.restype = C.c_size_t
    L.rp_batch_samples_per_frame.argtypes = [vp]
    L.rp_resample_to_16k.restype = C.c_int64
    L.rp_resample_to_16k.argtypes = [C.c_uint32, f32p, C.c_size_t, f32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rp_batch_add_wakeword_from_buffer.argtypes = [vp, C.c_char_p, u8p, C.c_size_t]
    L.rp_batch_add_wakeword_from_file.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.rp_batch_remove_wakewords.argtypes = [vp]
    L.rp_batch_set_cuda_stream.argtypes = [vp, vp]
    L.rp_batch_process.argtypes = [vp, vp, C.c_int64, C.c_int, C.POINTER(C.POINTER(CBatchDetection)), C.POINTER(C.c_int64)]
    L.rp_batch_update_config.argtypes = [vp, cfgp]
    L.rp_batch_reset.argtypes = [vp]
    L.rp_batch_windows_scored.restype = C.c_uint64
    L.rp_batch_windows_scored.argtypes = [vp]
    L.rp_batch_n_streams.restype = C.c_int64
    L.rp_batch_n_streams.argtypes = [vp]
    L.rp_batch_max_mfcc_frames.argtypes = [vp]
    L.rp_batch_last_timings.argtypes = [vp, f32p, C.c_int]
    L.rp_batch_last_launches.argtypes = [vp]
    L.rp_batch_copy_last_scores.restype = C.c_int64
    L.rp_batch_copy_last_scores.argtypes = [vp, f32p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rp_mfcc_frames.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, vp, vp]
    L.rp_dtw_scores.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp]
    L.rp_set_dtw_variant.argtypes = [C.c_int]
    L.rp_set_mfcc_variant.argtypes = [C.c_int]
    L.rp_wakeword_inspect.argtypes = [u8p, C.c_size_t, C.POINTER(WakewordInfo)]
    L.rp_wakeword_template.argtypes = [u8p, C.c_size_t, C.c_int, C.c_char_p, f32p, C.c_size_t]
    L.rp_host_replay.argtypes = [cfgp, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, f32p, C.c_int64, C.c_int, f32p,
                                 C.POINTER(CBatchDetection), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _check(code: int, handle=None) -> int:
    if code < 0:
        raise RustpotterError(code, lib().rp_last_error(handle).decode(errors="replace"))
    return code


def default_config(**kw) -> Config:
    """RustpotterConfig::default() with keyword overrides (strings accepted for the enums)."""
    c = Config()
    lib().rp_config_default(C.byref(c))
    for k, v in kw.items():
        if k == "score_mode" and isinstance(v, str):
            v = SCORE_MODES[v.lower()]
        elif k == "sample_format" and isinstance(v, str):
            v = SAMPLE_FORMATS[v.lower()]
        elif k == "endianness" and isinstance(v, str):
            v = ENDIANNESS[v.lower()]
        elif k == "vad_mode" and (v is None or isinstance(v, str)):
            v = VAD_MODES[v]
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def device_count() -> int:
    return lib().rp_device_count()


class Rustpotter:
    """Drop-in twin of the reference's `Rustpotter` (src/detector.rs)."""

    def __init__(self, config: Config | None = None, device: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        cfg = config if config is not None else default_config()
        _check(self._L.rp_create(C.byref(cfg), device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rp_destroy(self._h)
            self._h = None

    __del__ = close

    def add_wakeword_from_buffer(self, key: str, buf: bytes):
        _check(self._L.rp_add_wakeword_from_buffer(self._h, key.encode(), buf, len(buf)), self._h)

    def add_wakeword_from_file(self, key: str, path: str):
        _check(self._L.rp_add_wakeword_from_file(self._h, key.encode(), path.encode()), self._h)

    def remove_wakeword(self, key: str) -> bool:
        return bool(_check(self._L.rp_remove_wakeword(self._h, key.encode()), self._h))

    def remove_wakewords(self) -> bool:
        return bool(_check(self._L.rp_remove_wakewords(self._h), self._h))

    def get_samples_per_frame(self) -> int:
        return self._L.rp_get_samples_per_frame(self._h)

    def get_bytes_per_frame(self) -> int:
        return self._L.rp_get_bytes_per_frame(self._h)

    def get_partial_detection(self):
        d = CDetection()
        return d.to_dict() if _check(self._L.rp_get_partial_detection(self._h, C.byref(d)), self._h) else None

    def get_rms_level(self):
        return np.float32(self._L.rp_get_rms_level(self._h))

    def get_gain(self):
        return np.float32(self._L.rp_get_gain(self._h))

    def get_rms_level_ref(self):
        return np.float32(self._L.rp_get_rms_level_ref(self._h))

    def process_bytes(self, audio_bytes: bytes):
        d = CDetection()
        r = _check(self._L.rp_process_bytes(self._h, audio_bytes, len(audio_bytes), C.byref(d)), self._h)
        return d.to_dict() if r else None

    def process_samples(self, samples):
        a = np.ascontiguousarray(samples)
        table = {np.dtype(np.int8): ("i8", C.c_int8), np.dtype(np.int16): ("i16", C.c_int16),
                 np.dtype(np.int32): ("i32", C.c_int32), np.dtype(np.float32): ("f32", C.c_float)}
        if a.dtype not in table:
            raise TypeError("samples must be int8/int16/int32/float32 (the reference's Sample impls)")
        nm, ct = table[a.dtype]
        d = CDetection()
        r = _check(getattr(self._L, f"rp_process_samples_{nm}")(self._h, a.ctypes.data_as(C.POINTER(ct)), a.size, C.byref(d)), self._h)
        return d.to_dict() if r else None

    def update_config(self, config: Config):
        _check(self._L.rp_update_config(self._h, C.byref(config)), self._h)

    def update_detector_config(self, config: Config):
        _check(self._L.rp_update_detector_config(self._h, C.byref(config)), self._h)

    def update_filters_config(self, config: Config):
        _check(self._L.rp_update_filters_config(self._h, C.byref(config)), self._h)

    def reset(self):
        self._L.rp_reset(self._h)

    def windows_scored(self) -> int:
        return int(self._L.rp_windows_scored(self._h))


class RustpotterBatch:
    """N independent streams scored together on one device (rp_batch_*)."""

    def __init__(self, n_streams: int, config: Config | None = None, device: int = 0, devices=None):
        """devices: list of CUDA device ids -> rp_batch_create_multi (streams sharded contiguously, host audio only)."""
        self._L = lib()
        self._h = C.c_void_p()
        cfg = config if config is not None else default_config()
        if devices is None:
            _check(self._L.rp_batch_create(C.byref(cfg), n_streams, device, C.byref(self._h)))
        else:
            ids = (C.c_int * len(devices))(*devices)
            _check(self._L.rp_batch_create_multi(C.byref(cfg), n_streams, ids, len(devices), C.byref(self._h)))
        self.n_streams = n_streams
        self.channels = int(cfg.channels)

    def close(self):
        if getattr(self, "_h", None):
            self._L.rp_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def add_wakeword_from_buffer(self, key: str, buf: bytes):
        _check(self._L.rp_batch_add_wakeword_from_buffer(self._h, key.encode(), buf, len(buf)), self._h)

    def add_wakeword_from_file(self, key: str, path: str):
        _check(self._L.rp_batch_add_wakeword_from_file(self._h, key.encode(), path.encode()), self._h)

    def remove_wakeword(self, key: str) -> bool:
        return bool(_check(self._L.rp_batch_remove_wakeword(self._h, key.encode()), self._h))

    def remove_wakewords(self) -> bool:
        return bool(_check(self._L.rp_batch_remove_wakewords(self._h), self._h))

    def n_devices(self) -> int:
        return self._L.rp_batch_n_devices(self._h)

    def get_samples_per_frame(self) -> int:
        return int(self._L.rp_batch_samples_per_frame(self._h))

    def set_cuda_stream(self, stream_handle: int):
        _check(self._L.rp_batch_set_cuda_stream(self._h, C.c_void_p(stream_handle)), self._h)

    def _dets(self, dets, n):
        return [(int(dets[i].stream), int(dets[i].chunk), dets[i].det.to_dict()) for i in range(n.value)]

    def process_ptr(self, ptr: int, samples_per_stream: int, on_device: bool, fmt: str = "f32", parse: bool = True):
        """audio at raw address `ptr` ([n_streams][samples_per_stream] of `fmt`: i8/i16/i32/f32, interleaved channels)."""
        dets = C.POINTER(CBatchDetection)()
        n = C.c_int64()
        if fmt == "f32" and self.channels == 1:
            _check(self._L.rp_batch_process(self._h, C.c_void_p(ptr), samples_per_stream, int(on_device), C.byref(dets), C.byref(n)), self._h)
        else:
            _check(self._L.rp_batch_process_samples(self._h, C.c_void_p(ptr), SAMPLE_FORMATS[fmt], samples_per_stream, int(on_device),
                                                    C.byref(dets), C.byref(n)), self._h)
        return self._dets(dets, n) if parse else int(n.value)

    def process_count(self, audio) -> int:
        """process() without building Python objects: returns the number of detections of the call (the
        rp_batch_detection records stay readable through the C ABI until the next call)."""
        return self.process(audio, parse=False)

    def process(self, audio, parse: bool = True):
        """process_samples<T> for every stream. audio: [n_streams][S] int8/int16/int32/float32 — a host numpy array or a
        torch tensor (CPU pinned/pageable or CUDA)."""
        if hasattr(audio, "data_ptr"):  # torch tensor
            import torch
            fmt = {torch.float32: "f32", torch.int16: "i16", torch.int32: "i32", torch.int8: "i8"}[audio.dtype]
            assert audio.is_contiguous() and audio.shape[0] == self.n_streams
            return self.process_ptr(audio.data_ptr(), int(audio.shape[1]), audio.is_cuda, fmt, parse)
        a = np.ascontiguousarray(audio)
        fmt = {np.dtype(np.float32): "f32", np.dtype(np.int16): "i16", np.dtype(np.int32): "i32", np.dtype(np.int8): "i8"}.get(a.dtype)
        if fmt is None:
            a, fmt = np.ascontiguousarray(audio, dtype=np.float32), "f32"
        assert a.shape[0] == self.n_streams
        return self.process_ptr(a.ctypes.data, a.shape[1], False, fmt, parse)

    def process_bytes(self, audio_bytes, bytes_per_stream: int | None = None, on_device: bool = False):
        """process_bytes for every stream: raw bytes in the config's sample format / endianness / channels.
        audio_bytes: bytes-like / uint8 numpy [n_streams][bytes_per_stream], or a raw address with bytes_per_stream."""
        dets = C.POINTER(CBatchDetection)()
        n = C.c_int64()
        if isinstance(audio_bytes, int):
            ptr, per = audio_bytes, int(bytes_per_stream)
        else:
            a = np.ascontiguousarray(np.frombuffer(audio_bytes, np.uint8) if isinstance(audio_bytes, (bytes, bytearray)) else audio_bytes)
            a = a.view(np.uint8).reshape(self.n_streams, -1)
            ptr, per = a.ctypes.data, a.shape[1]
        _check(self._L.rp_batch_process_bytes(self._h, C.c_void_p(ptr), per, int(on_device), C.byref(dets), C.byref(n)), self._h)
        return self._dets(dets, n)

    def last_gate_stats(self):
        """(tiles, passed) of the avg gate in the last process() — see rp_batch_last_gate_stats."""
        t, p = C.c_int64(), C.c_int64()
        _check(self._L.rp_batch_last_gate_stats(self._h, C.byref(t), C.byref(p)), self._h)
        return int(t.value), int(p.value)

    def update_config(self, config: Config):
        _check(self._L.rp_batch_update_config(self._h, C.byref(config)), self._h)

    def reset(self):
        self._L.rp_batch_reset(self._h)

    def windows_scored(self) -> int:
        return int(self._L.rp_batch_windows_scored(self._h))

    def max_mfcc_frames(self) -> int:
        return self._L.rp_batch_max_mfcc_frames(self._h)

    def last_timings(self) -> dict:
        ms = (C.c_float * 5)()
        n = self._L.rp_batch_last_timings(self._h, ms, 5)
        keys = ["h2d_ms", "mfcc_ms", "dtw_ms", "d2h_ms", "host_ms"]
        return {keys[i]: float(ms[i]) for i in range(n)}

    def last_launches(self) -> int:
        return self._L.rp_batch_last_launches(self._h)

    def last_scores(self, n_new: int, n_slots: int) -> np.ndarray:
        """Dense [n_streams][n_new][n_slots] window scores of the last process() (parity-test tap)."""
        out = np.zeros((self.n_streams, n_new, n_slots), np.float32)
        a, b = C.c_int32(), C.c_int32()
        n = self._L.rp_batch_copy_last_scores(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(a), C.byref(b))
        _check(int(n) if n < 0 else 0, self._h)
        assert (a.value, b.value) == (n_new, n_slots), (a.value, b.value)
        return out


# ---------------------------------------------------------------- raw kernels (torch CUDA tensors)
def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def mfcc_frames(audio, mfcc_size: int, stream=None):
    """K1. audio: CUDA float32 [n_streams][S] -> [n_streams][S/160-3][mfcc_size] (fresh-extractor semantics)."""
    import torch
    assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 2
    B, S = audio.shape
    frames = max(S // 160 - 3, 0)
    out = torch.empty((B, frames, mfcc_size), dtype=torch.float32, device=audio.device)
    with torch.cuda.device(audio.device):
        _check(lib().rp_mfcc_frames(C.c_void_p(audio.data_ptr()), B, S, mfcc_size, C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def dtw_scores(tmpl, win, band: int = 5, score_ref: float = 0.22, cmn: bool = False, tmpl_off=None, tmpl_len=None,
               win_off=None, win_len=None, max_tmpl_len=None, max_win_len=None, d=None, n_pairs=None, out=None, stream=None):
    """K2. Dense form: tmpl [P][m][d], win [P][n][d] CUDA float32. Ragged form: flat tmpl/win plus
    int64 offsets (floats) and int32 lengths (CUDA tensors) and the maximum lengths."""
    import torch
    dev = tmpl.device
    if tmpl_off is None:
        P, m, dd = tmpl.shape
        n = win.shape[1]
        assert win.shape[0] == P and win.shape[2] == dd and tmpl.is_contiguous() and win.is_contiguous()
        args = (None, None, m, None, None, n)
    else:
        P, dd, m, n = n_pairs, d, max_tmpl_len, max_win_len
        args = (tmpl_off, tmpl_len, m, win_off, win_len, n)
    if out is None:
        out = torch.empty(P, dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    with torch.cuda.device(dev):
        _check(lib().rp_dtw_scores(p(tmpl), p(args[0]), p(args[1]), args[2], p(win), p(args[3]), p(args[4]), args[5], P, dd,
                                   band, score_ref, int(cmn), p(out), _stream_ptr(stream)))
    return out


def set_dtw_variant(v: int):
    lib().rp_set_dtw_variant(v)


def set_mfcc_variant(v: int):
    lib().rp_set_mfcc_variant(v)


def resample_to_16k(samples, sample_rate_in: int):
    """The AudioEncoder's resampling stage alone (host code; rubato FftFixedInOut restated): mono f32 at sample_rate_in ->
    16 kHz, whole chunks only. Returns (output float32 array, input chunk length)."""
    a = np.ascontiguousarray(samples, np.float32)
    chunk = C.c_size_t()
    n = _check(int(lib().rp_resample_to_16k(sample_rate_in, a.ctypes.data_as(C.POINTER(C.c_float)), a.size, None, 0, C.byref(chunk))))
    out = np.zeros(n, np.float32)
    _check(int(lib().rp_resample_to_16k(sample_rate_in, a.ctypes.data_as(C.POINTER(C.c_float)), a.size,
                                        out.ctypes.data_as(C.POINTER(C.c_float)), out.size, None)))
    return out, int(chunk.value)


def set_avg_gate(mode: int):
    """1 / -1 (default): avg gate first, templates only where it can pass; 0: dense scoring (parity taps, A/B)."""
    lib().rp_set_avg_gate(mode)


# ---------------------------------------------------------------- wakeword builder
def build_wakeword(name: str, samples: list[tuple[str, bytes]], mfcc_size: int = 16, threshold: float | None = None,
                   avg_threshold: float | None = None, from_files: bool = True, device: int = 0) -> bytes:
    """`WakewordRef::new_from_sample_files` (from_files: rms_level = median over the samples) or
    `new_from_sample_buffers` (max) + `save_to_buffer` (reference wakeword_ref_build.rs:9-110).
    samples: [(sample name, whole 16 kHz WAV file bytes)]. Returns the .rpw bytes."""
    L = lib()
    L.rp_wakeword_build.restype = C.c_int64
    L.rp_wakeword_build.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_char_p),
                                    C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    names = (C.c_char_p * len(samples))(*[n.encode() for n, _ in samples])
    bufs = (C.c_char_p * len(samples))(*[b for _, b in samples])
    lens = (C.c_size_t * len(samples))(*[len(b) for _, b in samples])
    args = (name.encode(), int(threshold is not None), float(threshold or 0.0), int(avg_threshold is not None),
            float(avg_threshold or 0.0), len(samples), names, bufs, lens, mfcc_size, int(from_files), device)
    n = _check(L.rp_wakeword_build(*args, None, 0))
    out = C.create_string_buffer(n)
    _check(L.rp_wakeword_build(*args, out, n))
    return out.raw[:n]


# ---------------------------------------------------------------- host-logic hooks (no GPU)
def stream4_ctl(m: int, n: int, band: int):
    """Control words of the streaming DTW kernel's consumer warps: uint32 [4][steps + 1] (column 0 unused), or None."""
    L = lib()
    L.rp_debug_stream4_ctl.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_size_t]
    L.rp_debug_stream4_ctl.restype = C.c_int
    out = np.zeros(4 * 256, np.uint32)
    steps = L.rp_debug_stream4_ctl(m, n, band, out.ctypes.data_as(C.POINTER(C.c_uint32)), out.size)
    return out[:4 * (steps + 1)].reshape(4, steps + 1).copy() if steps > 0 else None


def stream4_schedule(m: int, n: int, band: int):
    """Producer schedule of the streaming DTW kernel: uint16 [batches][4], or None if the kernel does not take the shape."""
    L = lib()
    L.rp_debug_stream4_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_size_t]
    L.rp_debug_stream4_schedule.restype = C.c_int
    out = np.zeros((128, 4), np.uint16)
    nb = L.rp_debug_stream4_schedule(m, n, band, out.ctypes.data_as(C.POINTER(C.c_uint16)), out.size)
    return out[:nb].copy() if nb > 0 else None


def wakeword_inspect(buf: bytes) -> dict:
    info = WakewordInfo()
    _check(lib().rp_wakeword_inspect(buf, len(buf), C.byref(info)))
    return {k: (getattr(info, k).decode() if k == "name" else getattr(info, k)) for k, _ in WakewordInfo._fields_}


def wakeword_template(buf: bytes, t: int, mfcc_size: int):
    rows = _check(lib().rp_wakeword_template(buf, len(buf), t, None, None, 0))
    if rows == 0:
        return "", None
    name = C.create_string_buffer(NAME_MAX)
    out = np.zeros((rows, mfcc_size), np.float32)
    _check(lib().rp_wakeword_template(buf, len(buf), t, name, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
    return name.value.decode(), out


def wakeword_from_features(name: str, templates: list[tuple[str, "np.ndarray"]], rms_level: float, threshold: float | None = None,
                           avg_threshold: float | None = None) -> bytes:
    """Averages already normalised template matrices (MfccAverager) and serialises the WakewordRef."""
    L = lib()
    L.rp_wakeword_from_features.restype = C.c_int64
    FP = C.POINTER(C.c_float)
    L.rp_wakeword_from_features.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int,
                                            C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(FP), C.c_float,
                                            C.c_void_p, C.c_size_t]
    mats = [np.ascontiguousarray(m, np.float32) for _, m in templates]
    d = mats[0].shape[1] if mats else 1
    names = (C.c_char_p * len(mats))(*[n.encode() for n, _ in templates])
    frames = (C.c_int32 * len(mats))(*[m.shape[0] for m in mats])
    data = (FP * len(mats))(*[m.ctypes.data_as(FP) for m in mats])
    args = (name.encode(), int(threshold is not None), float(threshold or 0.0), int(avg_threshold is not None),
            float(avg_threshold or 0.0), d, len(mats), names, frames, data, float(rms_level))
    n = _check(L.rp_wakeword_from_features(*args, None, 0))
    out = C.create_string_buffer(n)
    _check(L.rp_wakeword_from_features(*args, out, n))
    return out.raw[:n]


def host_replay(config: Config, rpws: list[bytes], scores, vad_values=None, max_out: int = 64):
    """Runs the product's host state machine over a dense [n_frames][n_slots] score tensor."""
    s = np.ascontiguousarray(scores, np.float32)
    bufs = (C.c_char_p * len(rpws))(*rpws)
    lens = (C.c_size_t * len(rpws))(*[len(r) for r in rpws])
    out = (CBatchDetection * max_out)()
    n = C.c_int64()
    ws = C.c_uint64()
    vv = None
    if vad_values is not None:
        vv = np.ascontiguousarray(vad_values, np.float32)
    _check(lib().rp_host_replay(C.byref(config), bufs, lens, len(rpws), s.ctypes.data_as(C.POINTER(C.c_float)), s.shape[0],
                                s.shape[1], vv.ctypes.data_as(C.POINTER(C.c_float)) if vv is not None else None, out, max_out,
                                C.byref(n), C.byref(ws)))
    dets = [(int(out[i].chunk), out[i].det.to_dict()) for i in range(min(n.value, max_out))]
    return dets, int(ws.value)
