"""ctypes binding of the CPU oracle (oracle/rp_oracle.cpp). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module; nothing under rustpotter_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LOCK = threading.Lock()
_LIBS: dict[str, C.CDLL] = {}

NAME_MAX = 128
MAX_SCORES = 64

SCORE_MODES = {"average": 0, "max": 1, "median": 2, "p25": 3, "p50": 4, "p75": 5, "p80": 6, "p90": 7, "p95": 8}
SAMPLE_FORMATS = {"i8": 0, "i16": 1, "i32": 2, "f32": 3}
VAD_MODES = {None: -1, "easy": 0, "medium": 1, "hard": 2}


class Config(C.Structure):
    """Mirror of rpo_config / rp_config (same layout)."""

    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("sample_format", C.c_uint32),
        ("channels", C.c_uint32),
        ("endianness", C.c_uint32),
        ("avg_threshold", C.c_float),
        ("threshold", C.c_float),
        ("min_scores", C.c_uint64),
        ("eager", C.c_uint32),
        ("score_ref", C.c_float),
        ("band_size", C.c_uint32),
        ("score_mode", C.c_uint32),
        ("vad_mode", C.c_int32),
        ("gain_normalizer_enabled", C.c_uint32),
        ("gain_ref_set", C.c_uint32),
        ("gain_ref", C.c_float),
        ("min_gain", C.c_float),
        ("max_gain", C.c_float),
        ("band_pass_enabled", C.c_uint32),
        ("low_cutoff", C.c_float),
        ("high_cutoff", C.c_float),
    ]


class CDetection(C.Structure):
    _fields_ = [
        ("name", C.c_char * NAME_MAX),
        ("avg_score", C.c_float),
        ("score", C.c_float),
        ("counter", C.c_uint64),
        ("gain", C.c_float),
        ("n_scores", C.c_uint32),
        ("score_names", (C.c_char * NAME_MAX) * MAX_SCORES),
        ("score_values", C.c_float * MAX_SCORES),
    ]

    def to_dict(self):
        return {
            "name": self.name.decode(),
            "avg_score": np.float32(self.avg_score),
            "score": np.float32(self.score),
            "counter": int(self.counter),
            "gain": np.float32(self.gain),
            "scores": {self.score_names[i].value.decode(): np.float32(self.score_values[i]) for i in range(self.n_scores)},
        }


def build(native: bool = False) -> str:
    """Compile the oracle (idempotent). Returns the path of the shared library."""
    name = "librp_oracle_native.so" if native else "librp_oracle.so"
    out = os.path.join(_HERE, "build", name)
    src = [os.path.join(_HERE, "rp_oracle.cpp"), os.path.join(_HERE, "rp_oracle.h")]
    stamp = out + ".host"
    host_id = _host_id() if native else "generic"
    fresh = os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in src)
    if fresh and os.path.exists(stamp) and open(stamp).read() == host_id:
        return out
    with _LOCK:
        subprocess.run(["make", "-C", _HERE, "-B", os.path.join("build", name)], check=True, capture_output=True)
        with open(stamp, "w") as f:
            f.write(host_id)
    return out


def _host_id() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def lib(native: bool = False) -> C.CDLL:
    key = "native" if native else "generic"
    if key in _LIBS:
        return _LIBS[key]
    L = C.CDLL(build(native))
    f32p = C.POINTER(C.c_float)
    L.rpo_config_default.argtypes = [C.POINTER(Config)]
    L.rpo_mfcc_num_frames.restype = C.c_size_t
    L.rpo_mfcc_num_frames.argtypes = [C.c_size_t]
    L.rpo_mfcc_stream.restype = C.c_size_t
    L.rpo_mfcc_stream.argtypes = [f32p, C.c_size_t, C.c_int, f32p]
    L.rpo_mfcc_frame.argtypes = [f32p, C.c_int, f32p]
    L.rpo_mel_centres.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.rpo_hamming.argtypes = [f32p]
    L.rpo_mel_bank.argtypes = [C.c_int, f32p]
    L.rpo_dtw_cost.restype = C.c_float
    L.rpo_dtw_cost.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int]
    L.rpo_compare.restype = C.c_float
    L.rpo_compare.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_float]
    L.rpo_normalize.argtypes = [f32p, C.c_int, C.c_int]
    L.rpo_compare_pairs.argtypes = [f32p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), f32p, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, f32p, C.c_int]
    L.rpo_aggregate.restype = C.c_float
    L.rpo_aggregate.argtypes = [f32p, C.c_int, C.c_int]
    L.rpo_wakeword_load.restype = C.c_void_p
    L.rpo_wakeword_load.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.rpo_wakeword_free.argtypes = [C.c_void_p]
    L.rpo_wakeword_name.restype = C.c_char_p
    L.rpo_wakeword_name.argtypes = [C.c_void_p]
    L.rpo_wakeword_mfcc_size.argtypes = [C.c_void_p]
    L.rpo_wakeword_num_templates.argtypes = [C.c_void_p]
    L.rpo_wakeword_template_name.restype = C.c_char_p
    L.rpo_wakeword_template_name.argtypes = [C.c_void_p, C.c_int]
    L.rpo_wakeword_template_frames.argtypes = [C.c_void_p, C.c_int]
    L.rpo_wakeword_template_data.restype = f32p
    L.rpo_wakeword_template_data.argtypes = [C.c_void_p, C.c_int]
    L.rpo_wakeword_rms_level.restype = C.c_float
    L.rpo_wakeword_rms_level.argtypes = [C.c_void_p]
    L.rpo_wakeword_threshold.argtypes = [C.c_void_p, f32p]
    L.rpo_wakeword_avg_threshold.argtypes = [C.c_void_p, f32p]
    L.rpo_wakeword_encode.restype = C.c_size_t
    L.rpo_wakeword_encode.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                      C.POINTER(f32p), C.c_int, f32p, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float,
                                      C.c_int, C.c_char_p, C.c_size_t]
    L.rpo_wakeword_build.restype = C.c_size_t
    L.rpo_wakeword_build.argtypes = [C.c_char_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_char_p),
                                     C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_char_p, C.c_size_t,
                                     C.c_char_p, C.c_size_t]
    L.rpo_resample_to_16k.restype = C.c_int64
    L.rpo_resample_to_16k.argtypes = [C.c_uint32, f32p, C.c_size_t, f32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rpo_detector_new.restype = C.c_void_p
    L.rpo_detector_new.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t]
    L.rpo_detector_free.argtypes = [C.c_void_p]
    L.rpo_detector_add_wakeword_from_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    L.rpo_detector_remove_wakeword.argtypes = [C.c_void_p, C.c_char_p]
    L.rpo_detector_remove_wakewords.argtypes = [C.c_void_p]
    L.rpo_detector_samples_per_frame.restype = C.c_size_t
    L.rpo_detector_samples_per_frame.argtypes = [C.c_void_p]
    L.rpo_detector_bytes_per_frame.restype = C.c_size_t
    L.rpo_detector_bytes_per_frame.argtypes = [C.c_void_p]
    L.rpo_detector_process_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(CDetection)]
    for nm, ct in (("f32", C.c_float), ("i16", C.c_int16), ("i32", C.c_int32), ("i8", C.c_int8)):
        getattr(L, f"rpo_detector_process_{nm}").argtypes = [C.c_void_p, C.POINTER(ct), C.c_size_t, C.POINTER(CDetection)]
    L.rpo_detector_get_partial.argtypes = [C.c_void_p, C.POINTER(CDetection)]
    for nm in ("rms_level", "gain", "rms_level_ref"):
        getattr(L, f"rpo_detector_{nm}").restype = C.c_float
        getattr(L, f"rpo_detector_{nm}").argtypes = [C.c_void_p]
    L.rpo_detector_update_config.argtypes = [C.c_void_p, C.POINTER(Config)]
    L.rpo_detector_reset.argtypes = [C.c_void_p]
    L.rpo_detector_windows_scored.restype = C.c_uint64
    L.rpo_detector_windows_scored.argtypes = [C.c_void_p]
    L.rpo_trace_window_scores.restype = C.c_size_t
    L.rpo_trace_window_scores.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t, f32p, C.c_size_t, f32p, C.c_size_t]
    L.rpo_run_streams.restype = C.c_uint64
    L.rpo_run_streams.argtypes = [C.POINTER(Config), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_int, f32p, C.c_int64,
                                  C.c_int64, C.c_int, C.POINTER(C.c_int32), C.POINTER(CDetection), C.c_int]
    _LIBS[key] = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_config(**kw) -> Config:
    c = Config()
    lib().rpo_config_default(C.byref(c))
    for k, v in kw.items():
        if k == "score_mode" and isinstance(v, str):
            v = SCORE_MODES[v.lower()]
        if k == "sample_format" and isinstance(v, str):
            v = SAMPLE_FORMATS[v.lower()]
        if k == "vad_mode" and (v is None or isinstance(v, str)):
            v = VAD_MODES[v]
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


# ---------------------------------------------------------------- MFCC / DTW
def mfcc_stream(audio, mfcc_size: int) -> np.ndarray:
    a = _f32(audio)
    n = lib().rpo_mfcc_num_frames(a.size)
    out = np.zeros((n, mfcc_size), np.float32)
    got = lib().rpo_mfcc_stream(_p(a), a.size, mfcc_size, _p(out))
    assert got == n, (got, n)
    return out


def mfcc_frame(samples480, mfcc_size: int) -> np.ndarray:
    a = _f32(samples480)
    assert a.size == 480
    out = np.zeros(mfcc_size, np.float32)
    lib().rpo_mfcc_frame(_p(a), mfcc_size, _p(out))
    return out


def mel_centres(mfcc_size: int) -> np.ndarray:
    out = (C.c_int * (mfcc_size + 3))()
    lib().rpo_mel_centres(mfcc_size, out)
    return np.array(list(out))


def hamming() -> np.ndarray:
    out = np.zeros(480, np.float32)
    lib().rpo_hamming(_p(out))
    return out


def mel_bank(mfcc_size: int) -> np.ndarray:
    out = np.zeros((mfcc_size + 1, 240), np.float32)
    lib().rpo_mel_bank(mfcc_size, _p(out))
    return out


def dtw_cost(a, b, band: int) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().rpo_dtw_cost(_p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], band))


def compare(a, b, band: int = 5, score_ref: float = 0.22) -> np.float32:
    a, b = _f32(a), _f32(b)
    return np.float32(lib().rpo_compare(_p(a), a.shape[0], _p(b), b.shape[0], a.shape[1], band, score_ref))


def normalize(frames) -> np.ndarray:
    f = _f32(frames).copy()
    lib().rpo_normalize(_p(f), f.shape[0], f.shape[1])
    return f


def compare_pairs(a, a_off, a_len, b, b_off, b_len, d, band=5, score_ref=0.22, cmn=False, n_threads=1, native=False):
    a, b = _f32(a), _f32(b)
    a_off = np.ascontiguousarray(a_off, np.int64)
    b_off = np.ascontiguousarray(b_off, np.int64)
    a_len = np.ascontiguousarray(a_len, np.int32)
    b_len = np.ascontiguousarray(b_len, np.int32)
    n = a_off.size
    out = np.zeros(n, np.float32)
    i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    lib(native).rpo_compare_pairs(_p(a), a_off.ctypes.data_as(i64p), a_len.ctypes.data_as(i32p), _p(b),
                                  b_off.ctypes.data_as(i64p), b_len.ctypes.data_as(i32p), n, d, band, score_ref,
                                  1 if cmn else 0, _p(out), n_threads)
    return out


def aggregate(scores, mode) -> np.float32:
    s = _f32(scores)
    if isinstance(mode, str):
        mode = SCORE_MODES[mode.lower()]
    return np.float32(lib().rpo_aggregate(_p(s), s.size, mode))


# ---------------------------------------------------------------- wakeword files
class Wakeword:
    def __init__(self, buf: bytes):
        err = C.create_string_buffer(256)
        self._h = lib().rpo_wakeword_load(buf, len(buf), err, 256)
        if not self._h:
            raise ValueError(err.value.decode())
        L = lib()
        self.name = L.rpo_wakeword_name(self._h).decode()
        self.mfcc_size = L.rpo_wakeword_mfcc_size(self._h)
        self.rms_level = np.float32(L.rpo_wakeword_rms_level(self._h))
        t = C.c_float()
        self.threshold = np.float32(t.value) if L.rpo_wakeword_threshold(self._h, C.byref(t)) else None
        self.avg_threshold = np.float32(t.value) if L.rpo_wakeword_avg_threshold(self._h, C.byref(t)) else None
        self.templates: list[tuple[str, np.ndarray]] = []
        for i in range(L.rpo_wakeword_num_templates(self._h)):
            self.templates.append((L.rpo_wakeword_template_name(self._h, i).decode(), self._mat(i)))
        self.avg_features = self._mat(-1) if L.rpo_wakeword_template_frames(self._h, -1) > 0 else None
        L.rpo_wakeword_free(self._h)
        self._h = None

    def _mat(self, t):
        L = lib()
        n = L.rpo_wakeword_template_frames(self._h, t)
        ptr = L.rpo_wakeword_template_data(self._h, t)
        return np.ctypeslib.as_array(ptr, shape=(n, self.mfcc_size)).copy()


def encode_wakeword(name, templates, avg=None, rms_level=0.05, threshold=None, avg_threshold=None, v2=False) -> bytes:
    """Serialise a WakewordRef (.rpw CBOR). templates: list of (name, [frames][D] array)."""
    mats = [_f32(m) for _, m in templates]
    d = mats[0].shape[1]
    names = (C.c_char_p * len(mats))(*[n.encode() for n, _ in templates])
    frames = (C.c_int32 * len(mats))(*[m.shape[0] for m in mats])
    data = (C.POINTER(C.c_float) * len(mats))(*[_p(m) for m in mats])
    avgm = _f32(avg) if avg is not None else None
    args = [name.encode(), d, len(mats), names, frames, data, 0 if avgm is None else avgm.shape[0],
            None if avgm is None else _p(avgm), float(rms_level), int(threshold is not None), float(threshold or 0),
            int(avg_threshold is not None), float(avg_threshold or 0), int(v2)]
    n = lib().rpo_wakeword_encode(*args, None, 0)
    out = C.create_string_buffer(n)
    lib().rpo_wakeword_encode(*args, out, n)
    return out.raw


def build_wakeword(name: str, samples: list[tuple[str, bytes]], mfcc_size: int, threshold=None, avg_threshold=None,
                   from_files: bool = True) -> bytes:
    """WakewordRef::new_from_sample_files / _buffers + save_to_buffer. samples: [(file name, wav bytes)]."""
    names = (C.c_char_p * len(samples))(*[n.encode() for n, _ in samples])
    bufs = (C.c_char_p * len(samples))(*[b for _, b in samples])
    lens = (C.c_size_t * len(samples))(*[len(b) for _, b in samples])
    cap = 64 + sum(len(b) for _, b in samples) * 4
    out = C.create_string_buffer(cap)
    err = C.create_string_buffer(256)
    n = lib().rpo_wakeword_build(name.encode(), int(threshold is not None), float(threshold or 0), int(avg_threshold is not None),
                                 float(avg_threshold or 0), len(samples), names, bufs, lens, mfcc_size, int(from_files), out, cap,
                                 err, 256)
    if n == 0:
        raise ValueError(err.value.decode())
    return out.raw[:n]


# ---------------------------------------------------------------- detector
class Detector:
    """Oracle twin of `Rustpotter` (reference src/detector.rs)."""

    def __init__(self, config: Config):
        err = C.create_string_buffer(256)
        self._L = lib()
        self._h = self._L.rpo_detector_new(C.byref(config), err, 256)
        if not self._h:
            raise ValueError(err.value.decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.rpo_detector_free(self._h)
            self._h = None

    def add_wakeword_from_buffer(self, key: str, buf: bytes):
        err = C.create_string_buffer(256)
        if self._L.rpo_detector_add_wakeword_from_buffer(self._h, key.encode(), buf, len(buf), err, 256) != 0:
            raise ValueError(err.value.decode())

    def add_wakeword_from_file(self, key: str, path: str):
        self.add_wakeword_from_buffer(key, open(path, "rb").read())

    def remove_wakeword(self, key):
        return bool(self._L.rpo_detector_remove_wakeword(self._h, key.encode()))

    def remove_wakewords(self):
        return bool(self._L.rpo_detector_remove_wakewords(self._h))

    def get_samples_per_frame(self):
        return self._L.rpo_detector_samples_per_frame(self._h)

    def get_bytes_per_frame(self):
        return self._L.rpo_detector_bytes_per_frame(self._h)

    def process_bytes(self, b: bytes):
        d = CDetection()
        return d.to_dict() if self._L.rpo_detector_process_bytes(self._h, b, len(b), C.byref(d)) else None

    def process_samples(self, samples):
        a = np.ascontiguousarray(samples)
        fn, ct = {np.dtype(np.float32): ("f32", C.c_float), np.dtype(np.int16): ("i16", C.c_int16),
                  np.dtype(np.int32): ("i32", C.c_int32), np.dtype(np.int8): ("i8", C.c_int8)}[a.dtype]
        d = CDetection()
        r = getattr(self._L, f"rpo_detector_process_{fn}")(self._h, a.ctypes.data_as(C.POINTER(ct)), a.size, C.byref(d))
        return d.to_dict() if r else None

    def get_partial_detection(self):
        d = CDetection()
        return d.to_dict() if self._L.rpo_detector_get_partial(self._h, C.byref(d)) else None

    def get_rms_level(self):
        return np.float32(self._L.rpo_detector_rms_level(self._h))

    def get_gain(self):
        return np.float32(self._L.rpo_detector_gain(self._h))

    def get_rms_level_ref(self):
        return np.float32(self._L.rpo_detector_rms_level_ref(self._h))

    def update_config(self, config: Config):
        self._L.rpo_detector_update_config(self._h, C.byref(config))

    def reset(self):
        self._L.rpo_detector_reset(self._h)

    def windows_scored(self):
        return int(self._L.rpo_detector_windows_scored(self._h))


def resample_to_16k(samples, sample_rate_in: int):
    """rubato FftFixedInOut restated: mono f32 at sample_rate_in -> 16 kHz (whole chunks). Returns (out, input chunk)."""
    a = _f32(samples)
    chunk = C.c_size_t()
    n = lib().rpo_resample_to_16k(sample_rate_in, _p(a), a.size, None, 0, C.byref(chunk))
    out = np.zeros(n, np.float32)
    lib().rpo_resample_to_16k(sample_rate_in, _p(a), a.size, _p(out), out.size, None)
    return out, int(chunk.value)


def trace_window_scores(config: Config, rpw: bytes, audio, n_templates: int) -> np.ndarray:
    a = _f32(audio)
    maxw = a.size // 160 + 8
    out = np.zeros((maxw, n_templates + 2), np.float32)
    n = lib().rpo_trace_window_scores(C.byref(config), rpw, len(rpw), _p(a), a.size, _p(out), maxw)
    return out[:n]


def run_streams(config: Config, rpws: list[bytes], audio, n_threads: int = 1, max_det: int = 0, native: bool = False):
    """B independent detectors over audio[B][S]. Returns (windows_scored, det_counts, detections)."""
    a = _f32(audio)
    B, S = a.shape
    L = lib(native)
    bufs = (C.c_char_p * len(rpws))(*rpws)
    lens = (C.c_size_t * len(rpws))(*[len(r) for r in rpws])
    counts = np.zeros(B, np.int32)
    dets = (CDetection * (B * max_det))() if max_det else None
    total = L.rpo_run_streams(C.byref(config), bufs, lens, len(rpws), _p(a), B, S, n_threads,
                              counts.ctypes.data_as(C.POINTER(C.c_int32)), dets, max_det)
    out = []
    if max_det:
        for b in range(B):
            out.append([dets[b * max_det + i].to_dict() for i in range(min(int(counts[b]), max_det))])
    return int(total), counts, out
