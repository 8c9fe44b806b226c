"""Shared helpers for the parity tests (mirrors the helpers at the bottom of the reference's
tests/detector.rs:296-435)."""
import os
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name: str) -> str:
    return os.path.join(GOLDEN, name)


def read_wav_i16(path: str) -> np.ndarray:
    with wave.open(path) as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1 and w.getframerate() == 16000
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()


def read_wav_buffer(path: str, gain: float) -> bytes:
    """tests/detector.rs:405-426 — strips a 44-byte header and applies an i16 gain."""
    raw = open(path, "rb").read()[44:]
    raw = raw[: len(raw) // 2 * 2]
    s = np.frombuffer(raw, dtype="<i2").astype(np.float32) * np.float32(gain)
    # f32::round = half away from zero
    r = np.where(s >= 0, np.floor(s + np.float32(0.5)), np.ceil(s - np.float32(0.5)))
    return np.clip(r, -32768, 32767).astype("<i2").tobytes()


def two_wakeword_stream(gain1: float = 1.0, gain2: float = 1.0) -> bytes:
    """tests/detector.rs:372-400: 5 s silence + sample 1 + 5 s + sample 2 + 5 s, 16 kHz i16 LE."""
    sil = bytes(16000 * 2 * 5)
    return sil + read_wav_buffer(golden("oye_casa_g_1.wav"), gain1) + sil + read_wav_buffer(golden("oye_casa_g_2.wav"), gain2) + sil


def run_detection_simulation(detector, stream: bytes):
    """tests/detector.rs:355-363: feed get_bytes_per_frame() chunks, collect detections."""
    n = detector.get_bytes_per_frame()
    out = []
    for off in range(0, len(stream) - n + 1, n):
        d = detector.process_bytes(stream[off : off + n])
        if d is not None:
            out.append(d)
    return out


def synth_audio(n_streams: int, n_samples: int, seed: int = 0x5EED) -> np.ndarray:
    """SURVEY §8d synthetic audio: 0.1*N(0,1) noise + a per-stream chirp 200->3000 Hz at 0.3,
    clipped to [-1,1]; never exactly zero."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    dur = n_samples / 16000.0
    out = np.empty((n_streams, n_samples), np.float32)
    for b in range(n_streams):
        noise = 0.1 * rng.standard_normal(n_samples)
        f0 = 200.0 + 37.0 * (b % 13)
        k = (3000.0 - f0) / max(dur, 1e-9)
        chirp = 0.3 * np.sin(2 * np.pi * (f0 * t + 0.5 * k * t * t) + 0.1 * b)
        x = np.clip(noise + chirp, -1.0, 1.0).astype(np.float32)
        x[x == 0] = np.float32(1e-4)
        out[b] = x
    return out


def synth_utterance(seed: int, n_frames: int) -> np.ndarray:
    """A deterministic speech-like burst: three gliding formants with an amplitude envelope plus a
    little noise; (n_frames + 3) hops long so a fresh MfccExtractor emits exactly n_frames frames."""
    rng = np.random.default_rng(seed)
    n = (n_frames + 3) * 160
    t = np.arange(n) / 16000.0
    x = np.zeros(n)
    for k in range(3):
        f0 = rng.uniform(250, 900) * (k + 1)
        f1 = f0 * rng.uniform(0.6, 1.6)
        ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / t[-1] * t * t)
        am = 0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(2, 7) * t + rng.uniform(0, 6))
        x += (0.25 / (k + 1)) * am * np.sin(ph)
    env = np.sin(np.pi * np.arange(n) / n) ** 0.5
    x = x * env + 0.01 * rng.standard_normal(n)
    x = np.clip(x, -1, 1).astype(np.float32)
    x[x == 0] = np.float32(1e-4)
    return x


def wakeword_utterances(lengths=(88, 92, 96, 100, 100, 96, 92, 100), seed=1234):
    """The synthetic utterances a wakeword's templates are made from: variations of one "word" — time-trimmed, slightly
    noised copies of one base utterance; utterance i gives exactly lengths[i] MFCC frames. Pure numpy (deterministic)."""
    utts = []
    base = synth_utterance(seed, max(lengths))
    for i, n in enumerate(lengths):
        rng = np.random.default_rng(seed + 17 * i + 1)
        off = int(rng.integers(0, max(lengths) - n + 1)) * 160
        u = base[off: off + (n + 3) * 160].copy()
        u = np.clip(u * np.float32(rng.uniform(0.8, 1.1)) + 0.004 * rng.standard_normal(u.size).astype(np.float32), -1, 1)
        u = u.astype(np.float32)
        u[u == 0] = np.float32(1e-4)
        utts.append(u)
    return utts


def make_wakeword(oracle, name="hey b200", d=16, lengths=(88, 92, 96, 100, 100, 96, 92, 100), seed=1234,
                  threshold=None, avg_threshold=None, with_avg=True):
    """Synthetic WakewordRef: templates are the ORACLE's MFCC + CMN of synthetic
    utterances (so they are realistic mean-normalised cepstra). Returns (rpw bytes, utterances)."""
    utts = wakeword_utterances(lengths, seed)
    tmpl = [(f"sample_{i}.wav", oracle.normalize(oracle.mfcc_stream(u, d))) for i, u in enumerate(utts)]
    avg = None
    if with_avg:
        longest = max(tmpl, key=lambda t: t[1].shape[0])[1]
        avg = longest.copy()
    rpw = oracle.encode_wakeword(name, tmpl, avg=avg, rms_level=0.05, threshold=threshold, avg_threshold=avg_threshold)
    return rpw, utts


def splice(audio: np.ndarray, utt: np.ndarray, hop_offset: int) -> None:
    """Overwrites audio[hop_offset*160 : ...] with an utterance (in place)."""
    s = hop_offset * 160
    audio[s: s + utt.size] = utt


# BASELINE configs[4] ("config 5"): four WakewordRefs, 8 templates (+ avg_features) each, D = 16, ~1 s templates
CONFIG5_NAMES = ("hey b200", "ok blackwell", "wake nvlink", "hello hbm")
CONFIG5_LENGTHS = ((88, 92, 96, 100, 100, 96, 92, 100), (84, 90, 94, 98, 100, 96, 88, 92),
                   (90, 100, 86, 94, 98, 92, 96, 100), (96, 88, 100, 92, 90, 98, 94, 86))
CONFIG5_SEEDS = (1234, 2345, 3456, 4567)


def make_config5_wakewords(oracle, d=16):
    """Returns ([rpw bytes] * 4, [[utterances of wakeword w]] * 4)."""
    rpws, utts = [], []
    for name, lengths, seed in zip(CONFIG5_NAMES, CONFIG5_LENGTHS, CONFIG5_SEEDS):
        r, u = make_wakeword(oracle, name=name, d=d, lengths=lengths, seed=seed)
        rpws.append(r)
        utts.append(u)
    return rpws, utts


def read_wav_f32(path: str):
    """Minimal RIFF reader for the reference's float fixtures (WAVE_FORMAT_EXTENSIBLE / IEEE float, which the `wave`
    module rejects). Returns (sample_rate, channels, float32 samples)."""
    import struct
    f = open(path, "rb").read()
    assert f[:4] == b"RIFF" and f[8:12] == b"WAVE"
    p, rate, ch, data = 12, None, None, None
    while p + 8 <= len(f):
        cid, sz = f[p:p + 4], struct.unpack("<I", f[p + 4:p + 8])[0]
        if cid == b"fmt ":
            tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", f[p + 8:p + 24])
            assert bits == 32 and tag in (3, 65534)
        elif cid == b"data":
            data = np.frombuffer(f[p + 8:p + 8 + sz], "<f4").copy()
            break
        p += 8 + sz + (sz & 1)
    return rate, ch, data


def real_sample_stream():
    """tests/detector.rs:296-326 run_detection_with_audio_file input: real_sample.wav (48 kHz f32 mono) + 5 s of silence."""
    rate, ch, x = read_wav_f32(golden("real_sample.wav"))
    assert (rate, ch) == (48000, 1)
    return rate, np.concatenate([x, np.zeros(rate * 5, np.float32)])


# tests/detector.rs:162-214 — the reference's goldens on a 48 kHz recording (through its rubato resampler)
REAL_SAMPLE_GOLDENS = [
    (dict(avg_threshold=0.3, threshold=0.47, score_mode="max", min_scores=5),
     [(0.4676845, 0.527971, 24), (0.32865646, 0.48120698, 7), (0.30807483, 0.5164661, 35)]),
    (dict(avg_threshold=0.3, threshold=0.49, score_mode="max", min_scores=5, gain_normalizer_enabled=1, min_gain=0.4,
          band_pass_enabled=1, low_cutoff=210.0, high_cutoff=700.0),
     [(0.45496628, 0.5380342, 23), (0.336222, 0.5001262, 5), (0.3049497, 0.5189481, 31)]),
]
