"""Pins the CPU oracle to the reference: every golden of the reference's tests/detector.rs that does
not need the (un-vendored) rubato resampler, and the template matrices inside the reference's .rpw
fixtures (the reference's own MFCC + CMN output for the fixture wavs, tests/wakeword.rs:26-54).
Expected values are the f32 literals asserted by the reference (tests/detector.rs, cited per test).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import golden, read_wav_i16, run_detection_simulation, two_wakeword_stream

f32 = np.float32


def _sim(rpw, gains=(1.0, 1.0), **cfg):
    c = O.default_config(sample_rate=16000, sample_format="i16", channels=1, **cfg)
    det = O.Detector(c)
    det.add_wakeword_from_file("wakeword", golden(rpw))
    return run_detection_simulation(det, two_wakeword_stream(*gains))


def _check(dets, expected):
    assert len(dets) == len(expected)
    for d, e in zip(dets, expected):
        for k, v in e.items():
            # the reference asserts exact f32 equality; the oracle's FFT differs from rustfft in
            # rounding order only, so allow 2 ulp-ish (5e-7 relative)
            assert abs(float(d[k]) - v) <= 5e-7 * max(1.0, abs(v)), (k, d[k], v)


BASE = dict(avg_threshold=0.2, threshold=0.5, gain_normalizer_enabled=0, band_pass_enabled=0)


def test_v2_file():  # tests/detector.rs:9-22
    _check(_sim("oye_casa_g_v2.rpw", score_mode="max", **BASE),
           [dict(avg_score=0.6495044, score=0.7310586), dict(avg_score=0.5804737, score=0.721843)])


def test_max_score_mode():  # tests/detector.rs:25-38
    _check(_sim("oye_casa_g.rpw", score_mode="max", **BASE),
           [dict(avg_score=0.6495044, score=0.7310586), dict(avg_score=0.5804737, score=0.721843)])


def test_median_score_mode():  # tests/detector.rs:41-54
    _check(_sim("oye_casa_g.rpw", score_mode="median", **BASE),
           [dict(avg_score=0.64608675, score=0.60123634), dict(avg_score=0.5288923, score=0.63968724)])


def test_average_score_mode():  # tests/detector.rs:57-70
    _check(_sim("oye_casa_g.rpw", score_mode="average", **BASE),
           [dict(avg_score=0.64608675, score=0.60458726), dict(avg_score=0.5750509, score=0.6313083)])


def test_vad_mode():  # tests/detector.rs:73-87
    _check(_sim("oye_casa_g.rpw", score_mode="max", vad_mode="easy", **BASE),
           [dict(avg_score=0.6495044, score=0.7310586), dict(avg_score=0.5804737, score=0.721843)])


def test_ignore_words():  # tests/detector.rs:90-100
    assert _sim("alexa.rpw", score_mode="max", avg_threshold=0.0, threshold=0.45, min_scores=0) == []


def test_ignore_words_with_filters():  # tests/detector.rs:102-112
    assert _sim("alexa.rpw", score_mode="max", avg_threshold=0.0, threshold=0.45, min_scores=0,
                gain_normalizer_enabled=1, band_pass_enabled=1) == []


def test_band_pass_filter():  # tests/detector.rs:114-127
    _check(_sim("oye_casa_g.rpw", score_mode="max", avg_threshold=0.0, threshold=0.5, band_pass_enabled=1,
                low_cutoff=80.0, high_cutoff=400.0),
           [dict(score=0.6858197), dict(score=0.66327363)])


def test_gain_normalizer_filter():  # tests/detector.rs:130-142
    _check(_sim("oye_casa_g.rpw", gains=(0.2, 5.0), score_mode="max", avg_threshold=0.0, threshold=0.5,
                gain_normalizer_enabled=1),
           [dict(score=0.7304294), dict(score=0.71067876)])


def test_gain_normalizer_and_band_pass():  # tests/detector.rs:145-159
    _check(_sim("oye_casa_g.rpw", gains=(0.2, 5.0), score_mode="median", avg_threshold=0.0, threshold=0.5,
                gain_normalizer_enabled=1, band_pass_enabled=1, low_cutoff=80.0, high_cutoff=500.0),
           [dict(score=0.5775406), dict(score=0.5828697)])


@pytest.mark.parametrize("rpw,wavs", [
    ("oye_casa_g.rpw", [f"oye_casa_g_{i}.wav" for i in range(1, 6)]),
    ("alexa.rpw", ["alexa.wav", "alexa2.wav", "alexa3.wav"]),
])
def test_mfcc_matches_stored_templates(rpw, wavs):
    """The .rpw templates were produced by the reference's MfccWavFileExtractor (wav -> 480-sample
    chunks -> MfccExtractor -> whole-file CMN; src/mfcc/wav_file_extractor.rs:18-68)."""
    ww = O.Wakeword(open(golden(rpw), "rb").read())
    assert ww.mfcc_size == 5
    by_name = dict(ww.templates)
    for w in wavs:
        s = read_wav_i16(golden(w)).astype(np.float32) / f32(32767.0)
        s = s[: len(s) // 480 * 480]
        m = O.normalize(O.mfcc_stream(s, 5))
        t = by_name[w]
        assert m.shape == t.shape, (w, m.shape, t.shape)
        assert np.max(np.abs(m - t)) < 5e-5, (w, np.max(np.abs(m - t)))


BUILD_CASES = [
    ("oye_casa_g.rpw", [f"oye_casa_g_{i}.wav" for i in range(1, 6)]),
    ("alexa.rpw", ["alexa.wav", "alexa2.wav", "alexa3.wav"]),
]


@pytest.mark.parametrize("rpw,wavs", BUILD_CASES)
def test_builder_reproduces_reference_rpw(rpw, wavs):
    """WakewordRef::new_from_sample_files (wakeword_ref_build.rs:47-91) over the fixture wavs: the oracle's
    wav reader + extractor + CMN + MfccAverager + median rms must rebuild the reference's own .rpw."""
    fixture = O.Wakeword(open(golden(rpw), "rb").read())
    built = O.Wakeword(O.build_wakeword(fixture.name, [(w, open(golden(w), "rb").read()) for w in wavs], 5))
    assert built.name == fixture.name and built.mfcc_size == 5
    assert built.threshold is None and built.avg_threshold is None
    assert float(built.rms_level) == float(fixture.rms_level)           # median of per-file median chunk rms: exact
    got, want = dict(built.templates), dict(fixture.templates)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape and np.max(np.abs(got[k] - want[k])) < 5e-5, k
    assert built.avg_features.shape == fixture.avg_features.shape
    assert np.max(np.abs(built.avg_features - fixture.avg_features)) < 5e-5


def test_builder_buffers_variant_and_errors():
    """new_from_sample_buffers takes the MAX rms level (wakeword_ref_build.rs:28-30); one sample -> no average
    (:97-99); no samples -> error (wakeword_ref.rs:52-54)."""
    wavs = [(w, open(golden(w), "rb").read()) for w in BUILD_CASES[1][1]]
    files = O.Wakeword(O.build_wakeword("a", wavs, 5, from_files=True))
    bufs = O.Wakeword(O.build_wakeword("a", wavs, 5, threshold=0.4, avg_threshold=0.1, from_files=False))
    assert float(bufs.rms_level) >= float(files.rms_level)
    assert abs(bufs.threshold - 0.4) < 1e-7 and abs(bufs.avg_threshold - 0.1) < 1e-7
    single = O.Wakeword(O.build_wakeword("a", wavs[:1], 5))
    assert single.avg_features is None and len(single.templates) == 1
    with pytest.raises(ValueError):
        O.build_wakeword("a", [], 5)
    wide = O.Wakeword(O.build_wakeword("a", wavs, 16))
    assert wide.mfcc_size == 16 and wide.avg_features.shape[1] == 16


def test_fixture_shapes():
    """SURVEY §4 fixture shapes."""
    ww = O.Wakeword(open(golden("oye_casa_g.rpw"), "rb").read())
    assert ww.name == "oye casa" and [t.shape[0] for _, t in ww.templates].count(108) == 1
    assert sorted(t.shape[0] for _, t in ww.templates) == [90, 93, 96, 102, 108]
    assert ww.avg_features.shape == (108, 5) and ww.threshold is None and ww.avg_threshold is None
    assert abs(float(ww.rms_level) - 0.05257) < 1e-4
    v2 = O.Wakeword(open(golden("oye_casa_g_v2.rpw"), "rb").read())
    assert v2.mfcc_size == 5 and len(v2.templates) == 5
    real = O.Wakeword(open(golden("oye_casa_real.rpw"), "rb").read())
    assert sorted(t.shape[0] for _, t in real.templates) == [144, 147, 153, 159, 165, 168]
    al = O.Wakeword(open(golden("alexa.rpw"), "rb").read())
    assert sorted(t.shape[0] for _, t in al.templates) == [99, 117, 126] and al.avg_features.shape == (126, 5)


def test_mel_centres():
    """SURVEY §8a a4 (probe-validated centre indices)."""
    assert list(O.mel_centres(5)) == [0, 9, 22, 41, 68, 106, 161, 240]
    assert list(O.mel_centres(16)) == [0, 3, 7, 11, 16, 21, 28, 35, 43, 53, 64, 77, 92, 109, 128, 150, 176, 206, 240]


def test_dtw_quirks():
    """dtw.rs:56-105: result cell D[m-1][n]; +inf (score 0) iff n - m >= band - 1."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal((100, 16)).astype(np.float32)
    b = rng.standard_normal((120, 16)).astype(np.float32)
    assert np.isinf(O.dtw_cost(a, b, 5)) and O.compare(a, b) == 0.0       # window 120 vs template 100
    assert np.isfinite(O.dtw_cost(b, a, 5))                                  # template 120 vs window 100
    assert np.isinf(O.dtw_cost(a, a, 1)) and np.isfinite(O.dtw_cost(a, a, 2))
    # identical sequences: the diagonal costs 0, and the returned cell D[m-1][n] sits one step off
    # it, so the cost is exactly distance(a[m-2], a[m-1])
    an = a / np.linalg.norm(a, axis=1, keepdims=True)
    assert abs(float(O.dtw_cost(a, a, 5)) - (1.0 - float(an[-2] @ an[-1]))) < 1e-5
    c = a.copy(); c[-2] = c[-1]
    assert O.dtw_cost(c, c, 5) < 1e-6 and abs(float(O.compare(c, c)) - 0.7310586) < 2e-6
    z = np.zeros((10, 4), np.float32)
    # zero vectors: similarity 0 -> distance 1 per cell (comparator.rs:42-47)
    assert abs(float(O.dtw_cost(z, z, 5)) - 10.0) < 1e-6  # max(m-1, n) cells on the cheapest path


def test_percentiles():
    """wakeword_comp.rs:38-49."""
    s = np.array([0.1, 0.5, 0.3, 0.9, 0.7], np.float32)
    assert O.aggregate(s, "max") == f32(0.9)
    assert O.aggregate(s, "median") == f32(0.5) == O.aggregate(s, "p50")
    assert abs(float(O.aggregate(s, "p25")) - 0.3) < 1e-7
    assert abs(float(O.aggregate(s, "p90")) - (0.7 * 0.4 + 0.9 * 0.6)) < 1e-6
    assert abs(float(O.aggregate(s, "average")) - 0.5) < 1e-7
    s4 = np.array([0.2, 0.4, 0.6, 0.8], np.float32)
    assert abs(float(O.aggregate(s4, "median")) - 0.5) < 1e-7


def test_encode_roundtrip():
    rng = np.random.default_rng(3)
    tmpl = [(f"t{i}.wav", rng.standard_normal((n, 16)).astype(np.float32)) for i, n in enumerate([88, 100, 96])]
    avg = rng.standard_normal((100, 16)).astype(np.float32)
    for v2 in (False, True):
        buf = O.encode_wakeword("hey", tmpl, avg=avg, rms_level=0.07, threshold=0.55, v2=v2)
        ww = O.Wakeword(buf)
        assert ww.name == "hey" and ww.mfcc_size == 16 and ww.threshold == f32(0.55) and ww.avg_threshold is None
        for (n0, m0), (n1, m1) in zip(tmpl, ww.templates):
            assert n0 == n1 and np.array_equal(m0, m1)
        assert np.array_equal(ww.avg_features, avg)


def test_process_wrong_length_returns_none():  # detector.rs:235-237,249-251
    det = O.Detector(O.default_config(sample_format="i16"))
    det.add_wakeword_from_file("w", golden("oye_casa_g.rpw"))
    assert det.get_samples_per_frame() == 480 and det.get_bytes_per_frame() == 960
    assert det.process_bytes(bytes(100)) is None
    assert det.process_samples(np.zeros(479, np.int16)) is None
    det2 = O.Detector(O.default_config())
    assert det2.process_samples(np.zeros(480, np.float32)) is None  # no wakeword loaded


def test_bench_template_fixture_is_the_oracles_mfcc():
    """tests/golden/bench_templates.npz (what both arms of bench.py score against) equals the oracle's MFCC + CMN of
    the deterministic utterances it was generated from (tools/make_bench_templates.py)."""
    from tests.helpers import CONFIG5_LENGTHS, CONFIG5_SEEDS, wakeword_utterances
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_templates.npz"))
    assert len(z.files) == 32
    for w in (0, 3):
        for i, u in enumerate(wakeword_utterances(CONFIG5_LENGTHS[w], CONFIG5_SEEDS[w])):
            want = O.normalize(O.mfcc_stream(u, 16))
            got = z[f"w{w}_t{i}"]
            assert got.shape == want.shape == (CONFIG5_LENGTHS[w][i], 16)
            assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("case", [0, 1], ids=["record_with_noise", "record_with_noise_using_filters"])
def test_golden_48khz_record_through_the_restated_resampler(case):
    """tests/detector.rs:162-214: real_sample.wav is 48 kHz, so the reference feeds it through rubato's FftFixedInOut.
    The oracle's restatement reproduces both goldens (scores, avg scores and counters) — the second golden was not used to
    calibrate the restatement's cutoff."""
    from tests.helpers import REAL_SAMPLE_GOLDENS, real_sample_stream
    kw, want = REAL_SAMPLE_GOLDENS[case]
    rate, x = real_sample_stream()
    det = O.Detector(O.default_config(sample_rate=rate, sample_format="f32", channels=1, **kw))
    det.add_wakeword_from_file("wakeword", golden("oye_casa_real.rpw"))
    n = det.get_samples_per_frame()
    assert n == 1440
    got = []
    for i in range(0, len(x) - n + 1, n):
        d = det.process_samples(x[i:i + n])
        if d is not None:
            got.append(d)
    assert len(got) == len(want)
    for d, (avg, score, counter) in zip(got, want):
        assert d["counter"] == counter
        assert abs(float(d["avg_score"]) - avg) <= 2e-6 * avg and abs(float(d["score"]) - score) <= 2e-6 * score, (d, avg, score)
