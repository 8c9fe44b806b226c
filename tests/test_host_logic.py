"""CPU tests of the product's host side (no GPU, no compute calls): the C-ABI library loads and
exports every symbol the header declares; the .rpw reader agrees with the oracle's; the per-stream
state machine (detector.rs:377-454), fed the oracle's per-window scores, emits exactly the oracle
detector's detections."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rustpotter_b200 as rp
from oracle import oracle as O
from rustpotter_b200 import api
from tests.helpers import ROOT, golden, read_wav_i16, two_wakeword_stream

RPWS = ["oye_casa_g.rpw", "oye_casa_g_v2.rpw", "alexa.rpw", "oye_casa_real.rpw"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rustpotter_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = rp.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(api.EXPORTED) == declared


def test_defaults_match_reference():  # src/config.rs Default impls, src/constants.rs
    c = rp.default_config()
    assert (c.sample_rate, c.sample_format, c.channels, c.endianness) == (16000, 3, 1, 0)
    assert abs(c.avg_threshold - 0.2) < 1e-7 and abs(c.threshold - 0.5) < 1e-7 and c.min_scores == 5 and c.eager == 0
    assert abs(c.score_ref - 0.22) < 1e-7 and c.band_size == 5 and c.score_mode == 1 and c.vad_mode == -1
    assert c.gain_normalizer_enabled == 0 and abs(c.min_gain - 0.1) < 1e-7 and c.max_gain == 1.0
    assert c.band_pass_enabled == 0 and c.low_cutoff == 80.0 and c.high_cutoff == 400.0
    o = O.default_config()
    assert bytes(c) == bytes(o)  # identical layout and defaults as the oracle's twin struct


@pytest.mark.skipif(rp.device_count() > 0, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    h = C.c_void_p()
    cfg = rp.default_config()
    assert rp.lib().rp_create(C.byref(cfg), 0, C.byref(h)) == -2  # RP_ERR_CUDA
    assert b"no CPU fallback" in rp.lib().rp_last_error(None)
    with pytest.raises(rp.RustpotterError):
        rp.Rustpotter()
    with pytest.raises(rp.RustpotterError):
        rp.RustpotterBatch(4)


@pytest.mark.parametrize("name", RPWS)
def test_rpw_reader_matches_oracle(name):
    buf = open(golden(name), "rb").read()
    info = rp.wakeword_inspect(buf)
    ow = O.Wakeword(buf)
    assert info["name"] == ow.name and info["mfcc_size"] == ow.mfcc_size and info["n_templates"] == len(ow.templates)
    assert info["is_v2"] == (1 if "v2" in name else 0)
    assert np.float32(info["rms_level"]) == ow.rms_level
    assert info["has_threshold"] == int(ow.threshold is not None)
    assert info["max_frames"] == max(t.shape[0] for _, t in ow.templates)
    for t, (tn, tm) in enumerate(ow.templates):
        n, m = rp.wakeword_template(buf, t, ow.mfcc_size)
        assert n == tn and np.array_equal(m, tm)
    _, avg = rp.wakeword_template(buf, -1, ow.mfcc_size)
    assert np.array_equal(avg, ow.avg_features)


def test_rpw_reader_rejects_other_files():
    L = rp.lib()
    info = api.WakewordInfo()
    assert L.rp_wakeword_inspect(b"\x00\x01\x02", 3, C.byref(info)) == -3          # RP_ERR_FORMAT
    assert L.rp_wakeword_inspect(b"", 0, C.byref(info)) == -3
    # a WakewordModel-shaped map (labels/weights) is outside this path
    model = bytes([0xA2, 0x66]) + b"labels" + bytes([0x80, 0x67]) + b"weights" + bytes([0xA0])
    assert L.rp_wakeword_inspect(model, len(model), C.byref(info)) == -4            # RP_ERR_UNSUPPORTED
    good = open(golden("alexa.rpw"), "rb").read()
    assert L.rp_wakeword_inspect(good[: len(good) // 2], len(good) // 2, C.byref(info)) == -3  # truncated
    # options round trip through the oracle's writer (threshold Some / avg None)
    rng = np.random.default_rng(0)
    tm = [("a.wav", rng.standard_normal((50, 16)).astype(np.float32)), ("b.wav", rng.standard_normal((60, 16)).astype(np.float32))]
    buf = O.encode_wakeword("x", tm, avg=None, threshold=0.61, avg_threshold=None, rms_level=0.1)
    i2 = rp.wakeword_inspect(buf)
    assert i2["has_threshold"] == 1 and abs(i2["threshold"] - 0.61) < 1e-7 and i2["avg_frames"] == 0 and i2["max_frames"] == 60


def _stream_f32():
    return np.frombuffer(two_wakeword_stream(), dtype="<i2").astype(np.float32) / np.float32(32767.0)


def _dense_scores(cfg, rpw, audio):
    """Oracle per-window scores (no gates) laid out per emitted frame: [n_frames][n_slots]."""
    ww = O.Wakeword(rpw)
    T = len(ww.templates)
    tr = O.trace_window_scores(cfg, rpw, audio, T)  # rows: [avg, aggregate, s_0..s_T-1] per scored window
    n_frames = audio.size // 160 - 3
    maxf = max(t.shape[0] for _, t in ww.templates)
    assert tr.shape[0] == n_frames - maxf + 1
    has_avg = ww.avg_features is not None
    dense = np.zeros((n_frames, T + (1 if has_avg else 0)), np.float32)
    if has_avg:
        dense[maxf - 1:, 0] = tr[:, 0]
        dense[maxf - 1:, 1:] = tr[:, 2:]
    else:
        dense[maxf - 1:, :] = tr[:, 2:]
    return dense


REPLAY_CASES = [
    ("oye_casa_g.rpw", dict(score_mode="max")),
    ("oye_casa_g.rpw", dict(score_mode="median")),
    ("oye_casa_g.rpw", dict(score_mode="average")),
    ("oye_casa_g.rpw", dict(score_mode="p90", min_scores=2)),
    ("oye_casa_g.rpw", dict(score_mode="max", eager=1, min_scores=3)),
    ("oye_casa_g.rpw", dict(score_mode="max", min_scores=30)),          # partials dropped, never emitted
    ("oye_casa_g.rpw", dict(score_mode="max", avg_threshold=0.0, threshold=0.3, min_scores=0)),
    ("oye_casa_g.rpw", dict(score_mode="max", vad_mode="easy")),
    ("oye_casa_g.rpw", dict(score_mode="max", vad_mode="hard", threshold=0.4)),
    ("alexa.rpw", dict(score_mode="max", avg_threshold=0.0, threshold=0.45, min_scores=0)),
    ("alexa.rpw", dict(score_mode="max", avg_threshold=0.0, threshold=0.2, min_scores=1)),
    ("oye_casa_real.rpw", dict(score_mode="p25", threshold=0.3, avg_threshold=0.1, min_scores=1)),
]


@pytest.mark.parametrize("rpw_name,kw", REPLAY_CASES)
def test_host_state_machine_matches_oracle_detector(rpw_name, kw):
    rpw = open(golden(rpw_name), "rb").read()
    audio = _stream_f32()
    audio = audio[: audio.size // 480 * 480]
    cfg_o = O.default_config(**kw)
    dense = _dense_scores(O.default_config(**{**kw, 'vad_mode': None}), rpw, audio)  # trace scores every window
    vad = None
    if kw.get("vad_mode"):
        mf = O.mfcc_stream(audio, 5)
        # mean |mfcc| summed in coefficient order, f32 (vad.rs:12)
        acc = np.zeros(mf.shape[0], np.float32)
        for k in range(mf.shape[1]):
            acc = (acc + np.abs(mf[:, k])).astype(np.float32)
        vad = (acc / np.float32(mf.shape[1])).astype(np.float32)
    # oracle detector on the same audio
    det = O.Detector(cfg_o)
    det.add_wakeword_from_buffer("w", rpw)
    want = []
    for c in range(audio.size // 480):
        d = det.process_samples(audio[480 * c: 480 * (c + 1)])
        if d is not None:
            want.append((c, d))
    got, scored = rp.host_replay(rp.default_config(**kw), [rpw], dense, vad_values=vad)
    assert scored == det.windows_scored()
    assert len(got) == len(want), (got, want)
    for (gc, gd), (wc, wd) in zip(got, want):
        assert gc == wc
        assert gd["name"] == wd["name"] and gd["counter"] == wd["counter"]
        assert gd["score"] == wd["score"] and gd["avg_score"] == wd["avg_score"]
        assert gd["scores"] == wd["scores"]


@pytest.mark.parametrize("name", ["oye_casa_g.rpw", "alexa.rpw"])
def test_averager_and_cbor_writer_reproduce_reference_file_bytes(name):
    """rp_wakeword_from_features (MfccAverager::average, averager.rs:5-37, with the unbanded DTW path of
    dtw.rs:11-55,106-138, + the ciborium layout of WakewordRef) fed the templates stored in a reference-built
    .rpw must give back that file byte for byte: same averaged matrix bits, same field order, same float widths."""
    buf = open(golden(name), "rb").read()
    info = rp.wakeword_inspect(buf)
    templates = [rp.wakeword_template(buf, t, info["mfcc_size"]) for t in range(info["n_templates"])]
    out = rp.wakeword_from_features(info["name"], templates, info["rms_level"])
    assert out == buf
    # and the oracle reads the product-written file
    ww = O.Wakeword(out)
    assert ww.name == info["name"] and len(ww.templates) == info["n_templates"]


def test_cbor_writer_options_and_float_widths():
    """Option fields (threshold / avg_threshold / avg_features) and ciborium's smallest-lossless float width."""
    rng = np.random.default_rng(5)
    t0 = rng.standard_normal((7, 3)).astype(np.float32)
    t0[0] = [0.0, 0.5, -2.0]                 # exact halves -> 3-byte floats
    t0[1, 0] = np.float32(6.1035156e-05)     # smallest normal half
    t0[1, 1] = np.float32(5.9604645e-08)     # smallest subnormal half
    t0[1, 2] = np.float32(65504.0)           # largest half
    t0[2, 0] = np.float32(65536.0)           # not a half
    t0[2, 1] = np.float32(1e-30)
    out = rp.wakeword_from_features("w", [("only", t0)], 0.25, threshold=0.5, avg_threshold=0.125)
    info = rp.wakeword_inspect(out)
    assert info["avg_frames"] == 0 and info["has_threshold"] and info["has_avg_threshold"]
    assert info["threshold"] == 0.5 and info["avg_threshold"] == 0.125 and info["rms_level"] == 0.25
    name, back = rp.wakeword_template(out, 0, 3)
    assert name == "only" and np.array_equal(back, t0)
    ww = O.Wakeword(out)                     # the oracle's reader agrees
    assert np.array_equal(dict(ww.templates)["only"], t0) and ww.avg_features is None
    # 7*3 floats: 6 of them fit a half (3 bytes instead of 5)
    ref = O.encode_wakeword("w", [("only", t0)], None, 0.25, threshold=0.5, avg_threshold=0.125)
    assert len(out) <= len(ref)
    with pytest.raises(rp.RustpotterError):
        rp.wakeword_from_features("w", [], 0.1)


@pytest.mark.skipif(rp.device_count() > 0, reason="checks the no-GPU failure mode")
def test_builder_fails_loudly_without_gpu():
    wavs = [(w, open(golden(w), "rb").read()) for w in ("alexa.wav", "alexa2.wav")]
    with pytest.raises(rp.RustpotterError):
        rp.build_wakeword("alexa", wavs, 5)


def test_judgement_modes_match_oracle_aggregate():
    """score_logic.h percentile/aggregate vs the oracle's (wakeword_comp.rs:38-49,108-139)."""
    rng = np.random.default_rng(5)
    for T in (1, 2, 3, 5, 8, 13):
        tm = [(f"t{i}", rng.standard_normal((4, 16)).astype(np.float32)) for i in range(T)]
        rpw = O.encode_wakeword("w", tm, avg=None)
        for mode in O.SCORE_MODES:
            s = rng.uniform(0.05, 0.7, size=T).astype(np.float32)
            dense = np.zeros((10, T), np.float32)
            dense[3] = s  # window ending at frame 3 (4 frames = longest template) is the only hit
            cfg = rp.default_config(score_mode=mode, threshold=0.0, avg_threshold=0.0, min_scores=0)
            # countdown = 4/2 = 2 => the partial detection fires two windows later
            got, _ = rp.host_replay(cfg, [rpw], dense)
            assert len(got) == 1, mode
            d = got[0][1]
            assert d["score"] == O.aggregate(s, mode), (mode, T, d["score"], O.aggregate(s, mode))
            assert d["counter"] == 1 and d["avg_score"] == 0 and list(d["scores"].values()) == list(s)


def test_mfcc_tables_build_for_every_mfcc_size(tmp_path):
    """Host table construction (mel centres, segment chunks of the two-frames-per-warp kernel) terminates
    and covers all 240 bins for every mfcc_size the path accepts (1..31)."""
    import subprocess
    import textwrap
    src = tmp_path / "t.cpp"
    src.write_text(textwrap.dedent('''
        #include "rustpotter_b200/csrc/mfcc_tables.h"
        #include <cstdio>
        int main() {
            for (int d = 1; d <= 31; d++) {
                auto t = rp::build_mfcc_tables(d);
                int bins = 0;
                for (int c = 0; c < t.n_chunks; c++) bins += t.chunks[4 * c + 2] - t.chunks[4 * c + 1];
                if (t.n_chunks > 32 || (t.n_chunks > 0 && bins != 240) || (d <= 16 && t.n_chunks == 0)) { printf("bad %d\\n", d); return 1; }
                if ((int)t.centres.size() != d + 3 || t.centres.front() != 0 || t.centres.back() != 240) { printf("centres %d\\n", d); return 1; }
            }
            printf("ok\\n");
            return 0;
        }
    '''))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-I", ROOT, "-o", str(exe), str(src), os.path.join(ROOT, "rustpotter_b200/csrc/mfcc_tables.cpp")],
                   check=True, timeout=120)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout


# ------------------------------------------------------------------ K2s v4 producer schedule (host-built, no GPU)
def _stream4_consumer_reads(m, n, band):
    """Re-derives, from the kernel's documented geometry only, which ring slots and staged blocks the consumer
    warps read in which super-step (dtw_stream4_kernel.cu: block B handles row pair u = st - 2B at step st while
    4B + 1 - w/2 <= u <= 4B + 4 + (w+1)/2; it reads row pair u, looks one row pair ahead, and switches to
    block B+4 after its last step)."""
    w = max(band, abs(m - n))
    nb = (n + 7) // 8
    half = m // 2
    steps = half + 2 * (nb - 1)
    fin0 = 4 + (w + 1) // 2
    rows, blocks = {}, {}          # row pair -> (first super-step read, last), block -> super-step of the switch read
    for B in range(nb):
        ufirst, ulast = max(1, 4 * B + 1 - w // 2), 4 * B + fin0
        for st in range(1, steps + 1):
            u = st - 2 * B
            if ufirst <= u <= ulast:
                S = (st + 1) // 2
                for k in (u, u + 1) if u + 1 <= ulast else (u,):
                    lo, hi = rows.get(k, (S, S))
                    rows[k] = (min(lo, S), max(hi, S))
                if u == ulast and st < steps and B + 4 < nb:
                    blocks[B + 4] = S
    return w, nb, steps, rows, blocks


def test_stream4_schedule_invariants():
    from rustpotter_b200 import api
    checked = 0
    for m in list(range(2, 40)) + [57, 64, 99, 100, 101, 119, 120, 121, 150, 168, 200]:
        for n in sorted({max(1, m + d) for d in (-20, -19, -13, -8, -7, -3, -1, 0, 1, 2, 5, 8, 12, 17, 20)}):
            for band in (3, 5, 8, 12, 16, 17, 20):
                sched = api.stream4_schedule(m, n, band)
                w = max(band, abs(m - n))
                if sched is None:
                    assert not (3 <= w <= 20 and m // 2 + 2 * ((n + 7) // 8 - 1) <= 236), (m, n, band)
                    continue
                w, nb, steps, rows, blocks = _stream4_consumer_reads(m, n, band)
                n_super = (steps + 1) // 2
                assert len(sched) == n_super
                row_batch, quarter_batch = {}, {}
                for c, units in enumerate(sched):
                    for code in units:
                        code = int(code)
                        if code == 0:
                            continue
                        if code & 0x8000:
                            quarter_batch[((code & 0x7fff) >> 2, code & 3)] = c
                        else:
                            assert code not in row_batch
                            row_batch[code] = c
                kmax = (m + 1) // 2
                for k, (first, last) in rows.items():
                    if k > kmax:
                        continue            # beyond the template: only rows > m-1 of non-final blocks touch these slots
                    # stored (batch c is visible from super-step c+1) before the first read ...
                    assert k in row_batch and row_batch[k] + 1 <= first, (m, n, band, k, row_batch.get(k), first)
                    # ... and not overwritten (slot k % 16) until after the last one: the producers run up to two
                    # super-steps ahead, so batch c may be stored as soon as super-step c-2 is over
                    if k + 16 in row_batch:
                        assert row_batch[k + 16] - 2 >= last, (m, n, band, k, row_batch[k + 16], last)
                for B, S in blocks.items():
                    for j in range(4):
                        assert quarter_batch.get((B, j), 1 << 30) + 1 <= S, (m, n, band, B, j, S)
                        if B + 2 in blocks or (B + 2, j) in quarter_batch:
                            assert quarter_batch.get((B + 2, j), 1 << 30) - 2 >= S, (m, n, band, B, j)
                checked += 1
    assert checked > 1500


def test_stream4_control_words():
    """The consumers' control words against the band geometry of dtw.rs:56-105 restated here: for every (warp, step) the
    active flag, the cost mask (cells (r, c) with r - w <= c <= r + w - 1, 1 <= c <= .., r >= 1 handled by +inf
    induction), the left-neighbour flags and the block sequence (b, b+4, ..) must match an independent derivation."""
    from rustpotter_b200 import api
    A, FRESH, FULL, SWITCH, OK1, OK2, OK2P, DRAIN, LEFT, NEXT, NPAR, NLAST = (1 << i for i in range(12))
    n_checked = 0
    for m, n, band in [(120, 100, 5), (100, 100, 5), (100, 100, 20), (121, 101, 5), (100, 120, 5), (41, 40, 3), (64, 72, 9),
                       (33, 17, 16), (96, 97, 20), (300, 290, 12), (9, 8, 7), (2, 1, 5), (7, 3, 4)]:
        ctl = api.stream4_ctl(m, n, band)
        w = max(band, abs(m - n))
        if not 3 <= w <= 20:
            assert ctl is None
            continue
        assert ctl is not None, (m, n, band)
        nb = (n + 7) // 8
        steps = m // 2 + 2 * (nb - 1)
        assert ctl.shape == (4, steps + 1)
        in_band = lambda r, c: r - w <= c <= r + w - 1   # noqa: E731  (dtw.rs:73-78)
        for wq in range(4):
            B, prev_active = wq, False
            for st in range(1, steps + 1):
                c = int(ctl[wq, st])
                u = st - 2 * B
                rows = (2 * u - 1, 2 * u)
                cols = range(8 * B + 1, 8 * B + 9)
                any_cell = B < nb and u >= 1 and any(in_band(r, cc) for r in rows for cc in cols)
                # active exactly while one of the block's 16 cells of this step lies inside the band
                assert bool(c & A) == any_cell, (m, n, band, wq, st, B, u)
                if not c & A:
                    prev_active = False
                    continue
                n_checked += 1
                assert bool(c & FRESH) == (not prev_active)
                mask = (c >> 12) & 0x3ff
                for j, cc in enumerate(cols):
                    assert bool((mask >> (7 - j)) & 1) == in_band(rows[0], cc), (m, n, band, wq, st, j)
                    assert bool((mask >> (8 - j)) & 1) == in_band(rows[1], cc), (m, n, band, wq, st, j)
                assert bool(c & FULL) == all(in_band(r, cc) for r in rows for cc in cols)
                assert bool(c & LEFT) == (B > 0)
                assert bool(c & OK1) == (B > 0 and in_band(rows[0], 8 * B))
                assert bool(c & OK2) == (B > 0 and in_band(rows[1], 8 * B))
                if c & FRESH:
                    assert bool(c & OK2P) == (B > 0 and in_band(rows[0] - 1, 8 * B))
                assert ((c >> 22) & 15) == u % 16
                prev_active = True
                if c & SWITCH:
                    assert st < steps
                    # last step of the block: the next row pair has no cell of this block inside the band
                    assert not any(in_band(r, cc) for r in (2 * u + 1, 2 * u + 2) for cc in cols)
                    assert bool(c & NEXT) == (B + 4 < nb) and bool(c & NPAR) == bool((B + 4) & 1) and bool(c & NLAST) == (B + 4 == nb - 1)
                    B += 4
                    prev_active = False
    assert n_checked > 2000


def test_cbor_reader_rejects_deep_nesting_without_recursing():
    """1 MB of nested arrays / tags must fail with RP_ERR_FORMAT, not overflow the host stack (advisor finding)."""
    for byte in (0x81, 0xC0, 0x9F, 0x7F):
        with pytest.raises(rp.RustpotterError) as e:
            rp.wakeword_inspect(bytes([0xA1, 0x61, 0x78]) + bytes([byte]) * (1 << 20))
        assert e.value.code == -3


@pytest.mark.parametrize("rate", [48000, 44100, 32000, 8000, 22050, 96000])
def test_resampler_matches_oracle_and_a_float64_reference(rate):
    """The product's host resampler (csrc/resampler.h, Stockham FFTs) against the oracle's restatement of rubato's
    FftFixedInOut (recursive FFT) and against the same algorithm evaluated with numpy in float64."""
    rng = np.random.default_rng(rate)
    t = np.arange(rate) / rate
    x = (0.4 * np.sin(2 * np.pi * 440.0 * t) + 0.2 * np.sin(2 * np.pi * 3100.0 * t + 1.0) + 0.05 * rng.standard_normal(rate)).astype(np.float32)
    got, chunk = rp.resample_to_16k(x, rate)
    want, chunk_o = O.resample_to_16k(x, rate)
    assert chunk == chunk_o and got.shape == want.shape and got.size > 0
    assert np.abs(got - want).max() < 3e-6
    # float64 evaluation of the published algorithm
    g = np.gcd(rate, 16000)
    k = -(-480 // (16000 // g))
    nin, nout = k * rate // g, k * 16000 // g
    assert chunk == nin and got.size == (x.size // nin) * nout
    n = np.arange(nin, dtype=np.float64)
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * n / nin) + 0.14128 * np.cos(4 * np.pi * n / nin) - 0.01168 * np.cos(6 * np.pi * n / nin)) ** 2
    rel = 1.0 / (1.0 + 42.08 / nin)
    cutoff = rel * nout / nin if nin > nout else rel
    y = w * np.sinc((n - nin // 2) * cutoff)
    y /= y.sum()
    ft = np.zeros(2 * nin)
    ft[:nin] = y / (2 * nin)
    F = np.fft.rfft(ft)
    new_len = nin + 1 if nin < nout else nout
    ref = np.zeros(got.size)
    ov = np.zeros(nout)
    for c in range(x.size // nin):
        b = np.zeros(2 * nin)
        b[:nin] = x[c * nin:(c + 1) * nin]
        X = np.fft.rfft(b)
        Y = np.zeros(nout + 1, complex)
        Y[:new_len] = X[:new_len] * F[:new_len]
        if new_len > nout:
            Y[nout] = Y[nout].real
        o = np.fft.irfft(Y, 2 * nout) * (2 * nout)
        ref[c * nout:(c + 1) * nout] = o[:nout] + ov
        ov = o[nout:]
    assert np.abs(got - ref).max() < 5e-6
    # a 440 Hz tone survives with its amplitude (unit passband gain)
    mid = ref[nout:-nout]
    assert 0.3 < np.abs(mid).max() < 0.8


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) works without a GPU and prints ONE JSON line with
    our arm's metric string, unit and config plus the keys the contract names."""
    import json
    import subprocess
    import sys

    import bench

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == bench.METRIC and j["unit"] == "windows/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert "configs[4]" in j["config"]["workload"]


def test_rpw_reader_survives_mutated_files():
    """Untrusted input across the C ABI: a few thousand corruptions of real .rpw files (byte flips, header rewrites,
    truncations, splices, length fields blown up) either parse or fail with a format error -- never crash, hang or report
    shapes the buffer cannot hold."""
    L = rp.lib()
    info = api.WakewordInfo()
    rng = np.random.default_rng(20261017)
    seeds = [open(golden(n), "rb").read() for n in ("alexa.rpw", "oye_casa_g.rpw", "oye_casa_g_v2.rpw")]
    ok = bad = 0
    for it in range(3000):
        src = bytearray(seeds[it % len(seeds)])
        kind = it % 5
        if kind == 0:       # a handful of random byte flips, biased to the head (map header, keys, array headers)
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, min(len(src), 4096 if rng.random() < 0.7 else len(src))))
                src[pos] = int(rng.integers(0, 256))
        elif kind == 1:     # truncation
            src = src[: int(rng.integers(0, len(src)))]
        elif kind == 2:     # a CBOR length byte turned into a huge 64-bit length
            pos = int(rng.integers(0, min(len(src), 2048)))
            src[pos:pos + 1] = bytes([0x9B if rng.random() < 0.5 else 0x5B]) + bytes([0x7F] + [0xFF] * 7)
        elif kind == 3:     # splice of two files
            other = seeds[(it + 1) % len(seeds)]
            cut = int(rng.integers(0, len(src)))
            src = src[:cut] + other[int(rng.integers(0, len(other))):]
        else:               # random garbage after a plausible head
            src = src[: int(rng.integers(1, 64))] + bytes(rng.integers(0, 256, int(rng.integers(0, 512)), dtype=np.uint8))
        buf = bytes(src)
        rc = L.rp_wakeword_inspect(buf, len(buf), C.byref(info))
        if rc == 0:
            ok += 1
            assert 0 < info.mfcc_size <= 0xFFFF and 0 <= info.n_templates and 0 <= info.max_frames
            # what it claims to hold must fit the buffer (f32 rows at the very least 2 bytes per float in CBOR half floats)
            assert info.n_templates * info.max_frames * info.mfcc_size * 2 <= max(len(buf), 1) * 8 or info.n_templates == 0
            for t in range(-1, min(info.n_templates, 3)):
                rows = L.rp_wakeword_template(buf, len(buf), t, None, None, 0)
                assert rows >= -8
        else:
            bad += 1
            assert rc in (-3, -4, -1), rc      # RP_ERR_FORMAT, RP_ERR_UNSUPPORTED, RP_ERR_INVALID
    assert bad > 1000 and ok >= 0
