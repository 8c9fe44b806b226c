"""BASELINE configs[4] ("config 5": many streams x 4 WakewordRefs, ScoreMode::Median) through the batched C ABI against
N oracle detectors, and regressions for the round-1 advisor findings. Runs on the B200 box (`-m gpu`).

Reference behaviour compared: best-of-wakewords pick src/detector.rs:433-447, Median = P50
src/wakewords/comp/wakeword_comp.rs:38-49,108-139, max_mfcc_frames over ALL wakewords src/detector.rs:328-335.
"""
import numpy as np
import pytest

import rustpotter_b200 as rp
from oracle import oracle as O
from tests.helpers import CONFIG5_LENGTHS, golden, make_config5_wakewords, make_wakeword, splice, synth_audio

pytestmark = pytest.mark.gpu
SCORE_RTOL = 1e-4


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-12)


def _config5_case(B=64, n_chunks=120, seed=55):
    rpws, utts = make_config5_wakewords(O)
    audio = synth_audio(B, n_chunks * 480, seed=seed)
    # every third stream holds an utterance of one of the four wakewords; two streams hold two different wakewords
    for b in range(0, B, 3):
        w = (b // 3) % 4
        splice(audio[b], utts[w][(b // 12) % len(utts[w])], 110 + (5 * b) % 130)
    splice(audio[1], utts[2][1], 30)
    splice(audio[1], utts[0][5], 215)
    splice(audio[2], utts[3][0], 60)
    splice(audio[2], utts[1][7], 230)
    return rpws, audio


@pytest.mark.parametrize("gated", [True, False], ids=["avg_gate_first", "dense"])
@pytest.mark.parametrize("kw", [dict(score_mode="median"), dict(score_mode="median", min_scores=2, eager=1),
                                dict(score_mode="p90", threshold=0.45)])
def test_config5_four_wakewords_median_matches_oracle(kw, gated):
    """64 streams x 4 WakewordRefs (8 templates + avg each, D=16), Median: scores, per-template scores, avg scores,
    names, counters, chunks and windows_scored equal those of 64 oracle detectors."""
    rpws, audio = _config5_case()
    cfg_o = O.default_config(**kw)
    total, counts, want = O.run_streams(cfg_o, rpws, audio, n_threads=8, max_det=8)
    rp.set_avg_gate(1 if gated else 0)
    try:
        bt = rp.RustpotterBatch(audio.shape[0], rp.default_config(**kw))
        for i, r in enumerate(rpws):
            bt.add_wakeword_from_buffer(f"w{i}", r)
        assert bt.max_mfcc_frames() == max(max(l) for l in CONFIG5_LENGTHS)
        got = bt.process(audio)
    finally:
        rp.set_avg_gate(-1)
    names = {w["name"] for b in range(audio.shape[0]) for w in want[b]}
    assert counts.sum() >= 12 and len(names) == 4, (counts.sum(), names)   # every wakeword fires somewhere
    per = {b: [] for b in range(audio.shape[0])}
    for s, c, d in got:
        per[s].append((c, d))
    worst = 0.0
    for b in range(audio.shape[0]):
        assert len(per[b]) == int(counts[b]), (b, per[b], want[b])
        for (c, d), w in zip(per[b], want[b]):
            assert d["name"] == w["name"] and d["counter"] == w["counter"], (b, d, w)
            assert set(d["scores"]) == set(w["scores"])
            for x, y in [(d["score"], w["score"]), (d["avg_score"], w["avg_score"])] + [(d["scores"][k], w["scores"][k]) for k in w["scores"]]:
                worst = max(worst, _rel(x, y))
    assert worst < SCORE_RTOL, worst
    assert bt.windows_scored() == total


def test_config5_device_audio_i16_and_sharded_handles_agree():
    """The same streams as i16 (Sample::into_f32 on the device: audio_types.rs:98-137) on one handle, and split over
    two handles (contiguous shards, what a multi-GPU run does per device): identical detections."""
    rpws, audio = _config5_case(B=32)
    pcm = np.clip(np.round(audio * 32767.0), -32768, 32767).astype(np.int16)
    want_audio = pcm.astype(np.float32) / np.float32(32767.0)
    cfg = rp.default_config(score_mode="median")
    total, counts, want = O.run_streams(O.default_config(score_mode="median"), rpws, want_audio, n_threads=8, max_det=8)

    def run(lo, hi, samples):
        bt = rp.RustpotterBatch(hi - lo, cfg)
        for i, r in enumerate(rpws):
            bt.add_wakeword_from_buffer(f"w{i}", r)
        return [(s + lo, c, d) for s, c, d in bt.process(samples[lo:hi])], bt.windows_scored()

    whole, w_all = run(0, 32, pcm)
    a, wa = run(0, 16, pcm)
    b, wb = run(16, 32, pcm)
    assert w_all == total == wa + wb
    assert [(s, c, d["name"], d["counter"]) for s, c, d in whole] == [(s, c, d["name"], d["counter"]) for s, c, d in a + b]
    flat = [w for bb in range(32) for w in want[bb]]
    assert len(whole) == len(flat) >= 6
    for (s, c, d), w in zip(whole, flat):
        assert d["name"] == w["name"] and d["counter"] == w["counter"]
        assert _rel(d["score"], w["score"]) < SCORE_RTOL and _rel(d["avg_score"], w["avg_score"]) < SCORE_RTOL


def test_batch_sample_formats_match_per_stream_conversion():
    """rp_batch_process_samples for i8 / i16 / i32 / f32 host buffers equals feeding the converted f32."""
    rpw, utts = make_wakeword(O, d=16, seed=91)
    audio = synth_audio(4, 90 * 480, seed=17)
    splice(audio[1], utts[0], 100)
    splice(audio[3], utts[3], 120)
    for fmt, dt, mx in (("i16", np.int16, 32767.0), ("i32", np.int32, 2147483647.0), ("i8", np.int8, 127.0)):
        q = np.clip(np.round(audio.astype(np.float64) * mx), -mx - 1, mx).astype(dt)
        f = (q.astype(np.float32) / np.float32(mx)) if fmt != "i32" else (q.astype(np.float32) / np.float32(mx))
        a = rp.RustpotterBatch(4)
        a.add_wakeword_from_buffer("w", rpw)
        b = rp.RustpotterBatch(4)
        b.add_wakeword_from_buffer("w", rpw)
        ga, gb = a.process(q), b.process(f)
        assert [(s, c, d["counter"]) for s, c, d in ga] == [(s, c, d["counter"]) for s, c, d in gb], fmt
        for (_, _, x), (_, _, y) in zip(ga, gb):
            assert float(x["score"]) == float(y["score"]), fmt
        assert a.windows_scored() == b.windows_scored()
        if fmt == "i16":
            assert len(ga) >= 1


# ------------------------------------------------------------------ advisor findings (round 1)
def test_partial_detection_survives_wakeword_removal():
    """A pending partial detection keeps its own name/template names (the reference keeps them by value): removing or
    replacing wakewords while it is pending must not read stale indices."""
    rpw_a, utts_a = make_wakeword(O, name="alpha", d=16, seed=100)
    rpw_b, utts_b = make_wakeword(O, name="beta", d=16, seed=200, lengths=(70, 80, 96, 76))
    det = rp.Rustpotter(rp.default_config(sample_format="f32", min_scores=1))
    det.add_wakeword_from_buffer("a", rpw_a)
    det.add_wakeword_from_buffer("b", rpw_b)
    audio = synth_audio(1, 120 * 480, seed=4)[0]
    splice(audio, utts_b[2], 150)
    partial = None
    for c in range(120):
        assert det.process_samples(audio[c * 480:(c + 1) * 480]) is None or True
        p = det.get_partial_detection()
        if p is not None:
            partial = p
            break
    assert partial is not None and partial["name"] == "beta" and len(partial["scores"]) == 4
    assert det.remove_wakeword("a") is True          # beta moves from index 1 to index 0
    p2 = det.get_partial_detection()
    assert p2 is not None and p2["name"] == "beta" and set(p2["scores"]) == set(partial["scores"])
    det.add_wakeword_from_buffer("b", make_wakeword(O, name="beta2", d=16, seed=201, lengths=(60, 64))[0])   # fewer templates
    p3 = det.get_partial_detection()
    assert p3 is not None and p3["name"] == "beta" and len(p3["scores"]) == 4
    assert det.remove_wakewords() is True
    p4 = det.get_partial_detection()
    assert p4 is not None and p4["name"] == "beta"


def test_avg_features_longer_than_every_template():
    """The reference scores avg_features against a window of max_mfcc_frames (= longest TEMPLATE) rows, so an avg matrix
    longer than every template meets n < m (wakeword_comp.rs:22-27, dtw.rs:62-67). Same scores as the oracle; no
    out-of-bounds window rows."""
    base, utts = make_wakeword(O, name="longavg", d=16, seed=321, lengths=(60, 64, 58))
    ww = O.Wakeword(base)
    avg = np.concatenate([ww.avg_features, ww.avg_features[-9:]], axis=0)        # 73 rows > 64
    rpw = O.encode_wakeword("longavg", ww.templates, avg=avg, rms_level=0.05)
    audio = synth_audio(3, 80 * 480, seed=8)
    splice(audio[2], utts[1], 90)
    total, counts, want = O.run_streams(O.default_config(), [rpw], audio, n_threads=2, max_det=4)
    bt = rp.RustpotterBatch(3)
    bt.add_wakeword_from_buffer("w", rpw)
    assert bt.max_mfcc_frames() == 64
    got = bt.process(audio)
    assert len(got) == int(counts.sum()) and bt.windows_scored() == total
    dense = bt.last_scores(80 * 3, 4)
    first = 64 + 2
    for b in range(3):
        tr = O.trace_window_scores(O.default_config(), rpw, audio[b], 3)
        w = np.concatenate([tr[:, :1], tr[:, 2:]], axis=1)
        g = dense[b, first:]
        assert g.shape == w.shape
        rel = np.abs(g - w) / np.maximum(np.abs(w), 1e-12)
        assert rel.max() < SCORE_RTOL, (b, rel.max())


def test_update_config_rejects_out_of_range_enums():
    det = rp.Rustpotter(rp.default_config(sample_format="i16"))
    det.add_wakeword_from_file("w", golden("oye_casa_g.rpw"))
    bad = rp.default_config(sample_format="i16")
    bad.score_mode = 77
    with pytest.raises(rp.RustpotterError) as e:
        det.update_config(bad)
    assert e.value.code == -1
    bad = rp.default_config(sample_format="i16")
    bad.vad_mode = 9
    with pytest.raises(rp.RustpotterError):
        det.update_config(bad)
    det.update_config(rp.default_config(sample_format="i16", score_mode="p95", vad_mode="hard"))   # still usable


# ------------------------------------------------------------------ tuned window kernel: every mfcc width <= 16, every band <= 20
@pytest.mark.parametrize("d,band", [(5, 5), (13, 5), (10, 4), (16, 1), (16, 2), (16, 3), (16, 6), (16, 8), (16, 9), (16, 12),
                                    (16, 13), (16, 20), (5, 3), (5, 16), (16, 21), (20, 5)])
def test_window_kernel_any_width_and_band(d, band):
    """Dense per-window scores of the batched front-end (tuned kernel for d <= 16 and band <= 20, generic kernel
    otherwise) against the reference-order generic kernel and the oracle's per-window trace
    (config.rs:193-208 band_size, wakeword_comp.rs:22-37)."""
    lengths = (44, 57, 50)
    rpw, utts = make_wakeword(O, d=d, lengths=lengths, seed=70 + d)
    n_chunks = 110
    audio = synth_audio(3, n_chunks * 480, seed=40 + band)
    splice(audio[0], utts[1], 120)
    splice(audio[2], utts[0], 201)
    T, maxf = len(lengths), max(lengths)
    cfg = dict(band_size=band)
    res = {}
    rp.set_avg_gate(0)
    try:
        for variant in (1, 0):
            rp.set_dtw_variant(variant)
            bt = rp.RustpotterBatch(3, rp.default_config(**cfg))
            bt.add_wakeword_from_buffer("w", rpw)
            bt.process(audio)
            res[variant] = bt.last_scores(n_chunks * 3, T + 1)
    finally:
        rp.set_dtw_variant(0)
        rp.set_avg_gate(-1)
    first = maxf + 2
    worst = 0.0
    for b in range(3):
        tr = O.trace_window_scores(O.default_config(**cfg), rpw, audio[b], T)
        want = np.concatenate([tr[:, :1], tr[:, 2:]], axis=1)
        for variant, tol in ((1, 5e-6), (0, SCORE_RTOL)):
            got = res[variant][b, first:]
            assert got.shape == want.shape
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-12)
            assert rel.max() < tol, (d, band, b, variant, rel.max(), np.unravel_index(rel.argmax(), rel.shape))
            worst = max(worst, rel.max()) if variant == 0 else worst
    if band >= 2:
        assert (res[0][:, first:] > 0).any()
    # the whole detector on the same input: detections equal the oracle's
    total, counts, want_d = O.run_streams(O.default_config(**cfg), [rpw], audio, n_threads=3, max_det=4)
    bt = rp.RustpotterBatch(3, rp.default_config(**cfg))
    bt.add_wakeword_from_buffer("w", rpw)
    got = bt.process(audio)
    assert len(got) == int(counts.sum()) and bt.windows_scored() == total
    for (s, c, dd), w in zip(got, [w for b in range(3) for w in want_d[b]]):
        assert dd["counter"] == w["counter"] and _rel(dd["score"], w["score"]) < SCORE_RTOL


def test_reference_fixture_runs_on_the_tuned_kernel():
    """The reference's own fixtures are mfcc_size 5 (tests/wakeword.rs:6-24): the batched front-end scores them with
    the tuned kernel (zero-padded to 16), same detections as the per-stream goldens (tests/detector.rs:25-38)."""
    from tests.helpers import two_wakeword_stream
    x = np.frombuffer(two_wakeword_stream(), dtype="<i2").copy()
    x = x[: x.size // 480 * 480]
    pcm = np.stack([x, x, x])
    bt = rp.RustpotterBatch(3, rp.default_config(score_mode="max"))
    bt.add_wakeword_from_file("wakeword", golden("oye_casa_g.rpw"))
    got = bt.process(pcm)
    assert len(got) == 6
    for i, (s, c, d) in enumerate(got):
        avg, sc = [(0.6495044, 0.7310586), (0.5804737, 0.721843)][i % 2]
        assert _rel(d["score"], sc) < SCORE_RTOL and _rel(d["avg_score"], avg) < SCORE_RTOL, d
    tiles, passed = bt.last_gate_stats()
    assert tiles > 0 and 0 < passed <= tiles     # the avg gate ran, i.e. the tuned kernel took the mfcc_size-5 fixture


def test_multi_device_handle_matches_single_device():
    """rp_batch_create_multi: streams sharded contiguously over devices (here the same device twice when the box has one
    GPU), host audio in, detections merged in stream order — identical to the single-device handle."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = [0, 1 % n_dev, 0] if n_dev >= 1 else [0]
    rpws, audio = _config5_case(B=20)
    cfg = rp.default_config(score_mode="median")
    one = rp.RustpotterBatch(20, cfg)
    multi = rp.RustpotterBatch(20, cfg, devices=devices)
    assert multi.n_devices() == 3
    for i, r in enumerate(rpws):
        one.add_wakeword_from_buffer(f"w{i}", r)
        multi.add_wakeword_from_buffer(f"w{i}", r)
    a, b = one.process(audio), multi.process(audio)
    assert len(a) >= 4 and [(s, c, d["name"], d["counter"], float(d["score"])) for s, c, d in a] == \
        [(s, c, d["name"], d["counter"], float(d["score"])) for s, c, d in b]
    assert one.windows_scored() == multi.windows_scored()
    # second call continues every shard's state
    more = synth_audio(20, 30 * 480, seed=77)
    assert [(s, c, d["counter"]) for s, c, d in one.process(more)] == [(s, c, d["counter"]) for s, c, d in multi.process(more)]
    with pytest.raises(rp.RustpotterError):
        multi.process(torch.from_numpy(audio).cuda())      # device audio cannot be split by the library
    assert multi.remove_wakeword("w1") is True and one.remove_wakeword("w1") is True
    assert multi.max_mfcc_frames() == one.max_mfcc_frames()


# ------------------------------------------------------------------ 48 kHz ingest
@pytest.mark.parametrize("case", [0, 1], ids=["record_with_noise", "record_with_noise_using_filters"])
def test_golden_48khz_record(case):
    """tests/detector.rs:162-214 through the C ABI: a 48 kHz recording, resampled by the restated rubato FftFixedInOut on the
    host, then the CUDA path. Counters exact, scores within the parity bar (observed ~1e-6)."""
    from tests.helpers import REAL_SAMPLE_GOLDENS, real_sample_stream
    kw, want = REAL_SAMPLE_GOLDENS[case]
    rate, x = real_sample_stream()
    det = rp.Rustpotter(rp.default_config(sample_rate=rate, sample_format="f32", channels=1, **kw))
    det.add_wakeword_from_file("wakeword", golden("oye_casa_real.rpw"))
    n = det.get_samples_per_frame()
    assert n == 1440
    got = [d for d in (det.process_samples(x[i:i + n]) for i in range(0, len(x) - n + 1, n)) if d is not None]
    assert len(got) == len(want)
    for d, (avg, score, counter) in zip(got, want):
        assert d["counter"] == counter
        assert _rel(d["avg_score"], avg) < SCORE_RTOL and _rel(d["score"], score) < SCORE_RTOL, (d, avg, score)


def test_batch_48khz_streams_match_the_per_stream_handle():
    """The batched front-end with sample_rate 48000: every stream through its own host resampler, then one device call."""
    from tests.helpers import REAL_SAMPLE_GOLDENS, real_sample_stream
    kw, want = REAL_SAMPLE_GOLDENS[0]
    rate, x = real_sample_stream()
    x = x[: x.size // 1440 * 1440]
    shifted = np.concatenate([np.zeros(1440 * 7, np.float32), x[:-1440 * 7]])
    audio = np.stack([x, shifted, (x * np.float32(0.5)).astype(np.float32)])
    bt = rp.RustpotterBatch(3, rp.default_config(sample_rate=rate, **kw))
    assert bt.get_samples_per_frame() == 1440
    bt.add_wakeword_from_file("wakeword", golden("oye_casa_real.rpw"))
    half = (x.size // 1440 // 2) * 1440
    got = bt.process(audio[:, :half]) + [(s, c + half // 1440, d) for s, c, d in bt.process(audio[:, half:])]
    per = {s: [d for s2, c, d in got if s2 == s] for s in range(3)}
    assert [d["counter"] for d in per[0]] == [w[2] for w in want]
    for d, (avg, score, counter) in zip(per[0], want):
        assert _rel(d["score"], score) < SCORE_RTOL and _rel(d["avg_score"], avg) < SCORE_RTOL
    assert len(per[1]) >= 2 and len(per[2]) == 3
    for a, b in zip(per[0], per[2]):   # a pure gain change leaves the cosine/CMN scores (nearly) unchanged
        assert a["counter"] == b["counter"] and _rel(a["score"], b["score"]) < 1e-3
    with pytest.raises(rp.RustpotterError):
        import torch
        bt.process(torch.zeros((3, 1440), device="cuda"))


# ------------------------------------------------------------------ short calls: the cadence kernel (30 ms chunks)
@pytest.mark.parametrize("d,band,chunks_per_call", [(16, 5, 1), (16, 5, 2), (16, 5, 8), (5, 5, 1), (16, 3, 1), (13, 1, 3), (16, 5, 9)])
def test_cadence_kernel_dense_scores_vs_generic_and_pipeline(d, band, chunks_per_call):
    """Calls of a few 30 ms chunks take the warp-per-window-triple kernel (variant 0); its dense per-window scores equal the
    reference-order generic kernel's (variant 1) and the pipeline kernel's (variant 9) on the same calls, call by call."""
    lengths = (50, 44, 57, 50)
    rpw, utts = make_wakeword(O, d=d, lengths=lengths, seed=500 + d + band)
    n_calls = 70 // chunks_per_call
    S = chunks_per_call * 480
    audio = synth_audio(5, n_calls * S, seed=31)
    splice(audio[1], utts[2], 70)
    splice(audio[4], utts[0], 100)
    res = {}
    rp.set_avg_gate(0)
    try:
        for variant in (1, 9, 0):
            rp.set_dtw_variant(variant)
            bt = rp.RustpotterBatch(5, rp.default_config(band_size=band))
            bt.add_wakeword_from_buffer("w", rpw)
            out, dets = [], []
            for c in range(n_calls):
                dets += [(s, c * chunks_per_call + ch, dd["counter"], float(dd["score"])) for s, ch, dd in bt.process(audio[:, c * S:(c + 1) * S])]
                out.append(bt.last_scores(chunks_per_call * 3, len(lengths) + 1).copy())
            res[variant] = (np.concatenate(out, axis=1), dets, bt.windows_scored())
    finally:
        rp.set_dtw_variant(0)
        rp.set_avg_gate(-1)
    first = max(lengths) + 2
    ref, got9, got0 = (res[v][0][:, first:] for v in (1, 9, 0))
    assert np.isfinite(ref).all() and np.isfinite(got0).all()
    for got in (got9, got0):
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12)
        assert rel.max() < 3e-5, (d, band, chunks_per_call, rel.max())
    # same detections (stream, chunk, counter) from all three, scores within the parity bar
    assert res[0][2] == res[1][2] == res[9][2]
    assert [x[:3] for x in res[0][1]] == [x[:3] for x in res[1][1]] == [x[:3] for x in res[9][1]]
    assert len(res[0][1]) >= 1 or band < 2      # (band 1 leaves the result cell outside the band: every score is 0)
    for x, y in zip(res[0][1], res[1][1]):
        assert _rel(x[3], y[3]) < SCORE_RTOL
