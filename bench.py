#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wakeword-scoring path.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...           the reference algorithm on the host cores (CPU oracle)

Workload (default, every N): BASELINE.json configs[4] — the configuration the metric "windows scored/sec at 1/2/4/8
B200" is quoted on: 65 536 streams x 10.02 s of synthetic 16 kHz f32 mono, 4 WakewordRefs (8 templates + avg_features
each, D = 16, ~1 s), ScoreMode::Median, sharded contiguously over the N GPUs (STRONG scaling: total work is fixed; no
collective on the data path, torch.distributed only for the barrier and the max-over-ranks time). `--config 2` runs
BASELINE configs[1] (4096 streams per GPU, 1 WakewordRef, Max, weak scaling) — also reported as the `config2` object of
the default N=1 line.

A "step" is one pass of the hot path over the whole batch, starting from freshly reset stream state, so it scores
exactly the windows that many fresh `Rustpotter`s would. `value` = windows scored/s with the f32 audio
resident in HBM (CUDA events on the launching stream); `e2e` = the same through the public batched C ABI with the f32
audio in pinned HOST memory, H2D copy and D2H of the detections inside the timed region; `e2e_i16` = the same with an
i16 source (Sample::into_f32 on the device: half the H2D bytes). Extra objects: `roofline` (the DTW pairs kernel on
BASELINE configs[3]), `mfcc_microbench` (configs[2]: 10 M frames), `cadence` (30 ms / 300 ms calls), `cpu_baseline`
(the oracle timed on this box's host cores, with the in-run parity check of the GPU detections against it), `clocks`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "audio windows scored/sec (batched streams)"   # identical in both arms (BASELINE.json `metric`, first clause)
N_CHUNKS = 334                      # 334 * 480 = 160320 samples = 10.02 s (a whole number of 30 ms chunks)
SAMPLES = N_CHUNKS * 480
MFCC_SIZE = 16
SPLICE_EVERY = 50                   # every 50th stream contains an utterance one of the templates was made from
SEED = 0x5EED
GEN_ROWS = 256                      # streams per RNG block: stream b's audio depends only on (SEED, b // GEN_ROWS)

CONFIGS = {
    5: dict(name="BASELINE configs[4]", total_streams=65536, n_wakewords=4, score_mode="median", scaling="strong"),
    2: dict(name="BASELINE configs[1]", streams_per_gpu=4096, n_wakewords=1, score_mode="max", scaling="weak"),
}


# ---------------------------------------------------------------------------------------------- workload
def cbor_rpw(name: str, templates, avg, rms_level: float) -> bytes:
    """Serialises a WakewordRef as the reference does (serde struct -> CBOR; wakeword_ref.rs:12-20)."""
    import struct

    def head(major, v):
        if v < 24:
            return bytes([major << 5 | v])
        if v < 256:
            return bytes([major << 5 | 24, v])
        if v < 65536:
            return bytes([major << 5 | 25]) + struct.pack(">H", v)
        return bytes([major << 5 | 26]) + struct.pack(">I", v)

    def text(s):
        b = s.encode()
        return head(3, len(b)) + b

    def matrix(m):
        m = np.asarray(m, np.float32)
        out = [head(4, m.shape[0])]
        row_head = head(4, m.shape[1])
        for r in m:
            out.append(row_head + b"".join(b"\xfa" + struct.pack(">f", float(x)) for x in r))
        return b"".join(out)

    out = [head(5, 7), text("name"), text(name), text("avg_features"), matrix(avg) if avg is not None else b"\xf6",
           text("samples_features"), head(5, len(templates))]
    for n, m in templates:
        out += [text(n), matrix(m)]
    out += [text("threshold"), b"\xf6", text("avg_threshold"), b"\xf6", text("rms_level"),
            b"\xfa" + struct.pack(">f", rms_level), text("mfcc_size"), head(0, int(np.asarray(templates[0][1]).shape[1]))]
    return b"".join(out)


def load_wakewords(n_wakewords: int):
    """The benchmark's WakewordRefs: template matrices from the committed fixture (tests/golden/bench_templates.npz,
    made by tools/make_bench_templates.py), utterances regenerated from their seeds. Identical in both arms."""
    from tests.helpers import CONFIG5_LENGTHS, CONFIG5_NAMES, CONFIG5_SEEDS, wakeword_utterances
    z = np.load(os.path.join(ROOT, "tests", "golden", "bench_templates.npz"))
    rpws, utts = [], []
    for w in range(n_wakewords):
        tmpl = [(f"sample_{i}.wav", z[f"w{w}_t{i}"]) for i in range(len(CONFIG5_LENGTHS[w]))]
        avg = max(tmpl, key=lambda t: t[1].shape[0])[1].copy()
        rpws.append(cbor_rpw(CONFIG5_NAMES[w], tmpl, avg, 0.05))
        utts.append(wakeword_utterances(CONFIG5_LENGTHS[w], CONFIG5_SEEDS[w]))
    return rpws, utts


def synth_streams(torch, lo: int, hi: int, device, utts):
    """Streams [lo, hi) of the synthetic workload: 0.1*N(0,1) noise + a per-stream chirp 200->3000 Hz at
    amplitude 0.3, clipped, never 0; every 50th stream holds the utterance of one template. Stream b depends only on
    (SEED, b), so any arm / rank generates the same samples for the same stream on the same kind of device."""
    assert lo % GEN_ROWS == 0
    t = torch.arange(SAMPLES, device=device, dtype=torch.float64) / 16000.0
    dur = SAMPLES / 16000.0
    out = torch.empty((hi - lo, SAMPLES), dtype=torch.float32, device=device)
    for b0 in range(lo, hi, GEN_ROWS):
        nb = min(GEN_ROWS, hi - b0)
        g = torch.Generator(device=device).manual_seed(SEED + b0 // GEN_ROWS)
        ids = torch.arange(b0, b0 + nb, device=device, dtype=torch.float64)[:, None]
        f0 = 200.0 + 37.0 * (ids % 13)
        k = (3000.0 - f0) / dur
        chirp = 0.3 * torch.sin(2 * np.pi * (f0 * t + 0.5 * k * t * t) + 0.1 * ids)
        noise = 0.1 * torch.randn((GEN_ROWS, SAMPLES), generator=g, device=device, dtype=torch.float32)[:nb]
        x = torch.clamp(noise + chirp.float(), -1.0, 1.0)
        x[x == 0] = 1e-4
        out[b0 - lo:b0 - lo + nb] = x
    first = (lo + SPLICE_EVERY - 1) // SPLICE_EVERY * SPLICE_EVERY
    for b in range(first, hi, SPLICE_EVERY):
        w = (b // SPLICE_EVERY) % len(utts)
        u = utts[w][(b // (SPLICE_EVERY * len(utts))) % len(utts[w])]
        hop = 150 + (7 * b) % 600
        out[b - lo, hop * 160: hop * 160 + u.size] = torch.from_numpy(u).to(device)
    return out


def workload_text(cfg_id: int, n_streams_total: int, world: int) -> str:
    c = CONFIGS[cfg_id]
    refs = "4 WakewordRefs (8 templates 84..100 frames + avg_features each, D=16)" if c["n_wakewords"] == 4 else \
        "1 WakewordRef (8 templates 88..100 frames + avg_features, D=16)"
    return (f"{c['name']}: {n_streams_total} streams x {SAMPLES} samples (10.02 s) 16 kHz f32 mono over {world} GPU(s), {refs}, "
            f"band 5, score_ref 0.22, avg_threshold 0.2, threshold 0.5, ScoreMode {c['score_mode']}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []   # (host time when the line arrived, text)
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        inside = [ln for (t, ln) in self.lines if self.t0 is None or (self.t0 - 0.06 <= t <= (self.t1 or t) + 0.06)]
        for ln in inside:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def pin_to_gpu_numa_node(torch, local: int, world: int) -> dict:
    """Best effort: run this rank (and first-touch its pinned staging buffer) on the cores of the GPU's NUMA node, split
    between the ranks that share the node. Returns what was done (reported on the JSON line)."""
    info = {"numa_node": None, "cpus": None}
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = []
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if cpus and world > 1:
            per = max(1, len(cpus) // world)
            mine = cpus[(local * per) % len(cpus):][:per] or cpus
            os.sched_setaffinity(0, mine)
            cpus = mine
        info = {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001 — topology files are optional
        info["note"] = f"{type(e).__name__}"
    return info


# ---------------------------------------------------------------------------------------------- micro-benchmarks (rank 0)
def dtw_roofline(torch, rp):
    """BASELINE configs[3]: 1M independent pairs, template 120x16 vs window 100x16, band 5 (effective
    window 20), distinct data per pair in HBM (14.08 GB > L2). Algorithmic bytes/pair = (m+n)*D*4 + 4."""
    P, m, n, d = 1_000_000, 120, 100, 16
    g = torch.Generator(device="cuda").manual_seed(1234)
    scale = torch.tensor([8, 4, 3, 2, 2, 1.5] + [1.0] * 10, device="cuda")
    a = torch.randn((P, m, d), device="cuda", generator=g) * scale
    w = torch.randn((P, n, d), device="cuda", generator=g) * scale
    out = torch.empty(P, device="cuda")
    for _ in range(3):
        rp.dtw_scores(a, w, band=5, out=out)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        rp.dtw_scores(a, w, band=5, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_per_launch = P * ((m + n) * d * 4 + 4)
    peak, how = measured_peaks()
    ach = bytes_per_launch / (ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "dtw_stream_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    assert bool(torch.isfinite(out).all())
    del a, w
    return {"bound": "hbm", "kernel": "dtw pairs kernel (rp_dtw_scores)", "workload": "1M pairs 120x16 vs 100x16, band 5",
            "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
            "traffic_source": traffic_src, "ms_per_launch": round(ms, 4), "bytes_per_launch": bytes_per_launch,
            "peak_source": how, "launches": reps + 3}


def mfcc_microbench(torch, rp):
    """BASELINE configs[2]: 10 M 10 ms frames @16 kHz f32, D = 16, through the fused MFCC kernel (rp_mfcc_frames).
    Algorithmic bytes per frame: 640 B in (160 new samples) + 64 B out = 704 B."""
    B, hops = 4000, 2503                      # 4000 streams x 2500 frames = 10.0 M frames; 6.4 GB of audio (> L2)
    g = torch.Generator(device="cuda").manual_seed(99)
    audio = torch.empty((B, hops * 160), device="cuda")
    for b0 in range(0, B, 500):
        audio[b0:b0 + 500] = 0.2 * torch.randn((500, hops * 160), generator=g, device="cuda")
    frames = B * (hops - 3)
    for _ in range(2):
        out = rp.mfcc_frames(audio, MFCC_SIZE)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = rp.mfcc_frames(audio, MFCC_SIZE)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak, how = measured_peaks()
    gbs = frames * 704 / (ms * 1e-3) / 1e9
    assert bool(torch.isfinite(out).all())
    del audio, out
    return {"workload": f"{frames} frames ({B} streams x {hops - 3}), D=16", "ms_per_launch": round(ms, 3),
            "frames_per_s": round(frames / (ms * 1e-3), 1), "algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4),
            "bytes_per_frame": 704, "note": "issue/FP32-bound (about 1.5 k warp instructions per frame), see profiles/"}


def cadence_bench(torch, rp, rpws, score_mode: str, device_index: int):
    """The reference's calling cadence (detector.rs:347-376: one 30 ms chunk per call): 4096 streams fed chunk by chunk
    (S = 480) and in 300 ms calls (S = 4800) from pinned host memory; per-call latency and windows/s in steady state."""
    out = {}
    n = 4096
    bt = rp.RustpotterBatch(n, rp.default_config(score_mode=score_mode), device=device_index)
    for i, r in enumerate(rpws):
        bt.add_wakeword_from_buffer(f"w{i}", r)
    for S, calls in ((480, 150), (4800, 30)):
        g = torch.Generator(device="cpu").manual_seed(7)
        host = (0.2 * torch.randn((n, S), generator=g)).pin_memory()
        bt.reset()
        warm = (bt.max_mfcc_frames() + 3) * 160 // S + 3
        for _ in range(warm):
            bt.process(host)
        torch.cuda.synchronize()
        w0 = bt.windows_scored()
        lat = []
        stage = {}
        t0 = time.perf_counter()
        for _ in range(calls):
            t1 = time.perf_counter()
            bt.process_count(host)
            lat.append((time.perf_counter() - t1) * 1e3)
            for k, v in bt.last_timings().items():
                stage[k] = stage.get(k, 0.0) + v / calls
        dt = time.perf_counter() - t0
        w = bt.windows_scored() - w0
        out[f"S{S}"] = {"streams": n, "calls": calls, "windows_per_s": round(w / dt, 1), "ms_per_call_median": round(float(np.median(lat)), 3),
                        "ms_per_call_p95": round(float(np.percentile(lat, 95)), 3), "launches_per_call": bt.last_launches(),
                        "audio_ms_per_call": S / 16.0, "real_time_factor": round(S / 16.0 / float(np.median(lat)), 2),
                        "stage_ms_per_call": {k: round(v, 3) for k, v in stage.items()}}
    del bt
    # the per-stream drop-in (`Rustpotter::process_samples`, one 30 ms chunk per call, host f32 in, detection out)
    det = rp.Rustpotter(rp.default_config(score_mode=score_mode, sample_format="f32"), device=device_index)
    for i, r in enumerate(rpws):
        det.add_wakeword_from_buffer(f"w{i}", r)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = (0.2 * torch.randn(480 * 400, generator=g)).numpy()
    for c in range(150):
        det.process_samples(x[c * 480:(c + 1) * 480])
    lat = []
    w0 = det.windows_scored()
    t0 = time.perf_counter()
    for c in range(150, 400):
        t1 = time.perf_counter()
        det.process_samples(x[c * 480:(c + 1) * 480])
        lat.append((time.perf_counter() - t1) * 1e3)
    dt = time.perf_counter() - t0
    out["single_stream_handle"] = {"calls": len(lat), "ms_per_call_median": round(float(np.median(lat)), 4),
                                   "ms_per_call_p95": round(float(np.percentile(lat, 95)), 4), "audio_ms_per_call": 30.0,
                                   "real_time_factor": round(30.0 / float(np.median(lat)), 1),
                                   "windows_per_s": round((det.windows_scored() - w0) / dt, 1)}
    del det
    return out


def cpu_baseline(rpws, score_mode: str, audio_host: np.ndarray, gpu_dets, target_seconds: float = 12.0):
    """The oracle (C++ restatement of the reference algorithm) on this box's host cores, on a bounded sample of the same
    streams — and the in-run parity check: the GPU detections of those streams against the oracle's."""
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    cfg = O.default_config(score_mode=score_mode)
    probe = min(audio_host.shape[0], threads)
    O.run_streams(cfg, rpws, audio_host[:probe], n_threads=threads, native=True)   # warm-up (library load, tables, threads)
    t0 = time.perf_counter()
    O.run_streams(cfg, rpws, audio_host[:probe], n_threads=threads, native=True)
    dt = time.perf_counter() - t0
    n = int(min(audio_host.shape[0], max(probe, probe * target_seconds / max(dt, 1e-3))))
    n = max(threads, n // threads * threads)
    n = min(n, audio_host.shape[0])
    t0 = time.perf_counter()
    w, counts, want = O.run_streams(cfg, rpws, audio_host[:n], n_threads=threads, max_det=16, native=True)
    dt = time.perf_counter() - t0
    # parity: same detections (stream, name, counter) and scores within 1e-4 relative (north_star)
    got = {}
    for s, c, d in gpu_dets:
        if s < n:
            got.setdefault(s, []).append(d)
    n_det, worst, mismatches = 0, 0.0, 0
    for b in range(n):
        g = got.get(b, [])
        if len(g) != len(want[b]):
            mismatches += 1
            continue
        for x, y in zip(g, want[b]):
            n_det += 1
            if x["name"] != y["name"] or x["counter"] != y["counter"]:
                mismatches += 1
                continue
            for a, bb in [(x["score"], y["score"]), (x["avg_score"], y["avg_score"])] + [(x["scores"][k], y["scores"][k]) for k in y["scores"]]:
                worst = max(worst, abs(float(a) - float(bb)) / max(abs(float(bb)), 1e-12))
    parity = {"streams": n, "detections": n_det, "mismatched_streams_or_detections": mismatches, "max_rel_score_err": worst,
              "tolerance": 1e-4, "ok": bool(mismatches == 0 and worst <= 1e-4)}
    return {"value": round(w / dt, 1), "unit": "windows/s", "cores": threads, "kind": "port",
            "sample": f"{n} of this workload's streams ({w} windows, {dt:.1f} s), {threads} host threads, "
                      "C++ oracle built -O3 -march=native (the Rust reference cannot be built in this image)",
            "parity": parity}


def dist_setup():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _stdout_to_stderr() -> int:
    """Everything but the JSON line goes to stderr: NCCL prints its version banner to stdout when the first
    communicator is created, and the contract is ONE line on stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _restore_stdout(saved: int) -> None:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


# ---------------------------------------------------------------------------------------------- our arm
def run_legs(torch, rp, bt, audio_dev, args, barrier, device, want_i16=True, want_dense=True):
    """Resident, dense, end-to-end f32 and end-to-end i16 legs for one batch handle. Returns per-rank numbers."""
    res = {}

    def one_step(src, parse=False):
        """One pass over the batch from reset state. The timed loops take the detection COUNT from the C ABI
        (rp_batch_detection structs are filled by the call either way); parse=True also builds the Python dicts."""
        bt.reset()
        w0 = bt.windows_scored()
        dets = bt.process(src) if parse else bt.process_count(src)
        return bt.windows_scored() - w0, dets

    # ---- resident leg (value): f32 audio already in HBM, CUDA events on the launching stream
    for _ in range(args.warmup):
        one_step(audio_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    windows = launches = n_dets = 0
    stage = {}
    e0.record()
    for _ in range(args.steps):
        w, nd = one_step(audio_dev)
        windows += w
        n_dets += nd
        launches += bt.last_launches()
        for k, v in bt.last_timings().items():
            stage[k] = stage.get(k, 0.0) + v
    e1.record()
    barrier()
    res.update(ms_res=e0.elapsed_time(e1), windows=windows, launches=launches, dets_per_step=n_dets / max(args.steps, 1),
               stage={k: v / args.steps for k, v in stage.items()}, gate=bt.last_gate_stats())
    # ---- dense leg: every template of every window (the avg gate off), fewer steps
    if want_dense:
        rp.set_avg_gate(0)
        ds = max(1, min(args.steps, 2))
        one_step(audio_dev)
        barrier()
        e0.record()
        wd = 0
        for _ in range(ds):
            w, nd_dense = one_step(audio_dev)
            wd += w
        e1.record()
        barrier()
        rp.set_avg_gate(-1)
        res.update(ms_dense=e0.elapsed_time(e1), windows_dense=wd, dense_steps=ds, dets_dense=nd_dense, dets_gated=nd)
    # ---- end-to-end legs: pinned host audio in, detections out, through the public API
    host = torch.empty(audio_dev.shape, dtype=torch.float32, pin_memory=True)   # first touch on this rank's cores
    host.copy_(audio_dev)
    torch.cuda.synchronize()
    for _ in range(min(args.warmup, 2)):
        one_step(host)
    barrier()
    t0 = time.perf_counter()
    we = 0
    h2d_ms = 0.0
    for _ in range(args.steps):
        w, _nd = one_step(host)
        we += w
        h2d_ms += bt.last_timings().get("h2d_ms", 0.0)
    torch.cuda.synchronize()
    res.update(ms_e2e=(time.perf_counter() - t0) * 1e3, windows_e2e=we, h2d_ms_per_step=h2d_ms / max(args.steps, 1),
               h2d_bytes=int(host.numel()) * 4)
    barrier()
    res["dets_e2e"] = one_step(host, parse=True)[1]   # (untimed) the detections themselves, for the in-run parity check
    if want_i16:
        B, S = audio_dev.shape
        if B >= 8192:   # reuse the second half of the f32 pinned buffer (only its first streams are needed afterwards)
            host16 = host.view(-1).view(torch.int16)[B * S:].view(B, S)
        else:
            host16 = torch.empty((B, S), dtype=torch.int16, pin_memory=True)
        for b0 in range(0, B, 1024):   # (chunked: no batch-sized temporaries on the device)
            host16[b0:b0 + 1024].copy_(torch.clamp(torch.round(audio_dev[b0:b0 + 1024] * 32767.0), -32768, 32767).to(torch.int16))
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        for _ in range(min(args.warmup, 2)):
            one_step(host16)
        barrier()
        t0 = time.perf_counter()
        w16 = 0
        steps16 = max(1, min(args.steps, 5))   # (a secondary leg: fewer steps keep the default run within minutes)
        for _ in range(steps16):
            w, _d = one_step(host16)
            w16 += w
        torch.cuda.synchronize()
        res.update(ms_e2e_i16=(time.perf_counter() - t0) * 1e3, windows_e2e_i16=w16, h2d_bytes_i16=int(host16.numel()) * 2, steps_i16=steps16)
        barrier()
        del host16
    res["host"] = host
    return res


def run_ours(args):
    import torch
    saved_stdout = _stdout_to_stderr()

    import rustpotter_b200 as rp
    from rustpotter_b200.sharding import shard_range
    rank, world, local = dist_setup()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(torch, local, world)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    cfg_id = args.config
    c = CONFIGS[cfg_id]
    if c["scaling"] == "strong":
        total = args.streams or c["total_streams"]
        lo, hi = shard_range(total, rank, world)
    else:
        per = args.streams or c["streams_per_gpu"]
        total = per * world
        lo, hi = rank * per, (rank + 1) * per
    if args.dtw_variant:
        rp.set_dtw_variant(args.dtw_variant)
    rpws, utts = load_wakewords(c["n_wakewords"])
    audio = synth_streams(torch, lo, hi, device, utts)
    torch.cuda.synchronize()

    def make_batch(n, n_ww, mode):
        b = rp.RustpotterBatch(n, rp.default_config(score_mode=mode), device=local)
        for i in range(n_ww):
            b.add_wakeword_from_buffer(f"w{i}", rpws[i])
        b.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
        return b

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    bt = make_batch(hi - lo, c["n_wakewords"], c["score_mode"])
    sampler = ClockSampler(local)
    sampler.start()
    sampler.mark_begin()
    r = run_legs(torch, rp, bt, audio, args, barrier, device)
    sampler.mark_end()
    clocks = sampler.stop()
    host = r.pop("host")

    keys = ["ms_res", "ms_dense", "ms_e2e", "ms_e2e_i16", "h2d_ms_per_step"]
    t_max = torch.tensor([r.get(k, 0.0) for k in keys], dtype=torch.float64, device=device)
    sums = torch.tensor([r["windows"], r.get("windows_dense", 0), r["windows_e2e"], r.get("windows_e2e_i16", 0), r["launches"],
                         r["gate"][0], r["gate"][1], r["dets_per_step"]], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_res, ms_dense, ms_e2e, ms_e2e_i16, h2d_ms = (float(x) for x in t_max)
    windows, windows_dense, windows_e2e, windows_i16, launches, tiles, passed, dets_step = (float(x) for x in sums)

    if rank == 0:
        steps = args.steps
        h2d_bytes_all = total * SAMPLES * 4
        line = {
            "metric": METRIC,
            "value": round(windows / (ms_res * 1e-3), 1),
            "unit": "windows/s",
            "n_gpus": world,
            "steps": steps,
            "warmup": args.warmup,
            "ms_per_step": round(ms_res / steps, 3),
            "higher_is_better": True,
            "scaling": c["scaling"],
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_text(cfg_id, total, world), "streams_total": total, "streams_rank0": hi - lo,
                       "samples_per_stream": SAMPLES, "windows_per_step_all_gpus": windows / steps,
                       "l2": "inputs (641 KB of audio per stream, GBs per GPU) exceed the 126 MB L2; no flush needed",
                       "sharding": f"streams sharded contiguously over {world} GPU(s) (rustpotter_b200.sharding.shard_range), "
                                   "no collective on the data path",
                       "roofline_metric": "DTW HBM GB/s vs peak: see `roofline`"},
            "e2e": {"value": round(windows_e2e / (ms_e2e * 1e-3), 1), "unit": "windows/s", "h2d_bytes_per_step": h2d_bytes_all,
                    "d2h_bytes_per_step": int(world * 4 + dets_step * 64), "ms_per_step": round(ms_e2e / steps, 3),
                    "source": "f32 pinned host audio", "h2d_ms_per_step_slowest_rank": round(h2d_ms, 3),
                    "h2d_GBps_per_gpu": round(h2d_bytes_all / world / max(h2d_ms, 1e-6) / 1e6, 2)},
            "e2e_i16": {"value": round(windows_i16 / (ms_e2e_i16 * 1e-3), 1), "unit": "windows/s", "h2d_bytes_per_step": h2d_bytes_all // 2,
                        "ms_per_step": round(ms_e2e_i16 / r["steps_i16"], 3), "steps": r["steps_i16"],
                        "source": "i16 pinned host audio, Sample::into_f32 on the device"},
            "value_dense": round(windows_dense / (ms_dense * 1e-3), 1) if ms_dense > 0 else None,
            "avg_gate": {"tiles": int(tiles), "passed": int(passed), "pass_fraction": round(passed / tiles, 4) if tiles else None,
                         "tile": "128 consecutive windows of one stream x one wakeword",
                         "detections_gated_vs_dense_rank0": [r.get("dets_gated"), r.get("dets_dense")]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host": {"numa": numa, "cpus_online": os.cpu_count()},
            "stage_ms_per_step_rank0": {k: round(v, 3) for k, v in r["stage"].items()},
            "detections_per_step_all_gpus": dets_step,
        }
        gpu_dets = r["dets_e2e"]
        del audio, bt
        torch.cuda.empty_cache()
        if world == 1 and not args.no_extras:
            if cfg_id == 5:   # BASELINE configs[1] on the same box, for continuity with round 1
                a2 = argparse.Namespace(**vars(args))
                a2.steps, a2.warmup = max(3, min(args.steps, 10)), 3
                u1 = [utts[0]]
                audio2 = synth_streams(torch, 0, 4096, device, u1)
                b2 = make_batch(4096, 1, "max")
                r2 = run_legs(torch, rp, b2, audio2, a2, barrier, device, want_i16=True, want_dense=True)
                r2.pop("host")
                line["config2"] = {"workload": workload_text(2, 4096, 1), "steps": a2.steps,
                                   "value": round(r2["windows"] / (r2["ms_res"] * 1e-3), 1),
                                   "value_dense": round(r2["windows_dense"] / (r2["ms_dense"] * 1e-3), 1),
                                   "ms_per_step": round(r2["ms_res"] / a2.steps, 3),
                                   "e2e": round(r2["windows_e2e"] / (r2["ms_e2e"] * 1e-3), 1),
                                   "e2e_i16": round(r2["windows_e2e_i16"] / (r2["ms_e2e_i16"] * 1e-3), 1),
                                   "dense_steps": r2["dense_steps"], "e2e_i16_steps": r2["steps_i16"],
                                   "stage_ms_per_step": {k: round(v, 3) for k, v in r2["stage"].items()},
                                   "avg_gate_pass_fraction": round(r2["gate"][1] / r2["gate"][0], 4) if r2["gate"][0] else None}
                del audio2, b2
                torch.cuda.empty_cache()
            line["mfcc_microbench"] = mfcc_microbench(torch, rp)
            line["cadence"] = cadence_bench(torch, rp, rpws[:c["n_wakewords"]], c["score_mode"], local)
        if not args.no_roofline:
            line["roofline"] = dtw_roofline(torch, rp)
        if world == 1 and not args.no_cpu:
            n_cpu = min(hi - lo, 2048)
            line["cpu_baseline"] = cpu_baseline(rpws[:c["n_wakewords"]], c["score_mode"], host[:n_cpu].numpy(), gpu_dets)
        _restore_stdout(saved_stdout)
        print(json.dumps(line), flush=True)
        saved_stdout = _stdout_to_stderr()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """--impl reference: the reference algorithm (CPU oracle, all host threads) on the same config and inputs.
    Each step is a bounded sample of the workload's streams (its first streams)."""
    rank, world, local = dist_setup()
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    c = CONFIGS[args.config]
    total = (args.streams or c["total_streams"]) if c["scaling"] == "strong" else (args.streams or c["streams_per_gpu"]) * world
    rpws, utts = load_wakewords(c["n_wakewords"])
    n = max(threads, min(total, threads * args.ref_streams_per_thread))
    n = min(total, (n + GEN_ROWS - 1) // GEN_ROWS * GEN_ROWS) if n > GEN_ROWS else n
    # the same streams the GPU arm scores (same generator, same seeds) when a CUDA device is there to generate them
    dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    audio = synth_streams(torch, 0, max(n, min(GEN_ROWS, total)), dev, utts)[:n].cpu().numpy()
    cfg = O.default_config(score_mode=c["score_mode"])
    for _ in range(min(args.warmup, 1)):
        O.run_streams(cfg, rpws, audio[:threads], n_threads=threads, native=True)
    t0 = time.perf_counter()
    windows = 0
    for _ in range(args.steps):
        w, _, _ = O.run_streams(cfg, rpws, audio, n_threads=threads, native=True)
        windows += w
    dt = time.perf_counter() - t0
    val = round(windows / dt, 1)
    sample = (f"the first {n} streams x {SAMPLES} samples per step ({windows // max(args.steps, 1)} windows/step), {threads} host threads, "
              f"audio generated on {dev.type}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / args.steps, 3),
        "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.config, total, world) + f" — bounded sample: {sample}"},
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": threads, "kind": "port", "sample": sample +
                         "; C++ oracle (restatement of the reference algorithm; the Rust crate cannot be built in this image)"},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 5], help="BASELINE config: 5 = configs[4] (default), 2 = configs[1]")
    ap.add_argument("--streams", type=int, default=0, help="total streams (config 5) / streams per GPU (config 2); 0 = the config's own")
    ap.add_argument("--ref-streams-per-thread", type=int, default=2)
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-2, MFCC and cadence legs")
    ap.add_argument("--dtw-variant", type=int, default=0, help="rp_set_dtw_variant for A/B measurements (default 0 = automatic)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
