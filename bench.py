#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 wakeword-scoring path.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...           the reference algorithm on the host cores (CPU oracle)

A "step" is one pass of the hot path over one batch of synthetic streams: BASELINE.json configs[1]
(4096 streams x 10.02 s of 16 kHz f32 mono, 1 WakewordRef with 8 templates + avg_features, D=16,
defaults otherwise) per GPU; each step starts from freshly reset stream state, so it scores exactly
the windows 4096 fresh `Rustpotter`s would.  `value` = windows scored/s with the audio
already resident in HBM (CUDA events on the launching stream); `e2e` = the same through the public
batched API with the audio in pinned HOST memory, H2D copy and D2H of the detections inside the timed
region.  Per-rank work is fixed (weak scaling): streams shard across GPUs with no collective on the
data path; torch.distributed is used only for the barrier and the max-over-ranks time.

Extra objects on the JSON line: `roofline` (the DTW kernel on BASELINE configs[3]: 1M independent
(120x16 template, 100x16 window) pairs streamed from HBM, timed live with CUDA events), `cpu_baseline`
(the oracle timed on this box's host cores, rank 0, N=1 only) and `clocks`.
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

# ---- workload (BASELINE.json configs[1]) ------------------------------------------------
N_STREAMS = 4096
N_CHUNKS = 334                      # 334 * 480 = 160320 samples = 10.02 s (a whole number of 30 ms chunks)
SAMPLES = N_CHUNKS * 480
MFCC_SIZE = 16
TEMPLATE_FRAMES = (88, 92, 96, 100, 100, 96, 92, 100)
SPLICE_EVERY = 50                   # every 50th stream contains the utterance of one template
SEED = 0x5EED


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


def synth_audio_gpu(torch, n_streams: int, n_samples: int, seed: int, device):
    """0.1*N(0,1) noise + a per-stream chirp 200->3000 Hz at amplitude 0.3, clipped, never 0."""
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / 16000.0
    dur = n_samples / 16000.0
    out = torch.empty((n_streams, n_samples), dtype=torch.float32, device=device)
    rows = 256
    for b0 in range(0, n_streams, rows):
        nb = min(rows, n_streams - b0)
        ids = torch.arange(b0, b0 + nb, device=device, dtype=torch.float64)[:, None]
        f0 = 200.0 + 37.0 * (ids % 13)
        k = (3000.0 - f0) / dur
        chirp = 0.3 * torch.sin(2 * np.pi * (f0 * t + 0.5 * k * t * t) + 0.1 * ids)
        noise = 0.1 * torch.randn((nb, n_samples), generator=g, device=device, dtype=torch.float32)
        x = torch.clamp(noise + chirp.float(), -1.0, 1.0)
        x[x == 0] = 1e-4
        out[b0:b0 + nb] = x
    return out


def synth_utterance_gpu(torch, seed: int, n_frames: int, device):
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
    x = x * np.sin(np.pi * np.arange(n) / n) ** 0.5 + 0.01 * rng.standard_normal(n)
    x = np.clip(x, -1, 1).astype(np.float32)
    x[x == 0] = np.float32(1e-4)
    return x


def make_workload(torch, rp, device, n_streams: int, rank: int):
    """Templates come from OUR MFCC kernel + CMN over synthetic utterances (realistic cepstra)."""
    base = synth_utterance_gpu(torch, 1234, max(TEMPLATE_FRAMES), device)
    utts, tmpl = [], []
    for i, n in enumerate(TEMPLATE_FRAMES):
        rng = np.random.default_rng(1234 + 17 * i + 1)
        off = int(rng.integers(0, max(TEMPLATE_FRAMES) - n + 1)) * 160
        u = base[off: off + (n + 3) * 160].copy()
        u = np.clip(u * np.float32(rng.uniform(0.8, 1.1)) + 0.004 * rng.standard_normal(u.size).astype(np.float32), -1, 1).astype(np.float32)
        u[u == 0] = np.float32(1e-4)
        m = rp.mfcc_frames(torch.from_numpy(u[None]).to(device), MFCC_SIZE)[0]
        m = (m - m.mean(dim=0, keepdim=True)).cpu().numpy()
        utts.append(u)
        tmpl.append((f"sample_{i}.wav", m))
    avg = max(tmpl, key=lambda t: t[1].shape[0])[1].copy()
    rpw = cbor_rpw("hey b200", tmpl, avg, 0.05)
    audio = synth_audio_gpu(torch, n_streams, SAMPLES, SEED + 7919 * rank, device)
    for b in range(0, n_streams, SPLICE_EVERY):
        u = utts[(b // SPLICE_EVERY) % len(utts)]
        hop = 150 + (7 * b) % 600
        audio[b, hop * 160: hop * 160 + u.size] = torch.from_numpy(u).to(device)
    return rpw, audio


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


def dtw_roofline(torch, rp):
    """BASELINE configs[3]: 1M independent pairs, template 120x16 vs window 100x16, band 5 (effective
    window 20), distinct data per pair in HBM (14.08 GB > L2). Algorithmic bytes/pair = (m+n)*D*4 + 4."""
    P, m, n, d = 1_000_000, 120, 100, 16
    g = torch.Generator(device="cuda").manual_seed(1234)
    scale = torch.tensor([8, 4, 3, 2, 2, 1.5] + [1.0] * 10, device="cuda")
    a = torch.randn((P, m, d), device="cuda", generator=g) * scale
    w = torch.randn((P, n, d), device="cuda", generator=g) * scale
    out = torch.empty(P, device="cuda")
    for _ in range(2):
        rp.dtw_scores(a, w, band=5, out=out)
    torch.cuda.synchronize()
    reps = 3
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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dtw_stream_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    assert bool(torch.isfinite(out).all())
    del a, w
    return {"bound": "hbm", "kernel": "dtw pairs kernel (rp_dtw_scores)", "workload": "1M pairs 120x16 vs 100x16, band 5",
            "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
            "ms_per_launch": round(ms, 4), "bytes_per_launch": bytes_per_launch, "peak_source": how, "launches": reps + 2}


def cpu_baseline(rpw: bytes, audio_host: np.ndarray, target_seconds: float = 12.0):
    """The oracle (C++ restatement of the reference algorithm) on this box's host cores."""
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    cfg = O.default_config()
    probe = min(audio_host.shape[0], threads)
    O.run_streams(cfg, [rpw], audio_host[:probe], n_threads=threads, native=True)   # warm-up (library load, tables, threads)
    t0 = time.perf_counter()
    w0, _, _ = O.run_streams(cfg, [rpw], audio_host[:probe], n_threads=threads, native=True)
    dt = time.perf_counter() - t0
    n = int(min(audio_host.shape[0], max(probe, probe * target_seconds / max(dt, 1e-3))))
    n = max(threads, n // threads * threads)
    n = min(n, audio_host.shape[0])
    t0 = time.perf_counter()
    w, _, _ = O.run_streams(cfg, [rpw], audio_host[:n], n_threads=threads, native=True)
    dt = time.perf_counter() - t0
    return {"value": round(w / dt, 1), "unit": "windows/s", "cores": threads, "kind": "port",
            "sample": f"{n} of the {audio_host.shape[0]} streams of this workload ({w} windows, {dt:.1f} s), {threads} host threads, "
                      "C++ oracle built -O3 -march=native (the Rust reference cannot be built in this image)"}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


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


def run_ours(args):
    import torch
    saved_stdout = _stdout_to_stderr()

    import rustpotter_b200 as rp
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    n_streams = args.streams
    if args.dtw_variant:
        rp.set_dtw_variant(args.dtw_variant)
    rpw, audio = make_workload(torch, rp, device, n_streams, rank)
    audio_host = torch.empty(audio.shape, dtype=torch.float32, pin_memory=True)
    audio_host.copy_(audio)
    torch.cuda.synchronize()

    bt = rp.RustpotterBatch(n_streams, rp.default_config(), device=local)
    bt.add_wakeword_from_buffer("wakeword", rpw)
    stream = torch.cuda.current_stream()
    bt.set_cuda_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(src):
        bt.reset()
        w0 = bt.windows_scored()
        dets = bt.process(src)
        return bt.windows_scored() - w0, len(dets)

    # ---- resident leg (value) ----
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        one_step(audio)
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    windows = dets_total = launches = 0
    stage = {}
    e0.record()
    for _ in range(args.steps):
        w, nd = one_step(audio)
        windows += w
        dets_total += nd
        launches += bt.last_launches()
        for k, v in bt.last_timings().items():
            stage[k] = stage.get(k, 0.0) + v
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms_res = e0.elapsed_time(e1)
    # ---- end-to-end leg (host pinned audio through the public API) ----
    for _ in range(min(args.warmup, 2)):
        one_step(audio_host)
    barrier()
    t0 = time.perf_counter()
    windows_e2e = 0
    d2h_bytes = 0
    for _ in range(args.steps):
        w, nd = one_step(audio_host)
        windows_e2e += w
        d2h_bytes += 4 + nd * 0  # hit-list bytes are added below from the engine's record stride
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3

    t_res = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=device)
    w_all = torch.tensor([windows, windows_e2e, launches], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(w_all, op=dist.ReduceOp.SUM)
    ms_res, ms_e2e = float(t_res[0]), float(t_res[1])
    windows, windows_e2e, launches = (float(x) for x in w_all)

    if rank == 0:
        line = {
            "metric": "audio windows scored/sec (batched streams); DTW HBM GB/s vs peak in `roofline`",
            "value": round(windows / (ms_res * 1e-3), 1),
            "unit": "windows/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": round(ms_res / args.steps, 3),
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {n_streams} streams/GPU x {SAMPLES} samples (10.02 s) 16 kHz f32 mono, "
                                   "1 WakewordRef (8 templates 88..100 frames + avg_features, D=16), band 5, score_ref 0.22, "
                                   "avg_threshold 0.2, threshold 0.5, Max",
                       "streams_per_gpu": n_streams, "samples_per_stream": SAMPLES, "windows_per_step_all_gpus": windows / args.steps,
                       "l2": "inputs (2.6 GB audio/GPU) exceed the 126 MB L2; no flush needed",
                       "sharding": f"streams sharded contiguously over {world} GPU(s), no collective on the data path"},
            "e2e": {"value": round(windows_e2e / (ms_e2e * 1e-3), 1), "unit": "windows/s",
                    "h2d_bytes_per_step": int(n_streams) * SAMPLES * 4 * world,
                    "d2h_bytes_per_step": int(world * (4 + dets_total / max(args.steps, 1) * 64)),
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms_per_step_rank0": {k: round(v / args.steps, 3) for k, v in stage.items()},
            "detections_per_step_rank0": dets_total / max(args.steps, 1),
        }
        del audio
        torch.cuda.empty_cache()
        if not args.no_roofline:
            line["roofline"] = dtw_roofline(torch, rp)
        if world == 1 and not args.no_cpu:
            n_cpu = min(n_streams, 2048)
            line["cpu_baseline"] = cpu_baseline(rpw, audio_host[:n_cpu].numpy())
        _restore_stdout(saved_stdout)
        print(json.dumps(line), flush=True)
        saved_stdout = _stdout_to_stderr()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference algorithm (CPU oracle, all host threads) on the same config.
    Each step is a bounded sample of the workload's streams."""
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    # the same synthetic streams, generated on the host with numpy (no GPU on this arm)
    from tests.helpers import make_wakeword, splice, synth_audio
    rpw, utts = make_wakeword(O, d=MFCC_SIZE, lengths=TEMPLATE_FRAMES, seed=1234)
    n = max(threads, min(args.streams, threads * args.ref_streams_per_thread))
    audio = synth_audio(n, SAMPLES, seed=SEED)
    for b in range(0, n, SPLICE_EVERY):
        splice(audio[b], utts[(b // SPLICE_EVERY) % len(utts)], 150 + (7 * b) % 600)
    cfg = O.default_config()
    for _ in range(min(args.warmup, 1)):
        O.run_streams(cfg, [rpw], audio[:threads], n_threads=threads, native=True)
    t0 = time.perf_counter()
    windows = 0
    for _ in range(args.steps):
        w, _, _ = O.run_streams(cfg, [rpw], audio, n_threads=threads, native=True)
        windows += w
    dt = time.perf_counter() - t0
    val = round(windows / dt, 1)
    sample = f"{n} streams x {SAMPLES} samples per step ({windows // max(args.steps, 1)} windows/step), {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "audio windows scored/sec (batched streams)", "value": val, "unit": "windows/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] (bounded sample): {sample}; 1 WakewordRef (8 templates + avg, D=16), defaults"},
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
    ap.add_argument("--streams", type=int, default=N_STREAMS, help="streams per GPU (default: BASELINE configs[1])")
    ap.add_argument("--ref-streams-per-thread", type=int, default=2)
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--dtw-variant", type=int, default=0, help="rp_set_dtw_variant for A/B measurements (default 0 = automatic)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
