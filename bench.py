#!/usr/bin/env python
"""bench.py — audio-seconds/second of the per-chunk hot path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ragged4096|streams256|streams1024|streams4096|streams10240|lowlat4096|fbank1024]
  python bench.py --impl reference ...      # the reference's CPU implementation (torch/torchaudio port) on the host cores

A step = one pass of the hot path (PCM -> fbank -> 20-layer Emformer chunk forward with K/V rings -> CTC log-softmax ->
greedy / prefix beam) over one batch of `streams` concurrent 640 ms stream-chunks per GPU.  `value` times K steps with the PCM
batch already resident in HBM (CUDA events on the engine's own stream); `e2e` times the public API with host int16 buffers, H2D of
the PCM and D2H of the token ids inside the timed region.  Default workload = BASELINE configs[3] (the largest single-GPU
configuration): 4096 sessions at mixed progress driven by SessionScheduler (VAD gate, pinned gather, one pre-staged tick per pass, endpoint
rules), prefix beam 10.  N > 1: one process per GPU (torchrun), sessions partitioned per GPU, no data-path collective (weak
scaling), time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WEIGHT_SEED = 1234
WORKLOADS = {
    # name: (streams per GPU, description)
    "streams256": (256, "configs[2]: 256 concurrent synthetic 16 kHz streams, lightspeech encoder (random-init), chunk_size=16, greedy CTC"),
    "streams1024": (1024, "1024 concurrent streams, chunk_size=16, greedy CTC"),
    "streams4096": (4096, "configs[3] batch size: 4096 concurrent streams, chunk_size=16, greedy CTC"),
    "streams10240": (10240, "north-star target: 10,240 concurrent streams, chunk_size=16, greedy CTC"),
    "fbank1024": (1024, "configs[1]: fbank-only, 1024 streams x 640 ms, 80-bin Kaldi fbank"),
    "ragged4096": (4096, "configs[3]: 4096 concurrent streams ragged-batched by the session scheduler (mixed progress: fresh / 1-chunk / steady-state "
                         "left context), chunk_size=16, CTC prefix beam search (beam 10, 8 candidates/frame) + greedy, energy-gate VAD standing in for "
                         "Silero (model absent) with ~20 % of chunks gated out, endpoints at ~1 %/s per stream + online_endpoint rules"),
    "realtime": (10240, "north-star operating point: N real-time 16 kHz streams (default 10,240; --streams N) delivering 640 ms of audio every 640 ms "
                        "at random phases, served by SessionScheduler ticks (two in flight, greedy CTC); reports per-chunk latency p50 / p99"),
    "lowlat4096": (4096, "configs[4] per-GPU share: 4096 concurrent streams per GPU, chunk_size=8 low-latency mode (320 ms chunks), greedy CTC"),
    "longform": (4096, "configs[4] long form: 4096 concurrent streams per GPU (32k on 8 GPUs), chunk_size=8 low-latency mode, 10 minutes of audio per stream "
                       "(1875 chunks of 320 ms; --long-chunks), SessionScheduler with one pre-staged full-size tick per pass, energy-gate VAD, endpoint rules (forced rule4 endpoints at 40 s)"),
}
FLOP_PER_STREAM_CHUNK = 2_583_363_584          # SURVEY.md §8a (L_valid = 32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "bf16_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_burst": 1590.0, "source": "fallback"}


def ncu_traffic(workload: str, family: str):
    """dram bytes per launch of `family` from the committed ncu --set full capture of the same workload, else None."""
    p = os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(workload, {}).get(family)
    except (OSError, ValueError):
        return None


def synth_pcm(n_streams: int, n_samples: int, first_id: int = 0) -> np.ndarray:
    """int16 audio ~ N(0, (0.1*32768)^2) + a 440 Hz tone, seeded per stream id."""
    out = np.empty((n_streams, n_samples), np.int16)
    t = np.arange(n_samples) / 16000.0
    for i in range(n_streams):
        rng = np.random.Generator(np.random.PCG64(1234 + first_id + i))
        x = 0.1 * rng.standard_normal(n_samples) + 0.05 * np.sin(2 * np.pi * 440.0 * t)
        out[i] = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.sm_max = index, threading.Event(), [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:      # NVML unavailable: report that, do not invent clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        self.stop_flag.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def dist_env():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(steps: int, warmup: int, sample_streams: int, threads: int):
    """The reference's CPU path (oracle/torch_ref_port.py: torchaudio Emformer + the reference glue), batch 1 per stream
    as the live server does (streaming_server.py:420-422), greedy_search on the accumulated emission every chunk (:431-433)."""
    import torch
    from oracle import lightspeech_oracle as O
    from oracle.torch_ref_port import TorchRefPort, greedy_search
    torch.set_num_threads(threads)
    W = O.make_weights(WEIGHT_SEED)
    m = TorchRefPort(W)
    geo = O.CANONICAL
    vocab = ["-", "|"] + [f"<{i}>" for i in range(2, geo.vocab)]
    pcm = synth_pcm(sample_streams, geo.chunk_length).astype(np.float32) / np.float32(32768.0)
    speeches = [torch.from_numpy(pcm[i])[None] for i in range(sample_streams)]
    states = [m.init_state() for _ in range(sample_streams)]
    emissions = [torch.zeros(0, geo.vocab) for _ in range(sample_streams)]

    def one_step():
        for i in range(sample_streams):
            em, ln, st = m.stream([speeches[i]], 16000, [states[i]])
            states[i] = st[0]
            emissions[i] = torch.cat((emissions[i], em[0]))[-160:]       # bounded segment (an endpoint every 10 chunks)
            greedy_search(emissions[i], vocab)

    for _ in range(warmup):
        one_step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one_step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    audio_s = steps * sample_streams * 0.64
    return audio_s / total, 1000.0 * total / steps, times


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = args.ref_streams
    value, ms, _ = cpu_reference_run(args.steps, args.warmup, sample, threads)
    streams, desc = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": desc, "streams_per_gpu": streams, "chunk_ms": 640},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the {streams} streams per step, batch 1 per stream (the live server's call pattern), "
                                   f"torch.set_num_threads({threads}); torchaudio Emformer/MelSpectrogram + restated reference glue"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ configs[3]: scheduler-driven workload
class RaggedWorkload:
    """4096 sessions with a scripted speech / silence pattern (2-state Markov chain per session, ~20 % of chunks silent, an
    endpoint at every speech -> silence transition: ~1 % of the streams per second), driven through SessionScheduler with two
    ticks in flight.  Audio for every pass is pre-loaded into the scheduler's per-session rings (the websocket receive path is
    not part of the measured hot path); everything from 'which sessions are ready' to 'token ids on the host' is timed."""

    P_END, P_START = 0.008, 0.032

    def __init__(self, eng, cfg, streams, passes, seed, pool, device_gather=None):
        from asr_streaming_b200 import SessionScheduler
        from asr_streaming_b200.endpoint import EndpointRules
        from asr_streaming_b200.scheduler import native_energy_gate
        self.eng, self.cfg, self.n, self.passes = eng, cfg, streams, passes
        self.sch = SessionScheduler(eng, capacity=streams, backlog_chunks=passes, endpoint_rules=EndpointRules(), device_gather=device_gather)
        self.gate = native_energy_gate()
        self.sess = [self.sch.open() for _ in range(streams)]
        rng = np.random.Generator(np.random.PCG64(seed))
        speech = np.empty((streams, passes), bool)
        state = rng.random(streams) >= 0.2                          # stationary start: 20 % of the sessions in silence
        for k in range(passes):
            speech[:, k] = state
            flip = rng.random(streams)
            state = np.where(state, flip >= self.P_END, flip < self.P_START)
        self.speech = speech
        seg = cfg.segment_length
        block = np.empty((streams, passes * seg), np.int16)
        shift = rng.integers(0, pool.shape[1] - passes * seg, size=streams)
        for i in range(streams):
            block[i] = pool[i % pool.shape[0], shift[i]:shift[i] + passes * seg]
        block.reshape(streams, passes, seg)[~speech] = 0
        self.sch.accept_block(np.arange(streams), block)
        self.sch.wr[:] = cfg.buffer_length                          # nothing has "arrived" yet: feed_pass() releases 640 ms at a time
        self.fed = 0
        self.run_chunks = self.skipped_chunks = self.endpoints = 0
        self.t_submit = self.t_collect = self.t_after = 0.0       # host seconds spent in submit_tick / collect_tick / scripted endpoints
        self.n_ticks = 0
        self._prev = None

    def feed_pass(self):
        """The next 640 ms of every stream arrive (the samples are already in the rings; only the write pointers move)."""
        assert self.fed < self.passes, "RaggedWorkload: out of pre-loaded audio"
        self.sch.wr += self.cfg.segment_length
        self.fed += 1

    def _after(self, res):
        """Scripted endpoints: a session whose just-decoded chunk was the last speech chunk of an utterance is reset."""
        self.skipped_chunks += int(res.skipped_rows.size)
        rows = res.rows
        if rows.size == 0:
            return
        k = self.sch.chunk_processed_total[rows] - 1               # index of the chunk just decoded
        nxt = np.minimum(k + 1, self.passes - 1)
        ended = self.speech[rows, k] & ~self.speech[rows, nxt] & ~res.final
        self.sch.reset_rows(rows[ended])
        self.run_chunks += int(rows.size)
        self.endpoints += int(ended.sum()) + int(res.final.sum())

    def run_pipelined(self, n_passes, max_rows):
        """n_passes x 640 ms for every session; ticks of <= max_rows sessions, two in flight, never draining between passes."""
        prev = None
        for _ in range(n_passes):
            self.feed_pass()
            while True:
                t0 = time.perf_counter()
                p = self.sch.submit_tick(gate=self.gate, max_rows=max_rows)
                self.t_submit += time.perf_counter() - t0
                if p.rows.size == 0 and p.res.skipped_rows.size == 0:
                    break                                           # nothing left to launch in this pass; `prev` stays in flight
                self.n_ticks += 1
                if prev is not None:
                    t0 = time.perf_counter()
                    res = self.sch.collect_tick(prev)
                    t1 = time.perf_counter()
                    self._after(res)
                    self.t_collect += t1 - t0
                    self.t_after += time.perf_counter() - t1
                    prev = None
                if p.rows.size:
                    prev = p
                else:
                    self._after(p.res)                              # VAD-skipped only: nothing to collect
        if prev is not None:
            self._after(self.sch.collect_tick(prev))

    def run_prestaged(self, n_passes):
        """n_passes x 640 ms for every session, ONE full-size tick per pass: as soon as a pass's audio is there its chunks are gathered and
        their H2D copy starts (SessionScheduler.prestage) while the previous tick's kernels still run; that tick is then collected (its
        bookkeeping, endpoint rules and resets decide which of the staged chunks run) and the next tick launches on the staged data."""
        for _ in range(n_passes):
            self.feed_pass()
            t0 = time.perf_counter()
            self.sch.prestage(self.gate)
            dt = time.perf_counter() - t0
            self.t_submit += dt
            self.t_prestage = getattr(self, "t_prestage", 0.0) + dt
            if self._prev is not None:
                t0 = time.perf_counter()
                res = self.sch.collect_tick(self._prev)
                t1 = time.perf_counter()
                self._after(res)
                self.t_collect += t1 - t0
                self.t_after += time.perf_counter() - t1
                self._prev = None
            t0 = time.perf_counter()
            p = self.sch.submit_tick(gate=self.gate)
            self.t_submit += time.perf_counter() - t0
            self.n_ticks += 1
            if p.rows.size:
                self._prev = p
            else:
                self._after(p.res)

    def drain(self):
        if self._prev is not None:
            self._after(self.sch.collect_tick(self._prev))
            self._prev = None

    def run_sync(self):
        """One pass, one synchronous tick over every ready session: the per-chunk latency a session sees."""
        self.feed_pass()
        t0 = time.perf_counter()
        res = self.sch.tick(gate=self.gate)
        dt = time.perf_counter() - t0
        self._after(res)
        return dt, len(res)


def realtime_measure(local: int, rank: int, n: int, chunks: int):
    """Wall-clock simulation of N real-time streams: chunk k of stream i becomes ready at phase_i + k * 0.64 s; latency = time from
    'ready' to 'token ids of that chunk on the host'.  Everything between (ready scan, batch assembly, H2D, kernels, D2H,
    bookkeeping, endpoint rules) is inside.  Not a throughput benchmark: the GPU idles between ticks when N is small.
    Returns (audio-s/s, mean tick ms, clocks, report dict)."""
    from asr_streaming_b200 import Engine, ModelConfig, PRECISION_FAST, SessionScheduler, pack_weights, random_weights
    from asr_streaming_b200.endpoint import EndpointRules
    cfg = ModelConfig(precision=PRECISION_FAST, max_batch=min(n, 4096), max_sessions=n)
    eng = Engine(cfg, pack_weights(random_weights(WEIGHT_SEED, cfg), cfg), local)
    sch = SessionScheduler(eng, capacity=n, backlog_chunks=chunks + 2, endpoint_rules=EndpointRules())
    for _ in range(n):
        sch.open()
    seg, chunk_s = cfg.segment_length, cfg.segment_length / cfg.sample_rate
    total = (chunks + 2) * seg
    pool = synth_pcm(16, total + 8192, first_id=rank * 16)
    rng = np.random.Generator(np.random.PCG64(17 + rank))
    shift = rng.integers(0, 8192, size=n)
    for i in range(n):
        sch.audio[i, cfg.buffer_length:cfg.buffer_length + total] = pool[i % 16, shift[i]:shift[i] + total]
    sch.wr[:] = cfg.buffer_length
    phase = rng.random(n) * chunk_s
    # warm-up: two chunks for everybody as fast as possible (left context filled, kernels / allocator warm)
    sch.wr += 2 * seg
    prev = None
    while True:
        p = sch.submit_tick()
        if prev is not None:
            sch.collect_tick(prev)
        prev = p if p.rows.size else None
        if prev is None and not sch.ready_rows().size:
            break
    eng.sync()
    released = np.zeros(n, np.int64)
    base_done = sch.chunk_processed_total.copy()
    lat, batch_sizes, tick_ms = [], [], []
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    prev, prev_t = None, 0.0
    while True:
        now = time.perf_counter() - t0
        k = np.clip(np.floor((now - phase) / chunk_s).astype(np.int64) + 1, 0, chunks)
        newly = k > released
        if newly.any():
            sch.wr[newly] += (k - released)[newly] * seg
            released = np.maximum(released, k)
        ts = time.perf_counter()
        p = sch.submit_tick()
        if prev is not None:
            res = sch.collect_tick(prev)
            t_done = time.perf_counter() - t0
            c = sch.chunk_processed_total[res.rows] - base_done[res.rows] - 1        # index of the chunk just decoded
            lat.append(t_done - (phase[res.rows] + c * chunk_s))
            batch_sizes.append(res.rows.size)
            tick_ms.append(1e3 * (time.perf_counter() - prev_t))
        prev, prev_t = (p, ts) if p.rows.size else (None, 0.0)
        if prev is None:
            if (released >= chunks).all() and (sch.chunk_processed_total - base_done >= chunks).all():
                break
            time.sleep(0.0002)
    wall = time.perf_counter() - t0
    clocks = sampler.result()
    lat = np.concatenate(lat) * 1e3
    audio = float((sch.chunk_processed_total - base_done).sum()) * chunk_s
    eng.close()
    rep = {"streams": n, "chunks_per_stream": chunks, "wall_s": wall, "keeps_up": bool(wall < chunks * chunk_s + 1.0),
           "chunk_latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "p999": float(np.percentile(lat, 99.9)),
                                "max": float(lat.max()), "what": "chunk ready (its last sample arrived) -> its token ids on the host"},
           "ticks": len(batch_sizes), "mean_sessions_per_tick": float(np.mean(batch_sizes)), "max_sessions_per_tick": int(np.max(batch_sizes)),
           "mean_tick_ms": float(np.mean(tick_ms)), "budget_ms": 40.0}
    return audio / wall, float(np.mean(tick_ms)), clocks, rep


def run_realtime(args):
    import torch
    rank, world, local = dist_env()
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    n = args.streams or WORKLOADS["realtime"][0]
    chunks = max(4, args.steps)
    value, tick_ms, clocks, rep = realtime_measure(local, rank, n, chunks)
    line = {"metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": 1, "steps": chunks, "warmup": 2, "ms_per_step": tick_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS["realtime"][1], "streams_per_gpu": n, "chunk_ms": 640, "weights": f"random-init seed {WEIGHT_SEED}"},
            "clocks": clocks, "realtime": rep}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    return 0



# ------------------------------------------------------------------------------------------ rooflines per kernel family
def ncu_traffic_table():
    """dram bytes per launch per kernel family from the committed ncu --set full captures (profiles/r02_roofline_traffic.json)."""
    for name in ("r02_roofline_traffic.json", "r01_roofline_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
                d["_file"] = name
                return d
        except (OSError, ValueError):
            continue
    return {}


def family_rooflines(fam, cfg, streams, precision_exact, pk, workload):
    """Per kernel family: algorithmic FLOPs (tensor-bound GEMMs) or algorithmic HBM bytes (memory-bound kernels) per launch divided by the
    mean launch time measured live with CUDA events on the engine stream (asr_profile_*), against the measured peaks.  The per-unit
    figures are DESIGN.md section 4's: M = rows * streams GEMM rows; attention reads each stream's left-context K and V ring rows + q and
    writes its rows of the out_proj operand; fbank reads the int16 chunk and writes the bf16 operand."""
    M, d, f, V = streams * cfg.rows, cfg.d_model, cfg.ffn_dim, cfg.vocab
    esz = 4 if precision_exact else 2                   # bytes per K/V / q element (EXACT: pre-split hi|lo bf16 rows, fp32 q)
    traffic = ncu_traffic_table().get(workload, {})
    tens = {"gemm_qkv": 2 * M * 3 * d * d, "gemm_out_proj": 2 * M * d * d, "gemm_ffn1": 2 * M * f * d, "gemm_ffn2": 2 * M * d * f,
            "gemm_input_linear": 2 * streams * cfg.frames * (d // cfg.stride) * cfg.n_mels,
            "gemm_ctc1": 2 * streams * cfg.seg_rows * d * cfg.ctc_hidden, "gemm_ctc2": 2 * streams * cfg.seg_rows * cfg.ctc_hidden * V}
    # attention: the left-context K and V rows of every stream (written a step ago: 8 GB of rings never survive in L2) + q in + out_proj operand
    # out; the segment / right-context K and V rows come straight from the QKV kernel before it and are not counted
    hbm = {"attention": streams * (2 * cfg.left_context * d * esz + cfg.rows * d * esz + cfg.rows * d * (4 if precision_exact else 2)),
           "fbank": streams * (cfg.chunk_length * 2 + cfg.frames * cfg.n_mels * (4 if precision_exact else 2)),
           "layernorm": M * d * (4 + esz), "ctc_greedy": streams * cfg.seg_rows * V * 4,
           # prefix beam: per frame the 8 candidates (id + log-prob), the row's (max, lse), the blank logit and one gathered logit per beam
           # entry (a 32-byte sector each); the kernel is a serial recursion over the chunk's frames per stream, bound by latency, not bytes
           "beam": streams * cfg.seg_rows * (8 * 8 + 8 + 32 * (1 + 10))}
    out = {}
    for k, v in fam.items():
        if not v["launches_per_step"]:
            continue
        t = v["ms_per_step"] / v["launches_per_step"] / 1e3
        if k in tens:
            ach = tens[k] / t / 1e12
            out[k] = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                      "us_per_launch": 1e6 * t, "traffic": traffic.get(k), "executed_flops_multiplier": 3 if precision_exact else 1}
        elif k in hbm:
            ach = hbm[k] / t / 1e9
            out[k] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "us_per_launch": 1e6 * t,
                      "traffic": traffic.get(k), "algorithmic_bytes_per_launch": hbm[k]}
    return out


def device_leg(local, blob, streams, precision, low_latency, steps, warmup, beam=0):
    """Device-resident throughput of `streams` steady-state streams per step on a fresh engine: W warm-up + K timed steps, CUDA events on
    the engine stream, then the per-family profile.  Returns (audio-s/s, ms/step, families, cfg, launches)."""
    import torch
    from asr_streaming_b200 import Engine, ModelConfig
    cfg = ModelConfig(precision=precision, max_batch=streams, max_sessions=streams, segment_size=32 if low_latency else 64)
    eng = Engine(cfg, blob, local)
    try:
        if beam:
            eng.set_beam(beam, 8)
        ext = torch.cuda.ExternalStream(eng.cuda_stream, device=local)
        sl = [eng.open_session() for _ in range(streams)]
        eng.stage(sl, synth_pcm(streams, cfg.chunk_length))
        for _ in range(max(warmup, 3)):
            eng.run_staged(streams)
        eng.sync()
        l0 = eng.stats()["kernel_launches"]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ext)
        for _ in range(steps):
            eng.run_staged(streams)
        b.record(ext)
        eng.sync()
        ms = a.elapsed_time(b) / steps
        launches = eng.stats()["kernel_launches"] - l0
        eng.profile_enable(True)
        ps = min(steps, 3)
        for _ in range(ps):
            eng.run_staged(streams)
        prof = eng.profile_read()
        eng.profile_enable(False)
        fam = {k: {"ms_per_step": v[0] / ps, "launches_per_step": v[1] / ps} for k, v in prof.items() if v[1]}
        return streams * (cfg.segment_length / cfg.sample_rate) / (ms / 1e3), ms, fam, cfg, launches
    finally:
        eng.close()


def fbank_cpu_baseline(threads: int, sample_streams: int = 64):
    """BASELINE.md section 3, configs[1]: the reference's own front-ends on the host — extract_filterbank (audio.py:9-30: MelSpectrogram
    rebuilt every call, as there) and torchaudio.compliance.kaldi.fbank(num_mel_bins=80, dither=0) — batch 1 per stream, 640 ms each."""
    import torch
    import torchaudio
    torch.set_num_threads(threads)
    pcm = synth_pcm(sample_streams, 10240 + 240).astype(np.float32)
    x = [torch.from_numpy(pcm[i])[None] for i in range(sample_streams)]

    def kaldi():
        for xi in x:
            torchaudio.compliance.kaldi.fbank(xi, num_mel_bins=80, dither=0.0, sample_frequency=16000.0)

    def melspec():
        for xi in x:
            tr = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=800, win_length=400, hop_length=160, n_mels=128, center=False)
            torch.transpose(tr(xi / 32768.0).clamp(1e-5).log(), 2, 1)
    out = {}
    for name, fn in (("kaldi80", kaldi), ("melspec128", melspec)):
        fn()
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 3.0:
            fn()
            reps += 1
        out[name] = reps * sample_streams * 0.64 / (time.perf_counter() - t0)
    return out


def fbank_leg(local, blob, pk, streams=1024, steps=50):
    """BASELINE configs[1]: 80-bin Kaldi fbank of 1024 streams x 640 ms (10,240 new samples + 240 of framing history), PCM resident in HBM."""
    import torch
    from asr_streaming_b200 import Engine, ModelConfig, PRECISION_FAST
    cfg = ModelConfig(precision=PRECISION_FAST, max_batch=streams, max_sessions=streams)
    eng = Engine(cfg, blob, local)
    try:
        ext = torch.cuda.ExternalStream(eng.cuda_stream, device=local)
        n_samples = 10240 + 240
        pcm = synth_pcm(streams, n_samples)
        eng.stage_raw(pcm)
        for _ in range(5):
            eng.fbank_staged(1, streams, 0, n_samples)
        eng.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ext)
        for _ in range(steps):
            eng.fbank_staged(1, streams, 0, n_samples)
        b.record(ext)
        eng.sync()
        us = 1e3 * a.elapsed_time(b) / steps
        t0 = time.perf_counter()
        for _ in range(10):
            eng.fbank(pcm, kind=1)
        e2e = 10 * streams * 0.64 / (time.perf_counter() - t0)
        alg = pcm.nbytes + streams * 64 * 80 * 4
        ach = alg / (us / 1e6) / 1e9
        traffic = ncu_traffic_table().get("fbank1024", {}).get("fbank")
        return {"workload": WORKLOADS["fbank1024"][1], "value": streams * 0.64 / (us / 1e6), "unit": "audio-s/s", "us_per_launch": us,
                "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm.nbytes), "d2h_bytes_per_step": streams * 64 * 80 * 4},
                "roofline": {"kernel": "fbank_kernel<256,int16,kaldi>", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                             "traffic": traffic, "algorithmic_bytes_per_launch": alg, "peak_source": pk["source"]}}
    finally:
        eng.close()


def run_longform(args):
    """BASELINE configs[4], long form.  Every stream delivers `--long-chunks` chunks of 320 ms (default 1875 = 10 minutes); the timed region is
    the whole run through the public API (ready scan -> gate -> gather -> H2D -> kernels -> D2H -> bookkeeping -> endpoint rules), audio
    arriving from host memory pass by pass.  `value` is the device-resident low-latency step for the same stream count."""
    import torch
    from asr_streaming_b200 import Engine, ModelConfig, PRECISION_FAST, SessionScheduler, pack_weights, random_weights
    from asr_streaming_b200.endpoint import EndpointRules
    from asr_streaming_b200.scheduler import native_energy_gate
    rank, world, local = dist_env()
    if world > 1:
        os.environ.setdefault("ASR_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    n = args.streams or WORKLOADS["longform"][0]
    chunks = args.long_chunks
    pk = peaks()
    cfg0 = ModelConfig(segment_size=32)
    blob = pack_weights(random_weights(WEIGHT_SEED, cfg0), cfg0)
    value, ms_step, fam, cfg, _ = device_leg(local, blob, n, PRECISION_FAST, True, 10, 3)
    eng = Engine(ModelConfig(precision=PRECISION_FAST, max_batch=n, max_sessions=n, segment_size=32), blob, local)
    dev_gather = world > 1 and (os.cpu_count() or 8) // world < 8
    sch = SessionScheduler(eng, capacity=n, backlog_chunks=3, endpoint_rules=EndpointRules(), device_gather=dev_gather)
    for _ in range(n):
        sch.open()
    seg = cfg.segment_length
    P = 64
    pool = synth_pcm(32, P * seg + seg, first_id=rank * 32)
    rng = np.random.Generator(np.random.PCG64(5 + rank))
    speech = rng.random((n, P)) < 0.8                          # periodic speech / silence pattern (P chunks), sticky
    for k in range(1, P):
        speech[:, k] = np.where(rng.random(n) < 0.85, speech[:, k - 1], speech[:, k])
    src = np.arange(n) % 32                                    # 32 distinct audio sources fanned out over the streams
    shift = rng.integers(0, seg, size=32)
    rows = np.arange(n)
    # what "arrives" every 320 ms, prepared up front for one period of the pattern (the websocket receive path is not the measured path;
    # handing a block to the scheduler — asr_sched_accept_block: ring compaction + copy — is, and it is timed)
    blocks = []
    for kk in range(min(P, chunks)):
        base = np.stack([pool[j, shift[j] + kk * seg: shift[j] + kk * seg + seg] for j in range(32)])
        block = base[src]
        block[~speech[:, kk]] = 0
        blocks.append(block)
    gate = native_energy_gate()
    lat = []
    two_ticks = os.environ.get("ASR_BENCH_E2E_ROWS") is not None

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    state = {"prev": None}

    def run(n_chunks, timed):
        """n_chunks x 320 ms for every session; ticks of <= n / 2 sessions, two in flight, never draining between passes."""
        done = ends = skips = 0

        def collect():
            nonlocal done, ends
            p, t0 = state["prev"]
            res = sch.collect_tick(p)
            if timed:
                lat.append(1e3 * (time.perf_counter() - t0))
            done += len(res)
            ends += len(res.final_tokens)
            state["prev"] = None
        for k in range(n_chunks):
            sch.accept_block(rows, blocks[k % len(blocks)])
            if not two_ticks:
                # one full-size tick per pass: gather + H2D of this pass start now, under the previous tick's kernels (SessionScheduler.prestage)
                sch.prestage(gate)
                if state["prev"] is not None:
                    collect()
                t0 = time.perf_counter()
                p = sch.submit_tick(gate=gate)
                skips += int(p.res.skipped_rows.size)
                if p.rows.size:
                    state["prev"] = (p, t0)
                else:
                    ends += len(p.res.final_tokens)
                continue
            while True:                                        # ASR_BENCH_E2E_ROWS: two ticks of <= n / 2 sessions per pass, two in flight
                t0 = time.perf_counter()
                p = sch.submit_tick(gate=gate, max_rows=n // 2)
                skips += int(p.res.skipped_rows.size)
                if p.rows.size == 0:
                    ends += len(p.res.final_tokens)
                    if p.res.skipped_rows.size == 0:
                        break                                   # nothing left to launch in this pass; the tick in flight stays in flight
                    continue
                if state["prev"] is not None:
                    collect()
                state["prev"] = (p, t0)
        return done, ends, skips

    def drain():
        if state["prev"] is not None:
            p, _ = state["prev"]
            res = sch.collect_tick(p)
            state["prev"] = None
            return len(res), len(res.final_tokens)
        return 0, 0
    run(4, False)                                              # warm-up: left context filled, kernels warm
    drain()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    done, ends, skips = run(chunks, True)
    d2, e2 = drain()
    done, ends = done + d2, ends + e2
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.result()
    if world > 1:
        t = torch.tensor([wall, -float(done), -float(ends), -float(skips)], dtype=torch.float64, device=f"cuda:{local}")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        wall = float(t[0].item())
        tot = torch.tensor([float(done), float(ends), float(skips)], dtype=torch.float64, device=f"cuda:{local}")
        torch.distributed.all_reduce(tot)
        done_all, ends_all, skips_all = (float(x) for x in tot.tolist())
    else:
        done_all, ends_all, skips_all = float(done), float(ends), float(skips)
    chunk_s = seg / cfg.sample_rate
    if rank == 0:
        lat_a = np.asarray(lat)
        line = {"metric": "audio-sec/sec", "value": value * world, "unit": "audio-s/s", "n_gpus": world, "steps": chunks, "warmup": 4, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOADS["longform"][1], "streams_per_gpu": n, "chunk_ms": 320, "chunks_per_stream": chunks,
                           "audio_minutes_per_stream": chunks * chunk_s / 60.0, "weights": f"random-init seed {WEIGHT_SEED}",
                           "parallelism": f"sessions partitioned per GPU x{world}, no collective"},
                "clocks": clocks,
                "e2e": {"value": done_all * chunk_s / wall, "unit": "audio-s/s", "h2d_bytes_per_step": int(done / max(chunks, 1) * (cfg.chunk_length * 2 + 4)),
                        "d2h_bytes_per_step": int(done / max(chunks, 1) * (cfg.seg_rows * 8 + 20)), "wall_s": wall,
                        "decoded_chunks": done_all, "vad_skipped_chunks": skips_all, "endpoints": ends_all,
                        "realtime_factor_per_stream": (chunks * chunk_s) / wall,
                        "what": "decoded audio-seconds of all streams / wall time of the whole long-form run (VAD-skipped chunks not counted)"},
                "gpu_launches": None,
                "chunk_latency_ms": {"p50": float(np.percentile(lat_a, 50)), "p99": float(np.percentile(lat_a, 99)), "max": float(lat_a.max()),
                                     "what": "submit_tick -> collect_tick of one full-size tick (one per pass, batch pre-staged under the previous tick)"},
                "kernel_rooflines": family_rooflines(fam, cfg, n, False, pk, "lowlat4096")}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        torch.distributed.destroy_process_group()
    sch.close_scheduler()
    eng.close()
    return 0


def run_ours(args):
    import torch
    from asr_streaming_b200 import Engine, ModelConfig, PRECISION_EXACT, PRECISION_FAST, pack_weights, random_weights
    rank, world, local = dist_env()
    if world > 1:                                           # ranks share the host: split its cores between their gather / gate threads
        os.environ.setdefault("ASR_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
    # Only the final JSON line may reach stdout: NCCL / torchrun chatter is routed to stderr for the duration of the run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    streams, desc = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    precision = PRECISION_EXACT if args.precision == "exact" else PRECISION_FAST
    pk = peaks()
    fbank_only = args.workload.startswith("fbank")

    ragged = args.workload.startswith("ragged")
    low_latency = args.workload.startswith("lowlat")
    cfg = ModelConfig(precision=precision, max_batch=streams, max_sessions=streams, segment_size=32 if low_latency else 64)
    blob = pack_weights(random_weights(WEIGHT_SEED, cfg), cfg)
    eng = Engine(cfg, blob, local)
    ext = torch.cuda.ExternalStream(eng.cuda_stream, device=local)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    extra = {}
    if fbank_only:
        n_samples = 10240 + 240
        pcm = synth_pcm(streams, n_samples, first_id=rank * streams)
        eng.stage_raw(pcm)
        run = lambda: eng.fbank_staged(1, streams, 0, n_samples)
        e2e_call = lambda: eng.fbank(pcm, kind=1)
        audio_per_step = streams * 0.64
        h2d, d2h = pcm.nbytes, streams * 64 * 80 * 4
        alg_bytes = pcm.nbytes + streams * 64 * 80 * 4
    elif ragged:
        # configs[3]: prefix beam 10 on every step, sessions at mixed progress, a trickle of endpoints (1 %/s per stream)
        eng.set_beam(10, 8)
        lat_passes = min(args.steps, 8)
        e2e_steps = min(args.steps, 24)                        # bounded host memory: every pass pre-loads 84 MB of PCM into the rings
        passes = 2 + lat_passes + 2 + e2e_steps + 1
        pool = synth_pcm(32, cfg.buffer_length + passes * cfg.segment_length + 4096, first_id=rank * 32)
        # batch assembly: host gather + DMA unless the ranks leave each other fewer than 8 host threads (see SessionScheduler.__init__)
        dev_gather = os.environ.get("ASR_B200_DEVICE_GATHER")
        dev_gather = (dev_gather == "1") if dev_gather is not None else (world > 1 and (os.cpu_count() or 8) // world < 8)
        wl = RaggedWorkload(eng, cfg, streams, passes, seed=99 + rank, pool=pool, device_gather=dev_gather)
        slots = wl.sch.slot.copy()
        pcm = np.stack([pool[i % 32, :cfg.chunk_length] for i in range(streams)])
        eng.stage(slots, pcm)
        rs = np.random.Generator(np.random.PCG64(7 + rank))
        n_reset = max(1, int(round(streams * 0.01 * cfg.segment_length / cfg.sample_rate)))

        def run():
            eng.run_staged(streams)
            eng.reset_sessions(slots[rs.choice(streams, n_reset, replace=False)])
        eng.run_staged(streams)                                  # mixed progress before the timed region: thirds at 0 / 16 / 32 rows
        eng.reset_sessions(slots[: streams // 3])
        eng.run_staged(streams)
        eng.reset_sessions(slots[streams // 3: 2 * streams // 3])
        audio_per_step = streams * cfg.segment_length / cfg.sample_rate
        h2d = d2h = 0
    else:
        pcm = synth_pcm(streams, cfg.chunk_length, first_id=rank * streams)
        slots = [eng.open_session() for _ in range(streams)]
        eng.stage(slots, pcm)
        run = lambda: eng.run_staged(streams)
        e2e_call = lambda: eng.step(slots, pcm)
        audio_per_step = streams * cfg.segment_length / cfg.sample_rate
        h2d = pcm.nbytes + 4 * streams
        d2h = streams * cfg.seg_rows * 4 * 2 + 3 * 4 * streams

    # ---- device-resident leg: W warm-up, K timed, CUDA events on the engine stream
    for _ in range(max(args.warmup, 3)):
        run()
    barrier()
    l0 = eng.stats()["kernel_launches"]
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps):
        run()
    ev1.record(ext)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.stats()["kernel_launches"] - l0

    # ---- per-kernel-family device time (same workload, events around every launch) -> roofline of the dominant kernel
    eng.profile_enable(True)
    prof_steps = min(args.steps, 5)
    for _ in range(prof_steps):
        run()
    prof = eng.profile_read()
    eng.profile_enable(False)

    # ---- end-to-end leg through the public API: inputs in pinned host memory, H2D + kernels + D2H of the results every step.
    # (a) synchronous Engine.step: per-chunk compute latency (enqueue of a ready batch -> token ids on the host), p50 / p99
    # (b) pipelined Engine.submit/collect (two steps in flight: the H2D of step k+1 overlaps the kernels of step k): throughput
    lat_ms = []
    if fbank_only:
        for _ in range(2):
            e2e_call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_call()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    elif ragged:
        eng.reset_sessions(slots)                            # the device-resident leg ran on the scheduler's slots: back to fresh state
        barrier()
        for _ in range(2):
            wl.run_sync()
        for _ in range(lat_passes):
            dt, n_run = wl.run_sync()
            lat_ms.append(1e3 * dt)
        # ONE full-size tick per pass with pre-staging — the gather + H2D (or the device gather out of pinned rings) of pass k + 1 overlap the
        # kernels of pass k; only the bookkeeping / endpoint rules between collect and the next launch are exposed.
        # ASR_BENCH_E2E_ROWS=<rows> forces the older form: two ticks of that size per pass, two in flight.
        forced = os.environ.get("ASR_BENCH_E2E_ROWS")
        prestaged = forced is None
        e2e_rows = streams if prestaged else int(forced or streams // 2)
        runner = wl.run_prestaged if prestaged else (lambda k: wl.run_pipelined(k, e2e_rows))
        runner(2)
        if prestaged:
            wl.drain()
        barrier()
        c0 = (wl.run_chunks, wl.skipped_chunks, wl.endpoints)
        h0 = (wl.t_submit, wl.t_collect, wl.t_after, wl.n_ticks, getattr(wl, "t_prestage", 0.0))
        eng.pipeline_gpu_time(reset=True)
        t0 = time.perf_counter()
        runner(e2e_steps)
        if prestaged:
            wl.drain()
        barrier()
        gpu_busy_ms, gpu_busy_n = eng.pipeline_gpu_time(reset=True)
        e2e_s = max_over_ranks(time.perf_counter() - t0) * (args.steps / e2e_steps)      # normalised to K steps (e2e_steps of them were run)
        run_c, skip_c, end_c = wl.run_chunks - c0[0], wl.skipped_chunks - c0[1], wl.endpoints - c0[2]
        per_chunk_in = cfg.chunk_length * 2 + 4
        per_chunk_out = cfg.seg_rows * 4 * 2 + 3 * 4 + 4 * 256 + 8
        h2d, d2h = run_c * per_chunk_in // e2e_steps, run_c * per_chunk_out // e2e_steps
        extra["ragged"] = {"sessions": streams, "batch_assembly": "device gather from pinned rings" if dev_gather else "host gather + DMA",
                           "pipelining": ("one full-size tick per pass, gather + H2D pre-staged under the previous tick's kernels" if prestaged else "two ticks per pass, two in flight"),
                           "max_rows_per_tick": e2e_rows, "e2e_steps_run": e2e_steps,
                           "decoded_chunks_per_pass": run_c / e2e_steps, "vad_skipped_chunks_per_pass": skip_c / e2e_steps,
                           "endpoints_per_pass": end_c / e2e_steps,
                           "host_ms_per_tick": {"prestage + submit_tick (ready + gate + gather + enqueue)": 1e3 * (wl.t_submit - h0[0]) / max(1, wl.n_ticks - h0[3]),
                                                "collect_tick (wait + bookkeeping + rules)": 1e3 * (wl.t_collect - h0[1]) / max(1, wl.n_ticks - h0[3]),
                                                "scripted endpoints": 1e3 * (wl.t_after - h0[2]) / max(1, wl.n_ticks - h0[3]),
                                                "of which prestage (overlaps the running tick)": 1e3 * (getattr(wl, "t_prestage", 0.0) - h0[4]) / max(1, wl.n_ticks - h0[3])},
                           "gpu_busy_ms_per_pass": gpu_busy_ms / e2e_steps, "gpu_steps_per_pass": gpu_busy_n / e2e_steps,
                           "gpu_busy_counts": "first to last kernel of each pipelined step (asr_pipeline_gpu_time); the speculative fbank of the pre-staged chunks (~0.75 ms per pass at 4096 sessions) runs between steps and is not included",
                           "e2e_counts": "audio-seconds of the chunks actually decoded (VAD-skipped chunks are excluded from e2e.value)"}
        e2e_audio_per_step = run_c * (cfg.segment_length / cfg.sample_rate) / e2e_steps
    else:
        for _ in range(2):                                   # fill both pinned staging buffers once (the receive path writes here)
            view = eng.pinned_pcm(np.int16)
            view[:streams] = pcm
            eng.step(slots, view[:streams])
        barrier()
        for _ in range(args.steps):
            view = eng.pinned_pcm(np.int16)
            t1 = time.perf_counter()
            eng.step(slots, view[:streams])
            lat_ms.append(1e3 * (time.perf_counter() - t1))
        barrier()
        t0 = time.perf_counter()
        prev = None
        for _ in range(args.steps):
            view = eng.pinned_pcm(np.int16)
            tk = eng.submit(slots, view[:streams])
            if prev is not None:
                eng.collect(prev)
            prev = tk
        eng.collect(prev)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.result()

    value = world * args.steps * audio_per_step / (dev_ms / 1e3)
    e2e_value = world * args.steps * (e2e_audio_per_step if ragged else audio_per_step) / e2e_s

    # ---- other batch sizes on the same GPU (device-resident leg only; explains where the headline sits on the curve)
    sweep = []
    extra_legs = world == 1 and not fbank_only and not args.no_sweep
    if extra_legs:
        if ragged:
            wl.sch.close_scheduler()
        eng.close()
        for s_n in (256, 1024, 4096):
            if s_n == streams and not ragged:
                continue
            v_, ms_, _, _, _ = device_leg(local, blob, s_n, precision, low_latency, 5, 3)
            sweep.append({"streams": s_n, "ms_per_step": ms_, "audio_s_per_s": v_})

    if rank == 0:
        fam = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] / prof_steps} for k, v in prof.items() if v[1]}
        if fbank_only:
            t = fam["fbank"]["ms_per_step"] / fam["fbank"]["launches_per_step"]
            ach = alg_bytes / (t / 1e3) / 1e9
            roof = {"kernel": "fbank_kernel<256,int16,kaldi>", "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic_table().get("fbank1024", {}).get("fbank"), "peak_source": pk["source"],
                    "algorithmic_bytes_per_launch": alg_bytes}
        else:
            M = streams * cfg.rows
            gemm_flops = {"gemm_qkv": 2 * M * 1536 * 512, "gemm_out_proj": 2 * M * 512 * 512, "gemm_ffn1": 2 * M * 2048 * 512,
                          "gemm_ffn2": 2 * M * 512 * 2048}
            keys = cfg.rc_rows + cfg.left_context + cfg.seg_rows
            flop_sc = (20 * sum(gemm_flops.values()) // streams + 2 * cfg.frames * 128 * 128 + 2 * cfg.seg_rows * (512 * 512 + 512 * 804)
                       + 20 * 8 * 2 * 2 * cfg.rows * keys * 64)
            dom = max(gemm_flops, key=lambda k: fam[k]["ms_per_step"])
            t = fam[dom]["ms_per_step"] / fam[dom]["launches_per_step"]
            mult = 3 if precision == PRECISION_EXACT else 1
            ach = gemm_flops[dom] / (t / 1e3) / 1e12
            all_gemm_ms = sum(v["ms_per_step"] for k, v in fam.items() if k.startswith("gemm"))
            all_gemm_flop = streams * (flop_sc - 20 * 8 * 2 * 2 * cfg.rows * keys * 64)
            fused = streams >= 96 and dom in ("gemm_out_proj", "gemm_ffn2") and not os.environ.get("ASR_B200_NO_FUSED_LN")
            kname = "gemm_ln_kernel" if fused else "gemm_tc_kernel"
            dom_name = dom
            roof = {"kernel": f"{kname} ({dom}, M={M})", "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops"], "traffic": ncu_traffic(args.workload if not args.streams else "", dom_name), "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                    "algorithmic_flops_per_launch": gemm_flops[dom], "executed_flops_multiplier": mult,
                    "all_gemms": {"ms_per_step": all_gemm_ms, "tflops": all_gemm_flop / (all_gemm_ms / 1e3) / 1e12}}
            extra["path_roofline"] = {"flop_per_stream_chunk": flop_sc,
                                      "achieved_tflops": streams * flop_sc * world * args.steps / (dev_ms / 1e3) / 1e12,
                                      "frac_of_tensor_peak": streams * flop_sc * args.steps / (dev_ms / 1e3) / 1e12 / pk["bf16_tflops"]}
            extra["kernel_rooflines"] = family_rooflines(fam, cfg, streams, precision == PRECISION_EXACT, pk, args.workload if not args.streams else "")
        line = {
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == PRECISION_FAST else "bf16x3-split (fp32-equivalent)", "data": "synthetic",
            "config": {"workload": desc, "streams_per_gpu": streams, "chunk_ms": int(1000 * cfg.segment_length / cfg.sample_rate), "weights": f"random-init seed {WEIGHT_SEED}",
                       "l2": "no explicit flush: per-step working set (bf16 weights 128 MB + K/V rings %.0f MB + activations) exceeds the 126 MB L2"
                             % (streams * 1.97), "parallelism": f"sessions partitioned per GPU x{world}, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "kernel_families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in fam.items()},
            "chunk_latency_ms": ({"p50": float(np.percentile(lat_ms, 50)), "p99": float(np.percentile(lat_ms, 99)), "max": float(max(lat_ms)),
                                  "what": ("synchronous SessionScheduler.tick over every ready session: VAD gate + pinned gather + H2D + kernels + D2H + bookkeeping + endpoint rules"
                                           if ragged else "synchronous Engine.step of the whole batch from pinned host memory: H2D + kernels + D2H of ids")}
                                 if lat_ms else None),
            # real-time capacity: every stream needs one chunk per 640 ms; ticks of this batch size back to back
            "realtime_streams_per_gpu": (int(e2e_value / world) if not fbank_only else None),
        }
        line.update(extra)
        if extra_legs:
            line["stream_sweep"] = sweep
        if extra_legs and ragged:
            # the second half of the metric, measured directly: 10,240 real-time streams (the north-star count), per-chunk latency
            _, _, _, line["realtime_10240_streams"] = realtime_measure(local, rank, 10240, 6)
            # ---- the parity-conformant precision at the same size: EXACT = split-bf16 operands, three tensor-core passes per product,
            # token-exact against the reference (tests/test_parity_gpu.py); FAST is token-exact on peaked posteriors, see DESIGN.md
            ev, ems, efam, ecfg, _ = device_leg(local, blob, streams, PRECISION_EXACT, False, 5, 3, beam=10)
            eroof = family_rooflines(efam, ecfg, streams, True, pk, "exact4096")
            edom = max((k for k in eroof if eroof[k]["bound"] == "tensor"), key=lambda k: efam[k]["ms_per_step"])
            line["exact"] = {"value": ev, "unit": "audio-s/s", "ms_per_step": ems, "dtype": "bf16x3-split (fp32-equivalent), token-exact vs the reference",
                             "streams_per_gpu": streams, "prefix_beam": 10,
                             "roofline": dict(eroof[edom], kernel=edom, note="algorithmic FLOPs; the kernel executes 3x that on the tensor pipe"),
                             "path_frac_of_tensor_peak": streams * flop_sc / (ems / 1e3) / 1e12 / pk["bf16_tflops"],
                             "kernel_families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in efam.items()}}
            # ---- the other BASELINE configs, each on its own engine (device-resident; parity of every one is a GPU test)
            confs = {}
            v2, ms2, fam2, cfg2, _ = device_leg(local, blob, 256, PRECISION_FAST, False, 20, 5)
            confs["streams256"] = {"workload": WORKLOADS["streams256"][1], "value": v2, "unit": "audio-s/s", "ms_per_step": ms2,
                                   "kernel_families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in fam2.items()}}
            v4, ms4, fam4, cfg4, _ = device_leg(local, blob, 4096, PRECISION_FAST, True, 10, 3)
            r4 = family_rooflines(fam4, cfg4, 4096, False, pk, "lowlat4096")
            confs["lowlat4096"] = {"workload": WORKLOADS["lowlat4096"][1], "value": v4, "unit": "audio-s/s", "ms_per_step": ms4, "chunk_ms": 320,
                                   "kernel_rooflines": {k: {"frac": round(v["frac"], 4), "bound": v["bound"], "us_per_launch": round(v["us_per_launch"], 2)} for k, v in r4.items()}}
            confs["fbank1024"] = fbank_leg(local, blob, pk)
            line["configs"] = confs
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            if fbank_only:
                fb = fbank_cpu_baseline(threads)
                line["cpu_baseline"] = {"value": fb["kaldi80"], "unit": "audio-s/s", "cores": threads, "kind": "reference",
                                        "sample": f"torchaudio.compliance.kaldi.fbank(num_mel_bins=80, dither=0) on 64 streams x 640 ms, batch 1 per stream, ~3 s, {threads} torch threads",
                                        "melspec128_extract_filterbank": fb["melspec128"]}
            else:
                v, ms, _ = cpu_reference_run(steps=2, warmup=1, sample_streams=args.ref_streams, threads=threads)
                v2t, _, _ = cpu_reference_run(steps=2, warmup=1, sample_streams=4, threads=2)      # the reference's deployment setting (TORCH_THREAD=2, docker-compose.yml:22)
                line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                        "sample": f"{args.ref_streams} streams x 2 steps of the same workload, batch 1 per stream, {threads} torch threads",
                                        "value_2_threads": v2t}
                if extra_legs and ragged:
                    fb = fbank_cpu_baseline(threads)
                    line["configs"]["fbank1024"]["cpu_baseline"] = {"value": fb["kaldi80"], "unit": "audio-s/s", "cores": threads, "kind": "reference",
                                                                    "sample": "torchaudio.compliance.kaldi.fbank on 64 streams x 640 ms, batch 1 per stream, ~3 s",
                                                                    "melspec128_extract_filterbank": fb["melspec128"]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        torch.distributed.destroy_process_group()
    eng.close()                                              # (idempotent)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ragged4096", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact"])
    ap.add_argument("--ref-streams", type=int, default=16, help="streams per step in the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the extra legs of the default line (batch-size sweep, real-time leg, exact leg, configs legs)")
    ap.add_argument("--long-chunks", type=int, default=1875, help="chunks per stream of the longform workload (1875 x 320 ms = 10 minutes)")
    args = ap.parse_args()
    if args.impl == "reference":
        sys.exit(run_reference_arm(args))
    if args.workload == "longform":
        sys.exit(run_longform(args))
    sys.exit(run_realtime(args) if args.workload == "realtime" else run_ours(args))


if __name__ == "__main__":
    main()
