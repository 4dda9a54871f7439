#!/usr/bin/env python3
"""bench.py -- throughput of the post-parse signal path on B200 (BASELINE.json metric: channel-samples/s).

Headline workload at N=1 = BASELINE.json configs[1]: batched IMDCT + windowing/overlap-add, 1024-sample frames,
256 streams x 24 channels (22.2), synthetic spectra, fp32.  One "step" = one frame for every stream of one
batch = 6144 channel-frames through mpeghb200_fd_synth (one kernel launch).

  value      device-resident inputs, CUDA-event timed over K back-to-back steps.  To keep every step's
             16 KB/unit of traffic in HBM and not in the 126 MB L2, the timed loop rotates over G stream
             groups (each with its own spectra, overlap state and output: 3 x 25 MB), G*75 MB >> L2.
  e2e        the same step through the host-buffer C ABI (mpeghb200_fd_submit / _wait, two frames in flight; the
             synchronous mpeghb200_fd_execute is reported beside it): pinned host spectra -> device, kernel,
             PCM -> host, every step, all inside the timed region.
  roofline   algorithmic bytes (16384 B per channel-frame) / mean step duration, vs the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the unmodified reference's ia_core_coder_fd_frm_dec (oracle/_ref, kind "reference"; the
             plain-C port if that build is absent) on one host core, bounded sample.
  configs    the other BASELINE.json configurations, each as ONE device-resident chain behind one host-buffer call
             (mpeghb200_chain_*, include/mpeghb200.h): "hoa" = configs[2], "binaural" = configs[3], "objects" =
             configs[4] (512 streams per GPU = 4096 over 8), "fd_4096" = configs[1]'s shape at the north-star stream
             count.  Each carries value (resident inputs, mpeghb200_chain_run_dev), roofline (algorithmic bytes of
             the whole chain unit / step time of ALL its kernels), e2e (mpeghb200_chain_submit / _wait from pinned
             host buffers: core signals up, packed PCM down) and cpu_baseline (the unmodified reference's stage
             functions chained in C on one pinned host thread, oracle/ref_chain.c).
  --impl reference   the reference arm: the same CPU functions with one persistent pinned worker thread per host
             core (oracle/ref_chain.c), headline workload first, then the configs.

Multi-GPU (torchrun, one rank per GPU): streams are sharded by rank, no collective on the data path
(weak scaling: the per-GPU stream counts above on every rank); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")

STREAMS, CHANNELS, FRAME = 256, 24, 1024
UNITS = STREAMS * CHANNELS
BYTES_PER_UNIT = 16384  # coef 4096 + overlap r/w 2*4096 + out 4096
FRAMES_PER_LAUNCH = 8  # a step = this many consecutive frames of every stream, one launch
METRIC = "channel-samples/s (FD synthesis: IMDCT+window+OLA, 256 streams x 24 ch x 1024)"
UNIT = "channel-samples/s"
CHAIN_CONFIGS = ("hoa", "binaural", "objects", "fd_4096")


def set_streams(n, frames=None):
    global STREAMS, UNITS, METRIC, FRAMES_PER_LAUNCH
    STREAMS, UNITS = n, n * CHANNELS
    if frames:
        FRAMES_PER_LAUNCH = max(1, frames)
    METRIC = f"channel-samples/s (FD synthesis: IMDCT+window+OLA, {n} streams x 24 ch x 1024)"


def workload_name():
    tag = "configs[1]" if STREAMS == 256 else "configs[1] shape at a different stream count"
    return (f"{tag}: IMDCT+windowing/OLA, 1024-sample frames, {STREAMS} streams x 24 ch, synthetic spectra, "
            "ONLY_LONG with random sine/KBD shapes")


def config_dict():
    """Identical in both arms (the driver compares the two lines' config)."""
    return {"workload": workload_name(), "streams_per_gpu": STREAMS, "channels": CHANNELS, "frame": FRAME,
            "frames_per_step": FRAMES_PER_LAUNCH, "sharding": "streams by rank, no collective"}


def make_side(rng, n_units, mixed=False):
    """SURVEY.md section 8(d) config 2.  Baseline: ONLY_LONG with random sine/KBD shapes.  mixed=True: the
    mixed-sequence variant, ~8 % LONG_START / EIGHT_SHORT / LONG_STOP frames."""
    seq = np.zeros(n_units, np.int32)
    r = rng.random(n_units)
    if not mixed:
        r[:] = 0.0
    seq[r > 0.92] = 1
    seq[r > 0.95] = 2
    seq[r > 0.98] = 3
    shape = (rng.random(n_units) < 0.5).astype(np.int32)
    shape_prev = (rng.random(n_units) < 0.5).astype(np.int32)
    z = np.zeros(n_units, np.int32)
    return np.stack([seq, shape, shape_prev, z, z], axis=-1)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe).

    nvidia-smi needs a few hundred ms to deliver its first row, so the caller starts it before the warm-up,
    keeps the same kernel running until the first row has arrived (wait_first), and marks the timed region with
    mark_begin/mark_end; rows are time-stamped on arrival.  `samples` counts the rows that arrived inside the
    timed region, `samples_under_load` all rows taken while the identical load was running (warm-up tail,
    timed region and, if the region was too short for 3 rows, an untimed continuation of the same load).
    """

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def n_rows(self):
        return len(self.rows)

    def alive(self):
        return self.proc is not None and self.proc.poll() is None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, inside = [], [], set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
                if self.t_begin is not None and self.t_begin <= ts <= (self.t_end or ts):
                    inside += 1
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": inside, "samples_under_load": len(sm), "reasons": sorted(reasons)}


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# =====================================================================================================================
# The BASELINE configurations as chains: shapes, synthetic inputs, byte accounting.  Used by both arms.
# =====================================================================================================================
class ChainSpec:
    """One BASELINE configuration: what the GPU chain and the CPU reference chain are built from."""

    def __init__(self, name, S):
        self.name, self.S = name, S
        import bin_cases  # noqa: F401  (tests/ fixtures: seeded synthetic inputs shared with the parity tests)
        import hoa_cases
        import hoasp_cases
        if name == "hoa":
            self.hoasp_cfg, self.ren_cfg = hoasp_cases.CONFIGS[0], hoa_cases.CONFIGS[0]
            self.n_in, self.n_out, self.fpr = 16, 24, 1
            self.workload = ("configs[2]: HOA order 4 (25 coeffs), 16 transport channels -> spatial decode (4-ish vector, "
                             "2-ish direction signals, ambience) -> NFC -> render matrix to 22.2 -> limiter -> 24-bit PCM, "
                             f"{S} streams")
            # per stream-frame: transport in + side record, PCM out, limiter delay line + history r/w
            self.bytes_per_unit = 16 * 4096 + hoasp_cases.SIDE_DTYPE.itemsize + 24 * 1024 * 3 + 2 * 240 * 25 * 4
            self.bytes_formula = "T*4096 + side 3.4 KB + 24*1024*3 PCM + limiter state r/w 2*240*25*4"
            self.value_samples = 24 * 1024      # rendered output channel-samples per unit
            self.value_what = "rendered output channel-samples/s (24 x 1024 per stream-frame)"
            self.h2d = 16 * 4096 + hoasp_cases.SIDE_DTYPE.itemsize
            self.d2h = 24 * 1024 * 3 + 4
        elif name == "binaural":
            self.n_in, self.n_out, self.fpr, self.taps = 24, 2, 4, 4096
            self.workload = ("configs[3]: binaural rendering 22.2 -> 2 ears, synthetic 4096-tap BRIRs (FFT 8192, "
                             f"per-stream filter sets), block overlap-add fast convolution -> 24-bit PCM, {S} streams; "
                             "one step = one 4096-sample block = 4 frames")
            self.bytes_per_unit = 24 * 16384 + 2 * 24 * 8192 * 4 + 2 * 2 * 16384 + 2 * 4096 * 3
            self.bytes_formula = "in 24*16384 + per-stream filter spectra 2*24*8192*4 + OLA tail r/w 2*2*16384 + PCM 2*4096*3"
            self.value_samples = 24 * 4096
            self.value_what = "input channel-samples/s (24 x 4096 per stream-block)"
            self.h2d = 24 * 16384
            self.d2h = 2 * 4096 * 3 + 4
        elif name == "objects":
            self.n_in, self.n_out, self.fpr = 32, 12, 1
            self.workload = ("configs[4]: 32 objects, positions changing every frame -> VBAP gains + ramped mix to 7.1.4 "
                             f"-> limiter -> 24-bit PCM, {S} streams per GPU (4096 streams over 8 GPUs)")
            self.bytes_per_unit = 32 * 4096 + 32 * 64 + 12 * 1024 * 3 + 2 * 240 * 13 * 4 + 2 * 32 * 12 * 4
            self.bytes_formula = "objects 32*4096 + side 32*64 + PCM 12*1024*3 + limiter state r/w 2*240*13*4 + gains r/w"
            self.value_samples = 32 * 1024
            self.value_what = "object-samples/s (32 x 1024 per stream-frame)"
            self.h2d = 32 * 4096 + 32 * 64
            self.d2h = 12 * 1024 * 3 + 4
        elif name == "fd_4096":
            self.n_in, self.n_out, self.fpr = 24, 24, 1
            self.workload = (f"configs[1] shape at the north-star stream count: {S} streams x 24 ch, spectra -> FD "
                             "synthesis (value: that kernel alone, up to 4 frames of every stream per launch; e2e: -> limiter "
                             "-> 24-bit PCM as one chain)")
            self.bytes_per_unit = 24 * BYTES_PER_UNIT
            self.bytes_formula = ("24 channel-frames x (coef 4096 + overlap r/w 8192 + out 4096), SURVEY.md 8(d)'s "
                                  "figure; with F frames per launch the overlap stays on the SM between frames "
                                  "(own minimum 8192 + 8192 / F per channel-frame)")
            self.value_samples = 24 * 1024
            self.value_what = "channel-samples/s (24 x 1024 per stream-frame)"
            self.h2d = 24 * 4096 + 24 * 4
            self.d2h = 24 * 1024 * 3 + 4
        else:
            raise ValueError(name)

    # ---- seeded synthetic inputs (numpy, shared by both arms) ----
    def hoa_sides(self, n_streams, n_frames, seed=2345):
        import hoasp_cases
        gens = [hoasp_cases.SideGen(self.hoasp_cfg, np.random.default_rng(seed + s), pred_prob=0.0)
                for s in range(min(n_streams, 64))]
        out = np.zeros((n_frames, n_streams), hoasp_cases.SIDE_DTYPE)
        for f in range(n_frames):
            a = np.array([g.next() for g in gens], hoasp_cases.SIDE_DTYPE)
            out[f] = np.tile(a, (n_streams + len(gens) - 1) // len(gens))[:n_streams]
        return out

    def obj_params(self, n_streams, n_frames, seed=4567):
        import obj_cases
        rng = np.random.default_rng(seed)
        return obj_cases.random_params(rng, n_frames, n_streams * 32, spread_prob=0.0).reshape(n_frames, n_streams, 32, 7)

    def brirs(self, n_sets, seed=3456):
        rng = np.random.default_rng(seed)
        L = self.taps
        env = np.exp(-6.0 * np.arange(L) / L).astype(np.float32)
        td = rng.standard_normal((n_sets, 2, 24, L), dtype=np.float32) * np.float32(0.2) * env
        return {"taps_direct": td, "taps_diffuse": np.zeros((n_sets, 0, 2, L), np.float32),
                "direct_fc": np.ones((n_sets, 2, 24), np.float32), "diffuse_fc": np.zeros((n_sets, 2, 0), np.float32),
                "weight": np.ones((n_sets, 24), np.float32)}


def cpu_chain_baseline(spec, n_threads, sample_streams, n_frames, repeats=1):
    """The unmodified reference's stage functions chained in C (oracle/ref_chain.c) on n_threads pinned workers over a
    bounded sample of the configuration.  Returns (value in spec's unit, seconds per sample, kind, sample text)."""
    from oracle import loader
    if loader.load_ref_ns() is None:
        if spec.name != "fd":
            return None
        # oracle/_ref did not travel: the plain-C port of the FD synthesis (oracle/oracle_fd.c), one thread
        lib = loader.load_oracle()
        U, F_ = sample_streams * 24, n_frames
        rng = np.random.default_rng(99)
        coef = rng.normal(0, 500, (F_, U, FRAME)).astype(np.float32)
        side = np.ascontiguousarray(make_side(rng, F_ * U).reshape(F_, U, 5))
        ov, out = np.zeros((U, FRAME), np.float32), np.zeros((F_, U, FRAME), np.float32)
        t0 = time.perf_counter()
        assert lib.oracle_fd_frm_dec_batch(coef, ov, out, side, F_, U) == 0
        t = time.perf_counter() - t0
        return {"value": F_ * U * FRAME / t, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"{U} channel-frames x {F_} frames on 1 host thread; oracle_fd_frm_dec (plain-C port)"}
    S = sample_streams
    rng = np.random.default_rng(99)
    name = spec.name
    side = None
    if name == "hoa":
        import hoasp_cases
        ch = loader.RefChain("hoa", S, n_threads, cfg=spec.hoasp_cfg, ren=spec.ren_cfg,
                             side_bytes=hoasp_cases.SIDE_DTYPE.itemsize, start_delay=240)
        x = rng.normal(0, 2000.0, size=(n_frames, S, 16, 1024)).astype(np.float32)
        side = spec.hoa_sides(S, n_frames)
        what = "impeghd_hoa_spatial_process stages + impeghd_hoa_ren_block_process (NFC) + limiter + samples_sat"
    elif name == "binaural":
        f = spec.brirs(min(S, 4))
        filt = [{k: v[i] for k, v in f.items()} for i in range(f["taps_direct"].shape[0])]
        ch = loader.RefChain("bin", S, n_threads, filters=filt, sets=np.arange(S) % len(filt))
        x = rng.normal(0, 3000.0, size=(n_frames, S, 24, 1024)).astype(np.float32)
        what = "impeghd_binaural_struct_process_ld_ola + samples_sat"
    elif name == "objects":
        ch = loader.RefChain("obj", S, n_threads, n_obj=32, cicp=19, n_spk=12, start_delay=240)
        x = rng.normal(0, 3000.0, size=(n_frames, S, 32, 1024)).astype(np.float32)
        side = spec.obj_params(S, n_frames)
        what = "impeghd_obj_renderer_dec + limiter + samples_sat"
    else:
        tail = 0 if name == "fd" else 1
        ch = loader.RefChain("fd", S, n_threads, n_ch=24, tail=tail, start_delay=240)
        x = rng.normal(0, 500.0, size=(n_frames, S, 24, 1024)).astype(np.float32)
        side = np.ascontiguousarray(make_side(rng, n_frames * S * 24).reshape(n_frames, S, 24, 5))
        what = "ia_core_coder_fd_frm_dec" + (" + limiter + samples_sat" if tail else "")
    ch.run(x[:max(spec.fpr, 1)], None if side is None else side[:max(spec.fpr, 1)], want_pcm=False)  # warm-up
    t = 0.0
    for _ in range(repeats):
        dt, _, _ = ch.run(x, side, want_pcm=False)
        t += dt
    used = ch.n_threads
    ch.close()
    units = repeats * S * n_frames / spec.fpr
    return {"value": units * spec.value_samples / t, "unit": UNIT, "cores": used, "kind": "reference",
            "sample": f"{S} streams x {n_frames * repeats} frames on {used} pinned host thread(s) "
                      f"(box has {cpu_threads()}); {what}",
            "ms_per_unit": t / units * 1e3, "audio_s_per_wall_s": S * n_frames * repeats * 1024 / 48000.0 / t}


def cpu_testbench_baseline(n_proc, repeats=8, stream_name="sine_1khz_000_cicp1.mhas", channels=1):
    """BASELINE.json configs[0]: the UNMODIFIED reference testbench (oracle/_ref/ia_mpeghd_testbench, built from
    /root/reference by oracle/Makefile) decoding smoke_test_suite/inp/sine_1khz_000_cicp1.mhas on the host cores, one
    process per core, `repeats` decodes each.  None when the build did not travel."""
    tb = os.path.join(ROOT, "oracle", "_ref", "ia_mpeghd_testbench")
    stream = os.path.join(ROOT, "oracle", "_ref", "smoke", stream_name)
    if not (os.path.exists(tb) and os.path.exists(stream)):
        return None
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        cmd = "for i in $(seq %d); do %s -ifile:%s -ofile:%s/o_$$.wav > /dev/null 2>&1 || exit 1; done" % (
            repeats, tb, stream, tmp)
        subprocess.run(["sh", "-c", cmd.replace("seq %d" % repeats, "seq 1")], check=False)  # page the binary in
        t0 = time.perf_counter()
        procs = [subprocess.Popen(["sh", "-c", cmd]) for _ in range(n_proc)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
    if any(rcs):
        return None
    audio_s = n_proc * repeats * 469 * 1024 / 48000.0   # 469 output frames of one channel per decode
    return {"workload": f"configs[0]: {stream_name} through the unmodified ia_mpeghd_testbench (parse + "
                        "signal path + WAV write), one process per host core",
            "processes": n_proc, "decodes_per_process": repeats, "wall_s": dt, "audio_s_per_wall_s": audio_s / dt,
            "value": audio_s * 48000.0 * channels / dt, "unit": UNIT, "kind": "reference"}


# bounded CPU samples: (streams per thread, frames) sized for about a second of work per thread
CPU_SAMPLE = {"hoa": (2, 24), "binaural": (1, 32), "objects": (4, 48), "fd_4096": (4, 48), "fd": (16, 200)}


def reference_arm(args, rank):
    """--impl reference: the unmodified reference's CPU functions, one persistent pinned worker per host core."""
    if rank != 0:
        return
    from oracle import loader
    n_threads = cpu_threads()
    spec = ChainSpec("fd_4096", STREAMS)
    spec.name = "fd"  # headline: the FD synthesis stage alone, like the b200 arm's value
    if loader.load_ref_ns() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmpeghref_ns.so not built"}), flush=True)
        return
    # A step = one frame of the full 6144 channel-frames unless K + W such steps would exceed ~90 s on this box; then
    # each step is a proportionally smaller slice of whole streams (stated in cpu_baseline.sample).
    ch = loader.RefChain("fd", STREAMS, n_threads, n_ch=CHANNELS, tail=0)
    rng = np.random.default_rng(1234)
    x = rng.normal(0, 500.0, size=(1, STREAMS, CHANNELS, FRAME)).astype(np.float32)
    side = np.ascontiguousarray(make_side(rng, UNITS).reshape(1, STREAMS, CHANNELS, 5))
    ch.run(x, side, want_pcm=False)
    t_full, _, _ = ch.run(x, side, want_pcm=False)
    sample_streams = STREAMS
    total = t_full * (args.steps + args.warmup)
    if total > 90.0:
        sample_streams = max(n_threads, int(STREAMS * 90.0 / total))
        ch.close()
        ch = loader.RefChain("fd", sample_streams, n_threads, n_ch=CHANNELS, tail=0)
        x, side = x[:, :sample_streams].copy(), side[:, :sample_streams].copy()
    for _ in range(args.warmup):
        ch.run(x, side, want_pcm=False)
    t = 0.0
    for _ in range(args.steps):
        dt, _, _ = ch.run(x, side, want_pcm=False)
        t += dt
    used = ch.n_threads
    ch.close()
    sps = t / args.steps
    value = sample_streams * CHANNELS * FRAME / sps
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
        "realtime_factor": value / CHANNELS / 48000.0,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "reference",
                         "sample": f"one frame of {sample_streams} of the workload's {STREAMS} streams x 24 ch per step "
                                   f"(the GPU arm's step hands {FRAMES_PER_LAUNCH} frames to one launch; a CPU core "
                                   "walks them one by one either way), "
                                   f"{args.steps} steps, ia_core_coder_fd_frm_dec per channel-frame on {used} "
                                   "persistent pinned worker threads (oracle/ref_chain.c), timed inside C"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["configs0_cpu_testbench"] = cpu_testbench_baseline(n_threads)
    if not args.no_configs:
        line["configs"] = {}
        for name in CHAIN_CONFIGS:
            per, frames = CPU_SAMPLE[name]
            cs = ChainSpec(name, per * n_threads)
            r = cpu_chain_baseline(cs, n_threads, per * n_threads, frames)
            line["configs"][name] = {"workload": cs.workload, "value": r["value"], "unit": UNIT,
                                     "value_is": cs.value_what, "cpu_baseline": r,
                                     "audio_s_per_wall_s": r["audio_s_per_wall_s"]}
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
def bench_chain(lib, spec, args, rank, world, dist, torch, peak):
    """One BASELINE configuration as a chain on this rank's shard: device-resident value, roofline, e2e."""
    import ctypes
    import hoasp_cases
    from oracle import loader as fixtures  # layout / matrix fixtures only (tests/golden), nothing is computed there
    S, name = spec.S, spec.name
    st = torch.cuda.current_stream()
    handles = []
    G = 2  # input groups rotated through; the chain's intermediates alone exceed the L2 for every configuration
    kw = dict(pcm_bits=24, limiter=1, start_delay=240, max_frames_per_call=spec.fpr)
    rng = np.random.default_rng(1000 + rank)
    side_d = side_h = None
    if name == "hoa":
        g, gr = np.load(os.path.join(GOLD, "hoasp_golden.npz")), np.load(os.path.join(GOLD, "hoa_ren_golden.npz"))
        import hoa_cases
        k = hoa_cases.key(spec.ren_cfg)
        hsp = lib.hoa_spatial_create(spec.hoasp_cfg, hoasp_cases.golden_tables(g, spec.hoasp_cfg))
        hren = lib.hoa_renderer_create(gr[k + "_M"], gr[k + "_nfc"])
        handles = [(lib.hoa_renderer_destroy, hren), (lib.hoa_spatial_destroy, hsp)]
        chain = lib.chain_open(S, n_hoa_transport=16, hoa_spatial=hsp, hoa_renderer=hren, **kw)
        sides = spec.hoa_sides(S, 4, seed=2345 + 100 * rank)
        side_h = [np.ascontiguousarray(sides[i]) for i in range(4)]
        side_key = "hoa_side"
        sigma = 2000.0
    elif name == "binaural":
        f = spec.brirs(S, seed=3456 + rank)
        hb = lib.binaural_create(f["taps_direct"], f["taps_diffuse"], f["direct_fc"], f["diffuse_fc"], f["weight"])
        del f
        handles = [(lib.binaural_destroy, hb)]
        chain = lib.chain_open(S, n_bed_ch=24, binaural=hb, binaural_set_of_stream=np.arange(S, dtype=np.int32), **kw)
        side_key = None
        sigma = 3000.0
    elif name == "objects":
        lay = lib.obj_layout_create(fixtures.obj_layout_golden(19))
        handles = [(lib.obj_layout_destroy, lay)]
        chain = lib.chain_open(S, n_obj=32, obj_layout=lay, **kw)
        p = spec.obj_params(S, 2, seed=4567 + rank)
        side_h = [lib.obj_side_from_polar(p[i]) for i in range(2)]
        side_key = "obj_side"
        sigma = 3000.0
    else:  # fd_4096
        chain = lib.chain_open(S, core_is_spectra=1, n_bed_ch=24, **kw)
        import fd_cases
        side_h = [np.ascontiguousarray(fd_cases.pack_side(make_side(rng, S * 24)).astype(np.uint32)) for _ in range(2)]
        side_key = "fd_side"
        sigma = 500.0
    F = spec.fpr
    gen = torch.Generator(device="cuda").manual_seed(4321 + rank)
    core_d = [torch.randn(F, S, spec.n_in, 1024, device="cuda", generator=gen) * sigma for _ in range(G)]
    if side_h is not None:
        side_d = [torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda() for a in side_h]
    pcm_d = torch.empty(S, chain.record_bytes, dtype=torch.uint8, device="cuda")
    nb_d = torch.zeros(S, dtype=torch.int32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: resident inputs ----
    if name == "fd_4096":
        # the FD synthesis kernel alone at 4096 streams (north-star stream count), state per group
        FDF = min(FRAMES_PER_LAUNCH, 4)  # frames of every stream per launch (403 MB of spectra per frame)
        ovl = [torch.zeros(S * 24, 1024, device="cuda") for _ in range(G)]
        outb = torch.empty(FDF, S * 24, 1024, device="cuda")
        core_f = [torch.randn(FDF, S * 24, 1024, device="cuda", generator=gen) * sigma for _ in range(G)]
        side_f = [torch.from_numpy(np.stack([fd_cases.pack_side(make_side(rng, S * 24)) for _ in range(FDF)])
                                   .astype(np.int32)).cuda() for _ in range(G)]

        def step(i):
            g_ = i % G
            lib.fd_synth_frames_dev(core_f[g_].data_ptr(), side_f[g_].data_ptr(), ovl[g_].data_ptr(), outb.data_ptr(),
                                    S * 24, FDF, st.cuda_stream)
    else:
        frames = []
        for i in range(max(G, len(side_d) if side_d else 1) * 2):
            kwf = {side_key: side_d[i % len(side_d)].data_ptr()} if side_key else {}
            frames.append(chain.frame(core=core_d[i % G].data_ptr(), pcm=pcm_d.data_ptr(), pcm_bytes=nb_d.data_ptr(),
                                      **kwf))

        def step(i):
            chain.run_dev(frames[i % len(frames)], F, st.cuda_stream)

    steps, warm = args.chain_steps, max(3, args.chain_steps // 10)
    for i in range(warm):
        step(i)
    barrier()
    l0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for i in range(steps):
        step(warm + i)
    e1.record(st)
    barrier()
    launches = lib.launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    value_frames = 1
    if name == "fd_4096":
        del ovl, outb, core_f, side_f
        value_frames = FDF

    # ---- e2e: pinned host buffers through submit / wait, two calls in flight ----
    n_host = 2
    h_core = [torch.empty(F, S, spec.n_in, 1024, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    for h in h_core:
        h.normal_(0, sigma)
    h_pcm = [torch.empty(S, chain.record_bytes, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
    h_nb = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(n_host)]
    h_side = None
    if side_h is not None:
        h_side = []
        for a in side_h[:2]:
            b = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
            b.copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
            h_side.append(b)
    hframes = []
    for k_ in range(n_host):
        kwf = {side_key: h_side[k_ % len(h_side)].data_ptr()} if side_key else {}
        hframes.append(chain.frame(core=h_core[k_].data_ptr(), pcm=h_pcm[k_].data_ptr(), pcm_bytes=h_nb[k_].data_ptr(),
                                   **kwf))

    def run_e2e(n):
        chain.submit(hframes[0], F)
        for i in range(1, n):
            chain.submit(hframes[i % n_host], F)
            chain.wait()
        chain.wait()

    run_e2e(3)
    barrier()
    n_e2e = args.chain_e2e_steps
    t0 = time.perf_counter()
    run_e2e(n_e2e)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert int(h_nb[0][0]) > 0, "chain produced no PCM"
    chain.close()
    for fn, h in handles:
        fn(h)
    del core_d, h_core, h_pcm

    units = S * value_frames  # per step and rank: stream-frames (stream-blocks for the binaural chain)
    achieved = spec.bytes_per_unit * units / (ms * 1e-3) / 1e9
    value = world * units * spec.value_samples / (ms * 1e-3)
    h2d, d2h = spec.h2d * S, spec.d2h * S
    return {
        "workload": spec.workload, "streams_per_gpu": S, "value": value, "unit": UNIT, "value_is": spec.value_what,
        "ms_per_step": ms, "steps": steps, "units_per_step_per_gpu": units,
        "audio_s_per_wall_s": world * S * F * value_frames * 1024 / 48000.0 / (ms * 1e-3),
        "gpu_launches": int(launches), "launches_per_step": launches / steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "scope": "all kernels of the chain step (not only the dominant one)",
                     "algorithmic_bytes_per_unit": spec.bytes_per_unit, "bytes_formula": spec.bytes_formula},
        "value_frames_per_launch": value_frames,
        "e2e": {"value": world * S * spec.value_samples / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_GBps": h2d / e2e_s / 1e9, "d2h_GBps": d2h / e2e_s / 1e9,
                "audio_s_per_wall_s": world * S * F * 1024 / 48000.0 / e2e_s,
                "api": "mpeghb200_chain_submit / mpeghb200_chain_wait, two calls in flight, pinned host buffers: core "
                       "signals + side records up, packed 24-bit PCM down"},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--groups", type=int, default=8, help="stream groups rotated through (working set > L2)")
    ap.add_argument("--frames-per-launch", type=int, default=FRAMES_PER_LAUNCH,
                    help="consecutive frames of every stream handed to one launch (mpeghb200_fd_synth_frames_dev)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip the chain configurations)")
    ap.add_argument("--only", default="", help="comma-separated subset of: " + ",".join(CHAIN_CONFIGS))
    ap.add_argument("--chain-steps", type=int, default=100)
    ap.add_argument("--chain-e2e-steps", type=int, default=30)
    ap.add_argument("--streams", type=int, default=STREAMS,
                    help="streams per GPU (default 256 = BASELINE.json configs[1]; 4096 = the north-star stream count)")
    args = ap.parse_args()
    set_streams(args.streams, args.frames_per_launch)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from libmpegh_b200.lib import Lib
    import fd_cases

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    # the host side of a rank lives next to its GPU: CPU affinity first, so that the pinned staging buffers
    # allocated below land on that NUMA node (first touch)
    from libmpegh_b200 import shard
    numa_cpus = shard.bind_to_gpu_numa(local_rank) if world > 1 else None
    # Libraries print banners on stdout (NCCL's version line under NCCL_DEBUG=VERSION/WARN does); the contract is ONE
    # JSON line there.  File descriptor 1 is pointed at stderr for the duration of the run and the line is written
    # to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = Lib().create(local_rank)
    stream = torch.cuda.current_stream()
    # weak scaling: world x 256 streams in total, rank r owns the contiguous block shard.stream_range gives it
    s_begin, s_end = shard.stream_range(rank, world, world * STREAMS)
    assert s_end - s_begin == STREAMS

    # ---- device-resident synthetic inputs: G groups of 256 streams (this rank's shard) ----------
    G = max(1, args.groups)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    rng = np.random.default_rng(1234 + rank)
    FPL = FRAMES_PER_LAUNCH
    GF = max(2, -(-G // FPL))  # groups for the FPL-frame launch: GF * FPL * 50 MB of spectra + PCM, >> L2 as well
    coef = [torch.randn(UNITS, FRAME, device="cuda", generator=gen) * 500.0 for _ in range(max(G, GF * FPL))]
    ovl = [torch.zeros(UNITS, FRAME, device="cuda") for _ in range(G)]
    outb = [torch.empty(UNITS, FRAME, device="cuda") for _ in range(G)]
    # the FPL-frame launch reads [FPL][UNITS][1024] blocks
    coef_f = [torch.stack(coef[g * FPL:(g + 1) * FPL]) for g in range(GF)] if FPL > 1 else None
    outb_f = [torch.empty(FPL, UNITS, FRAME, device="cuda") for _ in range(GF)] if FPL > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mixed, with_clocks, frames):
        """W warm-up + K timed steps of one workload variant; returns (ms_per_step, launches, clocks).  frames = 1:
        a step is one frame of every stream (mpeghb200_fd_synth_dev); frames > 1: that many consecutive frames of
        every stream in one launch (mpeghb200_fd_synth_frames_dev)."""
        n_g = G if frames == 1 else GF
        side = [torch.from_numpy(np.stack([fd_cases.pack_side(make_side(rng, UNITS, mixed)) for _ in range(frames)])
                                 .astype(np.int32)).cuda() for _ in range(n_g)]

        def step(i):
            g = i % n_g
            if frames == 1:
                lib.fd_synth_dev(coef[g].data_ptr(), side[g].data_ptr(), ovl[g].data_ptr(), outb[g].data_ptr(), UNITS,
                                 stream.cuda_stream)
            else:
                lib.fd_synth_frames_dev(coef_f[g].data_ptr(), side[g].data_ptr(), ovl[g].data_ptr(),
                                        outb_f[g].data_ptr(), UNITS, frames, stream.cuda_stream)

        sampler = ClockSampler(local_rank)
        sampling = rank == 0 and with_clocks
        if sampling:
            sampler.start()
        for i in range(args.warmup):
            step(i)
        n_extra = 0
        if sampling:
            # same load, untimed, until nvidia-smi has delivered its first row (bounded: 5 s)
            t_wait = time.perf_counter()
            while sampler.alive() and sampler.n_rows() == 0 and time.perf_counter() - t_wait < 5.0:
                for _ in range(64):
                    step(n_extra)
                    n_extra += 1
                torch.cuda.synchronize()
        barrier()
        launches0 = lib.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.mark_begin()
        ev0.record(stream)
        for i in range(args.steps):
            step(args.warmup + i)
        ev1.record(stream)
        barrier()
        sampler.mark_end()
        ms = ev0.elapsed_time(ev1)
        n_launch = lib.launch_count() - launches0
        clk = None
        if sampling:
            # a timed region shorter than three sampling periods: continue the identical load, untimed
            t_wait = time.perf_counter()
            while sampler.alive() and sampler.n_rows() < 3 and time.perf_counter() - t_wait < 3.0:
                for _ in range(64):
                    step(n_extra)
                    n_extra += 1
                torch.cuda.synchronize()
            clk = sampler.stop()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, n_launch, clk

    ms_per_step, launches, clocks = timed(False, True, FPL)
    value = world * UNITS * FPL * FRAME / (ms_per_step * 1e-3)
    ms_mixed, _, _ = timed(True, False, FPL)
    single = None
    if FPL > 1:  # the same workload one frame per launch (round 1's step)
        k_keep = args.steps
        args.steps = max(1, min(args.steps * FPL, 20000))
        ms_1, _, _ = timed(False, False, 1)
        ms_1m, _, _ = timed(True, False, 1)
        args.steps = k_keep
        single = (ms_1, ms_1m)
    del coef_f, outb_f

    # ---- e2e: host buffers through the session API, copies inside the timed region --------------
    sess = lib.fd_open(STREAMS, CHANNELS)
    n_host = 4
    h_coef = [torch.empty(STREAMS, CHANNELS, FRAME, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    for h in h_coef:
        h.copy_(torch.from_numpy(rng.normal(0, 500, (STREAMS, CHANNELS, FRAME)).astype(np.float32)))
    h_out = [torch.empty(STREAMS, CHANNELS, FRAME, dtype=torch.float32).pin_memory() for _ in range(n_host)]
    h_side = np.ascontiguousarray(fd_cases.pack_side(make_side(rng, UNITS)).reshape(STREAMS, CHANNELS))

    def e2e_submit(i):
        k = i % n_host
        lib.check(lib.c.mpeghb200_fd_submit(sess.h, h_coef[k].data_ptr(), h_side.ctypes.data, h_out[k].data_ptr()))

    def e2e_wait():
        lib.check(lib.c.mpeghb200_fd_wait(sess.h))

    # two frames in flight: frame i is submitted, then frame i - 1 is waited for (its PCM is in host memory);
    # every frame's spectra go host -> device and its PCM device -> host inside the timed region
    e2e_submit(0)
    for i in range(1, 4):
        e2e_submit(i)
        e2e_wait()
    e2e_wait()
    barrier()
    t0 = time.perf_counter()
    e2e_submit(0)
    for i in range(1, args.e2e_steps):
        e2e_submit(i)
        e2e_wait()
    e2e_wait()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    # the same through the synchronous call (one frame at a time: upload, kernel and download of a frame only
    # overlap chunk-wise inside the call)
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        k = i % n_host
        lib.check(lib.c.mpeghb200_fd_execute(sess.h, h_coef[k].data_ptr(), h_side.ctypes.data, h_out[k].data_ptr()))
    e2e_sync_s = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * UNITS * FRAME / e2e_s
    sess.close()
    del coef, ovl, outb, h_coef, h_out
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, one chain each ----------------------------------------
    peak, peak_src = hbm_peak()
    configs = {}
    if not args.no_configs:
        only = [s for s in args.only.split(",") if s] or list(CHAIN_CONFIGS)
        per_gpu = {"hoa": 1024, "binaural": 1024, "objects": 512, "fd_4096": 4096}
        for name in only:
            spec = ChainSpec(name, per_gpu[name])
            res = bench_chain(lib, spec, args, rank, world, dist, torch, peak)
            torch.cuda.empty_cache()
            if rank == 0 and not args.no_cpu_baseline:
                per, frames = CPU_SAMPLE[name]
                res["cpu_baseline"] = cpu_chain_baseline(spec, 1, per, frames)
            configs[name] = res

    if rank == 0:
        achieved = BYTES_PER_UNIT * UNITS * FPL / (ms_per_step * 1e-3) / 1e9
        own_min = (8192 + 8192 / FPL) * UNITS * FPL  # spectra in + PCM out every frame, overlap r/w once per launch
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fd_synth_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        traffic_f = None
        tpath = os.path.join(ROOT, "profiles", "fd_synth_frames_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("frames_per_launch") == FPL and tj.get("n_units") == UNITS:
                traffic_f = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
            "notes": {"l2_policy": f"inputs larger than L2: {GF if FPL > 1 else G} stream groups rotated, "
                      f"{(GF * (2 * FPL + 1) if FPL > 1 else G * 3) * UNITS * FRAME * 4 / 1e6:.0f} MB working set, "
                      "state carried per group",
                      "host_numa_binding": (f"rank 0 on {len(numa_cpus)} CPUs of its GPU's NUMA node" if numa_cpus
                                            else "none")},
            "realtime_factor": value / CHANNELS / 48000.0,
            "mixed_sequence_variant": {
                "note": "same shape with ~8 % LONG_START/EIGHT_SHORT/LONG_STOP frames (SURVEY.md 8d variant)",
                "ms_per_step": ms_mixed, "value": world * UNITS * FPL * FRAME / (ms_mixed * 1e-3),
                "roofline_frac": BYTES_PER_UNIT * UNITS * FPL / (ms_mixed * 1e-3) / 1e9 / peak},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic if FPL == 1 else traffic_f,
                         "traffic_source": ("profiles/fd_synth_traffic.json" if FPL == 1 else
                                            "profiles/fd_synth_frames_traffic.json") +
                                           " (one ncu --set full capture of this kernel, committed; not measured in "
                                           "this run)",
                         "peak_source": peak_src,
                         "kernel": "fd_synth_ws_kernel" if FPL == 1 else "fd_synth_ws_frames_kernel",
                         "algorithmic_bytes_per_unit": BYTES_PER_UNIT,
                         "units_per_launch": UNITS * FPL,  # channel-frames: streams x channels x frames_per_step
                         "algorithmic_bytes_per_launch": BYTES_PER_UNIT * UNITS * FPL,
                         "note": "SURVEY.md 8(d): 16 KB per channel-frame (spectra in, overlap read + write, PCM out), "
                                 "what ia_core_coder_fd_frm_dec moves per call" +
                                 ("" if FPL == 1 else
                                  f"; with {FPL} frames per launch a unit's overlap stays on the SM between frames, so "
                                  f"the launch's own minimum is {own_min / UNITS / FPL:.0f} B per channel-frame: "
                                  "frac_of_own_minimum is the fraction of the HBM peak on that smaller figure (the "
                                  "kernel is then bound by issue / shared-memory throughput, not HBM); frac, on the "
                                  "per-call figure, exceeds 1 when the launch is faster than a kernel that moved all "
                                  "16 KB per channel-frame could be"),
                         "frac_of_own_minimum": own_min / (ms_per_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": UNITS * FRAME * 4 + UNITS * 4,
                    "d2h_bytes_per_step": UNITS * FRAME * 4, "ms_per_step": e2e_s * 1e3,
                    "h2d_GBps": (UNITS * FRAME * 4 + UNITS * 4) / e2e_s / 1e9, "d2h_GBps": UNITS * FRAME * 4 / e2e_s / 1e9,
                    "api": "mpeghb200_fd_submit / mpeghb200_fd_wait, two frames in flight (pinned host buffers)",
                    "synchronous_call_ms_per_step": e2e_sync_s * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if single:
            line["single_frame_launch_variant"] = {
                "note": "the same workload one frame of every stream per launch (mpeghb200_fd_synth_dev; the step of "
                        "round 1's bench line)",
                "ms_per_step": single[0], "value": world * UNITS * FRAME / (single[0] * 1e-3),
                "roofline_frac": BYTES_PER_UNIT * UNITS / (single[0] * 1e-3) / 1e9 / peak,
                "mixed_sequence_ms_per_step": single[1],
                "mixed_sequence_roofline_frac": BYTES_PER_UNIT * UNITS / (single[1] * 1e-3) / 1e9 / peak}
        if configs:
            line["configs"] = configs
        if not args.no_cpu_baseline:
            fd = ChainSpec("fd_4096", STREAMS)
            fd.name = "fd"
            per, frames = CPU_SAMPLE["fd"]
            line["cpu_baseline"] = cpu_chain_baseline(fd, 1, per, frames)
            line["configs0_cpu_testbench"] = cpu_testbench_baseline(cpu_threads())
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    lib.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
