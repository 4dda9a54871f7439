#!/usr/bin/env python3
"""bench.py -- throughput of the post-parse signal path on B200 (BASELINE.json metric: channel-samples/s).

Workload at N=1 = BASELINE.json configs[1]: batched IMDCT + windowing/overlap-add, 1024-sample frames,
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
             plain-C port if that build is absent) on the host cores, bounded sample.
  --impl reference   the reference arm: that same CPU function on the full workload with all host threads.

Multi-GPU (torchrun, one rank per GPU): streams are sharded by rank, no collective on the data path
(weak scaling: 256 streams per GPU); time = max over ranks.
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

STREAMS, CHANNELS, FRAME = 256, 24, 1024
UNITS = STREAMS * CHANNELS
BYTES_PER_UNIT = 16384  # coef 4096 + overlap r/w 2*4096 + out 4096
METRIC = "channel-samples/s (FD synthesis: IMDCT+window+OLA, 256 streams x 24 ch x 1024)"
UNIT = "channel-samples/s"


def set_streams(n):
    global STREAMS, UNITS, METRIC
    STREAMS, UNITS = n, n * CHANNELS
    METRIC = f"channel-samples/s (FD synthesis: IMDCT+window+OLA, {n} streams x 24 ch x 1024)"


def workload_name():
    tag = "configs[1]" if STREAMS == 256 else "configs[1] shape at a different stream count"
    return (f"{tag}: IMDCT+windowing/OLA, 1024-sample frames, {STREAMS} streams x 24 ch, synthetic spectra, "
            "ONLY_LONG with random sine/KBD shapes")


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


def cpu_reference_step(kind_lib, coef, side, overlap, out, n_threads, handles):
    """One frame for all units of coef [U,1024] on n_threads host threads (disjoint unit ranges)."""
    from concurrent.futures import ThreadPoolExecutor
    kind, lib = kind_lib
    U = coef.shape[0]
    bounds = np.linspace(0, U, n_threads + 1).astype(int)

    def work(t):
        a, b = int(bounds[t]), int(bounds[t + 1])
        if b <= a:
            return 0
        if kind == "reference":
            return lib.ref_fd_frm_dec_batch(handles[t], coef[a:b], overlap[a:b], out[a:b], side[a:b], 1, b - a)
        return lib.oracle_fd_frm_dec_batch(coef[a:b], overlap[a:b], out[a:b], side[a:b], 1, b - a)

    with ThreadPoolExecutor(n_threads) as ex:
        errs = list(ex.map(work, range(n_threads)))
    assert not any(errs)


def load_cpu_impl():
    from oracle import loader
    ref = loader.load_ref()
    if ref is not None:
        return ("reference", ref)
    return ("port", loader.load_oracle())


def run_cpu(args, sample_units, steps, warmup, n_threads):
    """Times the CPU implementation; returns (channel-samples/s, seconds per step, kind)."""
    kind_lib = load_cpu_impl()
    rng = np.random.default_rng(1234)
    coef = rng.normal(0, 500, (sample_units, FRAME)).astype(np.float32)
    side = np.ascontiguousarray(make_side(rng, sample_units))
    overlap = np.zeros((sample_units, FRAME), np.float32)
    out = np.zeros((sample_units, FRAME), np.float32)
    handles = [kind_lib[1].ref_fd_create() for _ in range(n_threads)] if kind_lib[0] == "reference" else None
    for _ in range(warmup):
        cpu_reference_step(kind_lib, coef, side, overlap, out, n_threads, handles)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(kind_lib, coef, side, overlap, out, n_threads, handles)
    dt = time.perf_counter() - t0
    if handles:
        for h in handles:
            kind_lib[1].ref_fd_destroy(h)
    return sample_units * FRAME * steps / dt, dt / steps, kind_lib[0]


def reference_arm(args, rank):
    if rank != 0:
        return
    n_threads = cpu_threads()
    # bounded sample: a step is the full 6144 channel-frames unless K+W full steps would take more than ~90 s on
    # this box, in which case each step is a proportionally smaller slice of the same workload (whole streams).
    _, t_full, _ = run_cpu(args, UNITS, 1, 1, n_threads)
    sample_units = UNITS
    total = t_full * (args.steps + args.warmup)
    if total > 90.0:
        sample_units = max(CHANNELS * n_threads, int(UNITS * 90.0 / total) // CHANNELS * CHANNELS)
    value, sps, kind = run_cpu(args, sample_units, args.steps, args.warmup, n_threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "streams": STREAMS, "channels": CHANNELS, "frame": FRAME},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": kind,
                         "sample": f"{sample_units} of the workload's {UNITS} channel-frames per step, "
                                   f"{args.steps} steps, ia_core_coder_fd_frm_dec per channel-frame"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--groups", type=int, default=8, help="stream groups rotated through (working set > L2)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=STREAMS,
                    help="streams per GPU (default 256 = BASELINE.json configs[1]; 4096 = the north-star stream count)")
    args = ap.parse_args()
    set_streams(args.streams)
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
    coef = [torch.randn(UNITS, FRAME, device="cuda", generator=gen) * 500.0 for _ in range(G)]
    ovl = [torch.zeros(UNITS, FRAME, device="cuda") for _ in range(G)]
    outb = [torch.empty(UNITS, FRAME, device="cuda") for _ in range(G)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mixed, with_clocks):
        """W warm-up + K timed steps of one workload variant; returns (ms_per_step, launches, clocks)."""
        side = [torch.from_numpy(fd_cases.pack_side(make_side(rng, UNITS, mixed)).astype(np.int32)).cuda()
                for _ in range(G)]

        def step(i):
            g = i % G
            lib.fd_synth_dev(coef[g].data_ptr(), side[g].data_ptr(), ovl[g].data_ptr(), outb[g].data_ptr(), UNITS,
                             stream.cuda_stream)

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

    ms_per_step, launches, clocks = timed(False, True)
    value = world * UNITS * FRAME / (ms_per_step * 1e-3)
    ms_mixed, _, _ = timed(True, False)

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

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = BYTES_PER_UNIT * UNITS / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fd_synth_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "streams_per_gpu": STREAMS, "channels": CHANNELS,
                       "frame": FRAME, "l2_policy": f"inputs larger than L2: {G} stream groups rotated, "
                       f"{G * 3 * UNITS * FRAME * 4 / 1e6:.0f} MB working set, state carried per group",
                       "sharding": "streams by rank, no collective",
                       "host_numa_binding": (f"rank 0 on {len(numa_cpus)} CPUs of its GPU's NUMA node" if numa_cpus
                                             else "none")},
            "realtime_factor": value / CHANNELS / 48000.0,
            "mixed_sequence_variant": {
                "note": "same shape with ~8 % LONG_START/EIGHT_SHORT/LONG_STOP frames (SURVEY.md 8d variant)",
                "ms_per_step": ms_mixed, "value": world * UNITS * FRAME / (ms_mixed * 1e-3),
                "roofline_frac": BYTES_PER_UNIT * UNITS / (ms_mixed * 1e-3) / 1e9 / peak},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "fd_synth_ws_kernel", "algorithmic_bytes_per_launch": BYTES_PER_UNIT * UNITS},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": UNITS * FRAME * 4 + UNITS * 4,
                    "d2h_bytes_per_step": UNITS * FRAME * 4, "ms_per_step": e2e_s * 1e3,
                    "api": "mpeghb200_fd_submit / mpeghb200_fd_wait, two frames in flight (pinned host buffers)",
                    "synchronous_call_ms_per_step": e2e_sync_s * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            n_threads = cpu_threads()
            sample_units = 24 * 16  # 16 streams x 24 ch
            v, sps, kind = run_cpu(args, sample_units, 2500, 5, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": f"{sample_units} channel-frames x 2500 frames on 1 host thread "
                                              f"(box has {n_threads}); ia_core_coder_fd_frm_dec"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    lib.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
