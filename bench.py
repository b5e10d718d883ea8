#!/usr/bin/env python3
"""bench.py - the hot path of BASELINE.json on N GPUs of one node, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]

Workload (config.workload): BASELINE.json configs[4] - 1024 channels of CHAIN12 (12-stage chain, 4x oversampling on the
four nonlinear stages, 8192-tap power-amp IR, spatializer mix-down, tuner on channel 0) at 192 kHz, block 8192
(controller.go:36).  SURVEY.md 8e: the channels shard over the ranks (STRONG scaling: 1024 / N channels per GPU, 128 at
N = 8) with no data-path exchange except the stereo-master sum (NCCL reduce).  `--scaling weak` (1024 channels per GPU)
is kept as an alternative and reported under extras at N > 1.

A STEP = one "file" of SURVEY.md 8d: 128 consecutive blocks of 8192 samples of every channel (1 048 576 samples per
channel, controller.go:3076-3107 walks a file block by block) through tuner feed -> chains -> spatializer; the stereo master
of the whole step is summed over the ranks ONCE per step (the master never feeds back into the chains).
  value : Msamples/s with the inputs resident in HBM (a ring of distinct block sets larger than L2), CUDA events on the
          launching stream, max over ranks.
  e2e   : the same through HOST buffers (pinned), H2D of every block and D2H of every channel's output + master inside the
          timed region, through the C-ABI pair gdg_controller_submit() / gdg_controller_wait() with two blocks in flight;
          e2e.synchronous_call is the blocking gdg_controller_process().
  roofline : one chain pass of a block (the segment kernels k_chain -> k_os -> k_chain of CHAIN12, timed together with
          CUDA events) against the measured HBM copy bandwidth (16 algorithmic bytes per sample) and - because the path is
          FP64-pipe-bound by construction - against the FP64 FMA rate measured on this part.
  cpu_baseline / --impl reference : the restated reference (C oracle; no Go toolchain exists here or on the GPU box, see
          profiles/r02_toolchain_probe.txt) on a pinned pthread pool, one worker per channel like the reference's
          goroutines (oracle/orc_bench.c).
  extras : BASELINE configs 2 (1 channel, the JACK drop-in gdg_chain_process, latency per block), 3 (IR sweep 256..65536
          taps, 64 channels) and 4 (8 channels + tuner + spatializer); PCM16 end-to-end variants; level meter.
The product arm imports nothing from tests/ or oracle/.
"""
import argparse
import ctypes as Ct
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FS = 192000
N = 8192
CHANNELS_TOTAL = int(os.environ.get("GDG_BENCH_CHANNELS", "1024"))
BLOCKS = int(os.environ.get("GDG_BENCH_BLOCKS", "128"))        # blocks per step = one file of SURVEY.md 8d
IR_TAPS = 8192
ALGO_BYTES_PER_SAMPLE = 16.0                                    # 8 B in + 8 B out
ALGO_FLOP_PER_SAMPLE = 2200.0
FP64_PEAK_TFLOPS = 33.8                                         # measured: tools/fp64_peak.cu -> profiles/fp64_peak_r01.json
METRIC = "Msamples/sec full 12-stage chain @192kHz 4xOS"
WORKLOAD = "cfg5: %d-ch CHAIN12 (4xOS, 8192-tap IR, spatializer, tuner), 192 kHz, block 8192, %d blocks per step"


def chain_traffic():
    """dram bytes of one chain pass (all segment kernels of one block) from the committed ncu capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "kchain_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(self.rows)}


# ======================================================================================================================
# reference arm / cpu_baseline: the restated reference on the host cores (the one place bench.py executes oracle/)
# ======================================================================================================================
def time_oracle(channels, blocks, threads, reps, budget_s=150.0):
    """CHAIN12 @192 kHz + spatializer for `channels` channels x `blocks` blocks on `threads` pinned workers; returns the
    list of Msamples/s over `reps` repetitions (the stream continues from one repetition to the next)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as orc
    presets = importlib.import_module("go-dsp-guitar_b200.presets")
    ob = orc.IRBank()
    ob.add_raw("syn8192", 0, presets.synthetic_ir(IR_TAPS), FS)
    chains = [presets.build_chain(orc.Chain(ob), presets.CHAIN12, "syn8192") for _ in range(channels)]
    spat = orc.Spatializer(channels)
    spat.set_sample_rate(FS)
    for c in range(channels):
        spat.set_azimuth(c, -70.0 + 140.0 * c / max(1, channels - 1))
        spat.set_distance(c, 1.0 + 0.25 * (c % 8))
    base = np.stack([presets.synth_fast(c, FS, N * 4) for c in range(min(channels, 36))])
    x = np.ascontiguousarray(np.tile(base[np.arange(channels) % len(base)], (1, (blocks + 3) // 4))[:, :N * blocks])
    orc.bench_controller(chains[:1], None, x[:1, :N], N, FS, 1)            # builds the FFT tables, faults the pages in
    vals, t0 = [], time.perf_counter()
    for _ in range(reps):
        secs = orc.bench_controller(chains, spat, x, N, FS, threads)
        vals.append(channels * N * blocks / secs / 1e6)
        if time.perf_counter() - t0 > budget_s:
            break
    return vals


def cpu_quota():
    """CPUs this container may actually use: the affinity mask capped by the cgroup CPU quota (the 1-GPU boxes of this pool
    show 128 logical CPUs under cpu.max = 16 CPUs: 128 runnable threads are then throttled to 16 CPUs' worth of time)"""
    n = len(os.sched_getaffinity(0))
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    quota = float(txt[0]) / float(txt[1])
            else:
                q = float(txt[0])
                if q > 0:
                    quota = q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            break
        except Exception:
            continue
    return n, quota


def cpu_throttle_stat():
    """(nr_throttled, throttled_usec) of this container's cgroup: when they move during the CPU arm, the quota - not the code -
    set the rate"""
    for path in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.stat"):
        try:
            kv = dict(l.split()[:2] for l in open(path).read().splitlines() if l.strip())
            return int(kv.get("nr_throttled", 0)), int(kv.get("throttled_usec", int(kv.get("throttled_time", 0)) // 1000))
        except Exception:
            continue
    return None


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def cpu_baseline_record(reps, blocks=64):
    host_cpus, quota = cpu_quota()
    cores = max(1, min(host_cpus, int(math.ceil(quota)) if quota else host_cpus))
    th0 = cpu_throttle_stat()
    vals = time_oracle(cores, blocks, cores, reps)
    th1 = cpu_throttle_stat()
    one = time_oracle(1, 4, 1, 2)
    v = float(np.median(vals))
    sample = "%d channels (one pinned worker each, like the reference's goroutine per channel) x %d blocks of %d, CHAIN12 @192 kHz + spatializer; median of %d repetitions" % (cores, blocks, N, len(vals))
    return {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample,
            "host_logical_cpus": host_cpus, "cgroup_cpu_quota": quota, "blocks_per_repetition": blocks,
            "repetitions": [round(a, 3) for a in vals], "spread": (max(vals) - min(vals)) / v if v else None,
            "per_core_single_thread": float(np.median(one)), "per_core_in_full_run": v / cores,
            "best_repetition": float(max(vals)), "cpu_model": cpu_model(),
            "cgroup_throttled_during_run": ({"periods": th1[0] - th0[0], "usec": th1[1] - th0[1]} if th0 and th1 else None),
            "note": "threads = min(logical CPUs, cgroup CPU quota): the pool's boxes expose 128 logical CPUs but schedule the container on a quota (16 CPUs per GPU); per_core_single_thread x physical cores is the figure to extrapolate to a whole host"}


def run_reference(args):
    """--impl reference: the restated reference CPU path on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    rec = cpu_baseline_record(max(5, min(args.steps, 12)))
    v = rec["value"]
    blocks = rec["blocks_per_repetition"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": len(rec["repetitions"]),
        "warmup": args.warmup, "ms_per_step": 1e3 * rec["cores"] * N * blocks / (v * 1e6), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (CHANNELS_TOTAL, BLOCKS), "sample": rec["sample"],
                   "note": "a bounded sample of the workload: as many channels as usable host threads (the reference runs one goroutine per channel), %d blocks per step" % blocks},
        "cpu_baseline": rec,
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ======================================================================================================================
# product arm
# ======================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--scaling", default=os.environ.get("GDG_BENCH_SCALING", "strong"), choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    gdg = importlib.import_module("go-dsp-guitar_b200")
    presets = importlib.import_module("go-dsp-guitar_b200.presets")

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path")
    torch.cuda.set_device(local)
    # run this rank's host thread (and first-touch its pinned buffers) on the CPUs next to its GPU
    all_cpus = os.sched_getaffinity(0)
    numa_note = "unbound"
    if os.environ.get("GDG_BENCH_NUMA", "1") != "0":
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(all_cpus) // 64) + 1)
            near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1} & all_cpus
            if near:
                os.sched_setaffinity(0, near)
                numa_note = "%d CPUs near GPU %d" % (len(near), local)
        except Exception as e:                                   # affinity is an optimisation, never a requirement
            numa_note = "unbound (%s)" % type(e).__name__
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = gdg.load()
    gdg.init(local, lib)

    def shard(total):
        """contiguous channel range of this rank (controller.go:3339-3341: one worker per channel)"""
        per, rem = divmod(total, world)
        lo = rank * per + min(rank, rem)
        return lo, per + (1 if rank < rem else 0)

    strong = args.scaling == "strong"
    first_channel, C = shard(CHANNELS_TOTAL) if strong else (rank * CHANNELS_TOTAL, CHANNELS_TOTAL)
    total_channels = CHANNELS_TOTAL if strong else CHANNELS_TOTAL * world

    def make_set(first, count):
        pb = gdg.IRBank(lib)
        pb.add_raw("syn8192", 0, presets.synthetic_ir(IR_TAPS), FS)
        chains = [presets.build_chain(gdg.Chain(pb, lib=lib), presets.CHAIN12, "syn8192") for _ in range(count)]
        spat = gdg.Spatializer(count, lib=lib)
        spat.set_sample_rate(FS)
        for c in range(count):
            g = first + c
            spat.set_azimuth(c, -70.0 + 140.0 * g / max(1, total_channels - 1))
            spat.set_distance(c, 1.0 + 0.25 * (g % 8))
        return pb, chains, spat

    bank, chains, spat = make_set(first_channel, C)
    runner = gdg.BatchRunner(chains)
    tuner = gdg.Tuner(lib=lib) if rank == 0 else None             # the tuner lives with channel 0's GPU (controller.go:2668-2672)

    # ---- synthetic input ring in HBM (SURVEY.md 8d signal, generated on the device), larger than L2 with the outputs ----
    ring_blocks = max(4, int(math.ceil(300e6 / (2.0 * C * N * 8))))
    ring_blocks = min(ring_blocks, 64)
    g = torch.Generator(device=dev); g.manual_seed(1337 + rank)

    def synth(first, count, nblk):
        T = nblk * N
        t = torch.arange(T, device=dev, dtype=torch.float64)
        ch = torch.arange(count, device=dev, dtype=torch.float64) + first
        f = 82.4069 * torch.pow(torch.tensor(2.0, device=dev, dtype=torch.float64), torch.remainder(ch, 36) / 12.0)
        ph = 2 * np.pi * f[:, None] * t[None, :] / FS
        x = 0.4 * torch.sin(ph) + 0.2 * torch.sin(2.01 * ph) + 0.05 * (2 * torch.rand((count, T), device=dev, dtype=torch.float64, generator=g) - 1)
        return [x[:, k * N:(k + 1) * N].contiguous() for k in range(nblk)]

    ring_in = synth(first_channel, C, ring_blocks)
    ring_out = [torch.empty((C, N), device=dev, dtype=torch.float64) for _ in range(ring_blocks)]
    master = torch.empty((2, BLOCKS * N), device=dev, dtype=torch.float64)        # the step's stereo master [2][T]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream

    def step_device(s):
        for k in range(BLOCKS):
            j = (s * BLOCKS + k) % ring_blocks
            xin, out = ring_in[j], ring_out[j]
            if tuner is not None:
                tuner.process_device(xin.data_ptr(), N, FS, sh)                                         # controller.go:2668-2672
            runner.process_device(xin.data_ptr(), out.data_ptr(), N, N, FS, sh)                          # controller.go:2682-2705
            spat.process_device(out.data_ptr(), N, master.data_ptr() + 8 * k * N, BLOCKS * N, N, None, False, sh)   # controller.go:2761
        if world > 1:
            dist.reduce(master, dst=0, op=dist.ReduceOp.SUM)                                             # the one exchange step, once per file

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for k in range(warmup):
            fn(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            fn(warmup + k)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = gdg.launch_count(lib)
    if sampler:
        sampler.start()
    ms_total = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    launches = gdg.launch_count(lib) - launches0
    launches_per_step = launches / float(args.steps + args.warmup)
    samples_per_step = total_channels * N * BLOCKS
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- the chain pass alone (segment kernels of one block), for the roofline ----
    def step_kernel(s):
        for k in range(BLOCKS):
            j = (s * BLOCKS + k) % ring_blocks
            runner.process_device(ring_in[j].data_ptr(), ring_out[j].data_ptr(), N, N, FS, sh)
    ksteps = max(1, min(args.steps, 4))
    ms_k = timed(step_kernel, ksteps, 1) / (ksteps * BLOCKS)
    hbm_peak, peak_kind = peaks()
    achieved = ALGO_BYTES_PER_SAMPLE * C * N / (ms_k * 1e-3) / 1e9
    tfl = ALGO_FLOP_PER_SAMPLE * C * N / (ms_k * 1e-3) / 1e12
    seg_launches = 3                                             # CHAIN12: k_chain(gate, compressor) -> k_os(4 oversampled units) -> k_chain(6 units)

    # ---- end to end through host buffers ----
    h_in = [torch.empty((C, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_out = [torch.empty((C + 2, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_master = torch.empty((2, BLOCKS * N), dtype=torch.float64).pin_memory()
    for k in range(2):
        h_in[k].copy_(ring_in[k])
    torch.cuda.synchronize()
    handles = (Ct.c_void_p * C)(*[c_._h for c_ in chains])
    ptrs_in = [(Ct.c_void_p * C)(*[h_in[k].data_ptr() + 8 * N * c for c in range(C)]) for k in range(2)]
    ptrs_out = [(Ct.c_void_p * (C + 2))(*[h_out[k].data_ptr() + 8 * N * c for c in range(C + 2)]) for k in range(2)]
    tun_h, tun_ch = (tuner._h if tuner is not None else None), (0 if tuner is not None else -1)

    def keep_master(slot, k):
        """block k's master pair: stays on the host at N = 1; goes back to the device for the per-step sum at N > 1"""
        if world > 1:
            master[:, k * N:(k + 1) * N].copy_(h_out[slot][C:], non_blocking=True)
        else:
            h_master[:, k * N:(k + 1) * N] = h_out[slot][C:]

    def finish_step():
        if world > 1:
            dist.reduce(master, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                h_master.copy_(master, non_blocking=True)
            stream.synchronize()

    def check(rc):
        if rc < 0:
            raise RuntimeError(lib.gdg_last_error().decode())
        return rc

    def run_sync(steps):
        for _s in range(steps):
            for k in range(BLOCKS):
                check(lib.gdg_controller_process(handles, C, tun_h, tun_ch, spat._h, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS))
                keep_master(k % 2, k)
            finish_step()

    def run_pipelined(steps):
        # the batch loop of controller.go:2991-3106 with two blocks in flight: block k+1 is submitted (its H2D starts) before
        # block k is waited for (its D2H finishes); every block's inputs go up and every block's outputs come back in the loop
        for _s in range(steps):
            prev = None
            for k in range(BLOCKS):
                t = check(lib.gdg_controller_submit(handles, C, tun_h, tun_ch, spat._h, None, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS))
                if prev is not None:
                    check(lib.gdg_controller_wait(prev[0]))
                    keep_master(prev[1] % 2, prev[1])
                prev = (t, k)
            check(lib.gdg_controller_wait(prev[0]))
            keep_master(prev[1] % 2, prev[1])
            finish_step()

    def wall(fn):
        barrier()
        t0 = time.perf_counter()
        fn()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())
    e2e_api = ("gdg_controller_submit / gdg_controller_wait on pinned host buffers: one submit per block, two blocks in flight"
               + ("; the step's [2][%d] master goes back up, is summed with ONE NCCL reduce per step and comes down on rank 0" % (BLOCKS * N) if world > 1 else ""))
    esteps = max(1, min(args.steps, 6))
    run_sync(1)
    e2e_sync = samples_per_step * esteps / wall(lambda: run_sync(esteps)) / 1e6
    run_pipelined(1)
    e2e_value = samples_per_step * esteps / wall(lambda: run_pipelined(esteps)) / 1e6
    h2d_step = BLOCKS * C * N * 8 + (BLOCKS * 2 * N * 8 if world > 1 else 0)
    d2h_step = BLOCKS * (C + 2) * N * 8 + (BLOCKS * 2 * N * 8 if (world > 1 and rank == 0) else 0)

    extras = {}
    xsteps = max(1, min(args.steps, 3))
    if not args.no_extras:
        # ---- the same chains fed with PCM16 payloads, codecs inside the segment kernels ----
        try:
            p_in = [torch.empty((C, N), dtype=torch.int16).pin_memory() for _ in range(2)]
            p_out = [torch.empty((C, N), dtype=torch.int16).pin_memory() for _ in range(2)]
            for k in range(2):
                p_in[k].copy_((ring_in[k] * 32767.0).clamp(-32768, 32767).to(torch.int16))
            rows = [((Ct.c_void_p * C)(*[p_in[k].data_ptr() + 2 * N * c for c in range(C)]),
                     (Ct.c_void_p * C)(*[p_out[k].data_ptr() + 2 * N * c for c in range(C)])) for k in range(2)]

            def run_pcm(steps):                                      # two blocks in flight, like the headline e2e
                prev = None
                for k in range(steps * BLOCKS):
                    t = check(lib.gdg_process_batch_encoded_submit(handles, C, rows[k % 2][0], 1, rows[k % 2][1], 1, N, FS))
                    if prev is not None:
                        check(lib.gdg_controller_wait(prev))
                    prev = t
                check(lib.gdg_controller_wait(prev))
            run_pcm(1)
            extras["e2e_pcm16_payload"] = {"value": samples_per_step * xsteps / wall(lambda: run_pcm(xsteps)) / 1e6, "unit": "Msamples/s",
                                           "api": "gdg_process_batch_encoded_submit / gdg_controller_wait (PCM16 in / PCM16 out, chains only: no tuner / spatializer)",
                                           "h2d_bytes_per_step": BLOCKS * 2 * C * N, "d2h_bytes_per_step": BLOCKS * 2 * C * N}
            q_out = [torch.empty((C + 2, N), dtype=torch.int16).pin_memory() for _ in range(2)]
            qo = [(Ct.c_void_p * (C + 2))(*[q_out[k].data_ptr() + 2 * N * c for c in range(C + 2)]) for k in range(2)]

            def run_full_pcm(steps):
                prev = None
                for k in range(steps * BLOCKS):
                    t = check(lib.gdg_controller_submit_encoded(handles, C, tun_h, tun_ch, spat._h, None, rows[k % 2][0], 1, qo[k % 2], 1, C + 2, N, FS))
                    if prev is not None:
                        check(lib.gdg_controller_wait(prev))
                    prev = t
                check(lib.gdg_controller_wait(prev))
            run_full_pcm(1)
            extras["e2e_pcm16_controller"] = {"value": samples_per_step * xsteps / wall(lambda: run_full_pcm(xsteps)) / 1e6, "unit": "Msamples/s",
                                              "api": "gdg_controller_submit_encoded / gdg_controller_wait (PCM16 rows in and out, tuner feed + chains + spatializer; this rank's partial master, no NCCL reduce)",
                                              "h2d_bytes_per_step": BLOCKS * 2 * C * N, "d2h_bytes_per_step": BLOCKS * 2 * (C + 2) * N}
        except Exception as e:                                        # never let an extra break the contract line
            extras["e2e_pcm16"] = {"error": str(e)}

        # ---- the headline e2e call with the level meter of controller.go:2653-2781 fed on the device ----
        try:
            meter = gdg.Meter(2 * C + 3, lib=lib)
            meter.set_enabled(True)

            def run_metered(steps):
                for k in range(steps * BLOCKS):
                    check(lib.gdg_controller_process_metered(handles, C, tun_h, tun_ch, spat._h, meter._h, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS))
            run_metered(1)
            extras["e2e_with_level_meter"] = {"value": samples_per_step * xsteps / wall(lambda: run_metered(xsteps)) / 1e6, "unit": "Msamples/s",
                                              "api": "gdg_controller_process_metered, blocking (%d meter channels: inputs, outputs, master)" % (2 * C + 3),
                                              "input_0_level_peak_db": list(meter.analyze(0)), "master_left_level_peak_db": list(meter.analyze(2 * C + 1))}
        except Exception as e:
            extras["e2e_with_level_meter"] = {"error": str(e)}

        if rank == 0:
            try:
                extras.update(config_extras(torch, gdg, presets, lib, dev, stream, args))
            except Exception as e:
                extras["configs_2_3_4"] = {"error": "%s: %s" % (type(e).__name__, e)}

        # ---- weak scaling (1024 channels per GPU) beside the strong-scaling headline, N > 1 only ----
        if world > 1 and strong:
            try:
                del ring_in, ring_out
                _b2, chains_w, spat_w = make_set(rank * CHANNELS_TOTAL, CHANNELS_TOTAL)
                runner_w = gdg.BatchRunner(chains_w)
                rb = 4
                win = synth(rank * CHANNELS_TOTAL, CHANNELS_TOTAL, rb)
                wout = [torch.empty((CHANNELS_TOTAL, N), device=dev, dtype=torch.float64) for _ in range(rb)]

                def step_weak(s):
                    for k in range(BLOCKS):
                        j = (s * BLOCKS + k) % rb
                        runner_w.process_device(win[j].data_ptr(), wout[j].data_ptr(), N, N, FS, sh)
                        spat_w.process_device(wout[j].data_ptr(), N, master.data_ptr() + 8 * k * N, BLOCKS * N, N, None, False, sh)
                    dist.reduce(master, dst=0, op=dist.ReduceOp.SUM)
                msw = timed(step_weak, xsteps, 1)
                extras["weak_scaling"] = {"value": world * CHANNELS_TOTAL * N * BLOCKS * xsteps / (msw * 1e-3) / 1e6, "unit": "Msamples/s",
                                          "channels_per_gpu": CHANNELS_TOTAL, "note": "device-resident, same step definition, one NCCL reduce per step"}
            except Exception as e:
                extras["weak_scaling"] = {"error": str(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                        # the CPU baseline gets every host core
        cpu_baseline = cpu_baseline_record(5)

    # ---- extra: the SAME N GPUs driven by ONE process through gdg_cluster_* (what a Go host would do: no torch.distributed,
    # no NCCL; the stereo master is summed on device 0 by one kernel reading the other devices over NVLink peer addresses).
    # Measured by rank 0 AFTER the other ranks have left (their CUDA contexts would share the GPUs with it otherwise). ----
    if world > 1:
        barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
        if not args.no_extras:
            time.sleep(4.0)                                       # the other ranks tear their contexts down
            try:
                extras["single_process_cluster"] = cluster_extra(torch, gdg, presets, lib, world, xsteps)
            except Exception as e:
                extras["single_process_cluster"] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % (total_channels, BLOCKS), "channels_total": total_channels, "channels_per_gpu": C, "block": N,
                       "blocks_per_step": BLOCKS, "sample_rate": FS,
                       "cache": "inputs larger than L2: ring of %d distinct in/out block sets (%.0f MB) cycled" % (ring_blocks, ring_blocks * 2 * C * N * 8 / 1e6),
                       "parallelism": "channels sharded %d-way (%s scaling), one NCCL reduce of the [2][%d] stereo master per step" % (world, args.scaling, BLOCKS * N)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step, "api": e2e_api,
                    "host_affinity": numa_note, "steps": esteps,
                    "synchronous_call": {"value": e2e_sync, "unit": "Msamples/s", "api": "gdg_controller_process (one blocking C-ABI call per block)"}},
            "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
            "roofline": {"kernel": "one chain pass of a block = the %d segment kernels of CHAIN12: k_chain(noise_gate, compressor) -> k_os(excess, fuzz, overdrive, distortion) -> k_chain(tone_stack .. cabinet)" % seg_launches,
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": chain_traffic() if C == 1024 else None,
                         "traffic_note": "dram bytes of the pass from the committed ncu capture (profiles/kchain_traffic.json): ~1 MB of unit state per channel-block (delay / reverb / chorus rings, FFT tail) and two float64 round trips between the segments on top of the 134 MB of audio",
                         "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * C * N,
                         "unfused_byte_model": {"bytes_per_sample": 16.0 * seg_launches, "achieved_gbs": 16.0 * seg_launches * C * N / (ms_k * 1e-3) / 1e9},
                         "note": "FP64-pipe-bound by construction (~136 FLOP/B, SURVEY.md 8d); see fp64",
                         "fp64": {"achieved_tflops": tfl, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tfl / FP64_PEAK_TFLOPS,
                                  "flop_per_sample": ALGO_FLOP_PER_SAMPLE, "peak_kind": "measured (tools/fp64_peak.cu)"}},
            "cpu_baseline": cpu_baseline, "extras": extras}))


def cluster_extra(torch, gdg, presets, lib, ndev, steps):
    """CHANNELS_TOTAL channels over `ndev` GPUs from ONE process: gdg_cluster_submit / gdg_cluster_wait on pinned host rows,
    two blocks in flight; wall clock, every block's H2D and D2H inside."""
    cl = gdg.Cluster(list(range(ndev)), lib=lib)
    per, rem = divmod(CHANNELS_TOTAL, ndev)
    chains, spats, keep = [], [], []
    g0 = 0
    for m in range(ndev):
        cnt = per + (1 if m < rem else 0)
        cl.select(m)
        bank = gdg.IRBank(lib)
        bank.add_raw("syn8192", 0, presets.synthetic_ir(IR_TAPS), FS)
        keep.append(bank)
        sp = gdg.Spatializer(cnt, lib=lib)
        sp.set_sample_rate(FS)
        for c in range(cnt):
            g = g0 + c
            chains.append(presets.build_chain(gdg.Chain(bank, lib=lib), presets.CHAIN12, "syn8192"))
            sp.set_azimuth(c, -70.0 + 140.0 * g / max(1, CHANNELS_TOTAL - 1))
            sp.set_distance(c, 1.0 + 0.25 * (g % 8))
        spats.append(sp)
        g0 += cnt
    cl.select(0)
    tuner = gdg.Tuner(lib=lib)
    C = CHANNELS_TOTAL
    h_in = [torch.empty((C, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_out = [torch.empty((C + 2, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    base = torch.tensor(np.stack([presets.synth_fast(c, FS, N) for c in range(36)]))
    for k in range(2):
        h_in[k].copy_(base[torch.arange(C) % 36])
    handles = (Ct.c_void_p * C)(*[c_._h for c_ in chains])
    sph = (Ct.c_void_p * ndev)(*[s_._h for s_ in spats])
    pin = [(Ct.c_void_p * C)(*[h_in[k].data_ptr() + 8 * N * c for c in range(C)]) for k in range(2)]
    pout = [(Ct.c_void_p * (C + 2))(*[h_out[k].data_ptr() + 8 * N * c for c in range(C + 2)]) for k in range(2)]

    def check(rc):
        if rc < 0:
            raise RuntimeError(lib.gdg_last_error().decode())
        return rc

    def run(nblocks):
        prev = None
        for k in range(nblocks):
            t = check(lib.gdg_cluster_submit(cl._h, handles, C, tuner._h, 0, sph, pin[k % 2], pout[k % 2], C + 2, N, FS))
            if prev is not None:
                check(lib.gdg_cluster_wait(cl._h, prev))
            prev = t
        check(lib.gdg_cluster_wait(cl._h, prev))
    run(8)
    nb = steps * BLOCKS
    t0 = time.perf_counter()
    run(nb)
    dt = time.perf_counter() - t0
    gdg.use_device(0, lib)
    return {"value": C * N * nb / dt / 1e6, "unit": "Msamples/s", "devices": ndev, "channels_per_device": per,
            "api": "gdg_cluster_submit / gdg_cluster_wait from ONE process (host rows, two blocks in flight); master = peer-memory sum kernel on device 0, no NCCL",
            "blocks": nb}


def config_extras(torch, gdg, presets, lib, dev, stream, args):
    """BASELINE.json configs 2, 3 and 4 (rank 0, one GPU): latency / throughput beside the parity tests that cover them."""
    out = {}
    sh = stream.cuda_stream

    def check(rc):
        if rc < 0:
            raise RuntimeError(lib.gdg_last_error().decode())
        return rc

    # ---- config 2: ONE channel, CHAIN12 with a real guitar-cabinet IR at 96 kHz through the JACK drop-in gdg_chain_process ----
    bank = gdg.IRBank(lib)
    ir_name = "syn3839"
    fix = os.path.join(ROOT, "tests", "golden", "ir_bank.npz")
    if os.path.exists(fix):                                        # the reference's own IR (fixture data: PCM16 frames of ir/guitar/tweed-center.wav)
        z = np.load(fix)
        names = [str(s) for s in z["names"]]
        if "Guitar: American Vintage (Center)" in names:
            k = names.index("Guitar: American Vintage (Center)")
            bank.add("Guitar: American Vintage (Center)", int(z["compensation"][k]), (2.0 / 65535.0) * z["pcm_%02d" % k].astype(np.float64), int(z["sample_rate"]))
            ir_name = "Guitar: American Vintage (Center)"
    if ir_name == "syn3839":
        bank.add_raw(ir_name, 0, presets.synthetic_ir(3839), 96000)
    lat = {}
    for n in (1024, 8192):
        ch = presets.build_chain(gdg.Chain(bank, lib=lib), presets.CHAIN12, ir_name)
        x = presets.synth_fast(0, 96000, n * 48)
        y = np.zeros(n)
        for k in range(8):
            blk = np.ascontiguousarray(x[k * n:(k + 1) * n])
            check(lib.gdg_chain_process(ch._h, blk.ctypes.data, y.ctypes.data, n, 96000))
        ts = []
        for k in range(8, 48):
            blk = np.ascontiguousarray(x[k * n:(k + 1) * n])
            t0 = time.perf_counter()
            check(lib.gdg_chain_process(ch._h, blk.ctypes.data, y.ctypes.data, n, 96000))
            ts.append(time.perf_counter() - t0)
        ts.sort()
        lat["n=%d" % n] = {"median_us": 1e6 * ts[len(ts) // 2], "p90_us": 1e6 * ts[int(0.9 * len(ts))], "block_duration_us": 1e6 * n / 96000.0,
                           "realtime_factor": (n / 96000.0) / ts[len(ts) // 2], "msamples_per_s": n / ts[len(ts) // 2] / 1e6}
    out["cfg2_one_channel_chain12_96k"] = {"api": "gdg_chain_process (blocking, host buffers: H2D, 3 segment kernels on one CTA each, D2H)", "ir": ir_name, "latency": lat,
                                           "note": "one channel cannot fill a GPU: this is a latency number (the JACK period at 96 kHz / 1024 frames is 10.7 ms)"}

    # ---- config 3: IR sweep 256 .. 65536 taps, 64 channels, power_amp only, 192 kHz, device-resident ----
    Cn = 64
    g = torch.Generator(device=dev); g.manual_seed(7)
    xin = [(0.3 * (2 * torch.rand((Cn, N), device=dev, dtype=torch.float64, generator=g) - 1)) for _ in range(4)]
    yout = torch.empty((Cn, N), device=dev, dtype=torch.float64)
    sweep = {}
    for L in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        b2 = gdg.IRBank(lib)
        b2.add_raw("syn", 0, presets.synthetic_ir(L), FS)
        chs = [presets.build_chain(gdg.Chain(b2, lib=lib), [("power_amp", {"level_1": 0}, {"filter_order": "1048576", "filter_1": "syn"})]) for _ in range(Cn)]
        r = gdg.BatchRunner(chs)
        for k in range(4):
            r.process_device(xin[k].data_ptr(), yout.data_ptr(), N, N, FS, sh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record(stream)
        for k in range(reps):
            r.process_device(xin[k % 4].data_ptr(), yout.data_ptr(), N, N, FS, sh)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        Fs = 2 * min(8192, 1 << max(0, (L - 1).bit_length()))              # FFT size per hop (partitions of 8192 taps beyond that)
        M = Fs // 2
        K = max(1, (L + 8191) // 8192)
        hops = max(1, N // M)
        flop = hops * (2 * (5 * M * math.log2(M) + 10 * M) + 6 * M * K + Fs) / N        # per sample, SURVEY 8d power_amp row generalised
        sweep["L=%d" % L] = {"ms_per_block": ms, "msamples_per_s": Cn * N / ms / 1e3, "fft_size": Fs, "partitions": K,
                             "fp64_tflops": flop * Cn * N / (ms * 1e-3) / 1e12, "algorithmic_gbs": 16.0 * Cn * N / (ms * 1e-3) / 1e9}
    out["cfg3_ir_sweep_64ch_192k"] = {"api": "gdg_process_batch_device (power_amp only, 64 channels = 64 CTAs on 148 SMs, one wave)", "sweep": sweep}

    # ---- config 4: 8 channels CHAIN12 + tuner + spatializer, 192 kHz, blocking controller call on host buffers ----
    C4 = 8
    b4 = gdg.IRBank(lib)
    b4.add_raw("syn8192", 0, presets.synthetic_ir(IR_TAPS), FS)
    ch4 = [presets.build_chain(gdg.Chain(b4, lib=lib), presets.CHAIN12, "syn8192") for _ in range(C4)]
    sp4 = gdg.Spatializer(C4, lib=lib)
    sp4.set_sample_rate(FS)
    for c in range(C4):
        sp4.set_azimuth(c, -70.0 + 20.0 * c)
        sp4.set_distance(c, 1.0 + 0.25 * c)
    tu4 = gdg.Tuner(lib=lib)
    hin = torch.empty((C4, N), dtype=torch.float64).pin_memory()
    hout = torch.empty((C4 + 2, N), dtype=torch.float64).pin_memory()
    hin.copy_(torch.tensor(np.stack([presets.synth_fast(c, FS, N) for c in range(C4)])))
    h4 = (Ct.c_void_p * C4)(*[c_._h for c_ in ch4])
    pi = (Ct.c_void_p * C4)(*[hin.data_ptr() + 8 * N * c for c in range(C4)])
    po = (Ct.c_void_p * (C4 + 2))(*[hout.data_ptr() + 8 * N * c for c in range(C4 + 2)])
    ts = []
    for k in range(48):
        t0 = time.perf_counter()
        check(lib.gdg_controller_process(h4, C4, tu4._h, 0, sp4._h, pi, po, C4 + 2, N, FS))
        if k >= 8:
            ts.append(time.perf_counter() - t0)
    ts.sort()
    t0 = time.perf_counter()
    res = tu4.analyze()
    t_an = time.perf_counter() - t0
    out["cfg4_8ch_tuner_spatializer_192k"] = {"api": "gdg_controller_process (blocking, host buffers)", "median_us_per_block": 1e6 * ts[len(ts) // 2],
                                              "block_duration_us": 1e6 * N / FS, "realtime_factor": (N / FS) / ts[len(ts) // 2],
                                              "msamples_per_s": C4 * N / ts[len(ts) // 2] / 1e6, "tuner_analyze_ms": 1e3 * t_an, "tuner_note": str(res.get("note") if isinstance(res, dict) else res)}
    return out


if __name__ == "__main__":
    main()
