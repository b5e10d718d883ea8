#!/usr/bin/env python3
"""bench.py - the hot path of BASELINE.json on N GPUs of one node, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[4] - 1024-channel CHAIN12 (12-stage
chain, 4x oversampling on the four nonlinear stages, 8192-tap power-amp IR, spatializer
mix-down) at 192 kHz, block 8192 (controller.go:36).  Channels shard over ranks with
no data-path exchange except the stereo-master sum (NCCL reduce); scaling is WEAK:
every rank owns 1024 channels.

A step = one audio block of every channel through tuner feed -> chains ->
spatializer (-> NCCL reduce of the [2][8192] master when N > 1).
  value : Msamples/s with the inputs resident in HBM (a ring of distinct blocks larger
          than L2), CUDA events on the launching stream, max over ranks.
  e2e   : the same through host buffers (pinned), H2D of every block and D2H of every
          channel's output + master inside the timed region, through the C-ABI pair
          gdg_controller_submit() / gdg_controller_wait() with two blocks in flight (the
          batch loop's natural double buffering); e2e.synchronous_call is the blocking
          one-call-per-block gdg_controller_process().  At N>1 the NCCL reduce of the
          stereo master follows each block's wait.
  roofline : the fused chain kernel (k_chain) against the measured HBM copy bandwidth
          (16 algorithmic bytes per sample) - and, because this path is FP64-pipe-bound
          by construction, against the measured FP64 FMA rate as well.
  cpu_baseline / --impl reference : the CPU oracle (restated reference; no Go toolchain
          exists here or on the GPU box), one thread per channel on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FS = 192000
N = 8192
CH_PER_GPU = int(os.environ.get("GDG_BENCH_CHANNELS", "1024"))
IR_TAPS = 8192
RING_BLOCKS = int(os.environ.get("GDG_BENCH_RING", "4"))      # x (in + out) x 67 MB each > 126 MB L2
ALGO_BYTES_PER_SAMPLE = 16.0                                    # 8 B in + 8 B out
ALGO_FLOP_PER_SAMPLE = 2200.0
FP64_PEAK_TFLOPS = 33.8                                         # measured: tools/fp64_peak.cu -> profiles/fp64_peak_r01.json
METRIC = "Msamples/sec full 12-stage chain @192kHz 4xOS"
WORKLOAD = "cfg5: %d-ch/GPU CHAIN12 (4xOS, 8192-tap IR, spatializer), 192 kHz, block 8192"


def kchain_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_chain launch from the committed ncu capture."""
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def oracle_chains(count, with_spat=True):
    import common
    from common import orc
    ob = orc.IRBank()
    ob.add_raw("syn8192", 0, common.synthetic_ir(IR_TAPS), FS)
    chains = [common.build_chain(orc.Chain(ob), common.CHAIN12, "syn8192") for _ in range(count)]
    spat = None
    if with_spat:
        spat = orc.Spatializer(count)
        spat.set_sample_rate(FS)
        for c in range(count):
            spat.set_azimuth(c, -70.0 + 140.0 * c / max(1, count - 1))
            spat.set_distance(c, 1.0 + 0.25 * (c % 8))
    return ob, chains, spat


def time_oracle(threads, blocks):
    """Restated reference on `threads` host threads, one channel each (the reference runs one goroutine per
    channel, controller.go:3339-3341), `blocks` blocks per channel + the serial spatializer pass."""
    import common
    from common import orc
    ob, chains, spat = oracle_chains(threads)
    x = np.stack([common.synth_fast(c, FS, N * blocks) for c in range(threads)])
    out = np.zeros((threads, N))

    def work(c, k):
        out[c] = chains[c].process(x[c, k * N:(k + 1) * N], FS)
    t0 = time.perf_counter()
    for k in range(blocks):
        ts = [threading.Thread(target=work, args=(c, k)) for c in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        spat.process(out)
    dt = time.perf_counter() - t0
    return threads * N * blocks / dt / 1e6, dt


def run_reference(args):
    """--impl reference: the restated reference CPU path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    blocks = 2
    for _ in range(max(1, args.warmup // 3)):
        time_oracle(cores, 1)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, _dt = time_oracle(cores, blocks)
        vals.append(v)
        if time.perf_counter() - t0 > 150:
            break
    v = float(np.median(vals))
    sample = "%d channels (one per host thread) x %d blocks of %d per step, CHAIN12 @192 kHz + spatializer" % (cores, blocks, N)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": 1e3 * cores * N * blocks / (v * 1e6), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % CH_PER_GPU, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import common
    from common import gdg

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path")
    torch.cuda.set_device(local)
    # run this rank's host thread (and first-touch its pinned buffers) on the CPUs next to its GPU: the e2e path moves
    # 134 MB per block per GPU over PCIe and a rank on the far socket pays for it at N = 8
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

    C = CH_PER_GPU
    pb = gdg.IRBank(lib)
    pb.add_raw("syn8192", 0, common.synthetic_ir(IR_TAPS), FS)
    chains = [common.build_chain(gdg.Chain(pb, lib=lib), common.CHAIN12, "syn8192") for _ in range(C)]
    runner = gdg.BatchRunner(chains)
    spat = gdg.Spatializer(C, lib=lib)
    spat.set_sample_rate(FS)
    for c in range(C):
        spat.set_azimuth(c, -70.0 + 140.0 * c / max(1, C - 1))
        spat.set_distance(c, 1.0 + 0.25 * (c % 8))
    tuner = gdg.Tuner(lib=lib) if rank == 0 else None

    # ---- synthetic input ring in HBM (SURVEY.md 8d signal, generated on the device) ----
    g = torch.Generator(device=dev); g.manual_seed(1337 + rank)
    T = RING_BLOCKS * N
    t = torch.arange(T, device=dev, dtype=torch.float64)
    ch = torch.arange(C, device=dev, dtype=torch.float64) + rank * C
    f = 82.4069 * torch.pow(torch.tensor(2.0, device=dev, dtype=torch.float64), torch.remainder(ch, 36) / 12.0)
    ph = 2 * np.pi * f[:, None] * t[None, :] / FS
    x_all = 0.4 * torch.sin(ph) + 0.2 * torch.sin(2.01 * ph) + 0.05 * (2 * torch.rand((C, T), device=dev, dtype=torch.float64, generator=g) - 1)
    del ph
    ring_in = [x_all[:, k * N:(k + 1) * N].contiguous() for k in range(RING_BLOCKS)]
    del x_all
    ring_out = [torch.empty((C + 2, N), device=dev, dtype=torch.float64) for _ in range(RING_BLOCKS)]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream

    def step_device(k):
        xin, out = ring_in[k % RING_BLOCKS], ring_out[k % RING_BLOCKS]
        if tuner is not None:
            tuner.process_device(xin.data_ptr(), N, FS, sh)                         # controller.go:2668-2672
        runner.process_device(xin.data_ptr(), out.data_ptr(), N, N, FS, sh)          # controller.go:2682-2705
        spat.process_device(out.data_ptr(), N, out.data_ptr() + 8 * C * N, N, N, None, False, sh)   # controller.go:2761
        if world > 1:
            dist.reduce(out[C:], dst=0, op=dist.ReduceOp.SUM)                        # the one exchange step

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
    samples_per_step = world * C * N
    value = samples_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- dominant kernel alone (k_chain), for the roofline ----
    def step_kernel(k):
        runner.process_device(ring_in[k % RING_BLOCKS].data_ptr(), ring_out[k % RING_BLOCKS].data_ptr(), N, N, FS, sh)
    ms_k = timed(step_kernel, args.steps, 1) / args.steps
    hbm_peak, peak_kind = peaks()
    achieved = ALGO_BYTES_PER_SAMPLE * C * N / (ms_k * 1e-3) / 1e9
    tfl = ALGO_FLOP_PER_SAMPLE * C * N / (ms_k * 1e-3) / 1e12

    # ---- end to end through host buffers ----
    h_in = [torch.empty((C, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    h_out = [torch.empty((C + 2, N), dtype=torch.float64).pin_memory() for _ in range(2)]
    for k in range(2):
        h_in[k].copy_(ring_in[k])
    torch.cuda.synchronize()
    import ctypes as Ct
    handles = (Ct.c_void_p * C)(*[c_._h for c_ in chains])
    ptrs_in = [(Ct.c_void_p * C)(*[h_in[k].data_ptr() + 8 * N * c for c in range(C)]) for k in range(2)]
    ptrs_out = [(Ct.c_void_p * (C + 2))(*[h_out[k].data_ptr() + 8 * N * c for c in range(C + 2)]) for k in range(2)]
    d_master = torch.empty((2, N), device=dev, dtype=torch.float64)

    tun_h, tun_ch = (tuner._h if tuner is not None else None), (0 if tuner is not None else -1)

    def finish_block(k):                                         # the one exchange step: sum the partial masters onto rank 0
        if world > 1:
            d_master.copy_(h_out[k % 2][C:], non_blocking=True)
            dist.reduce(d_master, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                h_out[k % 2][C:].copy_(d_master, non_blocking=True)
            stream.synchronize()

    def step_sync(k):
        # one C-ABI call on HOST buffers: H2D, tuner feed, chains, spatializer (this rank's partial master), D2H - pipelined inside
        rc = lib.gdg_controller_process(handles, C, tun_h, tun_ch, spat._h, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS)
        if rc < 0:
            raise RuntimeError(lib.gdg_last_error().decode())
        finish_block(k)

    def run_pipelined(steps):
        # the batch loop of controller.go:2991-3106 with two blocks in flight: block k+1 is submitted (its H2D starts) before
        # block k is waited for (its D2H finishes); every block's inputs go up and every block's outputs come back in the loop
        prev = None
        for k in range(steps):
            t = lib.gdg_controller_submit(handles, C, tun_h, tun_ch, spat._h, None, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS)
            if t < 0:
                raise RuntimeError(lib.gdg_last_error().decode())
            if prev is not None:
                if lib.gdg_controller_wait(prev[0]) < 0:
                    raise RuntimeError(lib.gdg_last_error().decode())
                finish_block(prev[1])
            prev = (t, k)
        if lib.gdg_controller_wait(prev[0]) < 0:
            raise RuntimeError(lib.gdg_last_error().decode())
        finish_block(prev[1])

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
               + (" + NCCL reduce of the [2][8192] master" if world > 1 else ""))
    wall(lambda: [step_sync(k) for k in range(max(1, args.warmup // 2))])
    e2e_sync = samples_per_step * args.steps / wall(lambda: [step_sync(k) for k in range(args.steps)]) / 1e6
    run_pipelined(2)
    e2e_value = samples_per_step * args.steps / wall(lambda: run_pipelined(args.steps)) / 1e6

    # ---- extra (not the headline): the same chains fed with PCM16 payloads, codecs fused into the kernel ----
    extras = {}
    try:
        p_in = torch.empty((C, N), dtype=torch.int16).pin_memory()
        p_out = torch.empty((C, N), dtype=torch.int16).pin_memory()
        p_in.copy_((ring_in[0] * 32767.0).clamp(-32768, 32767).to(torch.int16))
        pi = (Ct.c_void_p * C)(*[p_in.data_ptr() + 2 * N * c for c in range(C)])
        po = (Ct.c_void_p * C)(*[p_out.data_ptr() + 2 * N * c for c in range(C)])
        p_in2, p_out2 = torch.empty_like(p_in).pin_memory(), torch.empty_like(p_out).pin_memory()
        p_in2.copy_(p_in)
        pi2 = (Ct.c_void_p * C)(*[p_in2.data_ptr() + 2 * N * c for c in range(C)])
        po2 = (Ct.c_void_p * C)(*[p_out2.data_ptr() + 2 * N * c for c in range(C)])
        rows = [(pi, po), (pi2, po2)]

        def run_pcm(steps):                                      # two blocks in flight, like the headline e2e
            prev = None
            for k in range(steps):
                t = lib.gdg_process_batch_encoded_submit(handles, C, rows[k % 2][0], 1, rows[k % 2][1], 1, N, FS)
                if t < 0:
                    raise RuntimeError(lib.gdg_last_error().decode())
                if prev is not None and lib.gdg_controller_wait(prev) < 0:
                    raise RuntimeError(lib.gdg_last_error().decode())
                prev = t
            if lib.gdg_controller_wait(prev) < 0:
                raise RuntimeError(lib.gdg_last_error().decode())
        run_pcm(2)
        dtp = wall(lambda: run_pcm(args.steps))
        extras["e2e_pcm16_payload"] = {"value": samples_per_step * args.steps / dtp / 1e6, "unit": "Msamples/s",
                                       "api": "gdg_process_batch_encoded_submit / gdg_controller_wait (PCM16 in / PCM16 out, chains only: no tuner / spatializer)",
                                       "h2d_bytes_per_step": 2 * C * N, "d2h_bytes_per_step": 2 * C * N}
    except Exception as e:                                        # never let an extra break the contract line
        extras["e2e_pcm16_payload"] = {"error": str(e)}

    # ---- extra: the WHOLE controller step (tuner feed, chains, spatializer -> master) on PCM16 rows, two blocks in flight ----
    try:
        q_out = [torch.empty((C + 2, N), dtype=torch.int16).pin_memory() for _ in range(2)]
        qo = [(Ct.c_void_p * (C + 2))(*[q_out[k].data_ptr() + 2 * N * c for c in range(C + 2)]) for k in range(2)]
        qi = [pi, pi2]

        def run_full_pcm(steps):
            prev = None
            for k in range(steps):
                t = lib.gdg_controller_submit_encoded(handles, C, tun_h, tun_ch, spat._h, None, qi[k % 2], 1, qo[k % 2], 1, C + 2, N, FS)
                if t < 0:
                    raise RuntimeError(lib.gdg_last_error().decode())
                if prev is not None and lib.gdg_controller_wait(prev) < 0:
                    raise RuntimeError(lib.gdg_last_error().decode())
                prev = t
            if lib.gdg_controller_wait(prev) < 0:
                raise RuntimeError(lib.gdg_last_error().decode())
        run_full_pcm(2)
        dtq = wall(lambda: run_full_pcm(args.steps))
        extras["e2e_pcm16_controller"] = {"value": samples_per_step * args.steps / dtq / 1e6, "unit": "Msamples/s",
                                          "api": "gdg_controller_submit_encoded / gdg_controller_wait (PCM16 rows in and out, tuner feed + chains + spatializer; this rank's partial master, no NCCL reduce)",
                                          "h2d_bytes_per_step": 2 * C * N, "d2h_bytes_per_step": 2 * (C + 2) * N}
    except Exception as e:
        extras["e2e_pcm16_controller"] = {"error": str(e)}

    # ---- extra: the headline e2e call with the level meter of controller.go:2653-2781 fed on the device ----
    try:
        meter = gdg.Meter(2 * C + 3, lib=lib)
        meter.set_enabled(True)

        def step_metered(k):
            if lib.gdg_controller_process_metered(handles, C, tuner._h if tuner is not None else None, 0 if tuner is not None else -1,
                                                  spat._h, meter._h, ptrs_in[k % 2], ptrs_out[k % 2], C + 2, N, FS) < 0:
                raise RuntimeError(lib.gdg_last_error().decode())
        for k in range(2):
            step_metered(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            step_metered(k)
        barrier()
        dtm = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtm, op=dist.ReduceOp.MAX)
        extras["e2e_with_level_meter"] = {"value": samples_per_step * args.steps / float(dtm.item()) / 1e6, "unit": "Msamples/s",
                                          "api": "gdg_controller_process_metered (%d meter channels: inputs, outputs, master)" % (2 * C + 3),
                                          "input_0_level_peak_db": list(meter.analyze(0)), "master_left_level_peak_db": list(meter.analyze(2 * C + 1))}
    except Exception as e:
        extras["e2e_with_level_meter"] = {"error": str(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                        # the CPU baseline gets every host core
        cores = os.cpu_count() or 1
        v, secs = time_oracle(cores, 4)
        cpu_baseline = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                        "sample": "%d channels (one per host thread) x 4 blocks of %d, CHAIN12 @192 kHz + spatializer, %.1f s wall" % (cores, N, secs)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % C, "channels_per_gpu": C, "block": N, "sample_rate": FS,
                       "cache": "inputs larger than L2: ring of %d distinct in/out block sets (%.0f MB) cycled" % (RING_BLOCKS, RING_BLOCKS * (2 * C + 2) * N * 8 / 1e6),
                       "parallelism": "channels sharded %d-way, NCCL reduce of the stereo master" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": C * N * 8, "d2h_bytes_per_step": (C + 2) * N * 8, "api": e2e_api, "host_affinity": numa_note,
                    "synchronous_call": {"value": e2e_sync, "unit": "Msamples/s", "api": "gdg_controller_process (one blocking C-ABI call per block)"}},
            "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
            "roofline": {"kernel": "k_chain (fused 12-unit segment, one CTA per channel)", "bound": "hbm", "achieved": achieved,
                         "peak": hbm_peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": kchain_traffic() if C == 1024 else None,
                         "traffic_note": "bytes per launch from the committed ncu capture (profiles/kchain_traffic.json); ~1 MB of unit state per channel-block (delay/reverb/chorus rings, FFT tail) on top of the 134 MB of audio",
                         "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * C * N,
                         "note": "FP64-pipe-bound by construction (~136 FLOP/B, SURVEY.md 8d); see fp64",
                         "fp64": {"achieved_tflops": tfl, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tfl / FP64_PEAK_TFLOPS,
                                  "flop_per_sample": ALGO_FLOP_PER_SAMPLE, "peak_kind": "measured (tools/fp64_peak.cu)"}},
            "cpu_baseline": cpu_baseline, "extras": extras}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
