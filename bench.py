#!/usr/bin/env python
"""Benchmark of the 2D TM+PML time-stepping hot path (BASELINE.json metric: Mcell-updates/s and HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full FDTD time step (D, E, Hx, Hy and both PML integrals of every cell).  Workload at
N GPUs: BASELINE config 5 weak-scaled -- every GPU owns a 32768 x 32768 fp32 row slab of an
(N*32768) x 32768 grid with npml=80 and the point sinusoid of program 3_2, ghost rows exchanged over
NVLink every time block.  One JSON line on stdout (rank 0).

--impl reference times the REFERENCE's own C/OpenMP step functions (oracle/_ref, compiled from
/root/reference/fd2d/clang/test_3_2.c; falls back to the numpy oracle port when absent) on all host cores,
on a bounded sample of the same workload.  The oracle is used only there and in the cpu_baseline leg --
never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 32768
NPML = 80
BYTES_PER_CELL_UPDATE = 48          # dz,hx,hy,ihx,ihy R+W; naz R; ez W (fp32) -- SURVEY.md 8(d)
METRIC = "Mcell-updates/s, 2D TM+PML fp32"
E2E_REPEATS = 3                     # the end-to-end job is a single ~0.15 s shot whose overlap depends on launch timing: median of 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period),
    falling back to an `nvidia-smi -lms` subprocess when pynvml is unavailable."""
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop_flag, self.thread = index, [], set(), False, None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.BAD.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is None:
            return
        import threading
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=float(np.median(self.samples)), reasons=sorted(self.reasons), samples=len(self.samples))
        return out


def pin_to_gpu_numa_node(index):
    """Run this process on the CPUs next to its GPU (NVML's ideal affinity), so that the pinned host buffers of the e2e
    leg are allocated on the GPU's NUMA node: on a two-socket box a buffer on the far socket halves the PCIe rate.
    FDTD_NO_AFFINITY=1 leaves the affinity alone."""
    if os.environ.get("FDTD_NO_AFFINITY"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


# ------------------------------------------------------------------------------------ CPU legs
def cpu_numpy_port(n, steps):
    """The numpy oracle (bit-identical restatement of fd2d/program/fd2d_3_2.py) on one core."""
    from oracle import fdtd_oracle as orc
    g = orc.Grid2D(n, n, NPML, np.float32, point=(n // 2 - 5, n // 2 - 5))
    src = orc.source_table("sine", steps + 1, freq=1500e6)
    orc.step_2d(g, np.int32(1), src[0])                       # warm-up (page faults, temporaries)
    t0 = time.perf_counter()
    for k in range(1, steps + 1):
        orc.step_2d(g, np.int32(k + 1), src[k])
    dt = time.perf_counter() - t0
    return n * n * steps / dt / 1e6, dt


def cpu_reference_c(nx, ny, steps, warm=1):
    """The reference's own C/OpenMP dfield/efield/hfield (oracle/_ref/libref_fd2d_3_2.so), all host threads."""
    import ctypes as C
    from oracle import fdtd_oracle as orc
    from oracle import ref_c
    lib = ref_c.load("3_2")
    g = orc.Grid2D(nx, ny, NPML, np.float32, point=(nx // 2 - 5, ny // 2 - 5))
    ps = ref_c.pml_struct(g.pml)
    p = ref_c._p

    def one(t):
        lib.dfield(C.c_int(t), nx, ny, C.byref(ps), p(g.dz), p(g.hx), p(g.hy))
        lib.efield(nx, ny, p(g.naz), p(g.dz), p(g.ez))
        lib.hfield(nx, ny, C.byref(ps), p(g.ez), p(g.ihx), p(g.ihy), p(g.hx), p(g.hy))
    for t in range(1, warm + 1):
        one(t)
    t0 = time.perf_counter()
    for t in range(warm + 1, warm + steps + 1):
        one(t)
    dt = time.perf_counter() - t0
    return nx * ny * steps / dt / 1e6, dt


def run_reference(args, emit=print):
    """--impl reference: rank 0 only; a bounded sample (8192 x 8192 of the 32768 x 32768 workload) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_c
    cores = os.cpu_count() or 1
    n = 8192
    steps, warm = max(1, args.steps), max(1, args.warmup)
    # keep the whole run within a few minutes whatever K the driver asks for
    steps, warm = min(steps, 40), min(warm, 3)
    if ref_c.available("3_2"):
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        v, dt = cpu_reference_c(n, n, steps, warm)
        kind, how = "reference", "reference C/OpenMP step functions (fd2d/clang/test_3_2.c) via oracle/_ref"
    else:
        n, steps = 4096, min(steps, 10)
        v, dt = cpu_numpy_port(n, steps)
        cores, kind, how = 1, "port", "numpy oracle port (oracle/_ref not built)"
    sample = f"{n}x{n} fp32 sub-grid of the {N_FULL}x{N_FULL} workload, {steps} steps after {warm} warm-up; {how}"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mcell-updates/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"fd2d TM+PML {N_FULL}x{N_FULL} fp32 npml={NPML} point sinusoid (BASELINE config 5)",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args, emit=print):
    # run_streamed (the e2e leg) drives one stream per pass level: give the driver enough hardware queues that they do
    # not alias (default 8).  Must be set before the CUDA context exists.
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import torch.distributed as dist
    from simulation_b200 import fd2d, surface

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n, K, W, T = args.size, args.steps, args.warmup, args.tblock
    knobs = ("FDTD_FORCE_V", "FDTD_CHUNK_ROWS", "FDTD_WARPS", "FDTD_RING", "FDTD_CAREFUL")
    if any(k in os.environ for k in knobs):                                                 # tuning sweeps only
        from simulation_b200 import _lib
        _lib.lib().fdtd2d_tune(*[int(os.environ.get(k, "0")) for k in knobs])
    nx_global = n * world                       # weak scaling: one n x n slab per GPU
    wave = surface.Sinusoid(1500e6)
    src = fd2d.PointSource(nx_global // 2 - 5, n // 2 - 5, wave, hard=True)

    if world > 1:
        from simulation_b200 import slab
        sim = slab.SlabFdtd2D(nx_global, n, NPML, np.float32, source=src, tblock=T)
        barrier = dist.barrier
    else:
        sim = fd2d.Fdtd2D(n, n, NPML, np.float32, source=src, tblock=T)
        barrier = lambda: None

    def sync():
        barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K steps, inputs already in HBM
    sim.advance(W)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    sim.advance(K)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = (K + T - 1) // T                  # passes; each pass = interior kernel + edge kernel
    cells = float(nx_global) * n
    value = cells * K / (ms * 1e-3) / 1e6
    peak, peak_src = peaks()
    per_gpu_cells = float(n) * n
    achieved = BYTES_PER_CELL_UPDATE * per_gpu_cells * K / (ms * 1e-3) / 1e9      # per GPU, algorithmic GB/s
    traffic = measured_traffic()

    # ---- end to end through the public API with HOST buffers: naz up (pinned), K steps, ez down (pinned)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(sim, world, rank, n, nx_global, K, T, src, sync)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nn, ss = 4096, 40                                   # ~13 s of single-core numpy: a bounded sample of the workload
        v, dt = cpu_numpy_port(nn, ss)
        cpu = {"value": v, "unit": "Mcell-updates/s", "cores": 1, "kind": "port",
               "sample": f"{nn}x{nn} fp32 npml={NPML}, {ss} steps after 1 warm-up ({dt:.1f} s); numpy oracle == "
                         f"fd2d/program/fd2d_3_2.py statements (single-threaded numpy), host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"fd2d TM+PML {nx_global}x{n} fp32 npml={NPML} point sinusoid 1500 MHz "
                                       f"(BASELINE config 5{'' if world == 1 else ', weak-scaled row slabs'})",
                           "grid_per_gpu": [n, n], "tblock": T, "parallelism": f"slab{world}",
                           "halo_exchange": getattr(sim, "halo_mode", "none"),
                           "l2": "inputs (52 GB per GPU) far exceed the 126 MB L2; no flush needed",
                           "timing": "CUDA events on the launch stream, barrier+sync both sides, max over ranks"},
                "gpu_launches": 2 * launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None if not traffic else traffic.get("dram_bytes_per_launch"),
                             "kernel": f"k_march<float,V=4,T={T},interior> (+ edge kernel)", "launches": launches,
                             "avg_launch_ms": ms / launches,
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL_UPDATE * per_gpu_cells * T,
                             "peak_source": peak_src,
                             "note": "achieved = 48 B x cells x steps / time; a T-step pass moves ~48 B per cell once, "
                                     "so frac may exceed 1 (traffic = real DRAM bytes per launch from ncu)"},
                "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(sim, world, rank, n, nx_global, K, T, src, sync):
    """Same metric through the user-facing call with host buffers: upload the medium (naz) from pinned host
    memory, run K steps, read Ez back into pinned host memory; all inside the timed region.  N > 1: every rank runs
    the same block wavefront on its slab with a K-row ghost band that is consumed instead of exchanged
    (communication-avoiding: K steps, no exchange; SlabFdtd2D.run_streamed)."""
    import gc
    import torch
    import torch.distributed as dist
    if world > 1:
        from simulation_b200 import slab
        sim.close()                                  # unmap the peers before this rank's arrays are freed
        del sim
        gc.collect()
        torch.cuda.empty_cache()
        sim = slab.SlabFdtd2D(nx_global, n, NPML, np.float32, source=src, tblock=T, ghost=K, halo="nccl")
        stored = sim.engine.rows_alloc
    else:
        stored = n
    rows = sim.row_hi - sim.row_lo
    host_naz = torch.ones((stored, n), dtype=torch.float32).pin_memory()
    host_ez = torch.empty((rows, n), dtype=torch.float32).pin_memory()

    def fresh():
        for name in ("dz", "hx", "hy", "ihx", "ihy", "ez"):
            sim.tensor(name, stored=True).zero_()
        sim.t = 0
    sim.run_streamed(K, host_naz, host_ez)                         # untimed warm-up of the same call: every pass-level stream is
                                                                   # created and used once (a stream's first use costs tens of ms)
    runs = []
    for _ in range(E2E_REPEATS):                                   # the whole job, E2E_REPEATS times; the median is reported
        fresh()                                                    # fresh problem: fields start at zero
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        sim.run_streamed(K, host_naz, host_ez)                     # transfers overlapped with the passes
        e1.record()
        sync()
        dt = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        runs.append(dt)
    dt = float(np.median(runs))
    return {"value": float(nx_global) * n * K / dt / 1e6, "unit": "Mcell-updates/s",
            "h2d_bytes_per_step": stored * n * 4 / K, "d2h_bytes_per_step": rows * n * 4 / K,
            "runs_ms": [round(x * 1e3, 2) for x in runs],
            "what": f"pinned naz H2D ({stored * n * 4 / 2**30:.1f} GiB/GPU) + {K} steps + pinned Ez D2H, CUDA events on the launch "
                    f"stream, max over ranks; run_streamed: row blocks uploaded in order (512..3072 rows), every block stepped through "
                    f"all its passes as soon as it has arrived (skewed space-time tiling), one stream per pass level, Ez of "
                    f"finished blocks downloaded behind the stepping (after one untimed run of the same call); "
                    f"median of {E2E_REPEATS} whole jobs (runs_ms)"
                    + ("" if world == 1 else f"; per rank a {K}-row ghost band consumed instead of exchanged (no communication in {K} steps)")}


def main():
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner on communicator
    # creation): park fd 1 on stderr for the whole run and emit the line on the saved descriptor at the end.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    lines = []
    try:
        _main(lines.append)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    for line in lines:
        print(line, flush=True)


def _main(emit):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=N_FULL, help="rows and columns per GPU (default: BASELINE config 5)")
    ap.add_argument("--tblock", type=int, default=6)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, emit)
    else:
        run_ours(args, emit)


if __name__ == "__main__":
    main()
