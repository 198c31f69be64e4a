#!/usr/bin/env python
"""Benchmark of the 2D TM+PML time-stepping hot path (BASELINE.json metric: Mcell-updates/s and HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full FDTD time step (D, E, Hx, Hy and both PML integrals of every cell).  Workload at
N GPUs: BASELINE config 5 -- a 32768-column fp32 grid with npml=80 and the point sinusoid of program 3_2, cut into
row slabs, ghost rows exchanged over NVLink every time block.  Default (--scaling weak): every GPU owns 32768 rows of
an (N*32768) x 32768 grid; --scaling strong: ONE 32768 x 32768 grid over the N GPUs (BASELINE.json configs[4] as
worded).  At N > 1 the weak run also measures the strong split and reports it under "configs".  One JSON line on
stdout (rank 0).

At N > 1 the line carries "parity_check": before the timed region the N ranks run a small slab problem with the
kernels and launch plan of the bench (4-wide vectors, 128-row chunks, fused peer-store halo exchange) and every rank
compares its rows of all six arrays, byte for byte, with a single-device run of the same problem; a mismatch exits
non-zero.

--impl reference times the REFERENCE's own C/OpenMP step functions (oracle/_ref, compiled from
/root/reference/fd2d/clang/test_3_2.c; falls back to the numpy oracle port when absent) on all host cores, on the
full 32768 x 32768 grid when the host has the memory for it (28 GiB), else on an 8192 x 8192 sample.  The oracle is
used only there and in the cpu_baseline leg -- never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL = 32768
NPML = 80
BYTES_PER_CELL_UPDATE = 48          # dz,hx,hy,ihx,ihy R+W; naz R; ez W (fp32) -- SURVEY.md 8(d)
BYTES_LOSSY = 60                    # + iz R+W, nbz R (program 3_4, DFT off)
BYTES_1D_LOSSY = 24                 # ex,hy R+W; ca,cb R
METRIC = "Mcell-updates/s, 2D TM+PML fp32"
E2E_REPEATS = 3                     # the end-to-end job is a single ~0.1 s shot whose overlap depends on launch timing: median of 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """DRAM bytes per pass of the dominant kernel from the committed ncu --set full captures, by pass depth."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def traffic_for(depths, writes_ez=True):
    """DRAM bytes one advance() with these pass depths moves over 32768^2 cells, from the committed ncu captures
    -> (bytes or None, bytes of the deepest pass or None, provenance string)"""
    t = measured_traffic()
    if not t:
        return None, None, "no committed ncu capture"
    by = t.get("by_depth", {})
    if "6" not in by and "dram_bytes_per_launch" in t:
        by = dict(by, **{"6": {"dram_bytes_per_pass": t["dram_bytes_per_launch"], "source": t.get("source", "profiles/roofline_traffic.json")}})
    total, used = 0.0, {}
    for d in depths:
        per = by.get(str(d)) or (by.get("6") if d < 6 else None)       # a shallow pass moves the state once as well
        if per is None:
            return None, None, f"no ncu capture for pass depth {d} in profiles/roofline_traffic.json"
        total += float(per["dram_bytes_per_pass"]) + float(t.get("careful_bytes_per_pass", {}).get(str(d if str(d) in by else 6), 0.0))
        used[str(d if str(d) in by else 6)] = per
    if writes_ez:
        total += float(t.get("ez_store_bytes_last_pass", 0.0))
    top = by.get(str(max(depths))) or by.get("6")
    src = "; ".join(f"depth {k}: {v.get('source', '?')}, commit {v.get('commit', t.get('commit', '?'))}" for k, v in sorted(used.items()))
    return total, float(top["dram_bytes_per_pass"]), "from profiles/roofline_traffic.json -- " + src


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period)."""
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


def host_mem_available_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 2**20
    except Exception:
        pass
    return 0.0


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


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: ASSIGN the variable before libgomp reads it, set the count
    through the OpenMP API as well, and return what the runtime will really use (omp_get_max_threads)."""
    import ctypes as C
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        gomp.omp_set_dynamic(0)
        gomp.omp_set_num_threads(cores)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return cores


def cpu_reference_c(nx, ny, steps, warm=1):
    """The reference's own C/OpenMP dfield/efield/hfield (oracle/_ref/libref_fd2d_3_2.so), all host threads.
    -> (Mcell/s, seconds, OpenMP threads used)"""
    import ctypes as C
    from oracle import fdtd_oracle as orc
    from oracle import ref_c
    threads = use_all_host_threads()
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
    return nx * ny * steps / dt / 1e6, dt, threads


def run_reference(args, emit=print):
    """--impl reference: rank 0 only.  The full 32768 x 32768 grid of the GPU arm's N=1 workload (28 GiB of host
    arrays) when the host has the memory, else an 8192 x 8192 sample of it; the step count is capped so that the run
    ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_c
    steps, warm = min(max(1, args.steps), 40), min(max(1, args.warmup), 3)
    n = args.size
    have = host_mem_available_gib()
    need = 7 * n * n * 4 / 2**30 * 1.15
    why = ""
    if have and have < need:
        why = f" (host has {have:.0f} GiB available, the {n}x{n} grid needs {need:.0f})"
        n = 8192
    if ref_c.available("3_2"):
        v, dt, cores = cpu_reference_c(n, n, steps, warm)
        kind, how = "reference", "reference C/OpenMP step functions (fd2d/clang/test_3_2.c) via oracle/_ref"
    else:
        n, steps = min(n, 4096), min(steps, 10)
        v, dt = cpu_numpy_port(n, steps)
        cores, kind, how = 1, "port", "numpy oracle port (oracle/_ref not built)"
    same = n == N_FULL
    sample = (f"the whole {n}x{n} fp32 grid" if same else f"{n}x{n} fp32 sub-grid of the {N_FULL}x{N_FULL} workload{why}") + \
        f", {steps} steps after {warm} warm-up, {cores} OpenMP threads (omp_get_max_threads; host has {os.cpu_count()} CPUs); {how}"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mcell-updates/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"fd2d TM+PML {N_FULL}x{N_FULL} fp32 npml={NPML} point sinusoid 1500 MHz (BASELINE config 5)",
                       "grid": [n, n], "same_config": same, "sample": sample,
                       "note": "the CPU arm always runs the one-GPU grid: at N > 1 the GPU arm's weak-scaled grid "
                               "(N x 28 GiB of host arrays) does not fit a host, and throughput per cell does not depend on it"},
            "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def _timed(fn, sync, world):
    """CUDA events on the launch stream around fn(), barrier + synchronize on both sides, max over ranks -> ms"""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    fn()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def parity_check(world, rank, T):
    """N-rank slab run (fused peer-store halo exchange, the bench's kernels: 4-wide vectors, 128-row chunks, depth T)
    against a single-device run of the same problem on EVERY rank's own GPU: the rank's rows of all six arrays must
    be byte-identical.  -> dict for the JSON line (rank 0) / None; raises SystemExit on mismatch (all ranks)."""
    import torch
    import torch.distributed as dist
    from simulation_b200 import _lib, fd2d, slab, surface
    nx, ny, npml, ns = 1024 * world, 1536, 40, 61
    rng = np.random.default_rng(20261018)
    naz = rng.uniform(0.25, 1.0, size=(nx, ny)).astype(np.float32)      # random medium: every cell matters
    src = fd2d.PointSource(nx // 2 - 5, ny // 2 - 5, surface.Sinusoid(1500e6), hard=True)
    _lib.lib().fdtd2d_tune(4, 128, 0, 0, 0)                              # the >= 100 M-cell launch plan, on a small grid
    try:
        s = slab.SlabFdtd2D(nx, ny, npml, np.float32, tblock=T, source=src, naz=naz, halo="p2p")
        s.advance(7)                       # ragged split: several calls, several epochs of the flag handshake
        s.advance(ns - 7)
        s.synchronize()
        one = fd2d.Fdtd2D(nx, ny, npml, np.float32, source=src, naz=naz, tblock=T)
        one.advance(ns)
        one.synchronize()
        names = ("dz", "ez", "hx", "hy", "ihx", "ihy")
        h, equal, peak = hashlib.sha256(), True, 0.0
        for name in names:
            mine = s.tensor(name).contiguous()
            ref = one.tensor(name)[s.row_lo:s.row_hi].contiguous()
            same = torch.equal(mine.view(torch.int32), ref.view(torch.int32))     # bit view: -0.0 != +0.0
            equal = equal and bool(same)
            h.update(mine.cpu().numpy().tobytes())
            if name == "ez":
                peak = float(ref.abs().max().item())
        mode, exchanges = s.halo_mode, s.exchanges
        digests = [None] * world
        dist.all_gather_object(digests, (bool(equal), h.hexdigest(), peak))
        s.close()
        del s, one
    finally:
        _lib.lib().fdtd2d_tune(0, 0, 0, 0, 0)
    all_equal = all(d[0] for d in digests)
    out = {"ranks": world, "halo": mode, "equal": bool(all_equal and mode == "p2p"), "grid": [nx, ny], "steps": ns,
           "tblock": T, "plan": "V=4, 128-row chunks (fdtd2d_tune), random medium, point sinusoid",
           "exchanges": exchanges, "arrays": list(names), "peak_abs_ez": max(d[2] for d in digests),
           "against": "a single-device run of the same problem on every rank's own GPU, owned rows, bit view",
           "sha256": hashlib.sha256("".join(d[1] for d in digests).encode()).hexdigest(),
           "per_rank_equal": [d[0] for d in digests]}
    if not out["equal"]:
        if rank == 0:
            sys.stderr.write("bench.py: N-GPU PARITY CHECK FAILED: " + json.dumps(out) + "\n")
        raise SystemExit(3)
    return out


def run_ours(args, emit=print):
    # run_streamed (the e2e leg) drives one stream per pass level: give the driver enough hardware queues that they do
    # not alias (default 8).  Must be set before the CUDA context exists.
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import gc
    import torch
    import torch.distributed as dist
    from simulation_b200 import _lib, fd2d, surface

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
        _lib.lib().fdtd2d_tune(*[int(os.environ.get(k, "0")) for k in knobs])
    for key, env in ((_lib.TUNE_DEEP, "FDTD_DEEP"), (_lib.TUNE_VARIANT, "FDTD_VARIANT"), (_lib.TUNE_EDGE_CHUNKS, "FDTD_EDGE_CHUNKS"), (_lib.TUNE_COL_FAST, "FDTD_COL_FAST")):
        if env in os.environ:
            _lib.lib().fdtd2d_tune2(key, int(os.environ[env]))
    barrier = dist.barrier if world > 1 else (lambda: None)

    def sync():
        barrier()
        torch.cuda.synchronize()

    def release(sim):
        if hasattr(sim, "close"):
            sim.close()                          # unmap the peers before this rank's arrays are freed
        del sim
        gc.collect()
        torch.cuda.empty_cache()

    # ---- N > 1: the ranks must agree with one device, bit for bit, before anything is timed
    parity = parity_check(world, rank, T if T else 6) if (world > 1 and not args.no_parity) else None

    wave = surface.Sinusoid(1500e6)
    peak, peak_src = peaks()

    def measure(nx_global, steps, warm):
        """K steps device-resident on an nx_global x n grid cut into `world` row slabs -> (sim, ms, depths)"""
        src = fd2d.PointSource(nx_global // 2 - 5, n // 2 - 5, wave, hard=True)
        if world > 1:
            from simulation_b200 import slab
            sim = slab.SlabFdtd2D(nx_global, n, NPML, np.float32, source=src, tblock=T)
        else:
            sim = fd2d.Fdtd2D(nx_global, n, NPML, np.float32, source=src, tblock=T)
        sim.advance(warm)
        sync()
        ms = _timed(lambda: sim.advance(steps), sync, world)
        eng = getattr(sim, "engine", sim)
        return sim, src, ms, eng.pass_depths(steps, T)

    strong = args.scaling == "strong"
    nx_global = n if strong else n * world
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sim, src, ms, depths = measure(nx_global, K, W)
    clocks = sampler.stop() if rank == 0 else None
    launches = len(depths)                       # passes; each pass = interior kernel + edge kernel (+ the column variant of
    kernels = sum(3 if d >= 8 else 2 for d in depths)   # the interior kernel for the PML-column strips in the deep passes)
    cells = float(nx_global) * n
    value = cells * K / (ms * 1e-3) / 1e6
    per_gpu_cells = cells / world
    achieved = BYTES_PER_CELL_UPDATE * per_gpu_cells * K / (ms * 1e-3) / 1e9      # per GPU, algorithmic GB/s
    job_bytes, traffic, traffic_src = traffic_for(depths)
    dram_frac = None
    if job_bytes is not None:
        # real DRAM bytes of the K-step job (ncu captures per pass depth, interior + edge kernel, + the one ez store;
        # measured over 32768^2 cells, scaled to this rank's cells) / time / measured copy peak
        dram_frac = job_bytes * (per_gpu_cells / (float(N_FULL) * N_FULL)) / (ms * 1e-3) / 1e9 / peak
    halo_mode = getattr(sim, "halo_mode", "none")

    # ---- end to end through the public API with HOST buffers: naz up (pinned), K steps, ez down (pinned)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(sim, world, rank, n, nx_global, K, T, src, sync, release)
    else:
        release(sim)
    sim = None

    # ---- the other BASELINE configs, one short measurement each (same timing discipline)
    configs = None
    if not args.no_configs:
        configs = {}
        if world > 1 and not strong:
            s2, _, ms2, d2 = measure(n, K, W)                    # configs[4] as worded: ONE 32768^2 grid over the N GPUs
            configs["c5_strong_32768_over_N"] = {
                "value": float(n) * n * K / (ms2 * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms2 / K,
                "scaling": "strong", "grid": [n, n], "rows_per_gpu": n // world, "pass_depths": d2,
                "halo_exchange": getattr(s2, "halo_mode", "none"),
                "frac": BYTES_PER_CELL_UPDATE * float(n) * n / world * K / (ms2 * 1e-3) / 1e9 / peak,
                "note": "frac = algorithmic 48 B x this GPU's cells x steps / time / measured copy peak"}
            release(s2)
        if rank == 0:
            configs.update(other_configs(peak, sync=lambda: torch.cuda.synchronize()))
        if world > 1:
            dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nn, ss = 4096, 40                                   # ~13 s of single-core numpy: a bounded sample of the workload
        v, dt = cpu_numpy_port(nn, ss)
        cpu = {"value": v, "unit": "Mcell-updates/s", "cores": 1, "kind": "port",
               "sample": f"{nn}x{nn} fp32 npml={NPML}, {ss} steps after 1 warm-up ({dt:.1f} s); numpy oracle == "
                         f"fd2d/program/fd2d_3_2.py statements (single-threaded numpy), host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"fd2d TM+PML {nx_global}x{n} fp32 npml={NPML} point sinusoid 1500 MHz "
                                       f"(BASELINE config 5{'' if world == 1 else (', one grid over row slabs' if strong else ', weak-scaled row slabs')})",
                           "grid_per_gpu": [nx_global // world, n], "tblock": T, "pass_depths": depths,
                           "parallelism": f"slab{world}", "halo_exchange": halo_mode,
                           "l2": "inputs (52 GB per GPU) far exceed the 126 MB L2; no flush needed" if per_gpu_cells >= 2**28
                                 else "state per GPU exceeds the 126 MB L2 (>= 6 GB); no flush needed",
                           "timing": "CUDA events on the launch stream, barrier+sync both sides, max over ranks"},
                "gpu_launches": kernels,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "dram_frac": dram_frac, "traffic": traffic, "traffic_source": traffic_src,
                             "kernel": f"interior kernel of a pass, fp32 V=4 (depth 8 / 12: k_march_chain, TMA-fed warp chain; depth <= 6: k_march register pipeline), "
                                       f"pass depths {sorted(set(depths), reverse=True)} (+ edge kernel)",
                             "launches": launches, "avg_launch_ms": ms / launches,
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL_UPDATE * per_gpu_cells * K / launches,
                             "peak_source": peak_src,
                             "note": "frac = 48 B x cells x steps / time / peak (SURVEY 8d): a T-step pass moves the state "
                                     "through HBM once, so frac may exceed 1; dram_frac = real DRAM bytes per pass (ncu capture "
                                     "named in traffic_source) x passes / time / peak -- the physical utilisation"},
                "clocks": clocks, "e2e": e2e, "cpu_baseline": cpu}
        if parity is not None:
            line["parity_check"] = parity
        if configs:
            line["configs"] = configs
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(peak, sync):
    """BASELINE configs 2 and 4 on one GPU (rank 0): device-resident, CUDA events, 3+ warm-up passes."""
    import torch
    from simulation_b200 import fd1d, fd2d, surface
    out = {}

    def timed(sim, steps, warm):
        sim.advance(warm)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.advance(steps)
        e1.record()
        sync()
        return e0.elapsed_time(e1)

    # config 2: 1D lossy slab, 1e6 cells, soft sinusoid 700 MHz, ABC
    nx, ns = 1_000_000, 10_000
    ca, cb = surface.dielectric_fdtd(nx, surface.DT, 4.0, 0.04, np.float32, start=nx // 2, stop=nx // 2 + nx // 4)
    line = fd1d.Fdtd1D(nx, np.float32, source=fd1d.LineSource(1, surface.Sinusoid(700e6)), ca=ca, cb=cb)
    ms = timed(line, ns, 256)
    out["c2_1d_lossy_1e6"] = {"value": nx * ns / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms / ns,
                              "steps": ns, "frac": BYTES_1D_LOSSY * nx * ns / (ms * 1e-3) / 1e9 / peak, "dram_frac": None,
                              "note": "24 B per cell-update algorithmic; the 24 MB state lives in L2 and a pass is 32 steps, "
                                      "so HBM is not the bound (instruction-bound, DESIGN 4.3)"}
    del line
    # config 4: 4096^2 TFSF plane wave + lossy dielectric cylinder (radius 6 m -> 599 cells), npml 80, DFT off
    n, ns = 4096, 2000
    md = fd2d.dielectric(n, n, NPML, int(6.0 / surface.DS - 1), surface.DT, 30.0, 0.30, np.float32)
    grid = fd2d.Fdtd2D(n, n, NPML, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=md.naz, nbz=md.nbz)
    ms = timed(grid, ns, 48)
    # the same job end to end with HOST buffers: the medium (naz, nbz) up from pinned memory, ns steps from zero fields,
    # Ez back into pinned memory -- what the reference main() of fd2d/python/fd2d_3_4.py does around its loop
    h_naz, h_nbz = md.naz.cpu().pin_memory(), md.nbz.cpu().pin_memory()
    h_ez = torch.empty((n, n), dtype=torch.float32).pin_memory()
    job = fd2d.Fdtd2D(n, n, NPML, np.float32, source=fd2d.IncidentWave(surface.Gaussian(20, 8.0)), naz=md.naz, nbz=md.nbz)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job.naz.copy_(h_naz, non_blocking=True)
    job.nbz.copy_(h_nbz, non_blocking=True)
    job.invalidate_lossy_box()
    job.advance(ns)
    h_ez.copy_(job.tensor("ez"), non_blocking=True)
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1)
    out["c4_tfsf_lossy_4096"] = {"value": float(n) * n * ns / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms / ns,
                                 "steps": ns, "frac": BYTES_LOSSY * float(n) * n * ns / (ms * 1e-3) / 1e9 / peak, "dram_frac": None,
                                 "pass_depths": sorted(set(grid.pass_depths(ns, None)), reverse=True),
                                 "e2e": {"value": float(n) * n * ns / (ms_e2e * 1e-3) / 1e6, "unit": "Mcell-updates/s",
                                         "h2d_bytes_per_step": 2 * n * n * 4 / ns, "d2h_bytes_per_step": n * n * 4 / ns,
                                         "what": "pinned naz + nbz H2D, 2000 steps, pinned Ez D2H (192 MiB over PCIe against 0.1 s "
                                                 "of stepping: copies around advance(), no streaming needed)"},
                                 "note": "60 B per cell-update algorithmic (iz R+W, nbz R on top of 48); the 0.4 GB state is "
                                         "3x the L2, launches are a few waves"}
    del grid, md, job
    torch.cuda.empty_cache()
    return out


def run_e2e(sim, world, rank, n, nx_global, K, T, src, sync, release):
    """Same metric through the user-facing call with host buffers: upload the medium (naz) from pinned host
    memory, run K steps, read Ez back into pinned host memory; all inside the timed region.  N > 1: every rank runs
    the same block schedule on its slab with a K-row ghost band that is consumed instead of exchanged
    (communication-avoiding: K steps, no exchange; SlabFdtd2D.run_streamed)."""
    import torch
    import torch.distributed as dist
    if world > 1:
        from simulation_b200 import slab
        release(sim)
        sim = slab.SlabFdtd2D(nx_global, n, NPML, np.float32, source=src, tblock=T, ghost=K, halo="nccl")
        stored = sim.engine.rows_alloc
    else:
        stored = nx_global
    rows = sim.row_hi - sim.row_lo
    host_naz = torch.ones((stored, n), dtype=torch.float32).pin_memory()
    host_ez = torch.empty((rows, n), dtype=torch.float32).pin_memory()

    def fresh():
        for name in ("dz", "hx", "hy", "ihx", "ihy", "ez"):
            sim.tensor(name, stored=True).zero_()
        sim.t = 0
    sim.run_streamed(K, host_naz, host_ez)                         # untimed warm-up of the same call: every pass-level stream is
                                                                   # created and used once (a stream's first use costs tens of ms)
    runs, mine = [], []
    for _ in range(E2E_REPEATS):                                   # the whole job, E2E_REPEATS times; the median is reported
        fresh()                                                    # fresh problem: fields start at zero
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        sim.run_streamed(K, host_naz, host_ez)                     # transfers overlapped with the passes
        e1.record()
        sync()
        dt = e0.elapsed_time(e1) * 1e-3
        mine.append(dt)
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        runs.append(dt)
    dt = float(np.median(runs))
    h2d, d2h = stored * n * 4, rows * n * 4
    # what limits it: every rank's own PCIe rate during its median run (its bytes each way / its own time)
    rate = (h2d + d2h) / float(np.median(mine)) / 1e9
    rates = [rate]
    if world > 1:
        rates = [None] * world
        dist.all_gather_object(rates, rate)
    out = {"value": float(nx_global) * n * K / dt / 1e6, "unit": "Mcell-updates/s",
           "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "runs_ms": [round(x * 1e3, 2) for x in runs],
           "pcie_gbs_per_rank": [round(float(r), 1) for r in rates],
           "pcie_bound_ms": round(max(h2d, d2h) / 44e9 * 1e3, 1),
           "what": f"pinned naz H2D ({h2d / 2**30:.1f} GiB/GPU) + {K} steps + pinned Ez D2H, CUDA events on the launch "
                   f"stream, max over ranks; run_streamed: row blocks uploaded in order (512..3072 rows), every block stepped through "
                   f"all its passes as soon as it has arrived (skewed space-time tiling), one stream per pass level, Ez of "
                   f"finished blocks downloaded behind the stepping (after one untimed run of the same call); "
                   f"median of {E2E_REPEATS} whole jobs (runs_ms); pcie_gbs_per_rank = (H2D + D2H bytes) / the rank's own time, "
                   f"pcie_bound_ms = either direction at the 44 GB/s these boxes sustain when BOTH directions are busy (55 GB/s alone; profiles/r1_box_topology_and_pcie.txt)"
                   + ("" if world == 1 else f"; per rank a {K}-row ghost band consumed instead of exchanged (no communication in {K} steps)")}
    release(sim)
    return out


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 32768 rows per GPU; strong: one 32768 x 32768 grid over the N GPUs")
    ap.add_argument("--size", type=int, default=N_FULL, help="rows (per GPU when weak) and columns (default: BASELINE config 5)")
    ap.add_argument("--tblock", type=int, default=0, help="pass depth (0: chosen by grid size and step count)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE configs 2, 4 and the strong split")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-GPU parity check (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args, emit)
    else:
        run_ours(args, emit)


if __name__ == "__main__":
    main()
