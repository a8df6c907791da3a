#!/usr/bin/env python
"""
bench.py -- throughput of the aMOF hot path on B200 (BASELINE.json metric: RDF frames/s and pair-evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--frames F] [--impl reference]

One step = one pass of the analysis over the whole synthetic trajectory of the workload:
  c2 (default)  9 792 atoms x 10 000 frames, all 16 partial RDFs (dr 0.01, rmax 10 -> 999 bins) + CN Zn-N/C-N/C-C
  c3            104 448-atom triclinic box x 2 000 frames, same analysis
  c4            48 960 atoms x 5 000 frames, bond angles N-Zn-N (dtheta 0.05)
  c5            979 200 atoms x 5 000 frames window MSD (needs ~120 GB of HBM; generated on the device)
`value`  frames/s of the whole job with the trajectory already resident in HBM (amofb_*_push_device),
`e2e`    the same through the public classes (amof_b200.rdf.rdf_and_cn / bad.Bad / msd.WindowMsd) from page-locked
         HOST arrays, host->device copies and the result read-back inside the timed region.
Multi-GPU (torchrun, one rank per GPU): frames are independent, every rank runs the same per-rank workload
(weak scaling) with no data-path collective; the histograms are all-reduced once per step.
Timing: CUDA events on the library's compute stream around the K timed steps, max over ranks.
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

CN_SETS = {'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}
# dram__bytes_read.sum + dram__bytes_write.sum of one k_pair_tiled launch over 107 frames of C2, ncu --set full,
# profiles/r01_k_pair_tiled_ncu_full_summary.txt: 56.24 MB + 2.35 MB (the kernel reads the 32-byte cell-sorted records);
# kept PER FRAME and scaled to the frames of a bench launch
PAIR_TRAFFIC_PER_FRAME = {"c2": (56.24e6 + 2.35e6) / 107.0}
FP64_NOFMA_GOPS = 18515.3      # measured on this pool's B200 with tools/microbench.cu (gpurun_out/microbench.json)


# ------------------------------------------------------------------------------------------------ utilities
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(ctx, world):
    ctx.sync()
    if world > 1:
        import torch
        import torch.distributed as td
        td.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as td
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ workloads
class PairWorkload:
    """c2 / c3: partial RDFs + coordination numbers."""

    def __init__(self, name, frames):
        from amof_b200 import synth
        self.name = name
        self.T = frames or synth.CONFIGS[name]["frames"]
        self.rmax, self.dr = 10.0, 0.01
        self.bins = int(self.rmax // self.dr)
        self.metric, self.unit, self.dtype = "rdf_cn_frames_per_s", "frames/s", "f64"

    def describe(self):
        return {"workload": "%s: synthetic a-ZIF %d atoms x %d frames, 16 partial RDFs (dr 0.01, rmax 10 -> %d bins) + CN %s"
                            % (self.name, self.n_atoms, self.T, self.bins, "/".join(CN_SETS)),
                "atoms": self.n_atoms, "frames_per_step": self.T, "l2": "inputs (%.2f GB/step) larger than L2" % (self.bytes_in / 1e9)}

    def setup(self, backend):
        from amof_b200 import atom as amatom, frames as fr, synth
        from amof_b200.frames import ArrayTrajectory
        ctx = backend.ctx
        numbers, _, cell = synth.base_frame(self.name)
        self.n_atoms = len(numbers)
        self.host = ctx.pinned_empty((self.T, self.n_atoms, 3))
        synth.fill_frames(self.name, 0, self.T, self.host)
        self.traj = ArrayTrajectory(numbers, self.host, cell, pinned=True)
        self.zs, self.spec = fr.species_index(numbers)
        self.cut = amatom.cutoff_matrix(amatom.format_cutoff(CN_SETS), self.zs)
        self.bytes_in = self.host.nbytes + self.T * 72
        self.dev = ctx.device_alloc(self.host.nbytes)
        ctx.h2d(self.dev, self.host)
        self.backend = backend
        self.chunk = max(1, (256 << 20) // (24 * self.n_atoms))

    def _device_chunks(self):
        fb = 24 * self.n_atoms
        for a in range(0, self.T, self.chunk):
            b = min(self.T, a + self.chunk)
            yield self.dev.value + a * fb, self.traj.cells[a:b]

    def step_resident(self):
        res = self.backend.pair_counts(self.spec, len(self.zs), self._device_chunks(), rmax=self.rmax, nbins=self.bins,
                                       cn_cutoff=self.cut)
        if self.world > 1:      # the one collective of the path: integer all-reduce of the histograms (NCCL)
            from amof_b200 import _dist
            self.total_hist = _dist.allreduce_sum(res["hist"])
        self.last = res
        return res

    def step_e2e(self):
        from amof_b200 import rdf
        r, c = rdf.rdf_and_cn(self.traj, CN_SETS, dr=self.dr, rmax=self.rmax, distributed=False)
        self.d2h = r.counts.nbytes + c.counts.nbytes
        return r, c

    def units(self, res):
        pairs = int(res["hist"].sum()) // 2
        return {"pair_evals_per_step": pairs}

    def algorithmic_bytes_per_frame(self):
        return 24 * self.n_atoms + 72

    def cpu_sample(self, threads, frames):
        from oracle import c_oracle as orc
        pos, cells = self.host[:frames], self.traj.cells[:frames]
        t0 = time.perf_counter()
        orc.rdf_traj(pos, cells, self.spec, len(self.zs), self.rmax, self.bins, threads=threads)
        orc.cn_traj(pos, cells, self.spec, len(self.zs), self.cut, threads=threads)
        return time.perf_counter() - t0


class BadWorkload:
    """c4: N-Zn-N bond-angle distribution."""

    def __init__(self, name, frames):
        from amof_b200 import synth
        self.name, self.T = name, frames or synth.CONFIGS[name]["frames"]
        self.metric, self.unit, self.dtype = "bad_frames_per_s", "frames/s", "f64"
        self.dtheta = 0.05

    def describe(self):
        return {"workload": "%s: synthetic a-ZIF %d atoms x %d frames, Bad({'Zn-N': 2.5}), dtheta 0.05 -> 3600 bins"
                            % (self.name, self.n_atoms, self.T), "atoms": self.n_atoms, "frames_per_step": self.T,
                "l2": "inputs (%.2f GB/step) larger than L2" % (self.bytes_in / 1e9)}

    def setup(self, backend):
        from amof_b200 import atom as amatom, frames as fr, synth
        from amof_b200.frames import ArrayTrajectory
        ctx = backend.ctx
        numbers, _, cell = synth.base_frame(self.name)
        self.n_atoms = len(numbers)
        self.host = ctx.pinned_empty((self.T, self.n_atoms, 3))
        synth.fill_frames(self.name, 0, self.T, self.host)
        self.traj = ArrayTrajectory(numbers, self.host, cell, pinned=True)
        self.zs, self.spec = fr.species_index(numbers)
        self.cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5}), self.zs)
        self.triples = [(self.zs.index(30), self.zs.index(7)), (self.zs.index(7), self.zs.index(30))]
        self.nbins = int(180 // self.dtheta) + 1
        self.bytes_in = self.host.nbytes + self.T * 72
        self.dev = ctx.device_alloc(self.host.nbytes)
        ctx.h2d(self.dev, self.host)
        self.backend = backend
        self.chunk = max(1, (256 << 20) // (24 * self.n_atoms))

    def _device_chunks(self):
        fb = 24 * self.n_atoms
        for a in range(0, self.T, self.chunk):
            b = min(self.T, a + self.chunk)
            yield self.dev.value + a * fb, self.traj.cells[a:b]

    def step_resident(self):
        hist, dropped, nf = self.backend.bad_counts(self.spec, len(self.zs), self._device_chunks(), self.cut, self.triples,
                                                    self.dtheta, self.nbins)
        if self.world > 1:
            from amof_b200 import _dist
            self.total_hist = _dist.allreduce_sum(hist)
        self.last = {"hist": hist}
        return self.last

    def step_e2e(self):
        from amof_b200 import bad
        b = bad.Bad.from_trajectory(self.traj, {'Zn-N': 2.5}, dtheta=self.dtheta, distributed=False)
        self.d2h = 2 * 33 * self.nbins * 8
        return b

    def units(self, res):
        return {"angles_per_step": int(res["hist"].sum())}

    def algorithmic_bytes_per_frame(self):
        return 24 * self.n_atoms + 72

    def cpu_sample(self, threads, frames):
        from oracle import c_oracle as orc
        t0 = time.perf_counter()
        for f in range(frames):
            for (A, B) in self.triples:
                orc.bad_hist(self.host[f], self.traj.cells[f], self.spec, len(self.zs), self.cut, A, B, self.dtheta, self.nbins)
        return time.perf_counter() - t0


class MsdWorkload:
    """c5: window MSD of an unwrapped random walk; the trajectory is generated on the device (117.5 GB at full
    size would not fit in host memory), the e2e leg streams a host-resident slab of frames repeatedly."""

    def __init__(self, name, frames, atoms=None):
        from amof_b200 import synth
        self.name, self.T = name, frames or synth.CONFIGS[name]["frames"]
        self.metric, self.unit, self.dtype = "msd_frames_per_s", "frames/s", "f64"
        self.atoms_override = atoms

    def describe(self):
        return {"workload": "%s: synthetic %d atoms x %d frames unwrapped random walk, WindowMsd(delta_time=100) -> %d windows"
                            % (self.name, self.n_atoms, self.T, len(self.window)), "atoms": self.n_atoms,
                "frames_per_step": self.T, "l2": "inputs (%.2f GB/step) larger than L2" % (self.bytes_in / 1e9)}

    def setup(self, backend):
        import torch
        from amof_b200 import frames as fr, synth
        from amof_b200.elements import atomic_masses
        numbers, pos0, cell = synth.base_frame(self.name)
        if self.atoms_override:
            numbers, pos0 = numbers[:self.atoms_override], pos0[:self.atoms_override]
        self.n_atoms = len(numbers)
        self.zs, self.spec = fr.species_index(numbers)
        self.masses = np.array([atomic_masses[z] for z in numbers])
        self.cells = np.broadcast_to(cell, (self.T, 3, 3)).copy()
        self.window = np.arange(0, self.T // 2, 100)
        self.bytes_in = self.T * self.n_atoms * 24
        self.backend = backend
        self.slab = max(1, min(self.T, (1 << 30) // (24 * self.n_atoms)))      # frames per generated slab
        dev = torch.device("cuda", backend.ctx.device)
        self.torch, self.dev = torch, dev
        self.base = torch.from_numpy(pos0).to(dev)
        self.host_slab = backend.ctx.pinned_empty((self.slab, self.n_atoms, 3))
        g = torch.Generator(device=dev); g.manual_seed(synth.SEED0 + 5)
        self.host_slab[...] = (self.base[None] + 0.05 * torch.randn((self.slab, self.n_atoms, 3), generator=g, device=dev,
                                                                    dtype=torch.float64).cumsum(0)).cpu().numpy()

    def _fill_device(self, session):
        """random walk generated slab by slab on the device, handed over with amofb_msd_load_device"""
        torch = self.torch
        g = torch.Generator(device=self.dev); g.manual_seed(20261023)
        cur = self.base.clone()
        for a in range(0, self.T, self.slab):
            b = min(self.T, a + self.slab)
            inc = 0.05 * torch.randn((b - a, self.n_atoms, 3), generator=g, device=self.dev, dtype=torch.float64)
            inc[0] += cur
            slab = inc.cumsum(0)
            cur = slab[-1].clone()
            torch.cuda.synchronize()
            session.load(a, (slab.data_ptr(), b - a))
            self.backend.ctx.sync()

    def _analyse(self, s):
        """atoms are sharded over ranks: the per-frame mass-weighted sums and the window sums are all-reduced"""
        from amof_b200 import _dist
        sums = _dist.allreduce_sum(s.com_sums())
        s.set_com(sums[:, :3] / sums[:, 3:4])
        return _dist.allreduce_sum(s.window(self.window.astype(np.int32)))

    def step_resident(self):
        # the device-resident leg times load_device (transpose) + COM + prepare + window; generation is outside
        raise NotImplementedError

    def units(self, res):
        return {"atom_frame_pairs_per_step": int(self.n_atoms * sum(self.T - m - 1 for m in self.window))}

    def algorithmic_bytes_per_frame(self):
        return 24 * self.n_atoms

    def cpu_sample(self, threads, frames):
        from oracle import c_oracle as orc
        n = min(self.n_atoms, 20000)
        T = min(frames, self.slab)
        pos = self.host_slab[:T, :n].copy()
        w = np.arange(0, T // 2, max(1, T // 50))
        t0 = time.perf_counter()
        orc.msd_window(pos, self.cells[:T], self.masses[:n], self.spec[:n], len(self.zs), w)
        dt = time.perf_counter() - t0
        # scale per atom.frame-pair to the full workload
        done = n * sum(T - m - 1 for m in w)
        full = self.n_atoms * sum(self.T - m - 1 for m in self.window)
        return dt * full / max(done, 1) * (frames / self.T)


# ------------------------------------------------------------------------------------------------ main
def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    if local and "AMOFB_DEVICE" not in os.environ:
        os.environ["AMOFB_DEVICE"] = str(local)
    from amof_b200 import _lib
    backend = _lib.get_backend()
    ctx = backend.ctx
    if args.workload in ("c2", "c3"):
        wl = PairWorkload(args.workload, args.frames)
    elif args.workload == "c4":
        wl = BadWorkload(args.workload, args.frames)
    else:
        return run_msd(args, backend, rank, world)
    wl.world = world
    wl.setup(backend)

    # ---- device-resident leg -------------------------------------------------------------------
    for _ in range(args.warmup):
        res = wl.step_resident()
    ctx.set_profiling(True)
    ctx.pair_kernel_time(reset=True)
    sampler = ClockSampler(local)
    barrier_sync(ctx, world)
    sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = wl.step_resident()
    ctx.timer_mark(1)
    barrier_sync(ctx, world)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = max_over_ranks(ctx.timer_elapsed(0, 1), world)
    launches = ctx.launch_count() - l0
    k_ms, k_n = ctx.pair_kernel_time(reset=True)
    ctx.set_profiling(False)

    # ---- end-to-end leg: public classes, page-locked host input, H2D and result D2H inside the timed region ----
    for _ in range(min(args.warmup, 2)):
        wl.step_e2e()
    barrier_sync(ctx, world)
    t1 = time.perf_counter()
    for _ in range(args.steps):
        wl.step_e2e()
    barrier_sync(ctx, world)
    e2e_s = max_over_ranks(time.perf_counter() - t1, world)

    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    frames_total = wl.T * args.steps * world
    value = frames_total / (dev_ms / 1e3)
    out = {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": wl.dtype, "data": "synthetic", "config": dict(wl.describe(), parallelism="frames x%d" % world),
        "clocks": clocks, "gpu_launches": int(launches), "wall_ms_per_step": wall / args.steps * 1e3,
        "e2e": {"value": frames_total / e2e_s, "unit": wl.unit, "h2d_bytes_per_step": int(wl.bytes_in),
                "d2h_bytes_per_step": int(wl.d2h), "api": "amof_b200.rdf.rdf_and_cn" if isinstance(wl, PairWorkload) else "amof_b200.bad.Bad.from_trajectory"},
    }
    u = wl.units(res)
    for k, v in u.items():
        out[k] = v
        out[k.replace("_per_step", "_per_s")] = v * args.steps * world / (dev_ms / 1e3)
    peak, how = peaks()
    if isinstance(wl, PairWorkload) and k_n:
        per_launch_ms = k_ms / k_n
        frames_per_launch = wl.T * args.steps / k_n
        ach = wl.algorithmic_bytes_per_frame() * frames_per_launch / (per_launch_ms / 1e3) / 1e9
        pairs = u["pair_evals_per_step"] * args.steps
        fp64 = 10.0 * pairs / (k_ms / 1e3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "k_pair_tiled (+ k_pair_plan)", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                           "traffic": (PAIR_TRAFFIC_PER_FRAME[wl.name] * frames_per_launch if wl.name in PAIR_TRAFFIC_PER_FRAME else None), "peak_source": how, "kernel_ms_per_launch": per_launch_ms, "launches": int(k_n),
                           "algorithmic_bytes_per_launch": wl.algorithmic_bytes_per_frame() * frames_per_launch,
                           "kernel_share_of_step": k_ms / dev_ms,
                           "note": "compute-bound kernel: ~130 in-range pairs per 24 B read, so the HBM fraction is small by design; "
                                   "the binding pipe is FP64 (see fp64)",
                           "fp64": {"achieved_gflops": fp64, "peak_gflops": FP64_NOFMA_GOPS, "frac": fp64 / FP64_NOFMA_GOPS,
                                    "flop_per_pair": 10, "peak_source": "tools/microbench.cu DADD/DMUL without FMA, measured on this pool"}}
    else:
        ach = wl.algorithmic_bytes_per_frame() * wl.T * args.steps / (dev_ms / 1e3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "k_bad + cell list", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "traffic": None, "peak_source": how}
    if world == 1:
        out["cpu_baseline"] = cpu_baseline(wl, args)
    print(json.dumps(out))


def host_threads():
    """threads the CPU arm may use: every core this process is allowed on (torchrun sets OMP_NUM_THREADS=1, which
    would otherwise silently serialise the OpenMP oracle)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(wl, args, frames=None):
    threads = host_threads()
    frames = frames or {"c2": 2048, "c3": 160, "c4": 192}.get(wl.name, 16)      # ~10-15 s of CPU work on 16 cores
    frames = min(frames, wl.T)
    dt = wl.cpu_sample(threads, frames)
    return {"value": frames / dt, "unit": wl.unit, "cores": threads, "kind": "port",
            "sample": "%d frames of the same workload, oracle/amof_oracle.c (ASAP/ASE stand-in), OpenMP over frames, %.1f s" % (frames, dt)}


def run_msd(args, backend, rank, world):
    ctx = backend.ctx
    wl = MsdWorkload("c5", args.frames, args.atoms)
    wl.setup(backend)
    n_of = None
    times, launches = [], 0
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    for it in range(args.warmup + args.steps):
        with backend.msd_open(wl.T, wl.masses, wl.spec, len(wl.zs), wl.cells) as s:
            wl._fill_device(s)                      # generation + transpose-in: not timed (inputs resident)
            barrier_sync(ctx, world)
            if it == args.warmup:
                sampler.start()
            l0 = ctx.launch_count()
            ctx.timer_mark(0)
            raw = wl._analyse(s)
            ctx.timer_mark(1)
            ctx.sync()
            if it >= args.warmup:
                times.append(ctx.timer_elapsed(0, 1))
                launches += ctx.launch_count() - l0
    clocks = sampler.stop()
    dev_ms = max_over_ranks(sum(times), world)
    # e2e: WindowMsd-equivalent through the C ABI from host memory: a pinned slab of frames streamed T/slab times
    t1 = time.perf_counter()
    for _ in range(args.steps):
        with backend.msd_open(wl.T, wl.masses, wl.spec, len(wl.zs), wl.cells) as s:
            for a in range(0, wl.T, wl.slab):
                b = min(wl.T, a + wl.slab)
                s.load(a, wl.host_slab[:b - a])
            raw = wl._analyse(s)
    ctx.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t1, world)
    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    frames_total = wl.T * args.steps * world
    peak, how = peaks()
    ach = wl.bytes_in * args.steps / (dev_ms / 1e3) / 1e9
    out = {"metric": wl.metric, "value": frames_total / (dev_ms / 1e3), "unit": wl.unit, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic (random walk generated on the device)",
           "config": dict(wl.describe(), parallelism="atoms x%d" % world), "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": frames_total / e2e_s, "unit": wl.unit, "h2d_bytes_per_step": int(wl.bytes_in),
                   "d2h_bytes_per_step": int(raw.nbytes + wl.T * 32), "api": "amofb_msd_* via GpuBackend.msd_open"},
           "roofline": {"bound": "hbm", "kernel": "k_msd_frame_sums + k_msd_window_ap (shift, wrap and running sum fused in)", "achieved": ach, "peak": peak,
                        "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": how,
                        "note": "algorithmic bytes = 24*N*T read once; the centre-of-mass pass and the window pass each read it once "
                                "(2x), nothing is written back; the window kernel itself is FP64/issue-bound (7 flop per frame pair)"}}
    u = wl.units(None)
    for k, v in u.items():
        out[k] = v
        out[k.replace("_per_step", "_per_s")] = v * args.steps * world / (dev_ms / 1e3)
    if "atom_frame_pairs_per_step" in u:
        # the tiled window kernel issues 6 FP64 instructions per frame pair (3 differences + 3 chained FMAs)
        ginstr = 6.0 * u["atom_frame_pairs_per_step"] * args.steps / (dev_ms / 1e3) / 1e9
        out["roofline"]["fp64"] = {"achieved_ginstr": ginstr, "peak_ginstr": FP64_NOFMA_GOPS, "frac": ginstr / FP64_NOFMA_GOPS,
                                   "instr_per_pair": 6, "peak_source": "tools/microbench.cu FP64 issue rate, measured on this pool"}
    if world == 1:
        from oracle import c_oracle as orc
        dt = wl.cpu_sample(1, min(wl.T, 400))
        out["cpu_baseline"] = {"value": min(wl.T, 400) / dt, "unit": wl.unit, "cores": 1, "kind": "port",
                               "sample": "oracle msd_window on 20 000 atoms x %d frames, scaled per atom.frame-pair" % min(wl.T, 400)}
    print(json.dumps(out))


def run_reference(args):
    """The reference's CPU path for the same metric/config.  asap3/ase cannot be installed here (no wheels, no
    network: SURVEY.md 8(c)), so this times the oracle port of it -- labelled as such -- with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = {"c2": 1024, "c3": 64, "c4": 96}.get(args.workload, 64)           # one step = a few seconds of CPU work
    if args.workload in ("c2", "c3"):
        wl = PairWorkload(args.workload, sample)
    elif args.workload == "c4":
        wl = BadWorkload(args.workload, sample)
    else:
        print(json.dumps({"impl": "reference", "unavailable": "MSD reference arm not implemented for c5"}))
        return
    from amof_b200 import atom as amatom, frames as fr, synth
    numbers, _, cell = synth.base_frame(args.workload)
    wl.n_atoms = len(numbers)
    wl.host = np.empty((sample, wl.n_atoms, 3))
    synth.fill_frames(args.workload, 0, sample, wl.host)
    wl.traj = fr.ArrayTrajectory(numbers, wl.host, cell)
    wl.zs, wl.spec = fr.species_index(numbers)
    if isinstance(wl, PairWorkload):
        wl.cut = amatom.cutoff_matrix(amatom.format_cutoff(CN_SETS), wl.zs)
    else:
        wl.cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5}), wl.zs)
        wl.triples = [(wl.zs.index(30), wl.zs.index(7)), (wl.zs.index(7), wl.zs.index(30))]
        wl.nbins = int(180 // wl.dtheta) + 1
    wl.bytes_in = wl.host.nbytes
    for _ in range(min(args.warmup, 1)):
        wl.cpu_sample(threads, min(sample, 8))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.cpu_sample(threads, sample)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    full = PairWorkload(args.workload, None) if isinstance(wl, PairWorkload) else BadWorkload(args.workload, None)
    full.n_atoms, full.bytes_in = wl.n_atoms, wl.bytes_in
    out = {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
           "config": dict(full.describe(), parallelism="cpu x%d threads" % threads),
           "cpu_baseline": {"value": value, "unit": wl.unit, "cores": threads, "kind": "port",
                            "sample": "each step = %d frames of the workload; oracle/amof_oracle.c stands in for asap3/ase "
                                      "(not installable here), OpenMP over frames" % sample},
           "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=None, help="frames per step (default: the workload's)")
    ap.add_argument("--atoms", type=int, default=None, help="c5 only: use the first ATOMS atoms")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
