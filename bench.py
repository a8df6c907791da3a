#!/usr/bin/env python
"""
bench.py -- throughput of the aMOF hot path on B200 (BASELINE.json metric: RDF frames/s and pair-evals/s at 1/2/4/8 GPUs,
MSD / BAD frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5] [--frames F]

One step = one pass of the analysis over the whole synthetic trajectory of the workload:
  c2  9 792 atoms x 10 000 frames, all 16 partial RDFs (dr 0.01, rmax 10 -> 999 bins) + CN Zn-N/C-N/C-C
  c3  104 448-atom triclinic box x 2 000 frames, same analysis
  c4  48 960 atoms x 5 000 frames, bond angles N-Zn-N (dtheta 0.05)
  c5  979 200 atoms x 5 000 frames window MSD (117.5 GB: generated on the device, slab by slab)

What one JSON line reports
  N = 1   headline = c2 (the configuration the metric is quoted on); sub-records "c3", "c4", "c5" carry the same fields
          (value, e2e, roofline, cpu_baseline) for the other configurations of BASELINE.json.
  N > 1   headline = c3 STRONG scaling: the fixed 2 000 frames are split over the ranks (frames.frame_range), every rank
          analyses its block, the integer histograms are all-reduced and the per-frame CN rows gathered (NCCL).  Rank 0
          then repeats the analysis of ALL frames alone and checks both legs bit for bit ("parity_checked").  Sub-records
          "c4" (frames sharded) and "c5" (atoms sharded: per-slab mass sums and the final window sums are all-reduced).
  value   frames/s of the whole job with the trajectory already resident in HBM (amofb_*_push_device / slab_*_device),
  e2e     the same through the public classes (amof_b200.rdf.rdf_and_cn / bad.Bad / msd.WindowMsd) from page-locked HOST
          arrays: host->device copies and the result read-back are inside the timed region.
Timing: CUDA events on the library's compute stream around the K timed steps, max over ranks; >= 3 warm-up steps; the inputs
of a step are larger than L2.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CN_SETS = {'Zn-N': 2.5, 'C-N': 1.728, 'C-C': 1.752}
FP64_NOFMA_GOPS = 18515.3      # measured on this pool's B200 with tools/microbench.cu (profiles/r01_microbench_fp64_atomics.json)
PAIR_NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_k_pair_tiled_ncu_full_summary.txt")     # one launch over 107 C2 frames
PAIR_NCU_FRAMES = 107


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


def dist_setup():
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


def pair_traffic_per_frame():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_pair_tiled launch (ncu --set full), per frame; None when the
    committed summary is missing"""
    try:
        txt = open(PAIR_NCU_SUMMARY).read()
    except OSError:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(r"^%s\s+([0-9.]+)\s+(\w+)" % re.escape(key), txt, re.M)
        if not m:
            return None, None
        tot += float(m.group(1)) * unit.get(m.group(2), 1.0)
    return tot / PAIR_NCU_FRAMES, os.path.relpath(PAIR_NCU_SUMMARY, ROOT)


def host_threads():
    """threads the CPU arm may use: every core this process is allowed on (torchrun sets OMP_NUM_THREADS=1, which
    would otherwise silently serialise the OpenMP oracle)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def shard(T, rank, world):
    return (T * rank) // world, (T * (rank + 1)) // world


# ------------------------------------------------------------------------------------------------ workloads
class PairWorkload:
    """c2 / c3: partial RDFs + coordination numbers.  world > 1: the frames are split over the ranks (strong scaling)."""
    metric, unit, dtype = "rdf_cn_frames_per_s", "frames/s", "f64"

    def __init__(self, name, frames=None, rank=0, world=1):
        from amof_b200 import synth
        self.name, self.rank, self.world = name, rank, world
        self.T = frames or synth.CONFIGS[name]["frames"]
        self.rmax, self.dr = 10.0, 0.01
        self.bins = int(self.rmax // self.dr)
        self.lo, self.hi = shard(self.T, rank, world)
        numbers, _, cell = synth.base_frame(name)
        self.numbers, self.cell, self.n_atoms = numbers, cell, len(numbers)
        self.bytes_in = 24 * self.n_atoms * self.T + 72 * self.T          # whole job

    def describe(self):
        return {"workload": "%s: synthetic a-ZIF %d atoms x %d frames, 16 partial RDFs (dr 0.01, rmax 10 -> %d bins) + CN %s"
                            % (self.name, self.n_atoms, self.T, self.bins, "/".join(CN_SETS)),
                "atoms": self.n_atoms, "frames_per_step": self.T,
                "l2": "inputs (%.2f GB/step over %d GPU%s) larger than L2" % (self.bytes_in / 1e9, self.world, "s" if self.world > 1 else ""),
                "parallelism": "frames x%d" % self.world}

    def setup(self, backend):
        from amof_b200 import atom as amatom, frames as fr, synth
        ctx = backend.ctx
        self.backend = backend
        n_loc = self.hi - self.lo
        self.host = ctx.pinned_empty((n_loc, self.n_atoms, 3))
        synth.fill_frames(self.name, self.lo, n_loc, self.host)
        self.traj = fr.ArrayTrajectory(self.numbers, self.host, self.cell, pinned=True, first_frame=self.lo, n_frames=self.T)
        self.zs, self.spec = fr.species_index(self.numbers)
        self.cut = amatom.cutoff_matrix(amatom.format_cutoff(CN_SETS), self.zs)
        self.dev = ctx.device_alloc(max(self.host.nbytes, 8))
        ctx.h2d(self.dev, self.host)
        self.chunk = max(1, (256 << 20) // (24 * self.n_atoms))

    def teardown(self):
        ctx = self.backend.ctx
        ctx.device_free(self.dev)
        ctx.pinned_free(self.host)
        self.host = self.traj = None

    def _device_chunks(self):
        fb = 24 * self.n_atoms
        for a in range(self.lo, self.hi, self.chunk):
            b = min(self.hi, a + self.chunk)
            yield self.dev.value + (a - self.lo) * fb, self.traj.cells[a:b]

    def step_resident(self):
        from amof_b200 import _dist
        res = self.backend.pair_counts(self.spec, len(self.zs), self._device_chunks(), rmax=self.rmax, nbins=self.bins,
                                       cn_cutoff=self.cut)
        if self.world > 1:      # the collectives of the path: integer all-reduce of the histograms, gather of the CN rows (NCCL)
            res["hist"] = _dist.allreduce_sum(res["hist"])
            counts = [shard(self.T, r, self.world)[1] - shard(self.T, r, self.world)[0] for r in range(self.world)]
            res["cn"] = _dist.allgather_rows(res["cn"], counts)
        self.last = res
        return res

    def step_e2e(self):
        from amof_b200 import rdf
        r, c = rdf.rdf_and_cn(self.traj, CN_SETS, dr=self.dr, rmax=self.rmax, distributed=self.world > 1)
        self.last_e2e = (r, c)
        return r.counts.nbytes + c.counts.nbytes
    e2e_api = "amof_b200.rdf.rdf_and_cn"

    def units(self):
        return {"pair_evals_per_step": int(self.last["hist"].sum()) // 2}

    def parity(self):
        """rank 0 alone over ALL frames (generated chunk by chunk), against the sharded results of both legs"""
        from amof_b200 import synth
        ctx = self.backend.ctx
        step = max(1, min(self.T, (128 << 20) // (24 * self.n_atoms)))
        buf = ctx.pinned_empty((step, self.n_atoms, 3))
        cells = np.broadcast_to(self.cell, (self.T, 3, 3))

        def chunks():
            for a in range(0, self.T, step):
                b = min(self.T, a + step)
                ctx.sync_copies()
                synth.fill_frames(self.name, a, b - a, buf)
                yield buf[:b - a], cells[a:b]
        one = self.backend.pair_counts(self.spec, len(self.zs), chunks(), rmax=self.rmax, nbins=self.bins, cn_cutoff=self.cut)
        ctx.pinned_free(buf)
        r, c = self.last_e2e
        ok = (np.array_equal(one["hist"], self.last["hist"]) and np.array_equal(one["cn"], self.last["cn"]) and
              np.array_equal(one["hist"], r.counts) and np.array_equal(one["cn"], c.counts) and one["n_frames"] == self.T)
        if not ok:
            raise SystemExit("bench.py: the %d-rank result differs from the single-rank result" % self.world)
        return True

    def roofline(self, dev_ms, steps, k_ms, k_n):
        peak, how = peaks()
        pairs = self.units()["pair_evals_per_step"] * steps
        per_launch_ms = k_ms / k_n
        frames_per_launch = (self.hi - self.lo) * steps / k_n
        alg = (24 * self.n_atoms + 72) * frames_per_launch
        hbm = alg / (per_launch_ms / 1e3) / 1e9
        tf = 10.0 * (pairs / self.world) / (k_ms / 1e3) / 1e12          # this rank's share of the pairs over its kernel time
        per_frame, src = pair_traffic_per_frame()
        return {"bound": "fp64", "kernel": "k_pair_tiled (+ k_pair_plan)", "achieved": tf, "peak": FP64_NOFMA_GOPS / 1e3, "unit": "TFLOP/s",
                "frac": tf / (FP64_NOFMA_GOPS / 1e3), "flop_per_pair": 10,
                "peak_source": "tools/microbench.cu DADD/DMUL without FMA, measured on this pool (profiles/r01_microbench_fp64_atomics.json)",
                "traffic": (per_frame * frames_per_launch * (self.n_atoms / 9792.0)) if per_frame else None,
                "traffic_source": (src + " (dram read+write of one launch over %d C2 frames, scaled per atom.frame)" % PAIR_NCU_FRAMES) if src else None,
                "kernel_ms_per_launch": per_launch_ms, "launches": int(k_n), "kernel_share_of_step": k_ms / dev_ms,
                "algorithmic_bytes_per_launch": alg,
                "hbm": {"achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "peak_source": how,
                        "note": "~130 in-range pairs per 24 B read: the HBM fraction is small by design"}}

    def cpu_sample(self, threads, frames):
        from amof_b200 import synth
        from oracle import c_oracle as orc
        pos = np.empty((frames, self.n_atoms, 3))
        synth.fill_frames(self.name, 0, frames, pos)
        cells = np.broadcast_to(self.cell, (frames, 3, 3)).copy()
        if not hasattr(self, "spec"):
            from amof_b200 import atom as amatom, frames as fr
            self.zs, self.spec = fr.species_index(self.numbers)
            self.cut = amatom.cutoff_matrix(amatom.format_cutoff(CN_SETS), self.zs)
        t0 = time.perf_counter()
        orc.rdf_traj(pos, cells, self.spec, len(self.zs), self.rmax, self.bins, threads=threads)
        orc.cn_traj(pos, cells, self.spec, len(self.zs), self.cut, threads=threads)
        return time.perf_counter() - t0
    cpu_frames = {"c2": 2048, "c3": 160}
    ref_frames = {"c2": 1024, "c3": 64}


class BadWorkload:
    """c4: N-Zn-N bond-angle distribution; world > 1: frames split over the ranks."""
    metric, unit, dtype = "bad_frames_per_s", "frames/s", "f64"

    def __init__(self, name, frames=None, rank=0, world=1):
        from amof_b200 import synth
        self.name, self.rank, self.world = name, rank, world
        self.T = frames or synth.CONFIGS[name]["frames"]
        self.dtheta = 0.05
        self.lo, self.hi = shard(self.T, rank, world)
        numbers, _, cell = synth.base_frame(name)
        self.numbers, self.cell, self.n_atoms = numbers, cell, len(numbers)
        self.bytes_in = 24 * self.n_atoms * self.T + 72 * self.T

    def describe(self):
        return {"workload": "%s: synthetic a-ZIF %d atoms x %d frames, Bad({'Zn-N': 2.5}), dtheta 0.05 -> 3600 bins"
                            % (self.name, self.n_atoms, self.T), "atoms": self.n_atoms, "frames_per_step": self.T,
                "l2": "inputs (%.2f GB/step) larger than L2" % (self.bytes_in / 1e9), "parallelism": "frames x%d" % self.world}

    def setup(self, backend):
        from amof_b200 import atom as amatom, frames as fr, synth
        ctx = backend.ctx
        self.backend = backend
        n_loc = self.hi - self.lo
        self.host = ctx.pinned_empty((n_loc, self.n_atoms, 3))
        synth.fill_frames(self.name, self.lo, n_loc, self.host)
        self.traj = fr.ArrayTrajectory(self.numbers, self.host, self.cell, pinned=True, first_frame=self.lo, n_frames=self.T)
        self.zs, self.spec = fr.species_index(self.numbers)
        self.cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5}), self.zs)
        self.triples = [(self.zs.index(30), self.zs.index(7)), (self.zs.index(7), self.zs.index(30))]
        self.nbins = int(180 // self.dtheta) + 1
        # the library gathers the atoms that can take part (Zn, N) on the host and copies only those (+ the cells)
        kept = int(np.count_nonzero(np.isin(self.numbers, (30, 7))))
        self.bytes_h2d = (24 * kept + 72) * self.T if 10 * kept < 6 * self.n_atoms and not os.environ.get("AMOFB_NO_HOST_GATHER") else self.bytes_in
        self.dev = ctx.device_alloc(max(self.host.nbytes, 8))
        ctx.h2d(self.dev, self.host)
        self.chunk = max(1, (256 << 20) // (24 * self.n_atoms))

    teardown = PairWorkload.teardown
    _device_chunks = PairWorkload._device_chunks

    def step_resident(self):
        from amof_b200 import _dist
        hist, dropped, nf = self.backend.bad_counts(self.spec, len(self.zs), self._device_chunks(), self.cut, self.triples,
                                                    self.dtheta, self.nbins)
        if self.world > 1:
            hist = _dist.allreduce_sum(hist)
        self.last = {"hist": hist}
        return self.last

    def step_e2e(self):
        from amof_b200 import bad
        self.last_e2e = bad.Bad.from_trajectory(self.traj, {'Zn-N': 2.5}, dtheta=self.dtheta, distributed=self.world > 1)
        return 2 * 33 * self.nbins * 8
    e2e_api = "amof_b200.bad.Bad.from_trajectory"

    def units(self):
        return {"angles_per_step": int(self.last["hist"].sum())}

    def parity(self):
        got = self.last["hist"][0].sum(axis=0)
        if not np.array_equal(got, self.last_e2e.counts["N-Zn-N"]):
            raise SystemExit("bench.py: the resident and the end-to-end angle histograms differ")
        return True

    def roofline(self, dev_ms, steps, k_ms, k_n):
        peak, how = peaks()
        ach = (24 * self.n_atoms + 72) * self.T * steps / (dev_ms / 1e3) / 1e9
        return {"bound": "hbm", "kernel": "k_bad_search + k_bad_angles + cell list", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": how,
                "note": "latency-bound neighbour walk over a sparse, species-filtered frame (profiles/r02_k_bad_search_ncu_full_summary.txt)"}

    def cpu_sample(self, threads, frames):
        from amof_b200 import synth
        from oracle import c_oracle as orc
        if not hasattr(self, "spec"):
            from amof_b200 import atom as amatom, frames as fr
            self.zs, self.spec = fr.species_index(self.numbers)
            self.cut = amatom.cutoff_matrix(amatom.format_cutoff({'Zn-N': 2.5}), self.zs)
            self.triples = [(self.zs.index(30), self.zs.index(7)), (self.zs.index(7), self.zs.index(30))]
            self.nbins = int(180 // self.dtheta) + 1
        pos = np.empty((frames, self.n_atoms, 3))
        synth.fill_frames(self.name, 0, frames, pos)
        t0 = time.perf_counter()
        for f in range(frames):
            for (A, B) in self.triples:
                orc.bad_hist(pos[f], self.cell, self.spec, len(self.zs), self.cut, A, B, self.dtheta, self.nbins)
        return time.perf_counter() - t0
    cpu_frames = {"c4": 96}
    ref_frames = {"c4": 48}


class MsdWorkload:
    """c5: window MSD of an unwrapped random walk.  The trajectory (117.5 GB at full size) is generated on the device slab by
    slab; a step = ingest of every slab (mass sums, centre-of-mass shift + wrap + running sum + transposition) + the window
    kernel.  world > 1: the ATOMS are split over the ranks."""
    metric, unit, dtype = "msd_frames_per_s", "frames/s", "f64"

    def __init__(self, name, frames=None, rank=0, world=1, atoms=None):
        from amof_b200 import synth
        self.name, self.rank, self.world = name, rank, world
        self.T = frames or synth.CONFIGS[name]["frames"]
        numbers, pos0, cell = synth.base_frame(name)
        if atoms:
            numbers, pos0 = numbers[:atoms], pos0[:atoms]
        self.n_total = len(numbers)
        self.lo, self.hi = shard(self.n_total, rank, world)
        self.numbers, self.pos0, self.cell = numbers, pos0, cell
        self.n_atoms = self.n_total
        self.window = np.arange(0, self.T // 2, 100)
        self.bytes_in = 24 * self.n_total * self.T

    def describe(self):
        return {"workload": "%s: synthetic %d atoms x %d frames unwrapped random walk, WindowMsd(delta_time=100) -> %d windows"
                            % (self.name, self.n_total, self.T, len(self.window)), "atoms": self.n_total, "frames_per_step": self.T,
                "l2": "inputs (%.2f GB/step) larger than L2" % (self.bytes_in / 1e9), "parallelism": "atoms x%d" % self.world}

    def setup(self, backend):
        import torch
        from amof_b200 import frames as fr
        from amof_b200.elements import atomic_masses
        self.backend = backend
        self.zs, spec = fr.species_index(self.numbers)
        self.spec_all = spec
        self.spec = spec[self.lo:self.hi]
        self.masses_all = np.array([atomic_masses[z] for z in self.numbers])
        self.masses = self.masses_all[self.lo:self.hi]
        self.cells = np.broadcast_to(self.cell, (self.T, 3, 3)).copy()
        n = self.hi - self.lo
        slab = max(1, min(self.T, 256, (3 << 29) // (24 * n)))
        self.slab = slab - slab % 32 if slab >= 32 else slab
        self.torch, self.dev = torch, torch.device("cuda", backend.ctx.device)
        self.base = torch.from_numpy(self.pos0[self.lo:self.hi]).to(self.dev)
        self.t_last = 0.0

    def teardown(self):
        self.base = None
        self.torch.cuda.empty_cache()

    def step_resident(self):
        """returns the device time of the step's library calls (the generation of the synthetic slabs is not part of it)"""
        from amof_b200 import _dist
        torch, ctx, n = self.torch, self.backend.ctx, self.hi - self.lo
        g = torch.Generator(device=self.dev); g.manual_seed(20261023 + self.rank)
        cur = self.base.clone()
        t_dev = 0.0
        begun = []
        with self.backend.msd_open(self.T, self.masses, self.spec, len(self.zs), self.cells) as s:
            def finish():
                a_, b_, _keep = begun.pop(0)
                sums = _dist.allreduce_sum(s.slab_sums_wait(b_ - a_))
                s.slab_commit(sums[:, :3] / sums[:, 3:4])
            for a in range(0, self.T, self.slab):
                b = min(self.T, a + self.slab)
                inc = 0.05 * torch.randn((b - a, n, 3), generator=g, device=self.dev, dtype=torch.float64)
                inc[0] += cur
                blk = inc.cumsum(0)
                cur = blk[-1].clone()
                del inc
                torch.cuda.synchronize()
                ctx.timer_mark(0)
                s.slab_sums_begin(a, (blk.data_ptr(), b - a))
                begun.append((a, b, blk))
                if len(begun) == 2:
                    finish()
                ctx.timer_mark(1)
                ctx.sync()
                t_dev += ctx.timer_elapsed(0, 1)
            ctx.timer_mark(0)
            while begun:
                finish()
            raw = _dist.allreduce_sum(s.window(self.window.astype(np.int32)))
            ctx.timer_mark(1)
            ctx.sync()
            t_dev += ctx.timer_elapsed(0, 1)
        self.last = {"raw": raw}
        self.t_last = t_dev
        return self.last

    def step_e2e(self):
        """WindowMsd.from_trajectory on a host trajectory whose frames are served from one page-locked slab (117.5 GB do not fit
        in host memory): every frame crosses PCIe, the analysis is the public class's"""
        from amof_b200 import frames as fr, msd
        if not hasattr(self, "traj"):
            ctx = self.backend.ctx
            n_local = -(-self.n_total // self.world)         # the library sizes its slabs by the atoms a rank holds
            nslab = max(1, min(self.T, 256, (3 << 29) // (24 * n_local)))
            nslab = min(nslab, max(32, (2 << 30) // (24 * self.n_total)))   # at most 2 GiB of page-locked host memory per rank
            nslab = nslab - nslab % 32 if nslab >= 32 else nslab          # whole rounds of the commit kernel; blocks are served as views
            rng = np.random.default_rng(5)
            slabbuf = ctx.pinned_empty((nslab, self.n_total, 3))
            slabbuf[...] = self.pos0[None] + np.cumsum(rng.normal(scale=0.05, size=(nslab, self.n_total, 3)), axis=0)

            class Cyclic(fr.ArrayTrajectory):
                def resident(self_, a, b):
                    return True

                def block(self_, a, b):
                    return slabbuf[:b - a] if b - a <= nslab else np.concatenate([slabbuf] * ((b - a) // nslab + 1))[:b - a]
            self.traj = Cyclic(self.numbers, slabbuf, self.cell, masses=self.masses_all, pinned=True, first_frame=0, n_frames=self.T)
            self.traj.block_frames = nslab                  # WindowMsd then asks for slabs of this many frames
            self._slabbuf = slabbuf
        m = msd.WindowMsd.from_trajectory(self.traj, delta_time=100, timestep=1, mutate=False, distributed=self.world > 1)
        self.last_e2e = m
        return int(m.data.to_numpy().nbytes + self.T * 32)
    e2e_api = "amof_b200.msd.WindowMsd.from_trajectory"

    def units(self):
        return {"atom_frame_pairs_per_step": int(self.n_total * sum(self.T - m - 1 for m in self.window))}

    def parity(self):
        # random-walk law (quirk Q4): MSD(m) = 3 sigma^2 m (T-m-1)/(T-m), sampling noise ~ 1/sqrt(atoms)
        raw = self.last["raw"]
        n_of = np.bincount(self.spec_all, minlength=len(self.zs)).astype(np.float64)
        m = self.window.astype(np.float64)
        got = (raw / n_of[:, None] / (self.T - m)[None, :]).mean(axis=0)
        law = 3 * 0.05 ** 2 * m * (self.T - m - 1) / (self.T - m)
        if not (got[0] == 0.0 and np.allclose(got[1:], law[1:], rtol=0.05)):
            raise SystemExit("bench.py: MSD does not follow the random-walk law")
        return True

    def roofline(self, dev_ms, steps, k_ms, k_n):
        peak, how = peaks()
        ach = self.bytes_in / self.world * steps / (dev_ms / 1e3) / 1e9
        pairs = self.units()["atom_frame_pairs_per_step"] / self.world
        return {"bound": "hbm", "kernel": "k_msd_slab_sums + k_msd_slab_commit + k_msd_window_wide", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": 4.0 * self.bytes_in / self.world, "peak_source": how,
                "note": "algorithmic bytes = 24*N*T read once; the path moves 4x that by construction (mass sums read, commit read + "
                        "write of the atom-major store, window read), so 0.25 would be the ceiling of this fraction",
                "fp64": {"achieved_ginstr": 3.0 * pairs * steps / (dev_ms / 1e3) / 1e9, "peak_ginstr": FP64_NOFMA_GOPS, "instr_per_pair": 3}}

    def cpu_sample(self, threads, frames):
        from amof_b200 import frames as fr
        from amof_b200.elements import atomic_masses
        from oracle import c_oracle as orc
        n = min(self.n_total, 20000)
        T = min(frames, 400)
        zs, spec = fr.species_index(self.numbers)
        masses = np.array([atomic_masses[z] for z in self.numbers])
        rng = np.random.default_rng(5)
        pos = self.pos0[None, :n] + np.cumsum(rng.normal(scale=0.05, size=(T, n, 3)), axis=0)
        w = np.arange(0, T // 2, max(1, T // 50))
        t0 = time.perf_counter()
        orc.msd_window(pos, np.broadcast_to(self.cell, (T, 3, 3)).copy(), masses[:n], spec[:n], len(zs), w)
        dt = time.perf_counter() - t0
        done = n * sum(T - m - 1 for m in w)
        full = self.n_total * sum(self.T - m - 1 for m in self.window)
        self.cpu_note = "oracle msd_window on %d atoms x %d frames (1 thread), scaled per atom.frame-pair" % (n, T)
        return dt * full / max(done, 1) * (frames / self.T)
    cpu_frames = {"c5": 400}
    ref_frames = {"c5": 400}


def make_workload(name, frames, rank, world, atoms=None):
    if name in ("c2", "c3"):
        return PairWorkload(name, frames, rank, world)
    if name == "c4":
        return BadWorkload(name, frames, rank, world)
    return MsdWorkload(name, frames, rank, world, atoms)


# ------------------------------------------------------------------------------------------------ measurement
T_START = time.perf_counter()


def note(rank, msg):
    if rank == 0:
        print("[bench %6.1f s] %s" % (time.perf_counter() - T_START, msg), file=sys.stderr, flush=True)


def measure(wl, backend, steps, warmup, rank, world, local, e2e_steps=None, with_cpu=True, check=False):
    """-> the record of one workload (rank 0), None on the other ranks"""
    ctx = backend.ctx
    note(rank, "%s: setup" % wl.name)
    wl.setup(backend)
    note(rank, "%s: resident leg, %d + %d steps" % (wl.name, warmup, steps))
    is_msd = isinstance(wl, MsdWorkload)
    for _ in range(warmup):
        wl.step_resident()
    ctx.set_profiling(True)
    ctx.pair_kernel_time(reset=True)
    sampler = ClockSampler(local)
    barrier_sync(ctx, world)
    sampler.start()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    if not is_msd:
        ctx.timer_mark(6)
    for _ in range(steps):
        wl.step_resident()
        if is_msd:
            dev_ms += wl.t_last
    if not is_msd:
        ctx.timer_mark(7)
    barrier_sync(ctx, world)
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    if not is_msd:
        dev_ms = ctx.timer_elapsed(6, 7)
    dev_ms = max_over_ranks(dev_ms, world)
    launches = ctx.launch_count() - l0
    k_ms, k_n = ctx.pair_kernel_time(reset=True)
    ctx.set_profiling(False)

    # ---- end-to-end leg: public classes, page-locked host input, H2D and result D2H inside the timed region ----
    e2e_steps = e2e_steps or steps
    note(rank, "%s: end-to-end leg" % wl.name)
    for _ in range(1 if is_msd else min(warmup, 2)):
        wl.step_e2e()
    barrier_sync(ctx, world)
    t1 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        d2h = wl.step_e2e()
    barrier_sync(ctx, world)
    e2e_s = max_over_ranks(time.perf_counter() - t1, world)
    parity = None
    if check:
        note(rank, "%s: parity check" % wl.name)
        parity = wl.parity() if rank == 0 else None
        barrier_sync(ctx, world)
    rec = None
    if rank == 0:
        value = wl.T * steps / (dev_ms / 1e3)
        rec = {"metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
               "dtype": wl.dtype, "data": "synthetic" + (" (random walk generated on the device)" if is_msd else ""),
               "config": wl.describe(), "clocks": clocks, "gpu_launches": int(launches), "wall_ms_per_step": wall / steps * 1e3,
               "e2e": {"value": wl.T * e2e_steps / e2e_s, "unit": wl.unit, "steps": e2e_steps, "h2d_bytes_per_step": int(getattr(wl, "bytes_h2d", wl.bytes_in)),
                       "d2h_bytes_per_step": int(d2h), "api": wl.e2e_api}}
        for k, v in wl.units().items():
            rec[k] = v
            rec[k.replace("_per_step", "_per_s")] = v * steps / (dev_ms / 1e3)
        rec["roofline"] = wl.roofline(dev_ms, steps, k_ms, max(k_n, 1))
        if parity is not None:
            rec["parity_checked"] = bool(parity)
        if with_cpu and world == 1:
            note(rank, "%s: cpu baseline" % wl.name)
            rec["cpu_baseline"] = cpu_baseline(wl)
    wl.teardown()
    return rec


def cpu_baseline(wl):
    threads = 1 if isinstance(wl, MsdWorkload) else host_threads()
    frames = min(wl.cpu_frames[wl.name], wl.T)            # ~10-15 s of CPU work
    dt = wl.cpu_sample(threads, frames)
    note = getattr(wl, "cpu_note", "%d frames of the same workload, oracle/amof_oracle.c (ASAP/ASE stand-in), OpenMP over frames" % frames)
    return {"value": frames / dt, "unit": wl.unit, "cores": threads, "kind": "port", "sample": "%s, %.1f s" % (note, dt)}


def sub_record(rec):
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "scaling", "config", "e2e", "roofline", "cpu_baseline",
            "gpu_launches", "clocks", "parity_checked", "pair_evals_per_s", "angles_per_s", "atom_frame_pairs_per_s")
    return {k: rec[k] for k in keep if k in rec}


def run_ours(args):
    rank, world, local = dist_setup()
    if local and "AMOFB_DEVICE" not in os.environ:
        os.environ["AMOFB_DEVICE"] = str(local)
    from amof_b200 import _lib
    backend = _lib.get_backend()
    if args.workload:            # one workload alone (tools/round_records.sh, profiling)
        wl = make_workload(args.workload, args.frames, rank, world, args.atoms)
        rec = measure(wl, backend, args.steps, args.warmup, rank, world, local, e2e_steps=1 if args.workload == "c5" else None,
                      check=world > 1)
    else:
        head = "c2" if world == 1 else "c3"
        rec = measure(make_workload(head, None, rank, world), backend, args.steps, args.warmup, rank, world, local, check=world > 1)
        s_steps, s_warm = min(args.steps, 5), 3
        for name in [n for n in ("c3", "c4", "c5") if n != head]:
            sub = measure(make_workload(name, None, rank, world), backend, 2 if name == "c5" else s_steps, s_warm, rank, world, local,
                          e2e_steps=1 if name == "c5" else min(s_steps, 3), check=(world > 1 or name != "c3"))
            if rank == 0:
                rec[name] = sub_record(sub)
    if world > 1:
        import torch.distributed as td
        td.barrier()
        td.destroy_process_group()
    if rank == 0:
        print(json.dumps(rec))


# ------------------------------------------------------------------------------------------------ reference arm
def probe_reference():
    """The unmodified reference (amof + ase + asap3) when it can be imported: baseline/_ref (the driver's install) or the
    interpreter's own site-packages.  -> module `amof` or None"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import ase  # noqa: F401
        import asap3  # noqa: F401
        import amof.rdf  # noqa: F401
        import amof.cn  # noqa: F401
        import amof
        return amof
    except Exception:
        return None


def run_reference(args):
    """The reference's CPU path for the same metric and config as our arm at this N.  With ase + asap3 + amof importable
    the unmodified classes are timed (kind "reference"); otherwise -- this image has neither package and no index -- the
    oracle port of them, labelled as such, with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    name = args.workload or ("c2" if world == 1 else "c3")
    wl = make_workload(name, None, 0, max(world, 1), args.atoms)
    sample = wl.ref_frames[name]
    amof = probe_reference() if name in ("c2", "c3") else None
    threads = host_threads()
    kind = "port"
    if amof is not None:
        import ase
        from amof_b200 import synth
        pos = np.empty((sample, wl.n_atoms, 3))
        synth.fill_frames(name, 0, sample, pos)
        traj = [ase.Atoms(numbers=wl.numbers, positions=pos[k], cell=wl.cell, pbc=True) for k in range(sample)]

        def step():
            t0 = time.perf_counter()
            amof.rdf.Rdf.from_trajectory(traj, dr=wl.dr, rmax=wl.rmax)
            amof.cn.CoordinationNumber.from_trajectory(traj, CN_SETS)
            return time.perf_counter() - t0
        kind, threads = "reference", 1          # amof.rdf is a serial loop over frames (amof/rdf.py:88-93)
        note = "each step = %d frames through the unmodified amof.rdf.Rdf + amof.cn.CoordinationNumber (asap3 / ase)" % sample
    else:
        def step():
            return wl.cpu_sample(1 if isinstance(wl, MsdWorkload) else threads, sample)
        if isinstance(wl, MsdWorkload):
            threads = 1
        note = ("each step = %d frames of the workload; oracle/amof_oracle.c stands in for asap3/ase (not importable here), "
                "OpenMP over frames" % sample)
    for _ in range(min(args.warmup, 1)):
        step()
    dt = sum(step() for _ in range(args.steps))
    value = sample * args.steps / dt
    out = {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
           "config": wl.describe(),
           "cpu_baseline": {"value": value, "unit": wl.unit, "cores": threads, "kind": kind, "sample": getattr(wl, "cpu_note", note)},
           "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c2", "c3", "c4", "c5"], help="measure this workload alone")
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
