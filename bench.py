#!/usr/bin/env python
"""bench.py -- site-updates/s of the Metropolis sweep (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE config C5): triangular 128 x 128, n = 1, SU(2)xSU(2) notebook couplings applied as a dense 16x16 matrix,
8192 replicas = 64 temperatures log-spaced in [0.02, 2] x 128 copies, FP32 state + FP64 accumulators, replicas split evenly over the ranks
(total work fixed -> "strong" scaling).  One step = one measurement interval of the reference's run! = `--sweeps` (10, measurement_rate,
src/MonteCarlo.jl:11) colour-ordered sweeps of every replica with 2 updates per site visit (= the reference's 2N updates per sweep,
src/MonteCarlo.jl:40-44); the e2e arm adds the measure! call that ends the interval (src/MonteCarlo.jl:82-86).

  value  : whole-job site-updates/s with everything resident in HBM, CUDA-event timed, max over ranks
  e2e    : the same through the public C-ABI calls with HOST buffers (scmc_sweep: beta/sigma H2D, counters D2H; scmc_measure: rows D2H)
  roofline / cpu_baseline : see DESIGN.md "Measurement"
`--impl reference` times the CPU restatement of the reference path (oracle/, all host cores, one independent chain per core).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NOTEBOOK_J = np.diag([0, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2, -1, 0.2, 0.2, 0.2]).astype(float)
SEED = 20230301
METRIC = "site-updates/sec (n=1, 128x128 triangular Metropolis sweep)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=128)
    ap.add_argument("--replicas", type=int, default=8192, help="total replicas over all ranks")
    ap.add_argument("--sweeps", type=int, default=10, help="sweeps per step (10 = the reference's measurement interval, MonteCarlo.jl:11)")
    ap.add_argument("--hits", type=int, default=2)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--coupling", default="notebook", choices=["notebook", "dense"])
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1/2 streaming/resident T-gather, 3/4 streaming/resident psi-gather, 5/6 TMA-pipelined psi-gather without / with the tensor-core mat-vec")
    ap.add_argument("--stream", type=int, default=2, choices=[1, 2], help="random stream version (DESIGN.md section 5): 1 = Philox4x32-10, 2 = Philox4x32-7")
    ap.add_argument("--batch", type=int, default=0, help="replicas per L2 batch (streaming path)")
    ap.add_argument("--colouring", default="auto", choices=["auto", "few", "balanced"], help="few = 3 large + seam classes, balanced = 2x2 parity (4 equal classes)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=1.2, help="a step holds as many measurement intervals as it takes for the timed region to last this long")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-random-J figure in `extra`")
    ap.add_argument("--extra-hits", type=int, default=8, help="also report the throughput at this many updates per site visit (0 = skip)")
    return ap.parse_args()


def couplings(kind):
    if kind == "notebook":
        return NOTEBOOK_J
    rng = np.random.default_rng(7)
    A = rng.uniform(-1, 1, (16, 16))
    J = 0.5 * (A + A.T)
    J[0, 0] = 0.0
    return J


def temperatures(R):
    """64 temperatures log-spaced in [0.02, 2], each repeated R/64 times (at least once)."""
    nT = min(64, R)
    Ts = np.geomspace(0.02, 2.0, nT)
    return np.repeat(Ts, -(-R // nT))[:R]


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1, t_load0=None):
        """Samples inside the timed region [t0, t1]; a region shorter than the sampling period (strong scaling at 8 GPUs) falls back to
        the samples taken under the same load since t_load0 (the warm-up steps run the identical kernels back to back)."""
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        out = self._summarise(t0, t1 + 0.05)
        if out is None and t_load0 is not None:
            out = self._summarise(t_load0, t1 + 0.2)
            if out is not None:
                out["window"] = "warm-up + timed region (timed region shorter than the sampling period)"
        return out

    def _summarise(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1):
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle)
_ORACLE_BUILT = False


def cpu_reference_rate(L, J, target_seconds, sweeps=None, max_cores=None):
    """The reference's CPU path (C restatement, -O3 -march=native), one independent chain per host core (the reference's
    job-farm model, src/Launcher.jl:286-289) on a single C5 replica each.  Returns (updates/s, cores, sample description)."""
    from oracle import oracle as orc, lattice as ol
    global _ORACLE_BUILT
    orc.build(fast=True, force=not _ORACLE_BUILT)   # rebuilt once on this host: -march=native must match the box's CPU
    _ORACLE_BUILT = True
    pos, nbr, lab, bl = ol.triangular(L)
    try:
        cores = len(os.sched_getaffinity(0))      # the cores this process may actually run on (cpu_count ignores affinity / cgroup masks)
    except AttributeError:
        cores = os.cpu_count() or 1
    if max_cores:
        cores = min(cores, max_cores)
    chains = []
    for c in range(cores):
        o = orc.Oracle(nbr, lab, J, np.zeros(16), 1, fast=True)
        o.seed(SEED + c); o.randomize()
        chains.append(o)
    t = time.perf_counter(); chains[0].sweeps_random(1, 1.0, 0.3); per_sweep = time.perf_counter() - t
    if sweeps is None:
        sweeps = max(1, int(target_seconds / max(per_sweep, 1e-6)))
    betas = 1.0 / temperatures(cores)

    def work(k):
        chains[k].sweeps_random(sweeps, float(betas[k % len(betas)]), 0.3)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(cores)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    updates = cores * sweeps * 2 * L * L
    return updates / dt, cores, f"{cores} independent chains x {sweeps} sweeps (2N random-site updates each) of one {L}x{L} triangular replica, FP64, {dt:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    J = couplings(args.coupling)
    rates = []
    per_step = max(0.5, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    t0 = time.perf_counter()
    for s in range(args.warmup + args.steps):
        rate, cores, sample = cpu_reference_rate(args.L, J, per_step)
        if s >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "site-updates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t0) / max(1, args.steps + args.warmup), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5: triangular {args.L}x{args.L}, n=1, {args.coupling} couplings; CPU restatement of the reference path "
                                   "(oracle/scmc_oracle.c, Julia unavailable)", "sample_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "site-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def kernel_source_hash():
    """Hash of the sources the headline kernel is compiled from: an ncu traffic figure is only reported while it belongs to this code."""
    import hashlib
    h = hashlib.sha256()
    for f in ("scmc_tma.cu", "device_core.cuh", "gen_code.h", "handle.h"):
        with open(os.path.join(ROOT, "semiclassicalmc.jl_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import scmc_b200 as scmc
    from scmc_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R_total = args.replicas
    R = R_total // world + (1 if rank < R_total % world else 0)
    roff = rank * (R_total // world) + min(rank, R_total % world)
    J = couplings(args.coupling)
    Ts_all = temperatures(R_total)
    Ts = Ts_all[roff:roff + R]
    colour = None
    if args.colouring != "auto":
        from scmc_b200 import lattice as _pl
        colour = _pl.colourSites(_pl.getLatticePeriodic(_pl.getUnitcell("triangular"), args.L), prefer_few_passes=(args.colouring == "few"))
    cfg = scmc.initializeCfg("nearest-neighbor", "triangular", [J], args.L, 1, n_replicas=R, precision=args.precision, device=local,
                             seed=SEED, replica_offset=roff, colour=colour, stream=args.stream)
    cfg.set_sweep_mode(args.mode, args.batch)
    lib, h = cfg._lib, cfg._h
    N = cfg.nSites
    beta = 1.0 / Ts
    # untimed thermalisation exactly as run! does it: cone width sigma starts at 60 and adapts towards 50 % acceptance
    # (src/MonteCarlo.jl:17, 47-54); 48 sweeps with an adaptation every 2 sweeps, so the timed steps run in the production regime
    therm = scmc.run_(cfg, None, Ts, np.full(R, 2.0), 48, 0, measurement_rate=2, hits=args.hits, keep_series=False)
    sigma = therm["sigma"]
    sweep_ctr = cfg.sweep_counter

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from scmc_b200 import distributed as sdist
    intervals = 1              # measurement intervals per step; raised below so that the timed region lasts >= ~1 s at every GPU count

    def device_step():
        nonlocal sweep_ctr
        for _ in range(intervals):
            _lib.check(lib.scmc_sweep_async(h, args.sweeps, args.hits, C.c_uint64(SEED), C.c_uint64(sweep_ctr)))
            sweep_ctr += args.sweeps

    acc_h = np.zeros(R, dtype=np.int64); att_h = np.zeros(R, dtype=np.int64); obs_h = np.zeros((R, _lib.NOBS))
    gathered = None

    def e2e_step():
        # what a user of the public API does per measurement interval: sweeps with host beta / sigma tables, measure! rows back on the
        # host, and -- with several GPUs -- the all-gather of the per-replica rows over NCCL (the only collective of the path)
        nonlocal sweep_ctr, gathered
        for _ in range(intervals):
            _lib.check(lib.scmc_sweep(h, args.sweeps, args.hits, _lib.dptr(beta), _lib.dptr(sigma), C.c_uint64(SEED), C.c_uint64(sweep_ctr),
                                      _lib.lptr(acc_h), _lib.lptr(att_h)))
            sweep_ctr += args.sweeps
            _lib.check(lib.scmc_measure(h, _lib.dptr(obs_h)))
            gathered = sdist.gather_replicas(obs_h, R_total) if world > 1 else obs_h

    # ---- device-resident timing
    sampler = ClockSampler(local) if rank == 0 else None       # started before the warm-up so nvidia-smi is already streaming
    t_load0 = time.time()
    for _ in range(args.warmup):
        device_step()
    barrier()
    # size the step: one more interval, timed, decides how many intervals a step holds (same number on every rank)
    tp = time.perf_counter(); device_step(); barrier(); t_one = time.perf_counter() - tp
    if world > 1:
        tt = torch.tensor([t_one], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_one = float(tt[0])
    intervals = max(1, int(np.ceil(args.min_seconds / max(1e-6, t_one * args.steps))))
    l0 = cfg.info()["launches"]
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1, t_load0) if sampler else None
    launches = cfg.info()["launches"] - l0
    ms = ev0.elapsed_time(ev1)
    # ---- end-to-end timing through the host-buffer API
    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        tot = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        launches = int(tot[0])
    # context figure: same kernel with more Metropolis hits per site visit (amortises the per-visit neighbour gather)
    extra = None
    if args.extra_hits and args.extra_hits != args.hits:
        def extra_step():
            nonlocal sweep_ctr
            _lib.check(lib.scmc_sweep_async(h, args.sweeps, args.extra_hits, C.c_uint64(SEED), C.c_uint64(sweep_ctr)))
            sweep_ctr += args.sweeps
        extra_step(); barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_extra = max(2, args.steps * intervals // 3)
        x0.record()
        for _ in range(n_extra):
            extra_step()
        x1.record(); barrier()
        ms_x = x0.elapsed_time(x1)
        if world > 1:
            tx = torch.tensor([ms_x], device="cuda", dtype=torch.float64)
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
            ms_x = float(tx[0])
        extra = {"hits_per_visit": args.extra_hits, "value": R_total * N * args.extra_hits * args.sweeps * n_extra / (ms_x * 1e-3), "unit": "site-updates/s"}
    info = cfg.info()
    # survey 8(d)(ii): the same workload with a dense random symmetric J (no structure to exploit) -- couplings are always applied as dense 16x16
    if not args.no_dense and args.coupling == "notebook":
        cfg.close()
        cfg2 = scmc.initializeCfg("nearest-neighbor", "triangular", [couplings("dense")], args.L, 1, n_replicas=R, precision=args.precision, device=local,
                                  seed=SEED, replica_offset=roff, colour=colour, stream=args.stream)
        cfg2.set_sweep_mode(args.mode, args.batch)
        th2 = scmc.run_(cfg2, None, Ts, np.full(R, 2.0), 48, 0, measurement_rate=2, hits=args.hits, keep_series=False)
        ctr2 = cfg2.sweep_counter
        def dense_step():
            nonlocal ctr2
            for _ in range(intervals):
                _lib.check(lib.scmc_sweep_async(cfg2._h, args.sweeps, args.hits, C.c_uint64(SEED), C.c_uint64(ctr2)))
                ctr2 += args.sweeps
        dense_step(); barrier()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_dense = max(2, args.steps // 3)
        d0.record()
        for _ in range(n_dense):
            dense_step()
        d1.record(); barrier()
        ms_d = d0.elapsed_time(d1)
        if world > 1:
            td = torch.tensor([ms_d], device="cuda", dtype=torch.float64)
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
            ms_d = float(td[0])
        extra = dict(extra or {})
        extra["dense_random_J"] = {"value": R_total * N * args.hits * args.sweeps * intervals * n_dense / (ms_d * 1e-3), "unit": "site-updates/s",
                                   "note": "same workload, dense random symmetric 16x16 J (J[0,0] = 0), device-resident"}
        cfg2.close()
    total_updates = R_total * N * args.hits * args.sweeps * intervals * args.steps
    value = total_updates / (ms * 1e-3)
    e2e_value = total_updates / (ms_e2e * 1e-3)
    acc_rate = float(acc_h.sum() / max(1, att_h.sum()))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        d, w = cfg.d, args.precision // 8
        n_launch_local = max(1, launches // world)
        psi_family = args.mode in (3, 4, 5, 6) or (args.mode == 0 and cfg.d == 4)
        small = 1 <= info["cluster"] <= 2 and (args.sweeps >= 4 or R * N / info["n_colours"] < 300000)
        uses_resident = args.mode in (2, 4) or (args.mode == 0 and small and (info["resident_psi"] if psi_family else info["resident"]))
        uses_tma = (not uses_resident) and psi_family and info.get("tma") and args.mode in (0, 5, 6) and args.precision == 32
        uses_tc = uses_tma and info.get("tma_tc")
        kernel_name = (f"replica-resident cluster({info['cluster']})" if uses_resident else "streaming colour-pass") + (", psi-gather" if psi_family else ", T-gather") + \
                      (", TMA-pipelined (cp.async.bulk + mbarrier ring; seam classes: direct gather)" if uses_tma else "") + \
                      (", 16x16 mat-vec on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, A in TMEM)" if uses_tc else "")
        # algorithmic bytes per kernel launch: state in + state out (4 d w bytes) of every site the launch visits
        sites_per_launch = R * N * args.sweeps * intervals * args.steps / n_launch_local
        alg_bytes = sites_per_launch * 4 * d * w
        launch_s = (ms * 1e-3) / n_launch_local
        achieved = alg_bytes / launch_s / 1e9
        # DRAM bytes of one launch from the committed ncu capture -- only while that capture belongs to the kernel sources of this tree
        traffic, traffic_src = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            same_config = uses_tma and args.L == 128 and args.stream == prof.get("stream") and args.hits == prof.get("hits_per_visit")
            if same_config and prof.get("kernel_source_hash") == kernel_source_hash():
                traffic = prof["dram_bytes_per_site_visit"] * sites_per_launch
                traffic_src = f"{prof['source']} @ git {prof.get('git_sha')}, kernel sources {prof.get('kernel_source_hash')}"
            else:
                traffic_src = "dropped: the ncu capture in profiles/traffic.json was taken on other kernel sources or another configuration"
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "site-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if args.precision == 32 else "f64", "data": "synthetic",
                "config": {"workload": f"C5: triangular {args.L}x{args.L} n=1, {R_total} replicas (64 T in [0.02,2] x copies), {args.coupling} J applied dense 16x16, "
                                       f"{intervals} measurement interval(s)/step x {args.sweeps} sweeps x {args.hits} hits/site-visit, {info['n_colours']}-colour schedule, "
                                       f"random stream v{args.stream}",
                           "replicas_per_gpu": R, "sites": N, "hits_per_visit": args.hits, "sweeps_per_step": args.sweeps * intervals, "intervals_per_step": intervals, "stream": args.stream,
                           "kernel": kernel_name,
                           "cache_note": f"state {info['device_bytes'] / 2**30:.1f} GiB per GPU >> 126 MB L2 (inputs larger than L2, no flush needed)",
                           "acceptance_rate": acc_rate, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "updates_per_site_per_launch": args.hits, "launch_us": launch_s * 1e6,
                             "note": "issue / latency bound, not HBM bound: see profiles/README.md (round 2) for the ncu summary of this kernel"},
                "e2e": {"value": e2e_value, "unit": "site-updates/s", "h2d_bytes_per_step": int(2 * R * 8) * world * intervals,      # beta and sigma cross the ABI as FP64 tables
                        "d2h_bytes_per_step": int(2 * R * 8 + R * _lib.NOBS * 8) * world * intervals, "ms_per_step": ms_e2e / args.steps,
                        "collective": (f"all_gather of the {_lib.NOBS}-value measure! rows of all {R_total} replicas over NCCL, once per interval" if world > 1 else None)},
                "gpu_launches": launches, "clocks": clocks, "extra": extra}
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample = cpu_reference_rate(args.L, J, args.cpu_seconds)
            rate1, _, sample1 = cpu_reference_rate(args.L, J, min(4.0, args.cpu_seconds), max_cores=1)
            line["cpu_baseline"] = {"value": rate, "unit": "site-updates/s", "cores": cores, "kind": "port", "sample": sample,
                                    "single_core": {"value": rate1, "unit": "site-updates/s", "cores": 1, "sample": sample1,
                                                    "note": "the reference's own execution model: one chain, one thread (no threaded code in src/)"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
