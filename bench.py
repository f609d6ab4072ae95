#!/usr/bin/env python
"""bench.py -- particle-steps/s of the svm psi-APF (BASELINE.json config 2).

Workload ("c2"): stochastic-volatility model, synthetic series T = 1000,
psi-auxiliary particle filter with N = 65536 particles, fp64.  One STEP = one
pass of the hot path over a batch of `--filters` independent replicate filters
of that configuration on each GPU (different random streams, same model) --
the shape of the reference's only parallel pattern (independent filters,
src/approx_mcmc.cpp:364-369) and of suggest_N's replications.  A single filter
cannot fill a B200 (its T = 1000 steps are strictly sequential), so the
single-filter figure is reported beside the batched one in `config`.

  value      whole-job particle-steps/s, inputs resident in HBM, CUDA events on
             the launching stream (pre-pass + filter kernels)
  e2e        the same metric through the C-ABI call bssm_b200_psi_filter with
             HOST buffers: model upload, kernels, and the device->host copy of
             the particle cube + weights the reference's .Call returns
  roofline   filter kernel against the HBM roofline: algorithmic bytes
             (8m + 12 = 20 B per particle-step: state, weight, ancestor written
             once) / kernel time
  cpu_baseline  the CPU oracle (a port of the reference's loops) on host cores

`--impl reference` times the CPU restatement of the reference's path (the
reference itself cannot be built here: no R/Rcpp/Armadillo) with all host
threads, one replicate filter per thread, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = {"N": 65536, "n": 1000, "m": 1, "rho": 0.98, "sd_ar": 0.2, "sigma": 1.0, "gen_seed": 2024}
BYTES_PER_PS = 20  # 8m (state) + 8 (weight) + 4 (ancestor), m = 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th, self.source = index, [], False, None, "nvidia-smi"

    def _nvml_handle(self):
        """NVML handle of CUDA device `index` (by PCI bus id: NVML and CUDA may enumerate differently)"""
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(self.index)
        bus = "%08x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())

    def _run_nvml(self, nv, h):
        # NVML is cheap enough to sample every few milliseconds, i.e. many times inside the timed region
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active"
                                                         for k in ("hw_slowdown", "hw_thermal_slowdown",
                                                                   "sw_thermal_slowdown", "sw_power_cap")])
            except Exception:
                pass
            time.sleep(0.005)

    def _run(self):
        try:
            nv, h = self._nvml_handle()
        except Exception:
            nv = None
        if nv is not None:
            self.source = "nvml"
            return self._run_nvml(nv, h)
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(int(float(r[0])))
                mx = max(mx, int(float(r[1])))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def make_model():
    import data
    w = WORKLOAD
    return data.svm_model(w["n"], w["rho"], w["sd_ar"], w["sigma"], w["gen_seed"])


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank, world):
    """CPU arm: the oracle (a port of the reference's loops) with all host threads."""
    if rank != 0:
        return
    import bssm_b200
    import data
    import orc
    orc.build()
    w = WORKLOAD
    # torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP, how many cores this process may use
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # the SAME workload as the GPU arm: the full T = 1000 series, N = 65536, one replicate filter per host
    # thread and step (~6 s per step); one warm-up step is enough for a CPU loop
    n_sample = w["n"]
    y = data.svm_series(w["n"], w["rho"], w["sd_ar"], w["sigma"], w["gen_seed"])[:n_sample]
    model = bssm_b200.svm(y, w["rho"], w["sd_ar"], sigma=w["sigma"])
    times = []
    cpu_warmup = min(args.warmup, 1)
    for it in range(cpu_warmup + args.steps):
        t0 = time.perf_counter()
        orc.filter_batch(model, w["N"], method="psi", seed=1, first_filter=it * cores, n_filters=cores,
                         arith=orc.ARITH_REFERENCE, threads=cores)
        dt = time.perf_counter() - t0
        if it >= cpu_warmup:
            times.append(dt)
    total = sum(times)
    value = cores * w["N"] * n_sample * args.steps / total
    sample = ("%d replicate filters (one per thread, OpenMP static), N=%d, the full series of T=%d time points per step "
              "(%d warm-up step); oracle port of the reference loops (reference not buildable here: no R/Rcpp/Armadillo); cpu=%s"
              % (cores, w["N"], w["n"], cpu_warmup, cpu_model_name()))
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": cpu_warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c2: svm psi-APF, T=1000, N=65536 particles, %d replicate filters per step (one per host thread)" % cores},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def other_configs_leg(bssm_b200):
    """BASELINE configs 3 and 5 at full size, device-timed (CUDA events inside the job): reported next to
    the headline so that every configuration of BASELINE.json has a measured number."""
    import data
    out = {}
    N3, n3 = 1 << 20, 200
    job = bssm_b200.FilterJob(data.growth_model(n3), N3, method="bsf", seed=1, n_filters=1, out_flags=0, post=2)
    for _ in range(3):
        ms = job.run()
    k = job.kernel_ms()
    out["c3"] = {"workload": "ssm_nlg growth model, bootstrap filter + particle smoother (backtrack + summary), "
                             "N=2^20, T=200, m=2", "filter_ms": k[1], "smoother_ms": k[2], "total_ms": ms,
                 "particle_steps_per_s": N3 * n3 / (ms * 1e-3)}
    job.close()
    N5, n5, L5 = 1 << 18, 500, 4
    job = bssm_b200.FilterJob(data.gbm_model(n5), N5, method="bsf", seed=1, n_filters=1, out_flags=7, post=0, L=L5)
    for _ in range(3):
        ms = job.run()
    out["c5"] = {"workload": "ssm_sde GBM, Milstein bootstrap filter, L=4 (16 sub-steps), N=2^18, T=500",
                 "total_ms": ms, "particle_steps_per_s": N5 * n5 / (ms * 1e-3),
                 "milstein_substeps_per_s": N5 * n5 * (1 << L5) / (ms * 1e-3)}
    job.close()
    return out


def cpu_baseline_leg():
    """single-thread oracle on a bounded sample (the reference never threads inside a filter)"""
    import bssm_b200
    import data
    import orc
    orc.build()
    w = WORKLOAD
    n_sample = 200
    y = data.svm_series(w["n"], w["rho"], w["sd_ar"], w["sigma"], w["gen_seed"])[:n_sample]
    model = bssm_b200.svm(y, w["rho"], w["sd_ar"], sigma=w["sigma"])
    t0 = time.perf_counter()
    orc.filter_batch(model, w["N"], method="psi", seed=1, n_filters=1, arith=orc.ARITH_REFERENCE, threads=1)
    dt = time.perf_counter() - t0
    return {"value": w["N"] * n_sample / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": "1 filter, N=%d, first %d of T=%d time points, 1 thread (the reference never threads inside "
                      "a filter); oracle port, not the reference binary; cpu=%s" % (w["N"], n_sample, w["n"], cpu_model_name())}


def bind_near_gpu(torch, local_rank):
    """With several ranks on one host every rank copies ~1 GB per end-to-end call into pinned memory: keep the rank
    (and with it the first touch of its pinned buffers) on the CPUs of the NUMA node its GPU hangs off
    (/sys/bus/pci/devices/<bus id>/local_cpulist).  Returns a note for the JSON line; does nothing if unknown."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "not bound (no local cpus listed)"
        os.sched_setaffinity(0, cpus)
        return "bound to the %d cpus local to GPU %s" % (len(cpus), bus)
    except Exception as e:  # no sysfs entry, no permission, old torch: run unbound
        return "not bound (%s)" % type(e).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--filters", type=int, default=int(os.environ.get("BENCH_FILTERS", "0")),
                    help="replicate filters per GPU per step (0 = one full wave of the device: 18)")
    ap.add_argument("--no-c4", action="store_true", help="skip the post_correct (config 4) leg")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 (default): the headline metric; c4: ONLY the post_correct leg, printed as its own "
                         "strong-scaling line (15000 draws in total whatever the number of GPUs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import bssm_b200
    from bssm_b200 import build
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    bssm_b200._lib.check(bssm_b200.lib().bssm_b200_set_device(local_rank))
    affinity = bind_near_gpu(torch, local_rank) if world > 1 else "not bound (one rank)"
    dist = None
    if world > 1:
        # NCCL's INFO lines (communicator, ranks, NVLS / ring choice) go to STDERR so that stdout stays the
        # one JSON line; BENCH_NCCL_DEBUG=off silences them
        # (the image exports NCCL_DEBUG=VERSION: it is overridden, not inherited)
        dbg = os.environ.get("BENCH_NCCL_DEBUG", "INFO")
        if dbg.lower() in ("off", "0", "none"):
            os.environ.pop("NCCL_DEBUG", None)
        else:
            os.environ["NCCL_DEBUG"] = dbg
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,COLL")
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    w = WORKLOAD
    N, n, B = w["N"], w["n"], args.filters
    model = make_model()
    launches = 0

    def c4_leg():
        """config 4: post_correct, 15000 draws x psi-APF N=256, sharded over the ranks"""
        nonlocal launches
        from bssm_b200.distributed import shard_range
        nd, Np = 15000, 256
        mo = bssm_b200.synthetic_svm_mcmc_output(model, nd)  # same inputs on every rank (seeded)
        b, e = shard_range(nd, rank, world)
        pj = bssm_b200.PostcorrJob(model, mo, Np, is_type="is2", seed=1, draw_range=(b, e))
        ll_loc = torch.empty(e - b, dtype=torch.float64, device="cuda")
        maxlen = max(shard_range(nd, r, world)[1] - shard_range(nd, r, world)[0] for r in range(world))
        pad = torch.zeros(maxlen, dtype=torch.float64, device="cuda")
        outs = [torch.empty_like(pad) for _ in range(world)]

        def c4_step():
            ms = pj.run()
            pj.logliks_to_device(ll_loc.data_ptr())
            if dist is not None:  # the single NCCL gather of per-draw log-likelihoods
                pad[:e - b] = ll_loc
                dist.all_gather(outs, pad)
            return ms
        for _ in range(args.warmup):
            c4_step()
        barrier()
        t0 = time.perf_counter()
        c4_dev = [c4_step() for _ in range(args.steps)]
        barrier()
        c4_wall = time.perf_counter() - t0
        tt = torch.tensor([sum(c4_dev), c4_wall * 1e3], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        c4_launches = pj.launches() * args.steps
        wsum = float(torch.exp(ll_loc).sum())
        c4 = {"workload": "c4: post_correct svm, 15000 draws x psi-APF N=256, T=1000, weights only, is2",
              "draws": nd, "draws_per_s": nd * args.steps / (float(tt[1]) * 1e-3),
              "particle_steps_per_s": nd * Np * n * args.steps / (float(tt[1]) * 1e-3),
              "ms_per_pass_wall": float(tt[1]) / args.steps,
              "ms_per_pass_device_max_rank": float(tt[0]) / args.steps,
              "timing": "wall clock over K passes incl. the NCCL all_gather, max over ranks",
              "sharding": "contiguous static blocks of draws per rank, one all_gather of log-likelihoods",
              "launches": c4_launches, "rank0_weight_sum": wsum}
        pj.close()
        launches += c4_launches
        return c4



    if args.workload == "c4":
        with ClockSampler(local_rank) as clk:
            c4 = c4_leg()
        if rank == 0:
            print(json.dumps({
                "metric": "post_correct draws/s", "value": c4["draws_per_s"], "unit": "draws/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": c4["ms_per_pass_wall"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": c4["workload"], "timing": c4["timing"], "sharding": c4["sharding"]},
                "post_correct": c4, "clocks": clk.summary(), "gpu_launches": launches}), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return
    if B <= 0:
        # one wave of the engine's throughput geometry: 148 SMs x 4 resident CTAs of 256 threads
        # = 592 CTAs, 32 CTAs (x 256 threads x 8 particles) per 65536-particle filter -> 18 filters
        B = 18
    job = bssm_b200.FilterJob(model, N, method="psi", seed=1, n_filters=B, first_filter=rank * B,
                              out_flags=7)
    # One step = what one psi_filter call of the reference does for these filters: update_model(theta),
    # approximate() (Newton sweeps), the smoother pre-pass, then the particle loop.  For replicate filters
    # the first three are ONE sequential O(n) chain and run on the host (job.set_theta: ~0.1 ms, DESIGN.md
    # section 3.2); they are inside the timed region, so the step time is the wall time between the
    # synchronisations (the CUDA-event time of the kernels is reported beside it).
    theta0 = np.asarray(model.theta, dtype=np.float64)
    for _ in range(args.warmup):
        job.set_theta(theta0)
        job.run()
    barrier()
    step_ms, filt_ms = [], []
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            job.set_theta(theta0)
            step_ms.append(job.run())
            filt_ms.append(job.kernel_ms()[1])
            launches += job.launches()
        torch.cuda.synchronize()
        wall_steps = time.perf_counter() - t0
        barrier()
        wall = time.perf_counter() - t0
    kernels_total_ms = float(sum(step_ms))
    dev_total_ms = max(kernels_total_ms, wall_steps * 1e3)  # host pre-pass + upload + kernels
    t = torch.tensor([dev_total_ms, wall * 1e3, kernels_total_ms], dtype=torch.float64, device="cuda")
    # the single NCCL collective of the data path: gather per-filter log-likelihoods
    ll_local = torch.empty(B, dtype=torch.float64, device="cuda")
    job.logliks_to_device(ll_local.data_ptr())
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ll_all = [torch.empty_like(ll_local) for _ in range(world)]
        dist.all_gather(ll_all, ll_local)
        ll = torch.cat(ll_all).cpu().numpy()
    else:
        ll = ll_local.cpu().numpy()
    dev_total_ms, wall_ms, kernels_ms_max = float(t[0]), float(t[1]), float(t[2])
    total_ps = world * B * N * n * args.steps
    value = total_ps / (dev_total_ms * 1e-3)

    # single-filter latency (same kernel, one cluster)
    job.close()
    single = bssm_b200.FilterJob(model, N, method="psi", seed=1, n_filters=1, out_flags=7)
    for _ in range(2):
        single.set_theta(theta0)
        single.run()

    def one_single():
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        single.set_theta(theta0)
        single.run()  # synchronises its stream
        return (time.perf_counter() - t1) * 1e3
    s_ms = min(one_single() for _ in range(3))
    single.close()

    # end to end through the one-shot C-ABI call with (pinned) host buffers
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        from bssm_b200._lib import FilterOut, c_double_p
        alpha_h = torch.empty((n + 1) * N, dtype=torch.float64).pin_memory()
        w_h = torch.empty((n + 1) * N, dtype=torch.float64).pin_memory()
        out = FilterOut()
        out.alpha = C.cast(alpha_h.data_ptr(), c_double_p)
        out.weights = C.cast(w_h.data_ptr(), c_double_p)
        ms_struct = model.c_struct()
        L = bssm_b200.lib()
        reps = 3
        L.bssm_b200_psi_filter(C.byref(ms_struct), N, 1, C.byref(out))  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            bssm_b200._lib.check(L.bssm_b200_psi_filter(C.byref(ms_struct), N, 1, C.byref(out)))
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = 8 * (11 * n + 8)  # model vectors (y, u, xbeta, mode x2, Z, T, R, a1, P1, D, C ...)
        d2h = 8 * 2 * (n + 1) * N + 32
        e2e = {"value": world * reps * N * n / float(te[0]), "unit": "particle-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "call": "bssm_b200_psi_filter (1 filter per call, alpha cube + weights to pinned host memory; finished blocks of columns are copied while the filter kernel runs)",
               "loglik_check": out.loglik}

    c4 = None
    if not args.no_c4:
        c4 = c4_leg()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    filt_ms_avg = float(np.mean(filt_ms))
    achieved = BYTES_PER_PS * B * N * n / (filt_ms_avg * 1e-3) / 1e9
    traffic, ncu_note = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("filter_kernel_dram_bytes_per_launch")
            ncu_note = dict(tj.get("ncu") or {})  # static: what the committed ncu capture says bounds the kernel
            ncu_note["traffic_measured"] = tj.get("measured")
        except Exception:
            traffic = None
    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "c2: svm psi-APF, T=1000, N=65536 particles, %d replicate filters per GPU per step" % B,
            "filters_per_gpu": B, "outputs": "state + weights + ancestors kept in HBM (20 B/particle-step)",
            "l2": "each step writes %.1f GB per GPU (>> 126 MB L2); no explicit flush" % (BYTES_PER_PS * B * N * n / 1e9),
            "single_filter_ms": s_ms, "single_filter_particle_steps_per_s": N * n / (s_ms * 1e-3),
            "wall_ms_per_step": wall_ms / args.steps,
            "cpu_affinity": affinity,
            "kernels_ms_per_step": kernels_ms_max / args.steps,
            "pre_pass": "host, inside the timed step: update_model + approximate() + smoother pre-pass + upload "
                        "(one sequential chain shared by the replicate filters)",
            "loglik_mean": float(np.mean(ll)), "loglik_sd": float(np.std(ll)),
            "replicas": "replicas only: filters are independent, one NCCL all_gather of log-likelihoods",
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "pf::filter_kernel<UngPsi1,P>", "kernel_ms": filt_ms_avg,
                     "algorithmic_bytes_per_launch": BYTES_PER_PS * B * N * n, "ncu": ncu_note},
        "clocks": clk.summary(),
        "gpu_launches": launches,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if c4 is not None:
        line["post_correct"] = c4
    if world == 1 and not args.no_c4:
        line["other_configs"] = other_configs_leg(bssm_b200)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
