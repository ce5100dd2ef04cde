#!/usr/bin/env python
"""bench.py -- unique ERIs/s of the int2e hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload h2o_64] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; ranks shard the packed quartet space, no
   collective on the data path; only the timing barrier / max-over-ranks use NCCL.)

A "step" is one full evaluation of this rank's shard of the packed 8-fold-unique ERI array of the
workload molecule: zero fill + six quartet-class kernels, inputs (shell-pair tables, Boys table)
already resident in HBM.  `value` = unique ERIs of the whole molecule / max-over-ranks step time.
`e2e` is the same metric through the host-buffer C-ABI call a user makes (plan build, H2D of the
pair tables, kernels, D2H of the packed slice into pinned host memory inside the timed region).

`--impl reference` times the CPU restatement of the reference's own loop (oracle/, the reference
is Fortran and cannot be built in this image) on all host cores over a bounded sample of the
same workload, extrapolated to the whole molecule.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")
METRIC = "unique_eris_per_s"
UNIT = "ERI/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_system(workload: str):
    import myqc_b200 as Q
    from myqc_b200 import molecules
    d = os.path.join(INPUTS, workload)
    zm = open(os.path.join(d, "ZMAT")).read() if os.path.isdir(d) else molecules.zmat(workload)
    with tempfile.TemporaryDirectory() as tmp:
        return Q.make_job(tmp, zm, INPUTS), zm


# ----------------------------------------------------------------------------------------------
def ordered_survivors(mol, b) -> int:
    """Number of ordered set quartets (a,b,c,d) that pass the reference's EIJ*EGH >= 1e-14 test
    (int2e.f90:257): the clmnew calls of a full reference run.  Exact, from the sorted prefactors."""
    setl = int(b.setinfo[1])
    n = b.nset
    cen = np.array([b.setinfo[1 + s * setl + 3] for s in range(n)])
    al = b.set[:n]
    xyz = mol.xyz[cen]
    r2 = ((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1)
    E = np.exp(-(al[:, None] * al[None, :]) * r2 / (al[:, None] + al[None, :])).ravel()
    Es = np.sort(E)
    # for each E_ab the number of E_cd with E_ab*E_cd >= 1e-14 (product test done in floating point on
    # the candidates next to the boundary would change a handful of counts out of ~1e10: ignored)
    with np.errstate(divide="ignore"):
        thr = 1.0e-14 / Es
    cnt = len(Es) - np.searchsorted(Es, thr, side="left")
    return int(cnt.sum())


def cpu_sample(zm: str, seconds: float, nthreads: int, offset: int = 0):
    """Bounded sample of the reference's loop on the host (oracle port).  Returns
    (unique ERIs/s extrapolated to the whole molecule, description, seconds spent)."""
    from oracle import oracle as O
    mol = O.parse_zmat(zm)
    b = O.build_basis(open(os.path.join(INPUTS, "mybasis")).read(), mol.atoms)
    ft = O.read_ftab(os.path.join(INPUTS, "Ftab"))
    n = b.nset
    total = n * n
    npair = b.norb * (b.norb + 1) // 2
    nunique = npair * (npair + 1) // 2
    # calibrate on a small stride sample, then size the real sample for `seconds`
    ncal = min(total, 8 * nthreads)
    idx = (np.arange(ncal, dtype=np.int64) * (total // ncal) + offset) % total
    dt, _ = O.int2e_sample(mol, b, ft, idx // n, idx % n, nthreads)
    per = max(dt / ncal, 1e-7)
    ns = int(min(total, max(ncal, seconds / per)))
    idx = (np.arange(ns, dtype=np.int64) * (total // ns) + offset) % total
    dt, surv = O.int2e_sample(mol, b, ft, idx // n, idx % n, nthreads)
    if ns == total:
        full, how = dt, "full run"
    else:
        # the loop's cost is dominated by the surviving quartets (one clmnew call each): extrapolate
        # by their exact count, which is far less noisy than the row count
        tot_surv = ordered_survivors(mol, b)
        full = dt * tot_surv / max(surv, 1)
        how = f"extrapolated by surviving-quartet count ({tot_surv} in the full loop)"
    desc = (f"oracle port of int2e.f90 loop: {ns} of {total} ordered (a,b) set-pair rows "
            f"(stride sample, {surv} surviving quartets, {dt:.1f} s on {nthreads} thread(s)), {how}")
    return nunique / full, desc, dt


def make_config(workload, s, nq_total, model_flops, rank0_elems, world):
    """The `config` object of the JSON line.  Both arms print the SAME object (the reference arm is timed on the GPU
    arm's config), so it is built from host-side quantities only."""
    need_flush = 8 * rank0_elems < 512e6
    return {"workload": workload, "norb": s.norb, "nset": s.nset, "unique_eris": s.nunique,
            "canonical_prim_quartets": int(nq_total), "model_flops": model_flops,
            "l2": "flushed between steps (256 MB scratch write)" if need_flush else
                  "output slice per step (%.1f GB) is larger than L2" % (8 * rank0_elems / 1e9),
            "parallelism": f"quartet-space row shards x{world}, no collective"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s, zm = build_system(args.workload)
    # the GPU arm's config (host-only library calls: canonical work statistics and the shard layout)
    import myqc_b200 as Q
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    nq, model_flops = Q.canonical_stats(s)
    off = Q.shard_layout(s, world)
    config = make_config(args.workload, s, nq.sum(), model_flops, int(off[1] - off[0]), world)
    nthreads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for w in range(args.warmup):
        cpu_sample(zm, min(per_step, 2.0), nthreads, offset=w)
    vals, descs, t_tot = [], [], 0.0
    for k in range(args.steps):
        v, d, dt = cpu_sample(zm, per_step, nthreads, offset=args.warmup + k)
        vals.append(v); descs.append(d); t_tot += dt
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s.nunique / value,
        "ms_per_step_is": "extrapolated from the timed sample to the whole molecule (see cpu_baseline.sample)",
        "sample_seconds_timed": t_tot,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": descs[-1]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(args._stdout, json.dumps(line))


# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import myqc_b200 as Q

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the ERI engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    s, zm = build_system(args.workload)
    nq, model_flops = Q.canonical_stats(s)
    plan = Q.Plan(s, device=local, shard=rank, nshards=world)
    out = torch.empty(max(plan.out_elems, 1), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    nlaunch = plan.stats()["nlaunch"]
    log(f"[rank {rank}] {args.workload}: norb={s.norb} unique={s.nunique} shard elems={plan.out_elems} "
        f"({8 * plan.out_elems / 1e9:.2f} GB) launches/step={nlaunch}")

    fp64_peak = Q.fp64_peak(local) if rank == 0 else 0.0
    hbm_peak, peak_src = measured_peaks()

    # L2: the slice each step writes is far larger than the 126 MB L2 for the headline workload;
    # for small workloads flush L2 between steps by writing a 256 MB scratch buffer.
    need_flush = 8 * plan.out_elems < 512e6
    scratch = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device="cuda") if need_flush else None

    def step():
        if scratch is not None:
            scratch.zero_()
        plan.execute(out.data_ptr(), stream)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.5)  # nvidia-smi needs a moment before its first sample
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    barrier()
    for k in range(args.steps):
        if scratch is not None:
            scratch.zero_()
        ev[2 * k].record()
        plan.execute(out.data_ptr(), stream)
        ev[2 * k + 1].record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms_local = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)) / args.steps
    ms = allmax(ms_local)
    value = s.nunique / (ms * 1e-3)
    if world > 1:
        tt = torch.tensor([ms_local, float(plan.out_elems)], dtype=torch.float64, device="cuda")
        gl = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(gl, tt)
        per_rank = [{"rank": r, "ms": float(g[0].item()), "slice_gb": 8e-9 * float(g[1].item())} for r, g in enumerate(gl)]
    else:
        per_rank = [{"rank": 0, "ms": ms_local, "slice_gb": 8e-9 * plan.out_elems}]

    # per-launch breakdown (CUDA events around every launch, on the launching stream)
    launches = plan.launches()
    acc = np.zeros(len(launches))
    nrep = max(1, min(args.steps, 5))
    for _ in range(nrep):
        if scratch is not None:
            scratch.zero_()
        acc += np.array(plan.execute_timed(out.data_ptr(), stream))
    acc /= nrep
    exq, schwarz_tau = plan.executed_quartets()  # device counters of the last execute
    barrier()
    checksum = float(out[:plan.out_elems].sum().item()) if plan.out_elems else 0.0
    checksum = allsum(checksum)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        off = Q.shard_layout(s, world)
        nloc = int(off[rank + 1] - off[rank])
        host = torch.empty(max(nloc, 1), dtype=torch.float64, pin_memory=True)
        harr = host.numpy()
        Q.eri_packed_shard(s, harr[:nloc], device=local, shard=rank, nshards=world)  # warm-up
        barrier()
        ne2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ne2e):
            Q.eri_packed_shard(s, harr[:nloc], device=local, shard=rank, nshards=world)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ne2e
        dt = allmax(dt)
        h2d = Q.plan_h2d_bytes(s)
        d2h = int(allsum(float(Q.last_d2h_bytes())))  # what crossed PCIe (sparse route: nonzero chunks + flags)
        e2e = {"value": s.nunique / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": d2h, "host_bytes_written_per_step": int(8 * s.nunique),
               "d2h_route": "sparse push of the nonzero 256-byte chunks by the GPU + zeros of the other chunks by host threads (streaming stores)"
               if d2h < 8 * s.nunique else "cudaMemcpy",
               "ms_per_step": dt * 1e3, "steps": ne2e,
               "host_checksum": float(allsum(float(harr[:nloc].sum())))}
    except Exception as exc:  # report, do not hide
        e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel breakdown and the rooflines of the step ------------------------------------------
    # One execute() = the zero fill + the six quartet-class kernels ((SP SP|SP SP) as four mu-slices).  Times are
    # CUDA events around every launch of a serialised pass; launches of one kernel family are added up.
    # the zero fill is the driver's cudaMemsetAsync by default (7.35 TB/s on a B200 against 6.2-6.55 TB/s for every SM
    # fill kernel tried; MYQC_FILL_ENGINE=kernel selects the repo's fill_zero_kernel): write-only, so its rate can
    # exceed MEASURED_PEAKS' copy (read + write) bandwidth, and it is not one of this repo's kernel launches
    fill_engine = os.environ.get("MYQC_FILL_ENGINE", "memset")
    if fill_engine not in ("kernel", "copy") or os.environ.get("MYQC_OUTPUT_MODE") == "compose":
        fill_engine = "memset" if os.environ.get("MYQC_OUTPUT_MODE") != "compose" else "kernel"
    own_launches = nlaunch - (0 if fill_engine == "kernel" else sum(1 for (cls, tri, rows) in launches if cls == -1))
    per_class_ms = {}
    fill_elems = 0
    for (cls, tri, rows), t in zip(launches, acc):
        per_class_ms[cls] = per_class_ms.get(cls, 0.0) + float(t)
        if cls < 0:
            fill_elems += int(rows)
    serial_ms = float(acc.sum())
    kernels = []
    for cls, t in sorted(per_class_ms.items()):
        if cls < 0:
            # cls -2: compose_kernel (writes every element of the slice once: algorithmic bytes = 8 B per unique ERI);
            # cls -1: fill_zero_kernel of the scatter mode (MYQC_OUTPUT_MODE=scatter)
            gb = 8.0 * fill_elems / 1e9
            kernels.append({"kernel": "compose" if cls == -2 else ("fill_zero" if fill_engine == "kernel" else
                                                                   "zero fill (cudaMemsetAsync)" if fill_engine == "memset" else
                                                                   "zero fill (device-to-device copies of a zero buffer)"), "elements_written": fill_elems, "ms": t, "bound": "hbm",
                            "share_of_step": t / serial_ms,
                            "achieved": gb / (t * 1e-3) if t > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s"})
        else:
            # canonical primitive quartets of the class x W(class); for a shard the class counts of the whole molecule
            # are scaled by this rank's share of the model flops (the plan cuts shards by that weight)
            fl = Q.CLASS_W[cls] * float(nq[cls]) / world
            kernels.append({"kernel": "eri_class" + Q.CLASS_NAMES[cls], "ms": t, "bound": "fp64", "share_of_step": t / serial_ms,
                            "achieved": fl / (t * 1e-3) / 1e12 if t > 0 else 0.0, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac_of_nominal_peak": fl / (t * 1e-3) / 1e12 / Q.FP64_NOMINAL_TFLOPS if t > 0 else None})
    for k in kernels:
        k["frac"] = k["achieved"] / k["peak"] if k["peak"] else None
    for k in kernels:
        if k["bound"] == "fp64":
            c = Q.CLASS_NAMES.index(k["kernel"][len("eri_class"):])
            k["prim_quartets_evaluated"] = int(exq[c])
            k["frac_evaluated"] = Q.CLASS_W[c] * exq[c] / (k["ms"] * 1e-3) / 1e12 / fp64_peak if k["ms"] > 0 and fp64_peak else None
    flops_eval = float(sum(Q.CLASS_W[c] * exq[c] for c in range(6)))
    cls_ms = sum(k["ms"] for k in kernels if k["bound"] == "fp64")
    family = {"kernel": "eri_class_kernel family (all class launches of one step)", "ms": cls_ms, "share_of_step": cls_ms / serial_ms,
              "fp64_tflops_model": (model_flops / world) / (cls_ms * 1e-3) / 1e12 if cls_ms > 0 else 0.0}
    family["frac"] = family["fp64_tflops_model"] / fp64_peak if fp64_peak else None
    family["frac_of_nominal_peak"] = family["fp64_tflops_model"] / Q.FP64_NOMINAL_TFLOPS
    # The roofline object is the STEP against its binding roofline (the larger of the two roofline times), with both
    # fractions spelled out; `dominant_family` is the kernel family with the largest share of the step.
    bytes_alg = 8.0 * plan.out_elems
    hbm_t = bytes_alg / (hbm_peak * 1e9)
    fp_t = (model_flops / world) / (fp64_peak * 1e12) if fp64_peak else 0.0
    bound = "hbm" if hbm_t >= fp_t else "fp64"
    if bound == "hbm":
        achieved, peak, unit = bytes_alg / 1e9 / (ms_local * 1e-3), hbm_peak, "GB/s"
    else:
        achieved, peak, unit = (model_flops / world) / (ms_local * 1e-3) / 1e12, fp64_peak, "TFLOP/s"
    traffic = None
    traffic_src = None
    tfile = os.path.join(ROOT, "profiles", "r2g_h2o_64_dram_traffic_bytes.json")
    if args.workload == "h2o_64" and world == 1 and os.path.exists(tfile):
        per_kernel = json.load(open(tfile))
        traffic = float(sum(per_kernel.values()))
        traffic_src = ("sum over the step's class-kernel launches of dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                       "ncu --set full (profiles/r2g_h2o_64_ncu_full.json)")
        if not any("fill" in k for k in per_kernel):
            # ncu does not see the driver's memset; it writes every byte of the slice once (the repo's fill kernel,
            # which ncu does see, moved 40.40 GB for 40.46 GB: profiles/r2f_h2o64_dram_traffic_bytes.json)
            traffic += bytes_alg
            traffic_src += " + 8 B per element for the cudaMemsetAsync zero fill, which ncu does not capture"
    roofline = {"kernel": "step (zero fill + all class launches of one execute)", "bound": bound, "achieved": achieved, "peak": peak,
                "unit": unit, "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src if bound == "hbm" else "measured DFMA microbenchmark in this run (myqc_fp64_peak)",
                "algorithmic_bytes": bytes_alg, "model_flops": model_flops / world,
                "hbm_frac": bytes_alg / 1e9 / (ms_local * 1e-3) / hbm_peak,
                "fp64_frac_measured_peak": (model_flops / world) / (ms_local * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                "fp64_frac_nominal_peak": (model_flops / world) / (ms_local * 1e-3) / 1e12 / Q.FP64_NOMINAL_TFLOPS,
                "fp64_peak_nominal_tflops": Q.FP64_NOMINAL_TFLOPS,
                "dominant_family": family, "serialised_launch_sum_ms": serial_ms,
                # model flops credit the reference's rule; the Schwarz skip evaluates fewer quartets
                "schwarz": {"tau": schwarz_tau, "prim_quartets_reference_rule": int(sum(nq)) // world if world > 1 else int(sum(nq)),
                            "prim_quartets_evaluated_this_rank": int(sum(exq)), "model_flops_evaluated_this_rank": flops_eval,
                            "fp64_frac_evaluated": flops_eval / (ms_local * 1e-3) / 1e12 / fp64_peak if fp64_peak else None}}
    whole = {"fp64_tflops_model": model_flops / (ms * 1e-3) / 1e12 / 1.0,
             "fp64_frac_of_measured_dfma_peak": model_flops / (ms * 1e-3) / 1e12 / (fp64_peak * world) if fp64_peak else None,
             "fp64_frac_of_nominal_peak": model_flops / (ms * 1e-3) / 1e12 / (Q.FP64_NOMINAL_TFLOPS * world),
             "hbm_gbs_algorithmic": 8.0 * s.nunique / 1e9 / (ms * 1e-3),
             "hbm_frac": 8.0 * s.nunique / 1e9 / (ms * 1e-3) / (hbm_peak * world)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        v, desc, _ = cpu_sample(zm, args.cpu_seconds, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": make_config(args.workload, s, nq.sum(), model_flops, plan.out_elems, world),
        "roofline": roofline, "kernels": kernels, "whole_step": whole,
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(own_launches * args.steps),
        "zero_fill_engine": fill_engine,
        "clocks": clk, "fp64_peak_tflops_measured": fp64_peak, "checksum": checksum, "per_rank": per_rank,
    }
    _emit(args._stdout, json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Route fd 1 to stderr until the JSON line is printed, so that library banners (NCCL prints its
    version to stdout) cannot end up in front of the one line the driver parses."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd, line: str):
    sys.stdout.flush()
    os.dup2(saved_fd, 1)
    os.close(saved_fd)
    print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="h2o_64")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="experiments only: skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    args._stdout = _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
