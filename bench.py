#!/usr/bin/env python
"""bench.py -- CDR::run throughput (QLT + CAAS) on synthetic cubed-sphere workloads.

Metric (BASELINE.json): cell.tracer updates/s of CDR::run, and the fraction of the
HBM roofline at 40 algorithmic bytes per update.

A STEP is one pass of the hot path over the whole workload: one QLT::run() plus one
CAAS::run() over the same ncells x nt inputs (2*ncells*nt updates). set_Qm/get_Qm
are outside the timed region, as in the reference API where they are caller kernels;
CAAS clips in place, so its inputs are restored between steps, untimed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ne120x128x40]
    python bench.py --impl reference ...   # the reference's own CPU path (oracle/_ref)

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.
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

METRIC = "CDR::run cell.tracer updates/s (QLT + CAAS)"
UNIT = "updates/s"
BYTES_PER_UPDATE = 40.0          # read min, Qm, max, prev; write Qm
CST = 7                          # conserve | shapepreserve | consistent
REF_SAMPLE_NT = 640              # tracer batch of the CPU baseline (16 levels x 40)


def workload_dims(name):
    from compose_b200.workloads import CONFIGS
    return CONFIGS[name]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm)//2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def host_cores():
    """Host threads this process may use (its affinity mask, else the CPU count)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def digest_np(a):
    """Wrap-around (mod 2^64) sum of the raw 64-bit patterns: order-independent, so ranks'
    digests of disjoint cell ranges add up to the digest of the whole array."""
    import numpy as np
    with np.errstate(over="ignore"):
        return int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64))


def cpu_reference_step(ncells, config_id, nt_sample, nrep, warm, digests=None):
    """Time the reference's own QLT::run + CAAS::run (oracle/_ref, all host threads) on a
    tracer batch of the workload. Returns (per-step seconds list, info dict). If `digests`
    is a dict it receives digest_np of the reference's outputs for that batch."""
    from oracle.oracle_py import Oracle, Ref, ref_available
    o = Oracle()
    rhom, lo, q, hi, prev = o.fill_headline(ncells, config_id, 0, nt_sample)
    pts = [CST]*nt_sample
    if ref_available(omp=True):
        r = Ref(omp=True)
        # torchrun exports OMP_NUM_THREADS=1: ask for the host's cores explicitly.
        cores = r.set_num_threads(host_cores())
        oq, _, sq = r.qlt(ncells, ("bisect", False), pts, rhom, lo, q, hi, prev, nrep=warm + nrep)
        _, sc = r.caas(ncells, pts, rhom, lo, q, hi, prev, nrep=warm + nrep)
        kind = "reference"
        secs = [float(a + b) for a, b in zip(sq[warm:], sc[warm:])]
        split = {"qlt_s": float(min(sq[warm:])), "caas_s": float(min(sc[warm:]))}
        if digests is not None:
            # CAAS digest: the reference CAAS driven through its own BfbTreeAllReducer
            # (tree-ordered sums, the b200 default mode; untimed). The timed CAAS above is
            # the stock sequential-sum path.
            oc, _ = r.caas(ncells, pts, rhom, lo, q, hi, prev, tree=("bisect", False))
            digests["qlt"] = "%016x" % digest_np(oq)
            digests["caas"] = "%016x" % digest_np(oc)
    else:
        # The plain-C restatement, OpenMP over tracers.
        tree = o.bisection_tree(ncells)
        cores = host_cores()
        os.environ["OMP_NUM_THREADS"] = str(cores)
        secs, split = [], {}
        for i in range(warm + nrep):
            t0 = time.perf_counter()
            oq = o.qlt(tree, pts, rhom, lo, q, hi, prev)
            t1 = time.perf_counter()
            oc = o.caas(ncells, pts, lo, q, hi, prev, tree=tree)
            t2 = time.perf_counter()
            if i >= warm:
                secs.append(t2 - t0)
                split = {"qlt_s": t1 - t0, "caas_s": t2 - t1}
        kind = "port"
        if digests is not None:
            digests["qlt"] = "%016x" % digest_np(oq)
            digests["caas"] = "%016x" % digest_np(oc)
    info = {"cores": cores, "kind": kind,
            "sample": "%d cells x %d of the workload's tracers (one tracer batch; tracers are "
                      "independent problems), QLT::run + CAAS::run only, %d warm-up + %d reps, "
                      "reference sources + stand-in Kokkos/MPI runtime, -O2 -fopenmp"
                      % (ncells, nt_sample, warm, nrep)}
    info.update(split)
    return secs, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncells, nt, cid = workload_dims(args.workload)
    nts = min(REF_SAMPLE_NT, nt)
    digests = {}
    secs, info = cpu_reference_step(ncells, cid, nts, args.steps, args.warmup, digests)
    ms = 1e3*sum(secs)/len(secs)
    value = 2.0*ncells*nts/(ms*1e-3)
    info["value"] = value
    info["unit"] = UNIT
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, ncells, nt),
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # Parity record: compare with the b200 arm's "sample_digest" (same tracers, any N).
        "sample_digest": dict(digests, tracers=nts,
                              how="mod-2^64 sum of the raw 64-bit output words of the first "
                                  "%d tracers; caas = reference CAAS + its BfbTreeAllReducer"
                                  % nts),
    }
    print(json.dumps(line))


def config_dict(args, ncells, nt):
    return {"workload": args.workload, "ncells": ncells, "cdr_tracers": nt,
            "problem_type": "conserve|shapepreserve|consistent",
            "tree": "recursive bisection (make_tree_over_1d_mesh)",
            "step": "QLT::run + CAAS::run (tree-ordered sums), 2*ncells*nt updates",
            "l2": "inputs (>= 0.6 GB per reconstructor) exceed the 126 MB L2; no flush",
            "parallelism": ("single GPU" if args.gpus == 1 else
                            "%d GPUs, cells partitioned by subtree (each rank owns "
                            "ncells/%d contiguous cells = whole tier-0 blocks); the block "
                            "roots are exchanged once per run() (%s), tiers above "
                            "replicated in fixed tree order"
                            % (args.gpus, args.gpus,
                               "NCCL all-gather" if os.environ.get("CEDR_B200_NO_P2P") else
                               "peer-to-peer stores over NVLink + epoch flags"))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ne120x128x40")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-launches", action="store_true",
                    help="extra untimed step with per-launch CUDA events")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import compose_b200 as cb
    from compose_b200.pipeline import HostStepPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Multi-GPU exchange: direct stores into the peers' buffers over NVLink (default), or
    # torch.distributed's NCCL all-gather (CEDR_B200_NO_P2P=1).
    P2P = world > 1 and not os.environ.get("CEDR_B200_NO_P2P")
    p2p_used = []
    ncells, nt, cid = workload_dims(args.workload)
    # Subtree partition: rank r owns cells [r*nl, (r+1)*nl) of every tracer.
    if ncells % world:
        raise SystemExit("bench.py: ncells must be divisible by the number of GPUs")
    nl = ncells//world
    nt_lcl = nt

    # ---- inputs, resident in HBM before any timed region
    rhom, lo, q, hi, prev = cb.fill_headline(ncells, nt, cid, cell0=rank*nl, nlclcells=nl)

    def make(kind):
        if kind == "qlt":
            c = cb.QLT(ncells, rank=rank, nranks=world)
        else:
            c = cb.CAAS(nl, cell0=rank*nl, ncells_global=ncells, rank=rank, nranks=world)
        for _ in range(nt_lcl):
            c.declare_tracer(CST)
        c.end_tracer_declarations()
        if world > 1:
            c.enable_distributed(world)
        c.finish_setup()
        if world > 1 and P2P:
            p2p_used.append(c.enable_p2p(world))
        c.set_rhom(rhom)
        c.set_Qm(q, lo, hi, prev)
        return c

    qlt, caas = make("qlt"), make("caas")
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one_step(timed):
        caas.set_Qm(q, lo, hi, prev)          # restore (CAAS clips in place); untimed
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        qlt.run()
        e1.record()
        caas.run()
        e2.record()
        return e0, e1, e2

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_qlt = [a.elapsed_time(b) for a, b, _ in evs]
    t_caas = [b.elapsed_time(c) for _, b, c in evs]
    ms_step = (sum(t_qlt) + sum(t_caas))/args.steps
    launches = (qlt.last_run_launches() + caas.last_run_launches())*args.steps
    t = torch.tensor([ms_step, sum(t_qlt)/args.steps, sum(t_caas)/args.steps],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_qlt, ms_caas = (float(x) for x in t.cpu())
    updates = float(ncells)*nt
    value = 2.0*updates/(ms_step*1e-3)

    # ---- parity record: digests of the outputs of the last timed step (all tracers, and
    # the first REF_SAMPLE_NT tracers = the reference arm's batch), summed over the ranks'
    # disjoint cell ranges. Identical at every N and equal to the reference arm's
    # sample_digest iff the results are bit-identical.
    nts = min(REF_SAMPLE_NT, nt)
    dig = []
    for c in (qlt, caas):
        o = c.get_Qm()
        dig += [o.view(torch.int64).sum(), o[:nts].contiguous().view(torch.int64).sum()]
        del o
    dig = torch.stack(dig)
    if world > 1:
        dist.all_reduce(dig, op=dist.ReduceOp.SUM)
    dig = ["%016x" % (int(v) & 0xffffffffffffffff) for v in dig.cpu()]
    output_digest = {"qlt": dig[0], "caas": dig[2]}
    sample_digest = {"qlt": dig[1], "caas": dig[3], "tracers": nts}

    # ---- per-launch breakdown (untimed extra step)
    kernels = None
    if args.profile_launches or True:
        for c in (qlt, caas):
            c.set_profiling(True)
        one_step(False)
        kernels = {"qlt": [(n, tr, round(ms, 4)) for n, tr, ms in qlt.launch_times()],
                   "caas": [(n, tr, round(ms, 4)) for n, tr, ms in caas.launch_times()]}
        for c in (qlt, caas):
            c.set_profiling(False)

    # ---- the caller's whole device-resident step: scatter + run + gather
    # through set_Qm / get_Qm against run() on bound arrays (no copies), QLT then CAAS.
    caller_step = None
    if nl % 2 == 0:
        def timed(fn, reps=3):
            fn()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)/reps
        outq = torch.empty_like(q)
        qc = q.clone()

        def copy_step():
            qlt.set_Qm(q, lo, hi, prev)
            qlt.run()
            qlt.get_Qm(out=outq)
            caas.set_Qm(q, lo, hi, prev)
            caas.run()
            caas.get_Qm(out=outq)
        ms_copy = timed(copy_step)
        qlt.bind_arrays(q, lo, hi, prev, out=outq)
        caas.bind_arrays(qc, lo, hi, prev)

        def bound_step():
            qlt.run()
            caas.run()      # in place on qc: later repetitions redistribute a solved field
        ms_bound = timed(bound_step)
        qlt.bind_arrays(None, None, None)
        caas.bind_arrays(None, None, None)
        caller_step = {"set_run_get_ms": ms_copy, "bound_run_ms": ms_bound,
                       "run_ms": ms_step,
                       "note": "device-resident caller arrays; bound = cedr_b200_bind_arrays "
                               "(run() reads the caller's SoA arrays, no set_Qm/get_Qm kernels)"}
        del outq, qc

    # ---- roofline of the dominant reconstructor pass (QLT run()): algorithmic bytes
    # (40 B x the updates one run() processes on this GPU) over the CUDA-event duration
    # of the run's launches; traffic = DRAM bytes of the same kernels from the committed
    # ncu capture (profiles/r02b_traffic.json, bytes per update x updates) -- NOT measured in
    # this run (ncu cannot run inside a timed bench): "traffic_source" says so.
    peak, peak_src = measured_peaks()
    upd_gpu = updates/world
    ach = BYTES_PER_UPDATE*upd_gpu/(ms_qlt*1e-3)/1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02b_traffic.json")
    bpu = json.load(open(tpath))["bytes_per_update"] if os.path.exists(tpath) else None
    per_kernel = {}
    if kernels:
        # Each kernel against the bytes it must move itself: up reads 4 rows (32 B), down
        # reads 3 rows and writes 1 (32 B), caas_adjust likewise; mid (the block tops of the
        # down-sweep) reads the depth-7 sums and writes the depth-7 masses: 4 KB per block x
        # tracer (a block is ncells / 2^k <= 1024 leaves: 675 at ne30 / ne120, 768 at ne256).
        nblk = 1
        while ncells/nblk > 1024:
            nblk *= 2
        mid_b = 4096.0/(ncells/nblk)
        for kind in ("qlt", "caas"):
            for name, tier, ms in kernels[kind]:
                if tier == 0 and name in ("up", "down", "caas_adjust", "fused", "mid") and ms > 0:
                    b = 40.0 if name == "fused" else mid_b if name == "mid" else 32.0
                    g = b*upd_gpu/(ms*1e-3)/1e9
                    per_kernel["%s.%s" % (kind, name)] = {
                        "ms": ms, "bytes_per_update": b, "achieved": g, "frac": g/peak}
    if bpu:
        traffic = bpu["qlt"]["run_total"]*upd_gpu
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach/peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "QLT::run(): fast::up_kernel + tier-1 sweep + fast::midT_kernel + "
                          "fast::down3_kernel (dominant: down3_kernel); per GPU; duration = "
                          "CUDA events around "
                          "run() on its stream",
                "algorithmic_bytes_per_update": BYTES_PER_UPDATE,
                "algorithmic_bytes": BYTES_PER_UPDATE*upd_gpu,
                "traffic_bytes_per_update": bpu["qlt"]["run_total"] if bpu else None,
                "traffic_source": ("committed ncu --set full capture of the same kernels "
                                   "(profiles/r02b_traffic.json, profiles/r02b_qlt_ne120_ncu.txt), "
                                   "bytes per update x this run's updates; not re-measured here"),
                "kernels": per_kernel,
                "kernels_note": "per-kernel frac is against the measured COPY bandwidth (read + "
                                "write); a read-mostly stream such as the up-sweep can exceed it",
                "caas": {"achieved": BYTES_PER_UPDATE*upd_gpu/(ms_caas*1e-3)/1e9,
                         "frac": BYTES_PER_UPDATE*upd_gpu/(ms_caas*1e-3)/1e9/peak,
                         "traffic": bpu["caas"]["run_total"]*upd_gpu if bpu else None}}

    # ---- end to end through the public API with HOST buffers
    e2e = None
    if not args.no_e2e:
        del qlt, caas
        torch.cuda.empty_cache()
        from compose_b200.pipeline import near_gpu
        # Pinned buffers and the copy-issuing thread on the GPU's NUMA node (restored
        # afterwards: the CPU baseline leg uses every core).
        with near_gpu(local_rank) as ng:
            pipe = HostStepPipeline(ncells, nt_lcl, rank=rank, nranks=world, p2p=P2P)
            pin = lambda x: x.cpu().pin_memory()
            rhom_h, lo_h, q_h, hi_h, prev_h = (pin(x) for x in (rhom, lo, q, hi, prev))
            out_h = {k: torch.empty((nt_lcl, nl), dtype=torch.float64).pin_memory()
                     for k in pipe.kinds}
            for _ in range(2):
                pipe.step(rhom_h, lo_h, q_h, hi_h, prev_h, out_h)       # warm-up
            barrier()
            t0 = time.perf_counter()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(args.steps):
                pipe.step(rhom_h, lo_h, q_h, hi_h, prev_h, out_h)
            e1.record()
            barrier()
            wall = time.perf_counter() - t0
        ms_e2e = max(e0.elapsed_time(e1), 0.0)/args.steps
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.cpu()[0])
        e2e = {"value": 2.0*updates/(ms_e2e*1e-3), "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes_per_step*world,
               "d2h_bytes_per_step": pipe.d2h_bytes_per_step*world,
               "ms_per_step": ms_e2e, "wall_ms_per_step": 1e3*wall/args.steps,
               "host_gb_per_s": (pipe.h2d_bytes_per_step + pipe.d2h_bytes_per_step)*world
                                /(ms_e2e*1e-3)/1e9,
               "numa_local_cpus": len(ng.cpus) if ng.cpus else None,
               "how": "pinned host SoA arrays -> chunked H2D / run() on the bound chunk arrays "
                      "/ D2H pipeline over %d tracer chunks on %d streams "
                      "(compose_b200.pipeline; bound=%s)"
                      % (pipe.nchunks, len(pipe.slots), pipe.bound)}
        launches += pipe.launches_per_step*args.steps

    # ---- config 5 flavour: many tiny runs (111 cells x 1 tracer), device time per run()
    small = None
    if rank == 0:
        small = {"workload": "111 cells x 1 tracer (cedr_test_1d_transport size), 200 "
                             "back-to-back run() calls, CUDA events"}
        r1, l1, q1, h1, p1 = cb.fill_headline(111, 1, 5)
        for kind in ("qlt", "caas"):
            c = cb.QLT(111) if kind == "qlt" else cb.CAAS(111)
            c.declare_tracer(3)
            c.end_tracer_declarations()
            c.finish_setup()
            c.set_rhom(r1)
            c.set_Qm(q1, l1, h1, p1)
            for _ in range(20):
                c.run()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(200):
                c.run()
            e1.record()
            torch.cuda.synchronize()
            small[kind + "_us_per_run"] = 1e3*e0.elapsed_time(e1)/200
            small[kind + "_launches_per_run"] = c.last_run_launches()
            # The whole transport step on the device (interpolation + set_Qm, run, get_Qm:
            # cedr_b200_transport1d_cycle), 351 steps, launched one by one and replayed
            # from a CUDA graph.
            import numpy as np
            y0 = np.linspace(0.1, 0.9, 112)
            y0[-1] = y0[0]
            # ... and as ONE launch for the whole cycle (a persistent CTA, t1d_cycle_kernel).
            for g, tag in ((False, "launches"), (True, "graph"), (2, "one_launch")):
                # (twice, the faster: a kernel's first launch loads its module)
                us = min(c.transport1d_cycle(351, y0, use_graph=g)[1] for _ in range(2))
                small["%s_t1d_step_us_%s" % (kind, tag)] = us

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_dig = {}
        secs, cpu = cpu_reference_step(ncells, cid, nts, 3, 1, ref_dig)
        cpu["value"] = 2.0*ncells*nts/(sum(secs)/len(secs))
        cpu["unit"] = UNIT
        cpu["sample_digest"] = ref_dig
        sample_digest["matches_cpu_reference"] = (ref_dig.get("qlt") == sample_digest["qlt"] and
                                                  ref_dig.get("caas") == sample_digest["caas"])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, ncells, nt),
            "qlt": {"value": updates/(ms_qlt*1e-3), "ms_per_run": ms_qlt},
            "caas": {"value": updates/(ms_caas*1e-3), "ms_per_run": ms_caas},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "output_digest": output_digest, "sample_digest": sample_digest,
            "gpu_launches": launches, "clocks": clocks, "kernels_ms": kernels,
            "small_problem": small, "caller_step": caller_step,
            "exchange": (None if world == 1 else
                         "p2p (NVLink stores + epoch flags)" if p2p_used and all(p2p_used)
                         else "nccl all-gather"),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
