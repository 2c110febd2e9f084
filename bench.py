#!/usr/bin/env python
"""Benchmark of the covariance hot path (BASELINE.json metric and configs).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference itself on host cores

One "step" is one pass of the hot path over BASELINE.json configs[1]: the dense float64
Gram matrices of Matern-5/2 AND RatQuad over N = 32768 points in d = 3 (2 * N^2 entries,
17.2 GB written).  With G > 1 GPUs (launched by torchrun, one rank per GPU) the units are
independent point clouds: each rank builds the two Gram matrices of its own 32768-point cloud
(seeded per rank) -- weak scaling, identical per-GPU work, no collective on the data path.  The JSON line also carries, under "extra", the single-kernel numbers (ExpQuad,
Matern-5/2, RatQuad) and the matrix-free K.v of configs[2] (N = 2^20, d = 8), which IS
row-sharded with an NCCL all-gather of the result vector.

Timing: CUDA events on the stream the kernels are launched on, W warm-up steps, exactly K timed
steps between barrier + synchronize, max over ranks.  Every step writes 8.6 GB per kernel, far
more than the 126 MB L2, so no L2 flush is needed between iterations.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GRAM = 32768
D_GRAM = 3
N_MV = 1 << 20
D_MV = 8
METRIC = "gram_entries_per_s"
UNIT = "entries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-matvec", action="store_true", help="skip the K.v extra (configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--n", type=int, default=N_GRAM, help="Gram size (default: BASELINE config)")
    ap.add_argument("--n-mv", type=int, default=N_MV)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (NumPy restatement of the reference algorithm)
# --------------------------------------------------------------------------------------------
def _oracle_specs():
    from oracle import covfuncs_oracle as orc  # noqa: PLC0415

    return [orc.KernelSpec("matern", (D_GRAM,), 1.0, nu=2.5),
            orc.KernelSpec("ratquad", (D_GRAM,), 1.0, alpha=1.0)]


def _import_reference():
    """The unmodified reference (pip-installed into baseline/_ref, which travels to the GPU box),
    or None.  Only the four NumPy-2 aliases of SURVEY.md 8c are added before the import."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "probnum")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    for name, val in (("float_", np.float64), ("Inf", np.inf),
                      ("find_common_type", lambda a, s: np.result_type(*a, *s)),
                      ("AxisError", np.exceptions.AxisError)):
        if not hasattr(np, name):
            setattr(np, name, val)
    try:
        import warnings  # noqa: PLC0415

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from probnum.randprocs import covfuncs  # noqa: PLC0415
        return covfuncs
    except Exception:  # noqa: BLE001 - any import problem means "not available"
        return None


def _cpu_rows(args):
    """Worker: both Gram kernels for rows [lo, hi) against all N points (oracle)."""
    lo, hi, n = args
    from oracle import covfuncs_oracle as orc  # noqa: PLC0415

    x = np.random.default_rng(42).normal(size=(n, D_GRAM))
    t0 = time.perf_counter()
    for spec in _oracle_specs():
        orc.gram(spec, x[lo:hi], x, row_block=512)
    return time.perf_counter() - t0


def _ref_rows(args):
    """Worker: the same rows through the reference's own ``CovarianceFunction.matrix``."""
    lo, hi, n = args
    cf = _import_reference()
    x = np.random.default_rng(42).normal(size=(n, D_GRAM))
    t0 = time.perf_counter()
    for k in (cf.Matern((D_GRAM,), nu=2.5), cf.RatQuad((D_GRAM,))):
        k.matrix(x[lo:hi], x)
    return time.perf_counter() - t0


def cpu_baseline_sample(n: int, rows: int = 2048):
    """Single-process CPU run on a bounded sample: `rows` rows of both Gram matrices, through the
    unmodified reference when baseline/_ref is importable, else through the oracle port."""
    have_ref = _import_reference() is not None
    dt = 0.0
    for lo in range(0, rows, 256):  # 256-row blocks bound the reference's (rows, N, d) temporary
        dt += (_ref_rows if have_ref else _cpu_rows)((lo, min(lo + 256, rows), n))
    entries = 2.0 * rows * n
    what = ("probnum (unmodified, baseline/_ref) CovarianceFunction.matrix" if have_ref
            else "oracle (NumPy restatement of the reference)")
    return {
        "value": entries / dt, "unit": UNIT, "cores": 1, "kind": "reference" if have_ref else "port",
        "sample": f"{what}, rows 0..{rows} of the Matern-5/2 and RatQuad Gram matrices in 256-row blocks, "
                  f"N={n}, d={D_GRAM}: {entries:.3g} entries in {dt:.1f} s",
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores, bounded
    sample per step.  Uses the unmodified reference from baseline/_ref (kind "reference") when it
    is importable, else the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp  # noqa: PLC0415

    n = args.n
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    rows_per_worker = 256
    jobs = [(w * rows_per_worker, (w + 1) * rows_per_worker, n) for w in range(workers)]
    entries_per_step = 2.0 * workers * rows_per_worker * n
    have_ref = _import_reference() is not None
    worker, kind = (_ref_rows, "reference") if have_ref else (_cpu_rows, "port")
    what = ("probnum (unmodified, baseline/_ref) CovarianceFunction.matrix" if have_ref
            else "oracle port (NumPy restatement of the reference)")
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in range(args.warmup):
            pool.map(worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(worker, jobs)
        dt = time.perf_counter() - t0
    value = entries_per_step * args.steps / dt
    sample = (f"{what}, {workers} processes x {rows_per_worker} rows x {n} columns of the Matern-5/2 "
              f"and RatQuad Gram matrices per step ({entries_per_step:.3g} entries/step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"Matern-5/2 + RatQuad Gram, N={n}, d={D_GRAM}, float64 (BASELINE configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def _bind_to_gpu_numa_node(index: int) -> None:
    """One process per GPU: run on the CPUs NVML names as local to this GPU, so that the pinned
    result buffers of the end-to-end leg are allocated on the GPU's own NUMA node and its D2H
    traffic does not cross the socket interconnect.  Best effort; silent if NVML is unavailable."""
    try:
        import pynvml  # noqa: PLC0415

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            print(f"[bench] rank on GPU {index}: bound to {len(cpus)} CPUs local to the GPU", file=sys.stderr)
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] no NUMA binding for GPU {index}: {exc}", file=sys.stderr)


def run_ours(args):
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    import probnum_b200 as pb  # noqa: PLC0415
    from probnum_b200 import _lib  # noqa: PLC0415
    from probnum_b200.sharded import ShardedCovarianceOperator  # noqa: PLC0415

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: probnum_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        _bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n, d = args.n, D_GRAM
    # synthetic inputs as in benchmarks/covfuncs.py:56 (default_rng(42), standard normal)
    x0_host = np.random.default_rng(42 + 1000 * rank).normal(size=(n, d))
    kernels = {
        "matern52": pb.Matern((d,), nu=2.5),
        "ratquad": pb.RatQuad((d,)),
        "expquad": pb.ExpQuad((d,)),
    }
    x0_dev = torch.from_numpy(x0_host).to(dev)
    ops = {}
    for name, k in kernels.items():
        ops[name] = _lib.Operator(k._kernel_params(), x0_dev.data_ptr(), None,
                                  device=local_rank, on_device=True, n0=n, n1=n, dtype=np.float64)
    out = torch.empty((n, n), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def gram(name):
        ops[name].gram_device(out.data_ptr(), n, stream=stream)

    step_kernels = ["matern52", "ratquad"]

    def step():
        for name in step_kernels:
            gram(name)

    sampler = ClockSampler(local_rank)
    sampler.start()  # samples every 100 ms through warm-up and the timed region
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _lib.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(step_kernels) + 1)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        ev[s][0].record()
        for i, name in enumerate(step_kernels):
            gram(name)
            ev[s][i + 1].record()
    barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][-1])
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    per_kernel_ms = {
        name: float(np.mean([ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(args.steps)]))
        for i, name in enumerate(step_kernels)
    }
    total_ms = max_over_ranks(total_ms)
    entries_per_step = 2.0 * n * n * world
    value = entries_per_step * args.steps / (total_ms * 1e-3)

    # roofline of the dominant launch (same kernel, slower epilogue)
    peaks, peak_kind = measured_peaks()
    dom = max(per_kernel_ms, key=per_kernel_ms.get)
    alg_bytes = 8.0 * n * n + 8.0 * d * (n + n)
    achieved = alg_bytes / (per_kernel_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tr = json.load(fh)
            traffic = tr.get("gram_sym_kernel_bytes_per_launch" if dom == "matern52" else "gram_fast_kernel_bytes_per_launch")
    roofline = {
        "bound": "hbm", "kernel": (f"gram_sym_kernel<D=3, {dom}> (upper triangle + mirrored stores)" if dom == "matern52"
                                    else f"gram_fast_kernel<D=3, {dom}>"),
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind}, copy read+write)",
        "algorithmic_bytes_per_launch": alg_bytes,
    }

    extra = {"per_kernel_ms": per_kernel_ms,
             "per_kernel_entries_per_s": {k: n * n / (v * 1e-3) for k, v in per_kernel_ms.items()}}

    # ExpQuad Gram alone (north_star's 70 % target is quoted on it)
    for _ in range(2):
        gram("expquad")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(reps):
        gram("expquad")
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
    extra["expquad_gram"] = {"ms": ms, "entries_per_s": n * n * world / (ms * 1e-3),
                             "hbm_gbs": alg_bytes / (ms * 1e-3) / 1e9,
                             "frac_of_peak": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}

    # ---- end to end through the public API: host arrays in, host ndarray out ------------------
    e2e = None
    if not args.no_e2e:
        ks = [kernels[name] for name in step_kernels]
        for k in ks:  # warm-up: pins the result buffers once
            k.matrix(x0_host)
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            for k in ks:
                K = k.matrix(x0_host)
                del K
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {
            "value": entries_per_step * e2e_steps / dt, "unit": UNIT,
            "h2d_bytes_per_step": 2 * 8 * d * n * world,
            "d2h_bytes_per_step": int(2 * 8 * n * n) * world,
            "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
            "api": "probnum_b200.covfuncs.{Matern,RatQuad}.matrix(X) -> numpy.ndarray (pinned result pool)",
        }
    del out
    torch.cuda.empty_cache()

    # ---- configs[2]: matrix-free ExpQuad K.v, N = 2^20, d = 8, row-sharded + all-gather -----------
    if not args.no_matvec:
        nm = args.n_mv
        xm = np.random.default_rng(42).normal(size=(nm, D_MV))
        kmv = pb.ExpQuad((D_MV,))
        if world > 1:
            sop = ShardedCovarianceOperator(kmv, xm, device=local_rank)
            local_op = sop.local
            run_product = lambda: sop.matmul_device(v, out=y)  # noqa: E731
        else:  # one GPU: the plain operator, no process group needed
            sop = None
            local_op = _lib.Operator(kmv._kernel_params(), np.ascontiguousarray(xm), None, device=local_rank)
            run_product = lambda: local_op.matmul_device(v.data_ptr(), 1, 1, y.data_ptr(), 1, stream=stream)  # noqa: E731
        v = torch.from_numpy(np.random.default_rng(7).normal(size=(nm, 1))).to(dev)
        y = torch.empty((nm, 1), dtype=torch.float64, device=dev)
        run_product()
        barrier()
        reps = 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(reps):
            run_product()
        e1.record()
        barrier()
        launches += _lib.launch_count() - l0
        ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        flop_per_pair = 3 * D_MV + 1 + 2  # 3d + 1 + 2K, FMA = 2
        path_op = sop.local_sym if (sop is not None and sop.local_sym is not None) else local_op
        symmetric = path_op.matmul_path().endswith("symmetric")
        extra["matvec"] = {
            "kernel": path_op.matmul_path(),
            "workload": f"ExpQuad K.v, N={nm}, d={D_MV}, K=1, float64, " + (
                ("symmetric halving (each off-diagonal tile evaluated once)" + (
                    "" if world == 1 else f", row blocks round-robin over {world} GPUs + all-reduce")) if symmetric
                else f"row-sharded over {world} GPU(s) + all-gather"),
            "ms": ms, "pairs_per_s": float(nm) * nm / (ms * 1e-3),
            "gflops": float(nm) * nm * flop_per_pair / (ms * 1e-3) / 1e9,
            "flop_per_pair": flop_per_pair, "scaling": "strong",
        }
        # FP64 roofline of this kernel: algorithmic flops against the DFMA rate measured on this
        # pool by tools/microbench.cu (profiles/r01_microbench.json), per GPU
        mb = os.path.join(ROOT, "profiles", "r01_microbench.json")
        if os.path.exists(mb):
            with open(mb) as fh:
                dfma = json.load(fh).get("dfma_tflops_sustained")
            if dfma:
                alg = extra["matvec"]["gflops"] / 1e3 / world
                # the symmetric kernel evaluates each off-diagonal 128 x 128 tile once (plus one more
                # FMA per pair for the mirrored sum): the roofline fraction is taken on the flops it
                # EXECUTES, the algorithmic figure (all N^2 pairs, BASELINE's metric) is kept beside it
                ach = alg * (0.5 * (1.0 + 128.0 / nm) * (flop_per_pair + 2) / flop_per_pair) if symmetric else alg
                extra["matvec"]["roofline"] = {"bound": "fp64 pipe", "achieved": ach, "algorithmic": alg, "peak": dfma,
                                               "unit": "TFLOP/s", "frac": ach / dfma,
                                               "peak_source": "profiles/r01_microbench.json dfma_tflops_sustained"}
        if sop is not None:
            sop.close()
        else:
            local_op.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(n)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"Matern-5/2 + RatQuad Gram, N={n}, d={d}, float64 (BASELINE configs[1])"
                            + ("" if world == 1 else f"; {world} independent point clouds, one per rank/GPU"),
                "l2": "each launch writes 8.6 GB >> 126 MB L2; no flush needed",
                "inputs": "default_rng(42).normal, resident in HBM",
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "extra": extra,
        }
        emit(line)
    if dist.is_initialized():
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict) -> None:
    """Print the one JSON line on the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # NCCL (and anything else native) may write banners to fd 1: point fd 1 at stderr for the
    # duration of the run and keep the original stdout for the one JSON line
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
