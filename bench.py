#!/usr/bin/env python
"""bench.py — C-layer construction throughput on B200 (BASELINE.json metric: C-layers/sec, Minkowski sum +
sweep-line free segments), synthetic scaling workload of BASELINE.json configs[4] / SURVEY.md §8(d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (K1 Minkowski + K2 sweep + K3 free segments) over a batch of
`--layers` orientations per rank.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torch initialises CUDA before libhrmgpu.so is loaded: ask for eager module loading here (see hrm_b200/csrc/ctx.cu)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

from hrm_b200 import workload as wl  # noqa: E402

METRIC = "C-layers/sec (Minkowski sum + sweep-line free segments)"
UNIT = "layers/s"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(layers, n_obs=wl.N_OBS):
    return {
        "workload": "synthetic-scaling (BASELINE.json configs[4]): 1 arena + %d superquadric obstacles x n=%d "
                    "(10^4 boundary pts) x %dx%d sweep lines, 1 ellipsoid body, per-orientation poses" % (n_obs, wl.N_GRID, wl.LX, wl.LY),
        "layers_per_step_per_gpu": layers, "surfaces_per_layer": n_obs + 1, "points_per_surface": wl.N_GRID ** 2,
        "sweep_lines": wl.LX * wl.LY, "seed": wl.SEED,
        "l2_policy": "outputs (>=248 MB per layer, fresh addresses per layer) far exceed the 126 MB L2; inputs are KB-sized tables",
    }


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (the reference itself cannot be compiled here: no Eigen)
# ------------------------------------------------------------------------------------------------
def cpu_layers_per_s(n_obs_sample, layers, threads):
    from oracle import hrmo

    rows = wl.scene_rows12(wl.N_OBS)[: 1 + n_obs_sample]
    q = wl.orientations(layers, seed=wl.SEED + 7).reshape(layers, 1, 4)
    off = np.zeros((layers, 1, 3))
    sec, _ = hrmo.time_layers3d(1, n_obs_sample, rows, wl.N_GRID, np.array(wl.BODY_SEMI), q, off, wl.LIM, wl.LX, wl.LY, threads=threads)
    frac = (1 + n_obs_sample) / float(1 + wl.N_OBS)  # share of one full layer's surfaces in the sample
    return layers * frac / sec, sec


def plan_ms_block():
    """Secondary metric of BASELINE.json ('end-to-end plan ms'): HRM3D::plan() through the C++ host layer + GPU kernels next to
    the CPU oracle planner on one core, on the bundled cluttered map (configs[1]: 1+7 superquadrics, rabbit robot,
    q_icosahedron_60, n=10, 11x5 lines) and on the maze map with dense sweep lines (configs[2]: 44x20 lines).  The wall clock
    of the GPU arm includes creating the context and uploading the scene.  Roadmaps are identical (tests/test_gpu_planner.py)."""
    import json as _json

    from hrm_b200 import hostlib
    from oracle import hrmo

    scenes = _json.load(open(os.path.join(ROOT, "tests", "golden", "scenes.json")))

    def one(name, Lx, Ly, label):
        sc = scenes[name]
        rows = np.array(sc["arena"] + sc["obstacle"])
        a = (1, len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], Lx or sc["Lx"], Ly or sc["Ly"], 5,
             sc["end_points"][0], sc["end_points"][1])
        t = []
        for _ in range(4):
            t0 = time.perf_counter()
            got = hostlib.plan3d(*a, per_orientation=True, precision=0)
            t.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        ref = hrmo.plan3d(*a, per_orientation=True)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        return {"scene": label, "gpu_plan_ms": sorted(t[1:])[len(t[1:]) // 2], "gpu_plan_ms_planner_clock": got["total_time_s"] * 1e3,
                "cpu_oracle_plan_ms": cpu_ms, "cpu_cores": 1, "solved": bool(got["solved"]), "vertices": int(len(got["vertex"])),
                "edges": int(len(got["edge"])), "edge_tests": int(got["edge_tests"]), "gpu_launches": int(got["gpu_launches"]),
                "identical_roadmap": bool(np.array_equal(got["vertex"], ref["vertex"]) and np.array_equal(got["edge"], ref["edge"]))}

    def timed(gpu_call, cpu_call, label, extra=None):
        t = []
        for _ in range(4):
            t0 = time.perf_counter()
            got = gpu_call()
            t.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        ref = cpu_call()
        cpu_ms = (time.perf_counter() - t0) * 1e3
        r = {"scene": label, "gpu_plan_ms": sorted(t[1:])[len(t[1:]) // 2], "gpu_plan_ms_planner_clock": got["total_time_s"] * 1e3,
             "cpu_oracle_plan_ms": cpu_ms, "cpu_cores": 1, "solved": bool(got["solved"]), "vertices": int(len(got["vertex"])),
             "edges": int(len(got["edge"])), "gpu_launches": int(got["gpu_launches"]),
             "identical_roadmap": bool(np.array_equal(got["vertex"], ref["vertex"]) and np.array_equal(got["edge"], ref["edge"]))}
        if extra:
            r.update(extra(got))
        return r

    def hrm2d():
        # configs[0]: TestHRM2D's scene (sparse map, rabbit robot, 50 points per curve) with the README bench's 20 slices x 30 lines
        sc = scenes["2D_superellipse_sparse_rabbit"]
        rows = np.array(sc["arena"] + sc["obstacle"])
        a = (len(sc["arena"]), len(sc["obstacle"]), rows, 50, np.array(sc["robot"]), 20, sc["lim"], 30, 5, sc["end_points"][0], sc["end_points"][1])
        return timed(lambda: hostlib.plan2d(*a, per_orientation=False, precision=0), lambda: hrmo.plan2d(*a, per_orientation=False),
                     "HRM2D sparse/rabbit (BASELINE.json configs[0]), 20 C-layers, 50 points per curve, 30 sweep lines, F64_STRICT")

    def prob3d():
        # configs[3]: ProbHRM3D on the home map with the snake robot (URDF chain, TFE bridge layers); the random draws are a fixed
        # Halton sequence injected into both planners (the reference seeds from the clock)
        from tests.test_oracle_kat import GOLDEN, prob_samples

        sc = scenes["3D_superquadrics_home_snake"]
        rows = np.array(sc["arena"] + sc["obstacle"])
        a = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), os.path.join(GOLDEN, "snake.urdf"), prob_samples(30, 3),
             sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1])
        return timed(lambda: hostlib.plan_prob3d(*a, per_orientation=True, precision=0, batch=8), lambda: hrmo.plan_prob3d(*a, per_orientation=True),
                     "ProbHRM3D home/snake (BASELINE.json configs[3]), n=10, %dx%d lines, injected draws, F64_STRICT" % (sc["Lx"], sc["Ly"]),
                     lambda got: {"slices": int(got["slices"])})

    out = one("3D_superquadrics_cluttered_rabbit", 0, 0,
              "HRM3D cluttered/rabbit (BASELINE.json configs[1]), 60 C-layers, n=10, 11x5 lines, per-orientation, F64_STRICT")
    out["dense"] = one("3D_superquadrics_maze_rabbit", 44, 20,
                       "HRM3D maze/rabbit with dense sweep lines (BASELINE.json configs[2]), 60 C-layers, n=10, 44x20 lines, per-orientation, F64_STRICT")
    out["hrm2d"] = hrm2d()
    out["prob3d"] = prob3d()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    n_obs_sample = 124  # arena + 124 obstacles = 125 of the 1001 surfaces of a layer
    for _ in range(args.warmup):
        cpu_layers_per_s(n_obs_sample, threads, threads)
    t = []
    v = []
    for _ in range(args.steps):
        val, sec = cpu_layers_per_s(n_obs_sample, threads, threads)
        v.append(val)
        t.append(sec)
    value = len(v) / sum(1.0 / x for x in v)  # total layer-equivalents / total time
    sample = ("per step: %d threads x 1 layer each over the arena + first %d obstacles (125/1001 of a layer's surfaces), "
              "scaled to full-layer equivalents; oracle = CPU restatement of the reference (Eigen absent: reference not buildable)"
              % (threads, n_obs_sample))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / len(t), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.layers),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, every ~5 ms; nvidia-smi fallback)."""
    REASONS = {0x08: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x04: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            self.sm.append(int(out[0]))
            self.max_mhz = int(out[1])

    def _run(self):
        while not self.stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self.stop.wait(0.005)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=10)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from hrm_b200 import capi
    from hrm_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    prec = {"f64": capi.F64, "f64_strict": capi.F64_STRICT, "f32": capi.F32}[args.precision]
    layers = args.layers
    n_obs = args.obstacles
    rows12 = wl.scene_rows12(wl.N_OBS)[: 1 + n_obs]
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_scene3d(1, n_obs, wl.scene_rows17(rows12), wl.N_GRID)
    ctx.set_sweep(wl.LIM, wl.LX, wl.LY)

    total_steps = args.warmup + args.steps
    # every step (and every rank) sees different orientations: nothing can be cached across steps
    quats = wl.orientations(total_steps * layers, seed=wl.SEED + 1 + 1000 * rank)
    R_all = wl.body_rotations(quats)                                   # [T*layers][1][9]
    off_all = np.zeros((total_steps * layers, 1, 3))
    semi_h = np.array(wl.BODY_SEMI)
    d_semi = torch.tensor(semi_h, dtype=torch.float64, device=dev)
    d_R = torch.tensor(R_all, dtype=torch.float64, device=dev)
    d_off = torch.tensor(off_all, dtype=torch.float64, device=dev)

    def step_resident(i):
        a = i * layers
        ctx.build_layers_dev(layers, 1, d_semi.data_ptr(), d_R[a:a + layers].data_ptr(), d_off[a:a + layers].data_ptr(), prec)
        if world > 1:
            hdist.allgather_segments(ctx, dist, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    barrier()
    ctx.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for i in range(args.warmup, total_steps):
            step_resident(i)
        e1.record(stream)
        barrier()
    launches = ctx.kernel_launches()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * layers * args.steps / (ms * 1e-3)

    # dominant-kernel duration: average over a few extra launches, CUDA events recorded by the library on
    # the launching stream right around the fused Minkowski+sweep kernel
    kt = []
    for i in range(min(args.steps, 5)):
        step_resident(args.warmup + i)
        kt.append(ctx.last_minksweep_ms())
    kernel_ms = sum(kt) / len(kt)
    d = ctx.dims()
    w = 4 if prec == capi.F32 else 8
    alg_bytes = wl.algorithmic_bytes_per_layer(d["S"], d["M"], d["L"], w) * layers
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_n%d_S%d_L%d" % (args.precision, wl.N_GRID, d["S"], d["L"])
        if key in tr:
            traffic = tr[key]["dram_bytes_per_layer"] * layers
    except Exception:
        pass

    # end to end through the C ABI with HOST buffers: H2D of the step's body poses, D2H of its free segments
    e2e_steps = max(1, min(args.steps, 5))
    pin_R = torch.tensor(R_all).pin_memory()
    pin_off = torch.tensor(off_all).pin_memory()
    R_np, off_np = pin_R.numpy(), pin_off.numpy()
    h2d = layers * 1 * (9 + 3) * 8 + semi_h.size * 8
    d2h = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        a = (args.warmup + i) * layers
        ctx.build_layers(semi_h, R_np[a:a + layers], off_np[a:a + layers], prec)
        off, xL, xU, xM = ctx.segments_all()
        d2h = off.nbytes + xL.nbytes + xU.nbytes + xM.nbytes
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * layers * e2e_steps / (float(t.item()) * 1e-3)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            v, sec = cpu_layers_per_s(wl.N_OBS if n_obs == wl.N_OBS else n_obs, threads, threads)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "%d full layers (1 per host thread, %d threads) of the same workload, %.1f s; oracle = CPU restatement "
                             "of the reference (reference not buildable here: Eigen absent)" % (threads, threads, sec)}
            # the reference itself is single-threaded: the same oracle on ONE core, bounded sample
            v1, sec1 = cpu_layers_per_s(124, 1, 1)
            cpu["single_core"] = {"value": v1, "unit": UNIT, "cores": 1,
                                  "sample": "1 layer over the arena + first 124 obstacles (125/1001 of a layer's surfaces), %.1f s, scaled "
                                            "to full-layer equivalents" % sec1}
        plan = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                plan = plan_ms_block()
            except Exception as e:  # the headline metric must not depend on the secondary one
                plan = {"error": str(e)}
        cfg = workload_config(layers, n_obs)
        cfg["precision_mode"] = args.precision
        cfg["exchange"] = "all-gather of free segments per step (NCCL)" if world > 1 else "none (1 GPU)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if prec == capi.F32 else "f64", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_minksweep3d", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks.summary(), "plan": plan,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if line is not None:
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=128, help="orientations (C-layers) per step per GPU")
    ap.add_argument("--obstacles", type=int, default=wl.N_OBS)
    ap.add_argument("--precision", default="f64", choices=["f64", "f64_strict", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
