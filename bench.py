#!/usr/bin/env python
"""bench.py — C-layer construction throughput on B200 (BASELINE.json metric: C-layers/sec, Minkowski sum +
sweep-line free segments), synthetic scaling workload of BASELINE.json configs[4] / SURVEY.md §8(d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (K1 Minkowski + K2 sweep + K3 free segments) over `--builds-per-step` batches of
`--layers` orientations per rank (defaults 4 x 125: with the driver's 20 steps one run covers the 10 000 orientations of
configs[4] per GPU).  Prints ONE JSON line (rank 0): the headline on the SURVEY §8(d) scene, plus `parity` (the timed mode
against the bit-exact mode and the CPU oracle, outside the timed region), `freespace` (the same measurement on a scene whose
layers really have free segments), `strong_scaling` (N > 1: one pass over the fixed 10 000 orientations) and `plan`.
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


def workload_config(layers, builds_per_step, n_obs=wl.N_OBS):
    """The SAME dict for the GPU arm and the reference arm (what each arm samples of it is said in its own keys)."""
    return {
        "workload": "synthetic-scaling (BASELINE.json configs[4]): 1 arena + %d superquadric obstacles x n=%d "
                    "(10^4 boundary pts) x %dx%d sweep lines, 1 ellipsoid body, per-orientation poses" % (n_obs, wl.N_GRID, wl.LX, wl.LY),
        "layers_per_step_per_gpu": layers * builds_per_step, "layers_per_build": layers, "builds_per_step": builds_per_step,
        "surfaces_per_layer": n_obs + 1, "points_per_surface": wl.N_GRID ** 2,
        "sweep_lines": wl.LX * wl.LY, "seed": wl.SEED,
        "l2_policy": "outputs (>=248 MB per layer, fresh addresses per layer) far exceed the 126 MB L2; inputs are KB-sized tables",
    }


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (the reference itself cannot be compiled here: no Eigen)
# ------------------------------------------------------------------------------------------------
def cpu_layers_per_s(n_obs_sample, layers, threads, fast=False):
    from oracle import hrmo

    rows = wl.scene_rows12(wl.N_OBS)[: 1 + n_obs_sample]
    q = wl.orientations(layers, seed=wl.SEED + 7).reshape(layers, 1, 4)
    off = np.zeros((layers, 1, 3))
    sec, _ = hrmo.time_layers3d(1, n_obs_sample, rows, wl.N_GRID, np.array(wl.BODY_SEMI), q, off, wl.LIM, wl.LX, wl.LY, threads=threads, fast=fast)
    frac = (1 + n_obs_sample) / float(1 + wl.N_OBS)  # share of one full layer's surfaces in the sample
    return layers * frac / sec, sec


def cpu_baseline_block(n_obs):
    threads = host_threads()
    v, sec = cpu_layers_per_s(wl.N_OBS if n_obs == wl.N_OBS else n_obs, threads, threads)
    cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d full layers (1 per host thread, %d threads) of the same workload, %.1f s; oracle = CPU restatement "
                     "of the reference, op-count faithful (reference not buildable here: Eigen absent)" % (threads, threads, sec)}
    # the reference itself is single-threaded: the same oracle on ONE core, bounded sample
    v1, sec1 = cpu_layers_per_s(124, 1, 1)
    cpu["single_core"] = {"value": v1, "unit": UNIT, "cores": 1,
                          "sample": "1 layer over the arena + first 124 obstacles (125/1001 of a layer's surfaces), %.1f s, scaled "
                                    "to full-layer equivalents" % sec1}
    # SURVEY.md §8(d) asks for both builds: `faithful` above keeps the reference's operation counts, `fast` hoists the loop-invariant
    # work (power function on n instead of n^2 values, mesh bounding boxes once per mesh); results are bit-identical
    vf, secf = cpu_layers_per_s(wl.N_OBS if n_obs == wl.N_OBS else n_obs, threads, threads, fast=True)
    cpu["fast"] = {"value": vf, "unit": UNIT, "cores": threads, "kind": "port (hoisted build, -DHRMO_FAST)",
                   "sample": "%d full layers (1 per host thread), %.1f s" % (threads, secf)}
    cpu["build"] = "faithful"
    return cpu


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
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.layers, args.builds_per_step),
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
def parity_block(ctx, capi, rows12, workload, layers, timed_prec, rank_seed):
    """Outside every timed region: the SAME `layers` orientations built in the bit-exact mode (HRMGPU_F64_STRICT, identical to
    the oracle: tests/test_gpu_fullsize.py) and in the timed mode, compared pair by pair -- hit / no-hit decisions of every
    (line, surface), interval and segment end points -- and layer 0 of the batch against the CPU oracle itself."""
    from oracle import hrmo  # the checker (test infrastructure), never inside a timed region

    quats = wl.orientations(layers, seed=wl.SEED + 5000 + rank_seed)
    R = wl.body_rotations(quats)
    off = np.zeros((layers, 1, 3))
    semi = np.array(wl.BODY_SEMI)
    out = {}

    def collect(prec):
        ctx.build_layers(semi, R, off, prec)
        seg = ctx.segments_all()
        iv = [ctx.intervals(k) for k in range(layers)]
        return seg, iv

    seg_s, iv_s = collect(capi.F64_STRICT)
    seg_t, iv_t = collect(timed_prec)
    flips, dmax, pairs = 0, 0.0, 0
    for (lo0, up0), (lo1, up1) in zip(iv_s, iv_t):
        n0, n1 = np.isnan(lo0), np.isnan(lo1)
        flips += int((n0 != n1).sum())
        pairs += lo0.size
        both = ~(n0 | n1)
        if both.any():
            dmax = max(dmax, float(np.max(np.abs(lo0[both] - lo1[both]))), float(np.max(np.abs(up0[both] - up1[both]))))
    same_structure = bool(np.array_equal(seg_s[0], seg_t[0]))
    seg_diff = max((float(np.max(np.abs(a - b), initial=0.0)) for a, b in zip(seg_s[1:], seg_t[1:])), default=0.0) if same_structure else None
    per_line = np.diff(seg_s[0])
    out.update({"layers": layers, "line_surface_pairs": pairs, "hit_decision_flips": flips, "max_interval_abs_diff": dmax,
                "same_segment_structure": same_structure, "max_segment_abs_diff": seg_diff,
                "segments_per_layer": float(seg_s[0][-1]) / layers, "lines_with_segments_frac": float((per_line > 0).mean()),
                "reference_mode": "HRMGPU_F64_STRICT (bit-identical to the CPU oracle)", "timed_mode": workload["precision"]})
    # layer 0 of the same batch against the oracle (all host threads over the surfaces)
    ref = hrmo.layer3d_parallel(1, len(rows12) - 1, rows12, wl.N_GRID, semi, quats[0].reshape(1, 4), off[0], wl.LIM, wl.LX, wl.LY)
    L = wl.LX * wl.LY
    o_s = seg_s[0][: L + 1]
    strict_ok = (np.array_equal(iv_s[0][0], ref["lo"], equal_nan=True) and np.array_equal(iv_s[0][1], ref["up"], equal_nan=True)
                 and np.array_equal(o_s, ref["off"]) and np.array_equal(seg_s[3][: o_s[-1]], ref["xM"]))
    lo1, up1 = iv_t[0]
    out["oracle_layer0"] = {"strict_bit_exact": bool(strict_ok), "timed_mode_flips": int((np.isnan(lo1) != np.isnan(ref["lo"])).sum()),
                            "timed_mode_max_interval_abs_diff": float(np.nanmax(np.abs(lo1 - ref["lo"]))) if (~np.isnan(ref["lo"])).any() else 0.0}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from hrm_b200 import capi
    from hrm_b200 import dist as hdist

    t_start = time.perf_counter()
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
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
    layers, bps, n_obs = args.layers, args.builds_per_step, args.obstacles
    ctx = capi.Context(local)
    # a real (non-default) stream: the library queues everything on it and returns, and the CUDA events below are recorded on it
    # (the legacy default stream has handle 0, which hrmgpu_set_stream reads as "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        hdist.init_exchange(ctx, dist, dev)  # NCCL communicator of the library; torch.distributed only broadcasts the id
    semi_h = np.array(wl.BODY_SEMI)
    d_semi = torch.tensor(semi_h, dtype=torch.float64, device=dev)
    L = wl.LX * wl.LY
    line_cap = layers * L

    def stage(msg):  # progress on stderr (HRM_BENCH_VERBOSE=1): where a multi-rank run is, should it ever stall
        if os.environ.get("HRM_BENCH_VERBOSE"):
            print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(workload, steps, warmup, seed_base):
        """One workload: device-resident throughput, dominant-kernel time, end-to-end through host buffers."""
        rows12 = wl.scene_rows12(wl.N_OBS, workload=workload)[: 1 + n_obs]
        ctx.set_scene3d(1, n_obs, wl.scene_rows17(rows12), wl.N_GRID)
        ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
        builds = (warmup + steps) * bps
        stage("measure(%s): setup" % workload)
        # every build (and every rank) sees different orientations: nothing can be cached across builds
        quats = wl.orientations(builds * layers, seed=wl.SEED + seed_base + 1000 * rank)
        R_all = wl.body_rotations(quats)                                   # [builds*layers][1][9]
        off_all = np.zeros((builds * layers, 1, 3))
        d_R = torch.tensor(R_all, dtype=torch.float64, device=dev)
        d_off = torch.tensor(off_all, dtype=torch.float64, device=dev)
        xcap = [4096]  # exchange / read-back capacity in segments (same on all ranks), fixed after the first warm-up build

        def build_resident(i):
            a = i * layers
            ctx.build_layers_dev(layers, 1, d_semi.data_ptr(), d_R[a:a + layers].data_ptr(), d_off[a:a + layers].data_ptr(), prec)
            if world > 1:
                ctx.allgather_layers(line_cap, xcap[0])

        # first build: learn the segment volume (grows the library's capacities once), agree on the exchange capacity
        build_resident(0)
        tot = ctx.dims()["segments"]
        xcap[0] = int(max_over_ranks(float(tot)) * 1.5) + 4096
        stage("first build done, capacity %d" % xcap[0])
        for i in range(1, warmup * bps):
            build_resident(i)
        ctx.join()
        barrier()
        stage("warm-up done")
        ctx.kernel_launches(reset=True)
        ctx.segment_redos(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            barrier()
            e0.record(stream)
            for i in range(warmup * bps, builds):
                build_resident(i)
            ctx.join()  # the last build's free-segment passes / exchange (side stream) are inside the timed region
            e1.record(stream)
            barrier()
        launches = ctx.kernel_launches()
        ms = max_over_ranks(e0.elapsed_time(e1))
        stage("timed region done: %.2f ms" % ms)
        res = {"value": world * layers * bps * steps / (ms * 1e-3), "ms_per_step": ms / steps, "launches": launches, "clocks": clocks.summary(),
               "segment_pass_redos_in_timed_region": ctx.segment_redos()}

        # dominant-kernel duration: CUDA events recorded by the library on the launching stream right around the fused
        # Minkowski+sweep kernel; the build before it is still running its free-segment passes (as in the timed loop)
        kt, st = [], []
        for j in range(min(steps * bps, 6)):
            build_resident(warmup * bps + j)
            build_resident(warmup * bps + j + 1 if warmup * bps + j + 1 < builds else warmup * bps)
            kt.append(ctx.last_minksweep_ms())
            st.append(ctx.last_segments_ms())
        res["kernel_ms"] = sum(kt) / len(kt)
        res["segments_ms"] = sum(st) / len(st)
        ki = []
        for j in range(3):  # the same kernel with nothing else on the GPU (no free-segment passes of a previous build beside it)
            ctx.synchronize()
            build_resident(warmup * bps + j)
            ki.append(ctx.last_minksweep_ms())
        res["kernel_ms_alone"] = sum(ki) / len(ki)
        d = ctx.dims()
        res["dims"] = d
        res["segments_per_layer"] = d["segments"] / float(layers)

        # end to end through the C ABI with (pinned) HOST buffers, pipelined the way a planner would drive it: H2D of the
        # build's body poses, the build, [the exchange], D2H of its free segments into one of two host buffers while the next
        # build already runs; every build's result is waited for and touched on the host
        stage("kernel timings done")
        e2e_builds = max(2, min(steps * bps, 12))
        pin_R = torch.tensor(R_all).pin_memory()
        pin_off = torch.tensor(off_all).pin_memory()
        R_np, off_np = pin_R.numpy(), pin_off.numpy()
        cap = xcap[0]
        hb = [[torch.zeros(line_cap + 1, dtype=torch.int64).pin_memory().numpy()] + [torch.zeros(cap, dtype=torch.float64).pin_memory().numpy() for _ in range(3)]
              for _ in range(2)]
        h2d = layers * 1 * (9 + 3) * 8 + semi_h.size * 8
        d2h = 0
        chk = 0.0
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_builds):
            a = (warmup * bps + i) * layers
            ctx.build_layers(semi_h, R_np[a:a + layers], off_np[a:a + layers], prec)
            if world > 1:
                ctx.allgather_layers(line_cap, xcap[0])
            if i > 0:
                n = ctx.wait_segments()          # result of build i-1 is on the host now
                chk += float(hb[(i - 1) % 2][3][:n].sum()) + float(hb[(i - 1) % 2][0][-1])
            copied = ctx.fetch_segments_async(*hb[i % 2])
            d2h = hb[i % 2][0].nbytes + 3 * 8 * copied
        n = ctx.wait_segments()
        chk += float(hb[(e2e_builds - 1) % 2][3][:n].sum())
        if world > 1:
            ctx.synchronize()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        res["e2e"] = {"value": world * layers * e2e_builds / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * bps,
                      "d2h_bytes_per_step": d2h * bps, "builds": e2e_builds,
                      "api": "hrmgpu_build_layers (host poses) -> hrmgpu_fetch_segments_async / hrmgpu_wait_segments (pinned host arrays)"
                             + (" with hrmgpu_allgather_layers per build" if world > 1 else ""),
                      "host_checksum": chk}
        res["allgather_bytes_per_build_per_rank"] = (8 * (2 + line_cap + 1) + 24 * xcap[0]) if world > 1 else 0
        res["rows12"] = rows12
        stage("e2e done")
        return res

    head = measure("scaling", args.steps, args.warmup, 1)
    value, ms_step, kernel_ms = head["value"], head["ms_per_step"], head["kernel_ms"]
    d = head["dims"]
    w = 4 if prec == capi.F32 else 8
    alg_bytes = wl.algorithmic_bytes_per_layer(d["S"], d["M"], d["L"], w) * layers
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_n%d_S%d_L%d" % (args.precision, wl.N_GRID, d["S"], d["L"])
        if key in tr:
            traffic = tr[key]["dram_bytes_per_layer"] * layers
            traffic_src = "static: " + tr[key].get("source", "profiles/traffic.json") + " (one ncu --set full capture, per layer x layers per launch; not re-measured in this run)"
    except Exception:
        pass
    parity = None
    if world == 1 and not args.no_parity:
        parity = parity_block(ctx, capi, head["rows12"], {"precision": args.precision}, layers, prec, 0)

    # second workload: the same generator with small obstacles, so that the layers really have free space (K3, D2H and the
    # exchange carry a payload); reported beside the headline, never as the headline
    free = None
    if not args.no_freespace:
        fr = measure("freespace", max(2, args.steps // 4), 3, 501)
        free = {"workload": "synthetic-freespace: same generator / seed / sizes as the headline, obstacle semi-axes U[0.5,2] instead of U[2,10]",
                "value": fr["value"], "unit": UNIT, "ms_per_step": fr["ms_per_step"], "kernel_ms": fr["kernel_ms"], "segments_ms": fr["segments_ms"],
                "segments_share_of_build": fr["segments_ms"] / (fr["ms_per_step"] / bps), "segments_per_layer": fr["segments_per_layer"],
                "roofline_frac": alg_bytes / (fr["kernel_ms"] * 1e-3) / 1e9 / peak, "e2e": fr["e2e"],
                "allgather_bytes_per_build_per_rank": fr["allgather_bytes_per_build_per_rank"],
                "segment_pass_redos_in_timed_region": fr["segment_pass_redos_in_timed_region"]}
        if world == 1 and not args.no_parity:
            free["parity"] = parity_block(ctx, capi, fr["rows12"], {"precision": args.precision}, min(layers, 32), prec, 7)

    # strong scaling on the fixed K = 10 000 orientations of BASELINE.json configs[4]: every rank builds its shard in batches
    strong = None
    if world > 1 and not args.no_strong:
        rows12 = head["rows12"]
        ctx.set_scene3d(1, n_obs, wl.scene_rows17(rows12), wl.N_GRID)
        ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
        K = 10000
        k0, k1 = hdist.shard_range(K, rank, world)
        quats = wl.orientations(K, seed=wl.SEED + 77)[k0:k1]
        dR = torch.tensor(wl.body_rotations(quats), dtype=torch.float64, device=dev)
        dO = torch.zeros((k1 - k0, 1, 3), dtype=torch.float64, device=dev)
        # the exchange capacity is part of the collective's size: it has to be the SAME number on every rank
        xc = int(max_over_ranks(head["segments_per_layer"]) * layers * 1.5) + 4096

        def one_pass():
            for a in range(0, k1 - k0, layers):
                nb = min(layers, k1 - k0 - a)
                ctx.build_layers_dev(nb, 1, d_semi.data_ptr(), dR[a:a + nb].data_ptr(), dO[a:a + nb].data_ptr(), prec)
                ctx.allgather_layers(line_cap, xc)
            ctx.join()

        stage("strong scaling: capacity %d" % xc)
        one_pass()
        barrier()
        stage("strong scaling: warm-up pass done")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_pass()
        e1.record(stream)
        barrier()
        sms = max_over_ranks(e0.elapsed_time(e1))
        strong = {"orientations": K, "per_gpu": k1 - k0, "ms": sms, "value": K / (sms * 1e-3), "unit": UNIT, "scaling": "strong",
                  "note": "one pass over BASELINE.json configs[4]'s 10 000 orientations, sharded by orientation, exchange per batch included"}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_block(n_obs)
        plan = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                plan = plan_ms_block()
            except Exception as e:  # the headline metric must not depend on the secondary one
                plan = {"error": str(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if prec == capi.F32 else "f64", "data": "synthetic", "config": workload_config(layers, bps, n_obs),
            "precision_mode": args.precision,
            "exchange": "hrmgpu_allgather_layers (one ncclAllGather of a fixed-capacity packed block per build, side stream)" if world > 1 else "none (1 GPU)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_minksweep3d", "kernel_ms": kernel_ms,
                         "launches_per_step": bps, "algorithmic_bytes_per_launch": alg_bytes,
                         "segments_ms_per_launch": head["segments_ms"], "kernel_ms_alone": head["kernel_ms_alone"], "step_minus_kernels_ms": ms_step - bps * kernel_ms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "cpu_baseline": cpu, "e2e": head["e2e"], "gpu_launches": head["launches"], "clocks": head["clocks"],
            "segments_per_layer": head["segments_per_layer"], "segment_pass_redos_in_timed_region": head["segment_pass_redos_in_timed_region"],
            "allgather_bytes_per_build_per_rank": head["allgather_bytes_per_build_per_rank"],
            # distinct orientations built inside the timed region, all ranks (defaults: 20 steps x 500 = the 10 000 orientations of configs[4] per GPU)
            "orientations_in_timed_region": args.steps * bps * layers * world,
            "parity": parity, "freespace": free, "strong_scaling": strong, "plan": plan,
        }
    if world > 1:
        dist.barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    if line is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=125, help="orientations (C-layers) per build call per GPU")
    ap.add_argument("--builds-per-step", type=int, default=4,
                    help="build calls per step: 20 steps x 4 x 125 layers = the 10 000 orientations of BASELINE.json configs[4] per GPU")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-freespace", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--obstacles", type=int, default=wl.N_OBS)
    ap.add_argument("--precision", default="f64", choices=["f64", "f64_strict", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--watchdog", type=int, default=int(os.environ.get("HRM_BENCH_WATCHDOG", "1500")),
                    help="seconds after which a stuck run dumps every thread's Python stack to stderr and exits (0 = off)")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler

        faulthandler.dump_traceback_later(args.watchdog, exit=True)
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
