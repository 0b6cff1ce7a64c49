#!/usr/bin/env python
"""bench.py — CFR subgame-iters/sec on the BASELINE.json workload.

One "step" = one wave: K concurrent 1x6f depth-2 root subgames (last_bid = -1, random beliefs) each solved with 1024 CFR
iterations, the leaf value net (Net2 256x2 + LayerNorm, random init seed 0) evaluated on every iteration — the reference's
`build_solver` + `multistep` + `update_value_network` for every subgame of the wave (subgame_solving.cc:791,666,672).

    python bench.py --gpus 1 --steps 3 --warmup 3                      # this repo's CUDA path (1 GPU)
    torchrun --nproc-per-node N ... bench.py --gpus N ...              # N GPUs, K subgames per GPU (weak scaling)
    python bench.py --impl reference ...                               # the reference's own CPU path on the host cores

Prints ONE JSON line (rank 0).  `value` = device-resident wave solve (state re-initialised on the device each step);
`e2e` = the same through the host-buffer C-ABI calls (cfrb_begin_wave H2D + cfrb_run + cfrb_examples D2H, + NCCL gather
of the examples to rank 0 when N > 1) timed end to end.
"""
import argparse
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

METRIC = "cfr_subgame_iters_per_sec"
UNIT = "subgame-iters/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--subgames", type=int, default=8192, help="concurrent subgames per GPU")
    ap.add_argument("--iters", type=int, default=1024)
    ap.add_argument("--dice", type=int, default=1)
    ap.add_argument("--faces", type=int, default=6)
    ap.add_argument("--net", default="auto", choices=["auto", "fp32", "tc", "tcx2"],
                    help="value-net kernel: tcx2 (default) = tcgen05 fp16 with packed-half GELU, tc = tcgen05 fp16 with fp32 GELU, fp32 = SIMT parity net")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="subgames in the CPU-baseline sample (0 = auto)")
    return ap.parse_args()


def dims(D, F):
    A = 1 + 2 * D * F
    H = F ** D
    return A, H, 2 + A + 2 * H


def workload_beliefs(n, H, first):
    """Synthetic random beliefs b_p[h] = u / sum(u), u ~ U(0,1); subgame g of the whole job uses counter-based stream g."""
    out = np.empty((n, 2, H), np.float64)
    for i in range(n):
        out[i] = np.random.Generator(np.random.Philox(key=first + i)).random((2, H))
    out /= out.sum(-1, keepdims=True)
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (nvidia_ml_py): polling the
    nvidia-smi CLI at 5 Hz was measured to slow the timed loop by ~9 % (driver-side query cost), NVML calls do not."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=float(os.environ.get("BENCH_SAMPLER_PERIOD", "0.1"))):
        self.index, self.period, self.rows, self.stop_flag, self.t, self.nv, self.h, self.max_mhz = index, period, [], False, None, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((mhz, reasons, time.perf_counter()))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def window(self, t0, t1):
        """Keep only the samples taken inside the timed region [t0, t1] (perf_counter)."""
        self.rows = [r for r in self.rows if t0 <= r[2] <= t1]

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        sm = [r[0] for r in self.rows]
        active = sorted(name for name, bit in self.REASONS.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": active, "samples": len(sm), "source": f"nvml, every {self.period:g} s during the timed steps"}


def effective_cores():
    """Host cores this process may actually use: min(cpu_count, affinity mask, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return n


_SCRIPT = {}


def script_net_file(D, F):
    if (D, F) in _SCRIPT:
        return _SCRIPT[(D, F)]
    _SCRIPT[(D, F)] = _script_net_file(D, F)
    return _SCRIPT[(D, F)]


def _script_net_file(D, F):
    import torch
    from rebel_b200.models import make_selfplay_net
    path = os.path.join(tempfile.mkdtemp(prefix="cfrb_bench_"), "net2.torchscript")
    torch.jit.script(make_selfplay_net(D, F, seed=0)).save(path)
    return path


def cpu_reference_rate(D, F, iters, n, beliefs, threads=None):
    """The reference's CPU implementation of the same workload sample (oracle/_ref when the reference compiled, else the
    C port) on the host cores.  Returns (rate, info)."""
    from oracle.oracle import Oracle, available
    cores = effective_cores()
    if available("ref_fast"):
        # all the host threads it can use: one solver thread per usable core, and 2x (SMT) — the better one is reported
        ref = Oracle("ref_fast")
        best = None
        for t in ([threads] if threads else [cores, 2 * cores]):
            m = max(t, (n // 2 // t) * t) if not threads else n
            m = min(m, n)
            s_ = ref.bench_solve(D, F, m, script_path=script_net_file(D, F), threads=t, num_iters=iters, beliefs=beliefs[:m])
            if best is None or m / s_ > best[0] / best[1]:
                best = (m, s_, t)
        n, secs, threads = best
        kind = "reference"
    else:
        import torch  # noqa: F401
        from rebel_b200.models import flatten_state_dict, make_selfplay_net
        port = Oracle("port")
        w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
        secs = port.bench_solve(D, F, n, net_w=w, num_iters=iters, beliefs=beliefs[:n])
        kind, threads = "port", 1
    rate = n * iters / secs
    sample = (f"{n} of the workload's root subgames x {iters} iters, build_solver+multistep with TorchScript Net2 on CPU, {threads} solver "
              f"threads on {cores} usable cores (os.cpu_count()={os.cpu_count()}), {secs:.1f} s wall" if kind == "reference"
              else f"{n} root subgames x {iters} iters, C port, 1 thread, {secs:.1f} s wall")
    return rate, {"value": rate, "unit": UNIT, "cores": cores if kind == "reference" else 1, "threads": threads, "kind": kind, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, F = args.dice, args.faces
    A, H, Q = dims(D, F)
    from oracle.oracle import available
    cores = effective_cores()
    per_core = 12 if available("ref_fast") else 1
    n = args.cpu_sample or max(cores * per_core if available("ref_fast") else 2, 2)
    beliefs = workload_beliefs(n, H, 0)
    for _ in range(args.warmup):
        cpu_reference_rate(D, F, args.iters, min(n, cores if available("ref_fast") else 1), beliefs)
    t0 = time.time()
    rates, info = [], None
    for _ in range(args.steps):
        r, info = cpu_reference_rate(D, F, args.iters, n, beliefs)
        rates.append(r)
    wall = time.time() - t0
    value = n * args.iters * args.steps / sum(n * args.iters / r for r in rates)
    info["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (CFR) / f32 (value net)", "data": "synthetic",
        # the workload is the B200 arm's (args.subgames concurrent subgames); each step times a bounded sample of it (cpu_baseline.sample)
        "config": dict(workload_config(args, args.subgames), parallelism="cpu threads"), "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def workload_config(args, k):
    return {"workload": f"{args.dice}x{args.faces}f Liar's Dice, depth-2 root subgames (last_bid=-1), {k} concurrent subgames per GPU, "
                        f"{args.iters} linear-CFR iterations each, Net2(256x2,LayerNorm) leaf value net every iteration",
            "subgames_per_gpu": k, "cfr_iters": args.iters, "max_depth": 2, "value_net": "Net2 n_hidden=256 n_layers=2 layer_norm, random init seed 0",
            "beliefs": "random (Philox counter streams)"}


def run_b200(args):
    import torch
    import rebel_b200 as rb
    from rebel_b200 import dist as rbdist
    from rebel_b200.models import flatten_state_dict, make_selfplay_net

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun on this node
        port = 29500 + os.getpid() % 1000
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", str(port)] + sys.argv)
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    D, F, K, iters = args.dice, args.faces, args.subgames, args.iters
    A, H, Q = dims(D, F)

    # ---- value-net weights: rank 0 owns them, NCCL broadcast to the other ranks (ModelLocker::updateModel analogue)
    nflat = 256 * Q + 3 * 256 + 256 * 256 + 3 * 256 + H * 256 + H
    w = rbdist.broadcast_weights(flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict()) if rank == 0 else None, nflat, dev)

    mode, mode_name = {"auto": (rb.NET_TC_F16X2, "tc_f16x2"), "tcx2": (rb.NET_TC_F16X2, "tc_f16x2"), "tc": (rb.NET_TC_F16, "tc_f16"),
                       "fp32": (rb.NET_FP32, "fp32")}[args.net]
    is_tc = mode in (rb.NET_TC_F16, rb.NET_TC_F16X2)
    S = rb.WaveSolver(D, F, K, num_iters=iters, net_mode=mode, device=local)
    S.set_weights(w, version=1)

    # ---- this rank's shard of the job: subgames [rank*K, (rank+1)*K); inputs staged in pinned host memory
    beliefs64 = workload_beliefs(K, H, rbdist.shard_range(rank, world, K)[0])
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    t_b = pin((K, 2, H), torch.float64); t_b.numpy()[:] = beliefs64
    t_lb = pin((K,), torch.int32); t_lb.fill_(-1)
    t_pl = pin((K,), torch.int32); t_pl.zero_()
    t_act = pin((K,), torch.int32); t_act.numpy()[:] = np.random.RandomState(rank).randint(0, iters + 1, size=K)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return rbdist.max_over_ranks(ms, dev)

    # ================= device-resident throughput (`value`) =================
    S.begin(t_lb.numpy(), t_pl.numpy(), t_b.numpy(), t_act.numpy())
    S.set_profiling(int(os.environ.get("BENCH_PROFILE_EVERY", "16")))   # CUDA-event pairs around every 16th value-net launch of the timed steps
    sampler = ClockSampler(local)           # NVML is initialised and polling before the warm-up: its start-up (driver locks) must not
    if rank == 0 and os.environ.get("BENCH_NO_SAMPLER") != "1":   # land in the timed region; only samples taken inside it are reported
        sampler.start()
    for _ in range(max(args.warmup, 3)):    # (at least 3: eager, graph capture, replay) exactly a timed step: the L2 flush too (the first launch of torch's fill kernel loads its
        flush.fill_(1)                      # module lazily, ~70 ms of host time), and the same run configuration (the run is replayed
        S.reset(stream); S.run(iters, stream)   # from a CUDA graph that is built the second time a configuration is requested)
    barrier()
    launches0 = S.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net_ms = 0.0
    run_ms, wall = [], []
    t_region0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        tw = time.perf_counter()
        flush.fill_(1)                      # evict L2 between steps (inside the timed region)
        S.reset(stream)
        S.run(iters, stream)
        tr, tn = S.last_run_ms()            # waits for this step; value-net kernel time from per-launch CUDA events
        net_ms += tn
        run_ms.append(round(tr, 2)); wall.append(round(1e3 * (time.perf_counter() - tw), 2))
    e1.record()
    barrier()
    if rank == 0:
        sampler.window(t_region0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = S.kernel_launches - launches0
    S.set_profiling(False)
    value = world * K * iters * args.steps / (ms * 1e-3)

    # ================= end to end through the host-buffer C ABI (`e2e`) =================
    ex_all = None
    def e2e_step():
        S.begin(t_lb.numpy(), t_pl.numpy(), t_b.numpy(), t_act.numpy())      # H2D from pinned memory
        S.run(iters, stream)
        q, v = S.examples()                                                   # D2H: training examples of the wave
        return rbdist.gather_examples(q, v, dev)                              # NCCL gather of the example blocks on rank 0
    for _ in range(2):                      # eager, then graph capture of the (unprofiled) run configuration
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        flush.fill_(1)
        ex_all = e2e_step()
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1))
    e2e_value = world * K * iters * args.steps / (ms_e2e * 1e-3)
    h2d = K * 2 * H * 8 + 3 * K * 4
    d2h = K * 2 * H * 4

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    # ================= roofline of the dominant kernel (value net) + CPU baseline =================
    peaks = load_peaks()
    rows = S.leaf_rows
    n_net = args.steps * iters
    flops_launch = 2.0 * rows * (256 * Q + 256 * 256 + 256 * H)
    avg_net_ms = net_ms / max(n_net, 1)
    traffic = None   # DRAM bytes per launch of the value-net kernel from the committed ncu --set full capture (same workload only)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if is_tc and (D, F, K) == (1, 6, 8192):
            traffic = [v["dram_bytes_per_launch"] for k, v in tj.items() if "leaf_mlp_tc_kernel" in k][0]
    except Exception:
        traffic = None
    achieved = flops_launch / (avg_net_ms * 1e-3) / 1e12 if avg_net_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "leaf value net (Net2 forward over all pseudo-leaf rows of the wave)",
                "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "peak_source": peaks["src"], "avg_launch_ms": avg_net_ms, "launch_timing": "CUDA events around every 16th launch inside the timed steps",
                "step_run_ms": run_ms, "step_wall_ms": wall,
                "rows_per_launch": rows,
                "flops_per_launch": flops_launch, "share_of_step": net_ms / ms if ms > 0 else None,
                "cfr_tables_algorithmic_GBps": (4 * H * (90 + 6 * 45) + 4 * 66 * (Q + H) + 8 * H) * K * iters * args.steps / max(ms - net_ms, 1e-9) / 1e6
                if (D, F) == (1, 6) else None}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (CFR tables) / " + ("f16 operands, f32 accumulate + LayerNorm, " + ("f16x2" if mode == rb.NET_TC_F16X2 else "f32") + " GELU (value net, tcgen05)" if is_tc else "f32 (value net)"),
        "data": "synthetic", "config": dict(workload_config(args, K), value_net_kernel=mode_name, parallelism=f"dp{world}",
                                           l2="256 MiB memset between steps, inside the timed region"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import available
        cores = effective_cores()
        n = args.cpu_sample or (cores * 24 if available("ref_fast") else 3)
        _, info = cpu_reference_rate(D, F, iters, n, beliefs64)
        out["cpu_baseline"] = info
    print(json.dumps(out), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
