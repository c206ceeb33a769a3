#!/usr/bin/env python
"""bench.py -- tiles/sec of the HoVer-Net tile hot path (infer_step + post_proc.process per patch).

    python bench.py --gpus N --steps K --warmup W            # this repo (libhvn, sm_100a kernels)
    python bench.py --impl reference --gpus N ...            # the reference path on the host cores

Workload (BASELINE.json configs[1], restated per SURVEY.md 8d because `fast` mode cannot ingest
270x270): a step = one batch of 64 synthetic 256x256x3 uint8 patches, `fast` mode, nr_types=6, each
patch through the CNN (-> 164x164x4 float map) and the instance post-processing (-> inst_map +
instance table).  Weights: seeded synthetic checkpoint (hover_net_b200.synth).  One process per GPU;
each rank runs its own batch (weak scaling) and rank 0 gathers the instance tables over NCCL.

`value`  : device-timed (CUDA events on the library's stream), inputs resident in HBM.
`e2e`    : same metric through the host-buffer C-ABI call (`hvn_forward_postproc`): pinned host
           uint8 in, inst_map + instance table out, copies inside the timed region.
`roofline`: dominant kernel = the convolution kernel class; algorithmic 2*MACs of the reference
           graph per launch / its measured launch time (per-launch CUDA events) vs the measured
           dense-bf16 peak in MEASURED_PEAKS.json.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLOCK_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
           "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
           "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks/throttle reasons for one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + CLOCK_Q,
                                  "--format=csv,noheader,nounits", "-lms", "200"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        try:
            while not self.stop_flag.is_set():
                line = p.stdout.readline()
                if not line:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        finally:
            p.terminate()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measured_traffic(kernel_class):
    """dram__bytes_read+write per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    try:
        import glob
        path = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1]  # latest committed capture
        t = json.load(open(path))
        k = t["kernels"].get("k_conv_tc" if kernel_class == "conv_tc" else "k_conv_ref")
        return {"dram_bytes_per_launch": k["dram_bytes"] / k["launches"], "launches": k["launches"], "batch": t["batch"],
                "source": "profiles/" + os.path.basename(path) + " (" + t["command"] + ")"}
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def np_from_addr(addr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * n).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


# ------------------------------------------------------------------------------------------------
def cpu_oracle_step(x, sd, mode, nt, pool, torch, O, P, device="cpu"):
    """One reference-style pass over the patches in x: forward, then pooled post-processing."""
    pred = O.infer_step(x, sd, mode, nt, device=device)
    if pool is None:
        res = [P.process_table(pred[i], nt) for i in range(pred.shape[0])]
    else:
        res = pool.map(_pp_worker, [(pred[i], nt) for i in range(pred.shape[0])])
    return pred, res


def _pp_worker(args):
    from oracle import postproc_oracle as P
    return P.process_table(args[0], args[1])


def run_reference(args):
    """Reference arm: the reference's path (torch graph + cv2/scipy-faithful post-proc restatement,
    pooled over the host cores as infer/tile.py:232-234 does) on a bounded sample per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import multiprocessing as mp
    import torch
    from hover_net_b200 import synth
    from oracle import hovernet_torch as O
    from oracle import postproc_oracle as P
    P.build()
    mode, nt = args.mode, args.nr_types
    cores = os.cpu_count() or 1
    threads = min(cores, args.cpu_threads)  # torch's CPU conv degrades badly past ~32 threads
    torch.set_num_threads(threads)
    sd = O.to_torch_state_dict(synth.make_state_dict(mode, nt, seed=0))
    sample = args.ref_sample
    x = synth.make_patches(sample, 256 if mode == "fast" else 270, seed=1)
    fwd_dev = "cuda" if (args.ref_forward == "cuda" and torch.cuda.is_available()) else "cpu"
    sdd = {k: v.to(fwd_dev) for k, v in sd.items()}
    workers = min(cores, sample)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    for _ in range(args.warmup):
        cpu_oracle_step(x, sdd, mode, nt, pool, torch, O, P, fwd_dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(x, sdd, mode, nt, pool, torch, O, P, fwd_dev)
    if fwd_dev == "cuda":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    if pool:
        pool.close()
    val = sample / dt
    line = {
        "impl": "reference", "metric": "tiles/sec", "value": val, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note="bounded sample of %d patches per step" % sample),
        "cpu_baseline": {"value": val, "unit": "tiles/s", "cores": max(threads, workers), "host_cores": cores, "kind": "port",
                         "sample": "%d patches/step: torch fp32 forward on %s (%d threads) + oracle post-proc in a "
                                   "%d-process pool" % (sample, fwd_dev, threads, workers)},
        "e2e": {"value": val, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, note=None):
    cfg = {"workload": "BASELINE configs[1]: batch=%d synthetic %s-mode patches %dx%dx3 uint8, nr_types=%s, "
                       "forward + instance post-processing per patch (output %s)" % (
                           args.batch, args.mode, args.patch, args.patch, args.nr_types,
                           "164x164" if args.mode == "fast" else "80x80"),
           "mode": args.mode, "nr_types": args.nr_types, "batch_per_gpu": args.batch, "patch": args.patch,
           "parallelism": "tiles sharded per rank (dp%d), end-of-batch NCCL gather of instance tables" % args.gpus,
           "l2": "per-step activations (~%.1f GB) exceed the 126 MB L2; a 256 MB buffer is also written "
                 "between timed steps" % (0.5 * args.batch)}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    import torch
    from hover_net_b200 import _lib, synth
    from hover_net_b200.dist import gather_tables
    from hover_net_b200.models.hovernet.net_desc import create_model

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    mode, nt, B, S = args.mode, args.nr_types, args.batch, args.patch
    net = create_model(mode=mode, input_ch=3, nr_types=nt, device=local)
    net.load_state_dict(synth.make_state_dict(mode, nt, seed=0), strict=True)
    ctx = net.ctx
    if args.chunk:
        ctx.set_option("chunk", args.chunk)
    if args.conv_path is not None:
        ctx.set_option("conv_path", args.conv_path)
    oh, ow, oc = ctx.out_shape(S, S)
    max_rows = max(16, oh * ow // 64)
    x = synth.make_patches(min(B, 8), S, seed=1 + rank)
    x = np.concatenate([x] * ((B + x.shape[0] - 1) // x.shape[0]))[:B]
    rng = np.random.default_rng(rank)
    x = np.ascontiguousarray(x[rng.permutation(B)])
    in_bytes = x.nbytes

    # device-resident buffers are torch tensors (so NCCL can gather them); libhvn gets raw pointers
    d_img = torch.from_numpy(x).cuda()
    d_inst = torch.empty((B, oh, ow), dtype=torch.int32, device="cuda")
    d_tab = torch.zeros((B, max_rows, 10), dtype=torch.int64, device="cuda")
    d_nr = torch.zeros((B,), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    gathered = [None, None]
    torch.cuda.synchronize()

    def step_resident():
        ctx.forward_postproc_dev(d_img.data_ptr(), B, S, S, None, d_inst.data_ptr(), d_tab.data_ptr(), max_rows,
                                 d_nr.data_ptr())

    def gather():
        if world > 1:
            ctx.sync()
            gathered[0], gathered[1] = gather_tables(d_tab, d_nr)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
        gather()
    barrier()
    launches0 = ctx.counter("kernel_launches")
    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed: K steps, device time per step (events on the library's stream), L2 flushed between
    step_ms = []
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_start()
        step_resident()
        ms = ctx.timer_stop()
        if world > 1:
            t0 = time.perf_counter()
            gather()
            torch.cuda.synchronize()
            ms += (time.perf_counter() - t0) * 1e3
        step_ms.append(ms)
    barrier()
    wall = time.perf_counter() - t_wall
    launches = ctx.counter("kernel_launches") - launches0
    total_ms = float(np.sum(step_ms))

    # ---- e2e through the host-buffer entry point (pinned host buffers)
    h_in = ctx.malloc_host(in_bytes)
    h_inst = ctx.malloc_host(B * oh * ow * 4)
    h_tab = ctx.malloc_host(B * max_rows * 10 * 8)
    h_nr = ctx.malloc_host(B * 4)
    np_from_addr(h_in, x.shape, np.uint8)[...] = x
    L = _lib.lib()

    def step_e2e():
        _lib.check(L.hvn_forward_postproc(ctx._h, ctypes.c_void_p(h_in), B, S, S, None, ctypes.c_void_p(h_inst),
                                          ctypes.c_void_p(h_tab), max_rows, ctypes.c_void_p(h_nr)))

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.timer_start()
        step_e2e()
        e2e_ms.append(ctx.timer_stop())
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    e2e_total = float(np.sum(e2e_ms))
    inst_e2e = np_from_addr(h_inst, (B, oh, ow), np.int32).copy()
    nr_e2e = np_from_addr(h_nr, (B,), np.int32).copy()

    # ---- per-kernel-class profile pass (untimed): per-launch CUDA events
    ctx.set_option("branch_streams", 0)  # isolated kernels: per-launch event times must not overlap
    ctx.set_option("profile", 2)
    step_resident()
    step_resident()
    ctx.sync()
    prof = {}
    for cls in ("conv_tc", "conv_ref", "conv0", "bnrelu", "head"):
        prof[cls] = {"ms": ctx.stage_ms(cls), "launches": ctx.counter("launches:" + cls),
                     "gflop": ctx.counter("flops:" + cls) / 1e9}
    prof["postproc"] = {"ms": ctx.stage_ms("postproc"), "launches": ctx.counter("pp_launches")}
    prof["cnn_total_ms"] = ctx.stage_ms("cnn")
    flops_step = float(ctx.counter("last_flops"))
    ctx.set_option("profile", 0)
    ctx.set_option("branch_streams", 1)

    # sanity: resident and e2e paths agree
    same = bool(np.array_equal(d_inst.cpu().numpy(), inst_e2e)) and bool(np.array_equal(d_nr.cpu().numpy(), nr_e2e))

    # max over ranks
    t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_total = float(t[0]), float(t[1])

    if rank == 0:
        pk, pk_kind = peaks()
        dom = "conv_tc" if prof["conv_tc"]["launches"] > 0 else "conv_ref"
        d = prof[dom]
        ach = (d["gflop"] / 1e3) / (d["ms"] / 1e3) if d["ms"] > 0 else 0.0  # TFLOP/s
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        value = world * B * args.steps / (total_ms / 1e3)
        e2e_v = world * B * args.steps / (e2e_total / 1e3)
        line = {
            "metric": "tiles/sec", "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16x3-split (fp32-equivalent products, fp32 accumulate); "
                                                              "post-proc int32/f64",
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_v, "unit": "tiles/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": int(B * oh * ow * 4 + B * max_rows * 80 + B * 4),
                    "ms_per_step": e2e_total / args.steps, "matches_resident_path": same},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": ach / peak if peak else None, "peak_source": pk_kind + " bf16_tflops_sustained",
                         "algorithmic_gflop_per_launch": d["gflop"] / max(1, d["launches"]),
                         "avg_launch_ms": d["ms"] / max(1, d["launches"]), "launches_per_step": d["launches"],
                         "traffic": measured_traffic(dom),
                         "note": "algorithmic 2*MACs of the reference graph; every product is formed three times in "
                                 "fp16 (hi*hi+hi*lo+lo*hi, issued as two MMA instructions per K-step), so frac <= 1/3 "
                                 "by construction"},
            "kernel_classes": prof,
            "algorithmic_gflop_per_tile": flops_step / B / 1e9,
            "postproc_ms_per_tile": prof["postproc"]["ms"] / B,
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for p in (h_in, h_inst, h_tab, h_nr):
        ctx.free_host(p)
    ctx.close()


def cpu_baseline(args):
    """Oracle (CPU port of the reference path) on the host cores, bounded sample."""
    import multiprocessing as mp
    import torch
    from hover_net_b200 import synth
    from oracle import hovernet_torch as O
    from oracle import postproc_oracle as P
    P.build()
    cores = os.cpu_count() or 1
    threads = min(cores, args.cpu_threads)
    torch.set_num_threads(threads)
    mode, nt = args.mode, args.nr_types
    sd = O.to_torch_state_dict(synth.make_state_dict(mode, nt, seed=0))
    n = args.cpu_sample
    x = synth.make_patches(n, args.patch, seed=1)
    workers = min(cores, n)
    pool = mp.get_context("fork").Pool(workers) if workers > 1 else None
    cpu_oracle_step(x[:2], sd, mode, nt, pool, torch, O, P)  # warm-up
    t0 = time.perf_counter()
    pred = O.infer_step(x, sd, mode, nt)
    t1 = time.perf_counter()
    if pool:
        pool.map(_pp_worker, [(pred[i], nt) for i in range(n)])
    else:
        [P.process_table(pred[i], nt) for i in range(n)]
    t2 = time.perf_counter()
    t3 = time.perf_counter()
    for i in range(min(n, 8)):
        P.process_table(pred[i], nt)
    pp1 = (time.perf_counter() - t3) / min(n, 8)
    if pool:
        pool.close()
    return {"value": n / (t2 - t0), "unit": "tiles/s", "cores": max(threads, workers), "host_cores": cores, "kind": "port",
            "sample": "%d patches: torch fp32 CPU forward (%d threads) %.2fs + oracle post-proc in a %d-process pool "
                      "%.3fs; single-core post-proc %.2f ms/tile" % (n, threads, t1 - t0, workers, t2 - t1, pp1 * 1e3),
            "forward_s_per_tile": (t1 - t0) / n, "postproc_ms_per_tile_1core": pp1 * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["fast", "original"])
    ap.add_argument("--nr-types", type=int, default=6)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--conv-path", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--ref-sample", type=int, default=8)
    ap.add_argument("--cpu-threads", type=int, default=32)
    ap.add_argument("--ref-forward", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.nr_types == 0:
        args.nr_types = None
    args.patch = 256 if args.mode == "fast" else 270
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
