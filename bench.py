#!/usr/bin/env python
"""Headline benchmark: fused depth-loss forward+backward throughput (Mpix/s) and fraction of the HBM
roofline (BASELINE.json `metric`), on synthetic SUN RGB-D-shaped inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode local|global]

A "step" is one pass of the hot path over one batch: phase A (reduce) + phase B (gradient) of the full
loss stack (SI + gradient matching + smoothness + reprojection) with both metric variants fused in,
B=32 images of 480x640 per GPU (BASELINE config 3).  Prints ONE JSON line (rank 0).

  value     whole-job Mpix/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the reference-facing C++ API (drop-in CombinedDepthLoss +
            DepthMetrics via host/libcadl_host.so) with pinned HOST buffers: H2D of pred/gt/rgb/K and the
            D2H of the loss scalar + metric blocks are inside the timed region
  roofline  for the dominant kernel (stream3_kernel, the full-resolution gradient pass): algorithmic bytes (24 B/px:
            read pred, gt, 3 x rgb, write grad) / its CUDA-event duration -- events recorded on the launching stream
            around 20 back-to-back cadl_stack_grad calls after cadl_stack_prepare + cadl_stack_reduce have completed
            (one launch each: the average launch duration) -- against MEASURED_PEAKS.json hbm_gbs
  also      config 2 (reprojection alone), config 5's per-GPU shape (B=16 960x1280, reprojection + both metric
            variants), and for N > 1 the exact-global-batch mode with both statistics exchanges
  cpu_baseline  the unmodified reference headers on LibTorch CPU (oracle/_ref) timed on this box

--impl reference times the reference's own CPU implementation (same harness source compiled against the
reference headers) on a bounded sample of the same workload; rank 0 only.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG_NAME = "camera-aware-neural-networks-for-few-view-depth-estimation_b200"

B_PER_GPU, H, W = 32, 480, 640          # BASELINE config 3 (and config 2 for the reprojection-only line)
ALGO_BYTES_PER_PX = 24                  # read pred 4 + gt 4 + rgb 12, write grad 4
ALGO_BYTES_PER_PX_REPROJ = 12           # config 2: read pred 4 + gt 4, write grad 4
FALLBACK_HBM_GBS = 6650.0               # /opt/skills/guides/B200_PROFILING.md fallback


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args, pkg):
    """The reference's own CPU implementation of the path, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libcadl_refharness.so")
    nthreads = os.cpu_count() or 1
    kind = "reference"
    if os.path.exists(ref_so):
        h = pkg.StepHarness(ref_so)
        h.set_num_threads(nthreads)

        def run_steps(Bs, warm, iters):
            b = pkg.synth.make_batch(Bs, H, W, seed=1234)
            z = {k: v.numpy() for k, v in b.items()}
            ms, _ = h.time_steps(pkg.StepCfg(device=-1, term=0), z["pred"], z["gt"], z["rgb"], z["K"],
                                 with_metrics=True, include_h2d=False, warmup=warm, iters=iters)
            return ms
    else:   # the oracle port (same op chain in Python torch) -- the one other place bench may execute oracle/
        kind = "port"
        import torch
        torch.set_num_threads(nthreads)
        oracle = importlib.import_module("oracle.oracle_torch")

        def run_steps(Bs, warm, iters):
            b = pkg.synth.make_batch(Bs, H, W, seed=1234)
            out = []
            for it in range(warm + iters):
                t0 = time.perf_counter()
                p = b["pred"].clone().requires_grad_(True)
                tot, _ = oracle.combined_loss(p, b["gt"], b["rgb"], b["K"])
                tot.sum().backward()
                oracle.metrics_eval(b["pred"], b["gt"]); oracle.metrics_train(b["pred"], b["gt"])
                float(tot)
                if it >= warm:
                    out.append((time.perf_counter() - t0) * 1e3)
            return out

    probe = statistics.median(run_steps(4, 1, 2))            # ms for 4 images
    budget_ms = 150e3
    Bs = 4
    for cand in (32, 16, 8, 4):
        if (args.steps + args.warmup) * probe * cand / 4.0 <= budget_ms:
            Bs = cand
            break
    ms = run_steps(Bs, args.warmup, args.steps)
    step_ms = sum(ms) / len(ms)
    px = Bs * H * W
    value = px / (step_ms * 1e-3) / 1e6
    sample = f"{Bs} of {B_PER_GPU} images at {H}x{W}, full stack fwd+bwd + both metric variants, {len(ms)} timed steps"
    line = {
        "impl": "reference", "metric": "fused depth-loss fwd+bwd throughput", "value": value, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"config3: SI+grad+smooth+reproj fwd+bwd + metrics, B={Bs} (sample of {B_PER_GPU}) "
                               f"{H}x{W}, LibTorch CPU"},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": nthreads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="mode global: statistics exchange by our peer-memory kernel over NVLink (default) or NCCL all-reduce")
    ap.add_argument("--mode", default="local", choices=["local", "global"],
                    help="multi-GPU: local = each rank evaluates its own batch (DDP semantics, no exchange); "
                         "global = exact global-batch loss, one all-reduce of the 32-double statistics vector")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    pkg = importlib.import_module(PKG_NAME)
    if args.impl == "reference":
        reference_arm(args, pkg)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    P = B_PER_GPU * H * W
    b = pkg.synth.make_batch(B_PER_GPU, H, W, seed=1234 + rank, device=dev)
    pred, gt, rgb, K = b["pred"], b["gt"], b["rgb"], b["K"]
    grad = torch.empty_like(pred)
    ws = pkg.Workspace(B_PER_GPU, H, W, dev)
    params = pkg.default_params(metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
    if args.mode == "global" and world > 1:
        params.global_B = B_PER_GPU * world

    p2p = None
    if args.mode == "global" and world > 1 and args.exchange == "p2p":
        p2p = pkg.multi.P2PStatsExchange(pkg, dev)
        # one-off check: the peer-memory exchange and the NCCL all-reduce give the same sums
        pkg.stack_reduce(pred, gt, None, params, ws)
        ref_stats = ws.stats_view().clone()
        dist.all_reduce(ref_stats)
        p2p.exchange(ws)
        torch.cuda.synchronize()
        rel = float(((ws.stats_view() - ref_stats).abs() / ref_stats.abs().clamp_min(1e-300)).max())
        assert not p2p.timed_out() and rel <= 1e-12, f"p2p exchange disagrees with NCCL: {rel}"

    def step():
        if args.mode == "global" and world > 1:
            pkg.stack_prepare(pred, gt, params, ws)          # pyramid kernels beside phase A and the exchange
            pkg.stack_reduce(pred, gt, None, params, ws)
            if p2p is not None:
                p2p.exchange(ws)                             # 32 doubles pushed into every rank's inbox over NVLink
            else:
                dist.all_reduce(ws.stats_view())             # the same through NCCL, stream-ordered
            pkg.stack_grad(pred, gt, rgb, K, None, params, grad, ws)
        else:
            pkg.stack_fwd_bwd(pred, gt, rgb, K, None, params=params, grad=grad, ws=ws)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps      # ms per step, max over ranks

    # ---- headline: K steps of the fused path, device-resident inputs (235 MB/step > 126 MB L2) ----
    with ClockSampler(local_rank) as clk:
        ms_step = timed(step, args.steps, args.warmup)
        # the dominant kernel alone: prepare (pyramid kernels) + reduce, wait for them, then ONE launch between two events
        # on the launching stream (the product library has no per-kernel timing hook)
        loc = pkg.default_params(metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
        ks, ka = [], []
        NREP = 20
        for it in range(2 + 10):
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pkg.stack_prepare(pred, gt, loc, ws)
            pkg.stack_reduce(pred, gt, None, loc, ws)
            a1.record()
            torch.cuda.synchronize()
            # the statistics and the coarse-scale field stay valid: the gradient kernel can be launched again and again
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(NREP):
                pkg.stack_grad(pred, gt, rgb, K, None, loc, grad, ws)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ks.append(e0.elapsed_time(e1) / NREP); ka.append(a0.elapsed_time(a1))
        ms_k, ms_a = statistics.median(ks), statistics.median(ka)
        also = []
        # config 2 (reprojection alone)
        p2 = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0)
        ms_rp = timed(lambda: pkg.stack_fwd_bwd(pred, gt, None, K, None, params=p2, grad=grad, ws=ws), args.steps, 3)
        # config 5, per-GPU shape: B=16 at 960x1280, reprojection + both metric variants
        b5 = pkg.synth.make_batch(16, 960, 1280, seed=77 + rank, device=dev, with_rgb=False)
        g5, ws5 = torch.empty_like(b5["pred"]), pkg.Workspace(16, 960, 1280, dev)
        p5 = pkg.default_params(terms=pkg.TERM_REPROJ, w_reproj=1.0, metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN)
        ms_c5 = timed(lambda: pkg.stack_fwd_bwd(b5["pred"], b5["gt"], None, b5["K"], None, params=p5, grad=g5, ws=ws5),
                      max(10, args.steps // 2), 3)
        del b5, g5, ws5
        # N > 1: the exact-global-batch mode, statistics exchanged by our peer-memory kernel and by NCCL
        ms_global = {}
        if world > 1:
            gp = pkg.default_params(metrics=pkg.METRICS_EVAL | pkg.METRICS_TRAIN, global_B=B_PER_GPU * world)
            ex = pkg.multi.P2PStatsExchange(pkg, dev)

            def gstep(how):
                pkg.stack_prepare(pred, gt, gp, ws)
                pkg.stack_reduce(pred, gt, None, gp, ws)
                if how == "p2p":
                    ex.exchange(ws)
                else:
                    dist.all_reduce(ws.stats_view())
                pkg.stack_grad(pred, gt, rgb, K, None, gp, grad, ws)
            for how in ("p2p", "nccl"):
                ms_global[how] = timed(lambda: gstep(how), args.steps, args.warmup)
            ex.check()
            ex.close()
        # the timed loops above last ~0.1 s in total: keep the same step running (untimed) for about a second so that
        # nvidia-smi (100 ms per query) sees the clocks UNDER this load, not an idle GPU
        t_end = time.perf_counter() + 1.2
        while time.perf_counter() < t_end:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
    clocks = clk.summary()
    value = world * P / (ms_step * 1e-3) / 1e6

    peak, peak_src = hbm_peak()
    achieved = ALGO_BYTES_PER_PX * P / (ms_k * 1e-3) / 1e9
    traffic = None
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            traffic = json.load(open(traffic_file)).get("stream3_kernel_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "stream3_kernel<15,false>",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": "ncu --set full capture of the same kernel (profiles/traffic.json), static: not measured in this run",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * P,
                "kernel_ms": ms_k, "first_stage_ms": ms_a,
                "how": "CUDA events on the launching stream around 20 back-to-back cadl_stack_grad calls (one launch each), median of 10 such groups",
                "step_frac": ALGO_BYTES_PER_PX * P / (ms_step * 1e-3) / 1e9 / peak}
    also.append({"workload": "config2: reprojection alone fwd+bwd, B=32/GPU 480x640", "ms_per_step": ms_rp,
                 "value": world * P / (ms_rp * 1e-3) / 1e6, "unit": "Mpix/s",
                 "roofline_frac": ALGO_BYTES_PER_PX_REPROJ * P / (ms_rp * 1e-3) / 1e9 / peak})
    P5 = 16 * 960 * 1280
    also.append({"workload": "config5 shape: reprojection fwd+bwd + DepthMetrics + computeDepthMetrics, B=16/GPU 960x1280",
                 "ms_per_step": ms_c5, "value": world * P5 / (ms_c5 * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world,
                 "roofline_frac": ALGO_BYTES_PER_PX_REPROJ * P5 / (ms_c5 * 1e-3) / 1e9 / peak})
    for how, ms in ms_global.items():
        also.append({"workload": "config3, exact global-batch mode (statistics of all ranks before the gradient pass)",
                     "stats_exchange": "peer-memory kernel over NVLink (cadl_stats_exchange)" if how == "p2p" else "NCCL all-reduce of 32 doubles",
                     "ms_per_step": ms, "value": world * P / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": world})

    # ---- e2e: the drop-in C++ API with pinned host buffers (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host = pkg.host_harness()
        z = {k: v.cpu().numpy() for k, v in b.items()}
        iters = max(5, min(args.steps, 20))
        ms, _ = host.time_steps(pkg.StepCfg(device=local_rank, term=0), z["pred"], z["gt"], z["rgb"], z["K"],
                                with_metrics=True, include_h2d=True, warmup=3, iters=iters)
        ms_e2e = max_over_ranks(sum(ms) / len(ms))
        e2e = {"value": world * P / (ms_e2e * 1e-3) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(4 * (pred.numel() + gt.numel() + rgb.numel() + K.numel())),
               "d2h_bytes_per_step": 4 + 2 * 232, "ms_per_step": ms_e2e,
               "api": "CombinedDepthLoss::forwardWithIntrinsics + backward + DepthMetrics::compute + "
                      "computeDepthMetrics (host/libcadl_host.so)"}

    # ---- CPU baseline: the unmodified reference on this box's host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_so = os.path.join(ROOT, "oracle", "_ref", "libcadl_refharness.so")
        nthreads = os.cpu_count() or 1
        bc = pkg.synth.make_batch(B_PER_GPU, H, W, seed=1234)
        zc = {k: v.numpy() for k, v in bc.items()}
        if os.path.exists(ref_so):
            h = pkg.StepHarness(ref_so)
            h.set_num_threads(nthreads)
            ms, _ = h.time_steps(pkg.StepCfg(device=-1, term=0), zc["pred"], zc["gt"], zc["rgb"], zc["K"],
                                 with_metrics=True, include_h2d=False, warmup=1, iters=6)
            kind = "reference"
        else:
            oracle = importlib.import_module("oracle.oracle_torch")
            torch.set_num_threads(nthreads)
            ms = []
            for it in range(4):
                t0 = time.perf_counter()
                pc = bc["pred"].clone().requires_grad_(True)
                tot, _ = oracle.combined_loss(pc, bc["gt"], bc["rgb"], bc["K"])
                tot.sum().backward()
                oracle.metrics_eval(bc["pred"], bc["gt"]); oracle.metrics_train(bc["pred"], bc["gt"])
                if it:
                    ms.append((time.perf_counter() - t0) * 1e3)
            kind = "port"
        med = statistics.median(ms)
        cpu = {"value": P / (med * 1e-3) / 1e6, "unit": "Mpix/s", "cores": nthreads, "kind": kind,
               "sample": f"full config-3 batch (32x480x640), fwd+bwd + both metric variants, median of {len(ms)} "
                         f"steps after 1 warm-up, {med:.0f} ms/step"}
        # the other CPU rows of BASELINE.md section 4 (reference build only): config 2 = reprojection alone on the same
        # batch; config 1 = BaselineUNet(3,64,10) + full stack + clip + Adam, B=4 at 240x320
        if kind == "reference":
            others = []
            ms2, _ = h.time_steps(pkg.StepCfg(device=-1, term=4), zc["pred"], zc["gt"], zc["rgb"], zc["K"],
                                  with_metrics=False, include_h2d=False, warmup=1, iters=5)
            m2 = statistics.median(ms2)
            others.append({"workload": "config2: ReprojectionLoss::forward + backward, B=32 480x640, LibTorch CPU",
                           "ms_per_step": m2, "value": P / (m2 * 1e-3) / 1e6, "unit": "Mpix/s"})
            if h.has_unet():
                b1 = pkg.synth.make_batch(4, 240, 320, seed=1234)
                z1 = {k: v.numpy() for k, v in b1.items()}
                r1 = h.unet_train(z1["rgb"], z1["gt"], z1["K"], device=-1, feats=64, warmup=1, iters=2)
                m1, l1 = statistics.median(r1["ms"]), statistics.median(r1["loss_ms"])
                others.append({"workload": "config1: BaselineUNet(3,64,10) training step (forward, forwardWithIntrinsics, "
                                           "backward, clip_grad_norm_, Adam), B=4 240x320, LibTorch CPU",
                               "ms_per_step": m1, "samples_per_s": 4 / (m1 * 1e-3), "loss_forward_ms": l1,
                               "value": 4 * 240 * 320 / (m1 * 1e-3) / 1e6, "unit": "Mpix/s", "params": r1["params"]})
            cpu["others"] = others

    if rank == 0:
        line = {
            "metric": "fused depth-loss fwd+bwd throughput", "value": value, "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config3: full loss stack (SI+grad+smooth+reproj) fwd+bwd + DepthMetrics + "
                                   "computeDepthMetrics, B=32/GPU 480x640",
                       "pixels_per_step_per_gpu": P, "multi_gpu_mode": args.mode,
                       "stats_exchange": (args.exchange if (args.mode == "global" and world > 1) else "none"),
                       "l2": "inputs+gradient 275 MB per step > 126 MB L2 (no flush needed)",
                       "seed": "1234 + rank"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": 4 * args.steps,
            "launches_per_step": ["phase_a_kernel<31,false>", "pyr_pool_kernel", "pyr_coef_kernel", "stream3_kernel<15,false>"],
            "clocks": clocks, "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
