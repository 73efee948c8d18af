#!/usr/bin/env python
"""bench.py -- FQSS ConvTasNet QAT train throughput (audio-seconds per second) on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one full QAT step of BASELINE.json configs[1]: fake-quantised ConvTasNet student forward,
float-teacher forward, FQSS KD SI-SDR loss, backward, gradient all-reduce, global-norm clip and Adam,
on synthetic 4 s / 8 kHz two-speaker mixtures.  `value` times steps whose inputs are already in HBM;
`e2e` times the same steps fed from pinned HOST buffers (H2D of mixture+sources and D2H of the loss
inside the timed region).  Scaling is weak: the per-GPU batch is fixed (32), global batch = 32*N.

`--impl reference` times the reference's CPU implementation of the same step -- the oracle port
(oracle/fqss_oracle.py, bit-identical to ssi-research/FQSS on CPU; the reference itself is a Python
tree that does not exist on the GPU box) -- on the box's host cores.
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

SEG_SECONDS = 4.0
SAMPLE_RATE = 8000
T = int(SEG_SECONDS * SAMPLE_RATE)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fqss_b200", choices=["fqss_b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=32)
    ap.add_argument("--global-batch", type=int, default=0, help=">0: strong scaling with this global batch")
    ap.add_argument("--workload", default="speech", choices=["speech", "music", "dptnet", "sepformer"],
                    help="speech (default): BASELINE configs[1], the graded line.  music: BASELINE configs[4] (ConvTasNetMusicQ, "
                         "4 stems, stereo 44.1 kHz, 80 000-sample segments, batch 8 per GPU); dptnet / sepformer: BASELINE configs[2] / [3] "
                         "(batch 1, 3 s / 4 s at 8 kHz) -- extra lines, single GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--small", action="store_true", help="reduced model (debug only; never a reported number)")
    ap.add_argument("--breakdown-file", default="", help="write the per-kernel attribution table of one step here")
    ap.add_argument("--cuda-graph", type=int, default=1,
                    help="1 (default): the timed steps replay ONE captured CUDA graph of the whole step "
                         "(fqss_b200.graph.GraphedStep); 0: eager launches")
    ap.add_argument("--verify-dp", type=int, default=-1,
                    help="data-parallel correctness check before the result line: one seeded global batch run sharded over the "
                         "N ranks (global-batch parity switches on) and again on rank 0 alone; reports the relative difference "
                         "of loss and averaged gradient arena.  -1 (default): on when N > 1")
    ap.add_argument("--strong", type=int, default=-1,
                    help="also time BASELINE configs[1] read literally (global batch 32 split over the N GPUs) and add it to "
                         "the line as `strong`.  -1 (default): on when N > 1")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket ONE extra step with cudaProfilerStart/Stop (for `ncu --profile-from-start off`); "
                         "numbers printed by such a run are never bench values")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, smax, reasons = [], 0, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ---------------------------------------------------------------------------------------------
def cpu_step_factory(small=False):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fqss_oracle as O
    from fqss_b200.testing import FULL_KW, SMALL_KW, _oracle_cfg, model_pair, oracle_params
    torch.set_num_threads(os.cpu_count() or 1)
    kw = SMALL_KW if small else FULL_KW
    cfg = _oracle_cfg(kw)
    model, fmodel = model_pair(kw, "cpu", seed=0)
    P, fP = oracle_params(model), oracle_params(fmodel)
    del model, fmodel
    gen = torch.Generator().manual_seed(0)

    def batch(B):
        src = torch.randn(B, 2, T, generator=gen) * 0.05
        return src.sum(1, keepdim=True), src
    mix, src = batch(1)
    st = O.calibrate(P, mix, cfg, passes=2)

    def step(B):
        mix, src = batch(B)
        P.leafify()
        t0 = time.perf_counter()
        O.qat_step(P, fP, mix, src, cfg, st, 0.1)
        return time.perf_counter() - t0
    return step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_step_factory(args.small)
    t_probe = step(1)                                   # also the first warm-up
    total = max(1, args.steps + max(args.warmup - 1, 0))
    B = 1
    for cand in (4, 2):
        if t_probe * cand * total <= 150.0:
            B = cand
            break
    for _ in range(max(args.warmup - 1, 0)):
        step(B)
    times = [step(B) for _ in range(args.steps)]
    dt = sum(times)
    val = B * SEG_SECONDS * args.steps / dt
    line = {"impl": "reference", "metric": "ConvTasNet QAT train audio-sec/sec", "value": val, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, args.per_gpu_batch, max(args.gpus, 1)), reference_sample_batch=B),
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "oracle port of the reference step (bit-identical to ssi-research/FQSS on CPU), "
                                       "batch %d x 4 s per step, %d steps" % (B, args.steps)},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch, world):
    return {"workload": "ConvTasNet 2spk 8 kHz full FQSS QAT (W8A8, splitter/combiner 2/2, KD SI-SDR vs float teacher, "
                        "lambda 0.1), 4 s segments" + (" [REDUCED MODEL - debug]" if args.small else ""),
            "global_batch": per_gpu_batch * world, "per_gpu_batch": per_gpu_batch, "segment_s": SEG_SECONDS,
            "sample_rate": SAMPLE_RATE, "parallelism": "dp%d" % world,
            "l2": "inputs+activations per step (>10 GB) exceed the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fqss_b200 import _native as N
    from fqss_b200.losses import fqss_kd_loss, fqss_training_step
    from fqss_b200.parallel import ParamArena
    from fqss_b200.qat.models.load_model import enable_observer
    from fqss_b200.testing import FULL_KW, SMALL_KW, model_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the fqss_b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.lib()
    B = args.per_gpu_batch if not args.global_batch else args.global_batch // world
    model, fmodel = model_pair(SMALL_KW if args.small else FULL_KW, dev, seed=0)
    gen = torch.Generator().manual_seed(1000 + rank)
    n_host = 2                                              # two pinned host batches, alternated
    host = []
    for _ in range(n_host):
        src = (torch.randn(B, 2, T, generator=gen) * 0.05).pin_memory()
        host.append((src.sum(1, keepdim=True).pin_memory(), src))
    dev_batches = [(m.to(dev), s.to(dev)) for m, s in host]
    # calibration: 2 observer passes, then steady state
    from fqss_b200 import parallel as PL
    with torch.no_grad():
        for _ in range(2):
            model(dev_batches[0][0][: min(B, 4)])
            if world > 1:       # every rank learns the ranges one process would learn from the global calibration batch
                PL.set_global_batch_parity(True)
                PL.sync_observer_ranges_(model)
                PL.set_global_batch_parity(False)
    enable_observer(model, False)
    arena = ParamArena(list(model.parameters()))
    loss_host = torch.zeros(1).pin_memory()

    def step(mix, src):
        arena.zero_grad()
        loss, _, _ = fqss_training_step(model, fmodel, mix, src, 0.1)   # teacher forward on a side stream next to the student's
        loss.backward()
        arena.gather_grads()
        scale = arena.allreduce_mean()
        arena.clip_and_step(pre_scale=scale, max_norm=5.0, lr=1e-3)
        return loss

    graphed, graph_error = None, None
    if args.cuda_graph and not args.profile_step:
        from fqss_b200.graph import GraphedStep
        try:
            graphed = GraphedStep(step, dev_batches[0], warmup=max(args.warmup, 3))
        except Exception as e:      # same kernels either way: fall back to eager launches and say so in the line
            graph_error = repr(e)[:200]
            torch.cuda.synchronize()

    def run_step(mix, src):
        return graphed(mix, src) if graphed is not None else step(mix, src)

    copy_stream = torch.cuda.Stream()
    staging = [tuple(torch.empty_like(t) for t in dev_batches[0]) for _ in range(2)]
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def h2d(i):
        """Enqueue the host->device copy of step i's batch on the copy stream (after the staging slot was consumed)."""
        m, s = host[i % n_host]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            staging[i % 2][0].copy_(m, non_blocking=True)
            staging[i % 2][1].copy_(s, non_blocking=True)
            staged[i % 2].record(copy_stream)

    def timed(n, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        if from_host and graphed is not None:
            # prefetching loader: the H2D copy of step i+1 (pinned host -> staging buffer, copy stream) runs under the
            # compute of step i; each step then moves its staged batch into the graph's static inputs (12 MB device copy)
            h2d(0)
        for i in range(n):
            if from_host and graphed is not None:
                cur = torch.cuda.current_stream()
                cur.wait_event(staged[i % 2])                       # this step's inputs have arrived
                mix, src = graphed.static_in
                mix.copy_(staging[i % 2][0], non_blocking=True)
                src.copy_(staging[i % 2][1], non_blocking=True)
                consumed[i % 2].record(cur)
                if i + 1 < n:
                    h2d(i + 1)
            elif from_host:
                m, s = host[i % n_host]
                mix, src = m.to(dev, non_blocking=True), s.to(dev, non_blocking=True)
            else:
                mix, src = dev_batches[i % n_host]
            loss = run_step(mix, src)
            if from_host:
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, t0, time.time()

    timed(max(args.warmup, 3), False)
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(*dev_batches[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    from fqss_b200 import roofline as R
    # the end-to-end loop (the headline number) is timed first, the HBM-resident loop second: the board heats up over the first
    # seconds of load and the second loop runs ~1.3 % slower whatever it is (measured by alternating the two loops:
    # FQSS_BENCH_DRIFT=1); the copies themselves cost ~0.1 ms per step
    ms_e2e, t0, _ = timed(args.steps, True)
    c0 = R.launch_count()
    ms, _, t2 = timed(args.steps, False)
    t1 = t2
    launches = R.launch_count() - c0          # kernels launched by libfqss_sm100 inside the timed region
    if graphed is not None:                   # replays launch the captured nodes without re-entering the library
        launches = graphed.kernels_per_replay * args.steps
    drift = None
    if os.environ.get("FQSS_BENCH_DRIFT"):      # diagnostic: alternate the two loops again (clock / thermal drift vs a real e2e cost)
        drift = [timed(args.steps, False)[0] / args.steps, timed(args.steps, True)[0] / args.steps,
                 timed(args.steps, False)[0] / args.steps, timed(args.steps, True)[0] / args.steps]
        if rank == 0:
            print("drift check (ms/step: resident, e2e, resident, e2e):", drift, flush=True)
    clocks = sampler.stop(t0, t2) if rank == 0 else None
    final_loss = float(loss_host.item())

    # ---- BASELINE configs[1] read literally: global batch 32 split over the N GPUs (strong scaling) ----
    strong = None
    want_strong = (args.strong == 1 or (args.strong < 0 and world > 1)) and not args.global_batch and not args.profile_step
    if want_strong and 32 % world == 0:
        Bs = 32 // world
        sb = [(m[:Bs].contiguous(), s_[:Bs].contiguous()) for m, s_ in dev_batches]
        g2 = None
        try:
            if args.cuda_graph:
                from fqss_b200.graph import GraphedStep
                g2 = GraphedStep(step, sb[0], warmup=3)
        except Exception:
            g2 = None
            torch.cuda.synchronize()
        run2 = (lambda m, s_: g2(m, s_)) if g2 is not None else step

        def timed2(n):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                run2(*sb[i % n_host])
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        timed2(3)
        ms_s = timed2(args.steps)
        strong = {"global_batch": 32, "per_gpu_batch": Bs, "ms_per_step": ms_s / args.steps,
                  "value": 32 * SEG_SECONDS * args.steps / (ms_s / 1e3), "unit": "audio-s/s", "scaling": "strong",
                  "launch_mode": "cuda graph" if g2 is not None else "eager"}
        if g2 is not None:
            g2.graph.reset()
            g2 = None

    # ---- data-parallel correctness: the same seeded global batch on N ranks and on rank 0 alone ----
    dp_check = None
    if (args.verify_dp == 1 or (args.verify_dp < 0 and world > 1)) and world > 1 and not args.profile_step:
        Gv = 8 if 8 % world == 0 else world
        g = torch.Generator().manual_seed(4242)
        src_g = (torch.randn(Gv, 2, T, generator=g) * 0.05).to(dev)
        mix_g = src_g.sum(1, keepdim=True)
        lo, hi = PL.shard_bounds(Gv, rank, world)

        def fwd_bwd(mix, src):
            arena.zero_grad()
            est = model(mix)
            with torch.no_grad():
                fest = fmodel(mix)
            loss, _, _ = fqss_kd_loss(est, fest, src, 0.1)
            loss.backward()
            arena.gather_grads()
            return loss.detach().clone()
        PL.set_global_batch_parity(True)           # splitter peak (MAX), loss means (SUM): global-batch semantics
        try:
            loss_dp = fwd_bwd(mix_g[lo:hi].contiguous(), src_g[lo:hi].contiguous())
            scale = arena.allreduce_mean()
            g_dp = arena.grad.clone().mul_(scale)
        finally:
            PL.set_global_batch_parity(False)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            loss_1 = fwd_bwd(mix_g, src_g)         # one process, the whole global batch, no collective
            g_1 = arena.grad
            num = (g_dp - g_1).double().norm().item()
            den = g_1.double().norm().item()
            dp_check = {"global_batch": Gv, "ranks": world, "loss_dp": float(loss_dp), "loss_single": float(loss_1),
                        "loss_rel_diff": abs(float(loss_dp) - float(loss_1)) / max(abs(float(loss_1)), 1e-30),
                        "grad_rel_l2_diff": num / max(den, 1e-30),
                        "grad_max_abs_diff_over_max": ((g_dp - g_1).abs().max() / g_1.abs().max().clamp_min(1e-30)).item(),
                        "what": "all-reduced mean gradient arena (5.1 M floats) and loss of one seeded global batch on %d ranks "
                                "vs the same batch on rank 0 alone; splitter peak and loss means synchronised "
                                "(fqss_b200.parallel.set_global_batch_parity)" % world}
        torch.cuda.synchronize()
        dist.barrier()

    def finish():
        """Multi-rank teardown: every captured graph is destroyed BEFORE the process group (a live graph that captured
        NCCL's all-reduce keeps the communicator busy and destroy_process_group then waits forever); a watchdog ends the
        process if the communicator teardown still does not return."""
        nonlocal graphed
        sys.stdout.flush()
        if world > 1:
            if graphed is not None:
                graphed.graph.reset()
            graphed = None
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            done = threading.Event()

            def watchdog():
                if not done.wait(30.0):
                    os._exit(0)         # result line is out and every rank is past its last collective
            threading.Thread(target=watchdog, daemon=True).start()
            dist.destroy_process_group()
            done.set()

    if rank != 0:
        finish()
        return
    gb = B * world
    val = gb * SEG_SECONDS * args.steps / (ms / 1e3)
    val_e2e = gb * SEG_SECONDS * args.steps / (ms_e2e / 1e3)
    line = {"metric": "ConvTasNet QAT train audio-sec/sec", "value": val, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
            "dtype": "f32 (8-bit fake-quant codes, fp32 accumulate)", "data": "synthetic",
            "config": workload_config(args, B, world),
            "e2e": {"value": val_e2e, "unit": "audio-s/s", "h2d_bytes_per_step": B * 3 * T * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss,
            "launch_mode": ("one CUDA graph per step (%d library kernels per replay)" % graphed.kernels_per_replay)
                           if graphed is not None else ("eager" + (" (graph capture failed: %s)" % graph_error if graph_error else ""))}
    if strong is not None:
        line["strong"] = strong
    elif world == 1 and not args.global_batch and B == 32:
        line["strong"] = {"global_batch": 32, "per_gpu_batch": 32, "ms_per_step": ms / args.steps, "value": val, "unit": "audio-s/s",
                          "scaling": "strong", "note": "N = 1: identical to the weak-scaling configuration"}
    if dp_check is not None:
        line["dp_check"] = dp_check
    if world == 1 and not args.no_roofline:
        try:
            Mfr = (T - 16) // 8 + 1
            line["roofline"], line["kernels"] = R.step_roofline(lambda: step(*dev_batches[0]), B, Mfr, ms / args.steps)
            if args.breakdown_file:
                with open(args.breakdown_file, "w") as f:
                    f.write(R.all_kernel_fractions(lambda: step(*dev_batches[0]), B, Mfr) + "\n")
        except Exception as e:      # never lose the bench line over the side measurement
            line["roofline"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        del model, fmodel, arena
        torch.cuda.empty_cache()
        cstep, cores = cpu_step_factory(args.small)
        cstep(1)
        Bc = 2
        ts = [cstep(Bc) for _ in range(2)]
        cv = Bc * SEG_SECONDS * len(ts) / sum(ts)
        line["cpu_baseline"] = {"value": cv, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                "sample": "oracle port of the reference QAT step on host cores: batch %d x 4 s, %d timed steps "
                                          "after 1 warm-up (%.1f s/step)" % (Bc, len(ts), sum(ts) / len(ts))}
    print(json.dumps(line), flush=True)
    finish()


# ---------------------------------------------------------------------------------------------
# extra workload: BASELINE configs[4], the music recipe (configs/convtasnet_music.yaml, musdbhq_train.py:45-137)
# ---------------------------------------------------------------------------------------------
def run_music(args):
    """One training step of the music recipe per timed step: fake-quantised ConvTasNetMusicQ forward (4 stems, stereo,
    40 blocks, dilation <= 512, 7 999 frames), float teacher forward, centre trim, L1 + new-SDR-weighted KD loss, backward,
    Adam (lr 1e-4, no clipping: musdbhq_train.py:120-128 computes the norm for display only).  Everything on libfqss_sm100
    kernels (the teacher is the same module tree with quantisation disabled, so it runs on the same wrappers)."""
    import copy
    import torch
    from fqss_b200 import _native as N
    from fqss_b200 import roofline as R
    from fqss_b200.losses import music_training_step
    from fqss_b200.parallel import ParamArena
    from fqss_b200.qat.models.convtasnetq_music import ConvTasNetMusicQ
    from fqss_b200.qat.models.load_model import enable_observer, quantize_model
    from fqss_b200.testing import RECIPE_QUANT
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the fqss_b200 arm has no CPU fallback)")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --workload music is a single-GPU line")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    N.lib()
    B = 8 if args.per_gpu_batch == 32 else args.per_gpu_batch        # recipe batch (convtasnet_music.yaml: batch_size 8)
    Tm, SR = 80000, 44100
    torch.manual_seed(0)
    kw = dict(n_repeats=1, n_blocks=3, n_filters=64, bn_chan=64, hid_chan=128) if args.small else {}
    model = ConvTasNetMusicQ(**kw)
    fmodel = quantize_model(copy.deepcopy(model), dict(RECIPE_QUANT, weight_quant=False, act_quant=False, out_quant=False,
                                                       n_splitter=1, n_combiner=1)).to(dev)
    model = quantize_model(model, dict(RECIPE_QUANT)).to(dev)
    for p in fmodel.parameters():
        p.requires_grad_(False)
    gen = torch.Generator().manual_seed(7)
    host = []
    for _ in range(2):
        src = (torch.randn(B, 4, 2, Tm, generator=gen) * 0.1).pin_memory()
        host.append((src.sum(1).pin_memory(), src))
    dev_batches = [(m.to(dev), s.to(dev)) for m, s in host]
    with torch.no_grad():
        for _ in range(2):
            model(dev_batches[0][0][:2])
            fmodel(dev_batches[0][0][:1])
    enable_observer(model, False)
    enable_observer(fmodel, False)
    arena = ParamArena(list(model.parameters()))

    def step(mix, src):
        arena.zero_grad()
        loss, _, _, _ = music_training_step(model, fmodel, mix, src, 0.1)
        loss.backward()
        arena.gather_grads()
        arena.clip_and_step(pre_scale=1.0, max_norm=float("inf"), lr=1e-4)
        return loss

    W = max(args.warmup, 3)
    for i in range(W):
        step(*dev_batches[i % 2])
    torch.cuda.synchronize()
    # ~12 000 small launches per step: the eager step is bound by host-side launch work, so replay it as ONE CUDA graph
    graphed, graph_error = None, None
    if args.cuda_graph:
        from fqss_b200.graph import GraphedStep
        try:
            graphed = GraphedStep(step, dev_batches[0], warmup=2)
        except Exception as e:
            graph_error = repr(e)[:200]
            torch.cuda.synchronize()
    run = (lambda m, s_: graphed(m, s_)) if graphed is not None else step
    c0 = R.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    tw0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(*dev_batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = graphed.kernels_per_replay * args.steps if graphed is not None else R.launch_count() - c0
    # end to end: the same steps fed from pinned host memory, loss read back every step
    loss_host = torch.zeros(1).pin_memory()
    e0.record()
    for i in range(args.steps):
        m, s_ = host[i % 2]
        loss = run(m.to(dev, non_blocking=True), s_.to(dev, non_blocking=True))
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop(tw0, tw1)
    secs = B * Tm / SR
    line = {"metric": "ConvTasNetMusic QAT train audio-sec/sec", "value": secs * args.steps / (ms / 1e3), "unit": "audio-s/s",
            "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (8-bit fake-quant codes)", "data": "synthetic",
            "config": {"workload": "ConvTasNet music 4-stem 44.1 kHz QAT (W8A8, splitter/combiner 2/2, L1 + new-SDR-weighted KD vs float "
                                   "teacher, lambda 0.1) on synthetic stereo 80 000-sample segments", "global_batch": B, "per_gpu_batch": B,
                       "segment_samples": Tm, "sample_rate": SR, "parallelism": "dp1", "small_debug_model": bool(args.small),
                       "l2": "activations per step exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": secs * args.steps / (ms_e2e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": B * 5 * 2 * Tm * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "final_loss": float(loss_host.item()),
            "launch_mode": ("one CUDA graph per step" if graphed is not None else "eager (graph capture failed: %s)" % graph_error
                            if args.cuda_graph else "eager") + " (per-layer wrappers; the fused TCN engine covers the speech model's blocks only)",
            "note": "extra line (BASELINE configs[4]); the graded line is the default speech workload"}
    if not args.no_roofline:
        try:
            prof = R.profile(lambda: step(*dev_batches[0]), 1)
            tot = sum(v["ms"] for v in prof.values())
            line["kernels"] = [{"kernel": k, "ms_per_step": round(v["ms"], 3), "launches_per_step": v["kernels"],
                                "share": round(v["ms"] / tot, 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:int(os.environ.get("FQSS_BENCH_TOPK", "10"))]]
            # HBM fraction of the fused row kernels at this workload's geometry (M = 7 999 frames, dilation <= 512: the row
            # plus its halo takes 38-42 KB of shared memory per CTA).  The byte formulas of the hidden-width kernels are those
            # of the speech blocks (roofline.algorithmic_bytes); the skip-less tail kernel moves the residual branch only.
            if not args.small:
                sep = model.separator.network
                Cio, Chid = sep[1].conv1d.out_channels, sep[2][0][0].net[0].conv1d.out_channels
                Mf = (Tm - model.kernel) // model.stride + 1
                table = R.algorithmic_bytes(B, Mf, Cio, Chid)
                table["tcn_tail_bwd"] = 4 * B * Cio * Mf * 3 + 4 * B * Cio * Mf + 2 * B * Cio * Mf      # g_x, res_y, x_in -> g_xd, dY2
                peak, how = R.measured_peaks()
                rows = {}
                for k in ("tcn_gln2_dw_bwd", "tcn_tail_bwd", "tcn_gln1_bwd", "tcn_dw_fwd", "tcn_dw_fwd(float)", "tcn_gln2_sums", "tcn_hidden_fq"):
                    v = prof.get(k)
                    if v and v["scopes"]:
                        us = v["ms"] / v["scopes"] * 1e3
                        rows[k] = {"avg_launch_us": round(us, 1), "frac": round(table[k] / (us * 1e-6) / 1e9 / peak, 3)}
                line["row_kernel_hbm_frac"] = {"peak_gbs": peak, "peak_source": how, "frames": Mf, "kernels": rows}
        except Exception as e:
            line["kernels"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# extra workloads: BASELINE configs[2] / [3], the sequence models (configs/dptnet_2spks_8k.yaml, sepformer_2spks_8k.yaml)
# ---------------------------------------------------------------------------------------------
def run_seq(args):
    """One QAT step of DPTNetQ / SepformerQ per timed step: fake-quantised student forward, float teacher forward, the FQSS
    KD SI-SDR loss (fused kernel), backward, global-norm clip + Adam on the flat arena.  Quantisers, filterbanks, norms and
    gates run on libfqss_sm100; attention / LSTM / Linear float math on torch (qat_layers_seq.py)."""
    import copy
    import torch
    from fqss_b200 import _native as N
    from fqss_b200 import roofline as R
    from fqss_b200.losses import fqss_training_step
    from fqss_b200.parallel import ParamArena
    from fqss_b200.qat.models.load_model import create_model, enable_observer, quantize_model
    from fqss_b200.testing import RECIPE_QUANT
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the fqss_b200 arm has no CPU fallback)")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --workload %s is a single-GPU line" % args.workload)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    N.lib()
    name, secs_seg, lr = ("DPTNet", 3.0, 4e-4) if args.workload == "dptnet" else ("Sepformer", 4.0, 1.5e-4)
    B = 1 if args.per_gpu_batch == 32 else args.per_gpu_batch          # recipe batch size: 1 (both YAMLs)
    Ts = int(secs_seg * SAMPLE_RATE)
    torch.manual_seed(0)
    cfg = dict(name=name, n_src=2, kernel_size=2) if name == "DPTNet" else dict(name=name, n_src=2, kernel_size=16, stride=8)
    model = create_model(cfg)
    fmodel = copy.deepcopy(model).to(dev)
    model = quantize_model(model, dict(RECIPE_QUANT)).to(dev)
    for p in fmodel.parameters():
        p.requires_grad_(False)
    gen = torch.Generator().manual_seed(7)
    host = []
    for _ in range(2):
        src = (torch.randn(B, 2, Ts, generator=gen) * 0.05).pin_memory()
        host.append((src.sum(1, keepdim=True).pin_memory(), src))
    dev_batches = [(m.to(dev), s_.to(dev)) for m, s_ in host]
    with torch.no_grad():
        model(dev_batches[0][0]); model(dev_batches[0][0])
    enable_observer(model, False)
    arena = ParamArena(list(model.parameters()))

    def step(mix, src):
        arena.zero_grad()
        loss, _, est = fqss_training_step(model, fmodel, mix, src[..., :], 0.1, overlap_teacher=False)
        loss.backward()
        arena.gather_grads()
        arena.clip_and_step(pre_scale=1.0, max_norm=5.0, lr=lr)
        return loss

    # the estimate is a few samples shorter / longer than the segment for some geometries: trim the targets once
    with torch.no_grad():
        Te = model(dev_batches[0][0]).shape[-1]
    dev_batches = [(m, s_[..., :Te].contiguous()) for m, s_ in dev_batches]
    W = max(args.warmup, 3)
    for i in range(W):
        step(*dev_batches[i % 2])
    torch.cuda.synchronize()
    # the eager step is bound by host-side launch work (thousands of small kernels per step): replay it as ONE CUDA graph
    graphed, graph_error = None, None
    if args.cuda_graph:
        from fqss_b200.graph import GraphedStep
        try:
            graphed = GraphedStep(step, dev_batches[0], warmup=2)
        except Exception as e:
            graph_error = repr(e)[:200]
            torch.cuda.synchronize()
    run = (lambda m, s_: graphed(m, s_)) if graphed is not None else step
    c0 = R.launch_count()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    tw0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = run(*dev_batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = graphed.kernels_per_replay * args.steps if graphed is not None else R.launch_count() - c0
    loss_host = torch.zeros(1).pin_memory()
    e0.record()
    for i in range(args.steps):
        m, s_ = host[i % 2]
        loss = run(m.to(dev, non_blocking=True), s_.to(dev, non_blocking=True)[..., :Te].contiguous())
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop(tw0, tw1)
    secs = B * secs_seg
    line = {"metric": "%s QAT train audio-sec/sec" % name, "value": secs * args.steps / (ms / 1e3), "unit": "audio-s/s",
            "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (8-bit fake-quant codes)", "data": "synthetic",
            "config": {"workload": "%s 2spk 8 kHz QAT (W8A8, splitter/combiner 2/2, KD SI-SDR vs float teacher, lambda 0.1), "
                                   "%.0f s segments" % (name, secs_seg), "global_batch": B, "per_gpu_batch": B,
                       "segment_s": secs_seg, "sample_rate": SAMPLE_RATE, "parallelism": "dp1",
                       "l2": "no explicit flush (extra line, not the graded workload)"},
            "e2e": {"value": secs * args.steps / (ms_e2e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": B * 3 * Ts * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "final_loss": float(loss_host.item()),
            "launch_mode": ("one CUDA graph per step" if graphed is not None else "eager (graph capture failed: %s)" % graph_error
                            if args.cuda_graph else "eager") + "; library kernels for quantisers / filterbanks / norms / gates / "
                           "LSTM recurrence / loss / optimizer, torch for attention and Linear float math",
            "note": "extra line (BASELINE configs[%d]); the graded line is the default speech workload" % (2 if name == "DPTNet" else 3)}
    if not args.no_roofline:
        try:
            prof = R.profile(lambda: step(*dev_batches[0]), 1)
            tot = sum(v["ms"] for v in prof.values())
            line["library_kernel_ms_per_step"] = round(tot, 3)
            line["kernels"] = [{"kernel": k, "ms_per_step": round(v["ms"], 3), "launches_per_step": v["kernels"]}
                               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]]
        except Exception as e:
            line["kernels"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "music":
        run_music(a)
    elif a.workload in ("dptnet", "sepformer"):
        run_seq(a)
    else:
        run_ours(a)
