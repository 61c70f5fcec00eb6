#!/usr/bin/env python
"""Benchmark of the hot path: CROG R50 batched referring-grasp inference (img + expression -> 5 maps
-> sigmoid/bicubic -> grasp decode -> Jaccard J@1/J@5), BASELINE.json configs[1]/[2].

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on host cores

One "step" = one pass over one batch of `--batch` synthetic samples per GPU (weak scaling: the batch per
GPU is fixed).  `value` is samples/s with inputs resident in HBM; `e2e` is the same metric through the public
API with pinned HOST buffers (H2D of img/word/GT and D2H of grasps/J flags inside the timed region).
Prints ONE JSON line on rank 0.
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
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from crog_b200 import synth  # noqa: E402

ALG_GFLOP_PER_SAMPLE = {17: 137.56, 20: 137.80}  # SURVEY.md §8(d) / BASELINE.md §3 (2*MAC of the reference forward)
METRIC = "samples/sec (img+expr -> grasps)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def stop(self, window=None):
        """`window` = (t0, t1) host times of the timed region: only samples taken inside it count (the sampler is
        started early because nvidia-smi needs a few hundred ms to come up); without samples inside, the nearest ones."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows
        if window is not None:
            inside = [r for r in rows if window[0] <= r[-1] <= window[1] + 0.06]
            rows = inside if inside else sorted(rows, key=lambda r: min(abs(r[-1] - window[0]), abs(r[-1] - window[1])))[:3]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = sm if window is not None else (sorted(sm)[len(sm) // 2:] if sm else [])  # no window: upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_STATE = {}


def workload_config(B: int, Lw: int, world: int):
    """The `config` object of the JSON line (identical for both arms)."""
    return {"workload": f"CROG R50 (crog_multiple_r50.yaml shapes) batched inference + grasp decode + Jaccard, batch {B}/GPU, "
                        f"416x416, L={Lw}, bf16, random-init (seeded) weights",
            "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": "no explicit flush: one step streams ~6 GB of activations, far above the 126 MB L2"}


def cpu_port_samples_per_sec(word_len: int, n_samples: int, threads: int, fwd_batch: int = 1):
    """The reference's CPU path restated (oracle port): fp32 torch forward on `threads` host threads, sigmoid/bicubic,
    then the serial per-sample decode / Jaccard loop (engine/crog_engine.py:478-527).  `fwd_batch` = samples per forward
    (the reference's test script uses 1, test_crog.py:62; 8 fills the cores better)."""
    from oracle import crog_forward as O
    from oracle import grasp_tail_c as TC

    torch.set_num_threads(threads)
    key = (word_len, n_samples)
    if key not in _CPU_STATE:
        cfg = synth.default_cfg(word_len)
        _CPU_STATE.clear()
        _CPU_STATE[key] = (cfg, synth.make_state_dict(cfg, 0, "perturbed"), synth.make_inputs(n_samples, word_len),
                           synth.make_gt_rects(n_samples, 64, seed=4))
    cfg, sd, (img, word), (gt, cnt) = _CPU_STATE[key]
    O.crog_forward(sd, cfg, img[:1], word[:1])  # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for b in range(0, n_samples, fwd_batch):
        e = min(b + fwd_batch, n_samples)
        maps, _ = O.crog_forward(sd, cfg, img[b:e], word[b:e])
        post = [p.numpy() for p in O.postprocess(maps, (416, 416))]
        for i in range(e - b):  # serial per-sample tail, as in the reference
            TC.tail_batch(post[1][i:i + 1], post[2][i:i + 1], post[3][i:i + 1], post[4][i:i + 1], gt[b + i:b + i + 1], cnt[b + i:b + i + 1])
    dt = time.perf_counter() - t0
    return n_samples / dt, dt


def cpu_baseline(word_len: int, n_samples: int, threads: int):
    """Best of the two ways of driving the CPU path (1 or 8 samples per forward) on a bounded sample."""
    best = None
    for fb in (1, 8):
        v, dt = cpu_port_samples_per_sec(word_len, n_samples, threads, fb)
        if best is None or v > best[0]:
            best = (v, dt, fb)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_samples
    vals = []
    n = max(8, min(n, 16))  # bounded per-step sample: K steps x 16 samples stays within a few minutes
    fb = 8
    for _ in range(max(args.warmup, 1)):
        cpu_port_samples_per_sec(args.word_len, n, threads, fb)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_port_samples_per_sec(args.word_len, n, threads, fb)
        vals.append(v)
    total = time.perf_counter() - t0
    value = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.word_len, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{n} samples per step ({fb} per forward): oracle/crog_forward.py (torch CPU fp32, all host threads) + "
                                   "oracle/grasp_tail.c serial decode/Jaccard loop; the reference itself is Python and cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gen_tail_maps_device(n, kind, seed, dev, size=416):
    """Config-5 maps generated on the device with the distributions of synth.make_tail_maps (bench input only)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    ys = torch.arange(size, device=dev, dtype=torch.float32).view(1, size, 1)
    xs = torch.arange(size, device=dev, dtype=torch.float32).view(1, 1, size)
    q = torch.empty((n, size, size), device=dev)
    for i0 in range(0, n, 256):
        m = min(256, n - i0)
        if kind == "blobs":
            acc = torch.zeros((m, size, size), device=dev)
            nb = torch.randint(1, 9, (m,), generator=g, device=dev)
            for j in range(8):
                a = (torch.rand((m, 1, 1), generator=g, device=dev) * 0.55 + 0.45) * (nb > j).view(m, 1, 1)
                sg = torch.rand((m, 1, 1), generator=g, device=dev) * 7 + 3
                mx = torch.rand((m, 1, 1), generator=g, device=dev) * (size - 20) + 10
                my = torch.rand((m, 1, 1), generator=g, device=dev) * (size - 20) + 10
                acc += a * torch.exp(-((xs - mx) ** 2 + (ys - my) ** 2) / (2 * sg * sg))
            acc += torch.randn((m, size, size), generator=g, device=dev) * 0.01
            q[i0:i0 + m] = acc.clamp_(0, 1)
        else:
            t = torch.rand((m, size, size), generator=g, device=dev)
            idx = torch.arange(i0, i0 + m, device=dev)
            quant = (idx % 20 == 19).view(m, 1, 1)
            q[i0:i0 + m] = torch.where(quant, torch.floor(t * 16) / 16, t)
    ph = torch.rand((n, 1, 1), generator=g, device=dev) * 6.28 - 3.14
    kx = torch.rand((n, 1, 1), generator=g, device=dev) * 0.04 - 0.02
    phi = kx * xs + kx.flip(0) * ys + ph
    s = torch.sin(2 * phi) + torch.randn((n, size, size), generator=g, device=dev) * 0.05
    c = torch.cos(2 * phi) + torch.randn((n, size, size), generator=g, device=dev) * 0.05
    w = 0.5 + 0.5 * torch.sin(0.01 * xs + 0.02 * ys + ph)
    return q, s, c, w.expand(n, size, size).contiguous()


def run_tail(args):
    """BASELINE.json configs[4]: grasp-decode / IoU tail micro-benchmark, 4096 maps x 64 GT rectangles."""
    from crog_b200.utils import grasp_eval as GE
    from oracle import grasp_tail_c as TC

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n, K = args.tail_maps, 5
    gt, cnt = synth.make_gt_rects(n, 64, seed=4)
    d_gt, d_cnt = torch.from_numpy(gt).to(dev), torch.from_numpy(cnt).to(dev)
    _, hbm_peak, which = peaks()
    res = {}
    for kind, seed in (("blobs", 7), ("stress", 8)):
        q, s, c, w = gen_tail_maps_device(n, kind, seed, dev)
        counters = torch.zeros(4, dtype=torch.int64, device=dev)

        def step():
            # sub-batches on two streams: the Jaccard rasterisation of one overlaps the peak scan of the next
            peaks_, npk, grasps, _ = GE.decode_and_score_batched(q, s, c, w, d_gt, d_cnt, K, counters=counters, chunks=args.tail_chunks)
            return peaks_, npk, grasps

        for _ in range(max(args.warmup, 3)):
            out = step()
        torch.cuda.synchronize()
        run = step
        if not args.tail_no_graph and args.tail_chunks == 1:
            # the four dependent launches of one step (scan, select, exact fallback, Jaccard) replayed as a CUDA graph:
            # no host launch gaps between them (the forward is replayed the same way)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = step()
            run = graph.replay
            for _ in range(3):
                run()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # spot parity at full size: 16 maps against the oracle, bit-exact peaks
        pk, npk = out[0].cpu().numpy(), out[1].cpu().numpy()
        ok = True
        for b in list(range(0, n, max(n // 16, 1)))[:16]:
            ref = TC.peak_local_max(q[b].cpu().numpy(), 0.4, K)
            ok = ok and npk[b] == len(ref) and np.array_equal(pk[b, :npk[b]], ref.astype(np.int32))
        alg_bytes = n * (416 * 416 * 4 + K * 3 * 4 + 64 * 6 * 8 + K * 5 * 8 + 2 * 4)  # 695 564 B/sample
        res[kind] = {"ms": ms, "samples_per_s": n / (ms / 1e3), "gbs": alg_bytes / (ms / 1e3) / 1e9, "parity_spot_check": bool(ok)}
        del q, s, c, w
    r = res["blobs"]
    line = {"metric": "samples/sec (grasp decode + Jaccard tail)", "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/int32", "data": "synthetic",
            "config": {"workload": f"tail micro-bench: {n} maps 416x416 (q,sin,cos,wid) x 64 GT rectangles, K=5; 'blobs' distribution "
                                   "(value) and 'stress' (iid uniform + plateaus) below", "l2": "inputs (2.8 GB of quality maps) exceed L2"},
            "roofline": {"bound": "hbm", "kernel": "peak_scan_kernel + peak_select + jaccard", "achieved": r["gbs"], "peak": hbm_peak,
                         "unit": "GB/s", "frac": r["gbs"] / hbm_peak, "peak_source": which, "traffic": None},
            "stress": res["stress"], "blobs": res["blobs"], "gpu_launches": 4 * args.tail_chunks * args.steps}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU per step")
    ap.add_argument("--word-len", type=int, default=17)
    ap.add_argument("--cpu-samples", type=int, default=48, help="bounded CPU-baseline sample (about 10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="forward", choices=["forward", "tail"])
    ap.add_argument("--tail-maps", type=int, default=4096)
    ap.add_argument("--tail-no-graph", action="store_true", help="launch the tail kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--tail-chunks", type=int, default=1, help="sub-batches of the tail (scan of i+1 overlaps Jaccard of i)")
    ap.add_argument("--ncu-range", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    ap.add_argument("--dump-ops", default=None, help="write the per-op CUDA-event table (eager replay) to this file")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "tail":
        return run_tail(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    from crog_b200.engine import GraspEvaluator
    from crog_b200.model import CROG

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, Lw, S = args.batch, args.word_len, 416
    cfg = synth.default_cfg(Lw)
    model = CROG(cfg, precision="bf16")
    model.load_state_dict(synth.make_state_dict(cfg, 0, "perturbed"), strict=True)
    model = model.to(dev)
    # every rank owns a contiguous shard of the global batch (different seeds per rank => different samples)
    img, word = synth.make_inputs(B, Lw, seed_img=1 + 1000 * rank, seed_txt=2 + 1000 * rank)
    gt, cnt = synth.make_gt_rects(B, 64, seed=4 + 1000 * rank)
    h_img, h_word = img.pin_memory(), word.pin_memory()
    h_gt, h_cnt = torch.from_numpy(gt).pin_memory(), torch.from_numpy(cnt).pin_memory()
    d_img, d_word, d_gt0, d_cnt = h_img.to(dev), h_word.to(dev), h_gt.to(dev), h_cnt.to(dev)
    d_gt = d_gt0.clone()
    ev = GraspEvaluator(model, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms.item())

    # ---- device-resident arm
    def step_resident():
        ev.step(d_img, d_word, d_gt, d_cnt)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # early: nvidia-smi takes a few hundred ms to deliver its first sample
    for _ in range(args.warmup):
        step_resident()
    if args.ncu_range:
        torch.cuda.profiler.start()
    t_w0 = time.time()
    ms = timed(step_resident, args.steps)
    t_w1 = time.time()
    if args.ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop((t_w0, t_w1)) if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)
    ev.reduce()
    counters = ev.counters.tolist()

    # ---- end-to-end arm: pinned host buffers in, grasps / flags out, every step
    e2e = None
    if not args.no_e2e:
        # the public streaming API: pinned host batches in, decoded grasps / J flags back on the host after EVERY step;
        # the H2D copy of batch k+1 rides on a copy stream behind the compute of batch k (GraspEvaluator.stream)
        h_batch = (h_img, h_word, h_gt, h_cnt)
        sink = []

        def run_e2e(steps):
            for n_, g_, f_ in ev.stream(h_batch for _ in range(steps)):
                sink.append(int(n_[0]))  # the host touches every step's result

        run_e2e(3)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        torch.cuda.synchronize()
        t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        sync_all()
        ms_e = float(t_ms.item())
        h_out_g, h_out_f, h_out_n = (torch.empty((B, 5, 5), dtype=torch.float64), torch.empty((B, 2), dtype=torch.int32),
                                     torch.empty((B,), dtype=torch.int32))
        h2d = h_img.numel() * 4 + h_word.numel() * 8 + h_gt.numel() * 8 + h_cnt.numel() * 4
        d2h = h_out_g.numel() * 8 + h_out_f.numel() * 4 + h_out_n.numel() * 4
        e2e = {"value": world * B * args.steps / (ms_e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / args.steps,
               "how": "GraspEvaluator.stream: pinned host batch -> H2D on a copy stream (double buffered, overlapped with the "
                      "previous step's kernels) -> forward + glue + decode + Jaccard -> D2H of grasps / counts / J flags, host "
                      "waits for every step's result"}

    # ---- roofline of the dominant kernel family (tcgen05 implicit GEMM): per-op CUDA events on an eager replay
    plan = model.plan_for(B, S)
    roof = op_table = None
    if rank == 0:
        tf_peak, hbm_peak, which = peaks()
        names, durs = plan.op_names, np.zeros(len(plan.ops))
        reps = 3
        from crog_b200 import _lib as L
        for _ in range(reps):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan.ops) + 1)]
            s = L.stream_ptr()
            # keep the GPU busy while the host enqueues the ~220 launches + events (a few ms of ctypes calls), so that the
            # event-to-event durations are kernel times, not host launch gaps (the 10-20 us text-tower kernels otherwise
            # wait for the host)
            torch.cuda._sleep(int(1.2e7))
            evs[0].record()
            for i, fn in enumerate(plan.ops):
                fn(s)
                evs[i + 1].record()
            torch.cuda.synchronize()
            durs += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(len(plan.ops))])
        durs /= reps
        is_gemm = np.array([n in plan.gemm_alg_flops for n in names])
        alg = sum(plan.gemm_alg_flops.values())
        t_gemm = float(durs[is_gemm].sum()) / 1e3
        achieved = alg / t_gemm / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")  # per-launch DRAM bytes of this kernel family (ncu)
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch_avg")
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit GEMM, all conv/linear layers)",
                "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak, "peak_source": which + " sustained",
                "traffic": traffic, "launches": int(is_gemm.sum()), "alg_gflop_per_launch_avg": alg / 1e9 / max(int(is_gemm.sum()), 1),
                "share_of_forward": t_gemm / (float(durs.sum()) / 1e3),
                "autotuned_layers": sum(1 for v in plan.tile_choice.values() if v[0] != 0),
                "whole_step_frac_of_peak": ALG_GFLOP_PER_SAMPLE.get(Lw, 137.56) * 1e9 * (value / world) / 1e12 / tf_peak}
        order = np.argsort(-durs)[:12]
        op_table = [{"op": names[i], "ms": round(float(durs[i]), 4)} for i in order]
        if args.dump_ops:
            with open(args.dump_ops, "w") as f:
                for i, n in enumerate(names):
                    gf = plan.gemm_alg_flops.get(n, 0) / 1e9
                    tc = plan.tile_choice.get(n)
                    tcs = f"  tile_cfg {tc[0]} ({tc[1]:.1f} us vs heuristic {tc[2]:.1f} us)" if tc else ""
                    f.write(f"{i:4d} {n:50s} {durs[i]:9.4f} ms {gf:10.2f} GF {gf / max(durs[i], 1e-9):9.1f} TF/s{tcs}\n")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, fb = cpu_baseline(Lw, args.cpu_samples, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_samples} samples of the same workload, {fb} per forward ({dt:.1f} s; best of 1 / 8 per forward): "
                         "oracle torch-CPU fp32 forward + sigmoid/bicubic + C decode/Jaccard serial loop"}

    if rank == 0:
        launches_per_step = plan.n_launches + 1 + 3 + 1  # forward + sigmoid/bicubic + (scan, select, exact) + jaccard
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": workload_config(B, Lw, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "j_counters": counters, "top_ops_ms": op_table,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
