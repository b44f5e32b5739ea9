#!/usr/bin/env python
"""Benchmark of the v5 hot path: MAML meta-steps/s (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload: per GPU 15 synthetic 21x21 region tasks (441 nodes, k = 8 kNN,
600 windows each), one meta-step = every task runs 3 inner SGD steps on support windows 0, 1, 2
plus the first query window (60 window forward+backward passes, 45 clip+SGD steps), the query
gradients are summed (one NCCL all-reduce when N > 1) and applied by one clip+AdamW step.
Scaling is weak: 15 tasks per GPU, so `value` counts 15-task meta-steps per second over the job.

One JSON line on stdout (rank 0); see the contract in the task statement.  `value`: features
resident in HBM, CUDA-graphed meta-step, device-timed.  `e2e`: same loop with the features in
pinned HOST memory -- every step uploads the rows its windows read and reads the loss back.
`roofline`: the dominant kernel family of an instrumented (un-graphed) step, algorithmic bytes
per launch / its CUDA-event time.  `cpu_baseline` / `--impl reference`: the oracle port with the
reference's execution shape (per-node nn.LSTM loop) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TASKS_PER_GPU = 15
NLAT = NLON = 21
KNN = 8
WINDOWS = 600
SUPPORT_ROWS = (0, 1, 2)
METRIC = "maml_meta_steps_per_sec"
UNIT = "meta-steps/s"
# what the path computes in: f32 storage and accumulation; every tensor-core product is a 3-term split of its operands into
# 16-bit hi/lo pairs (fp16 forward, bf16 where an operand is a gradient) -- FP32-class accuracy, not an FP32 FMA pipeline
DTYPE = "f32 (storage + accumulate; products = 3x 16-bit hi/lo operand splits: fp16 forward, bf16 backward)"


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's version banner, build
    messages, library chatter) is sent to stderr from here on."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_config(n_gpus):
    return {
        "workload": "config[1]: MAML meta-train step, 15 synthetic region tasks/GPU (441 nodes, k=8, ~600 windows, "
                    "75/25 support/query), 3 inner SGD steps + 1 query pass per task, outer AdamW",
        "tasks_per_gpu": TASKS_PER_GPU, "global_tasks": TASKS_PER_GPU * n_gpus, "nodes": NLAT * NLON, "k": KNN,
        "window": 24, "horizon": 8, "inner_steps": len(SUPPORT_ROWS), "window_passes_per_meta_step": 60 * n_gpus,
        "parallelism": f"task-sharded dp{n_gpus}, 1 all-reduce of 2.43 MB per meta-step" if n_gpus > 1 else "single GPU",
        "l2_policy": "working set per step (~4 GB of activations) exceeds the 126 MB L2; no explicit flush",
        "dropout": "off (parity configuration, SURVEY.md D11)",
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    """Sustained dense bf16 TFLOP/s (the step runs for many milliseconds): measured, else the recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0.0))) or None
    return 1500.0


# ----------------------------------------------------------------------------- CPU reference arm
def reference_window_pass_seconds(steps, warmup):
    """The UNMODIFIED reference on the host cores (oracle/build_ref.py: its own files over stand-ins for
    torch_geometric / xarray): graphBuilder.build_spatial_graph, dataset.WeatherGraphDataset, model.STGCN +
    hybrid_model.HybridSTGCN_LSTM, and one step of its inner loop per timed pass -- zero_grad, forward, nn.MSELoss,
    backward, clip_grad_norm_(1.0), SGD(lr=0.01).step() (train_hybrid_maml_v5.py:129-139) -- in train mode with the
    dropout probabilities constructed as 0, the same deterministic configuration the GPU arm's headline runs.
    Returns (median seconds per window pass, threads used) or None if the reference is not staged."""
    import contextlib
    import io

    import torch

    from oracle import build_ref
    from weatherforecast_stgcn_maml_b200 import synth

    if not build_ref.available():
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_ref.reference_modules(("graphBuilder", "model", "hybrid_model", "dataset"))
    lats, lons, feats, _ = synth.synth_task(0, num_windows=8, nlat=NLAT, nlon=NLON)
    with contextlib.redirect_stdout(io.StringIO()):
        edge_index, _, _ = m["graphBuilder"].build_spatial_graph(synth.GridCoords(lats, lons), k_neighbors=KNN)
    base = m["model"].STGCN(in_channels=24, hidden_channels=256, out_channels=12, window_size=24, forecast_horizon=8,
                            dropout_rate=0.0)
    hyb = m["hybrid_model"].HybridSTGCN_LSTM(base_stgcn=base, lstm_hidden_size=128, lstm_num_layers=4, lstm_dropout=0.0,
                                             out_channels=12, forecast_horizon=8, freeze_base=False)
    hyb.load_state_dict(synth.init_v5_state_dict(42))
    hyb.train()
    ds = m["dataset"].WeatherGraphDataset(feats, edge_index, window_size=24, forecast_horizon=8)
    opt = torch.optim.SGD(hyb.parameters(), lr=0.01)
    crit = torch.nn.MSELoss()
    times = []
    for it in range(warmup + steps):
        batch = ds[it % 4]
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = crit(hyb(batch.x, batch.edge_index), batch.y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(hyb.parameters(), max_norm=1.0)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return statistics.median(times), cores


def port_window_pass_seconds(steps, warmup, literal=True):
    """The oracle port's window pass (oracle/ref_port.py), kept beside the reference number: literal = the reference's
    execution shape (one nn.LSTM call per node), else the batched restatement the parity tests use."""
    import torch

    from oracle import ref_port as P
    from weatherforecast_stgcn_maml_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lats, lons, feats, _ = synth.synth_task(0, num_windows=8, nlat=NLAT, nlon=NLON)
    ei = P.knn_edges_ckdtree(lats, lons, KNN)
    sd = synth.init_v5_state_dict(42)
    times = []
    if literal:
        fwd, params = P.build_reference_like_module(sd, 24, 8, 4)
        opt = torch.optim.SGD(params, lr=0.01)
    for it in range(warmup + steps):
        x, y = P.window_xy(feats, it % 4, 24, 8)
        t0 = time.perf_counter()
        if literal:
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(fwd(x, ei), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        else:
            _, grads, _ = P.loss_and_grads(sd, x, y, ei, 24, 8, 1.0, 4)
            P.clip_grad_norm([g.clone() for g in grads.values()], 1.0)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return statistics.median(times), cores


def cpu_window_pass(steps, warmup):
    """(seconds per window pass, cores, kind): the staged reference if present, else the port with its execution shape."""
    r = reference_window_pass_seconds(steps, warmup)
    if r is not None:
        return r[0], r[1], "reference"
    sec, cores = port_window_pass_seconds(steps, warmup, literal=True)
    return sec, cores, "port"


CPU_SAMPLE = ("one window pass per step -- zero_grad, forward, MSE, backward, clip_grad_norm_, SGD.step of the reference's "
              "inner loop (train_hybrid_maml_v5.py:129-139, per-node nn.LSTM loop of hybrid_model.py:93-105) on one 441-node, "
              "k=8 window; a meta-step is 60 such passes per 15 tasks, strictly serial in the reference (:124-127,151), so "
              "the metric is extrapolated linearly")


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (oracle/_ref when
    staged: kind "reference"; else the oracle port: kind "port").  A step is a bounded sample of the workload."""
    if rank != 0:
        return
    sec, cores, kind = cpu_window_pass(args.steps, args.warmup)
    passes = 60 * args.gpus
    value = args.gpus / (sec * passes)  # 15-task meta-steps per second for the whole (N x 15)-task job
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": CPU_SAMPLE,
                         "sec_per_window_pass": sec},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = str(gpu_index), None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8 or c[0] != self.idx:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def build_tasks(rank, world):
    import torch  # noqa: F401

    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.dist import shard_tasks
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device

    tasks = []
    for t in shard_tasks(TASKS_PER_GPU * world, rank, world):
        lats, lons, feats, _ = synth.synth_task(t, num_windows=WINDOWS, nlat=NLAT, nlon=NLON)
        ei = knn_edge_index_device(lats, lons, KNN, "cuda")
        tasks.append((feats, ei))
    return tasks


def timed_steps(trainer, steps, warmup, world, sync_loss):
    """W untimed + K timed meta-steps; barrier + synchronize on both sides; device time, max over ranks."""
    import torch
    import torch.distributed as dist

    for _ in range(warmup):
        loss = trainer.meta_step()
        if sync_loss:
            loss.item()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(steps):
        loss = trainer.meta_step()
        if sync_loss:
            last = loss.item()  # the step's result read back to the host (reference: .item() at :170)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), (last if sync_loss else float(loss.item()))


def stage_breakdown(trainer, reps=3):
    """Un-graphed instrumented meta-steps: CUDA-event time of every launcher family."""
    import torch

    from weatherforecast_stgcn_maml_b200 import _lib

    acc, counts = {}, {}
    orig = _lib.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        if name == "wf_gcn_layer_fwd_ss":  # the 24-channel first layer and the three 256 -> 256 layers separately
            name += "_wide" if a[15] >= 128 else "_first"
        acc.setdefault(name, []).append((e0, e1))

    was, was_dist = trainer.use_graph, trainer.dist
    trainer.use_graph = False
    trainer.dist = None  # rank 0 only: an instrumented step must not enter a collective the other ranks do not
    for mod in (sys.modules["weatherforecast_stgcn_maml_b200.engine"],
                sys.modules["weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5"]):
        mod._lib.call = timed_call
    try:
        for _ in range(reps):
            trainer.meta_step()
        torch.cuda.synchronize()
    finally:
        _lib.call = orig
        trainer.use_graph, trainer.dist = was, was_dist
    out = {}
    for name, evs in acc.items():
        out[name] = {"ms_per_meta_step": sum(a.elapsed_time(b) for a, b in evs) / reps, "calls": len(evs) // reps}
    return out


def load_ncu_traffic():
    """{kernel name: dram__bytes_read.sum + dram__bytes_write.sum per launch} parsed from the newest committed
    `profiles/*_ncu_traffic.csv` (written by tools/ncu_traffic.py from an `ncu --set full` capture of this command);
    empty if no capture of the current kernels is committed -- `traffic` is then null, never a stale constant."""
    import csv
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.csv")))
    if not files:
        return {}, None
    out = {}
    with open(files[-1]) as fh:
        for row in csv.DictReader(fh):
            try:
                per = (float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])) / max(1.0, float(row["launches"]))
                out[row["kernel"]] = per
                out.setdefault(row["kernel"].split("<")[0], per)  # "wf_lstm_seq_fwd16_kernel<0>" -> the plain name
            except (KeyError, ValueError):
                continue
    return out, os.path.relpath(files[-1], ROOT)


def time_recurrence_kernels(trainer, reps=10):
    """CUDA-event time of ONE launch of each persistent LSTM kernel (layer 1: 128-wide input, as 3 of 4 layers) on the
    trainer's own buffers, on the stream the kernels are launched on."""
    import torch

    from weatherforecast_stgcn_maml_b200 import _lib

    e, d = trainer.engine, trainer.engine.dims
    Ls, L, st = d.lstm_layers, d.lstm_hidden, _lib.stream_ptr()
    out = {}

    def fwd():
        _lib.call("wf_lstm_seq_recur_fwd", _lib.ptr(e.gates[1]), _lib.ptr(e.c[1]), _lib.ptr(e.h16[1]), _lib.ptr(e.hb16[1]), None,
                  _lib.ptr(e.w16[0]), _lib.ptr(e.w16[1]), 1, Ls, L, d.window, d.num_nodes, e.G, e.Bw, _lib.ptr(e.err), st)

    def bwd():
        _lib.call("wf_lstm_seq_recur_bwd", _lib.ptr(e.gates[1]), _lib.ptr(e.c[1]), _lib.ptr(e.dg16), _lib.ptr(e.ws), 0,
                  _lib.ptr(e.w16[2]), _lib.ptr(e.w16[3]), 1, Ls, L, d.window, d.num_nodes, e.G, e.Bw, _lib.ptr(e.err), st)

    e.ws.zero_()  # dh from the layer above: zeros (timing does not depend on values)
    for name, fn in (("wf_lstm_seq_fwd16_kernel", fwd), ("wf_lstm_seq_bwd_kernel", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / reps
    e.check()
    return out


# Bytes per (row, step) of the two persistent LSTM kernels, L = 128 hidden units, f32 unless noted.
#   minimum = what ANY implementation of the step must move: forward writes the gates (4L), c (L) and h (L)
#             BPTT needs and reads its input x_t (L, for 3 of 4 layers); backward reads gates (4L), c[t] (L; c[t-1] is the
#             previous step's line), dh from above (L) and writes dG (4L).
#   design  = what THIS design moves: the input projection is a separate GEMM, so the forward reads 4L of pre-activations
#             instead of L of x_t, and h leaves twice as 16-bit hi/lo planes (fp16 for the next layer's projection, bf16 for
#             the weight gradients: 2 x L x 4 bytes); the backward reads c[t-1] again (dG leaves as bf16 hi/lo planes: the
#             same 4L x 4 bytes; no transposed copies since round 2).
KERNEL_BYTES = {
    "wf_lstm_seq_fwd16_kernel": {"minimum": (4 * 128 + 128 + 128) * 4 + 128 * 4,
                                 "design": 4 * 128 * 4 + (4 * 128 + 128 + 2 * 128) * 4},
    "wf_lstm_seq_bwd_kernel": {"minimum": (4 * 128 + 128 + 128) * 4 + 4 * 128 * 4,
                               "design": (4 * 128 + 2 * 128 + 128) * 4 + 4 * 128 * 4},
}


def roofline_report(stages, kernel_ms, G, peak, peak_src, ms_per_step, tf_peak):
    """Roofline of the dominant kernel (the slower of the two persistent LSTM kernels; both are reported).
    achieved = ALGORITHMIC bytes per launch (the minimum any implementation must move) / the kernel's
    CUDA-event time; the bytes this design actually asks for are reported beside it as design_bytes."""
    N, T, F, L = NLAT * NLON, 24, 256, 128
    R = T * N
    E = N * KNN + R
    csr = E * 8 + (R + 1) * 4
    traffic, traffic_src = load_ncu_traffic()
    calls_per_step = 16  # 4 layers x 4 window passes of a meta-step, each kernel
    kern = {}
    for name, ms in kernel_ms.items():
        b = KERNEL_BYTES[name]["minimum"] * G * R
        bd = KERNEL_BYTES[name]["design"] * G * R
        kern[name] = {"ms_per_launch": ms, "algorithmic_bytes": b, "design_bytes": bd, "GBps": b / (ms * 1e-3) / 1e9,
                      "frac": b / (ms * 1e-3) / 1e9 / peak, "design_frac": bd / (ms * 1e-3) / 1e9 / peak,
                      "share_of_step_ms": ms * calls_per_step, "traffic": traffic.get(name)}
    best = max(kern, key=lambda k: kern[k]["ms_per_launch"])
    # graph conv (BASELINE.json asks for it): one 256 -> 256 GCN layer call = read X, write Y, CSR, W
    gcn_bytes = G * (R * (F + F) * 4 + csr) + F * F * 4
    g = stages.get("wf_gcn_layer_fwd_ss_wide")
    gcn_gbps = gcn_bytes / (g["ms_per_meta_step"] / g["calls"] * 1e-3) / 1e9 if g else None
    # whole meta-step against both roofs (~262 MB and 41.81 GFLOP algorithmic per window pass)
    passes = 4 * G
    step_bytes, step_flop = passes * 262e6, passes * 41.81e9
    step = {"algorithmic_bytes": step_bytes, "algorithmic_flop": step_flop,
            "GBps": step_bytes / (ms_per_step * 1e-3) / 1e9, "hbm_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            "TFLOPs": step_flop / (ms_per_step * 1e-3) / 1e12,
            "tensor_frac": (step_flop / (ms_per_step * 1e-3) / 1e12 / tf_peak) if tf_peak else None,
            "tensor_peak_TFLOPs": tf_peak,
            "note": "the step is bound by neither roof: 24 sequential recurrence steps x 32 layer launches of per-step "
                    "latency; every tensor-core product is issued three times (hi/lo operand splits)"}
    return {"bound": "hbm", "kernel": best, "achieved": kern[best]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": kern[best]["frac"], "traffic": kern[best]["traffic"], "traffic_source": traffic_src,
            "peak_source": peak_src, "ms_per_launch": kern[best]["ms_per_launch"],
            "algorithmic_bytes_per_launch": kern[best]["algorithmic_bytes"],
            "design_bytes_per_launch": kern[best]["design_bytes"], "design_frac": kern[best]["design_frac"],
            "kernels": kern, "graph_conv_GBps": gcn_gbps, "graph_conv_frac": gcn_gbps / peak if gcn_gbps else None,
            "graph_conv_note": "one 256->256 GCN layer call (aggregation of the rows with neighbours + the resident-weight SS GEMM "
                               "with TMA-store epilogue), algorithmic bytes = read X + write Y as fp16 hi/lo planes (4 B/value)",
            "step_roofline": step}


def config1_windows_per_sec(sd, dims, batch=16, reps=5):
    """configs[0] shape on the GPU: forward + MSE + backward of a batch of 16 independent windows (batch-1 semantics,
    SURVEY.md D8) of one 441-node region, k = 8; device-timed."""
    import torch

    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.engine import HybridEngine, flatten_trainable, gcn_weights_from_state_dict
    from weatherforecast_stgcn_maml_b200.graph import RegionGraph
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device

    lats, lons, feats, _ = synth.synth_task(3, num_windows=batch + 8, nlat=NLAT, nlon=NLON)
    ei = knn_edge_index_device(lats, lons, KNN, "cuda")
    eng = HybridEngine(dims, 1, batch, "cuda")
    graph = RegionGraph(ei, dims.R, "cuda")
    fd = feats.cuda()
    per = dims.num_nodes * dims.in_channels
    xo = (torch.arange(batch, dtype=torch.long) * per).cuda()
    to = xo + (dims.window + 1) * per
    theta = flatten_trainable(sd, dims, "cuda")
    gw = gcn_weights_from_state_dict(sd, "cuda")
    run = lambda: eng.forward_backward(fd, dims.in_channels, 0, xo, gw, graph, theta, 0, feat=fd, tgt_off=to,
                                       feat_ld=dims.in_channels)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    eng.check()
    ms = e0.elapsed_time(e1) / reps
    return {"windows_per_sec": batch / (ms * 1e-3), "ms_per_batch": ms, "batch": batch,
            "workload": "configs[0] on the GPU: v5 hybrid forward + MSE + backward, batch of 16 independent windows of one "
                        "441-node region (k=8), features resident, tensor-core path"}


def config4_record(peak, tf_peak, batch=32, chunk=8, reps=2):
    """configs[3]: 121 x 121 = 14,641 nodes, k = 8, the graph-conv stack forward AND backward (STGCN.forward through the
    drop-in module API, model.py:30-52 -- the only differentiable use of the convolution), batch 32 as 4 chunks of 8
    windows (gradients accumulate in .grad).  GB/s and TFLOP/s on algorithmic bytes / FLOPs."""
    import torch

    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device
    from weatherforecast_stgcn_maml_b200.model import STGCN

    nlat = nlon = 121
    n, T, H, F = nlat * nlon, 24, 8, 256
    lats, lons = synth.region_grid(nlat, nlon)
    ei = knn_edge_index_device(lats, lons, KNN, "cuda")
    sd = synth.init_v5_state_dict(3, gcn_bias_scale=0.05)
    base = STGCN(24, F, out_channels=12, window_size=T, forecast_horizon=H, dropout_rate=0.0)
    base.load_state_dict({k[len("base_stgcn."):]: v for k, v in sd.items() if k.startswith("base_stgcn.")})
    base = base.cuda().train()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(chunk * T * n, 24, device="cuda", generator=g)

    def run():
        for _ in range(batch // chunk):
            out = base(x, ei)
            out.backward(torch.ones_like(out) / out.numel())

    run()
    torch.cuda.synchronize()
    base.zero_grad(set_to_none=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    WF.check()
    ms = e0.elapsed_time(e1) / reps
    rows = batch * T * n
    # per layer: forward read X + write Y; backward read dY, Y (ReLU mask), X (dW) and write dX -- f32
    byts = rows * 4 * ((24 + F) + 3 * (F + F) + (F + F + 24) + 3 * (F + F + F + F))
    flop = 3 * 2 * rows * (24 * F + 3 * F * F)  # forward + dX + dW
    del x, base
    torch.cuda.empty_cache()
    return {"ms_per_batch32": ms, "windows_per_sec": batch / (ms * 1e-3), "algorithmic_bytes": byts, "algorithmic_flop": flop,
            "GBps": byts / (ms * 1e-3) / 1e9, "hbm_frac": byts / (ms * 1e-3) / 1e9 / peak,
            "TFLOPs": flop / (ms * 1e-3) / 1e12, "tensor_frac": flop / (ms * 1e-3) / 1e12 / tf_peak if tf_peak else None,
            "workload": "configs[3]: 121x121 = 14,641 nodes, k=8, 4 x GCNConv(256)+ReLU forward+backward (STGCN.forward via "
                        "the drop-in modules), batch 32 = 4 chunks of 8 windows, 351,384 rows per window"}


def module_api_ms(sd, dims, reps=10):
    """The drop-in nn.Module route (``out = model(x, edge_index); loss.backward()``, hybrid_model.py:80-117) on one
    window: milliseconds per forward + MSE + backward, to set beside the task-batched engine's per-window time."""
    import torch

    from weatherforecast_stgcn_maml_b200 import functional as WF
    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device
    from weatherforecast_stgcn_maml_b200.hybrid_model import HybridSTGCN_LSTM
    from weatherforecast_stgcn_maml_b200.model import STGCN

    T, H, N = dims.window, dims.horizon, dims.num_nodes
    lats, lons, feats, _ = synth.synth_task(5, num_windows=8, nlat=NLAT, nlon=NLON)
    ei = knn_edge_index_device(lats, lons, KNN, "cuda")
    base = STGCN(24, 256, out_channels=12, window_size=T, forecast_horizon=H, dropout_rate=0.0)
    hyb = HybridSTGCN_LSTM(base, lstm_hidden_size=128, lstm_num_layers=4, lstm_dropout=0.0, out_channels=12,
                           forecast_horizon=H, freeze_base=False)
    hyb.load_state_dict(sd)
    hyb = hyb.cuda().train()
    fd = feats.cuda()
    x = fd[0:T].reshape(T * N, -1).contiguous()                                   # dataset.py:36-37
    y = torch.stack([fd[T + h, :, :12] for h in range(1, H + 1)]).reshape(H * N, 12)  # dataset.py:40-48
    crit = torch.nn.MSELoss()

    def run():
        hyb.zero_grad(set_to_none=True)
        crit(hyb(x, ei), y).backward()

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    WF.check()
    return e0.elapsed_time(e1) / reps


def validation_windows_per_sec(sd, dims, windows=240, reps=3):
    """configs[2], second half (adapt_hybrid_v5.py:216-231): eval-mode forward + MSE over the 240 validation windows of a
    region (16 windows per launch set); device-timed, loss read back once per sweep as the reference's average is."""
    import torch

    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import FineTuner
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device

    lats, lons, feats, _ = synth.synth_task(8, num_windows=windows + 8, nlat=NLAT, nlon=NLON)
    ei = knn_edge_index_device(lats, lons, KNN, "cuda")
    ft = FineTuner(sd, feats, ei, dims, "cuda", region_name="bench", max_samples=windows, train_frac=0.0, dropout=(0, 0, 0))
    idx = list(range(windows))
    ft.validate(idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ft.validate(idx)
    e1.record()
    torch.cuda.synchronize()
    return windows * reps / (e0.elapsed_time(e1) * 1e-3)


def reference_shape_record(sd, dims, reps=2):
    """configs[1] in the reference's own shape (train_hybrid_maml_v5.py:110-184,262-281): a meta-update
    over BATCH_SIZE = 4 sampled tasks, each adapted by 6 epochs x 15 support windows = 90 SGD steps + 1 query pass,
    optimiser step after every GRAD_ACCUMULATION_STEPS = 2 tasks -- two captured graphs of 2 tasks x 91 window passes."""
    import torch

    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer, reference_support_schedule

    tasks = build_tasks(0, 1)[:4]
    rows = tuple(reference_support_schedule(list(range(450))))
    trs = [MetaTrainer(sd, tasks[2 * i:2 * i + 2], dims, "cuda", use_cuda_graph=True, support_rows=rows, accum=2,
                       dropout=(0, 0, 0)) for i in range(2)]

    def update():
        for tr in trs:
            tr.theta.copy_(trs[0].theta)   # the second pair starts from the weights the first step produced
            tr.meta_step()
        trs[0].theta.copy_(trs[1].theta)

    update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        update()
    e1.record()
    torch.cuda.synchronize()
    for tr in trs:
        tr.check()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    out = {"sec_per_meta_update": sec, "meta_updates_per_sec": 1.0 / sec, "window_passes_per_sec": 4 * 91 / sec,
           "workload": "the reference's own meta-update: 4 tasks x (90 inner SGD steps over support windows 0..14 + 1 query "
                       "pass), optimiser step every 2 tasks; 441 nodes, k=8, dropout off"}
    del trs
    torch.cuda.empty_cache()
    return out


def finetune_windows_per_sec(sd, dims, steps=96, warmup=16, dropout=(0.0, 0.0, 0.0)):
    """configs[2] shape (regional adaptation, adapt_hybrid_v5.py:185-203): batch-1 Adam steps on one 441-node region,
    windows visited in a shuffled order; device-timed.  Sequential by construction (one optimiser step per window), so
    this is a latency number: 8 CTAs per LSTM launch."""
    import torch

    from weatherforecast_stgcn_maml_b200 import synth
    from weatherforecast_stgcn_maml_b200.adapt_hybrid_v5 import FineTuner
    from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device

    lats, lons, feats, _ = synth.synth_task(7, num_windows=steps + warmup + 8, nlat=NLAT, nlon=NLON)
    ei = knn_edge_index_device(lats, lons, KNN, "cuda")
    ft = FineTuner(sd, feats, ei, dims, "cuda", region_name="bench", max_samples=steps + warmup, train_frac=1.0,
                   dropout=dropout)
    order = torch.randperm(steps + warmup, generator=torch.Generator().manual_seed(0)).tolist()
    for i in order[:warmup]:
        ft.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in order[warmup:]:
        ft.step(i)
    e1.record()
    torch.cuda.synchronize()
    ft.engine.check()
    return steps / (e0.elapsed_time(e1) * 1e-3)


def run_gpu(args, rank, local, world):
    import torch

    import __graft_entry__ as entry

    if rank == 0:
        entry.build()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    from weatherforecast_stgcn_maml_b200 import _lib, synth
    from weatherforecast_stgcn_maml_b200.engine import V5Dims
    from weatherforecast_stgcn_maml_b200.train_hybrid_maml_v5 import MetaTrainer

    _lib.load()
    torch.cuda.set_device(local)
    dims = V5Dims(num_nodes=NLAT * NLON)
    sd = synth.init_v5_state_dict(42)
    tasks = build_tasks(rank, world)
    # headline = the deterministic parity configuration (dropout off); the reference's training
    # configuration (p = 0.2 at its three sites) is measured beside it as `dropout_on`
    kw = dict(support_rows=SUPPORT_ROWS, accum=TASKS_PER_GPU * world, dropout=(0.0, 0.0, 0.0))

    # ---- headline: device-resident, CUDA-graphed
    tr = MetaTrainer(sd, tasks, dims, "cuda", use_cuda_graph=True, **kw)
    tr.meta_step()  # capture
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, loss = timed_steps(tr, args.steps, args.warmup, world, sync_loss=False)
    clocks = sampler.stop() if sampler else None
    launches_per_step = tr.launches_per_step + 2  # + the sumsq and AdamW kernels of the outer update (graph nodes too)
    stages = stage_breakdown(tr) if rank == 0 else None
    kernel_ms = time_recurrence_kernels(tr) if rank == 0 else None
    del tr
    torch.cuda.empty_cache()

    # ---- the reference's TRAINING configuration: dropout 0.2 at its three sites (train_hybrid_maml_v5.py:197,205)
    from weatherforecast_stgcn_maml_b200.engine import REFERENCE_DROPOUT

    kw_on = dict(kw, dropout=REFERENCE_DROPOUT)
    tr_on = MetaTrainer(sd, tasks, dims, "cuda", use_cuda_graph=True, **kw_on)
    tr_on.meta_step()
    ms_on, loss_on = timed_steps(tr_on, max(3, args.steps // 2), 3, world, sync_loss=False)
    steps_on = max(3, args.steps // 2)
    tr_on.check()
    del tr_on
    torch.cuda.empty_cache()

    # ---- e2e: features in pinned host memory, upload per step, loss read back per step
    tr2 = MetaTrainer(sd, tasks, dims, "cuda", use_cuda_graph=True, host_staging=True, **kw)
    tr2.meta_step()
    ms2, loss2 = timed_steps(tr2, args.steps, args.warmup, world, sync_loss=True)
    h2d = tr2.stager.h2d_bytes + 32  # + the AdamW hyper-parameter block
    tr2.check()
    del tr2
    torch.cuda.empty_cache()

    # ---- configs[4] as BASELINE.json words it: 120 tasks in total over the N GPUs (strong scaling; N >= 2 only)
    config5 = None
    if world > 1 and 120 % world == 0:
        from weatherforecast_stgcn_maml_b200.dist import shard_tasks
        from weatherforecast_stgcn_maml_b200.graphBuilder import knn_edge_index_device

        tasks5 = []
        for t in shard_tasks(120, rank, world):
            lats, lons, feats, _ = synth.synth_task(t, num_windows=64, nlat=NLAT, nlon=NLON)
            tasks5.append((feats, knn_edge_index_device(lats, lons, KNN, "cuda")))
        tr5 = MetaTrainer(sd, tasks5, dims, "cuda", use_cuda_graph=True, support_rows=SUPPORT_ROWS, accum=120,
                          dropout=(0.0, 0.0, 0.0), query_row=40)
        tr5.meta_step()
        ms5, _ = timed_steps(tr5, 5, 3, world, sync_loss=False)
        tr5.check()
        config5 = {"global_tasks": 120, "tasks_per_gpu": 120 // world, "ms_per_step": ms5 / 5,
                   "meta_steps_per_sec": 1.0 / (ms5 * 1e-3 / 5), "task_passes_per_sec": 120 / (ms5 * 1e-3 / 5),
                   "scaling": "strong", "workload": "configs[4]: 120 synthetic region tasks across the N GPUs, NCCL "
                                                    "all-reduce of the meta-gradient, one 120-task meta-step"}
        del tr5, tasks5
        torch.cuda.empty_cache()

    if rank != 0:
        return
    sec_step = ms * 1e-3 / args.steps
    value = world / sec_step
    peak, peak_src = peaks()
    tf_peak = tensor_peak()
    roof = roofline_report(stages, kernel_ms, TASKS_PER_GPU, peak, peak_src, sec_step * 1e3, tf_peak)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": workload_config(world),
        "windows_per_sec": 60 * world / sec_step, "meta_loss": loss, "clocks": clocks,
        "dropout_on": {"value": world / (ms_on * 1e-3 / steps_on), "unit": UNIT, "ms_per_step": ms_on / steps_on,
                       "meta_loss": loss_on, "p": list(REFERENCE_DROPOUT),
                       "note": "same workload with the reference's training configuration: dropout 0.2 after GCN layers "
                               "1-3, between the LSTM layers and on the head input (fused counter-based masks, "
                               "regenerated in backward); the headline value keeps dropout off (parity configuration)"},
        "e2e": {"value": world / (ms2 * 1e-3 / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms2 / args.steps, "meta_loss": loss2,
                "note": "features in pinned host memory; every step uploads the rows its windows read (double buffered: the "
                        "copy for step t+1 runs on a copy stream while step t computes) and reads the loss back (.item())"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": roof,
        "stages_ms_per_meta_step": {k: round(v["ms_per_meta_step"], 4) for k, v in stages.items()},
    }
    line["finetune"] = {"windows_per_sec": finetune_windows_per_sec(sd, dims), "unit": "windows/s",
                        "windows_per_sec_dropout_on": finetune_windows_per_sec(sd, dims, dropout=REFERENCE_DROPOUT),
                        "validation_windows_per_sec": validation_windows_per_sec(sd, dims),
                        "workload": "configs[2]: batch-1 Adam fine-tuning steps on one 441-node region (k=8), "
                                    "forward + MSE + backward + clip + Adam per window; validation = eval-mode forward + "
                                    "MSE over 240 windows, 16 per launch set; 1 GPU"}
    if config5 is not None:
        line["config5"] = config5
    if world == 1:
        line["config1"] = config1_windows_per_sec(sd, dims)
        ms_mod = module_api_ms(sd, dims)
        line["module_api"] = {"ms_per_window_fwd_bwd": ms_mod, "engine_ms_per_window_fwd_bwd": sec_step * 1e3 / 60,
                              "note": "drop-in nn.Module route (HybridSTGCN_LSTM.forward + loss.backward(), one window, "
                                      "autograd, same tcgen05 kernels through a leased engine) vs the task-batched engine "
                                      "(meta-step time / 60 window passes)"}
        line["config4"] = config4_record(peak, tf_peak)
        line["config2_reference_shape"] = reference_shape_record(sd, dims)
    if world == 1 and not args.no_cpu_baseline:
        sec, cores, kind = cpu_window_pass(3, 1)
        sec_p, _ = port_window_pass_seconds(2, 1, literal=True)
        sec_b, _ = port_window_pass_seconds(3, 1, literal=False)
        line["cpu_baseline"] = {
            "value": 1.0 / (60 * sec), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "3 timed steps: " + CPU_SAMPLE, "sec_per_window_pass": sec,
            "port_sec_per_window_pass": sec_p, "batched_port_sec_per_window_pass": sec_b}
    else:
        line["cpu_baseline"] = None
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    from weatherforecast_stgcn_maml_b200.dist import init_from_env

    if world_env > 1:
        # a rank stuck in a collective must not hang the job: dump every thread's stack and exit
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ.get("WF_BENCH_WATCHDOG_S", "900")), exit=True)
    rank, local, world = init_from_env("nccl")
    if world != args.gpus and world_env > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    try:
        run_gpu(args, rank, local, world)
    finally:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
