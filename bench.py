#!/usr/bin/env python
"""Headline benchmark: summarised videos / second through the whole hot path

    features -> transformer scorer (bf16 tcgen05) -> sigmoid -> shot pooling -> knapsack 15 %
             -> keyshot mask -> F-score (-> F-score gather across ranks)

on BASELINE.json's throughput-sweep workload (config 5): synthetic videos with N log-uniform in
[128, 8192] frames, 1024-d fp32 features, 20 annotators, sharded over N GPUs of one node
(weak scaling: every rank gets --videos videos per step).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md "Measurement").
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
for _p in (ROOT, os.path.join(ROOT, "video-summarization_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

MODEL_KW = dict(num_heads=4, d_model=256, num_layers=4, sparsity=0., use_cls=False, dropout=0.3,
                num_classes=1, use_pos=True)                 # run_finetune.sh:1 / train.py:29-34
N_LO, N_HI, N_USERS = 128, 8192, 20
METRIC, UNIT = "summarized_videos_per_sec", "videos/s"
# CPU sample: one video per twelfth of the log-uniform length range (geometric midpoints of [128, 8192])
REF_SAMPLE_N = tuple(int(round(128 * 64 ** ((i + 0.5) / 12))) for i in range(12))


def workload_name(videos, lo=None, hi=None):
    lo, hi = lo or N_LO, hi or N_HI
    return (f"config5 throughput sweep: {videos} synthetic videos/GPU/step, N log-uniform [{lo},{hi}], "
            f"1024-d fp32 features, {N_USERS} users, scorer d256/h4/L4 + knapsack 15% + F-score(avg)")


def attention_flops(seqlens, d=256, layers=4):
    """QK^T + PV, per layer launch: 4 * N^2 * d per video."""
    return float(sum(4.0 * n * n * d for n in seqlens))


def scorer_flops(seqlens, d=256, layers=4, in_features=1024):
    return float(sum(n * (2 * in_features * d + layers * (24 * d * d + 4 * n * d) + 2 * d) for n in seqlens))


# --------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every ~20 ms through NVML (nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._t, self._nvml = threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: resolve through the UUID torch reports when possible
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [x.strip() for x in out.split(",")]
        if len(c) >= 6:
            self.sm.append(float(c[0]))
            self.max_mhz = float(c[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU algorithm (oracle port) on host cores
# --------------------------------------------------------------------------------------------
def _scores_of(sd, sample_videos, device="cpu"):
    """Scorer of the oracle port (torch fp32 eager, the reference's own module restated) + the caller's sigmoid."""
    import torch
    from oracle import scorer_ref
    out = []
    for v in sample_videos:
        x = torch.from_numpy(v.features).unsqueeze(0).to(device)
        logits, _ = scorer_ref.scorer_forward(sd, x, num_heads=4)
        out.append(torch.sigmoid(logits.view(1, -1)).squeeze(0).cpu().numpy())         # train.py:144,148
    return out


def cpu_reference_step(sd, sample_videos, pool=None):
    """One pass of the reference's val_step + eval_metrics body (minus scipy correlations) over the sample, in the
    reference's own execution model: torch fp32 on CPU (all host threads) for the scorer, pure Python loops for
    pooling / knapsack / F-score, builtin `sum()` counts included (oracle/ref_port.py).
    pool=None: single process, as the reference ships.  pool=multiprocessing pool: the pure-Python stages of the
    videos run in parallel worker processes (BASELINE.md 4.3 'fair')."""
    from oracle import ref_port
    scores = _scores_of(sd, sample_videos)
    jobs = [(v.change_points, sc, v.n_frames, v.picks, v.user_summary, "avg") for v, sc in zip(sample_videos, scores)]
    fs = pool.map(ref_port.pool_eval_video, jobs, chunksize=1) if pool is not None else [ref_port.pool_eval_video(j) for j in jobs]
    return float(np.mean(fs))


def reference_setup():
    import torch
    from vsum_b200.model import SimNet
    from vsum_b200.synthetic import make_video
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    sd = SimNet(**MODEL_KW).state_dict()
    sample = [make_video(5000 + i, n, n_users=N_USERS) for i, n in enumerate(REF_SAMPLE_N)]
    return sd, sample, cores


REF_BUDGET_S = 180.0      # the whole reference-arm run (warm-up + timed steps) should end within a few minutes


def pick_sample(sd, full, steps, warmup, pool=None):
    """Largest of the 12- / 6- / 3-video samples (every 1st / 2nd / 4th length class) whose estimated run time
    fits REF_BUDGET_S; the estimate comes from one untimed pass over the smallest sample (measured cost ratios
    of the three samples: about 9 : 4 : 1; rounded up)."""
    small = full[1::4]
    t0 = time.perf_counter()
    cpu_reference_step(sd, small, pool)
    dt = time.perf_counter() - t0
    n_steps = max(1, steps + warmup)
    if 10.0 * dt * n_steps <= REF_BUDGET_S:
        return full
    if 4.5 * dt * n_steps <= REF_BUDGET_S:
        return full[0::2]
    return small


def _make_pool(cores, n_jobs):
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(max(1, min(cores, n_jobs)))      # spawn: no fork of a process that owns torch / CUDA threads
    pool.map(abs, range(pool._processes))                                # workers up and imported before anything is timed
    return pool


def run_cpu_baseline(steps=1, warmup=0, fair_value=False, gpu_eager=False):
    """`value`: the sample's videos/s.  Two arrangements: 'as shipped' -- one process, torch on all
    host threads, the pure-Python stages one video after the other -- and 'fair' -- the same with the pure-Python
    stages of the videos spread over a process pool.  The reference arm (`fair_value`) reports the faster, fair one."""
    sd, sample, cores = reference_setup()
    pool = _make_pool(cores, len(sample))
    try:
        sample = pick_sample(sd, sample, 2 * steps, 2 * warmup, pool)
        for _ in range(warmup):
            cpu_reference_step(sd, sample, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(sd, sample, pool)
        dt_fair = (time.perf_counter() - t0) / steps
        n_ship = 1 if fair_value else steps
        t0 = time.perf_counter()
        for _ in range(n_ship):
            cpu_reference_step(sd, sample, None)
        dt_ship = (time.perf_counter() - t0) / n_ship
    finally:
        pool.terminate()
    desc = (f"{len(sample)} videos, N={[v.n_steps for v in sample]} (evenly spaced classes of the log-uniform length range; "
            f"mean N^2 {np.mean([v.n_steps ** 2 for v in sample]) / 1e6:.1f} M); oracle port of the reference (kind 'port': "
            "/root/reference is not on the GPU box): torch fp32 CPU scorer on all host threads + pure-Python pooling / knapsack / "
            "F-score with the reference's builtin sum() counts")
    dt = dt_fair if fair_value else dt_ship
    base = {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
            "arrangement": "fair" if fair_value else "as_shipped",
            "as_shipped": {"value": len(sample) / dt_ship, "s_per_step": dt_ship,
                           "what": "single process, torch intra-op threads = all cores, pure-Python stages sequential"},
            "fair": {"value": len(sample) / dt_fair, "s_per_step": dt_fair, "processes": min(cores, len(sample)),
                     "what": "scorer as above, pure-Python stages of the videos in a multiprocessing pool"}}
    if gpu_eager:
        base["reference_gpu_eager"] = run_gpu_eager(sd, sample)
    return base, dt


def run_gpu_eager(sd, sample):
    """BASELINE.md 4.7 (secondary): the reference's PyTorch module (oracle restatement, fp32 eager ops, one video per
    forward as train.py:139-143) with weights and features on the B200; scorer only, and with the pure-Python
    evaluation stages behind it (as shipped: sequential on the host)."""
    import torch
    from oracle import ref_port
    dev = torch.device("cuda", torch.cuda.current_device())
    sdg = {k: v.to(dev) for k, v in sd.items()}
    import oracle.scorer_ref as sr
    tab = sr.positional_table
    sr.positional_table = lambda n, d: tab(n, d).to(dev)        # the reference registers the table as a module buffer on the device
    try:
        _scores_of(sdg, sample[:2], dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scores = _scores_of(sdg, sample, dev)                   # .cpu() per video = the reference's own per-video sync
        dt_sc = time.perf_counter() - t0
    finally:
        sr.positional_table = tab
    t0 = time.perf_counter()
    for v, sc in zip(sample, scores):
        ref_port.pool_eval_video((v.change_points, sc, v.n_frames, v.picks, v.user_summary, "avg"))
    dt_ev = time.perf_counter() - t0
    return {"scorer_only_videos_per_s": len(sample) / dt_sc, "with_python_eval_videos_per_s": len(sample) / (dt_sc + dt_ev),
            "what": "torch fp32 eager scorer on cuda:0 (oracle restatement of simnet.py, without the reference's per-layer "
                    "attention_weight.detach().cpu() copy of simnet.py:164, i.e. a faster reference), one video per forward; evaluation stages "
                    "pure Python on the host"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = run_cpu_baseline(args.steps, args.warmup, fair_value=True)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.videos), "l2": "n/a (CPU)",
                       "sample": "each step is a bounded, length-stratified sample of the workload (see cpu_baseline.sample), "
                                 "not the 256-video batch"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def pin_rank_to_cores(local, world):
    """Ranks of one node share its host cores evenly (loader thread, pinned allocations and torch's own threads of a rank stay
    on its slice) instead of every rank spawning threads over all of them."""
    if world <= 1 or not hasattr(os, "sched_getaffinity"):
        return None
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // world)
    mine = cores[local * per:(local + 1) * per] or cores
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    return len(mine)


def scratch_dir(need_bytes):
    """Where the synthetic pack files go: VSUM_BENCH_DIR, else /dev/shm when it has room, else the temp directory."""
    import shutil
    import tempfile
    cands = [os.environ.get("VSUM_BENCH_DIR"), "/dev/shm", tempfile.gettempdir()]
    for d in cands:
        if d and os.path.isdir(d) and os.access(d, os.W_OK):
            try:
                if shutil.disk_usage(d).free > need_bytes * 1.1:
                    return d
            except OSError:
                pass
    return tempfile.gettempdir()


def write_bench_pack(path, videos, feats_of, compact):
    """One `.vspack` of the synthetic val videos (fp32 features + fp32 user summaries as the h5 files hold them, or the
    compact form: bf16 features + uint8 user summaries)."""
    from vsum_b200.data import write_pack

    def gen():
        for v in videos:
            yield dict(name=v.name, features=feats_of(v), picks=v.picks, change_points=v.change_points, n_frames=v.n_frames,
                       user_summary=v.user_summary)
    write_pack(path, gen(), user_summary_u8=compact, features_bf16=compact)


def train_records(world, rank, dev, steps=20, warmup=5):
    """BASELINE configs 3 and 4 beside the headline: one training step = forward with activation tape + masked MSE + backward
    through the C ABI + per-layer gradient all-reduce from a communication stream (world > 1) + fused Adam.  finetune: 4
    videos / GPU of 200-400 frames (run_finetune.sh); pretrain-shaped: 8 videos / GPU of 2048 frames.  Features resident on
    the device, CUDA events, max over ranks; weak scaling.  At world > 1 the data-parallel step is also checked against the
    single-process step on the concatenated batch (rank 0 regenerates every shard)."""
    import torch
    import torch.distributed as dist
    from vsum_b200 import _cabi
    from vsum_b200.model import SimNet
    from vsum_b200.sharding import DataParallel, scorer_cost
    from vsum_b200.utils import mse_with_mask_loss

    def shard(cfg, r):
        rng = np.random.default_rng(99 + r)
        lens = [int(n) for n in (rng.integers(200, 401, 4) if cfg == "finetune" else [2048] * 8)]
        g = torch.Generator(device=dev).manual_seed(7 + r)
        T = sum(lens)
        feats = torch.randn((T, 1024), device=dev, generator=g)
        tgt = torch.rand((1, T), device=dev, generator=g)
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
        return lens, feats, tgt, cu

    out = {}
    for cfg in ("finetune", "pretrain_8x2048"):
        lens, feats, tgt, cu = shard(cfg, rank)
        T, bs = sum(lens), len(lens)
        nopad = torch.zeros((1, T), dtype=torch.bool, device=dev)
        torch.manual_seed(1234)
        model = SimNet(**MODEL_KW).to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
        ddp = DataParallel(model)

        def step():
            opt.zero_grad(set_to_none=True)
            o, _ = model.forward_packed_train(feats, cu, lens)
            ddp.loss(o.view(1, T, 1), tgt, nopad, batch=bs, nmax=max(lens)).backward()
            loss = ddp.finish()
            opt.step()
            return loss
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        n0 = _cabi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        launches = (_cabi.launch_count() - n0) // steps
        v = torch.tensor([e0.elapsed_time(e1) / steps, float(T), float(bs), sum(scorer_cost(n) for n in lens)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = v.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            v[0] = mx[0]
        t_ms, frames, videos, fwd = (float(x) for x in v.tolist())
        rec = {"ms_per_step": t_ms, "frames_per_s": frames / t_ms * 1e3, "videos_per_s": videos / t_ms * 1e3,
               "train_tflops": 3.0 * fwd / t_ms / 1e9, "kernel_launches_per_step": int(launches), "final_loss": float(loss),
               "steps": steps, "warmup": warmup, "dtype": "bf16 (tcgen05 linears tf32/bf16, attention bf16; fp32 master weights)"}
        if world > 1 and cfg == "finetune":
            # equality with the single-process step: dropout off, DP gradient vs rank 0's gradient on the concatenated batch
            model.eval()
            model.train_precision = "fp32"      # the check is about the data-parallel logic: fp32 SIMT kernels, where scaling before / after the backward commutes up to rounding
            opt.zero_grad(set_to_none=True)
            o, _ = model.forward_packed_train(feats, cu, lens)
            ddp.loss(o.view(1, T, 1), tgt, nopad, batch=bs, nmax=max(lens)).backward()
            ddp.finish()
            got = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
            ddp.detach()
            err = torch.zeros(1, device=dev)
            if rank == 0:
                parts = [shard(cfg, r) for r in range(world)]
                all_lens = [n for p_ in parts for n in p_[0]]
                f_all, t_all = torch.cat([p_[1] for p_ in parts]), torch.cat([p_[2] for p_ in parts], dim=1)
                cu_all = torch.tensor(np.concatenate([[0], np.cumsum(all_lens)]), dtype=torch.int32, device=dev)
                model.zero_grad(set_to_none=True)
                o, _ = model.forward_packed_train(f_all, cu_all, all_lens)
                mse_with_mask_loss(o.view(1, -1, 1), t_all, torch.zeros_like(t_all, dtype=torch.bool),
                                   denom=float(len(all_lens) * max(all_lens))).backward()
                want = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
                err[0] = (got - want).abs().max() / want.abs().max()
            dist.broadcast(err, 0)
            rec["dp_equals_single_process"] = {"max_abs_err_over_max_abs_grad": float(err), "ok": bool(float(err) < 2e-4),
                                               "what": "all-reduced gradient of the sharded batch vs rank 0's gradient of the concatenated batch (dropout off, fp32 kernels)"}
        ddp.detach()
        out[cfg] = rec
        del model, opt, ddp
    return out


def main_b200(args):
    import torch
    import torch.distributed as dist
    from vsum_b200 import _cabi
    from vsum_b200.data import PackedDataset, PackedEvalLoader
    from vsum_b200.model import SimNet
    from vsum_b200.pipeline import DeviceBatch, Summarizer, pack_videos
    from vsum_b200.synthetic import make_video, video_length

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    rank_cores = None if args.no_pin else pin_rank_to_cores(local, world)
    if rank_cores:
        torch.set_num_threads(rank_cores)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"                      # no version banner: stdout is the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    if args.only_train:
        rec = train_records(world, rank, dev, steps=50, warmup=10)
        if rank == 0:
            print(json.dumps({"train": rec, "host_cores_per_rank": rank_cores, "n_gpus": world}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- workload: weak scaling, rank r owns videos [r*V*NB, (r+1)*V*NB) of the global id space: NB distinct batches of V
    # videos; the resident leg (`value`) runs on the first of them
    V, NB = args.videos, max(1, args.e2e_batches)
    try:
        import psutil
        avail = psutil.virtual_memory().available / max(world, 1)
        per_batch = V * 2108 * 4096 * 1.3 * 2.2          # pack file (tmpfs) + its page-locked copy, mean 2108 frames / video
        while NB > 2 and NB * per_batch > 0.6 * avail:
            NB -= 1
    except Exception:
        pass
    ids = [rank * V * NB + i for i in range(V * NB)]
    videos = [make_video(v, video_length(v, args.len_lo, args.len_hi), n_users=N_USERS, with_features=False) for v in ids]
    hb = pack_videos(videos[:V], pin=True, with_features=False)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats_dev = torch.rand((hb.n_steps, 1024), device=dev, generator=g)        # rng.random-like features in [0,1)
    torch.manual_seed(1234)
    model = SimNet(**MODEL_KW).to(dev).eval()
    summ = Summarizer(model, "avg", eval_sms=args.eval_sms)
    db = DeviceBatch(hb, dev, features=feats_dev)
    f_all = torch.empty(world * V, dtype=torch.float64, device=dev) if world > 1 else None

    def step_resident(i):
        # software pipeline over consecutive batches: scorer(i+1) on the main stream overlaps
        # pooling / knapsack / F-score (+ gather) of batch i on the side stream
        f = summ.submit_device(db, i & 1)
        if world > 1:
            with torch.cuda.stream(summ._side):
                dist.all_gather_into_tensor(f_all, f)       # the path's only exchange (F-score gather)
        return f

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        summ.drain(dev)                                     # the timed region ends when ALL streams are done
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: inputs resident in HBM (features 1024*4 B/frame >> L2, nothing to flush)
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    summ.drain(dev)
    launches0 = _cabi.launch_count()
    _cabi.profile_begin()
    with ClockSampler(local) as clocks:
        ms_total = timed(step_resident, args.steps)
    prof = _cabi.profile_end()
    launches = _cabi.launch_count() - launches0
    ms_step = ms_total / args.steps
    value = world * V / (ms_step * 1e-3)

    # ---- per-kernel times from ONE non-pipelined pass (single stream: side-stream kernels are not timed under the next
    # batch's scorer)
    torch.cuda.synchronize()
    _cabi.profile_begin()
    summ.run_device(db)
    torch.cuda.synchronize()
    prof_alone = _cabi.profile_end()
    del db, feats_dev

    # ---- e2e: starts where the reference starts -- a dataset file.  NB distinct batches of V videos in a `.vspack`,
    # read into page-locked memory once (dataset load, not timed), then per step, INSIDE the timed region: batch building
    # on the loader's background thread (native metadata gather), one H2D per video out of the dataset on the loader's copy
    # stream, the whole path, D2H of the F-scores.  Steady state: the loader runs up to `slots - 1` batches ahead both when
    # the timed region starts and when it ends, so K steps contain the copies of exactly K batches.
    def e2e_leg(compact):
        gg = torch.Generator(device=dev).manual_seed(4321 + rank)

        def feats_of(v):
            x = torch.rand((v.n_steps, 1024), device=dev, generator=gg)
            return (x.bfloat16().view(torch.int16).cpu().numpy().view(np.uint16) if compact else x.cpu().numpy())
        frames = sum(v.n_steps for v in videos)
        need = frames * (2048 if compact else 4096) + sum(v.user_summary.size for v in videos) * (1 if compact else 4)
        path = os.path.join(scratch_dir(need), f"vsum_bench_r{rank}_{os.getpid()}_{'compact' if compact else 'f32'}.vspack")
        t0 = time.perf_counter()
        write_bench_pack(path, videos, feats_of, compact)
        t_write = time.perf_counter() - t0
        t0 = time.perf_counter()
        try:
            ds = PackedDataset(path, split="val", resident="pinned")
        finally:
            os.unlink(path)                                 # the page-locked copy is the dataset from here on
        t_open = time.perf_counter() - t0
        loader = PackedEvalLoader(ds, batch_size=V, device=dev, slots=3, cycle=True)
        h2d = []
        state = {"it": None}

        def step(i):
            if state["it"] is None:                         # first step of a run: the loader's thread starts HERE, nothing is prefetched
                state["it"] = iter(loader)
            b = next(state["it"])
            h2d.append(b.h2d_bytes)
            return summ.submit_device(b, 2 + (i & 1), to_host=True)

        def stop():
            if state["it"] is not None:
                state["it"].close()
                state["it"] = None
        for i in range(max(2, NB)):
            step(i)
        summ.drain(dev)
        torch.cuda.synchronize()
        stop()
        h2d.clear()
        n_col = len(loader.collate_ms)
        steps = max(4 * NB, args.steps)                     # enough steps that the cold start (one exposed copy) weighs little
        ms = timed(step, steps) / steps                     # cold pipeline at the start, everything drained at the end
        col, iss = loader.collate_ms[n_col:n_col + steps], loader.issue_ms[n_col:n_col + steps]
        stop()
        out = {"value": world * V / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
               "h2d_bytes_per_step": int(np.mean(h2d)) * world, "d2h_bytes_per_step": 8 * V * world,
               "h2d_gb_per_s_per_gpu": float(np.mean(h2d)) / (ms * 1e-3) / 1e9,
               "distinct_batches": NB, "frames_per_gpu_per_step": frames / NB,
               "pipeline": "cold start: the loader thread starts with the first timed step (no batch prefetched before the timed region) and "
                           "the region ends when the last F-scores are on the host",
               "batch_build_ms": float(np.mean(col)) if col else None,
               "copy_issue_ms": float(np.mean(iss)) if iss else None,
               "batch_build": "1 background thread per rank: batch_build_ms = native metadata gather (vsum_pack_eval_collate) into a pinned "
                              "blob; copy_issue_ms = issuing one cudaMemcpyAsync per video and array out of the page-locked dataset "
                              "(vsum_pack_h2d; blocks while the copy queue is full); no host-side copy of features / user summaries",
               "dataset_load_s": {"write_pack": round(t_write, 2), "open_pinned": round(t_open, 2), "bytes": int(need)}}
        del loader
        ds.close()
        return out

    e2e = e2e_leg(compact=False)
    e2e_compact = e2e_leg(compact=True)
    e2e_compact["note"] = ("same pass from a pack with bf16 features + uint8 user summaries (write_pack features_bf16, user_summary_u8); "
                           "inputs are rounded once offline, scores stay within the 1e-2 bf16 tolerance")

    # ---- plain pinned host-to-device copy rate with every rank copying at once: the bound the fp32 e2e leg sits on
    probe_n = 1 << 30
    src = torch.empty(probe_n, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(probe_n, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    probe = torch.tensor([4 * probe_n / (p0.elapsed_time(p1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
    probe_all = [probe.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(probe_all, probe)
    probe_all = [float(x.item()) for x in probe_all]
    del src, dst

    train = None if args.no_train else train_records(world, rank, dev)

    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))       # kernel timed inside a long step
        peak_hbm = float(peaks.get("hbm_gbs", 6500.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:   # dram bytes per launch from the committed ncu --set full capture of this same workload
            tsrc = "profiles/r02_attn2_ncu_full_default_workload.json"
            rec = json.load(open(os.path.join(ROOT, tsrc)))
            if rec["frames_per_gpu_per_step"] == hb.n_steps:
                traffic, traffic_src = rec["dram_bytes_total"] / 1e9, tsrc
        except Exception:
            pass
        att_ms, att_n = prof.get("attention", (0.0, 0))
        att_flops = attention_flops(hb.seqlens)                            # per launch (one layer, whole batch)
        n_att = 4 * args.steps                                             # fast-pass launches (the exact pass returns at once)
        achieved = att_flops / (att_ms / n_att * 1e-3) / 1e12 if att_ms > 0 else None
        kernel_ms = {k: round(v[0], 4) for k, v in prof_alone.items() if v[1]}
        kernel_ms_pipelined = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}
        gpu_ms = sum(kernel_ms.values())
        scorer_ms = sum(v for k, v in kernel_ms.items() if k in ("embed_gemm", "qkv_gemm", "attention", "oproj_ln_gemm", "fc1_gemm", "fc2_ln_gemm", "ffn_fused"))
        m = hb.meta
        cells = float(sum((int(m.cu_shots[i + 1]) - int(m.cu_shots[i])) * (int((int(m.sum_offsets[i + 1]) - int(m.sum_offsets[i])) * 0.15) + 1)
                          for i in range(m.B)))
        us_bytes = float(m.user_summary.nbytes + m.sum_offsets[-1])
        ev = {}
        if kernel_ms.get("knapsack"):
            ev["knapsack"] = {"ms": kernel_ms["knapsack"], "dp_cells": cells, "gcells_per_s": cells / kernel_ms["knapsack"] / 1e6,
                              "bound": "shared-memory bandwidth + 2 barriers per shot (latency)"}
        if kernel_ms.get("overlap"):
            gbs = us_bytes / kernel_ms["overlap"] / 1e6
            ev["overlap"] = {"ms": kernel_ms["overlap"], "algorithmic_bytes": us_bytes, "gb_per_s": gbs, "hbm_peak_gb_per_s": peak_hbm,
                             "frac": gbs / peak_hbm, "bound": "hbm"}
        if kernel_ms.get("shot_mean"):
            sm_bytes = float(hb.n_steps * 8 + int(m.cu_shots[-1]) * 20)
            ev["shot_mean"] = {"ms": kernel_ms["shot_mean"], "algorithmic_bytes": sm_bytes, "gb_per_s": sm_bytes / kernel_ms["shot_mean"] / 1e6,
                               "bound": "latency (one thread per shot walks ~150 frames)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(V, args.len_lo, args.len_hi), "videos_per_gpu_per_step": V, "frames_per_gpu_per_step": hb.n_steps,
                       "l2": "inputs (fp32 features, %.2f GB/GPU) are larger than L2; no flush needed" % (hb.n_steps * 4096 / 1e9),
                       "parallelism": f"videos sharded over {world} GPU(s), F-score all-gather",
                       "pipelining": "scorer(batch i+1) overlaps pooling/knapsack/F-score(batch i) on a side stream; e2e adds the loader's copy stream",
                       "host_cores_per_rank": rank_cores or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()),
                       "reference_arm": "bench.py --impl reference times a 12-video length-stratified sample per step (cpu_baseline.sample), not this batch"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "e2e_compact_pack": e2e_compact,
            "h2d_probe": {"gb_per_s_per_gpu": probe_all, "aggregate_gb_per_s": sum(probe_all),
                          "what": "plain cudaMemcpyAsync of one 1 GiB page-locked buffer per rank, all ranks at once, 4 copies"},
            "gpu_launches": launches,
            "roofline": {"kernel": "attn2_tc05_kernel (varlen QK^T/softmax/PV, tcgen05, two query tiles per CTA, P in tensor memory)", "bound": "tensor",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                         "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch_gb": hb.n_steps * (768 + 256) * 2 / 1e9,
                         "peak_source": peak_src,
                         "flops_per_launch": att_flops, "launch_ms": att_ms / n_att,
                         "launch_ms_source": "CUDA events around the fast-pass + exact-pass launch pair on the main stream, inside the timed (pipelined) region",
                         "share_of_step": (kernel_ms.get("attention", 0.0)) / gpu_ms if gpu_ms else None},
            "scorer_tflops": scorer_flops(hb.seqlens) / (scorer_ms * 1e-3) / 1e12 if scorer_ms else None,
            "kernel_ms_per_step": kernel_ms,
            "kernel_ms_per_step_source": "one non-pipelined pass on a single stream (no contention between the scorer and the evaluation stream)",
            "kernel_ms_per_step_pipelined": kernel_ms_pipelined,
            "eval_kernels": ev,
            "train": train,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = run_cpu_baseline(1, 0, gpu_eager=True)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos", type=int, default=256, help="videos per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pin", action="store_true", help="do not restrict each rank to its slice of the host cores")
    ap.add_argument("--only-train", action="store_true", help="print only the training-step sub-record (experiments)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step sub-record (BASELINE configs 3 and 4)")
    ap.add_argument("--e2e-batches", type=int, default=4, help="distinct batches in the end-to-end legs' pack file (reduced when host memory is short)")
    ap.add_argument("--eval-sms", type=int, default=0, help="SMs left to the evaluation stream in pipelined mode (0 = no partition)")
    ap.add_argument("--len-lo", type=int, default=N_LO, help="shortest video (frames); default = BASELINE config 5")
    ap.add_argument("--len-hi", type=int, default=N_HI, help="longest video (frames)")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
