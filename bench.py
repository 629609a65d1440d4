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
def cpu_reference_step(sd, sample_videos):
    """One pass of the reference's val_step + eval_metrics body (minus scipy correlations) over
    the sample, in the reference's own execution model: torch fp32 on CPU for the scorer, pure
    Python loops for pooling / knapsack / F-score (oracle/ref_port.py)."""
    import torch
    from oracle import ref_port, scorer_ref
    fs = []
    for v in sample_videos:
        logits, _ = scorer_ref.scorer_forward(sd, torch.from_numpy(v.features).unsqueeze(0), num_heads=4)
        scores = torch.sigmoid(logits.view(1, -1)).squeeze(0).numpy()                  # train.py:144,148
        summary, *_ = ref_port.summarize_video(v.change_points, scores, v.n_frames, v.picks)
        fs.append(ref_port.fscore_video(summary, v.user_summary, "avg"))
    return float(np.mean(fs))


def reference_setup():
    import torch
    from vsum_b200.model import SimNet
    from vsum_b200.synthetic import make_video
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    sd = SimNet(**MODEL_KW).state_dict()
    sample = [make_video(5000 + i, n, n_users=N_USERS) for i, n in enumerate(REF_SAMPLE_N)]
    return sd, sample, cores


REF_BUDGET_S = 180.0      # the whole reference-arm run (warm-up + timed steps) should end within a few minutes


def pick_sample(sd, full, steps, warmup):
    """Largest of the 12- / 6- / 3-video samples (every 1st / 2nd / 4th length class) whose estimated run time
    fits REF_BUDGET_S; the estimate comes from one untimed pass over the smallest sample (measured cost ratios
    of the three samples: about 9 : 4 : 1; rounded up)."""
    small = full[1::4]
    t0 = time.perf_counter()
    cpu_reference_step(sd, small)
    dt = time.perf_counter() - t0
    n_steps = max(1, steps + warmup)
    if 10.0 * dt * n_steps <= REF_BUDGET_S:
        return full
    if 4.5 * dt * n_steps <= REF_BUDGET_S:
        return full[0::2]
    return small


def run_cpu_baseline(steps=1, warmup=0):
    sd, sample, cores = reference_setup()
    sample = pick_sample(sd, sample, steps, warmup)
    for _ in range(warmup):
        cpu_reference_step(sd, sample)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, sample)
    dt = (time.perf_counter() - t0) / steps
    return {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(sample)} videos, N={[v.n_steps for v in sample]} (evenly spaced classes of the log-uniform length range), "
                      f"{dt:.2f} s/step; torch fp32 CPU scorer ({cores} threads) + pure-Python pooling/knapsack/F-score "
                      "as the reference runs them"}, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = run_cpu_baseline(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.videos), "l2": "n/a (CPU)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def main_b200(args):
    import torch
    import torch.distributed as dist
    from vsum_b200 import _cabi
    from vsum_b200.model import SimNet
    from vsum_b200.pipeline import DeviceBatch, Summarizer, pack_videos
    from vsum_b200.synthetic import make_video, video_length

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"                      # no version banner: stdout is the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: weak scaling, rank r owns videos [r*V, (r+1)*V) of the global id space
    V = args.videos
    ids = [rank * V + i for i in range(V)]
    videos = [make_video(v, video_length(v, args.len_lo, args.len_hi), n_users=N_USERS, with_features=False) for v in ids]
    hb = pack_videos(videos, pin=True, with_features=False)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    feats_dev = torch.rand((hb.n_steps, 1024), device=dev, generator=g)        # rng.random-like features in [0,1)
    hb.features.copy_(feats_dev)                                                # pinned host copy for the e2e leg
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

    # ---- e2e: host (pinned) buffers in, F-scores out, copies inside the timed region
    dbe = [DeviceBatch(hb, dev, pin_meta=True) for _ in range(2)]     # double-buffered device landing zones

    def step_e2e(i):
        return summ.submit_host(dbe[i & 1], 2 + (i & 1))
    for i in range(2):
        step_e2e(i)
    summ.drain(dev)
    e2e_steps = max(2, min(args.steps, 6))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    e2e_value = world * V / (ms_e2e * 1e-3)

    # ---- the same end-to-end pass with the user summaries as the packed dataset stores them (uint8, lossless for
    # the 0/1 rows): a secondary figure -- `e2e` above keeps the h5 files' float32 rows the reference hands over
    from vsum_b200.evaluation import _engine
    from vsum_b200.pipeline import HostBatch
    vs = [videos[i] for i in hb.order]
    meta8 = _engine.HostEvalBatch.build([v.change_points for v in vs], [v.n_frames for v in vs], [v.picks for v in vs],
                                        [v.user_summary.astype(np.uint8) for v in vs])
    hb8 = HostBatch(hb.features, hb.seqlens, hb.cu_steps, meta8, hb.order, hb.names)
    dbe8 = [DeviceBatch(hb8, dev, pin_meta=True) for _ in range(2)]

    def step_e2e_u8(i):
        return summ.submit_host(dbe8[i & 1], 2 + (i & 1))
    for i in range(2):
        step_e2e_u8(i)
    summ.drain(dev)
    ms_e2e_u8 = timed(step_e2e_u8, e2e_steps) / e2e_steps
    h2d_u8 = int(dbe8[0].h2d_bytes)

    # ---- ... and with the features as a `features_bf16` pack stores them (rounded once, offline): half the feature bytes
    feats16 = torch.empty(hb.features.shape, dtype=torch.bfloat16, pin_memory=True)
    feats16.copy_(hb.features)
    hb16 = HostBatch(feats16, hb.seqlens, hb.cu_steps, meta8, hb.order, hb.names)
    del dbe8
    dbe16 = [DeviceBatch(hb16, dev, pin_meta=True) for _ in range(2)]

    def step_e2e_bf16(i):
        return summ.submit_host(dbe16[i & 1], 2 + (i & 1))
    for i in range(2):
        step_e2e_bf16(i)
    summ.drain(dev)
    ms_e2e_bf16 = timed(step_e2e_bf16, e2e_steps) / e2e_steps
    h2d_bf16 = int(dbe16[0].h2d_bytes)

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
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:   # dram bytes per launch from the committed ncu --set full capture of this same workload
            rec = json.load(open(os.path.join(ROOT, "profiles", "r01_attn_ncu_full_default_workload.json")))
            if rec["frames_per_gpu_per_step"] == hb.n_steps:
                traffic, traffic_src = rec["dram_bytes_total"] / 1e9, "profiles/r01_attn_ncu_full_default_workload.json"
        except Exception:
            pass
        att_ms, att_n = prof.get("attention", (0.0, 0))
        att_flops = attention_flops(hb.seqlens)                            # per launch (one layer, whole batch)
        achieved = att_flops / (att_ms / max(att_n, 1) * 1e-3) / 1e12 if att_ms > 0 else None
        kernel_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}
        gpu_ms = sum(kernel_ms.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(V, args.len_lo, args.len_hi), "videos_per_gpu_per_step": V, "frames_per_gpu_per_step": hb.n_steps,
                       "l2": "inputs (fp32 features, %.2f GB/GPU) are larger than L2; no flush needed" % (hb.n_steps * 4096 / 1e9),
                       "parallelism": f"videos sharded over {world} GPU(s), F-score all-gather",
                       "pipelining": "scorer(batch i+1) overlaps pooling/knapsack/F-score(batch i) on a side stream; e2e adds a copy stream"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(db.h2d_bytes) * world,
                    "d2h_bytes_per_step": 8 * V * world, "ms_per_step": ms_e2e},
            "e2e_u8_user_summaries": {"value": world * V / (ms_e2e_u8 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_u8,
                                      "h2d_bytes_per_step": h2d_u8 * world,
                                      "note": "same pass, user summaries as uint8 (PackedDataset user_summary_u8); e2e keeps float32"},
            "e2e_bf16_features": {"value": world * V / (ms_e2e_bf16 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_bf16,
                                  "h2d_bytes_per_step": h2d_bf16 * world,
                                  "note": "same pass from a pack with bf16 features + uint8 user summaries (write_pack features_bf16, "
                                          "user_summary_u8); inputs are rounded once offline, scores stay within the 1e-2 bf16 tolerance"},
            "gpu_launches": launches,
            "roofline": {"kernel": "attn_tc05_kernel (varlen QK^T/softmax/PV, tcgen05)", "bound": "tensor",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
                         "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch_gb": hb.n_steps * (768 + 256) * 2 / 1e9,
                         "peak_source": peak_src,
                         "flops_per_launch": att_flops, "launch_ms": att_ms / max(att_n, 1),
                         "share_of_step": (att_ms / args.steps) / gpu_ms if gpu_ms else None},
            "scorer_tflops": scorer_flops(hb.seqlens) / (sum(v for k, v in kernel_ms.items() if k in (
                "embed_gemm", "qkv_gemm", "attention", "oproj_ln_gemm", "fc1_gemm", "fc2_ln_gemm")) * 1e-3) / 1e12
            if gpu_ms else None,
            "kernel_ms_per_step": kernel_ms,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = run_cpu_baseline(1, 0)
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
    ap.add_argument("--eval-sms", type=int, default=0, help="SMs left to the evaluation stream in pipelined mode (0 = no partition)")
    ap.add_argument("--len-lo", type=int, default=N_LO, help="shortest video (frames); default = BASELINE config 5")
    ap.add_argument("--len-hi", type=int, default=N_HI, help="longest video (frames)")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
