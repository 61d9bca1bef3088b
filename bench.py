#!/usr/bin/env python
"""Benchmark of the early-exit inference hot path (BASELINE.json metric: documents/sec at a fixed
calibrated exit threshold).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU implementation (oracle port)

A step = one pass of the hot path over one batch of synthetic RVL-CDIP-shaped documents
(512 text tokens + boxes + 224x224 image, 16 classes).  Workload at every N: BASELINE.json configs[1] —
LayoutLMv3-base, learned exit gates after every layer (+ the text_visual_concat embedding exit), batch 256
per GPU, bf16 tensor-core math, entropy-threshold policy with per-exit temperature calibration.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "multi-modal-early-exit_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mmee import synth  # noqa: E402
from mmee.calibration import spread_temperatures, thresholds_for  # noqa: E402
from mmee.config import ExitConfig, ModelDims  # noqa: E402

METRIC = "docs/sec @ calibrated exit threshold"
BATCH_PER_GPU = 256
CONF_THRESHOLD = 0.7          # calibrated-confidence level the entropy threshold is derived from
WORKLOAD = ("LayoutLMv3-base, concat exit + learned exit gates after all 12 layers, entropy-threshold policy "
            "with temperature calibration, 512 tokens + 224x224 image, 16 classes, batch 256 per GPU")


def model_setup():
    dims = ModelDims.base()
    ee = ExitConfig.from_dict(dict(exits=["text_visual_concat"] + list(range(1, dims.layers + 1)),
                                   encoder_layer_strategy="gate", inference_strategy="entropy"))
    return dims, ee


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return dict(hbm=j.get("hbm_gbs", 6650.0), tf_burst=j.get("bf16_tflops", 1590.0),
                    tf_sust=j.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """Samples the SM clock + throttle reasons of one GPU while the timed region runs: in-process NVML (cheap enough
    for a 20 ms period, no driver-lock storms when 8 ranks share a box), `nvidia-smi` as the fallback."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index: int, enabled: bool = True):
        self.index = index
        self.enabled = enabled
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.source = None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _run(self):
        try:
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(nm)
                self._stop.wait(0.02)
            return
        except Exception:
            pass
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0]))
                self.max_mhz = float(o[1])
                for nm, v in zip(names, o[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.enabled:
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.enabled:
            self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def layer_flops(dims: ModelDims):
    S, H, I = dims.seq, dims.hidden, dims.inter
    return dict(linear=2.0 * S * (4 * H * H + 2 * H * I), attn=4.0 * S * S * H,
                patch=2.0 * dims.n_patch * (dims.channels * dims.patch ** 2) * H)


def cpu_port_docs_per_sec(dims, ee, sd, n_docs: int, threads: int):
    """The oracle port (reference algorithm, torch CPU fp32, all exits for all docs) on a bounded sample."""
    from oracle import port

    torch.set_num_threads(threads)
    docs = synth.make_docs(dims, n_docs, seed=77, pad=False)
    with torch.no_grad():
        t0 = time.perf_counter()
        port.forward(sd, dims, ee, docs)
        dt = time.perf_counter() - t0
    return n_docs / dt, dt


def run_reference(args):
    """`--impl reference`: the reference's own CPU path for the same workload.  /root/reference (pure Python on
    HF transformers) does not exist on the GPU box, so this times oracle/port.py — validated bit-exact against
    the unmodified reference by tests/test_oracle.py — with all host threads, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dims, ee = model_setup()
    sd = synth.make_state_dict(dims, ee, seed=0)
    threads = os.cpu_count() or 1
    n_docs = 2
    for _ in range(max(args.warmup, 0)):
        cpu_port_docs_per_sec(dims, ee, sd, 1, threads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_port_docs_per_sec(dims, ee, sd, n_docs, threads)
        times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    v = n_docs / (ms / 1000.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "docs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n_docs} documents per step, all exits computed (the reference never skips work)"},
        "cpu_baseline": {"value": v, "unit": "docs/s", "cores": threads, "kind": "port",
                         "sample": f"{n_docs} docs/step x {args.steps} steps, torch CPU fp32"},
        "e2e": {"value": v, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mmee", choices=["mmee", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--threshold", type=float, default=CONF_THRESHOLD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mmee" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mmee.model import B200EEForSequenceClassification

    dims, ee = model_setup()
    B = args.batch
    sd = synth.make_state_dict(dims, ee, seed=0)
    model = B200EEForSequenceClassification(dims, ee, sd, device=local, max_batch=B)
    kind = ee.inference_strategy
    E1 = model.n_exits + 1
    K = dims.n_labels

    # calibration pass (dense) on a held-out synthetic batch: per-exit temperatures, fixed threshold
    cal_docs = synth.make_docs(dims, min(B, 64), seed=12345, pad=False)
    cal = model.infer(**{k: v.to(dev) for k, v in cal_docs.items()}, exit_threshold=-1.0 if kind == "entropy" else 2.0,
                      early_exit=False, return_all=True)
    temps = spread_temperatures(cal.all_exit_logits.cpu().numpy(), kind)
    thr = thresholds_for(kind, args.threshold, K)

    docs = synth.make_docs(dims, B, seed=1 + 1000 * rank, pad=False)
    dev_docs = {k: v.to(dev) for k, v in docs.items()}
    pin_docs = {k: v.pin_memory() for k, v in docs.items()}

    packed_all = torch.empty((world * B, K + 2), dtype=torch.float32, device=dev)

    def step_device():
        if world == 1:
            return model.infer(**dev_docs, exit_threshold=thr, temperatures=temps)
        # data parallel: results stay on the device, one all_gather + one all_reduce, then ONE read-back of the
        # whole job's results (what a caller of the sharded job consumes)
        from mmee.dist import gather_results_fixed

        r = model.infer_device(**dev_docs, exit_threshold=thr, temperatures=temps)
        packed = torch.cat([r["logits"], r["exit_index"].to(torch.float32)[:, None], r["criterion"][:, None]], dim=1)
        job_hist = r["hist"].clone()
        gather_results_fixed(packed, job_hist, packed_all)
        host = packed_all.cpu()                      # synchronises: the step ends when the job's results are on the host
        r["exit_hist"] = r["hist"].cpu().numpy()
        r["job_results"] = host
        return r

    def step_host():
        return model.infer(**pin_docs, exit_threshold=thr, temperatures=temps)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, last

    model.set_profiling(True)
    with ClockSampler(local, enabled=(rank == 0)) as clk:      # rank 0 reports the clocks of its own GPU
        ms_step, res = timed(step_device, args.steps, args.warmup)
    stage = model.last_stage_ms()          # per-stage CUDA-event times of the last timed step
    launches = model.last_launch_count()
    model.set_profiling(False)
    value = world * B / (ms_step / 1000.0)

    e2e = None
    if not args.no_e2e:
        # end to end through the public host-buffer API.  Every step uploads its inputs from pinned host memory and
        # reads its results back inside the timed region.  `value`: the one-call form (infer: upload, forward and
        # read-back back to back; the pixel upload overlaps text embedding + bias build inside the call);
        # `pipelined_value`: infer_submit / infer_collect with two batches in flight (upload of batch n+1 overlaps
        # the forward of batch n), wall clock including pipeline fill and drain.
        n_e2e = max(2, min(args.steps, 5))
        ms_block, res_h = timed(step_host, n_e2e, 2)
        pending = []

        def step_pipelined():
            pending.append(model.infer_submit(**pin_docs, exit_threshold=thr, temperatures=temps))
            if len(pending) == 2:
                return model.infer_collect(pending.pop(0))
            return None

        def drain():
            while pending:
                model.infer_collect(pending.pop(0))

        for _ in range(2):
            step_pipelined()
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_pipelined()
        drain()                                    # every submitted step has been read back when the clock stops
        torch.cuda.synchronize()
        ms_pipe = (time.perf_counter() - t0) * 1000.0 / n_e2e
        tt = torch.tensor([ms_pipe], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_pipe = float(tt.item())
        h2d = sum(int(v.numel() * v.element_size()) for k, v in pin_docs.items() if k != "labels")
        d2h = B * K * 4 + B * 4 + B * 4 + E1 * 8
        e2e = {"value": world * B / (ms_block / 1000.0), "unit": "docs/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_block, "api": "B200EEForSequenceClassification.infer (blocking)",
               "pipelined_value": world * B / (ms_pipe / 1000.0), "pipelined_ms_per_step": ms_pipe,
               "pipelined_api": "infer_submit/infer_collect, 2 batches in flight, wall clock incl. pipeline fill and drain"}

    # ---- SURVEY.md §8(d) primary input recipe: text lengths ~ U{64..512}, padded to 512 (the headline above is its
    # all-512 no-pad variant, the harder case: no key tile is ever skipped)
    padded = None
    if world == 1 and not args.no_e2e:
        pdocs = synth.make_docs(dims, B, seed=7, pad=True)
        pdev = {k: v.to(dev) for k, v in pdocs.items()}
        ms_pad, res_p = timed(lambda: model.infer(**pdev, exit_threshold=thr, temperatures=temps), max(2, min(args.steps, 4)), 2)
        ph = np.asarray(res_p.exit_hist, dtype=np.float64)
        padded = {"value": B / (ms_pad / 1000.0), "unit": "docs/s", "ms_per_step": ms_pad,
                  "mean_real_text_tokens": float(pdocs["attention_mask"].sum(1).float().mean()),
                  "mean_exit_layer": float(np.dot(ph, np.array(model.exit_layers + [dims.layers])) / ph.sum()),
                  "note": "same engine, thresholds and temperatures; padded keys are masked and fully padded key tiles skipped"}

    # ---- roofline of the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of the docs that reached each layer
    hist = (res["exit_hist"] if isinstance(res, dict) else res.exit_hist).astype(np.int64)
    fl = layer_flops(dims)
    reached = []          # documents entering encoder layer l (1-based)
    active = B
    exit_layers = model.exit_layers
    for l in range(1, dims.layers + 1):
        left_before = sum(int(hist[i]) for i, el in enumerate(exit_layers) if el < l)
        reached.append(B - left_before)
    gemm_flops = sum(reached) * fl["linear"] + B * fl["patch"]
    attn_flops = sum(reached) * fl["attn"]
    pk = peaks()
    gemm_tf = gemm_flops / (stage["gemm"] / 1000.0) / 1e12 if stage.get("gemm") else None
    attn_tf = attn_flops / (stage["attention"] / 1000.0) / 1e12 if stage.get("attention") else None
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tj) and B == BATCH_PER_GPU:      # ncu --set full capture of the same workload (see profiles/)
        with open(tj) as f:
            tr = json.load(f)
        traffic = tr.get("gemm_mean_dram_bytes_per_launch")
        traffic_src = f"ncu dram__bytes_read.sum+dram__bytes_write.sum, mean of the 4 GEMM launches of one layer at {tr.get('docs')} docs (profiles/traffic.json, capture {tr.get('capture')})"
    roofline = {"bound": "tensor", "kernel": "gemm_tc_pair_kernel (QKV / out-proj / MLP-up+GELU / MLP-down; tcgen05 cta_group::2)",
                "achieved": gemm_tf, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": (gemm_tf / pk["tf_sust"]) if gemm_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"] + " (sustained: kernel timed inside a long step)",
                "attention": {"achieved": attn_tf, "unit": "TFLOP/s", "frac": (attn_tf / pk["tf_sust"]) if attn_tf else None},
                "stage_ms": stage,
                "whole_step_frac": (gemm_flops + attn_flops) / (ms_step / 1000.0) / 1e12 / pk["tf_sust"]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt = cpu_port_docs_per_sec(dims, ee, sd, 4, threads)
            cpu = {"value": v, "unit": "docs/s", "cores": threads, "kind": "port",
                   "sample": f"4 documents, one dense forward ({dt:.1f} s), torch CPU fp32, all exits computed"}
        mean_depth = float(np.dot(hist, np.array(exit_layers + [dims.layers])) / hist.sum())
        line = {
            "metric": METRIC, "value": value, "unit": "docs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world}",
                       "criterion": kind, "threshold": thr, "calibrated_confidence_level": args.threshold,
                       "temperatures": [round(float(t), 5) for t in temps],
                       "exit_hist_rank0": hist.tolist(), "mean_exit_layer_rank0": mean_depth,
                       "l2_policy": "working set (fp16 attention bias 3.1 GB + activations 2.6 GB per layer) exceeds the 126 MB L2"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "padded_variant": padded,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
