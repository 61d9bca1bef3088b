"""Developer experiment: one batch of 256 documents as two half-batches on two engines / two streams, so the
HBM-bound kernels of one half (LayerNorm, exits, embeddings) can overlap the tensor-bound GEMMs of the other.
    python tests/debug_dual_engine.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-early-exit_b200"))
import bench  # noqa: E402
from mmee import synth  # noqa: E402
from mmee.calibration import spread_temperatures, thresholds_for  # noqa: E402
from mmee.model import B200EEForSequenceClassification  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    dims, ee = bench.model_setup()
    B = 256
    sd = synth.make_state_dict(dims, ee, seed=0)
    kind = ee.inference_strategy
    full = B200EEForSequenceClassification(dims, ee, sd, device=0, max_batch=B)
    cal_docs = synth.make_docs(dims, 64, seed=12345, pad=False)
    cal = full.infer(**{k: v.to(dev) for k, v in cal_docs.items()}, exit_threshold=-1.0 if kind == "entropy" else 2.0,
                     early_exit=False, return_all=True)
    temps = spread_temperatures(cal.all_exit_logits.cpu().numpy(), kind)
    thr = thresholds_for(kind, bench.CONF_THRESHOLD, dims.n_labels)
    docs = {k: v.to(dev) for k, v in synth.make_docs(dims, B, seed=1, pad=False).items()}

    def timed(fn, steps=10, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def step_full():
        r = full.infer_device(**docs, exit_threshold=thr, temperatures=temps)
        return r["exit_index"].cpu()

    ms_full = timed(step_full)
    ref_exit = step_full()
    print(f"single engine, 256 docs: {ms_full:.2f} ms/step  {B / ms_full * 1000:.0f} docs/s")

    for parts in (2, 3, 4):
        n = B // parts
        engines = [B200EEForSequenceClassification(dims, ee, sd, device=0, max_batch=B - n * (parts - 1) if i == parts - 1 else n)
                   for i in range(parts)]
        streams = [torch.cuda.Stream(dev) for _ in range(parts)]
        bounds = [(i * n, B if i == parts - 1 else (i + 1) * n) for i in range(parts)]
        shards = [{k: v[a:b].contiguous() for k, v in docs.items()} for a, b in bounds]

        def step_split():
            cur = torch.cuda.current_stream(dev)
            outs = []
            for eng, st, sh in zip(engines, streams, shards):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    outs.append(eng.infer_device(**sh, exit_threshold=thr, temperatures=temps))
            for st in streams:
                cur.wait_stream(st)
            return torch.cat([o["exit_index"] for o in outs]).cpu()

        ms = timed(step_split)
        same = bool((step_split() == ref_exit).all())
        print(f"{parts} engines / streams: {ms:.2f} ms/step  {B / ms * 1000:.0f} docs/s  exits equal: {same}")
        for e_ in engines:
            e_.close()


if __name__ == "__main__":
    main()
