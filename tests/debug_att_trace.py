"""Developer tool: per-tile clock64 trace of CTA 0 of the first layer's attention kernel (base model, GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "multi-modal-early-exit_b200"))
from mmee import synth
from mmee.config import ExitConfig, ModelDims
from mmee.model import B200EEForSequenceClassification

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dims = ModelDims.base(layers=1)
ee = ExitConfig.from_dict(dict(exits=["text_visual_concat", 1], encoder_layer_strategy="ramp", inference_strategy="max_confidence"))
sd = synth.make_state_dict(dims, ee, seed=0)
docs = synth.make_docs(dims, n, seed=3, pad=False)
model = B200EEForSequenceClassification(dims, ee, sd, device=0, max_batch=n)
dev = {k: v.cuda() for k, v in docs.items()}
model.forward(**dev); torch.cuda.synchronize()
model._lib.mmee_set_profiling(model._h, 2)
model.forward(**dev); torch.cuda.synchronize()
tr = model.debug_read("ATT_TRACE", np.int64, 4096)
sm = tr[:768].reshape(96, 8); mm = tr[1024:1024 + 768].reshape(96, 8); pr = tr[2048:2048 + 768].reshape(96, 8)
t0 = sm[0, 0]
print("softmax warp 2: tile @start | s_full wait, tmem ld, exp+st, o_full wait(+item store), fence+arrive | tile period")
for t in range(1, 50):
    r = sm[t]
    if r[0] == 0: break
    d4 = r[4] if r[4] else r[3]
    print(f"{t:3d} @{r[0]-t0:8d} | {r[1]-r[0]:6d} {r[2]-r[1]:6d} {r[3]-r[2]:6d} {d4-r[3]:6d} {r[5]-d4:6d} | {sm[t+1,0]-r[0] if sm[t+1,0] else 0:6d}")
print("MMA thread: tile | S issue: @start, kv_full wait | PV: @start, p_full wait")
for t in range(1, 50):
    r = mm[t]; print(f"{t:3d} | @{r[0]-t0:8d} {r[1]-r[0]:6d} | @{r[2]-t0:8d} {r[3]-r[2]:6d}")
print("producer: tile | @start, kv_empty wait")
for t in range(1, 50):
    r = pr[t]; print(f"{t:3d} | @{r[0]-t0:8d} {r[1]-r[0]:6d}")
