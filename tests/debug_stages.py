"""Stage-by-stage comparison of engine buffers with the oracle port (developer tool, GPU box).

    python tests/debug_stages.py [tiny|base] [n_docs]

Runs a 1-layer model on UNPADDED documents (the encoder buffers use the ragged row layout: without padding a document's
rows are its 709 tokens in order, so the buffers compare one to one) so the activation buffers hold layer-0
intermediates after the forward.  X0 = dense embedding output, X1 = the same rows after the ragged gather."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multi-modal-early-exit_b200"))

from mmee import synth  # noqa: E402
from mmee.config import ExitConfig, ModelDims  # noqa: E402
from mmee.model import B200EEForSequenceClassification  # noqa: E402
from oracle import port  # noqa: E402


def bf16_to_f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


def rep(name, got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    d = np.abs(got - want)
    print(f"{name:12s} max|d|={np.nanmax(d):.4e} mean|d|={np.nanmean(d):.4e} max|ref|={np.abs(want).max():.3f} "
          f"nan={int(np.isnan(got).sum())} argmax={np.unravel_index(np.nanargmax(d), d.shape)}")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dims = ModelDims.tiny(layers=1) if which == "tiny" else ModelDims.base(layers=1)
    ee = ExitConfig.from_dict(dict(exits=["text_visual_concat", 1], encoder_layer_strategy="ramp",
                                   inference_strategy="max_confidence"))
    sd = synth.make_state_dict(dims, ee, seed=0)
    docs = synth.make_docs(dims, n, seed=3, pad=False)
    S, H, T, h = dims.seq, dims.hidden, dims.n_text, dims.heads
    model = B200EEForSequenceClassification(dims, ee, sd, device=0, max_batch=n)
    dev = {k: v.cuda() for k, v in docs.items()}
    out = model.forward(**dev)
    torch.cuda.synchronize()

    sdf = {k: v.float() for k, v in sd.items()}
    x0 = port.fused_embeddings(sdf, dims, docs)
    parts = {}
    mask = torch.cat([docs["attention_mask"].float(), torch.ones(n, dims.n_vis)], 1)
    ext = (1.0 - mask)[:, None, None, :] * torch.finfo(torch.float32).min
    bias = port.attention_bias(sdf, dims, docs["bbox"])
    y = port.encoder_layer(sdf, dims, 0, x0, bias, ext, parts)

    X0 = bf16_to_f32(model.debug_read("X0", np.uint16, n * S * H)).reshape(n, S, H)
    rep("X0 text", X0[:, :T], x0[:, :T].numpy())
    rep("X0 visual", X0[:, T:], x0[:, T:].numpy())
    pitch = ((S + 63) // 64) * 64
    LOG2E = 1.4426950408889634
    B_ = model.debug_read("BIAS", np.float16, n * h * S * pitch).reshape(n, h, S, pitch).astype(np.float32)
    want_b = (bias / 8.0).numpy() * LOG2E
    keep = (mask.numpy() > 0)[:, None, None, :] & np.ones_like(want_b, dtype=bool)
    rep("bias(fp16, log2 domain, unmasked keys)", np.where(keep, B_[..., :S], 0), np.where(keep, want_b, 0))
    print("bias on masked / padded keys: max", np.where(keep, -1e9, B_[..., :S]).max(), "pitch pad max", B_[..., S:].max())
    QK = bf16_to_f32(model.debug_read("QK", np.uint16, n * S * 2 * H)).reshape(n, S, 2 * H)
    q = parts["q"].transpose(1, 2).reshape(n, S, H).numpy() / 8.0 * LOG2E
    k = parts["k"].transpose(1, 2).reshape(n, S, H).numpy()
    rep("Q*log2e/8", QK[..., :H], q)
    rep("K", QK[..., H:], k)
    kvp = ((S + 127) // 128) * 128
    VT = bf16_to_f32(model.debug_read("VT", np.uint16, n * h * 64 * kvp)).reshape(n, h, 64, kvp)
    rep("V^T", VT[..., :S], parts["v"].transpose(-1, -2).numpy())
    print("V^T pad max", np.abs(VT[..., S:]).max())
    CTX = bf16_to_f32(model.debug_read("CTX", np.uint16, n * S * H)).reshape(n, S, H)
    rep("ctx", CTX, parts["ctx"].numpy())
    A1 = bf16_to_f32(model.debug_read("A1", np.uint16, n * S * H)).reshape(n, S, H)
    rep("attn_out", A1, parts["attn_out"].numpy())
    MID = bf16_to_f32(model.debug_read("MID", np.uint16, n * S * dims.inter)).reshape(n, S, dims.inter)
    rep("mlp gelu", MID, parts["mlp"].numpy())
    ref = port.forward(sd, dims, ee, docs)
    rep("exit logits", out.exit_logits.cpu().numpy(), ref["exit_logits"].numpy())
    print("stage ms", model.last_stage_ms() if False else "", "launches", model.last_launch_count())


if __name__ == "__main__":
    main()
