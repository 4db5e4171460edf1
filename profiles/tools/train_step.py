"""Training-step variant at BASELINE configs[2]'s shape (base model, 64 x 10 s + 577 x 768 image features per GPU):
forward (activations kept) + backward + gradient all-reduce + fairseq Adam, timed with CUDA events (eager launches
behind a device spin so host latency is not timed); per-kernel-family breakdown of one step.

    python profiles/tools/train_step.py [--steps 10] [--batch 64] [--seconds 10]
    torchrun --nproc-per-node N ... profiles/tools/train_step.py          (adds the NCCL gradient all-reduce)

Element-wise dropout is off by default (--dropout P switches every site on, masks generated in the fused kernels); modality dropout is on at 0.5
(image-drop branch, per-batch numpy draw like the reference).
"""
import argparse
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def model_step(a, enc, eng, wav, lens, imgs, dev, world, rank, rng):
    """Encoder + 6-layer S2UT unit decoder (V = 1004) + label-smoothed CE: the complete configs[2] step."""
    import torch.distributed as dist

    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.decoder_training import UnitDecoderTrainEngine
    from mm_s2ut_b200.graph import GraphedModelTrainStep
    from oracle.decoder import init_decoder        # weights only (seeded fairseq-style init); nothing is computed with it

    B, n = wav.shape
    d, ffn = enc.embed_dim, enc.ffn_dim
    dec = UnitDecoderTrainEngine(init_decoder(d, ffn, 6, 1004, seed=1), enc.num_heads, dev)
    dec.dropout_p = dec.activation_dropout_p = a.dropout
    dec.attention_dropout_p = a.dropout if not a.no_decoder_attention_dropout else 0.0
    gs = GraphedModelTrainStep(enc, dec, B, n, (577, 768), a.tgt_len, overlap_reduce=(world > 1 and not a.no_overlap))
    g = torch.Generator(device=dev).manual_seed(3 + rank)
    gs.wav.copy_(wav)
    gs.img.copy_(imgs)
    gs.prev_tokens.copy_(torch.randint(4, 1004, (B, a.tgt_len), device=dev, generator=g))
    gs.target.copy_(torch.randint(4, 1004, (B, a.tgt_len), device=dev, generator=g))
    gs._fwd_bwd(False)                     # first call: workspaces, kernel attributes
    torch.cuda.synchronize()
    n0 = K.launch_count
    K.timing = []
    gs._fwd_bwd(False)
    torch.cuda.synchronize()
    timing, K.timing = K.timing, None
    launches = K.launch_count - n0
    gs.capture()
    losses = []
    for _ in range(a.warmup):
        _, (loss, _) = gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    losses.append(loss.item())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        _, (loss, _) = gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    e1.record()
    torch.cuda.synchronize()
    losses.append(loss.item())
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        fam = defaultdict(lambda: [0.0, 0])
        for name, s, e, work in timing:
            fam[name][0] += s.elapsed_time(e)
            fam[name][1] += 1
        tot = sum(v[0] for v in fam.values())
        print(f"{'kernel family (forward + backward, eager)':44s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
        for name, (t, c) in sorted(fam.items(), key=lambda kv: -kv[1][0])[:24]:
            print(f"{name:44s} {c:8d} {t:9.3f} {100 * t / tot:6.1f}%")
        ntok = B * a.tgt_len
        line = dict(metric="audio-sec trained/sec (complete model step: encoder + unit decoder + label-smoothed CE)",
                    value=B * a.seconds * world / (ms * 1e-3), unit="audio-s/s", n_gpus=world, steps=a.steps, ms_per_step=ms,
                    fwd_bwd_launches=launches, eager_fwd_bwd_sum_ms=tot, batch_per_gpu=B, utt_seconds=a.seconds,
                    tgt_units_per_utt=a.tgt_len, encoder_params=int(eng.flat_p.numel()), decoder_params=int(dec.flat_p.numel()),
                    loss_per_unit_first=losses[0] / ntok, loss_per_unit_last=losses[-1] / ntok,
                    overlap_reduce=bool(world > 1 and not a.no_overlap),
                    dropout=a.dropout,
                    note=("element-wise dropout off" if a.dropout == 0 else
                          f"dropout / activation-dropout / attention-dropout {a.dropout} in encoder and decoder, masks "
                          "generated inside the fused kernels") + "; modality dropout 0.5; random target units")
        print(json.dumps(line), flush=True)
    if world > 1:
        del gs
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--preset", default="base")
    ap.add_argument("--no-graph", action="store_true", help="eager launches only (for an ncu launch list)")
    ap.add_argument("--no-overlap", action="store_true", help="all-reduce after the backward graph instead of inside it")
    ap.add_argument("--no-decoder-attention-dropout", action="store_true",
                    help="--model: keep the decoder's attention dropout off (its other dropout sites follow --dropout)")
    ap.add_argument("--dropout", type=float, default=0.0, help="value of --dropout / --activation-dropout / "
                    "--attention-dropout / SA_image_dropout / SA_attention_dropout")
    ap.add_argument("--model", action="store_true", help="complete model step: + 6-layer unit decoder + label-smoothed CE")
    ap.add_argument("--tgt-len", type=int, default=500, help="target units per utterance (50 Hz units x 10 s)")
    a = ap.parse_args()
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg.update(modality_dropout=0.5, audio_dropout=-0.5, SA_image_dropout=0.0, SA_attention_dropout=0.0)
    torch.manual_seed(0)
    args = make_args(a.preset, multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).to(dev).train()
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = a.dropout
    enc.SA_image_dropout = enc.SA_attention_dropout = a.dropout
    eng = enc.train_engine()
    B, n = a.batch, int(16000 * a.seconds)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    wav = (torch.randn(B, n, device=dev, generator=g) * 3000).clamp_(-32768, 32767)
    lens = torch.full((B,), n, dtype=torch.int64, device=dev)
    imgs = torch.randn(B, 577, 768, device=dev, generator=g)
    rng = np.random.RandomState(0)
    if a.model:
        return model_step(a, enc, eng, wav, lens, imgs, dev, world, rank, rng)

    def step(grad_out=None):
        drop_image = rng.random() < 0.5 and rng.random() >= -0.5
        out = eng.forward_train(wav, lens, [imgs], [None], drop_image=drop_image)
        y = out["encoder_out"][0]
        if grad_out is None:
            grad_out = torch.randn(y.shape, device=dev, generator=g) * 1e-3
        eng.backward(grad_out)
        ws = eng.all_reduce_grads()
        eng.adam_step(lr=5e-4, betas=(0.9, 0.98), clip_norm=10.0, grad_scale=1.0 / ws)
        return grad_out

    go = step()
    for _ in range(a.warmup):
        step(go)
    torch.cuda.synchronize()
    # ---- CUDA-graph replay of the same step (one graph for forward + backward per dropout outcome, one for the optimizer)
    from mm_s2ut_b200.graph import GraphedTrainStep
    if a.no_graph:
        for _ in range(a.steps):
            step(go)
        torch.cuda.synchronize()
        return
    gs = GraphedTrainStep(enc, B, n, (577, 768), overlap_reduce=(world > 1 and not a.no_overlap))
    gs.wav.copy_(wav)
    gs.img.copy_(imgs)
    gs.grad_out = go.clone()
    gs.capture()
    for _ in range(a.warmup):
        gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(a.steps):
        gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    g1.record()
    torch.cuda.synchronize()
    graph_ms = g0.elapsed_time(g1) / a.steps
    if world > 1:
        t = torch.tensor([graph_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        graph_ms = t.item()
    n0 = K.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(200_000_000)      # ~0.1 s device spin: the launches below queue up behind it
    e0.record()
    for _ in range(a.steps):
        step(go)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (K.launch_count - n0) // a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    # per-family breakdown of one step
    K.timing = []
    step(go)
    torch.cuda.synchronize()
    fam = defaultdict(lambda: [0.0, 0, 0.0])
    for name, s, e, work in K.timing:
        f = fam[name]
        f[0] += s.elapsed_time(e)
        f[1] += 1
        f[2] += work
    K.timing = None
    if rank == 0:
        tot = sum(v[0] for v in fam.values())
        print(f"{'kernel family':24s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'rate':>14s}")
        for name, (t, c, wk) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
            rate = f"{wk / t / 1e9:9.1f} TF/s" if name.startswith("gemm") or "attention" in name else f"{wk / t / 1e6:9.0f} GB/s"
            print(f"{name:24s} {c:8d} {t:9.3f} {100 * t / tot:6.1f}% {rate:>14s}")
        audio_s = B * a.seconds * world
        line = dict(metric="audio-sec trained/sec (fbank -> fused enc fwd + bwd + all-reduce + Adam)",
                    value=audio_s / (graph_ms * 1e-3), unit="audio-s/s", n_gpus=world, steps=a.steps,
                    ms_per_step=graph_ms, eager_ms_per_step=ms,
                    launches_per_step=launches, eager_sum_ms=tot, batch_per_gpu=B, utt_seconds=a.seconds,
                    preset=a.preset, params=int(eng.flat_p.numel()), dropout=a.dropout, overlap_reduce=bool(world > 1 and not a.no_overlap),
                    note="modality dropout 0.5; synthetic d loss/d encoder_out")
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that hold captured NCCL kernels must be gone before the communicator is torn down
        del gs
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
