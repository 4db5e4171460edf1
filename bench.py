#!/usr/bin/env python
"""Headline benchmark: audio-seconds encoded per second, raw waveform -> fused encoder states.

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [...]                     the reference's CPU path (oracle port) on host cores
  python bench.py --workload {encoder,train,frontend,large}  make another BASELINE configuration the headline line

Default workload ``encoder`` = BASELINE.json configs[1]: mm_s2ut_transformer base (12 enc layers, d=512, ffn 2048,
8 heads), bf16 operands / fp32 residual stream, encoder + SelectiveAttention fusion forward, batch 64 x 10 s of
synthetic 16 kHz audio + N(0,1) 577x768 image features, per GPU (weak scaling: every rank encodes its own shard of
utterances; the forward path has no collective).  The default run also carries short probes of the other BASELINE
configurations next to the headline, so that the driver's record holds them at every N:

  train_step     configs[2]  forward + backward + NCCL gradient all-reduce + Adam (and the whole model step)
  frontend       configs[3]  fbank + CMVN + Conv1dSubsampler sweep, batch 256, 1-30 s utterances
  large          configs[4]  16 layers, d=1024, DETR 100x256 image features, 40 000 fbank frames per GPU

One JSON line on stdout (rank 0).  The bulky per-kernel table comes FIRST and the results (e2e, train_step, ...) LAST,
because log tails are cut from the front.  See DESIGN.md "Measurement" for what each key means.
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
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH, DUR_S, SR = 64, 10.0, 16000
IMG_TOKENS, IMG_DIM = 577, 768
PRESET = "base"
METRIC = "audio-sec encoded/sec (fbank->fused enc)"
REF_SAMPLE = int(os.environ.get("MM_BENCH_REF_SAMPLE", "8"))    # utterances per step of the reference (CPU) arm


def _peer_exchange_active() -> bool:
    try:
        from mm_s2ut_b200 import peer
        return bool(peer._groups)
    except Exception:
        return False


def _peer_exchange_dtype() -> str:
    try:
        from mm_s2ut_b200 import peer
        return next(iter(peer._groups.values())).exchange_dtype
    except Exception:
        return "fp32"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def num_frames(n: int) -> int:
    return 1 + (n - 400) // 160


def algorithmic_flops(B: int, m: int, d: int, ffn: int, L: int, Tk: int, Dk: int, conv_mid: int = 1024) -> float:
    """conv + L encoder layers + fusion, per batch."""
    T1 = (m - 1) // 2 + 1
    T2 = (T1 - 1) // 2 + 1
    conv = 2 * 5 * 80 * conv_mid * T1 + 2 * 5 * (conv_mid // 2) * (2 * d) * T2
    layer = T2 * (8 * d * d + 4 * d * ffn) + 4 * T2 * T2 * d
    fusion = T2 * 8 * d * d + 4 * Tk * Dk * d + 4 * T2 * Tk * d
    return float(B) * (conv + L * layer + fusion)


def conv_flops(B: int, m: int, d: int, conv_mid: int = 1024) -> float:
    T1 = (m - 1) // 2 + 1
    T2 = (T1 - 1) // 2 + 1
    return float(B) * (2 * 5 * 80 * conv_mid * T1 + 2 * 5 * (conv_mid // 2) * (2 * d) * T2)


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference(sample_utts: int, steps: int, warmup: int, seed: int = 0):
    import torch

    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder
    from oracle import fbank as ofb, fusion as ofu

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    args = make_args(PRESET, multimodal_translation_config_yaml=str(DEFAULT_YAML))
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval()
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    cfg = load_mm_config(DEFAULT_YAML)
    wavs, _ = synth.synth_batch(1, sample_utts, DUR_S, ragged=False)
    imgs = synth.synth_images(1, sample_utts, IMG_TOKENS, IMG_DIM)

    def step():
        feats, flens = ofb.features_from_waveforms(wavs)          # torchaudio fbank + numpy CMVN, per utterance
        with torch.no_grad():
            return ofu.mm_encoder_forward(sd, cfg, torch.from_numpy(feats), torch.from_numpy(flens), [imgs], [None],
                                          args.encoder_attention_heads)

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return {"value": sample_utts * DUR_S / sec, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": f"{sample_utts} x {DUR_S:.0f} s utterances of the same workload (base model, fp32 PyTorch CPU "
                      f"oracle: torchaudio fbank + numpy CMVN + restated fairseq encoder + fusion), median of {steps} "
                      f"steps after {warmup} warm-up ({sum(ts):.1f} s of CPU work)", "ms_per_step": sec * 1e3}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(REF_SAMPLE, max(1, a.steps), max(1, min(a.warmup, 2)))
    cfg = workload_config(a.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"],
        "unit": "audio-s/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n, workload="encoder"):
    return {"workload": f"BASELINE configs[1]: mm_s2ut_transformer base (12 enc layers, d=512, ffn=2048, 8 heads) "
                        f"encoder + SelectiveAttention fusion forward, batch {BATCH} x {DUR_S:.0f} s 16 kHz + "
                        f"{IMG_TOKENS}x{IMG_DIM} image features per GPU",
            "batch_per_gpu": BATCH, "utt_seconds": DUR_S, "image_feats": [IMG_TOKENS, IMG_DIM],
            "sharding": f"utterance batches over {n} GPU(s), no forward collective",
            "l2_policy": "inputs_larger_than_L2 (41 MB waveform + 113 MB image features per step, two input sets alternated)",
            "reference_sample": f"--impl reference times a bounded sample per step: {REF_SAMPLE} of the {BATCH} utterances of a "
                                f"batch on the host cores (rate-normalised: audio-s of the sample / its CPU time)"}


# ---------------------------------------------------------------------------------------------------------
# shared run context
# ---------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        import mm_s2ut_b200  # noqa: F401
        from mm_s2ut_b200 import hostmem

        # pinned staging buffers must live on the GPU's own NUMA node (see hostmem.py): bind before they are allocated
        self.numa_node = hostmem.bind_to_gpu_numa_node(self.local) if os.environ.get("MM_BENCH_NUMA", "1") != "0" else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        """Write a buffer twice the size of the 126 MB L2 (between timed iterations whose inputs would fit in it)."""
        torch = self.torch
        if self.flush_buf is None:
            self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.flush_buf.zero_()

    def timed_replays(self, fns, steps, warmup, flush=False):
        """Device time of `steps` calls (fns alternated), CUDA events on the launch stream, max over ranks.  With
        flush=True every call is bracketed by its own event pair and the L2 is flushed between calls (outside them)."""
        torch = self.torch
        for i in range(warmup):
            fns[i % len(fns)]()
        self.barrier()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fns[i % len(fns)]()
            e1.record()
            self.barrier()
            return self.max_over_ranks(e0.elapsed_time(e1)) / steps
        evs = []
        for i in range(steps):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fns[i % len(fns)]()
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / steps


def make_encoder(ctx, preset=PRESET, cfg_over=None, train=False):
    torch = ctx.torch
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg.update(cfg_over or {})
    torch.manual_seed(0)
    args = make_args(preset, multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).to(ctx.dev)
    return (enc.train() if train else enc.eval()), args


def make_input_sets(ctx, batch=BATCH, dur=DUR_S, img_tokens=IMG_TOKENS, img_dim=IMG_DIM, n_sets=2):
    """Per rank: n_sets pinned host input sets (fp32 waveform in int16 range, lengths, fp32 image features, int16 PCM)."""
    torch = ctx.torch
    from mm_s2ut_b200 import synth

    n_samples = int(dur * SR)
    sets = []
    for s in range(n_sets):
        g = torch.Generator().manual_seed(100 + 10 * ctx.rank + s)
        n_real = 8                                   # 8 distinct synthetic utterances tiled to the batch
        wavs, _ = synth.synth_batch(1 + ctx.rank, n_real, dur, ragged=False)
        wav = torch.stack([torch.from_numpy(wavs[i % n_real]) for i in range(batch)])
        wav = wav * (0.5 + 0.5 * torch.rand(batch, 1, generator=g))          # distinct gains
        img = torch.randn(batch, img_tokens, img_dim, generator=g)
        lens = torch.full((batch,), n_samples, dtype=torch.int64)
        wav = wav.round().clamp_(-32768, 32767)      # 16-bit PCM values, as real audio files hold them
        sets.append(dict(wav=wav.pin_memory(), lens=lens.pin_memory(), img=img.pin_memory(),
                         pcm=wav.to(torch.int16).pin_memory()))
    return sets


# ---------------------------------------------------------------------------------------------------------
# configs[2]: the training step
# ---------------------------------------------------------------------------------------------------------
def train_step_probe(ctx, wav, img, steps=20, warmup=3, model_step=True):
    """BASELINE configs[2]: the training-step variant of the same path at the same shape (forward with activations kept
    + backward + NCCL gradient all-reduce + fairseq Adam), CUDA-graph replay, device timed, max over ranks."""
    import numpy as np

    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.graph import GraphedTrainStep

    enc, args = make_encoder(ctx, cfg_over=dict(modality_dropout=0.5, audio_dropout=-0.5, SA_image_dropout=0.0,
                                                SA_attention_dropout=0.0), train=True)
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = 0.0   # the probe runs without element-wise dropout; modality dropout is on
    gs = GraphedTrainStep(enc, wav.shape[0], wav.shape[1], tuple(img.shape[1:]), overlap_reduce=False)
    gs.wav.copy_(wav)
    gs.img.copy_(img)
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    T = (((1 + (wav.shape[1] - 400) // 160) - 1) // 2 + 1 - 1) // 2 + 1
    gs.grad_out = torch.randn(T, wav.shape[0], enc.embed_dim, device=dev, generator=g) * 1e-3
    n0 = K.launch_count
    gs.capture()
    launches = (K.launch_count - n0) // 2          # eager pass + capture pass over (keep, drop, optimizer)
    rng = np.random.RandomState(0)
    for _ in range(warmup):
        gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gs.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
    e1.record()
    torch.cuda.synchronize()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1) / steps)
    # the exchange alone: the same all-reduce of the flat gradient, timed by itself
    ar_ms = None
    if world > 1:
        ctx.barrier()
        e0.record()
        for _ in range(5):
            gs.eng._reduced = False
            gs.eng.all_reduce_grads()
        e1.record()
        torch.cuda.synchronize()
        ar_ms = ctx.max_over_ranks(e0.elapsed_time(e1) / 5)
    norm = gs.eng.norm_coef[0].item()
    audio_s = wav.shape[0] * DUR_S * world
    gbytes = gs.eng.flat_g.numel() * 4
    out = {"value": audio_s / (ms * 1e-3), "unit": "audio-s/s trained", "ms_per_step": ms, "steps": steps,
           "launches_per_step_approx": launches // 2, "params": int(gs.eng.flat_p.numel()), "grad_norm_last": norm,
           "allreduce_ms": ar_ms, "allreduce_bytes": gbytes if world > 1 else 0,
           "allreduce_busbw_gbs": (2 * (world - 1) / world * gbytes / (ar_ms * 1e-3) / 1e9) if ar_ms else None,
           "collective": (("two-shot in-place exchange kernel over NVLink peer memory (csrc/p2p.cu: device-side barrier, rank r "
                           "sums slice r of every rank's buffer and stores it back to all, barrier)"
                           + (", on a bf16 staging copy of the gradients (the recipe trains with --fp16: fairseq exchanges "
                              "16-bit gradients; fp32 accumulation in rank order, pack / unpack kernels inside the timed "
                              "exchange; MM_P2P_GRAD_DTYPE=fp32 exchanges the fp32 buffer) "
                              if _peer_exchange_dtype() == "bf16" else " ")
                           if _peer_exchange_active() else "one NCCL all-reduce ") +
                          "of the flat fp32 gradient (%d MB) between the backward and the optimizer graphs" % (gbytes >> 20)) if world > 1 else "none (1 GPU)",
           "allreduce_wire_dtype": (_peer_exchange_dtype() if _peer_exchange_active() else "fp32") if world > 1 else None,
           "what": "BASELINE configs[2]: forward (activations kept) + backward of every encoder/fusion/conv parameter + "
                   "gradient all-reduce + fairseq Adam with clip-norm, batch 64 x 10 s per GPU, modality dropout 0.5 "
                   "(image-drop branch), synthetic d loss/d encoder_out; element-wise dropout off"}
    del gs
    torch.cuda.empty_cache()
    # ---- the same step with the dropout of the reference's own recipe (scripts/textless/1_train.sh:112 --dropout 0.1
    # --attention-dropout 0.1 --relu-dropout 0.1; shipped YAML: SA_image_dropout 0.1, SA_attention_dropout 0.1), masks
    # generated inside the fused kernels
    try:
        enc_d, _ = make_encoder(ctx, cfg_over=dict(modality_dropout=0.5, audio_dropout=-0.5, SA_image_dropout=0.1,
                                                   SA_attention_dropout=0.1), train=True)
        enc_d.dropout_p = enc_d.activation_dropout_p = enc_d.attention_dropout_p = 0.1
        enc_d.SA_image_dropout = enc_d.SA_attention_dropout = 0.1
        gd = GraphedTrainStep(enc_d, wav.shape[0], wav.shape[1], tuple(img.shape[1:]), overlap_reduce=False)
        gd.wav.copy_(wav)
        gd.img.copy_(img)
        gd.grad_out = torch.randn(T, wav.shape[0], enc_d.embed_dim, device=dev, generator=g) * 1e-3
        gd.capture()
        for _ in range(warmup):
            gd.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
        ctx.barrier()
        e0.record()
        for _ in range(steps):
            gd.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
        e1.record()
        torch.cuda.synchronize()
        dms = ctx.max_over_ranks(e0.elapsed_time(e1) / steps)
        out["recipe_dropout"] = {
            "value": audio_s / (dms * 1e-3), "unit": "audio-s/s trained", "ms_per_step": dms, "steps": steps,
            "what": "the same step with the reference recipe's dropout on: --dropout 0.1 --attention-dropout 0.1 "
                    "--relu-dropout 0.1, SA_image_dropout 0.1, SA_attention_dropout 0.1 (counter-based masks generated "
                    "inside the fused GEMM / attention kernels and regenerated in the backward pass)"}
        del gd, enc_d
        torch.cuda.empty_cache()
    except Exception as e:
        out["recipe_dropout"] = {"error": f"{type(e).__name__}: {e}"}
    if not model_step:
        return out
    # ---- the complete model step: + 6-layer unit decoder (V = 1004, 500 target units per 10 s) + label-smoothed CE
    try:
        from mm_s2ut_b200.graph import GraphedModelTrainStep
        from mm_s2ut_b200.models.mm_s2ut_model import MM_S2UTTransformerModel

        torch.manual_seed(0)
        model = MM_S2UTTransformerModel(args, target_code_size=1000, build_unused_projections=False).to(dev).train()
        model.encoder.dropout_p = model.encoder.activation_dropout_p = model.encoder.attention_dropout_p = 0.0
        tgt_len = int(50 * DUR_S)
        deng = model.decoder_train_engine()
        deng.dropout_p = deng.attention_dropout_p = deng.activation_dropout_p = 0.0      # like the encoder probe above
        gm = GraphedModelTrainStep(model.encoder, deng, wav.shape[0], wav.shape[1],
                                   tuple(img.shape[1:]), tgt_len, overlap_reduce=False)
        gm.wav.copy_(wav)
        gm.img.copy_(img)
        gm.prev_tokens.copy_(torch.randint(4, 1004, (wav.shape[0], tgt_len), device=dev, generator=g))
        gm.target.copy_(torch.randint(4, 1004, (wav.shape[0], tgt_len), device=dev, generator=g))
        gm.capture()
        for _ in range(warmup):
            gm.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
        ctx.barrier()
        msteps = max(steps // 2, 5)
        e0.record()
        for _ in range(msteps):
            _, (loss, _) = gm.step(5e-4, drop_image=rng.random() < 0.5, clip_norm=10.0)
        e1.record()
        torch.cuda.synchronize()
        mms = ctx.max_over_ranks(e0.elapsed_time(e1) / msteps)
        out["model_step"] = {
            "value": audio_s / (mms * 1e-3), "unit": "audio-s/s trained", "ms_per_step": mms, "steps": msteps,
            "tgt_units_per_utt": tgt_len, "decoder_params": int(gm.dec.flat_p.numel()),
            "loss_per_unit": loss.item() / (wav.shape[0] * tgt_len),
            "what": "the same plus the 6-layer S2UT unit decoder (V = 1004) and fairseq's label-smoothed cross entropy "
                    "(0.2): waveform -> loss -> every parameter gradient -> joint-norm clipping -> Adam on both engines, "
                    "no autograd; random target units"}
        del gm, model
        torch.cuda.empty_cache()
    except Exception as e:
        out["model_step"] = {"error": f"{type(e).__name__}: {e}"}
    del enc
    return out


# ---------------------------------------------------------------------------------------------------------
# configs[3]: front-end sweep (fbank + CMVN + Conv1dSubsampler), batch 256, 1-30 s utterances
# ---------------------------------------------------------------------------------------------------------
def frontend_probe(ctx, enc, steps=10, warmup=3, durations=(1, 2, 5, 10, 20, 30), batch=256):
    torch, dev = ctx.torch, ctx.dev
    from mm_s2ut_b200 import kernels as K

    peaks, _ = _peaks()
    eng = enc.engine()
    rows = []
    for dur in durations:
        n = dur * SR
        m = num_frames(n)
        g = torch.Generator(device=dev).manual_seed(7 + ctx.rank + dur)
        wav = (torch.randn(batch, n, device=dev, generator=g) * 3000).round().clamp_(-32768, 32767).to(torch.int16)
        lens = torch.full((batch,), n, dtype=torch.int64, device=dev)

        def run():
            x1, mm_, seq_lens, _ = eng.frontend(wav, lens)
            eng.subsample(x1, mm_, seq_lens)

        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            run()
            run()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        ms = ctx.timed_replays([graph.replay], steps, warmup, flush=True)
        # fbank alone (the HBM-roofline part): eager launches, L2 flushed between them
        raw = eng.buf("fbank_raw", (batch, m, 80), torch.float32)
        fb_ms = ctx.timed_replays([lambda: K.fbank(wav, lens, raw, eng.fbank_tables)], max(3, steps // 2), 2, flush=True)
        fb_bytes = batch * (2 * n + 320 * m)
        rows.append({"utt_s": dur, "ms": ms, "audio_s_per_s": batch * dur * ctx.world / (ms * 1e-3),
                     "fbank_ms": fb_ms, "fbank_gbs": fb_bytes / (fb_ms * 1e-3) / 1e9,
                     "fbank_frac_hbm": fb_bytes / (fb_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]),
                     "conv_tflops_of_rest": conv_flops(batch, m, eng.d) / (max(ms - fb_ms, 1e-6) * 1e-3) / 1e12})
        del graph, wav, raw
        for key in [k for k in eng._buf if k[0] in ("fbank_raw", "cmvn_mean_std", "x1", "x2", "x", "seq_lens")]:
            del eng._buf[key]     # per-shape workspaces: the next duration allocates its own
        torch.cuda.empty_cache()
    return {"what": "BASELINE configs[3]: fbank + CMVN + Conv1dSubsampler(+GLU, sqrt(d), positions) front-end, batch 256 per "
                    "GPU, int16 PCM in HBM -> fp32 residual stream; CUDA-graph replay, L2 flushed between iterations; "
                    "audio_s_per_s is the whole job over all ranks; fbank_* = the fbank kernel alone against the HBM copy "
                    "peak (algorithmic bytes 2 n + 320 m per utterance)",
            "batch_per_gpu": batch, "sweep": rows}


# ---------------------------------------------------------------------------------------------------------
# configs[4]: large variant (16 layers, d = 1024, DETR 100 x 256 image features, 40 000 fbank frames per GPU)
# ---------------------------------------------------------------------------------------------------------
def large_probe(ctx, steps=10, warmup=3, batch=40, dur=10.0, img_tokens=100, img_dim=256):
    torch, dev = ctx.torch, ctx.dev
    from mm_s2ut_b200.graph import GraphedEncoder

    peaks, _ = _peaks()
    enc, args = make_encoder(ctx, "large", cfg_over=dict(image_feat_dim=[img_dim]))
    n = int(dur * SR)
    ges = []
    for s in range(2):
        g = torch.Generator(device=dev).manual_seed(31 + 10 * ctx.rank + s)
        wav = (torch.randn(batch, n, device=dev, generator=g) * 3000).round().clamp_(-32768, 32767).to(torch.int16)
        img = torch.randn(batch, img_tokens, img_dim, device=dev, generator=g)
        ge = GraphedEncoder(enc, batch, n, [(img_tokens, img_dim)], wav_dtype=torch.int16)
        ge.load_inputs(wav, torch.full((batch,), n, dtype=torch.int64, device=dev), [img])
        ge.capture()
        ges.append(ge)
    ms = ctx.timed_replays([g.replay for g in ges], steps, warmup, flush=True)
    flops = algorithmic_flops(batch, num_frames(n), args.encoder_embed_dim, args.encoder_ffn_embed_dim,
                              args.encoder_layers, img_tokens, img_dim)
    out = {"what": f"BASELINE configs[4]: large variant (16 enc layers, d=1024, ffn=4096, 16 heads), DETR-style "
                   f"{img_tokens}x{img_dim} image features, {batch} x {dur:.0f} s = {batch * num_frames(n)} fbank frames per "
                   f"GPU (max-tokens 40000), bf16 operands; CUDA-graph replay, L2 flushed between iterations",
           "value": batch * dur * ctx.world / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms,
           "frames_per_gpu": batch * num_frames(n), "tensor_flops_per_step": flops,
           "step_tensor_frac_of_peak": flops / (ms * 1e-3) / 1e12 / float(peaks.get("bf16_tflops_sustained", 1400.0))}
    del ges, enc
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------
# same-box GPU yardstick: the restated reference modules in eager PyTorch (bf16 autocast, cuBLASLt + SDPA)
# ---------------------------------------------------------------------------------------------------------
def gpu_eager_baseline(ctx, enc, args, wav_f32, lens, img, steps=10, warmup=3):
    """TEST-INFRASTRUCTURE leg, never on the product path: what stock PyTorch does with the same model on the same GPU.
    fbank = torchaudio.compliance.kaldi.fbank per utterance on the device (what fairseq calls, moved to the GPU) +
    torch CMVN; encoder = F.conv1d / F.glu / F.layer_norm / F.linear / F.scaled_dot_product_attention under
    torch.autocast(bfloat16); fusion = the reference's SelectiveAttention / gate math with torch ops."""
    import math

    torch, dev = ctx.torch, ctx.dev
    F = torch.nn.functional
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    H = args.encoder_attention_heads
    L = args.encoder_layers
    d = args.encoder_embed_dim
    B, n = wav_f32.shape

    def fbank_cmvn():
        try:
            import torchaudio.compliance.kaldi as ta_kaldi
        except Exception:
            return None
        feats = torch.stack([ta_kaldi.fbank(wav_f32[i:i + 1], num_mel_bins=80, sample_frequency=16000.0)
                             for i in range(B)])
        mean = feats.mean(1, keepdim=True)
        var = (feats ** 2).mean(1, keepdim=True) - mean ** 2
        return (feats - mean) / var.clamp_min(1e-10).sqrt()

    pos_cache = {}

    def encoder(feats):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            x = feats.transpose(1, 2)
            for i in range(2):
                x = F.glu(F.conv1d(x, sd[f"subsample.conv_layers.{i}.weight"], sd[f"subsample.conv_layers.{i}.bias"],
                                   stride=2, padding=2), dim=1)
            x = x.transpose(1, 2).float() * math.sqrt(d)          # [B, T, d], fp32 residual stream like ours
            T = x.shape[1]
            if T not in pos_cache:
                from mm_s2ut_b200.models.modules import SinusoidalPositionalEmbedding
                pos_cache[T] = SinusoidalPositionalEmbedding.get_embedding(T + 2, d, 1).to(dev)[2:T + 2]
            x = x + pos_cache[T]
            for i in range(L):
                p = f"transformer_layers.{i}."
                h = F.layer_norm(x, (d,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])
                q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])
                k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
                v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
                q, k, v = (t.view(B, T, H, d // H).transpose(1, 2) for t in (q, k, v))
                a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, d)
                x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"]).float()
                h = F.layer_norm(x, (d,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
                h = F.relu(F.linear(h, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
                x = x + F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"]).float()
            text = F.layer_norm(x, (d,), sd["layer_norm.weight"], sd["layer_norm.bias"])
            im = F.layer_norm(img, (img.shape[-1],), sd["image_pre_norm_module.weight"], sd["image_pre_norm_module.bias"])
            s = "selective_attns.0."
            q = F.linear(text, sd[s + "q_proj.weight"], sd[s + "q_proj.bias"])
            k = F.linear(im, sd[s + "k_proj.weight"], sd[s + "k_proj.bias"])
            v = F.linear(im, sd[s + "v_proj.weight"], sd[s + "v_proj.bias"])
            o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
            o = F.linear(o, sd[s + "proj.weight"], sd[s + "proj.bias"])
            gate = torch.sigmoid(F.linear(torch.cat([o, text.to(o.dtype)], -1), sd["gate_denses.0.weight"],
                                          sd["gate_denses.0.bias"]))
            res = (1 - gate.float()) * text.float() + gate.float() * o.float()
            return res.transpose(0, 1).contiguous()

    with torch.no_grad():
        feats = fbank_cmvn()
        have_fbank = feats is not None
        if not have_fbank:
            feats = torch.randn(B, num_frames(n), 80, device=dev)
        enc_ms = ctx.timed_replays([lambda: encoder(feats)], steps, warmup)
        full_ms = None
        if have_fbank:
            full_ms = ctx.timed_replays([lambda: encoder(fbank_cmvn())], max(2, steps // 3), 1)
    return {"kind": "PyTorch eager on the same GPU (test infrastructure, not the product path): restated reference "
                    "modules under torch.autocast(bfloat16) -- cuBLASLt GEMMs, cuDNN conv, SDPA attention, fp32 residual "
                    "stream -- on the same weights and the same 64 x 10 s batch",
            "encoder_fusion_ms": enc_ms, "encoder_fusion_value": B * DUR_S * ctx.world / (enc_ms * 1e-3),
            "with_torchaudio_fbank_on_gpu_ms": full_ms,
            "with_torchaudio_fbank_on_gpu_value": (B * DUR_S * ctx.world / (full_ms * 1e-3)) if full_ms else None,
            "unit": "audio-s/s"}


# ---------------------------------------------------------------------------------------------------------
# configs[1]: the headline
# ---------------------------------------------------------------------------------------------------------
def run_ours(a):
    ctx = Ctx()
    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    from mm_s2ut_b200 import kernels as K
    from mm_s2ut_b200.feature_store import ImageFeatureStore
    from mm_s2ut_b200.graph import GraphedEncoder

    enc, args = make_encoder(ctx)
    n_samples = int(DUR_S * SR)
    m = num_frames(n_samples)
    host_sets = make_input_sets(ctx)
    dev_sets = [dict(wav=s["wav"].to(dev), lens=s["lens"].to(dev), img=s["img"].to(dev), pcm=s["pcm"].to(dev))
                for s in host_sets]
    img_shapes = [(IMG_TOKENS, IMG_DIM)]

    # ---------------- graphs: fp32 waveform (kernel-only number), int16 PCM host-fed, int16 PCM + feature store ------
    ge = [GraphedEncoder(enc, BATCH, n_samples, img_shapes) for _ in range(2)]
    enc.engine()                                   # operand packing (one-off conversions) is not part of a step
    n0 = K.launch_count
    for j, g in enumerate(ge):
        g.load_inputs(dev_sets[j]["wav"], dev_sets[j]["lens"], [dev_sets[j]["img"]])
        g.capture()
    torch.cuda.synchronize()
    launches_per_fwd = (K.launch_count - n0) // (2 * 3)        # 2 warm-up + 1 capture pass per graph
    ge16 = [GraphedEncoder(enc, BATCH, n_samples, img_shapes, wav_dtype=torch.int16) for _ in range(2)]
    ge16h = [GraphedEncoder(enc, BATCH, n_samples, img_shapes, wav_dtype=torch.int16, img_dtype=torch.float16)
             for _ in range(2)]
    for j in range(2):
        ge16[j].load_inputs(dev_sets[j]["pcm"], dev_sets[j]["lens"], [dev_sets[j]["img"]])
        ge16[j].capture()
        ge16h[j].load_inputs(dev_sets[j]["pcm"], dev_sets[j]["lens"], [dev_sets[j]["img"].half()])
        ge16h[j].capture()
    # the "dataset": the two input sets' 2 x BATCH images held on the GPU in fp16 (the reference's ImageDataset keeps the
    # whole feature tensor of a split in host RAM, data/speech_to_speech_dataset.py:36-68; here it lives in HBM)
    store = ImageFeatureStore(torch.cat([host_sets[0]["img"], host_sets[1]["img"]], 0), dev)
    idx_host = [torch.arange(j * BATCH, (j + 1) * BATCH, dtype=torch.int64).pin_memory() for j in range(2)]
    ges = [GraphedEncoder(enc, BATCH, n_samples, img_shapes, wav_dtype=torch.int16, stores=[store]) for _ in range(2)]
    for j, g in enumerate(ges):
        g.load_inputs(dev_sets[j]["pcm"], dev_sets[j]["lens"], [idx_host[j]])
        g.capture()
    img16_host = [s["img"].half().pin_memory() for s in host_sets]
    torch.cuda.synchronize()

    # ---------------- kernel-only: inputs already resident in HBM, graph replay ----------------
    for i in range(a.warmup):
        ge[i & 1].replay()
    sampler = ClockSampler(ctx.local)
    ctx.barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        ge[i & 1].replay()
    e1.record()
    ctx.barrier()
    ms_dev = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    audio_s = BATCH * DUR_S * a.steps * world
    value = audio_s / (ms_dev * 1e-3)

    # ---------------- e2e: pinned host inputs -> H2D -> forward -> D2H of the result ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    result_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    T_out = ((m - 1) // 2 + 1 - 1) // 2 + 1
    full_host = [torch.empty(T_out, BATCH, args.encoder_embed_dim, dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_loop(graphs, inputs, n, full_d2h=False):
        """inputs(j) -> (wav, lens, [img or index]) pinned host tensors of set j; H2D of step i overlaps step i-1."""
        for j in range(2):
            free[j].record(main)
        for i in range(n):
            j = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[j])
                graphs[j].load_inputs(*inputs(j))
                ready[j].record(copy_stream)
            main.wait_event(ready[j])
            out = graphs[j].replay()
            if full_d2h:                                  # the whole [T, B, d] fp32 result back to pinned host memory
                full_host[j].copy_(out["encoder_out"][0], non_blocking=True)
            else:                                         # the consumer (decoder) is on the device: a checksum comes back
                result_host[j:j + 1].copy_(out["encoder_out"][0].sum().reshape(1), non_blocking=True)
            free[j].record(main)
        main.synchronize()

    def e2e_measure(graphs, inputs, full_d2h=False):
        e2e_loop(graphs, inputs, max(2, a.warmup), full_d2h)
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_loop(graphs, inputs, a.steps, full_d2h)
        ctx.barrier()
        ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
        return audio_s / (ms * 1e-3), ms / a.steps

    nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
    in_store = lambda j: (host_sets[j]["pcm"], host_sets[j]["lens"], [idx_host[j]])
    in_f32 = lambda j: (host_sets[j]["pcm"], host_sets[j]["lens"], [host_sets[j]["img"]])
    in_f16 = lambda j: (host_sets[j]["pcm"], host_sets[j]["lens"], [img16_host[j]])
    e2e_v, e2e_ms = e2e_measure(ges, in_store)
    full_v, full_ms = e2e_measure(ges, in_store, full_d2h=True)
    hf32_v, hf32_ms = e2e_measure(ge16, in_f32)
    hf16_v, hf16_ms = e2e_measure(ge16h, in_f16)
    b_store = nbytes([host_sets[0]["pcm"], host_sets[0]["lens"], idx_host[0]])
    b_f32 = nbytes([host_sets[0]["pcm"], host_sets[0]["lens"], host_sets[0]["img"]])
    b_f16 = nbytes([host_sets[0]["pcm"], host_sets[0]["lens"], img16_host[0]])
    e2e = {
        "value": e2e_v, "unit": "audio-s/s", "h2d_bytes_per_step": b_store, "d2h_bytes_per_step": 4,
        "ms_per_step": e2e_ms,
        "how": "public module API under CUDA-graph replay; per step: pinned-host int16 PCM + lengths + the batch's image "
               "INDEX vector -> cudaMemcpyAsync on a copy stream (double-buffered) -> forward -> checksum D2H.  The image "
               "features live on the GPU in fp16 (ImageFeatureStore = the reference's ImageDataset, which keeps a split's "
               "whole feature tensor in RAM, moved to HBM); hostfed_* = the features shipped from pinned host memory "
               "every step instead (fp32 as the reference's collater does, or fp16)",
        "full_d2h_value": full_v, "full_d2h_ms_per_step": full_ms, "full_d2h_bytes_per_step": full_host[0].numel() * 4,
        "hostfed_fp32_value": hf32_v, "hostfed_fp32_ms_per_step": hf32_ms, "hostfed_fp32_h2d_bytes_per_step": b_f32,
        "hostfed_fp32_h2d_gbs_all_ranks": b_f32 * world / (hf32_ms * 1e-3) / 1e9,
        "hostfed_fp16_value": hf16_v, "hostfed_fp16_ms_per_step": hf16_ms, "hostfed_fp16_h2d_bytes_per_step": b_f16,
        "hostfed_fp16_h2d_gbs_all_ranks": b_f16 * world / (hf16_ms * 1e-3) / 1e9,
    }
    del ge16, ge16h, ges, store
    torch.cuda.empty_cache()

    line = {}
    peaks, peak_src = _peaks()
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops = algorithmic_flops(BATCH, m, args.encoder_embed_dim, args.encoder_ffn_embed_dim, args.encoder_layers,
                              IMG_TOKENS, IMG_DIM)
    if rank == 0:
        # ---------------- instrumented pass: per-kernel device time (CUDA events on the launch stream) ------------
        K.timing = []
        reps = max(3, min(a.steps, 10))
        for i in range(reps):
            # The eager pass issues ~250 host calls per forward (launch + two event records each); park the GPU
            # behind a ~15 ms spin so the host has enqueued the whole forward before the first kernel runs, otherwise
            # the events time host launch latency instead of the kernels.
            if hasattr(torch.cuda, "_sleep"):
                torch.cuda._sleep(30_000_000)
            enc(dev_sets[i & 1]["wav"], dev_sets[i & 1]["lens"], None, None, None, imgs_list=[dev_sets[i & 1]["img"]],
                img_masks_list=[None])
            torch.cuda.synchronize()
        fam = {}
        for name, s0, s1, work in K.timing:
            f = fam.setdefault(name, [0.0, 0.0, 0])
            f[0] += s0.elapsed_time(s1)
            f[1] += work
            f[2] += 1
        K.timing = None
        total_ms = sum(v[0] for v in fam.values())
        TENSOR = ("gemm", "self_attention", "cross_attention")
        kern = {}
        for name, (ms, work, cnt) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
            tensor = name.startswith(TENSOR)
            ach = work / (ms * 1e-3) / (1e12 if tensor else 1e9) if ms > 0 else 0.0
            # [launches per step, ms per step, share of the step, achieved, unit, fraction of the measured peak]
            kern[name] = [cnt // reps, round(ms / reps, 5), round(ms / total_ms, 4), round(ach, 1),
                          "TFLOP/s" if tensor else "GB/s",
                          round(ach / (peak if tensor else float(peaks["hbm_gbs"])), 4)]
        # dominant kernel = gemm_kernel (one template, eight epilogue instantiations: every "gemm[...]" family);
        # the fused GEMM+LayerNorm kernel is a different kernel and is reported next to it
        gemm_ms = sum(v[0] for k, v in fam.items() if k.startswith("gemm["))
        gemm_fl = sum(v[1] for k, v in fam.items() if k.startswith("gemm["))
        gemm_n = sum(v[2] for k, v in fam.items() if k.startswith("gemm["))
        ach = gemm_fl / (gemm_ms * 1e-3) / 1e12
        traffic, traffic_ln, traffic_src = None, None, None
        for tpath in (ROOT / "profiles" / "r02" / "ncu_full_traffic.json", ROOT / "profiles" / "r01" / "ncu_full_v12_traffic.json"):
            if tpath.exists():   # dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture
                tj = json.loads(tpath.read_text())
                traffic = tj["gemm_kernel_per_launch_mean_mb"]["value"] * 1e6
                traffic_ln = tj["gemm_resid_ln_per_launch_mean_mb"]["value"] * 1e6
                traffic_src = f"{tpath.relative_to(ROOT)} (cold-L2 replays: reads are compulsory, writes stay in L2)"
                break
        ln = fam.get("gemm_resid_ln")
        roofline = {
            "bound": "tensor", "kernel": "gemm_kernel (tcgen05/TMA persistent GEMM, all epilogues)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
            "launches_per_step": gemm_n // reps, "share_of_step": gemm_ms / total_ms,
            "how": "algorithmic FLOPs (2*M*N*K per launch) / CUDA-event time per launch, eager instrumented pass "
                   "(each forward enqueued behind a device spin so host launch latency is not timed)",
            "second_kernel_name": "gemm_resid_ln_kernel (GEMM + residual + LayerNorm, 256x512 pair tiles)",
            "second_kernel_achieved": ln[1] / (ln[0] * 1e-3) / 1e12 if ln else None,
            "second_kernel_frac": ln[1] / (ln[0] * 1e-3) / 1e12 / peak if ln else None,
            "second_kernel_share_of_step": ln[0] / total_ms if ln else None,
            "second_kernel_traffic": traffic_ln,
            "step_tensor_frac_of_peak": flops / (ms_dev / a.steps * 1e-3) / 1e12 / peak,
        }
        for k in ("fbank", "cmvn_stats", "cmvn_apply", "layernorm", "softmax_rows"):
            if k in kern:
                roofline[f"hbm_{k}_gbs"] = kern[k][3]
                roofline[f"hbm_{k}_frac"] = kern[k][5]
        for k in ("self_attention", "cross_attention"):
            if k in kern:
                roofline[f"tensor_{k}_tflops"] = kern[k][3]
                roofline[f"tensor_{k}_frac"] = kern[k][5]
        line = {
            "kernels_columns": ["launches_per_step", "ms_per_step", "share", "achieved", "unit", "frac_of_measured_peak"],
            "kernels": kern,
            "metric": METRIC, "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "gpu_launches": launches_per_fwd * a.steps, "launches_per_step": launches_per_fwd,
            "clocks": clocks, "host_numa_node": ctx.numa_node,
            "tensor_flops_per_step": flops,
            "roofline": roofline,
        }

    # ---------------- probes of the other BASELINE configurations (every rank takes part) ----------------
    psteps = max(5, min(a.steps, 20))

    def probe(name, fn):
        if name in a.skip:
            return None
        try:
            return fn()
        except Exception as e:  # the headline line must not depend on a probe
            import traceback

            traceback.print_exc(file=sys.stderr)
            return {"error": f"{type(e).__name__}: {e}"}

    eager = probe("eager", lambda: gpu_eager_baseline(ctx, enc, args, dev_sets[0]["wav"], dev_sets[0]["lens"],
                                                      dev_sets[0]["img"], steps=min(psteps, 10)))
    del ge                      # the captured graphs own the engine's workspaces: gone before those are re-cut
    frontend = probe("frontend", lambda: frontend_probe(ctx, enc, steps=min(psteps, 10)))
    enc._engine = None
    torch.cuda.empty_cache()
    large = probe("large", lambda: large_probe(ctx, steps=min(psteps, 10)))
    train = probe("train", lambda: train_step_probe(ctx, dev_sets[0]["wav"], dev_sets[0]["img"], steps=psteps))
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(16, 30, 2)     # ~10 s of CPU work on the box's cores
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if eager and "encoder_fusion_ms" in eager:
            eager["ours_over_eager"] = eager["encoder_fusion_ms"] / (ms_dev / a.steps)
        line["gpu_eager_baseline"] = eager
        line["frontend"] = frontend
        line["large"] = large
        line["cpu_baseline"] = cpu
        line["e2e"] = e2e
        line["train_step"] = train
        print(json.dumps(line), flush=True)


def run_workload(a):
    """--workload train | frontend | large: that BASELINE configuration alone, as the headline line."""
    ctx = Ctx()
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    enc, args = make_encoder(ctx)
    base = {"n_gpus": world, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic"}
    if a.workload == "train":
        sets = make_input_sets(ctx, n_sets=1)
        r = train_step_probe(ctx, sets[0]["wav"].to(ctx.dev), sets[0]["img"].to(ctx.dev), steps=a.steps, warmup=a.warmup)
        line = dict(base, metric="audio-sec trained/sec (fbank->fused enc fwd+bwd+allreduce+Adam)", value=r["value"],
                    unit="audio-s/s", ms_per_step=r["ms_per_step"], config={"workload": r["what"]}, train_step=r)
    elif a.workload == "frontend":
        r = frontend_probe(ctx, enc, steps=a.steps, warmup=a.warmup)
        ten = next(x for x in r["sweep"] if x["utt_s"] == 10)
        line = dict(base, metric="audio-sec/sec through fbank+CMVN+Conv1dSubsampler (10 s row of the sweep)",
                    value=ten["audio_s_per_s"], unit="audio-s/s", ms_per_step=ten["ms"], dtype="f32",
                    config={"workload": r["what"]}, frontend=r)
    else:
        r = large_probe(ctx, steps=a.steps, warmup=a.warmup)
        line = dict(base, metric=METRIC + ", large variant", value=r["value"], unit="audio-s/s",
                    ms_per_step=r["ms_per_step"], config={"workload": r["what"]}, large=r)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="encoder", choices=["encoder", "train", "frontend", "large"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip", default="", help="comma list of probes to skip in the default run: eager,frontend,large,train")
    ap.add_argument("--no-train-step", action="store_true", help="same as --skip train")
    a = ap.parse_args()
    a.skip = set(filter(None, a.skip.split(",")))
    if a.no_train_step:
        a.skip.add("train")
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner, ...) go to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    buf = []
    real_print = print

    def capture(*args, **kw):
        buf.append(" ".join(str(x) for x in args))

    globals()["print"] = capture
    try:
        if a.impl == "reference":
            run_reference(a)
        elif a.workload == "encoder":
            run_ours(a)
        else:
            run_workload(a)
    finally:
        globals()["print"] = real_print
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    for line in buf:
        real_print(line, flush=True)


if __name__ == "__main__":
    main()
