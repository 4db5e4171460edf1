"""Eager per-kernel timing of other BASELINE configurations (large: 16 layers, d=1024, 100x256 image features;
base with long utterances) to see where the non-headline shapes spend their time."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402
from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args  # noqa: E402
from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder  # noqa: E402

dev = torch.device("cuda:0")


def run(name, preset, B, dur, img_tokens, img_dim, flops_note=""):
    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg["image_feat_dim"] = [img_dim]
    torch.manual_seed(0)
    args = make_args(preset, multimodal_translation_config_yaml=cfg)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False).eval().to(dev)
    n = int(dur * 16000)
    wav = (torch.randn(B, n, device=dev) * 3000).round()
    lens = torch.full((B,), n, dtype=torch.int64, device=dev)
    img = torch.randn(B, img_tokens, img_dim, device=dev)
    for _ in range(2):
        enc(wav, lens, None, None, None, imgs_list=[img], img_masks_list=[None])
    torch.cuda.synchronize()
    K.timing = []
    reps = 3
    for _ in range(reps):
        torch.cuda._sleep(40_000_000)
        enc(wav, lens, None, None, None, imgs_list=[img], img_masks_list=[None])
        torch.cuda.synchronize()
    fam = {}
    for nm, s0, s1, work in K.timing:
        f = fam.setdefault(nm, [0.0, 0.0, 0])
        f[0] += s0.elapsed_time(s1)
        f[1] += work
        f[2] += 1
    K.timing = None
    tot = sum(v[0] for v in fam.values()) / reps
    print(f"== {name}: {B} x {dur:.0f} s, sum of kernel times {tot:.3f} ms -> {B * dur / tot * 1e3:,.0f} audio-s/s (eager, serialised)")
    for nm, (ms, work, cnt) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        tensor = nm.startswith("gemm") or nm == "self_attention"
        ach = work / (ms * 1e-3) / (1e12 if tensor else 1e9) if ms > 0 else 0.0
        print(f"   {nm:18s} {cnt // reps:3d} launches {ms / reps:8.3f} ms  {ach:8.1f} {'TFLOP/s' if tensor else 'GB/s'}")


run("configs[4] large, 40000 frames/GPU", "large", 10, 40.0, 100, 256)
run("configs[4] large, 10 s utterances", "large", 40, 10.0, 100, 256)
run("base, 30 s utterances", "base", 21, 30.0, 577, 768)
