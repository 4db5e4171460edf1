"""Eager (no CUDA graph) forwards of the bench workload, for ncu: base model, B=64 x 10 s + 577x768 images."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200.config import DEFAULT_YAML, make_args  # noqa: E402
from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dur", type=float, default=10.0)
ap.add_argument("--profile", default="", help="name:count,... -> cudaProfilerStart/Stop around the first `count` "
                "launches of each named wrapper in the LAST iteration (use with ncu --profile-from-start off)")
a = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda:0")
enc = MM_S2STransformerEncoder(make_args("base", multimodal_translation_config_yaml=str(DEFAULT_YAML)),
                               build_unused_projections=False).eval().to(dev)
n = int(a.dur * 16000)
wav = (torch.randn(a.batch, n, device=dev) * 3000).contiguous()
lens = torch.full((a.batch,), n, dtype=torch.int64, device=dev)
img = torch.randn(a.batch, 577, 768, device=dev)
want = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in a.profile.split(",") if kv}
armed = [False]
if want:
    from mm_s2ut_b200 import kernels as K
    _enter, _exit = K._Launch.__enter__, K._Launch.__exit__

    def enter(self):
        if armed[0] and want.get(self.name, 0) > 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        return _enter(self)

    def exit_(self, *exc):
        r = _exit(self, *exc)
        if armed[0] and want.get(self.name, 0) > 0:
            want[self.name] -= 1
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return r

    K._Launch.__enter__, K._Launch.__exit__ = enter, exit_
for it in range(a.iters):
    armed[0] = it + 1 == a.iters
    out = enc(wav, lens, None, None, None, imgs_list=[img], img_masks_list=[None])
torch.cuda.synchronize()
print("ok", tuple(out["encoder_out"][0].shape), float(out["encoder_out"][0].abs().mean()))
