"""Teacher-forced unit-decoder forward + label-smoothed CE at the base shape: 64 utterances x 10 s -> 250 encoder
states, 500 target units each (50 Hz units), 6 layers, d = 512, V = 1004.  Eager launches, CUDA-event timed."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import mm_s2ut_b200  # noqa: E402,F401
from mm_s2ut_b200 import kernels as K  # noqa: E402
from mm_s2ut_b200.decoder import UnitDecoderEngine  # noqa: E402
from oracle import decoder as odec  # noqa: E402  (weights only: seeded fairseq-style init)

dev = torch.device("cuda:0")
B, L, T, d = 64, 500, 250, 512
sd = odec.init_decoder(d, 2048, 6, seed=0)
eng = UnitDecoderEngine(sd, 8, dev)
g = torch.Generator().manual_seed(0)
prev = torch.randint(4, 1004, (B, L), generator=g).to(dev)
tgt = torch.randint(4, 1004, (B * L,), generator=g).to(dev)
enc = (torch.randn(T, B, d, generator=g) * 0.5).to(dev)
mask = torch.zeros(B, T, dtype=torch.bool, device=dev)


def step():
    logits = eng.forward(prev, enc, mask)
    return K.label_smoothed_nll(logits.reshape(B * L, -1), 1004, tgt, 1, 0.2) if logits.is_contiguous() else logits


for _ in range(3):
    step()
torch.cuda.synchronize()
K.timing = []
reps = 5
for _ in range(reps):
    if hasattr(torch.cuda, "_sleep"):
        torch.cuda._sleep(40_000_000)
    eng.forward(prev, enc, mask)
    torch.cuda.synchronize()
fam = {}
for nm, s0, s1, work in K.timing:
    f = fam.setdefault(nm, [0.0, 0.0, 0])
    f[0] += s0.elapsed_time(s1)
    f[1] += work
    f[2] += 1
K.timing = None
tot = sum(v[0] for v in fam.values()) / reps
print(f"decoder forward B={B} L={L} T={T}: sum of kernel times {tot:.3f} ms ({B * L / tot * 1e3:,.0f} target units/s, "
      f"{B * 10.0 / tot * 1e3:,.0f} audio-s/s)")
for nm, (ms, work, cnt) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    tensor = nm.startswith("gemm") or nm == "attention"
    ach = work / (ms * 1e-3) / (1e12 if tensor else 1e9) if ms > 0 else 0.0
    print(f"   {nm:18s} {cnt // reps:3d} launches {ms / reps:8.3f} ms  {ach:8.1f} {'TFLOP/s' if tensor else 'GB/s'}")
