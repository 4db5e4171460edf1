"""Per-kernel SASS mnemonic counts of libmms2ut_b200.so: what proves the Blackwell-native path (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copy,
HMMA would be the legacy mma.sync path (none), LDGSTS = cp.async.   python profiles/tools/sass_summary.py > profiles/rNN/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[2] / "multimodal-s2ut_b200" / "libmms2ut_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "HGMMA", "LDGSTS", "MUFU.EX2", "SYNCS"]
per = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mm::", "")
        per[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[name]["_total"] += 1
        for w in WANT:
            if op.startswith(w):
                per[name][w] += 1
print(f"cuobjdump -sass {lib.name}: instruction counts per kernel (sm_100a)")
print("%-72s %7s " % ("kernel", "instrs") + " ".join("%8s" % w for w in WANT))
CORE = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "HGMMA"]
tot = collections.Counter()
for k, c in per.items():
    if not any(c[w] for w in CORE):
        continue
    print("%-72s %7d " % (k[:72], c["_total"]) + " ".join("%8d" % c[w] for w in WANT))
    tot.update(c)
print("%-72s %7d " % ("TOTAL (kernels listed)", tot["_total"]) + " ".join("%8d" % tot[w] for w in WANT))
others = [re.sub(r'<.*', '', k) for k, c in per.items() if not any(c[w] for w in CORE)]
print(f"\n{len(others)} CUDA-core kernels without any of these (fbank, CMVN, LayerNorm, softmax, Adam, ...): " + ", ".join(sorted(set(others)))[:1500])
