"""CPU: host-side logic -- config/YAML layer, state_dict contract, frame arithmetic, sharding (incl. a 2-process
gloo run), and that the C-ABI library loads and exports every symbol declared in include/mms2ut_b200.h."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_c_abi_exports_every_header_symbol():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import _lib

    _lib.build()
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 14 and set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.mm_abi_version() == _lib.ABI_VERSION
    n = lib.mm_fbank_table_floats()
    buf = np.zeros(n, dtype=np.float32)
    assert lib.mm_fbank_build_tables(buf.ctypes.data) == 0
    import torchaudio.compliance.kaldi as K
    win = K._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32).numpy()
    assert np.abs(buf[:400] - win).max() < 1e-6          # host-built povey window == torchaudio's
    # argument errors come back as cudaErrorInvalidValue without touching a device
    assert lib.mm_gemm(None, None) == 1 and b"null" in lib.mm_last_error()


def test_mel_table_matches_torchaudio_bank():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import _lib
    import torchaudio.compliance.kaldi as K

    lib = _lib.load()
    t = np.zeros(lib.mm_fbank_table_floats(), dtype=np.float32)
    lib.mm_fbank_build_tables(t.ctypes.data)
    # table layout of csrc/fbank.cu: window | W256 | W512 | mel weights | meta.  The weights are interleaved per group of
    # 16 filters: weight i of filter 16 g + l sits at group_offset(g) + 16 i + l (a half-warp reads 16 consecutive floats)
    gmax_expected = [4, 4, 6, 10, 16]                  # FB_GMAX: the longest filter of each group, rounded up to even
    o, nw = 400 + 512 + 512, 16 * sum(gmax_expected)   # 640
    k0, cnt, off = (t[o + nw + 80 * i: o + nw + 80 + 80 * i].view(np.int32) for i in range(3))
    gmax = t[o + nw + 240: o + nw + 245].view(np.int32)
    assert list(gmax) == gmax_expected
    mine = np.zeros((80, 256), dtype=np.float32)
    goff = 0
    for g in range(5):
        for l in range(16):
            m = 16 * g + l
            assert off[m] == goff + l and cnt[m] <= gmax[g]
            w = t[o + off[m]: o + off[m] + 16 * gmax[g]: 16]          # the filter's gmax[g] slots, stride 16
            mine[m, k0[m]: k0[m] + cnt[m]] = w[: cnt[m]]
            assert not w[cnt[m]:].any()                               # zero padding up to the group's longest filter
        goff += 16 * gmax[g]
    assert goff == nw
    bank, _ = K.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    assert int(cnt.sum()) == int((bank > 0).sum()) == 501
    assert np.abs(mine - bank.numpy()).max() < 2e-6


def test_no_cpu_fallback():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.config import DEFAULT_YAML, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    enc = MM_S2STransformerEncoder(make_args("small", multimodal_translation_config_yaml=str(DEFAULT_YAML)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(2, 16000), torch.tensor([16000, 16000]), None, None, None)


def test_state_dict_contract_and_yaml_keys():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder

    cfg = load_mm_config(DEFAULT_YAML)
    for k in ("SA_image_dropout", "SA_text_dropout", "SA_attention_dropout", "image_pre_norm", "is_fusion_top",
              "image_feat_dim", "modality_dropout", "audio_dropout", "use_selective_gate",
              "multimodal_attention_type", "is_merge_text_img"):
        assert k in cfg
    assert cfg.only_img is None                       # OmegaConf<2.1: missing key reads as None
    enc = MM_S2STransformerEncoder(make_args("base", multimodal_translation_config_yaml=str(DEFAULT_YAML)))
    keys = set(enc.state_dict())
    for k in ("subsample.conv_layers.0.weight", "subsample.conv_layers.1.bias",
              "transformer_layers.11.self_attn.q_proj.weight", "transformer_layers.0.self_attn.out_proj.bias",
              "transformer_layers.3.self_attn_layer_norm.weight", "transformer_layers.3.fc1.weight",
              "transformer_layers.3.fc2.bias", "transformer_layers.3.final_layer_norm.bias", "layer_norm.weight",
              "embed_positions._float_tensor", "selective_attns.0.q_proj.weight", "selective_attns.0.k_proj.bias",
              "selective_attns.0.v_proj.weight", "selective_attns.0.proj.weight", "gate_denses.0.weight",
              "image_pre_norm_module.weight", "proj_768_to_512.weight", "proj_1024_to_512.weight",
              "proj_1024_to_768.weight", "wav2vec2_adaptor.layers.0.weight"):
        assert k in keys, k
    sd = enc.state_dict()
    assert sd["subsample.conv_layers.0.weight"].shape == (1024, 80, 5)
    assert sd["subsample.conv_layers.1.weight"].shape == (1024, 512, 5)
    assert sd["gate_denses.0.weight"].shape == (512, 1024)
    assert sd["selective_attns.0.k_proj.weight"].shape == (512, 768)
    c2 = dict(cfg)
    c2["multimodal_attention_type"] = "multimodal_attention"
    enc2 = MM_S2STransformerEncoder(make_args("small", multimodal_translation_config_yaml=c2),
                                    build_unused_projections=False)
    k2 = set(enc2.state_dict())
    for k in ("multimodal_attns.0.q_proj_weight", "multimodal_attns.0.k_proj_weight", "multimodal_attns.0.v_proj_weight",
              "multimodal_attns.0.in_proj_bias", "multimodal_attns.0.bias_k", "multimodal_attns.0.bias_v",
              "multimodal_attns.0.out_proj.weight"):
        assert k in k2, k
    enc2.load_state_dict(enc2.state_dict())           # round-trips, and drops any packed device weights


def test_frame_arithmetic():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import synth
    from oracle import fbank as ofb, s2t

    for n in (399, 400, 559, 560, 16000, 80000, 160000, 480000):
        m = synth.num_frames(n)
        assert m == (ofb.kaldi_fbank_np(np.zeros(n, np.float32)).shape[0])
        if m:
            assert synth.subsampled_len(m) == int(s2t.out_seq_lens(torch.tensor([m]))[0])
    assert synth.num_frames(160000) == 998 and synth.subsampled_len(998) == 250
    assert synth.subsampled_len(498) == 125 and synth.subsampled_len(2998) == 750


def test_sharding_single_process():
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import sharding as S

    rng = np.random.RandomState(0)
    n_frames = rng.randint(100, 3000, size=500)
    order = S.ordered_indices(n_frames, seed=3)
    assert sorted(order.tolist()) == list(range(500))
    assert all(n_frames[order[i]] >= n_frames[order[i + 1]] for i in range(499))
    batches = S.batch_by_size(order, n_frames, max_tokens=40000)
    assert sorted(i for b in batches for i in b) == list(range(500))
    assert all(len(b) * n_frames[b].max() <= 40000 or len(b) == 1 for b in batches)
    assert S.padding_fraction(batches, n_frames) < 0.05                     # length bucketing keeps padding small
    unsorted = S.batch_by_size(np.arange(500), n_frames, max_tokens=40000)
    assert S.padding_fraction(unsorted, n_frames) > 0.25
    shards = [S.shard_batches(batches, 8, r) for r in range(8)]
    assert len({len(s) for s in shards}) == 1
    assert sorted(i for s in shards for b in s for i in b) == list(range(500))
    assert S.batch_by_size([], n_frames, 100) == []


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MM_ROOT"])
import mm_s2ut_b200
from mm_s2ut_b200 import sharding as S
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_frames = np.random.RandomState(0).randint(100, 3000, size=301)
batches = S.batch_by_size(S.ordered_indices(n_frames, seed=3), n_frames, max_tokens=30000)
mine = S.shard_batches(batches, world, rank)
ids = torch.full((301,), 0, dtype=torch.int64)
for b in mine:
    ids[b] += 1
frames = torch.tensor([float(sum(len(b) * n_frames[b].max() for b in mine if b)), float(len(mine))], dtype=torch.float64)
dist.all_reduce(ids)                          # every utterance is encoded by exactly one rank
allf = [torch.zeros_like(frames) for _ in range(world)]
dist.all_gather(allf, frames)
ok = bool((ids == 1).all()) and len({int(f[1]) for f in allf}) == 1
imb = max(float(f[0]) for f in allf) / (sum(float(f[0]) for f in allf) / world)
if rank == 0:
    print("GLOO_OK" if ok and imb < 1.1 else f"GLOO_FAIL ok={ok} imbalance={imb}")
dist.destroy_process_group()
'''


def test_sharding_two_process_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MM_ROOT=str(ROOT), OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert "GLOO_OK" in r.stdout, r.stdout + r.stderr


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints one JSON line with the tier's keys (tiny run)."""
    import json

    env = dict(os.environ, MM_BENCH_REF_SAMPLE="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=600)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_hostmem_cpulist_and_missing_gpu():
    """NUMA helper: sysfs cpulist parsing; without a driver (this container) it reports "unknown" and binds nothing."""
    import os

    from mm_s2ut_b200 import hostmem

    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    node = hostmem.bind_to_gpu_numa_node(0)
    assert node is None or isinstance(node, int)
    if node is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_committed_bench_line_has_contract_keys():
    """The round's committed default bench line (profiles/rNN/bench_*default_n1.json) carries every key of the bench
    contract, with the roofline fraction consistent with its own achieved / peak."""
    import glob
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import re

    files = sorted(glob.glob(os.path.join(root, "profiles", "r*", "bench_*default_n1.json")),
                   key=lambda f: [int(t) for t in re.findall(r"\d+", f)])   # r01 < r02, v4 < v19
    assert files, "no committed default bench line"
    d = json.load(open(files[-1]))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and r["traffic"] is not None
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]


def test_specaugment_draws_follow_the_reference_call_sequence():
    """data.specaugment draws the masks with the oracle's (= fairseq's) numpy calls in the same order, incl. early exits."""
    import numpy as np

    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.data.specaugment import POLICIES, SpecAugmentTransform
    from oracle.specaugment import spec_augment

    for name, pol in POLICIES.items():
        sa = SpecAugmentTransform.from_policy(name)
        kw = dict(freq_mask_n=pol["freq_mask_N"], freq_mask_f=pol["freq_mask_F"], time_mask_n=pol["time_mask_N"],
                  time_mask_t=pol["time_mask_T"], time_mask_p=pol["time_mask_p"], mask_value=0.0)
        for frames in (998, 37, 3, 0):
            x = np.random.RandomState(frames).randn(frames, 80).astype(np.float32) + 3.0
            ref = spec_augment(x, rng=np.random.RandomState(7), **kw)
            fm, tm = sa.draw(frames, 80, np.random.RandomState(7))
            got = x.copy()
            for f0, f in fm:
                got[:, f0:f0 + f] = 0.0
            for t0, t in tm:
                got[t0:t0 + t] = 0.0
            assert np.array_equal(got, ref), (name, frames)
    tab = SpecAugmentTransform.from_policy("ld").draw_batch([998, 500], 80, np.random.RandomState(0))
    assert tab.shape == (2, 8) and tab.dtype == np.int32
