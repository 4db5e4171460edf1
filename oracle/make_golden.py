"""Generate the committed golden vectors under tests/golden/ (run in the BUILD container only).

Four sources pin the oracle (the reference ships no tests of its own):
  1. the reference's OWN fusion code: /root/reference/mm_s2ut/models/fuse.py is imported through a two-symbol
     fairseq stub (its only fairseq imports are FairseqDataclass and with_incremental_state, fuse.py:13-14) and
     SelectiveAttention / MultimodalAttention are run on seeded inputs; the reference's fuse_img_feat method and the
     fusion-at-top statements of its forward() (modality dropout, per-image-type loop, sum) are cut out of
     mm_s2s_transformer.py with `ast` (the module itself needs fairseq / omegaconf / timm) and run bound to them;
  2. the real torchaudio.compliance.kaldi.fbank (what fairseq's _get_torchaudio_fbank executes) on seeded waveforms;
  3. HF transformers' Speech2TextEncoder -- an independent port of the same fairseq S2T encoder -- with copied weights.
/root/reference does not exist on the GPU box: only the .npz files produced here travel.

    python oracle/make_golden.py
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "tests" / "golden"
REF = Path("/root/reference")


def import_reference_fuse():
    fs = types.ModuleType("fairseq")
    dc = types.ModuleType("fairseq.dataclass")
    inc = types.ModuleType("fairseq.incremental_decoding_utils")

    class FairseqDataclass:  # noqa: D401 - stub
        pass

    dc.FairseqDataclass = FairseqDataclass
    inc.with_incremental_state = lambda cls: cls
    sys.modules.update({"fairseq": fs, "fairseq.dataclass": dc, "fairseq.incremental_decoding_utils": inc})
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_fuse", REF / "mm_s2ut" / "models" / "fuse.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_fuse():
    fuse = import_reference_fuse()
    torch.manual_seed(20251018)
    d, dk, Tq, Tk, B = 128, 192, 19, 61, 3
    sa = fuse.SelectiveAttention(qdim=d, kdim=dk, vdim=dk, attn_dim=d, intermediate_dim=d, output_dim=d, num_heads=1,
                                 attn_drop=0.1).eval()
    with torch.no_grad():
        for p in sa.parameters():          # non-zero biases so every term is pinned
            if p.dim() == 1:
                p.normal_(0, 0.1)
    q = torch.randn(Tq, B, d)
    img = torch.randn(Tk, B, dk)
    mask = torch.zeros(B, Tk, dtype=torch.bool)
    mask[1, 40:] = True
    with torch.no_grad():
        out, attn = sa(q.clone(), img, img, key_padding_mask=None)
        out_m, attn_m = sa(q.clone(), img, img, key_padding_mask=mask)
    np.savez_compressed(OUT / "fuse_selective_attention.npz", q=q.numpy(), img=img.numpy(), mask=mask.numpy(),
                        out=out.numpy(), attn=attn.numpy(), out_masked=out_m.numpy(), attn_masked=attn_m.numpy(),
                        **{"sd." + k: v.numpy() for k, v in sa.state_dict().items()})

    ma = fuse.MultimodalAttention(embed_dim=d, kdim=dk, vdim=dk, num_heads=1, dropout=0.1, add_bias_kv=True).eval()
    with torch.no_grad():
        ma.in_proj_bias.normal_(0, 0.1)
        ma.out_proj.bias.normal_(0, 0.1)
        ma.bias_k.normal_(0, 0.5)
        ma.bias_v.normal_(0, 0.5)
        text_mask = torch.zeros(B, Tq, dtype=torch.bool)
        o1, _ = ma(text=q, text_mask=text_mask, img=img, img_mask=None, is_merge_text_img=False)
        o2, _ = ma(text=q, text_mask=text_mask, img=img, img_mask=mask, is_merge_text_img=False)
    np.savez_compressed(OUT / "fuse_multimodal_attention.npz", q=q.numpy(), img=img.numpy(), mask=mask.numpy(),
                        out=o1.numpy(), out_masked=o2.numpy(),
                        **{"sd." + k: v.numpy() for k, v in ma.state_dict().items()})
    print("fuse goldens written")


def golden_fuse_grads():
    """Gradients of the reference's own fuse.py modules under PyTorch autograd (what the reference's training step
    computes for the fusion block): pins the oracle's differentiable path, which in turn is the yardstick of the CUDA
    backward pass (tests/test_gpu_training.py)."""
    fuse = import_reference_fuse()
    torch.manual_seed(20251019)
    d, dk, Tq, Tk, B = 128, 192, 17, 53, 2
    q = torch.randn(Tq, B, d)
    img = torch.randn(Tk, B, dk)
    R = torch.randn(Tq, B, d)
    sa = fuse.SelectiveAttention(qdim=d, kdim=dk, vdim=dk, attn_dim=d, intermediate_dim=d, output_dim=d, num_heads=1,
                                 attn_drop=0.0).eval()
    with torch.no_grad():
        for p in sa.parameters():
            if p.dim() == 1:
                p.normal_(0, 0.1)
    qg, ig = q.clone().requires_grad_(), img.clone().requires_grad_()
    out, _ = sa(qg, ig, ig, key_padding_mask=None)
    (out * R).sum().backward()
    np.savez_compressed(OUT / "fuse_selective_attention_grads.npz", q=q.numpy(), img=img.numpy(), R=R.numpy(),
                        dq=qg.grad.numpy(), dimg=ig.grad.numpy(),
                        **{"sd." + k: v.detach().numpy() for k, v in sa.state_dict().items()},
                        **{"grad." + k: v.grad.numpy() for k, v in sa.named_parameters()})
    ma = fuse.MultimodalAttention(embed_dim=d, kdim=dk, vdim=dk, num_heads=1, dropout=0.0, add_bias_kv=True).eval()
    with torch.no_grad():
        ma.in_proj_bias.normal_(0, 0.1)
        ma.out_proj.bias.normal_(0, 0.1)
        ma.bias_k.normal_(0, 0.5)
        ma.bias_v.normal_(0, 0.5)
    qg, ig = q.clone().requires_grad_(), img.clone().requires_grad_()
    o, _ = ma(text=qg, text_mask=torch.zeros(B, Tq, dtype=torch.bool), img=ig, img_mask=None, is_merge_text_img=False)
    (o * R).sum().backward()
    np.savez_compressed(OUT / "fuse_multimodal_attention_grads.npz", q=q.numpy(), img=img.numpy(), R=R.numpy(),
                        dq=qg.grad.numpy(), dimg=ig.grad.numpy(),
                        **{"sd." + k: v.detach().numpy() for k, v in ma.state_dict().items()},
                        **{"grad." + k: v.grad.numpy() for k, v in ma.named_parameters()})
    print("fuse gradient goldens written")


def _reference_encoder_methods():
    """``fuse_img_feat`` and ``f`` of the reference's MM_S2STransformerEncoder, plus the fusion-at-top statements of
    its ``forward`` (modality dropout :496-512, the per-image-type loop :513-556 and the sum :557-560), cut out of
    /root/reference/mm_s2ut/models/mm_s2s_transformer.py with ``ast`` -- the module itself cannot be imported here
    (fairseq, omegaconf, timm are absent) but these statements only need torch and numpy."""
    import ast

    src = (REF / "mm_s2ut" / "models" / "mm_s2s_transformer.py").read_text()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "MM_S2STransformerEncoder")
    fns = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}
    ns = {"torch": torch, "np": np}
    mod = ast.Module(body=[fns["fuse_img_feat"], fns["f"]], type_ignores=[])
    exec(compile(mod, "ref:mm_s2s_transformer.py", "exec"), ns)

    # the glue statements inside forward(): find `if self.training and not self.only_img:` and what follows it
    def walk(body):
        for i, st in enumerate(body):
            if isinstance(st, ast.If) and ast.unparse(st.test) == "self.training and (not self.only_img)":
                return body[i:]
            for sub in ("body", "orelse"):
                r = walk(getattr(st, sub, []) or []) if isinstance(getattr(st, sub, None), list) else None
                if r:
                    return r
        return None

    glue = walk(fns["forward"].body)
    assert glue is not None and isinstance(glue[1], ast.For) and isinstance(glue[2], ast.If), "reference layout changed"
    glue_code = compile(ast.Module(body=glue[:3], type_ignores=[]), "ref:mm_s2s_transformer.py:forward", "exec")
    return ns["fuse_img_feat"], ns["f"], glue_code


def golden_fuse_img_feat():
    """Pins oracle.fusion.fuse_img_feat and the modality-dropout / per-type-sum glue of mm_encoder_forward on the
    reference's OWN statements (ast-extracted, see above) bound to the reference's own fuse.py modules."""
    import types as _t

    fuse = import_reference_fuse()
    ref_fuse_img_feat, ref_f, glue_code = _reference_encoder_methods()
    torch.manual_seed(20251020)
    d, T, B = 64, 13, 3
    cases = {}

    def build(kind, dims, gate, pre_norm):
        me = _t.SimpleNamespace()
        me.multimodal_attention_type = kind
        me.use_selective_gate = gate
        me.is_merge_text_img = False
        me.only_img = False
        me.image_dropout_module = torch.nn.Dropout(0.3).eval()        # FairseqDropout in eval mode = identity
        me.text_dropout_module = torch.nn.Dropout(0.3).eval()
        me.image_pre_norm_module = torch.nn.LayerNorm(dims, 1e-5, True) if pre_norm else torch.nn.Identity()
        if kind == "selective_attention":
            me.selective_attns = torch.nn.ModuleList(
                fuse.SelectiveAttention(qdim=d, kdim=i, vdim=i, attn_dim=d, intermediate_dim=d, output_dim=d, num_heads=1,
                                        attn_drop=0.1) for i in dims).eval()
            mods = me.selective_attns
        else:
            me.multimodal_attns = torch.nn.ModuleList(
                fuse.MultimodalAttention(embed_dim=d, kdim=i, vdim=i, num_heads=1, dropout=0.1, add_bias_kv=True)
                for i in dims).eval()
            mods = me.multimodal_attns
        me.gate_denses = torch.nn.ModuleList(torch.nn.Linear(2 * d, d) for _ in dims).eval()
        with torch.no_grad():
            for m in list(mods.parameters()) + list(me.gate_denses.parameters()) + list(me.image_pre_norm_module.parameters()):
                if m.dim() == 1 or m.shape[0] == 1:
                    m.normal_(0, 0.3)
            if pre_norm:
                me.image_pre_norm_module.weight.add_(1.0)
        me.fuse_img_feat = _t.MethodType(ref_fuse_img_feat, me)
        me.f = _t.MethodType(ref_f, me)
        sd = {}
        name = "selective_attns" if kind == "selective_attention" else "multimodal_attns"
        for k, v in mods.state_dict().items():
            sd[f"{name}.{k}"] = v
        for k, v in me.gate_denses.state_dict().items():
            sd[f"gate_denses.{k}"] = v
        for k, v in me.image_pre_norm_module.state_dict().items():
            sd[f"image_pre_norm_module.{k}"] = v
        return me, sd

    for name, kind, dims, tks, gate, pre_norm in [
            ("sa_gate_prenorm", "selective_attention", [96], [37], True, True),
            ("sa_nogate", "selective_attention", [96], [37], False, False),
            ("ma_gate_prenorm", "multimodal_attention", [96], [37], True, True),
            ("sa_two_types", "selective_attention", [96, 48], [37, 21], True, False)]:
        me, sd = build(kind, dims, gate, pre_norm)
        text = torch.randn(T, B, d)
        imgs = [torch.randn(B, tk, dk) for tk, dk in zip(tks, dims)]
        masks = []
        for tk in tks:
            m = torch.zeros(B, tk, dtype=torch.bool)
            m[2, tk // 2:] = True
            masks.append(m)
        tmask = torch.zeros(B, T, dtype=torch.bool)
        tmask[1, 9:] = True
        rec = {"text": text.numpy(), "text_mask": tmask.numpy(), "n_types": np.array(len(dims)),
               "gate": np.array(gate), "pre_norm": np.array(pre_norm), "kind": np.array(kind)}
        for j, (im, mk) in enumerate(zip(imgs, masks)):
            rec[f"img{j}"], rec[f"mask{j}"] = im.numpy(), mk.numpy()
        with torch.no_grad():
            # (1) the method alone, first image type, without and with the image key mask
            r0, _ = me.fuse_img_feat(text, 0, imgs[0].transpose(0, 1), None, text_mask=tmask)
            r1, _ = me.fuse_img_feat(text, 0, imgs[0].transpose(0, 1), masks[0], text_mask=tmask)
            rec["res"], rec["res_masked"] = r0.numpy(), r1.numpy()
            # (2) the glue of forward(): eval; training without a drop; training with the image-drop branch
            for tag, training, draws in [("eval", False, None), ("keep", True, (0.9, 0.9)), ("drop_image", True, (0.1, 0.9))]:
                me.training = training
                me.modality_dropout, me.audio_dropout = 0.5, -0.5
                seq = list(draws or ())

                class _NP:   # the two per-batch np.random.random() draws of :497, replayed
                    class random:   # noqa: N801
                        @staticmethod
                        def random():
                            return seq.pop(0)

                out = {"encoder_out": [text.clone()], "encoder_padding_mask": [tmask.clone()], "encoder_states": []}
                env = {"self": me, "out": out, "imgs_list": [i.clone() for i in imgs], "img_masks_list": list(masks),
                       "xs": [], "idx": 0, "torch": torch, "np": _NP, "img_feat_list": None}
                exec(glue_code, env)
                rec[f"glue_{tag}"] = out["encoder_out"][0].numpy()
        rec.update({"sd." + k: v.numpy() for k, v in sd.items()})
        cases[name] = rec
    for name, rec in cases.items():
        np.savez_compressed(OUT / f"fuse_img_feat_{name}.npz", **rec)
    print("fuse_img_feat / glue goldens written:", ", ".join(cases))


def golden_fbank():
    from mm_s2ut_b200 import synth  # noqa: E402
    import torchaudio.compliance.kaldi as ta_kaldi

    data = {}
    for u, dur in enumerate((0.5, 1.0, 2.37)):
        w = synth.synth_waveform(7, u, dur, ragged=False)
        f = ta_kaldi.fbank(torch.from_numpy(w).unsqueeze(0), num_mel_bins=80, sample_frequency=16000).numpy()
        data[f"wav{u}"] = w
        data[f"fbank{u}"] = f
    z = np.zeros(8000, dtype=np.float32)
    data["wav_zero"] = z
    data["fbank_zero"] = ta_kaldi.fbank(torch.from_numpy(z).unsqueeze(0), num_mel_bins=80,
                                        sample_frequency=16000).numpy()
    np.savez_compressed(OUT / "fbank_torchaudio.npz", **data)
    print("fbank goldens written")


def golden_hf_encoder():
    """HF Speech2TextEncoder (port of fairseq's S2T encoder): copy weights into fairseq names, store in/out."""
    from transformers import Speech2TextConfig
    from transformers.models.speech_to_text.modeling_speech_to_text import Speech2TextEncoder

    torch.manual_seed(7)
    d, L, H, ffn = 64, 2, 2, 128
    cfg = Speech2TextConfig(d_model=d, encoder_layers=L, encoder_attention_heads=H, encoder_ffn_dim=ffn,
                            num_conv_layers=2, conv_kernel_sizes=(5, 5), conv_channels=128, input_feat_per_channel=80,
                            input_channels=1, max_source_positions=600, dropout=0.0, attention_dropout=0.0,
                            activation_dropout=0.0, encoder_layerdrop=0.0, activation_function="relu",
                            scale_embedding=True)
    enc = Speech2TextEncoder(cfg).eval()
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    B, T = 3, 83
    lens = torch.tensor([83, 61, 20])
    feats = torch.randn(B, T, 80)
    for i, n in enumerate(lens):
        feats[i, n:] = 0
    attn = (torch.arange(T)[None, :] < lens[:, None]).long()
    with torch.no_grad():
        out = enc(feats, attention_mask=attn).last_hidden_state        # [B, T', d]
    sd = {}
    hs = enc.state_dict()
    for i in range(2):
        sd[f"subsample.conv_layers.{i}.weight"] = hs[f"conv.conv_layers.{i}.weight"]
        sd[f"subsample.conv_layers.{i}.bias"] = hs[f"conv.conv_layers.{i}.bias"]
    for i in range(L):
        a, b = f"layers.{i}.", f"transformer_layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            for t in ("weight", "bias"):
                sd[f"{b}self_attn.{proj}.{t}"] = hs[f"{a}self_attn.{proj}.{t}"]
        for t in ("weight", "bias"):
            sd[f"{b}self_attn_layer_norm.{t}"] = hs[f"{a}self_attn_layer_norm.{t}"]
            sd[f"{b}final_layer_norm.{t}"] = hs[f"{a}final_layer_norm.{t}"]
            sd[f"{b}fc1.{t}"] = hs[f"{a}fc1.{t}"]
            sd[f"{b}fc2.{t}"] = hs[f"{a}fc2.{t}"]
    sd["layer_norm.weight"], sd["layer_norm.bias"] = hs["layer_norm.weight"], hs["layer_norm.bias"]
    np.savez_compressed(OUT / "hf_speech2text_encoder.npz", feats=feats.numpy(), lens=lens.numpy(), out=out.numpy(),
                        heads=np.array(H), **{"sd." + k: v.numpy() for k, v in sd.items()})
    print("HF encoder golden written", tuple(out.shape))


def golden_hf_decoder():
    """HF Speech2TextDecoder (port of fairseq's TransformerDecoder as the S2T / S2UT models use it: scaled embedding +
    fairseq sinusoidal positions, pre-LN layers with causal self-attention and encoder attention, final LayerNorm, output
    projection tied to the embedding): copy weights into fairseq names, store tokens / encoder states / logits.  Pins
    ``oracle/decoder.py: unit_decoder_forward`` (fairseq itself is not installed here)."""
    from transformers import Speech2TextConfig
    from transformers.models.speech_to_text.modeling_speech_to_text import Speech2TextDecoder

    torch.manual_seed(11)
    d, L, H, ffn, V = 64, 2, 2, 128, 57
    cfg = Speech2TextConfig(vocab_size=V, d_model=d, decoder_layers=L, decoder_attention_heads=H, decoder_ffn_dim=ffn,
                            max_target_positions=200, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
                            decoder_layerdrop=0.0, activation_function="relu", scale_embedding=True, pad_token_id=1,
                            bos_token_id=0, eos_token_id=2, decoder_start_token_id=2)
    dec = Speech2TextDecoder(cfg).eval()
    with torch.no_grad():
        for n, p in dec.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        dec.embed_tokens.weight.mul_(8.0)      # HF initialises with std 0.02: make the token term comparable to the positions
    B, Lt, T = 3, 23, 31
    tokens = torch.randint(4, V, (B, Lt))
    tokens[:, 0] = 2
    tokens[1, 17:] = 1                    # trailing target padding (fairseq pads prev_output_tokens on the right)
    tokens[2, 9:] = 1
    enc_lens = torch.tensor([31, 22, 13])
    enc = torch.randn(B, T, d)
    enc_mask = (torch.arange(T)[None, :] < enc_lens[:, None]).long()
    with torch.no_grad():
        hid = dec(input_ids=tokens, attention_mask=tokens.ne(1).long(), encoder_hidden_states=enc,
                  encoder_attention_mask=enc_mask).last_hidden_state           # [B, Lt, d]
        logits = hid @ dec.embed_tokens.weight.t()                             # tied output projection
    hs = dec.state_dict()
    sd = {"embed_tokens.weight": hs["embed_tokens.weight"]}
    for i in range(L):
        a = f"layers.{i}."
        for att in ("self_attn", "encoder_attn"):
            for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
                for t in ("weight", "bias"):
                    sd[f"{a}{att}.{proj}.{t}"] = hs[f"{a}{att}.{proj}.{t}"]
        for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            for t in ("weight", "bias"):
                sd[f"{a}{ln}.{t}"] = hs[f"{a}{ln}.{t}"]
        for t in ("weight", "bias"):
            sd[f"{a}fc1.{t}"] = hs[f"{a}fc1.{t}"]
            sd[f"{a}fc2.{t}"] = hs[f"{a}fc2.{t}"]
    sd["layer_norm.weight"], sd["layer_norm.bias"] = hs["layer_norm.weight"], hs["layer_norm.bias"]
    np.savez_compressed(OUT / "hf_speech2text_decoder.npz", tokens=tokens.numpy(), enc=enc.numpy(),
                        enc_lens=enc_lens.numpy(), logits=logits.numpy(), heads=np.array(H),
                        **{"sd." + k: v.numpy() for k, v in sd.items()})
    print("HF decoder golden written", tuple(logits.shape))


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    golden_fbank()
    golden_fuse()
    golden_fuse_grads()
    golden_fuse_img_feat()
    golden_hf_encoder()
    golden_hf_decoder()
