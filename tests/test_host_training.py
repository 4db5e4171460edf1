"""Host logic of the training-step variant, checked on CPU: the backward ORCHESTRATION of ``training.TrainEngine``
(buffers, strides, batch offsets, launch order, flat parameter layout) runs against emulated kernels
(``tests/_emul.py``: the same wrapper signatures restated in PyTorch) and must reproduce PyTorch autograd over the fp32
oracle.  The CUDA kernels themselves are covered by ``test_gpu_training.py``; plus the gradient all-reduce under gloo.
"""
import os
import socket

import pytest
import torch

import _emul


def _emulated(monkeypatch):
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import engine, training

    monkeypatch.setattr(engine, "K", _emul)
    monkeypatch.setattr(training, "K", _emul)
    monkeypatch.setattr(engine.EncoderEngine, "_require_cuda", False)


@pytest.mark.parametrize("attn_type,gate,drop_image", [("selective_attention", True, False),
                                                       ("multimodal_attention", True, False),
                                                       ("selective_attention", False, True)])
def test_backward_orchestration_matches_autograd_oracle(monkeypatch, attn_type, gate, drop_image):
    from test_gpu_training import REL, ZERO, _rel, _train_setup

    _emulated(monkeypatch)
    enc, wav, lens, imgs, R, ref_grads, out_ref, mask = _train_setup(attn_type, gate, B=2, dur=1.0,
                                                                      drop_image=drop_image)
    enc.train()
    eng = enc.train_engine()
    out = eng.forward_train(wav, lens, [imgs], [None], drop_image=drop_image)
    valid = (~mask).t().unsqueeze(-1)
    assert ((out["encoder_out"][0] - out_ref).abs() * valid).max().item() < 2e-2
    eng.backward(R)
    names = dict(enc.named_parameters())
    checked = 0
    for k, gref in ref_grads.items():
        if k not in names or gref.norm() < ZERO:
            continue
        rel = _rel(names[k].grad, gref)
        assert rel < REL, (k, rel)
        checked += 1
    assert checked >= 6 * 15 + 2 + 4
    # accumulate=True adds the same gradient once more
    g1 = eng.flat_g.clone()
    eng.backward(R, accumulate=True)
    assert _rel(eng.flat_g, 2 * g1) < 1e-2
    # Adam step + operand refresh: the linear functional of the output goes down
    l1 = (out["encoder_out"][0] * R).sum().item()
    eng.backward(R)
    eng.adam_step(lr=1e-3, clip_norm=10.0)
    out2 = eng.forward_train(wav, lens, [imgs], [None], drop_image=drop_image)
    assert (out2["encoder_out"][0] * R).sum().item() < l1


def test_autograd_function_boundary(monkeypatch):
    from test_gpu_training import REL, ZERO, _rel, _train_setup

    _emulated(monkeypatch)
    enc, wav, lens, imgs, R, ref_grads, out_ref, mask = _train_setup("selective_attention", True, B=2, dur=1.0)
    enc.train()
    out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    y = out["encoder_out"][0]
    assert y.requires_grad
    (y * R).sum().backward()
    k = "transformer_layers.5.fc2.weight"
    p = dict(enc.named_parameters())[k]
    assert _rel(p.grad, ref_grads[k]) < REL
    for q in enc.parameters():      # an optimizer's zero_grad(set_to_none=True)
        q.grad = None
    out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    (out["encoder_out"][0] * R).sum().backward()
    assert _rel(p.grad, ref_grads[k]) < REL
    out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    (out["encoder_out"][0] * R).sum().backward()                # no zero_grad in between: autograd accumulates
    assert _rel(p.grad, 2 * ref_grads[k]) < REL
    # a torch optimizer steps the fp32 parameters; the next forward re-derives the 16-bit operand copies from them
    opt = torch.optim.SGD(enc.parameters(), lr=1e-2)
    l0 = (out["encoder_out"][0].detach() * R).sum().item()
    opt.step()
    out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    assert (out["encoder_out"][0].detach() * R).sum().item() < l0


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200.training import all_reduce_flat

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    n = all_reduce_flat(flat, bucket_elems=256)
    # default path: the peer-memory exchange applies to NCCL / CUDA only; under gloo it must fall back to one collective
    from mm_s2ut_b200.peer import peer_group

    assert peer_group(flat) is None
    one = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    assert all_reduce_flat(one) == world and torch.equal(one, flat)
    q.put((rank, n, flat.sum().item()))
    dist.destroy_process_group()


def _overlap_worker(rank, world, port, q):
    import sys
    from pathlib import Path

    import torch.distributed as dist

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import _emul as emul
    import mm_s2ut_b200  # noqa: F401
    from mm_s2ut_b200 import engine, training
    from test_gpu_training import _train_setup

    engine.K = training.K = emul
    engine.EncoderEngine._require_cuda = False
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    enc, wav, lens, imgs, R, *_ = _train_setup("selective_attention", True, B=2, dur=1.0)
    enc.train()
    eng = enc.train_engine()
    R = R * (1.0 + rank)                      # different gradients on the two ranks
    eng.forward_train(wav, lens, [imgs], [None])
    eng.backward(R)
    local = eng.flat_g.clone()
    both = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(both, local)
    eng.backward(R, overlap_reduce=True)      # buckets reduced as the backward pass completes them
    ws = eng.all_reduce_grads()               # already reduced: no second collective
    cover = sorted([eng.bucket_conv, eng.bucket_top] + eng.bucket_layers)
    tiled = cover[0][0] == 0 and cover[-1][1] == eng.flat_g.numel() and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    err = (eng.flat_g - sum(both)).abs().max().item() / sum(both).abs().max().item()
    q.put((rank, ws, tiled, err))
    dist.destroy_process_group()


def test_overlapped_bucket_all_reduce_gloo_world2():
    """backward(overlap_reduce=True) under gloo, world_size 2: every bucket is reduced exactly once and the buckets
    tile the flat gradient buffer (the N > 1 path of SURVEY.md 8e, run on CPU with the emulated kernels)."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(60)
    for rank, ws, tiled, err in res:
        assert ws == 2 and tiled and err < 1e-6, (rank, ws, tiled, err)


def test_gradient_all_reduce_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    expect = float(sum(range(1000)) * 3)
    assert [r[1] for r in res] == [2, 2] and all(abs(r[2] - expect) < 1e-3 * expect for r in res)


def _decoder_setup(B=2, L=9, T=13, d=128, heads=2, ffn=256, layers=2, vocab=37, seed=0):
    from oracle import decoder as odec

    g = torch.Generator().manual_seed(seed)
    sd = odec.init_decoder(d, ffn, layers, vocab, seed)
    for k in sd:      # non-trivial biases / LayerNorm affine so that every gradient is exercised
        if k.endswith(".bias") or "layer_norm" in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    prev = torch.randint(4, vocab, (B, L), generator=g)
    target = torch.randint(4, vocab, (B, L), generator=g)
    target[0, -2:] = 1                                     # padded targets are ignored by the criterion
    enc = torch.randn(T, B, d, generator=g)
    mask = torch.zeros(B, T, dtype=torch.bool)
    mask[1, T - 4:] = True
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    encg = enc.clone().requires_grad_(True)
    logits = odec.unit_decoder_forward(sdg, prev, encg, mask, heads)
    loss, nll = odec.label_smoothed_nll_loss(logits, target, 0.2)
    loss.backward()
    ref = {k: v.grad for k, v in sdg.items()}
    return sd, prev, target, enc, mask, heads, logits.detach(), loss.item(), nll.item(), ref, encg.grad


def _check_decoder(eng, setup, REL, ZERO, _rel):
    sd, prev, target, enc, mask, heads, logits_ref, loss_ref, nll_ref, ref, denc_ref = setup
    dev = eng.device
    logits = eng.forward_train(prev.to(dev), enc.to(dev), mask.to(dev))
    assert (logits.cpu() - logits_ref).abs().max().item() < 0.15
    loss, nll, d_enc = eng.loss_backward(target.to(dev), 0.2)
    assert abs(loss.item() - loss_ref) / abs(loss_ref) < 2e-2 and abs(nll.item() - nll_ref) / abs(nll_ref) < 2e-2
    assert tuple(d_enc.shape) == tuple(enc.shape)
    valid = (~mask).t().unsqueeze(-1)
    assert _rel(d_enc.cpu() * valid, denc_ref * valid) < REL
    assert (d_enc.cpu() * ~valid).abs().max().item() == 0.0          # padded encoder states receive no gradient
    worst, checked = 0.0, 0
    for k, got in eng.grads().items():
        if ref[k].norm() < ZERO:
            continue
        r = _rel(got, ref[k])
        worst = max(worst, r)
        assert r < REL, (k, r)
        checked += 1
    assert checked >= len(ref) - 2 * eng.n_layers
    return worst


def test_decoder_backward_orchestration_matches_autograd_oracle(monkeypatch):
    """Unit decoder + label-smoothed CE: forward_train / loss_backward on the emulated kernels against autograd over
    the oracle decoder and criterion (parameter gradients, loss values, d loss / d encoder_out)."""
    from test_gpu_training import REL, ZERO, _rel

    _emulated(monkeypatch)
    import mm_s2ut_b200.decoder_training as dt

    monkeypatch.setattr(dt, "K", _emul)
    monkeypatch.setattr(dt.UnitDecoderTrainEngine, "_require_cuda", False)
    setup = _decoder_setup()
    eng = dt.UnitDecoderTrainEngine(setup[0], setup[5], "cpu")
    _check_decoder(eng, setup, REL, ZERO, _rel)
    # one Adam step lowers the loss on the same batch
    l0 = setup[7]
    eng.adam_step(lr=2e-3, clip_norm=10.0)
    eng.forward_train(setup[1], setup[3], setup[4])
    l1, _, _ = eng.loss_backward(setup[2], 0.2)
    assert l1.item() < l0


def test_backward_orchestration_d512_fused_forward(monkeypatch):
    """d_model = 512: the training forward takes the fused GEMM + residual + LayerNorm path (separate output buffers);
    same autograd check on the emulated kernels."""
    from test_gpu_training import REL, ZERO, _custom_setup, _rel

    _emulated(monkeypatch)
    enc, wav, lens, imgs, R, ref_grads = _custom_setup("base", dict(encoder_layers=2), 50, 768, 2, 1.0)
    enc.train()
    eng = enc.train_engine()
    assert eng.train_fused_ln
    eng.forward_train(wav, lens, [imgs], [None])
    eng.backward(R)
    names = dict(enc.named_parameters())
    for k, gref in ref_grads.items():
        if k in names and gref.norm() >= ZERO:
            assert _rel(names[k].grad, gref) < REL, k


def test_dropout_orchestration_matches_oracle_with_same_masks(monkeypatch):
    """Dropout sites of the training step (embedding, after attention, activation, after fc2, attention probabilities,
    SA_image_dropout, SA_attention_dropout): forward and backward regenerate the same masks; checked on the emulated kernels against autograd
    over the oracle run with those masks."""
    from test_gpu_training import _dropout_parity

    _emulated(monkeypatch)
    _dropout_parity(_emul, torch.device("cpu"), emulated=True)


def _decoder_dropout_parity(K, device, dt, seed=77, **shape):
    """Decoder training step with dropout at every fairseq site, against autograd over the oracle decoder + criterion run
    with the same masks (dumped from the kernels' mask function, re-laid-out to the oracle's tensors)."""
    from oracle import decoder as odec
    from test_gpu_training import REL, ZERO, _rel

    p_drop, p_attn, p_act = 0.1, 0.1, 0.15
    sd, prev, target, enc, mask, heads, *_ = _decoder_setup(**shape)

    def mask_for(site, p, rows, cols):
        ones = torch.ones(rows, cols, dtype=torch.float32, device=device)
        K.dropout(ones, ones, p, seed, site)
        return ones.cpu()

    def drop(site, x):
        kind = site[0]
        if kind in ("self_p", "enc_p"):                      # [B*H, L, Tk]; device [B*H][Lp][Tp]
            BH, L, Tk = x.shape
            Lp, Tp = (L + 63) // 64 * 64, (Tk + 63) // 64 * 64
            m = mask_for(dt.dsite_layer(site[1], 1 if kind == "self_p" else 3), p_attn, BH * Lp, Tp)
            return x * m.view(BH, Lp, Tp)[:, :L, :Tk]
        L, B, C = x.shape                                    # [L, B, C]; device [B * L, C]
        if kind == "embed":
            m = mask_for(dt.DSITE_EMBED, p_drop, B * L, C)
        elif kind == "act":
            m = mask_for(dt.dsite_layer(site[1], 4), p_act, B * L, C)
        else:
            m = mask_for(dt.dsite_layer(site[1], {"self": 0, "enc": 2, "ffn": 5}[kind]), p_drop, B * L, C)
        return x * m.view(B, L, C).transpose(0, 1)

    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    encg = enc.clone().requires_grad_(True)
    logits = odec.unit_decoder_forward(sdg, prev, encg, mask, heads, drop=drop)
    loss_ref, _ = odec.label_smoothed_nll_loss(logits, target, 0.2)
    loss_ref.backward()
    eng = dt.UnitDecoderTrainEngine(sd, heads, device)
    eng.dropout_p, eng.attention_dropout_p, eng.activation_dropout_p = p_drop, p_attn, p_act
    eng.forward_train(prev.to(device), enc.to(device), mask.to(device), dropout_seed=seed)
    loss, _, d_enc = eng.loss_backward(target.to(device), 0.2)
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 2e-2
    valid = (~mask).t().unsqueeze(-1)
    assert _rel(d_enc.cpu() * valid, encg.grad * valid) < REL
    worst = 0.0
    for k, got in eng.grads().items():
        if sdg[k].grad.norm() < ZERO:
            continue
        r = _rel(got, sdg[k].grad)
        worst = max(worst, r)
        assert r < REL, (k, r)
    return worst


def test_decoder_dropout_orchestration_matches_oracle_with_same_masks(monkeypatch):
    _emulated(monkeypatch)
    import mm_s2ut_b200.decoder_training as dt

    monkeypatch.setattr(dt, "K", _emul)
    monkeypatch.setattr(dt.UnitDecoderTrainEngine, "_require_cuda", False)
    _decoder_dropout_parity(_emul, torch.device("cpu"), dt)


def test_backward_without_images_matches_autograd_oracle(monkeypatch):
    """No image features in the batch (the plain S2T branch: fused states = final LayerNorm output): the gradient enters
    through the T x B x C -> token-major copy instead of the fusion backward."""
    from oracle import fbank as ofb, fusion as ofu
    from test_gpu_encoder import _build
    from test_gpu_training import REL, ZERO, _rel

    _emulated(monkeypatch)
    from mm_s2ut_b200 import synth

    enc, args, cfg = _build("small", "selective_attention", True)
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = 0.0
    wavs, _ = synth.synth_batch(9, 2, 1.0, ragged=True)
    feats, flens = ofb.features_from_waveforms(wavs)
    sd = {k: v.detach().clone().float().requires_grad_(v.is_floating_point()) for k, v in enc.state_dict().items()}
    ref = ofu.mm_encoder_forward(sd, cfg, torch.from_numpy(feats), torch.from_numpy(flens), [], [],
                                 args.encoder_attention_heads)
    out_ref, mask = ref["encoder_out"][0], ref["encoder_padding_mask"][0]
    R = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(3)) * (~mask).t().unsqueeze(-1)
    (out_ref * R).sum().backward()
    wav, lens = synth.pad_waveforms(wavs)
    enc.train()
    eng = enc.train_engine()
    eng.flat_g.fill_(7.0)                       # stale gradients of an earlier batch must not survive
    out = eng.forward_train(wav, lens, [], [])
    assert ((out["encoder_out"][0] - out_ref.detach()).abs() * (~mask).t().unsqueeze(-1)).max().item() < 2e-2
    eng.backward(R)
    names = dict(enc.named_parameters())
    checked = 0
    for k, v in sd.items():
        if v.grad is None or k not in names or v.grad.norm() < ZERO:
            continue
        assert _rel(names[k].grad, v.grad) < REL, k
        checked += 1
    assert checked >= 6 * 15 + 2 + 4
    assert all(names[k].grad.abs().max().item() == 0 for k in names if "selective_attns" in k or "gate_denses" in k)


def test_decoder_backward_d512_fused_forward(monkeypatch):
    """d_model = 512: the decoder's training forward runs the fused GEMM + residual + LayerNorm kernel (separate output)."""
    from test_gpu_training import REL, ZERO, _rel

    _emulated(monkeypatch)
    import mm_s2ut_b200.decoder_training as dt

    monkeypatch.setattr(dt, "K", _emul)
    monkeypatch.setattr(dt.UnitDecoderTrainEngine, "_require_cuda", False)
    setup = _decoder_setup(B=2, L=7, T=9, d=512, heads=8, ffn=256, layers=2, vocab=29)
    eng = dt.UnitDecoderTrainEngine(setup[0], setup[5], "cpu")
    _check_decoder(eng, setup, REL, ZERO, _rel)


def test_eval_after_external_optimizer_step_uses_the_updated_weights(monkeypatch):
    """training forward + backward -> torch optimizer step -> eval() forward: the 16-bit operand copies (and the
    re-laid-out conv weights) must be re-derived from the stepped fp32 parameters, not be one step old."""
    from oracle import fbank as ofb, fusion as ofu
    from test_gpu_training import _train_setup

    _emulated(monkeypatch)
    enc, wav, lens, imgs, R, ref_grads, out_ref, mask = _train_setup("selective_attention", True, B=2, dur=1.0)
    enc.train()
    out = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    (out["encoder_out"][0] * R).sum().backward()
    torch.optim.SGD(enc.parameters(), lr=0.005).step()                 # moves the output by O(5): stale operands would show
    enc.eval()
    with torch.no_grad():
        got = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None], return_all_hiddens=True)
    sd = {k: v.detach().clone().float() for k, v in enc.state_dict().items()}
    feats, flens = ofb.features_from_waveforms([w[:int(n)].numpy() for w, n in zip(wav, lens)])
    ref = ofu.mm_encoder_forward(sd, enc.mm_config, torch.from_numpy(feats), torch.from_numpy(flens), [imgs], [None],
                                 enc.num_heads)
    valid = (~ref["encoder_padding_mask"][0]).t().unsqueeze(-1)
    err = ((got["encoder_out"][0] - ref["encoder_out"][0]).abs() * valid).max().item()
    stale = ((out_ref - ref["encoder_out"][0]).abs() * valid).max().item()
    assert stale > 0.1, stale          # the step really moved the output ...
    assert err < 2e-2, err             # ... and the eval forward follows it
    assert len(got["encoder_states"]) == enc.num_layers


def test_second_training_forward_before_backward_raises(monkeypatch):
    """One set of saved activations per engine: backward of an OLDER forward must fail loudly, not return the
    gradients of the newer one.  Also: return_all_hiddens is honoured in training mode."""
    from test_gpu_training import _train_setup

    _emulated(monkeypatch)
    enc, wav, lens, imgs, R, *_ = _train_setup("selective_attention", True, B=2, dur=1.0)
    enc.train()
    o1 = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None], return_all_hiddens=True)
    assert len(o1["encoder_states"]) == enc.num_layers
    assert o1["encoder_states"][0].shape == o1["encoder_out"][0].shape
    o2 = enc(wav, lens, None, None, None, imgs_list=[imgs], img_masks_list=[None])
    with pytest.raises(RuntimeError, match="overwritten by a later training forward"):
        (o1["encoder_out"][0] * R).sum().backward()
    (o2["encoder_out"][0] * R).sum().backward()         # the latest forward is fine


def _multi_setup(kind, dim0=96):
    """Encoder + oracle gradients for the cases the backward pass used to refuse: 'mask' = image key mask,
    'two_types' = image_feat_dim [96, 64] (one attention + gate each, summed; image_pre_norm off as in the
    reference, whose single shared LayerNorm cannot serve two dims), 'store' = the batch lives in the feature store."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2s_transformer import MM_S2STransformerEncoder
    from oracle import fbank as ofb, fusion as ofu

    cfg = dict(load_mm_config(DEFAULT_YAML))
    dims = [dim0, 64] if kind == "two_types" else [dim0]
    cfg.update(image_feat_dim=dims, image_pre_norm=kind != "two_types",
               multimodal_attention_type="multimodal_attention" if kind == "mask" else "selective_attention")
    torch.manual_seed(3)
    args = make_args("small", multimodal_translation_config_yaml=cfg, encoder_layers=2)
    enc = MM_S2STransformerEncoder(args, build_unused_projections=False)
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = 0.0
    enc.SA_image_dropout = enc.SA_attention_dropout = enc.SA_text_dropout = 0.0
    B = 2
    wavs, _ = synth.synth_batch(4, B, 1.0, ragged=True)
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randn(B, tk, dk, generator=g) for tk, dk in zip([37, 21], dims)]
    masks = [None for _ in dims]
    if kind == "mask":
        m = torch.zeros(B, 37, dtype=torch.bool)
        m[1, 20:] = True
        m[0, ::5] = True
        masks = [m]
    if kind == "store":
        imgs = [imgs[0].half().float()]          # what a 16-bit store holds
    sd = {k: v.detach().clone().float().requires_grad_(v.is_floating_point()) for k, v in enc.state_dict().items()}
    feats, flens = ofb.features_from_waveforms(wavs)
    ref = ofu.mm_encoder_forward(sd, load_mm_config(cfg), torch.from_numpy(feats), torch.from_numpy(flens), imgs, masks,
                                 args.encoder_attention_heads)
    out_ref, pmask = ref["encoder_out"][0], ref["encoder_padding_mask"][0]
    R = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(11)) * (~pmask).t().unsqueeze(-1)
    (out_ref * R).sum().backward()
    ref_grads = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    wav, lens = synth.pad_waveforms(wavs)
    return enc, wav, lens, imgs, masks, R, ref_grads, out_ref.detach(), pmask


@pytest.mark.parametrize("kind", ["mask", "two_types", "store"])
def test_backward_with_key_mask_several_image_types_and_store_batches(monkeypatch, kind):
    from test_gpu_training import REL, ZERO, _rel

    _emulated(monkeypatch)
    enc, wav, lens, imgs, masks, R, ref_grads, out_ref, pmask = _multi_setup(kind)
    enc.train()
    eng = enc.train_engine()
    feed = imgs
    if kind == "store":
        from mm_s2ut_b200.feature_store import ImageFeatureStore, StoredImages

        monkeypatch.setattr(ImageFeatureStore, "__init__", lambda self, feats, device, dtype=torch.float16, chunk=1024:
                            self.__dict__.update(device=torch.device("cpu"), tokens=feats.shape[1], dim=feats.shape[2],
                                                 data=feats.to(dtype)))
        store = ImageFeatureStore(torch.cat([imgs[0].flip(0), imgs[0]], 0), "cpu")
        feed = [StoredImages(store, torch.tensor([2, 3]))]
    out = eng.forward_train(wav, lens, feed, masks)
    valid = (~pmask).t().unsqueeze(-1)
    assert ((out["encoder_out"][0] - out_ref).abs() * valid).max().item() < 3e-2
    eng.backward(R)
    names = dict(enc.named_parameters())
    checked = 0
    for k, gref in ref_grads.items():
        if k not in names or gref.norm() < ZERO:
            continue
        rel = _rel(names[k].grad, gref)
        assert rel < REL, (kind, k, rel)
        checked += 1
    want = {"mask": 10, "two_types": 2 * 9, "store": 11}[kind]          # fusion tensors with a non-zero gradient (k bias: 0)
    assert sum(1 for k in ref_grads if ("attns" in k or "gate" in k or "pre_norm" in k) and ref_grads[k].norm() >= ZERO) >= want
    assert checked >= 2 * 15 + 2 + want
