"""Unit decoder + label-smoothed cross entropy training step on the CUDA kernels, and the complete configs[2] chain:
waveform -> encoder (activations kept) -> fused states -> decoder -> loss -> decoder backward -> d loss / d encoder_out
-> encoder backward, against PyTorch autograd over the fp32 oracle (encoder + decoder + criterion)."""
import pytest
import torch

from _util import record
from test_gpu_training import REL, ZERO, _rel
from test_host_training import _check_decoder, _decoder_setup

pytestmark = pytest.mark.gpu


def test_ce_and_embedding_backward_kernels(cuda):
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(0)
    rows, vocab, ld = 77, 1004, 1008
    logits = (torch.randn(rows, ld, generator=g) * 3).requires_grad_()
    target = torch.randint(0, vocab, (rows,), generator=g)
    target[5] = 1
    lp = torch.log_softmax(logits[:, :vocab], -1)
    pad = target.eq(1)
    eps, eps_i = 0.2, 0.2 / (vocab - 1)
    loss = (1 - eps - eps_i) * (-lp.gather(1, target[:, None])).squeeze(1).masked_fill(pad, 0).sum() + \
        eps_i * (-lp.sum(1)).masked_fill(pad, 0).sum()
    loss.backward()
    dl = torch.full((rows, ld), 9.0, dtype=torch.bfloat16, device=cuda)
    K.label_smoothed_nll_bwd(logits.detach().cuda(), vocab, target.cuda(), 1, eps, dl, grad_scale=0.5)
    assert torch.allclose(dl.float().cpu(), 0.5 * logits.grad, atol=4e-3, rtol=1e-2)
    assert (dl[:, vocab:] == 0).all() and (dl[5] == 0).all()
    # embedding backward
    tokens = torch.randint(0, 50, (6, 11), generator=g)
    dx = torch.randn(66, 64, generator=g)
    tab = torch.zeros(50, 64, device=cuda)
    K.embed_tokens_bwd(tokens.cuda(), 1, dx.cuda(), 3.0, tab)
    ref = torch.zeros(50, 64)
    keep = tokens.view(-1).ne(1)
    ref.index_add_(0, tokens.view(-1)[keep], 3.0 * dx[keep])
    assert torch.allclose(tab.cpu(), ref, atol=1e-4)


def test_decoder_training_step_matches_autograd_oracle(cuda):
    from mm_s2ut_b200.decoder_training import UnitDecoderTrainEngine

    setup = _decoder_setup(B=3, L=40, T=50, d=256, heads=4, ffn=512, layers=2, vocab=104, seed=1)
    eng = UnitDecoderTrainEngine(setup[0], setup[5], cuda)
    worst = _check_decoder(eng, setup, REL, ZERO, _rel)
    record("configs[2] decoder + label-smoothed CE backward (2 layers, d=256, B=3 x 40 units, 50 encoder states): worst "
           "parameter-gradient relative L2 error", worst, REL)


def test_full_model_training_step_chain(cuda):
    """waveform -> encoder -> decoder -> loss -> backward through both engines == autograd over oracle encoder +
    oracle decoder + criterion (the complete configs[2] step, no autograd on the product side)."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.decoder_training import UnitDecoderTrainEngine
    from oracle import decoder as odec, fbank as ofb, fusion as ofu
    from test_gpu_encoder import _build

    enc, args, cfg = _build("small", "selective_attention", True)
    enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = 0.0
    enc.SA_image_dropout = enc.SA_attention_dropout = enc.SA_text_dropout = 0.0
    B, Lu, heads = 3, 30, args.encoder_attention_heads
    d = args.encoder_embed_dim
    wavs, _ = synth.synth_batch(3, B, 2.0, ragged=True)
    imgs = synth.synth_images(3, B)
    dsd = odec.init_decoder(d, 512, 2, 104, seed=2)
    g = torch.Generator().manual_seed(9)
    prev = torch.randint(4, 104, (B, Lu), generator=g)
    target = torch.randint(4, 104, (B, Lu), generator=g)
    # ---- oracle: autograd through encoder + decoder + criterion
    esd = {k: v.detach().clone().float().requires_grad_(v.is_floating_point()) for k, v in enc.state_dict().items()}
    dsg = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
    feats, flens = ofb.features_from_waveforms(wavs)
    ref = ofu.mm_encoder_forward(esd, cfg, torch.from_numpy(feats), torch.from_numpy(flens), [imgs], [None], heads)
    mask = ref["encoder_padding_mask"][0]
    logits = odec.unit_decoder_forward(dsg, prev, ref["encoder_out"][0], mask, heads)
    loss_ref, _ = odec.label_smoothed_nll_loss(logits, target, 0.2)
    loss_ref.backward()
    # ---- product: two engines chained by d loss / d encoder_out
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda().train()
    eeng = enc.train_engine()
    deng = UnitDecoderTrainEngine(dsd, heads, cuda)
    out = eeng.forward_train(wav.cuda(), lens.cuda(), [imgs.cuda()], [None])
    deng.forward_train(prev.cuda(), out["encoder_out"][0], out["encoder_padding_mask"][0])
    loss, nll, d_enc = deng.loss_backward(target.cuda(), 0.2)
    eeng.backward(d_enc)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 2e-2
    worst = 0.0
    names = dict(enc.named_parameters())
    for k, v in esd.items():
        if v.grad is None or k not in names or v.grad.norm() < ZERO:
            continue
        r = _rel(names[k].grad, v.grad)
        worst = max(worst, r)
        assert r < 0.12, (k, r)          # two bf16 engines in series: encoder-side tolerance widened from 8e-2
    for k, got in deng.grads().items():
        if dsg[k].grad.norm() < ZERO:
            continue
        r = _rel(got, dsg[k].grad)
        worst = max(worst, r)
        assert r < 0.12, (k, r)
    record("configs[2] full chain (small encoder + 2-layer decoder + label-smoothed CE), waveform -> loss -> every "
           "parameter gradient: worst relative L2 error vs autograd over the fp32 oracle", worst, 0.12)
    # one optimizer step on both engines lowers the loss of the same batch
    eeng.adam_step(lr=5e-4, clip_norm=10.0)
    deng.adam_step(lr=5e-4, clip_norm=10.0)
    out = eeng.forward_train(wav.cuda(), lens.cuda(), [imgs.cuda()], [None])
    deng.forward_train(prev.cuda(), out["encoder_out"][0], out["encoder_padding_mask"][0])
    loss2, _, _ = deng.loss_backward(target.cuda(), 0.2)
    assert loss2.item() < loss.item()


def test_graphed_model_train_step_matches_eager(cuda):
    """GraphedModelTrainStep (encoder + decoder + criterion + joint-norm clipping + Adam under CUDA-graph replay)
    is bit-identical to the same step issued eagerly (every reduction of the step, the embedding-gradient scatter
    included, has a fixed order)."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.decoder_training import UnitDecoderTrainEngine
    from mm_s2ut_b200.graph import GraphedModelTrainStep
    from oracle import decoder as odec
    from test_gpu_encoder import _build

    B, Lu = 2, 24
    wavs, _ = synth.synth_batch(4, B, 1.5, ragged=False)
    wav, lens = synth.pad_waveforms(wavs)
    imgs = synth.synth_images(4, B)
    g = torch.Generator().manual_seed(3)
    prev = torch.randint(4, 104, (B, Lu), generator=g)
    target = torch.randint(4, 104, (B, Lu), generator=g)
    res = {}
    for mode in ("eager", "graph"):
        enc, args, cfg = _build("small", "selective_attention", True)
        enc.dropout_p = enc.activation_dropout_p = enc.attention_dropout_p = 0.0
        enc.SA_image_dropout = enc.SA_attention_dropout = enc.SA_text_dropout = 0.0
        enc.cuda().train()
        eeng = enc.train_engine()
        deng = UnitDecoderTrainEngine(odec.init_decoder(args.encoder_embed_dim, 512, 2, 104, seed=2),
                                      args.encoder_attention_heads, cuda)
        losses = []
        if mode == "eager":
            for it in range(3):
                out = eeng.forward_train(wav.cuda(), lens.cuda(), [imgs.cuda()], [None], drop_image=(it == 1))
                deng.forward_train(prev.cuda(), out["encoder_out"][0], out["encoder_padding_mask"][0])
                loss, _, d_enc = deng.loss_backward(target.cuda(), 0.2)
                eeng.backward(d_enc)
                deng.grad_norm(grad_scale=0.5)
                eeng.adam_step(lr=1e-3, betas=(0.9, 0.98), clip_norm=1.0, weight_decay=0.01, grad_scale=0.5,
                               extra_norm=deng.norm_coef)
                deng.adam_apply(eeng.norm_coef, lr=1e-3, betas=(0.9, 0.98), weight_decay=0.01)
                losses.append(loss.item())
        else:
            gs = GraphedModelTrainStep(enc, deng, B, wav.shape[1], tuple(imgs.shape[1:]), Lu)
            gs.wav.copy_(wav.cuda())
            gs.img.copy_(imgs.cuda())
            gs.prev_tokens.copy_(prev.cuda())
            gs.target.copy_(target.cuda())
            gs.capture()
            for it in range(3):
                gs.forward_backward(drop_image=(it == 1))
                losses.append(gs.loss[it == 1][0].item())
                gs.optimizer_step(1e-3, weight_decay=0.01, clip_norm=1.0, grad_scale=0.5)
        torch.cuda.synchronize()
        res[mode] = (eeng.flat_p.clone(), deng.flat_p.clone(), losses, eeng.norm_coef[:2].clone())
    assert res["eager"][2] == res["graph"][2]
    assert torch.equal(res["eager"][0], res["graph"][0]) and torch.equal(res["eager"][1], res["graph"][1])
    assert torch.equal(res["eager"][3], res["graph"][3])
    assert res["eager"][2][2] < res["eager"][2][0]            # the loss goes down over the three steps


def test_model_train_step_api(cuda):
    """MM_S2UTTransformerModel.train_step / optimizer_step: the stand-alone model trains on the CUDA kernels (loss goes
    down on a fixed batch) and the decoder's trained parameters are written back for checkpoints."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.config import DEFAULT_YAML, load_mm_config, make_args
    from mm_s2ut_b200.models.mm_s2ut_model import MM_S2UTTransformerModel

    cfg = dict(load_mm_config(DEFAULT_YAML))
    cfg.update(SA_image_dropout=0.0, SA_attention_dropout=0.0)
    torch.manual_seed(0)
    args = make_args("small", multimodal_translation_config_yaml=cfg)
    args.decoder_layers = 2
    model = MM_S2UTTransformerModel(args, target_code_size=100, build_unused_projections=False).cuda().train()
    model.encoder.dropout_p = model.encoder.activation_dropout_p = model.encoder.attention_dropout_p = 0.0
    deng = model.decoder_train_engine()
    assert deng.dropout_p == args.dropout and deng.attention_dropout_p == args.attention_dropout   # taken from the args
    deng.dropout_p = deng.attention_dropout_p = deng.activation_dropout_p = 0.0                    # deterministic loss curve
    B, Lu = 2, 20
    wavs, _ = synth.synth_batch(5, B, 1.5, ragged=True)
    wav, lens = synth.pad_waveforms(wavs)
    imgs = synth.synth_images(5, B).cuda()
    g = torch.Generator().manual_seed(1)
    prev = torch.randint(4, 104, (B, Lu), generator=g).cuda()
    target = torch.randint(4, 104, (B, Lu), generator=g).cuda()
    before = model.decoder.layers[0].fc1.weight.detach().clone()
    losses = []
    for _ in range(4):
        loss, nll = model.train_step(wav.cuda(), lens.cuda(), prev, target, imgs_list=[imgs], img_masks_list=[None])
        model.optimizer_step(lr=1e-3, clip_norm=10.0)
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
    model.sync_decoder_parameters()
    after = model.decoder.layers[0].fc1.weight.detach()
    assert (after - before).abs().max().item() > 0
    assert torch.equal(after, model.decoder_train_engine().p("layers.0.fc1.weight"))
    with pytest.raises(NotImplementedError):
        model(wav.cuda(), lens.cuda(), prev, imgs_list=[imgs], img_masks_list=[None])


def test_decoder_training_step_with_dropout_matches_oracle_with_same_masks(cuda):
    import mm_s2ut_b200.decoder_training as dt
    from mm_s2ut_b200 import kernels as K
    from test_host_training import _decoder_dropout_parity

    worst = _decoder_dropout_parity(K, cuda, dt, B=3, L=40, T=50, d=256, heads=4, ffn=512, layers=2, vocab=104, seed=1)
    record("configs[2] decoder backward with dropout 0.1 / attention-dropout 0.1 / activation-dropout 0.15 (same masks in "
           "the oracle): worst parameter-gradient relative L2 error", worst, REL)


def test_decoder_training_step_with_dropout_in_fused_kernels_d512(cuda):
    """d_model 512: the residual-site masks are applied inside mm_gemm_resid_ln_drop, the activation mask in the fc1
    epilogue, and every LayerNorm backward emits the masked 16-bit gradient of the next branch."""
    import mm_s2ut_b200.decoder_training as dt
    from mm_s2ut_b200 import kernels as K
    from test_host_training import _decoder_dropout_parity

    worst = _decoder_dropout_parity(K, cuda, dt, B=2, L=70, T=90, d=512, heads=8, ffn=1024, layers=2, vocab=104, seed=2)
    record("configs[2] decoder backward, d_model 512, dropout inside the fused GEMM + LayerNorm / fc1 / LayerNorm-backward "
           "kernels (same masks in the oracle): worst parameter-gradient relative L2 error", worst, REL)


def test_decoder_training_step_with_fused_dgrad_layernorm_backward(cuda, monkeypatch):
    """The decoder's three LayerNorm backwards per layer through mm_gemm_ln_bwd (opt-in), d_model 512, dropout on."""
    import mm_s2ut_b200.decoder_training as dt
    from mm_s2ut_b200 import kernels as K
    from test_host_training import _decoder_dropout_parity

    monkeypatch.setattr(dt.UnitDecoderTrainEngine, "fused_ln_bwd", True)
    worst = _decoder_dropout_parity(K, cuda, dt, B=2, L=70, T=90, d=512, heads=8, ffn=1024, layers=2, vocab=104, seed=3)
    record("configs[2] decoder backward, d_model 512, fused dgrad + LayerNorm backward (opt-in), dropout on (same masks in "
           "the oracle): worst parameter-gradient relative L2 error", worst, REL)
