"""Device-resident image-feature store: gathered 16-bit LayerNorm kernel and the encoder path."""
import pytest
import torch

from _util import record
from test_gpu_encoder import TOL, _build, _compare, _oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,in_dt", [(768, torch.float16), (768, torch.bfloat16), (256, torch.float16)])
def test_layernorm_gather_kernel(cuda, dim, in_dt):
    from mm_s2ut_b200 import kernels as K

    g = torch.Generator().manual_seed(dim)
    N, Tk, B = 11, 37, 5
    store = (torch.randn(N, Tk, dim, generator=g) * 1.5 + 0.2).to(in_dt).to(cuda)
    idx = torch.tensor([7, 0, 10, 7, 3], dtype=torch.int64, device=cuda)
    gamma, beta = (1 + 0.2 * torch.randn(dim, generator=g)).to(cuda), (0.2 * torch.randn(dim, generator=g)).to(cuda)
    out = torch.zeros(B * Tk, dim, dtype=torch.bfloat16, device=cuda)
    K.layernorm_gather(store, idx, Tk, gamma, beta, out)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(store[idx].float().view(B * Tk, dim), (dim,), gamma, beta, 1e-5)
    assert (out.float() - ref).abs().max().item() < 4e-2
    out2 = torch.zeros(N * Tk, dim, dtype=torch.bfloat16, device=cuda)
    K.layernorm_gather(store, None, Tk, gamma, beta, out2)      # no index: rows in order
    torch.cuda.synchronize()
    ref2 = torch.nn.functional.layer_norm(store.float().view(N * Tk, dim), (dim,), gamma, beta, 1e-5)
    assert (out2.float() - ref2).abs().max().item() < 4e-2


def test_encoder_with_feature_store(cuda):
    """A batch drawn from an fp16 store gives the same fused states as the fp32 tensors it was built from (to fp16
    input rounding) and stays within the north-star tolerance of the fp32 oracle; also through the CUDA graph."""
    from mm_s2ut_b200 import synth
    from mm_s2ut_b200.feature_store import ImageFeatureStore
    from mm_s2ut_b200.graph import GraphedEncoder

    enc, args, cfg = _build("small")
    wavs, _ = synth.synth_batch(0, 4, 5.0, ragged=True)
    pool = synth.synth_images(3, 9)                       # "dataset" of 9 images; the batch uses 4 of them
    pick = [5, 0, 8, 5]
    imgs = pool[pick]
    ref = _oracle(enc, args, cfg, wavs, imgs)
    wav, lens = synth.pad_waveforms(wavs)
    enc.cuda()
    direct = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[imgs.cuda()], img_masks_list=[None])
    store = ImageFeatureStore(pool, cuda)
    out = enc(wav.cuda(), lens.cuda(), None, None, None, imgs_list=[store.batch(pick)], img_masks_list=[None])
    torch.cuda.synchronize()
    err = _compare(out, ref)
    record("configs[0] small, image features from the fp16 device store: fused states max-abs err", err, TOL)
    assert err < TOL
    d = (out["encoder_out"][0] - direct["encoder_out"][0]).abs().max().item()
    assert d < 1e-2, d
    # graph replay with an index vector as the only image input
    n = wav.shape[1]
    genc = GraphedEncoder(enc, 4, n, [tuple(imgs.shape[1:])], stores=[store])
    gout = genc(wav.cuda(), lens.cuda(), [torch.tensor(pick, dtype=torch.int64)])
    torch.cuda.synchronize()
    assert (gout["encoder_out"][0] - out["encoder_out"][0]).abs().max().item() < 1e-5
