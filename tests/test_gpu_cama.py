"""CAMA transformer forward (K5/K6/K7) against the reference's torch.nn.TransformerEncoder."""
import pytest
import torch
import torch.nn as nn

from oracle import cama_context as cc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,bias,gelu,splits", [(250, 3072, 1024, True, False, 1), (250, 4096, 1024, True, True, 1),
                                                    (250, 1024, 4096, False, False, 8), (250, 1024, 1024, False, False, 4),
                                                    (4000, 1024, 1024, True, False, 1), (37, 128, 64, False, False, 1),
                                                    (129, 256, 192, True, True, 3)])
def test_linear_matches_fp32_matmul(libmrag, M, N, K, bias, gelu, splits):
    from motionrag_b200.cama import linear
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g).bfloat16() if bias else None
    want = a.float() @ w.float().T
    got = linear(a.cuda(), w.cuda(), b.cuda() if bias else None, gelu, splits)
    if splits > 1:
        got = got.sum(0).cpu()
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-4)          # fp32 partial sums: only fp32 rounding
    else:
        if bias:
            want = want + b.float()
        if gelu:
            want = torch.nn.functional.gelu(want)
        assert torch.allclose(got.float().cpu(), want, rtol=1.6e-2, atol=1e-2)   # one bf16 rounding of the output
        assert torch.equal(got.cpu(), want.bfloat16()) or (got.float().cpu() - want).abs().max() < 2e-2


@pytest.mark.parametrize("M,N,K,bias,gelu,splits", [(250, 3072, 1024, True, False, 4), (250, 4096, 1024, True, True, 4),
                                                    (500, 3072, 1024, True, False, 2), (77, 256, 512, False, True, 8),
                                                    (250, 1024, 4096, True, False, 8)])
def test_linear_with_cluster_reduced_split_k(libmrag, M, N, K, bias, gelu, splits):
    """bf16 output of a split-K GEMM: the splits are summed through distributed shared memory in a fixed
    order, so the result equals the unsplit kernel's up to fp32 summation order — and is deterministic."""
    from motionrag_b200.cama import linear
    g = torch.Generator().manual_seed(M + N + K + splits)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    b = torch.randn(N, generator=g).bfloat16().cuda() if bias else None
    want = a.float() @ w.float().T
    if bias:
        want = want + b.float()
    if gelu:
        want = torch.nn.functional.gelu(want)
    got = linear(a, w, b, gelu, splits, cluster_reduce=True)
    again = linear(a, w, b, gelu, splits, cluster_reduce=True)
    assert torch.equal(got, again)
    assert (got.float() - want).abs().max() < 2e-2
    unsplit = linear(a, w, b, gelu, 1)
    assert (got.float() - unsplit.float()).abs().max() <= 2 ** -6 * max(1.0, float(want.abs().max()))   # <= 1 bf16 ulp


@pytest.mark.parametrize("M,N,K,bias,gelu,partial", [(4000, 4096, 1024, True, True, False),      # 128 x 256 tiles
                                                     (4000, 3072, 1024, True, False, False),
                                                     (6000, 2048, 256, False, False, True),
                                                     (12803, 384, 128, True, False, False),       # 128 x 128 tiles
                                                     (12803, 384, 192, False, False, True),
                                                     (2500, 1024, 4096, False, False, True),       # one tile per CTA
                                                     (4000, 1024, 4096, False, False, True),       # one round of 256 x 256 pair tiles
                                                     (3900, 1024, 512, True, True, False),         # same, last pair half empty
                                                     (3700, 1024, 1024, False, False, True)])
def test_persistent_linear_for_many_tiles(libmrag, M, N, K, bias, gelu, partial):
    """More than one wave of 128 x 128 tiles runs the persistent kernel (double-buffered TMEM accumulators)."""
    from motionrag_b200.cama import linear
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    b = torch.randn(N, generator=g).bfloat16().cuda() if bias else None
    want = a.float() @ w.float().T
    got = linear(a, w, b, gelu, 1, partial=partial)
    if partial:
        assert tuple(got.shape) == (1, M, N) and torch.allclose(got[0], want, rtol=1e-4, atol=1e-4)
        return
    if bias:
        want = want + b.float()
    if gelu:
        want = torch.nn.functional.gelu(want)
    assert (got.float() - want).abs().max() < 2e-2
    assert torch.equal(got, linear(a, w, b, gelu, 1))


def _encoder(d, heads, dff, layers, seed):
    torch.manual_seed(seed)
    layer = nn.TransformerEncoderLayer(d, heads, dff, 0.0, "gelu", batch_first=True, norm_first=False, bias=True)
    enc = nn.TransformerEncoder(layer, layers, enable_nested_tensor=False).eval()
    with torch.no_grad():                       # non-trivial norms / biases, bf16-representable weights
        for p in enc.parameters():
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
            p.copy_(p.bfloat16().float())
    return enc


@pytest.mark.parametrize("b,G,L,d,heads,dff,layers", [(1, 10, 25, 1024, 16, 4096, 4), (3, 10, 25, 1024, 16, 4096, 4),
                                                      (2, 4, 5, 256, 4, 512, 2), (2, 3, 40, 512, 8, 1024, 2),
                                                      (10, 10, 25, 1024, 16, 4096, 2),   # 2 500 rows: persistent GEMMs, warp-per-row LayerNorm
                                                      (9, 10, 25, 768, 12, 1536, 2),
                                                      # >= 176 (sample, head) pairs: one attention CTA per pair (K6w)
                                                      (16, 10, 25, 1024, 16, 4096, 1),   # 4 000 rows: pair GEMMs for N = 1024 too
                                                      (44, 3, 7, 256, 4, 512, 1),        # 21 tokens: second MMA tile mostly padding
                                                      (22, 7, 13, 512, 8, 512, 2),       # 91 tokens, groups cut MMA tiles anywhere
                                                      (22, 11, 64, 512, 8, 512, 1)])     # 704 tokens, the longest supported
def test_forward_matches_reference_encoder(libmrag, b, G, L, d, heads, dff, layers):
    """Reference configuration (configs/cogvideox/MotionRAG_open.yml:253-267) and a small one."""
    from motionrag_b200 import CamaTransformer
    enc = _encoder(d, heads, dff, layers, seed=b + G)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(b, G * L, d, generator=g).bfloat16()
    mask = cc.block_causal_mask(G, L)
    with torch.no_grad():
        want = enc(x.float(), mask)                                   # fp32 evaluation = the oracle
        torch_bf16 = enc.cuda().bfloat16()(x.cuda(), mask.cuda()).float().cpu()
    cama = CamaTransformer(enc, groups=G, group_tokens=L, max_batch=max(4, b), device=0)
    got = cama.forward(x.cuda()).float().cpu()
    got_nograph = cama.forward(x.cuda(), use_graph=False).float().cpu()
    assert torch.equal(got, got_nograph)                               # graph replay == direct launches
    err = (got - want).abs()
    err_torch = (torch_bf16 - want).abs()
    # outputs are LayerNorm-ed (unit scale): bf16 carries ~3 significant digits
    assert float(err.max()) < 6e-2 and float(err.mean()) < 6e-3, (float(err.max()), float(err.mean()))
    assert float(err.mean()) <= 1.25 * float(err_torch.mean()) + 1e-4, (float(err.mean()), float(err_torch.mean()))
    # predict(): the last layer computes the last group's rows only; same values up to the fp32 summation order
    # of the split-K GEMMs (the K split depends on the row count), same parity bar against the fp32 oracle
    pred = cama.predict(b=b).float().cpu()
    assert pred.shape == (b, L, d)
    assert float((pred - got[:, -L:]).abs().max()) < 4e-2
    perr = (pred - want[:, -L:]).abs()
    assert float(perr.max()) < 6e-2 and float(perr.mean()) < 6e-3, (float(perr.max()), float(perr.mean()))
    assert torch.equal(pred, cama.predict(x.cuda(), use_graph=False).float().cpu())
    # in-place input: the gather can write straight into the handle's buffer
    cama.input_view(b).copy_(x.cuda())
    assert torch.equal(cama.forward(b=b).float().cpu(), got)
    cama.close()


def test_block_causality(libmrag):
    """Changing a later group must not change the outputs of earlier groups (get_mask semantics)."""
    from motionrag_b200 import CamaTransformer
    G, L, d = 10, 25, 1024
    enc = _encoder(d, 16, 4096, 2, seed=3)
    cama = CamaTransformer(enc, groups=G, group_tokens=L, max_batch=2, device=0)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, G * L, d, generator=g).bfloat16().cuda()
    y0 = cama.forward(x).clone()
    x2 = x.clone()
    x2[:, 6 * L:] += 1.0
    y1 = cama.forward(x2)
    assert torch.equal(y0[:, :6 * L], y1[:, :6 * L]) and not torch.equal(y0[:, 6 * L:], y1[:, 6 * L:])
    cama.close()


def test_unsupported_shapes_are_loud(libmrag):
    from motionrag_b200 import CamaTransformer, MragError
    enc = _encoder(192, 2, 512, 1, seed=0)                            # head_dim 96
    with pytest.raises(MragError, match="head_dim 64"):
        CamaTransformer(enc, groups=2, group_tokens=4, max_batch=1)
    with pytest.raises(MragError, match="704 tokens"):
        CamaTransformer(_encoder(256, 4, 512, 1, seed=0), groups=40, group_tokens=25, max_batch=1)
    layer = nn.TransformerEncoderLayer(256, 4, 512, 0.0, "relu", batch_first=True)
    with pytest.raises(ValueError, match="gelu"):
        CamaTransformer(nn.TransformerEncoder(layer, 1), groups=2, group_tokens=4)


def test_attach_with_libmrag_transformer_and_cfg_predict(libmrag):
    """Row f-1 end to end: retrieval row ids -> K4 gather into the transformer's own input buffer ->
    K5/K6/K7 forward -> ActionTransformer.predict's slice and CFG concat (module.py:325-331)."""
    from motionrag_b200 import CamaTransformer, FeatureTable, MotionContext, attach
    L, C, K, b, n = 25, 1024, 9, 3, 64
    g = torch.Generator().manual_seed(5)
    table = torch.randn(n, L, C, generator=g).bfloat16()
    sos = (torch.randn(1, L, C, generator=g) / 32).bfloat16()
    un = torch.randn(L, C, generator=g).bfloat16()
    cond = torch.randn(b, (K + 1) * L, C, generator=g).bfloat16()
    idx = torch.randint(0, n, (b, K), generator=g)
    idx[2, 0] = -1

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.transformer = _encoder(C, 16, 4096, 4, seed=11)

        def encode_condition(self, images):
            return images

        def batch_forward(self, batch, return_loss=True, ignore_ref_loss=False):
            raise AssertionError("video path")

        def predict(self, batch, do_classifier_free_guidance=False):
            raise AssertionError("video path")

    model = Model().eval()
    ctx = MotionContext(FeatureTable(table.cuda()), sos, un, pe_max_length=256)
    cama = CamaTransformer(model.transformer, groups=K + 1, group_tokens=L, max_batch=4, device=0)
    x = cc.context_restatement(cc.gather_restatement(table, idx, un), sos, cc.sinusoid_table(256, C), cond.clone())
    with torch.no_grad():
        want = model.transformer(x.float(), cc.block_causal_mask(K + 1, L))[:, -L:]     # fp32 oracle
    attach(model, ctx, transformer=cama)
    batch = {"ref_index": idx.cuda(), "ref_images": cond.cuda()}
    got = model.predict(batch)
    assert torch.equal(cama.input_view(b).cpu(), x)            # the gather wrote the transformer's input in place
    assert got.shape == (b, L, C)
    err = (got.float().cpu() - want).abs()
    assert float(err.max()) < 6e-2 and float(err.mean()) < 6e-3, (float(err.max()), float(err.mean()))
    both = model.predict(batch, do_classifier_free_guidance=True)
    assert both.shape == (2 * b, L, C)
    assert torch.equal(both[b:], got) and torch.equal(both[:b].cpu(), un[None].expand(b, -1, -1))
    with pytest.raises(ValueError, match="groups"):
        model.batch_forward({"ref_index": idx[:, :4].cuda(), "ref_images": cond[:, :5 * L].cuda()}, return_loss=False)
    cama.close()


@pytest.mark.parametrize("seed", range(6))
def test_forward_random_shapes_match_reference_encoder(libmrag, seed):
    """Seeded random transformer shapes inside the kernels' envelope (head_dim 64, d in {256..1024},
    d_ff % 512 == 0, <= 704 tokens): same parity bar as the reference configuration."""
    from motionrag_b200 import CamaTransformer
    g = torch.Generator().manual_seed(900 + seed)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    d = [256, 512, 768, 1024][r(0, 3)]
    dff = 512 * r(1, 4)
    layers, G, L, b = r(1, 3), r(1, 12), r(1, 40), r(1, 5)
    while G * L > 704:
        L -= 1
    enc = _encoder(d, d // 64, dff, layers, seed=seed)
    x = torch.randn(b, G * L, d, generator=g).bfloat16()
    mask = cc.block_causal_mask(G, L)
    with torch.no_grad():
        want = enc(x.float(), mask)
    cama = CamaTransformer(enc, groups=G, group_tokens=L, max_batch=b, device=0)
    got = cama.forward(x.cuda()).float().cpu()
    err = (got - want).abs()
    cfg = dict(d=d, dff=dff, layers=layers, G=G, L=L, b=b)
    assert float(err.max()) < 8e-2 and float(err.mean()) < 8e-3, (cfg, float(err.max()), float(err.mean()))
    cama.close()
