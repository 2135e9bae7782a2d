"""Gather kernel (K4) against the reference-captured golden context and the restatement."""
import numpy as np
import pytest
import torch

from oracle import cama_context as cc

pytestmark = pytest.mark.gpu


def _bf16(a):
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)


@pytest.mark.parametrize("name,dt", [("bf16", torch.bfloat16), ("f32", torch.float32)])
def test_gather_reproduces_reference_context_bit_for_bit(libmrag, golden_dir, name, dt):
    """Fixture x was captured from the reference's real ActionTransformer.batch_forward."""
    from motionrag_b200 import FeatureTable, gather_context
    z = np.load(golden_dir / f"cama_context_{name}.npz")
    t = (lambda k: _bf16(z[k])) if dt == torch.bfloat16 else (lambda k: torch.from_numpy(z[k]))
    ref = t("ref_feats")
    b, K, L, C = ref.shape
    # put the retrieved rows at scattered positions of a bigger table
    g = torch.Generator().manual_seed(0)
    rows = torch.randperm(500, generator=g)[:b * K].view(b, K)
    table = torch.randn(500, L, C, generator=g).to(dt)
    table[rows.flatten()] = ref.reshape(b * K, L, C)
    ft = FeatureTable(table.cuda())
    pos = torch.from_numpy(z["pos_table"])[0].to(dt).cuda()
    x = gather_context(ft, rows.cuda(), t("sos").cuda(), torch.zeros(L, C, dtype=dt).cuda(), pos, t("cond").cuda())
    assert torch.equal(x.cpu(), t("x"))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("b,K,with_pe,with_cond", [(1, 9, True, True), (16, 9, True, False), (5, 3, False, True),
                                                   (2, 1, False, False)])
def test_gather_matches_restatement_cogvideox_shapes(libmrag, dt, b, K, with_pe, with_cond):
    """L=25, C=1024 (configs/cogvideox/MotionRAG_open.yml:228-238), missing refs -> uncond row."""
    from motionrag_b200 import FeatureTable, MotionContext
    L, C, n = 25, 1024, 300
    g = torch.Generator().manual_seed(b * 100 + K)
    table = torch.randn(n, L, C, generator=g).to(dt)
    idx = torch.randint(0, n, (b, K), generator=g)
    idx[0, K // 2] = -1
    idx[-1, 0] = -1
    sos = (torch.randn(1, L, C, generator=g) / 32).to(dt)
    un = torch.randn(L, C, generator=g).to(dt)
    cond = torch.randn(b, (K + 1) * L, C, generator=g).to(dt) if with_cond else None
    ctx = MotionContext(FeatureTable(table.cuda()), sos, un, pe_max_length=256 if with_pe else None)
    x = ctx.build(idx.cuda(), cond.cuda() if with_cond else None)
    feats = cc.gather_restatement(table, idx, un)
    want = cc.context_restatement(feats, sos, cc.sinusoid_table(256, C) if with_pe else None,
                                  cond.clone() if with_cond else None)
    assert torch.equal(x.cpu(), want)
    assert torch.equal(ctx.get_mask(K + 1, L).cpu(), cc.block_causal_mask(K + 1, L))
    assert torch.equal(ctx.uncond_action_emb(b).cpu(), un[None].expand(b, -1, -1))


def test_gather_reads_row_sharded_tables_through_pointer_table(libmrag):
    """Two shard blocks (same GPU here; peer-mapped over NVLink in the multi-process case)."""
    from motionrag_b200 import FeatureTable, gather_context
    L, C, rps = 25, 1024, 64
    g = torch.Generator().manual_seed(3)
    full = torch.randn(2 * rps, L, C, generator=g).to(torch.bfloat16)
    s0, s1 = full[:rps].contiguous().cuda(), full[rps:].contiguous().cuda()
    ft = FeatureTable(s0, rows_per_shard=rps, shard_rank=0, n_shards=2)
    assert not ft.complete
    with pytest.raises(RuntimeError):
        gather_context(ft, torch.zeros(1, 2, dtype=torch.int64).cuda(), full[0].cuda(), full[0].cuda())
    ft.set_peer_ptr(1, s1.data_ptr())
    idx = torch.tensor([[3, 100, 64, 63], [127, 0, -1, 65]])
    un = torch.zeros(L, C, dtype=torch.bfloat16)
    x = gather_context(ft, idx.cuda(), full[5].cuda(), un.cuda())
    want = cc.context_restatement(cc.gather_restatement(full, idx, un), full[5][None], None, None)
    assert torch.equal(x.cpu(), want)


def test_gather_never_dereferences_rows_past_the_table(libmrag):
    """A short last shard: indices inside its nominal range but beyond the rows the table really has (or
    stale ones) select the uncond row; validate=True raises instead."""
    from motionrag_b200 import FeatureTable, gather_context
    L, C, rps = 25, 1024, 64
    g = torch.Generator().manual_seed(4)
    full = torch.randn(rps + 10, L, C, generator=g).to(torch.bfloat16)           # second shard holds 10 rows only
    s0, s1 = full[:rps].contiguous().cuda(), full[rps:].contiguous().cuda()
    ft = FeatureTable(s0, rows_per_shard=rps, shard_rank=0, n_shards=2, n_rows=rps + 10)
    ft.set_peer_ptr(1, s1.data_ptr())
    un = torch.randn(L, C, generator=g).to(torch.bfloat16)
    idx = torch.tensor([[3, rps + 9, rps + 10, 127], [10_000_000, 0, -1, rps]])
    x = gather_context(ft, idx.cuda(), full[5].cuda(), un.cuda())
    safe = torch.where(idx >= rps + 10, torch.full_like(idx, -1), idx)
    want = cc.context_restatement(cc.gather_restatement(full, safe, un), full[5][None], None, None)
    assert torch.equal(x.cpu(), want)
    with pytest.raises(IndexError):
        gather_context(ft, idx.cuda(), full[5].cuda(), un.cuda(), validate=True)
    with pytest.raises(ValueError):
        FeatureTable(s0, rows_per_shard=rps, shard_rank=0, n_shards=2, n_rows=3 * rps)


def test_gather_argument_checks(libmrag):
    from motionrag_b200 import FeatureTable, gather_context
    ft = FeatureTable(torch.zeros(4, 25, 1024, dtype=torch.bfloat16).cuda())
    z = torch.zeros(25, 1024, dtype=torch.bfloat16).cuda()
    with pytest.raises(ValueError):
        gather_context(ft, torch.zeros(1, 2, dtype=torch.int32).cuda(), z, z)
    with pytest.raises(ValueError):
        gather_context(ft, torch.zeros(1, 2, dtype=torch.int64).cuda(), z.float(), z)


def test_attach_drives_a_cama_style_transformer_from_row_ids(libmrag):
    """Face 2 end to end: a model with the reference's ActionTransformer surface
    (encode_condition / transformer / batch_forward, src/projects/condition/module.py:270-323)
    takes batch['ref_index'] after attach() and returns transformer(x, mask) with x built by K4."""
    import torch.nn as nn
    from motionrag_b200 import FeatureTable, MotionContext, attach
    L, C, K, b, n = 25, 1024, 9, 2, 64
    g = torch.Generator().manual_seed(0)
    table = torch.randn(n, L, C, generator=g).bfloat16()
    sos = (torch.randn(1, L, C, generator=g) / 32).bfloat16()
    un = torch.randn(L, C, generator=g).bfloat16()
    cond = torch.randn(b, (K + 1) * L, C, generator=g).bfloat16()
    idx = torch.randint(0, n, (b, K), generator=g)
    idx[1, 4] = -1

    class Cama(nn.Module):          # the three members attach() touches
        def __init__(self):
            super().__init__()
            layer = nn.TransformerEncoderLayer(C, 16, 4096, 0.0, "gelu", batch_first=True, norm_first=False)
            self.transformer = nn.TransformerEncoder(layer, 1, enable_nested_tensor=False)

        def encode_condition(self, images):
            return images                                  # features supplied directly

        def batch_forward(self, batch, return_loss=True, ignore_ref_loss=False):
            raise AssertionError("the video path must not run when ref_index is given")

    model = Cama().cuda().bfloat16().eval()
    ctx = MotionContext(FeatureTable(table.cuda()), sos, un, pe_max_length=256)
    attach(model, ctx)
    with torch.no_grad():
        pred = model.batch_forward({"ref_index": idx.cuda(), "ref_images": cond.cuda()}, return_loss=False)
        x = cc.context_restatement(cc.gather_restatement(table, idx, un), sos, cc.sinusoid_table(256, C), cond.clone())
        want = model.transformer(x.cuda(), cc.block_causal_mask(K + 1, L).cuda())
    assert pred.shape == (b, K + 1, L, C)
    assert torch.equal(pred.reshape(b, -1, C), want)       # same x, same mask -> same bits
    # batch['ref_features'] (SURVEY §8b face 2): the same x from materialised features
    feats = cc.gather_restatement(table, idx, un)                      # [b, K, L, C], -1 -> uncond row
    with torch.no_grad():
        pred2 = model.batch_forward({"ref_features": feats.cuda(), "ref_images": cond.cuda()}, return_loss=False)
    assert torch.equal(pred2, pred)


@pytest.mark.parametrize("seed", range(10))
def test_gather_random_shapes_match_restatement(libmrag, seed):
    """Seeded random (b, K, L, C, dtype, pe, cond, share of missing references): bit-exact."""
    from motionrag_b200 import FeatureTable, MotionContext
    g = torch.Generator().manual_seed(500 + seed)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    b, K, L = r(1, 9), r(1, 12), r(1, 30)
    C = [64, 128, 384, 1024][r(0, 3)]
    dt = [torch.bfloat16, torch.float32][r(0, 1)]
    with_pe, with_cond = bool(r(0, 1)), bool(r(0, 1))
    n = r(1, 200)
    table = torch.randn(n, L, C, generator=g).to(dt)
    idx = torch.randint(0, n, (b, K), generator=g)
    idx[torch.rand(b, K, generator=g) < 0.25] = -1
    sos = (torch.randn(1, L, C, generator=g) / 8).to(dt)
    un = torch.randn(L, C, generator=g).to(dt)
    cond = torch.randn(b, (K + 1) * L, C, generator=g).to(dt) if with_cond else None
    ctx = MotionContext(FeatureTable(table.cuda()), sos, un, pe_max_length=(K + 1) * L + 3 if with_pe else None)
    x = ctx.build(idx.cuda(), cond.cuda() if with_cond else None)
    want = cc.context_restatement(cc.gather_restatement(table, idx, un), sos,
                                  cc.sinusoid_table((K + 1) * L + 3, C) if with_pe else None,
                                  cond.clone() if with_cond else None)
    assert torch.equal(x.cpu(), want), dict(b=b, K=K, L=L, C=C, dt=dt, pe=with_pe, cond=with_cond, n=n)


def test_loss_path_matches_the_reference_training_and_validation_forward(libmrag, golden_dir):
    """attach().batch_forward(return_loss=True) — what training_step / validation_step / test_step call
    (src/projects/condition/module.py:333-351) — with the K references read from the feature table and only the
    target clip's features supplied: losses and the sos_token gradient equal the recording of the reference's REAL
    ActionTransformer (tests/golden/cama_loss_f32.npz, oracle/make_golden.py::make_cama_loss)."""
    import types

    import torch.nn as nn
    import torch.nn.functional as F
    from motionrag_b200 import FeatureTable, MotionContext, attach
    gold = np.load(golden_dir / "cama_loss_f32.npz")
    C, heads, ff, layers = (int(v) for v in gold["shape"])
    ref, tgt = torch.from_numpy(gold["ref_feats"]), torch.from_numpy(gold["target"])
    cond, sos = torch.from_numpy(gold["cond"]), torch.from_numpy(gold["sos"])
    b, K, L, _ = ref.shape

    class Cama(nn.Module):          # the members of ActionTransformer the loss path touches
        def __init__(self):
            super().__init__()
            layer = nn.TransformerEncoderLayer(C, heads, ff, 0.0, "gelu", batch_first=True, norm_first=False)
            self.transformer = nn.TransformerEncoder(layer, layers, enable_nested_tensor=False)
            self.transformer.load_state_dict({k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")})
            self.sos_token = nn.Parameter(sos.clone())
            self.register_buffer("pos_table", cc.sinusoid_table(256, C))

        def vision_pe(self, x):                      # SinusoidPositionalEmbeddings.forward (position_embeddings.py:172-174)
            return x + self.pos_table[:, :x.size(-2)].type_as(x)

        def encode_condition(self, images):
            return images

        def encode_vision(self, videos):             # must only ever see the TARGET clip: [b, 1, ...]
            assert videos.shape[1] == 1
            return videos[:, :, 0]

        def get_loss(self, pred, emb):               # module.py:277-290
            pred, emb = pred.flatten(0, 1), emb.flatten(0, 1)
            mse = F.mse_loss(pred, emb)
            return types.SimpleNamespace(main=mse, mse=mse, smooth=F.smooth_l1_loss(pred, emb))

        def batch_forward(self, batch, return_loss=True, ignore_ref_loss=False):
            raise AssertionError("the video path must not run when ref_index is given")

    # the table holds the reference features at scattered rows; ref_index is in similarity order (0 = most similar)
    rows = torch.randperm(64, generator=torch.Generator().manual_seed(1))[:b * K].view(b, K)
    table = torch.zeros(64, L, C)
    table[rows.flatten()] = ref.reshape(b * K, L, C)
    un = torch.randn(L, C, generator=torch.Generator().manual_seed(2))
    for tag, ignore in (("train", False), ("val", True)):
        model = Cama().cuda()
        ctx = MotionContext(FeatureTable(table.cuda()), sos, un, pe_max_length=256)
        attach(model, ctx)
        for batch in ({"ref_index": rows.cuda(), "ref_images": cond.cuda(), "target_features": tgt.cuda()},
                      {"ref_index": rows.cuda(), "ref_images": cond.cuda(), "video": tgt.cuda()[:, None]},
                      {"ref_features": ref.cuda(), "ref_images": cond.cuda(), "target_features": tgt.cuda()}):
            model.zero_grad()
            loss = model.batch_forward(batch, return_loss=True, ignore_ref_loss=ignore)
            loss.main.backward()
            assert float(loss.mse) == pytest.approx(float(gold[f"{tag}_mse"]), rel=2e-5)
            assert float(loss.smooth) == pytest.approx(float(gold[f"{tag}_smooth"]), rel=2e-5)
            torch.testing.assert_close(model.sos_token.grad.cpu(), torch.from_numpy(gold[f"{tag}_sos_grad"]),
                                       rtol=1e-3, atol=1e-7)
