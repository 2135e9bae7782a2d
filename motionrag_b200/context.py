"""Context assembly for the CAMA causal motion transformer — face 2 of the drop-in boundary.

`gather_context` produces, with ONE kernel launch (K4, mrag_gather_context), the tensor `x` that
the reference's ActionTransformer.forward builds at src/projects/condition/module.py:298-301
from the K retrieved clips, without decoding or re-encoding them: rows of a precomputed
feature table are packed straight into the `[b, (K+1)*L, C]` layout, optionally with the
sinusoid position table and the condition embedding added in the reference's order/rounding.

`MotionContext` mirrors the pieces of ActionTransformer a caller touches (get_mask, the flipped
similarity order of batch_forward :318-319, the CFG "uncond" row of predict :326-330) and
`attach` lets a real ActionTransformer accept `batch['ref_index']` in place of
`batch['ref_videos']`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import check
from .store import FeatureTable, _stream_ptr


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """Same table as SinusoidPositionalEmbeddings (position_embeddings.py:159-170): float64
    angles, sin on even / cos on odd columns, stored float32. Shape [n_position, d_hid]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    t = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    t[:, 0::2] = np.sin(t[:, 0::2])
    t[:, 1::2] = np.cos(t[:, 1::2])
    return torch.from_numpy(t.astype(np.float32))


def block_causal_mask(num_groups: int, group_tokens: int, device=None) -> torch.Tensor:
    """ActionTransformer.get_mask (module.py:131-135): bool, True = blocked."""
    g = torch.arange(num_groups * group_tokens, device=device) // group_tokens
    return g[None, :] > g[:, None]


def _ptr(t: torch.Tensor | None):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def gather_context(table: FeatureTable, ref_index: torch.Tensor, sos: torch.Tensor,
                   uncond_row: torch.Tensor, pos_table: torch.Tensor | None = None,
                   cond: torch.Tensor | None = None, out: torch.Tensor | None = None,
                   validate: bool = False) -> torch.Tensor:
    """ref_index [b, K] int64 (similarity order, 0 = most similar, -1 = dropped/missing) ->
    x [b, (K+1)*L, C] in the table's dtype. sos [L, C] / [1, L, C]; uncond_row [L, C];
    pos_table [>= (K+1)*L, C] already in the table's dtype (the reference casts its fp32 table
    with .type_as(x) before adding); cond [b, (K+1)*L, C]. An index >= table.n_rows is never
    dereferenced (the kernel substitutes the uncond row); validate=True raises for one instead
    (costs a device synchronisation)."""
    lib = _cabi.load()
    dev, dt = table.device, table.local.dtype
    if ref_index.device != dev or ref_index.dtype != torch.int64 or ref_index.ndim != 2:
        raise ValueError("ref_index must be an int64 [b, K] tensor on the table's device")
    ref_index = ref_index.contiguous()
    if validate and ref_index.numel() and int(ref_index.max()) >= table.n_rows:
        raise IndexError(f"ref_index holds row {int(ref_index.max())} but the feature table has {table.n_rows} rows")
    b, K = ref_index.shape
    L, Cd = table.L, table.Cdim
    n_tok = (K + 1) * L

    def chk(name, t, shape):
        if t is None:
            return None
        if t.device != dev or t.dtype != dt:
            raise ValueError(f"{name} must be {dt} on {dev}")
        t = t.reshape(shape) if t.numel() == int(np.prod(shape)) else t
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
        return t.contiguous()

    sos = chk("sos", sos, (L, Cd))
    uncond_row = chk("uncond_row", uncond_row, (L, Cd))
    cond = chk("cond", cond, (b, n_tok, Cd))
    if pos_table is not None:
        if pos_table.device != dev or pos_table.dtype != dt or pos_table.ndim != 2 or pos_table.shape[0] < n_tok:
            raise ValueError(f"pos_table must be {dt} [>={n_tok}, {Cd}] on {dev}")
        pos_table = pos_table[:n_tok].contiguous()
    if out is None:
        out = torch.empty((b, n_tok, Cd), dtype=dt, device=dev)
    elif tuple(out.shape) != (b, n_tok, Cd) or out.dtype != dt or not out.is_contiguous():
        raise ValueError("bad out tensor")
    if not table.complete:
        raise RuntimeError("feature table has unmapped peer shards; call parallel.open_peer_tables first")
    check(lib.mrag_gather_context(
        _ptr(table.shard_ptrs), table.n_shards, table.rows_per_shard, _ptr(ref_index), _ptr(sos),
        _ptr(uncond_row), _ptr(pos_table), _ptr(cond), _ptr(out), b, K, L, Cd,
        0 if dt == torch.bfloat16 else 1, table.n_rows, _stream_ptr(dev)))
    return out


def select_refs(index: torch.Tensor, distance: torch.Tensor, ref_video_num: int,
                uncond_video_ratio: float = 0.0, generator: torch.Generator | None = None):
    """The per-sample reference selection of VideoDataset.get_ref_videos
    (src/data/dataset.py:285-312) on retrieval results instead of decoded clips:
    keep the first `ref_video_num` hits (`ref_videos[:K]`, :296), drop each one with probability
    `uncond_video_ratio` (`random.random() > ratio` keeps, :297) — a dropped or missing slot
    becomes index -1 (the all-zero clip -> uncond row) with `_distance` 1.0 (:305-310).

    index / distance: [b, k>=K] from a search (unused slots -1 / inf). Returns
    (ref_index int64 [b, K], ref_distance float32 [b, K]). One deliberate difference: when a search returned
    fewer than K rows the reference's distance LIST is simply shorter (:296 iterates over what exists) while its
    clip tensor still has K slots (zeros); the consumer of the distances, `condition_fusion(..., 'weight')`
    (src/projects/condition/utils.py:25-28: w ~ 1 - distance), needs one value per slot, so the tensor form
    here fills the missing slots with 1.0 — weight 0, exactly what a dropped clip gets."""
    K = int(ref_video_num)
    idx = index[:, :K].clone()
    dist = distance[:, :K].clone().to(torch.float32)
    if idx.shape[1] < K:   # fewer results than slots: the reference leaves zeros there
        pad = K - idx.shape[1]
        idx = torch.cat([idx, idx.new_full((idx.shape[0], pad), -1)], 1)
        dist = torch.cat([dist, dist.new_full((dist.shape[0], pad), 1.0)], 1)
    drop = idx < 0
    if uncond_video_ratio > 0:
        gdev = generator.device if generator is not None else idx.device
        u = torch.rand(idx.shape, generator=generator, device=gdev).to(idx.device)
        drop = drop | ~(u > uncond_video_ratio)
    idx = torch.where(drop, torch.full_like(idx, -1), idx)
    dist = torch.where(drop, torch.ones_like(dist), dist)
    return idx.contiguous(), dist.contiguous()


class MotionContext:
    """Holds what the gather needs besides the table: SOS block, uncond row, position table."""

    def __init__(self, table: FeatureTable, sos_token: torch.Tensor, uncond_row: torch.Tensor,
                 pe_max_length: int | None = 256):
        dt, dev = table.local.dtype, table.device
        self.table = table
        self.sos = sos_token.detach().to(dev, dt).reshape(table.L, table.Cdim).contiguous()
        self.uncond_row = uncond_row.detach().to(dev, dt).reshape(table.L, table.Cdim).contiguous()
        # the reference keeps an fp32 table and casts per call (.type_as(x)); cast once here
        self.pos_table = (sinusoid_table(pe_max_length, table.Cdim).to(dev).to(dt)
                          if pe_max_length else None)

    def get_mask(self, num_frames: int, frame_tokens: int) -> torch.Tensor:
        return block_causal_mask(num_frames, frame_tokens, self.table.device)

    def build(self, ref_index: torch.Tensor, condition_emb: torch.Tensor | None = None,
              out: torch.Tensor | None = None, with_pe: bool = True) -> torch.Tensor:
        """x of module.py:298-301 from row ids. with_pe=False leaves the position table out (the loss path adds
        it, and the condition embedding, with differentiable torch ops)."""
        return gather_context(self.table, ref_index, self.sos, self.uncond_row, self.pos_table if with_pe else None,
                              condition_emb, out)

    def build_from_features(self, ref_features: torch.Tensor, condition_emb: torch.Tensor | None = None,
                            out: torch.Tensor | None = None, with_pe: bool = True) -> torch.Tensor:
        """Same `x` from already-materialised reference features (`batch['ref_features']`,
        [b, K, L, C] in similarity order, 0 = most similar) instead of table rows: the batch itself
        is the table and slot (i, k) reads row i*K + k — still one K4 launch."""
        if ref_features.ndim != 4 or tuple(ref_features.shape[2:]) != (self.table.L, self.table.Cdim):
            raise ValueError(f"ref_features must be [b, K, {self.table.L}, {self.table.Cdim}]")
        b, K = ref_features.shape[:2]
        dt, dev = self.table.local.dtype, self.table.device
        feats = ref_features.detach().to(dev, dt).reshape(b * K, self.table.L, self.table.Cdim).contiguous()
        idx = torch.arange(b * K, dtype=torch.int64, device=dev).view(b, K)
        return gather_context(FeatureTable(feats), idx, self.sos, self.uncond_row, self.pos_table if with_pe else None,
                              condition_emb, out)

    def uncond_action_emb(self, b: int) -> torch.Tensor:
        """predict()'s CFG branch (module.py:327-329): encode_vision(zeros)[:, 0] per sample."""
        return self.uncond_row[None].expand(b, -1, -1)


def attach(model, ctx: MotionContext, transformer=None):
    """Teach a reference ActionTransformer to take `batch['ref_index']` ([b, K] int64 row ids
    from retrieval) or `batch['ref_features']` ([b, K, L, C] features in similarity order) instead
    of `batch['ref_videos']`, producing the same prediction tensor (`return_loss=False`) or the same
    loss (`return_loss=True`, see loss_forward below) as module.py:292-315 with encode_vision of the K
    references replaced by the table gather. The condition embedding is still computed by the
    model's own encode_condition from `batch['ref_images']` ([b, K+1, C, H, W], refs flipped +
    target first frame, as batch_forward builds them at :321).

    transformer: optionally a `CamaTransformer` built from `model.transformer`; the context is then
    gathered straight into its input buffer and the 4-layer forward runs on libmrag kernels
    (bf16 tables only). `predict` (module.py:325-331) is patched the same way: the CFG "uncond" rows
    come from the stored uncond row instead of encoding an all-zero clip."""
    orig = model.batch_forward
    orig_predict = getattr(model, 'predict', None)
    if transformer is not None and ctx.table.local.dtype != torch.bfloat16:
        raise ValueError("the libmrag transformer consumes a bf16 feature table")

    def loss_forward(batch, ignore_ref_loss: bool):
        """training_step / validation_step / test_step (module.py:333-351 -> forward :292-311 with
        return_loss=True): the K reference features come from the table (K4 gather, no autograd — vision_model
        and vision_proj are frozen in the CAMA configs), only the TARGET clip is encoded
        (`batch['target_features']` [b, L, C] when the caller has them, else one `encode_vision` pass over
        `batch['video']` instead of K+1). What can carry gradients stays in torch and differentiable: the SOS
        block is the model's own `sos_token` parameter, the position table and the condition embedding are added
        with torch ops in the reference's order (two roundings), and the model's own transformer and get_loss run."""
        by_index = 'ref_index' in batch
        L, C = ctx.table.L, ctx.table.Cdim
        if by_index:
            raw = ctx.build(batch['ref_index'], None, out=None, with_pe=False)      # [b, (K+1)L, C]: sos | refs flipped
            b, K = batch['ref_index'].shape
        else:
            raw = ctx.build_from_features(batch['ref_features'], None, with_pe=False)
            b, K = batch['ref_features'].shape[:2]
        refs = raw[:, L:].reshape(b, K, L, C)                  # similarity rank K-1 ... 0, as batch_forward flips them
        if 'target_features' in batch:
            target = batch['target_features'].to(raw.device, raw.dtype)
        else:
            target = model.encode_vision(batch['video'][:, None])[:, 0]
        vision_emb = torch.cat([refs, target[:, None]], dim=1)
        cond = model.encode_condition(batch['ref_images']) if 'ref_images' in batch else batch.get('condition_emb')
        sos = getattr(model, 'sos_token', None)
        sos = ctx.sos[None] if sos is None else sos
        x = torch.cat([sos.to(raw.dtype).repeat(b, 1, 1), raw[:, L:]], dim=1)
        vision_pe = getattr(model, 'vision_pe', None)
        if vision_pe is not None:
            x = vision_pe(x)
        elif ctx.pos_table is not None and not hasattr(model, 'vision_pe'):
            x = x + ctx.pos_table[:x.size(-2)]
        if cond is not None:
            x = x + cond
        pred = model.transformer(x, ctx.get_mask(K + 1, L))
        pred = pred.reshape(b, K + 1, L, -1)
        if ignore_ref_loss:
            return model.get_loss(pred[:, -1:], vision_emb[:, -1:])
        return model.get_loss(pred, vision_emb)

    def run(batch, last_only: bool):
        """Gather the context and run the transformer; last_only -> [b, L, C] (the slice predict keeps)."""
        cond = model.encode_condition(batch['ref_images']) if 'ref_images' in batch else batch.get('condition_emb')
        by_index = 'ref_index' in batch
        b, K = (batch['ref_index'] if by_index else batch['ref_features']).shape[:2]

        def build(out=None):
            if by_index:
                return ctx.build(batch['ref_index'], cond, out=out)
            return ctx.build_from_features(batch['ref_features'], cond, out=out)

        if transformer is None:
            x = build()
            pred = model.transformer(x, ctx.get_mask(K + 1, ctx.table.L))
            pred = pred.reshape(pred.shape[0], K + 1, ctx.table.L, -1)
            return pred[:, -1] if last_only else pred
        if (K + 1, ctx.table.L) != (transformer.groups, transformer.group_tokens):
            raise ValueError(f"transformer was built for {transformer.groups} groups of "
                             f"{transformer.group_tokens} tokens, got {K + 1} x {ctx.table.L}")
        build(out=transformer.input_view(b))
        if last_only:      # the last layer then only computes the last group's rows (mrag_cama_predict)
            return transformer.predict(b=b)
        pred = transformer.forward(b=b)
        return pred.reshape(pred.shape[0], K + 1, ctx.table.L, -1)

    def batch_forward(batch, return_loss: bool = True, ignore_ref_loss: bool = False):
        if 'ref_index' not in batch and 'ref_features' not in batch:
            return orig(batch, return_loss, ignore_ref_loss)
        if return_loss:
            return loss_forward(batch, ignore_ref_loss)
        return run(batch, last_only=False)

    def predict(batch, do_classifier_free_guidance: bool = False):
        if 'ref_index' not in batch and 'ref_features' not in batch:
            if orig_predict is None:
                raise AttributeError("the wrapped model has no predict() for the video path")
            return orig_predict(batch, do_classifier_free_guidance)
        action_emb = run(batch, last_only=True)
        if do_classifier_free_guidance:
            action_emb = torch.cat([ctx.uncond_action_emb(action_emb.shape[0]).to(action_emb.dtype), action_emb], dim=0)
        return action_emb

    model.batch_forward = batch_forward
    model.predict = predict
    return model
