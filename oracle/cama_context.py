"""ORACLE (test infrastructure) — the context tensor the reference's CAMA transformer consumes.

Two things live here:

1. `context_restatement(...)`: a literal restatement of ActionTransformer.forward lines
   src/projects/condition/module.py:298-301 plus the flip of batch_forward (:318-319):
       x = cat([sos.repeat(b,1,1), rearrange(vision_emb[:, :-1], 'b t l c -> b (t l) c')], 1)
       x = vision_pe(x)            # x + pos_table[:, :len].type_as(x)   (position_embeddings.py:174)
       x += condition_emb
   with vision_emb[:, :-1] == the K retrieved features in REVERSED similarity order, and the
   block-causal mask of get_mask (:131-135).

2. `reference_context(...)`: imports the reference's REAL `ActionTransformer` from
   /root/reference (stubbing the uninstalled lightning / diffusers / kornia / open_clip
   modules), feeds it precomputed features through monkey-patched
   encode_vision / encode_condition and captures the `(x, mask)` handed to
   `self.transformer`. PARITY PINNED: oracle/make_golden.py uses it to write
   tests/golden/cama_context_*.npz, and tests check restatement == captured tensors bit for
   bit. /root/reference only exists in the build container, hence the committed fixtures.

Only tests/, __graft_entry__.smoke(), bench.py's baseline legs and oracle/make_golden.py may
import this module.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """SinusoidPositionalEmbeddings.get_sinusoid_encoding_table
    (src/projects/condition/position_embeddings.py:159-170): float64 numpy angles
    pos / 10000^(2*(j//2)/d), sin on even / cos on odd columns, cast to float32. [1, n, d]"""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0)


def block_causal_mask(num_groups: int, group_tokens: int) -> torch.Tensor:
    """get_mask (module.py:131-135): bool [G*L, G*L], True = blocked; group i sees groups <= i."""
    n = num_groups * group_tokens
    mask = torch.ones(n, n, dtype=torch.bool)
    for i in range(num_groups):
        mask[i * group_tokens:(i + 1) * group_tokens, :(i + 1) * group_tokens] = False
    return mask


def context_restatement(ref_feats: torch.Tensor, sos: torch.Tensor, pos_table: torch.Tensor | None,
                        cond: torch.Tensor | None) -> torch.Tensor:
    """ref_feats [b, K, L, C] in similarity order (0 = most similar; dropped refs already
    replaced by the uncond row), sos [1, L, C], pos_table [1, max_len, C] float32, cond
    [b, (K+1)L, C]. Arithmetic runs in ref_feats.dtype exactly as torch does it."""
    b, K, L, C = ref_feats.shape
    prev = ref_feats.flip(1)                                  # batch_forward: reverse similarity
    x = torch.concat([sos.repeat(b, 1, 1), prev.reshape(b, K * L, C)], dim=1)
    if pos_table is not None:
        x = x + pos_table[:, :x.size(-2)].type_as(x)
    if cond is not None:
        x += cond
    return x


def gather_restatement(table: torch.Tensor, ref_idx: torch.Tensor, uncond_row: torch.Tensor) -> torch.Tensor:
    """[b, K, L, C] features for indices [b, K]; -1 -> uncond row (dataset.py:292,305-310:
    a dropped / unreadable reference is an all-zero clip whose encoding is the uncond row)."""
    b, K = ref_idx.shape
    out = uncond_row[None, None].expand(b, K, *uncond_row.shape).clone()
    ok = ref_idx >= 0
    out[ok] = table[ref_idx[ok]]
    return out


# --- the real reference module ---------------------------------------------------------------
def _install_stubs():
    import torch.nn as nn

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LM(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
        device = property(lambda self: torch.device("cpu"))

        def log(self, *a, **k):
            pass

    if "lightning" not in sys.modules:
        pl = stub("lightning.pytorch", LightningModule=LM, LightningDataModule=object, Callback=object)
        stub("lightning", pytorch=pl)
        stub("lightning.pytorch.utilities", grad_norm=None)
        stub("lightning.pytorch.utilities.types", STEP_OUTPUT=object)
    if "diffusers" not in sys.modules:
        stub("diffusers")
        stub("diffusers.models")
        stub("diffusers.models.lora", LoRALinearLayer=nn.Linear, LoRAConv2dLayer=nn.Conv2d,
             LoRACompatibleConv=type("A", (nn.Conv2d,), {}), LoRACompatibleLinear=type("B", (nn.Linear,), {}))
    for name in ("kornia", "open_clip"):
        if name not in sys.modules:
            stub(name)


def reference_context(ref_feats: torch.Tensor, target_feat: torch.Tensor, cond: torch.Tensor,
                      sos: torch.Tensor, reference_root: str = "/root/reference", max_len: int = 256):
    """Run the reference's own ActionTransformer.batch_forward and capture what reaches the
    causal transformer. ref_feats [b, K, L, C] similarity order, target_feat [b, L, C]
    (features of batch['video']), cond [b, (K+1)L, C], sos [1, L, C].
    Returns (x, mask, pos_table)."""
    import torch.nn as nn
    _install_stubs()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from src.projects.condition.module import ActionTransformer
    from src.projects.condition.position_embeddings import SinusoidPositionalEmbeddings

    b, K, L, C = ref_feats.shape
    dtype = ref_feats.dtype

    class Ident(nn.Module):
        num_queries, output_dim, cross_attention_dim, dim = L, C, C, C

        def forward(self, x):
            return x

    class Capture(nn.Module):
        def __init__(self):
            super().__init__()
            self.seen = None

        def forward(self, x, mask):
            self.seen = (x.clone(), mask.clone())
            return x

    cap = Capture()
    model = ActionTransformer(condition_model=Ident(), condition_proj=Ident(), vision_model=Ident(),
                              vision_proj=Ident(), transformer=cap, condition_pe=None,
                              vision_pe=SinusoidPositionalEmbeddings(C, max_len))
    with torch.no_grad():
        model.sos_token.copy_(sos.float())
    model = model.to(dtype)
    # the "videos" tensor only has to carry identity: encode_vision is replaced by a lookup of
    # the supplied features, in the order batch_forward hands the clips over (flipped refs, target)
    feats_by_slot = torch.cat([ref_feats, target_feat[:, None]], dim=1)      # slot K = target
    slots = torch.arange(K + 1).view(1, K + 1, 1, 1, 1, 1).expand(b, K + 1, 1, 1, 1, 1).to(dtype)

    def encode_vision(videos):
        s = videos[:, :, 0, 0, 0, 0].long()                                   # [b, K+1] slot ids
        return torch.stack([feats_by_slot[i, s[i]] for i in range(b)], 0)

    model.encode_vision = encode_vision
    model.encode_condition = lambda images: cond
    batch = {"ref_videos": slots[:, :K], "video": slots[:, K]}
    with torch.no_grad():
        model.batch_forward(batch, return_loss=False)
    x, mask = cap.seen
    return x, mask, model.vision_pe.pos_table


def tiny_encoder(C: int, heads: int, ff: int, layers: int, seed: int):
    """The CAMA transformer shape (configs/cogvideox/MotionRAG_open.yml:253-267: post-norm, GELU, batch_first) at
    toy size, seeded."""
    import torch.nn as nn
    torch.manual_seed(seed)
    layer = nn.TransformerEncoderLayer(C, heads, ff, 0.0, "gelu", batch_first=True, norm_first=False)
    return nn.TransformerEncoder(layer, layers, enable_nested_tensor=False)


def reference_loss(ref_feats: torch.Tensor, target_feat: torch.Tensor, cond: torch.Tensor, sos: torch.Tensor,
                   encoder, ignore_ref_loss: bool, reference_root: str = "/root/reference", max_len: int = 256):
    """The reference's REAL ActionTransformer.batch_forward(return_loss=True) (module.py:292-311, 317-323; what
    training_step / validation_step call) with a real transformer: returns (mse, smooth, d mse / d sos_token).
    Features are supplied through a patched encode_vision exactly as in reference_context."""
    import copy

    import torch.nn as nn
    _install_stubs()
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from src.projects.condition.module import ActionTransformer
    from src.projects.condition.position_embeddings import SinusoidPositionalEmbeddings
    b, K, L, C = ref_feats.shape

    class Ident(nn.Module):
        num_queries, output_dim, cross_attention_dim, dim = L, C, C, C

        def forward(self, x):
            return x

    model = ActionTransformer(condition_model=Ident(), condition_proj=Ident(), vision_model=Ident(),
                              vision_proj=Ident(), transformer=copy.deepcopy(encoder), condition_pe=None,
                              vision_pe=SinusoidPositionalEmbeddings(C, max_len))
    with torch.no_grad():
        model.sos_token.copy_(sos.float())
    feats_by_slot = torch.cat([ref_feats, target_feat[:, None]], dim=1)
    slots = torch.arange(K + 1).view(1, K + 1, 1, 1, 1, 1).expand(b, K + 1, 1, 1, 1, 1).float()
    model.encode_vision = lambda videos: torch.stack([feats_by_slot[i, videos[i, :, 0, 0, 0, 0].long()] for i in range(b)], 0)
    model.encode_condition = lambda images: cond
    loss = model.batch_forward({"ref_videos": slots[:, :K], "video": slots[:, K]}, return_loss=True,
                               ignore_ref_loss=ignore_ref_loss)
    loss.main.backward()
    return float(loss.mse), float(loss.smooth), model.sos_token.grad.detach().clone()
