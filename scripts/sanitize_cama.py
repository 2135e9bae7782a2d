"""compute-sanitizer target: one un-graphed forward through the multi-tile GEMM forms (pair kernel with ragged last
pair, bias strips, GELU), the per-head attention K6w (tokens not a multiple of 16) and the warp-per-row LayerNorm.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_cama.py"""
import sys
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import motionrag_b200 as m  # noqa: E402

torch.manual_seed(0)
for (b, G, L, d, heads, dff) in [(15, 10, 25, 1024, 16, 4096), (44, 3, 7, 256, 4, 512)]:
    layer = nn.TransformerEncoderLayer(d, heads, dff, 0.0, "gelu", batch_first=True, norm_first=False, bias=True)
    enc = nn.TransformerEncoder(layer, 1, enable_nested_tensor=False).eval()
    cama = m.CamaTransformer(enc, groups=G, group_tokens=L, max_batch=b, device=0)
    x = torch.randn(b, G * L, d).bfloat16().cuda()
    y = cama.forward(x, use_graph=False)
    p = cama.predict(x, use_graph=False)
    torch.cuda.synchronize()
    print("ok", b, G, L, d, float(y.float().abs().mean()), float(p.float().abs().mean()))
    cama.close()
