"""compute-sanitizer target for the retrieval kernels: K0 upload, K1 (single query with the fused K3 tail, pipelined
launches; 2-4 queries + K3 kernel), K2 single-CTA and CTA-pair forms, pre- and post-filter, fp32 path, top_k > 32 passes,
the two-stage text -> image search, K4 gather.
usage: compute-sanitizer --tool memcheck python scripts/sanitize_search.py"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import motionrag_b200 as m  # noqa: E402

rng = np.random.default_rng(0)
n, dim = 20011, 768
emb = rng.standard_normal((n, dim)).astype(np.float32)
emb /= np.linalg.norm(emb, axis=-1, keepdims=True)
img = rng.standard_normal((n, dim)).astype(np.float32)
table = {"text_embedding": emb, "image_embedding": img / np.linalg.norm(img, axis=-1, keepdims=True), "id": np.arange(n),
         "video": np.array([f"v{j // 3}" for j in range(n)]), "start_sec": (np.arange(n) % 3) * 2.0}
for prefilter in (False, True):
    for path in ("auto", "stream_f32"):
        db = m.RAGDatabase(None, None, "cuda", columns=table, prefilter=prefilter, path=path)
        for _ in range(3):                                         # back-to-back single-query searches pipeline
            r = db.text_search(emb[5] * 3, top_k=12, where='video != "v0"', select=["id"])
        assert r[0]["id"] == 5, r[0]
        for nq in (3, 64, 300):
            out = db.search_batch(emb[:nq] * 2, top_k=12, where=[f'video != "v{j // 3}"' for j in range(nq)], select=["id"])
            assert len(out) == nq
        assert len(db.text_search(emb[7], top_k=40, where="start_sec >= 2", select=["id"])) > 0
        if not prefilter and path == "auto":
            assert len(db.text_image_search(emb[9], table["image_embedding"][9], top_k=(21, 9), select=["id"])) == 9
feat = torch.randn(64, 25, 1024).bfloat16().cuda()
ctx = m.MotionContext(m.FeatureTable(feat), (torch.randn(1, 25, 1024) / 32).bfloat16(), torch.randn(25, 1024).bfloat16(),
                      pe_max_length=256)
idx = torch.randint(-1, 64, (3, 9)).cuda()
x = ctx.build(idx, torch.randn(3, 250, 1024).bfloat16().cuda())
torch.cuda.synchronize()
print("ok", tuple(x.shape))
