"""Build the motion-feature table for a RAG database (run inside the reference checkout).

    python tools/build_feature_table.py --config configs/cogvideox/MotionRAG_open.yml \
        --ckpt checkpoints/motion_transformer.ckpt --table datasets/rag/openvid.mrag/motion_caption \
        --loader my_loader:load_clip --out datasets/rag/openvid.features

`--loader module:function` names a callable `f(record: dict) -> Tensor[T, C, H, W]` that decodes one table
record ({video, start_sec, end_sec, ...}) exactly as the training data pipeline does — for the reference
that is a thin wrapper over `VideoDataset.get_video(video_info)['video']` (src/data/dataset.py:186-222)
built from the same data config, so that table rows equal what `get_ref_videos` would have decoded.

Needs the reference's own environment (its ActionTransformer, VideoMAE checkpoint and video decoder);
this repo only owns the table format and the driver loop (motionrag_b200/features.py).
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, help="reference yaml holding the condition transformer init args")
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--table", required=True, help="exported table directory (columns.parquet with video/start_sec/end_sec)")
    ap.add_argument("--loader", required=True, help="module:function decoding one record -> Tensor[T,C,H,W]")
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()

    import importlib

    import pandas as pd
    import yaml
    from motionrag_b200.features import build_feature_table
    # everything below is the reference's code, imported from its checkout (cwd)
    sys.path.insert(0, ".")
    from src.projects.condition.module import ActionTransformer   # noqa: E402
    mod, fn = a.loader.split(":")
    load_clip = getattr(importlib.import_module(mod), fn)

    cfg = yaml.safe_load(open(a.config))
    node = cfg
    for key in ("model", "init_args", "condition_transformer", "init_args"):
        node = node.get(key, node) if isinstance(node, dict) else node
    model = ActionTransformer(ckpt_path=a.ckpt, **node).cuda().eval()
    cols = pd.read_parquet(Path(a.table) / "columns.parquet")

    def clips():
        ids, vids = [], []
        for row, rec in enumerate(cols.to_dict("records")):
            try:
                v = load_clip(rec)
            except Exception:   # unreadable clip -> stays the uncond row, like dataset.py:305-310
                continue
            ids.append(row)
            vids.append(v)
            if len(ids) == a.batch:
                yield torch.tensor(ids), torch.stack(vids)
                ids, vids = [], []
        if ids:
            yield torch.tensor(ids), torch.stack(vids)

    print(build_feature_table(model, clips(), len(cols), a.out))


if __name__ == "__main__":
    main()
