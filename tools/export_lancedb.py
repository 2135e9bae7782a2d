#!/usr/bin/env python
"""Export a table of the reference's LanceDB directory (written by its
tools/build_rag_database.py:16-52) to the layout `motionrag_b200.RAGDatabase` opens.

Needs `lancedb` (the `lance` file format cannot be read with pyarrow alone), so run it where the
reference's environment is installed:

    python tools/export_lancedb.py datasets/rag/openvid.db motion_caption datasets/rag/openvid.mrag
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lancedb_path")
    ap.add_argument("table_name")
    ap.add_argument("out_path")
    args = ap.parse_args()
    try:
        import lancedb
    except ImportError as e:  # no silent alternative: the source format needs the lance reader
        raise SystemExit(f"lancedb is required to read {args.lancedb_path}: {e}")
    from motionrag_b200.rag import VECTOR_COLUMNS, save_table
    table = lancedb.connect(args.lancedb_path).open_table(args.table_name).to_arrow()
    cols = {}
    for name in table.column_names:
        col = table[name]
        if name in VECTOR_COLUMNS:
            cols[name] = np.stack(col.to_numpy(zero_copy_only=False)).astype(np.float32)
            norms = np.linalg.norm(cols[name], axis=-1)
            print(f"{name}: {cols[name].shape}, row norms in [{norms.min():.4f}, {norms.max():.4f}] "
                  "(the L2 ranking equals the cosine ranking only for unit rows)")
        else:
            cols[name] = col.to_numpy(zero_copy_only=False)
    out = save_table(args.out_path, args.table_name, cols)
    print("wrote", out)


if __name__ == "__main__":
    main()
