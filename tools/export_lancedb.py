#!/usr/bin/env python
"""Export a table of the reference's LanceDB directory (written by its
tools/build_rag_database.py:16-52) to a form `motionrag_b200.RAGDatabase(db_path, table_name)` opens
without the `lancedb` wheel.

Needs `lancedb` (the `.lance` file format is private to that library), so run it once where the
reference's environment is installed:

    python tools/export_lancedb.py datasets/rag/openvid.db motion_caption datasets/rag/openvid.mrag
    python tools/export_lancedb.py datasets/rag/openvid.db motion_caption datasets/rag/openvid.mrag --format parquet

`--format mrag` (default): <out>/<table>/{text_embedding.npy, columns.parquet} — the embedding matrix is
memory-mapped on open. `--format parquet` / `arrow`: the Arrow table as it is (`table.to_arrow()`),
FixedSizeList<f32>[768] column included, in one file <out>/<table>.parquet / .arrow.
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lancedb_path")
    ap.add_argument("table_name")
    ap.add_argument("out_path")
    ap.add_argument("--format", default="mrag", choices=["mrag", "parquet", "arrow"])
    args = ap.parse_args()
    try:
        import lancedb
    except ImportError as e:  # no silent alternative: the source format needs the lance reader
        raise SystemExit(f"lancedb is required to read {args.lancedb_path}: {e}")
    from motionrag_b200 import tables
    from motionrag_b200.rag import save_table
    table = lancedb.connect(args.lancedb_path).open_table(args.table_name).to_arrow()
    out = Path(args.out_path)
    out.mkdir(parents=True, exist_ok=True)
    if args.format == "parquet":
        import pyarrow.parquet as pq
        pq.write_table(table, out / f"{args.table_name}.parquet")
    elif args.format == "arrow":
        import pyarrow.feather as pf
        pf.write_feather(table, out / f"{args.table_name}.arrow", compression="uncompressed")
    else:
        save_table(out, args.table_name, tables.arrow_to_columns(table))
    facts = tables.check_table(tables.read_table(out, args.table_name))
    print("wrote", out, facts)
    for name in tables.VECTOR_COLUMNS:
        if name in facts and not facts[name]["unit_norm"]:
            print(f"note: {name} rows are not unit-norm (norms {facts[name]['norm_min']:.4f}..{facts[name]['norm_max']:.4f}, "
                  f"{facts[name]['zero_rows']} zero rows): LanceDB's L2 ranking then differs from the cosine ranking; the "
                  "store ranks by exact squared L2 of the rows as they are")
    if facts["id_is_row_number"] is False:
        print("note: `id` is not the row number — a feature table built in annotation order is not row-aligned")


if __name__ == "__main__":
    main()
