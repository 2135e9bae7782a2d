#!/usr/bin/env python
"""Write tests/golden/lancedb_golden.json from the REAL LanceDB — run it once on any machine that has the
reference's engine installed (`pip install lancedb==0.14.0`, reference requirements.txt:18), commit the file, and
the "parity unpinned" note of oracle/flat_search.py is closed: tests/test_lancedb_golden.py then replays the
recorded answers against the oracle (CPU) and against the CUDA drop-in (GPU).

    python tools/make_lancedb_golden.py            # writes tests/golden/lancedb_golden.json
    python tools/make_lancedb_golden.py --check    # also prints how the oracle compares, case by case

The cases (oracle/lancedb_golden.py::CASES) are the assumptions SURVEY §8(c) could not verify offline: default
metric, post-filter semantics of `.where`, nprobes/refine_factor on a flat table, tie order, non-unit and zero
rows, `select`. Tables are regenerated from seeds (numpy PCG64), so the file holds only queries' answers.
"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "lancedb_golden.json"))
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    try:
        import lancedb
    except ImportError as e:
        raise SystemExit(f"lancedb is not importable here ({e}); run this where the reference's requirements are installed")
    from oracle import lancedb_golden as lg
    runs = []
    with tempfile.TemporaryDirectory() as d:
        live = lg.lancedb_search_factory(d)
        for case in lg.CASES:
            runs.append(lg.run_engine(live, case))
            print(f"{case['name']}: {sum(len(r['ids']) for r in runs[-1]['results'])} rows recorded")
    gold = {"engine": "lancedb", "version": getattr(lancedb, "__version__", "?"), "runs": runs,
            "how": "tools/make_lancedb_golden.py; calls as in the reference's src/data/rag.py:54-61"}
    Path(args.out).write_text(json.dumps(gold, indent=1))
    print("wrote", args.out)
    if args.check:
        for run in runs:
            try:
                print(run["case"]["name"], lg.compare_runs(lg.run_engine(lg.oracle_search_one, run["case"]), run))
            except AssertionError as e:
                print(run["case"]["name"], "ORACLE DIFFERS:", e)


if __name__ == "__main__":
    main()
