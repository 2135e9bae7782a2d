"""The product's where-clause evaluator (motionrag_b200/where.py; the reference forwards the string to
LanceDB's `.where`, src/data/rag.py:56-57) against an independent SQL engine (SQLite, through
oracle.flat_search.where_mask) on a table with strings, numbers and NULLs."""
import numpy as np
import pytest

from motionrag_b200.where import WhereError, parse
from oracle import flat_search as fs


def table(n=200, seed=0):
    rng = np.random.default_rng(seed)
    video = np.array([f"video_{j // 3:04d}.mp4" for j in range(n)], dtype=object)
    video[rng.integers(0, n, 7)] = None
    start = rng.integers(0, 50, n).astype(np.float64)
    start[rng.integers(0, n, 5)] = np.nan
    return {"video": video, "start_sec": start, "end_sec": start + rng.integers(1, 5, n), "id": np.arange(n),
            "dataset": np.array(["openvid" if j % 4 else "webvid" for j in range(n)]),
            "text_embedding": rng.standard_normal((n, 4)).astype(np.float32)}


CLAUSES = [
    'video != "video_0007.mp4"',                 # the reference's own clause (src/data/datamodule.py:235)
    "video = 'video_0010.mp4'",
    "video <> 'video_0010.mp4' AND start_sec < 20",
    "start_sec >= 10 and start_sec <= 30 or id in (1, 2, 3, 199)",
    "NOT (dataset = 'webvid') AND id > 50",
    "not dataset = 'webvid' or not id > 50",
    "video IS NULL",
    "video IS NOT NULL AND start_sec IS NOT NULL",
    "start_sec is null or end_sec > 40",
    "id BETWEEN 20 AND 40",
    "id not between 20 and 180",
    "video LIKE 'video_001%'",
    "video not like '%7.mp4'",
    "video like 'video_00_1.mp4'",
    "dataset IN ('webvid') and (id < 10 or id >= 190)",
    "dataset not in ('webvid', 'other')",
    "3 < id and id < 9",
    "end_sec > start_sec",
    "end_sec >= start_sec and id != 5",
    "start_sec = 10.0",
    "(id < 5 or (id > 100 and id < 105)) and not video is null",
    "not (video = 'video_0001.mp4')",
    "not (video = 'video_0001.mp4' and start_sec > 3)",
    "video != 'it''s'",
]


@pytest.mark.parametrize("clause", CLAUSES)
def test_where_matches_sqlite(clause):
    cols = table()
    want = fs.where_mask(cols, clause)
    got = parse(clause).evaluate(cols)
    assert got.dtype == bool and np.array_equal(got, want), clause
    rows = np.array([5, 199, 0, 42, 42, 7])
    assert np.array_equal(parse(clause).evaluate(cols, rows), want[rows])


def test_random_clauses_match_sqlite():
    rng = np.random.default_rng(3)
    cols = table(300, seed=1)
    atoms = ["id < {a}", "id >= {a}", "start_sec > {b}", "end_sec <= {b}", "dataset = 'webvid'", "dataset != 'openvid'",
             "video != 'video_{c:04d}.mp4'", "video = 'video_{c:04d}.mp4'", "video is null", "start_sec is not null",
             "id in ({a}, {c}, 7)", "start_sec between {b} and {a}"]

    def gen(depth):
        if depth == 0 or rng.random() < 0.3:
            return rng.choice(atoms).format(a=int(rng.integers(0, 300)), b=int(rng.integers(0, 50)), c=int(rng.integers(0, 100)))
        op = rng.choice(["and", "or", "not"])
        if op == "not":
            return f"not ({gen(depth - 1)})"
        return f"({gen(depth - 1)}) {op} {gen(depth - 1)}"
    for _ in range(150):
        c = gen(3)
        assert np.array_equal(parse(c).evaluate(cols), fs.where_mask(cols, c)), c


def test_simple_exclusion_is_recognised_only_for_the_reference_shape():
    assert parse('video != "a b.mp4"').simple_exclusion() == ("video", "a b.mp4")
    assert parse("id <> 3").simple_exclusion() == ("id", 3)
    for c in ("video = 'x'", "video != 'x' and id > 3", "not video != 'x'", "video != NULL", "3 != id"):
        assert parse(c).simple_exclusion() is None


@pytest.mark.parametrize("clause", ["", "video !=", "video ! 'x'", "(id < 3", "id in 3", "id < 3 extra", "video like 3",
                                     "id between 3", "drop table t"])
def test_malformed_clauses_raise_value_error(clause):
    with pytest.raises(ValueError):
        parse(clause).evaluate(table(10))


def test_number_written_as_text_compares_with_a_numeric_column():
    """`where id != "12"`-style clauses (the reference formats every value into double quotes,
    src/data/datamodule.py:235): the literal is read as a number when the column is numeric."""
    cols = table(30)
    assert parse("id = '12'").evaluate(cols).nonzero()[0].tolist() == [12]
    assert parse('id != "12"').evaluate(cols).sum() == 29
    assert parse("id = 'abc'").evaluate(cols).sum() == 0 and parse("id != 'abc'").evaluate(cols).sum() == 30


def test_unknown_column_raises():
    with pytest.raises(WhereError):
        parse("nope = 3").evaluate(table(10))
