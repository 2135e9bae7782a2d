"""`where` clauses of the reference's search calls (src/data/rag.py:56-57 forwards an SQL string to
LanceDB's `.where(...)`), parsed and evaluated on the host over the table's scalar columns.

The reference's callers only ever write `video != "<own video>"` (src/data/datamodule.py:235,244);
that shape is recognised (`Predicate.simple_exclusion`) and runs on the device as a per-query
excluded group id. Every other predicate of the SQL subset below is evaluated here: in the default
post-filter mode on the k rows a query returned (what `.where()` without `prefilter=True` means in
LanceDB 0.14), in pre-filter mode once over the whole table into a pass/fail row mask that the scan
kernels read as group ids.

Grammar (case-insensitive keywords):
    expr    := term { OR term }
    term    := factor { AND factor }
    factor  := NOT factor | '(' expr ')' | operand tail
    tail    := cmp operand | [NOT] IN '(' literal {',' literal} ')' | IS [NOT] NULL
             | [NOT] LIKE string | [NOT] BETWEEN operand AND operand
    cmp     := = | == | != | <> | < | <= | > | >=
    operand := column | `column` | number | 'string' | "string" | TRUE | FALSE | NULL
A double-quoted token is a string literal, as in the reference's own clause. NULL (None / NaN cells)
follows SQL's three-valued logic: a comparison with NULL is unknown, NOT unknown stays unknown, and only
rows whose predicate is definitely true pass; IS [NOT] NULL tests for it.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Any

import numpy as np

_TOKEN = re.compile(r"""\s*(?:
    (?P<num>[+-]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<bq>`[^`]+`)
  | (?P<op><>|!=|<=|>=|==|=|<|>|\(|\)|,)
  | (?P<id>[A-Za-z_][A-Za-z_0-9.]*)
)""", re.X)
_KEYWORDS = {"AND", "OR", "NOT", "IN", "IS", "NULL", "LIKE", "BETWEEN", "TRUE", "FALSE"}
_CMP = {"=": "eq", "==": "eq", "!=": "ne", "<>": "ne", "<": "lt", "<=": "le", ">": "gt", ">=": "ge"}


class WhereError(ValueError):
    pass


def _tokens(s: str) -> list[tuple[str, Any]]:
    out, pos = [], 0
    s = s.rstrip()
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise WhereError(f"cannot parse where clause at {s[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("num") is not None:
            t = m.group("num")
            out.append(("lit", float(t) if re.search(r"[.eE]", t) else int(t)))
        elif m.group("str") is not None:
            t = m.group("str")
            out.append(("lit", t[1:-1].replace(t[0] * 2, t[0])))
        elif m.group("bq") is not None:
            out.append(("col", m.group("bq")[1:-1]))
        elif m.group("op") is not None:
            out.append(("op", m.group("op")))
        else:
            t = m.group("id")
            out.append(("kw", t.upper()) if t.upper() in _KEYWORDS else ("col", t))
    return out


@dataclass(frozen=True)
class Node:
    kind: str                 # or / and / not / cmp / in / isnull / like / between / const
    args: tuple = ()


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def take(self, kind=None, value=None):
        k, v = self.peek()
        if (kind is not None and k != kind) or (value is not None and v != value):
            raise WhereError(f"unexpected {v!r} in where clause" if k != "end" else "where clause ends early")
        self.i += 1
        return v

    def accept(self, kind, value):
        if self.peek() == (kind, value):
            self.i += 1
            return True
        return False

    def expr(self):
        n = self.term()
        while self.accept("kw", "OR"):
            n = Node("or", (n, self.term()))
        return n

    def term(self):
        n = self.factor()
        while self.accept("kw", "AND"):
            n = Node("and", (n, self.factor()))
        return n

    def operand(self):
        k, v = self.peek()
        self.i += 1
        if k == "col":
            return ("col", v)
        if k == "lit":
            return ("lit", v)
        if k == "kw" and v in ("TRUE", "FALSE"):
            return ("lit", v == "TRUE")
        if k == "kw" and v == "NULL":
            return ("lit", None)
        raise WhereError(f"expected a column or a literal, got {v!r}")

    def factor(self):
        if self.accept("kw", "NOT"):
            return Node("not", (self.factor(),))
        if self.accept("op", "("):
            n = self.expr()
            self.take("op", ")")
            return n
        a = self.operand()
        k, v = self.peek()
        if k == "op" and v in _CMP:
            self.i += 1
            return Node("cmp", (_CMP[v], a, self.operand()))
        neg = self.accept("kw", "NOT")
        if self.accept("kw", "IN"):
            self.take("op", "(")
            vals = [self.operand()]
            while self.accept("op", ","):
                vals.append(self.operand())
            self.take("op", ")")
            if any(x[0] != "lit" for x in vals):
                raise WhereError("IN (...) takes literals")
            n = Node("in", (a, tuple(x[1] for x in vals)))
        elif self.accept("kw", "LIKE"):
            pat = self.operand()
            if pat[0] != "lit" or not isinstance(pat[1], str):
                raise WhereError("LIKE takes a string pattern")
            n = Node("like", (a, pat[1]))
        elif self.accept("kw", "BETWEEN"):
            lo = self.operand()
            self.take("kw", "AND")
            n = Node("between", (a, lo, self.operand()))
        elif not neg and self.accept("kw", "IS"):
            isnot = self.accept("kw", "NOT")
            self.take("kw", "NULL")
            return Node("not", (Node("isnull", (a,)),)) if isnot else Node("isnull", (a,))
        elif not neg and a[0] == "lit" and isinstance(a[1], bool):
            return Node("const", (a[1],))
        elif not neg and a[0] == "col":
            return Node("cmp", ("eq", a, ("lit", True)))        # bare boolean column
        else:
            raise WhereError(f"unexpected {v!r} in where clause" if k != "end" else "where clause ends early")
        return Node("not", (n,)) if neg else n


def _is_null(a: np.ndarray) -> np.ndarray:
    if a.dtype.kind == "f":
        return np.isnan(a)
    if a.dtype.kind == "O":
        return np.fromiter((x is None or (isinstance(x, float) and x != x) for x in a), bool, len(a))
    return np.zeros(len(a), dtype=bool)


def _coerce(lit, like: np.ndarray):
    """A literal as something comparable with the cells of `like` (numbers written as strings compare
    with numeric columns; numbers compare with string columns through their text)."""
    if lit is None:
        return None
    kind = like.dtype.kind
    if kind in "iufb":
        if isinstance(lit, str):
            for cast in (int, float):
                try:
                    return cast(lit)
                except ValueError:
                    pass
            return None            # never equal to a number
        return lit
    if isinstance(lit, (int, float)) and not isinstance(lit, bool):
        return str(lit)
    return lit


class Predicate:
    """A parsed where clause. `evaluate(columns, rows)` -> bool [len(rows)] (True = row passes)."""

    def __init__(self, text: str):
        self.text = text
        try:
            p = _Parser(_tokens(text))
            self.root = p.expr()
            if p.peek()[0] != "end":
                raise WhereError(f"unexpected {p.peek()[1]!r}")
        except WhereError as e:
            raise WhereError(f"unsupported where clause {text!r}: {e}") from None

    def columns(self) -> set[str]:
        found: set[str] = set()

        def walk(x):
            if isinstance(x, Node):
                for a in x.args:
                    walk(a)
            elif isinstance(x, tuple) and len(x) == 2 and x[0] == "col":
                found.add(x[1])
        walk(self.root)
        return found

    def simple_exclusion(self):
        """(column, value) when the clause is exactly `<column> != <literal>`, else None."""
        r = self.root
        if r.kind == "cmp" and r.args[0] == "ne":
            a, b = r.args[1], r.args[2]
            if a[0] == "col" and b[0] == "lit" and b[1] is not None and not isinstance(b[1], bool):
                return a[1], b[1]
        return None

    def evaluate(self, columns: dict, rows=None) -> np.ndarray:
        missing = self.columns() - set(columns)
        if missing:
            raise WhereError(f"where clause names unknown column {sorted(missing)[0]!r}")
        n = len(rows) if rows is not None else len(next(iter(columns.values())))

        def value(opd):
            kind, v = opd
            if kind == "lit":
                return None, v
            col = np.asarray(columns[v])
            return (col if rows is None else col[rows]), None

        def compare(op, a, b):
            """-> (true, false) masks; a NULL on either side leaves both unset (SQL's unknown)."""
            xa, la = value(a)
            xb, lb = value(b)
            none = np.zeros(n, dtype=bool)
            if xa is None and xb is None:       # literal vs literal
                if la is None or lb is None:
                    return none, none
                xa = np.full(n, la)
            if xa is None:                      # literal on the left: mirror
                op = {"lt": "gt", "le": "ge", "gt": "lt", "ge": "le"}.get(op, op)
                xa, la, xb, lb = xb, lb, None, la
            null = _is_null(xa)
            if xb is None:
                if lb is None:
                    return none, none
                rhs = _coerce(lb, xa)
                if rhs is None or (isinstance(rhs, str) and xa.dtype.kind not in "OUS"):
                    # a text literal that is no number never equals a numeric cell
                    return (~null, none) if op == "ne" else (none, ~null)
            else:
                rhs = xb
                null = null | _is_null(xb)
            lhs = xa
            if lhs.dtype.kind == "O" and null.any():
                lhs = np.where(null, rhs if np.ndim(rhs) == 0 else "", lhs)
            with np.errstate(invalid="ignore"):
                res = {"eq": lambda: lhs == rhs, "ne": lambda: lhs != rhs, "lt": lambda: lhs < rhs,
                       "le": lambda: lhs <= rhs, "gt": lambda: lhs > rhs, "ge": lambda: lhs >= rhs}[op]()
            res = np.asarray(res, dtype=bool)
            return res & ~null, ~res & ~null

        def ev(x: Node):
            """Three-valued evaluation: (definitely true, definitely false); neither = unknown (NULL)."""
            if x.kind == "or":
                (ta, fa), (tb, fb) = ev(x.args[0]), ev(x.args[1])
                return ta | tb, fa & fb
            if x.kind == "and":
                (ta, fa), (tb, fb) = ev(x.args[0]), ev(x.args[1])
                return ta & tb, fa | fb
            if x.kind == "not":
                t, f = ev(x.args[0])
                return f, t
            if x.kind == "const":
                c = np.full(n, bool(x.args[0]))
                return c, ~c
            if x.kind == "cmp":
                return compare(*x.args)
            if x.kind == "in":
                t, f = np.zeros(n, dtype=bool), np.ones(n, dtype=bool)
                for lit in x.args[1]:
                    t1, f1 = compare("eq", x.args[0], ("lit", lit))
                    t, f = t | t1, f & f1
                return t, f
            if x.kind == "between":
                (ta, fa), (tb, fb) = compare("ge", x.args[0], x.args[1]), compare("le", x.args[0], x.args[2])
                return ta & tb, fa | fb
            if x.kind == "isnull":
                xa, la = value(x.args[0])
                t = np.full(n, la is None) if xa is None else _is_null(xa)
                return t, ~t
            if x.kind == "like":
                xa, la = value(x.args[0])
                if xa is None:
                    xa = np.full(n, la, dtype=object)
                null = _is_null(xa)
                rx = re.compile("".join(".*" if c == "%" else "." if c == "_" else re.escape(c) for c in x.args[1]) + r"\Z",
                                re.S)
                hit = np.fromiter((isinstance(s, str) and rx.match(s) is not None for s in xa.tolist()), bool, n)
                return hit & ~null, ~hit & ~null
            raise WhereError(f"unknown node {x.kind}")

        return ev(self.root)[0]


_CACHE: dict[str, Predicate] = {}


def parse(text: str) -> Predicate:
    p = _CACHE.get(text)
    if p is None:
        if len(_CACHE) > (1 << 16):
            _CACHE.clear()
        p = _CACHE[text] = Predicate(text)
    return p
