#!/usr/bin/env python3
"""Second, independent reader of the reference crate's Butcher tableaux (/root/reference/src/tableau/*.rs) for the Python
restatement (tests/py_restatement.py).  TEST INFRASTRUCTURE.

It shares nothing with tools/gen_tableau.py (which feeds the kernels and the C++ oracle through erk_tableau_data.h): no
regular expressions over assignments and no eval(); instead a tokenizer and a recursive-descent evaluator for the only
expression forms the constructors use (float literals with `_` separators, unary minus, + - * / and parentheses, evaluated
left to right in IEEE doubles exactly as rustc's constant folding does), driven by a statement scanner that understands
`name[i] = e;`, `name[i][j] = e;` and `let mut name = [[0.0; N]; M];`.

    python tests/support/reference_tableaux.py            rewrite tests/golden/reference_tableaux.json from /root/reference
The committed JSON (hex floats) is what travels to the GPU box; tests/test_host_logic.py re-derives it here and checks that
it agrees with the header the other extractor produced.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/src/tableau"
OUT = os.path.join(ROOT, "tests", "golden", "reference_tableaux.json")

# constructor -> (file, our key)
CONSTRUCTORS = {
    "dopri5": ("dorman_prince.rs", "DOPRI5"), "dop853": ("dorman_prince.rs", "DOP853"),
    "euler": ("runge_kutta.rs", "EULER"), "midpoint": ("runge_kutta.rs", "MIDPOINT"), "heun": ("runge_kutta.rs", "HEUN"),
    "ralston": ("runge_kutta.rs", "RALSTON"), "ssp_rk3": ("runge_kutta.rs", "SSP_RK3"), "rk4": ("runge_kutta.rs", "RK4"),
    "three_eighths": ("runge_kutta.rs", "THREE_EIGHTHS"), "rkf45": ("runge_kutta.rs", "RKF45"), "cash_karp": ("runge_kutta.rs", "CASH_KARP"),
    "rkv655e": ("verner.rs", "RKV655E"), "rkv656e": ("verner.rs", "RKV656E"), "rkv766e": ("verner.rs", "RKV766E"), "rkv767e": ("verner.rs", "RKV767E"),
    "rkv877e": ("verner.rs", "RKV877E"), "rkv878e": ("verner.rs", "RKV878E"), "rkv988e": ("verner.rs", "RKV988E"), "rkv989e": ("verner.rs", "RKV989E"),
}
# local array name in the Rust constructors -> tableau field
FIELDS = {"c": "C", "a": "A", "b": "B", "bh": "BH", "er": "ER", "bi": "BI", "bi4": "BI", "bi5": "BI", "bi6": "BI", "bi7": "BI", "bi8": "BI", "bi9": "BI"}


def strip_comments(src):
    out, i, n = [], 0, len(src)
    while i < n:
        if src.startswith("//", i):
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("/*", i):
            i = src.index("*/", i) + 2
        else:
            out.append(src[i])
            i += 1
    return "".join(out)


def function_body(src, name):
    key = "pub fn " + name + "()"
    at = src.index(key)
    i = src.index("{", at)
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                return src[i + 1:j]
        j += 1


def tokenize(expr):
    toks, i = [], 0
    while i < len(expr):
        ch = expr[i]
        if ch.isspace():
            i += 1
        elif ch in "+-*/()":
            toks.append(ch)
            i += 1
        elif ch.isdigit() or ch == ".":
            j = i
            while j < len(expr) and (expr[j].isdigit() or expr[j] in "._eE" or (expr[j] in "+-" and expr[j - 1] in "eE")):
                j += 1
            lit = expr[i:j].replace("_", "")
            if lit.endswith("f64"):
                lit = lit[:-3]
            toks.append(float(lit))
            i = j
        else:
            raise ValueError("unexpected character %r in %r" % (ch, expr))
    return toks


class Eval:
    """expr := term (('+'|'-') term)* ; term := unary (('*'|'/') unary)* ; unary := '-' unary | '(' expr ')' | literal"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self):
        v = self.t[self.i]
        self.i += 1
        return v

    def expr(self):
        v = self.term()
        while self.peek() in ("+", "-"):
            v = v + self.term() if self.take() == "+" else v - self.term()
        return v

    def term(self):
        v = self.unary()
        while self.peek() in ("*", "/"):
            v = v * self.unary() if self.take() == "*" else v / self.unary()
        return v

    def unary(self):
        p = self.peek()
        if p == "-":
            self.take()
            return -self.unary()
        if p == "+":
            self.take()
            return self.unary()
        if p == "(":
            self.take()
            v = self.expr()
            assert self.take() == ")"
            return v
        assert isinstance(p, float), p
        return self.take()


def evaluate(expr):
    e = Eval(tokenize(expr))
    v = e.expr()
    assert e.i == len(e.t), expr
    return v


def statements(body):
    """Split at the semicolons that are not inside brackets (`[[0.0; 7]; 7]` is one expression)."""
    depth, cur = 0, []
    for ch in body:
        if ch in "[({":
            depth += 1
        elif ch in "])}":
            depth -= 1
        if ch == ";" and depth == 0:
            yield "".join(cur)
            cur = []
        else:
            cur.append(ch)
    if cur:
        yield "".join(cur)


def shape_of(decl):
    """`[[0.0; 7]; 7]` -> (7, 7) (outer first), `[0.0; 7]` -> (7,)"""
    dims, depth_sizes = [], []
    body = decl.strip()
    while body.startswith("["):
        inner, _, size = body[1:-1].rpartition(";")
        depth_sizes.append(int(size.strip()))
        body = inner.strip()
    return tuple(depth_sizes)


def read_constructor(name):
    fname, key = CONSTRUCTORS[name]
    body = function_body(strip_comments(open(os.path.join(REF, fname)).read()), name)
    arrays = {}
    for stmt in statements(body):
        s = " ".join(stmt.split())
        if s.startswith("let mut ") and "= [" in s:
            var, _, decl = s[len("let mut "):].partition("=")
            var = var.split(":")[0].strip()
            if var in FIELDS:
                shp = shape_of(decl.strip())
                arrays[var] = [0.0] * shp[0] if len(shp) == 1 else [[0.0] * shp[1] for _ in range(shp[0])]
        elif s.startswith("let ") and "= [" in s and "T::" in s:
            # array literals of the one-stage constructor: `let b = [T::one()];`, `let a = [[T::zero()]];`
            var, _, lit = s[len("let "):].partition("=")
            var = var.strip()
            if var in FIELDS:
                lit = lit.strip().replace("T::zero()", "0.0").replace("T::one()", "1.0")
                nested = lit.startswith("[[")
                vals = [evaluate(x) for x in lit.strip("[] ").split(",") if x.strip()]
                arrays[var] = [vals] if nested else vals
        elif "[" in s and "=" in s and not s.startswith("let "):
            lhs, _, rhs = s.partition("=")
            var = lhs[:lhs.index("[")].strip()
            if var not in arrays or "map(" in rhs:
                continue
            idx = [int(x.split("]")[0]) for x in lhs.split("[")[1:]]
            v = evaluate(rhs)
            if len(idx) == 1:
                arrays[var][idx[0]] = v
            else:
                arrays[var][idx[0]][idx[1]] = v
    out = {}
    for var, val in arrays.items():
        out[FIELDS[var]] = val
    return key, out


def read_all():
    tabs = {}
    for name in CONSTRUCTORS:
        key, t = read_constructor(name)
        tabs[key] = t
    return tabs


def to_hex(v):
    return [to_hex(x) for x in v] if isinstance(v, list) else float(v).hex()


def from_hex(v):
    return [from_hex(x) for x in v] if isinstance(v, list) else float.fromhex(v)


def load_fixture():
    raw = json.load(open(OUT))
    return {k: {f: from_hex(v) for f, v in t.items()} for k, t in raw["tableaux"].items()}


if __name__ == "__main__":
    tabs = read_all()
    json.dump({"source": "/root/reference/src/tableau/{dorman_prince,runge_kutta,verner}.rs (crate v0.6.1), read by tests/support/reference_tableaux.py",
               "tableaux": {k: {f: to_hex(v) for f, v in t.items()} for k, t in tabs.items()}}, open(OUT, "w"), indent=0)
    print("wrote", OUT, {k: sorted(t) for k, t in tabs.items()}, file=sys.stderr)
