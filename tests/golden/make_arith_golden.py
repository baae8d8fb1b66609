"""Transcribes the numeric cases of the reference's scalar-function golden files into
tests/golden/arithmetic.json.  Run in the build container (reads /root/reference):

    python tests/golden/make_arith_golden.py

Sources: src/query/functions/tests/it/scalars/testdata/{arithmetic,cast,boolean,comparison}.txt — every case whose
"checked expr" uses only the functions libdbx evaluates (plus, minus, multiply, divide, div, modulo,
unary minus, and / or / not, the six comparisons, is_not_null, CAST / TRY_CAST between numbers and
booleans) over numeric / boolean columns.  For each
case: the typed expression (as a nested list), the input columns, the reference's output type and
its printed output column (values under NULL rows are not significant and not compared).
The three error cases of arithmetic.txt print only the SQL text; they are transcribed by hand with
their line numbers (columns from tests/it/scalars/arithmetic.rs:39-66)."""
import json
import os
import re
import sys

ROOT = "/root/reference/src/query/functions/tests/it/scalars/testdata"
TYPES = {"Int8": "I8", "Int16": "I16", "Int32": "I32", "Int64": "I64", "UInt8": "U8", "UInt16": "U16", "UInt32": "U32", "UInt64": "U64",
         "Float32": "F32", "Float64": "F64", "Boolean": "BOOL"}
SUFFIX = {"i8": "I8", "i16": "I16", "i32": "I32", "i64": "I64", "u8": "U8", "u16": "U16", "u32": "U32", "u64": "U64", "f32": "F32", "f64": "F64"}
FUNCS = {"plus", "minus", "multiply", "divide", "div", "modulo", "and", "or", "not", "eq", "noteq", "lt", "lte", "gt", "gte", "is_not_null"}


class Skip(Exception):
    pass


def parse_type(s):
    s = s.strip()
    nullable = s.endswith(" NULL")
    if nullable:
        s = s[:-5].strip()
    if s not in TYPES:
        raise Skip(s)
    return TYPES[s], nullable


class P:
    def __init__(self, s):
        self.s, self.i = s, 0

    def peek(self, n=1): return self.s[self.i:self.i + n]

    def eat(self, t):
        if not self.s.startswith(t, self.i):
            raise Skip(f"expected {t!r} at {self.s[self.i:]!r}")
        self.i += len(t)

    def until_balanced(self, close):
        depth, j = 0, self.i
        while j < len(self.s):
            c = self.s[j]
            if c in "(<":
                depth += 1
            elif c in ")>":
                if depth == 0 and c == close:
                    break
                depth -= 1
            j += 1
        out = self.s[self.i:j]
        self.i = j
        return out

    def expr(self):
        m = re.match(r"(TRY_CAST|CAST)<", self.s[self.i:])
        if m:
            self.i += len(m.group(0))
            self.until_balanced(">")
            self.eat(">(")
            inner = self.expr()
            self.eat(" AS ")
            t = self.until_balanced(")")
            self.eat(")")
            ty, _ = parse_type(t)
            return ["cast", inner, ty, 1 if m.group(1) == "TRY_CAST" else 0]
        m = re.match(r"([a-z_0-9]+)<", self.s[self.i:])
        if m:
            name = m.group(1)
            if name not in FUNCS:
                raise Skip(name)
            self.i += len(m.group(0))
            self.until_balanced(">")
            self.eat(">(")
            args = [self.expr()]
            while self.peek(2) == ", ":
                self.eat(", ")
                args.append(self.expr())
            self.eat(")")
            if name == "minus" and len(args) == 1:
                name = "negate"
            return ["call", name] + args
        m = re.match(r"(-?[0-9]+(?:\.[0-9]+)?(?:e-?[0-9]+)?)_([iuf][0-9]+)\b", self.s[self.i:])
        if m and not self.s.startswith("(", self.i + len(m.group(0))):
            self.i += len(m.group(0))
            t = SUFFIX[m.group(2)]
            return ["lit", float(m.group(1)) if t[0] == "F" else int(m.group(1)), t]
        m = re.match(r"(true|false)\b", self.s[self.i:])
        if m:
            self.i += len(m.group(0))
            return ["lit", m.group(1) == "true", "BOOL"]
        m = re.match(r"[a-z_][a-z_0-9]*", self.s[self.i:])
        if m:
            self.i += len(m.group(0))
            return ["colname", m.group(0)]
        raise Skip(self.s[self.i:])


def parse_values(t, body):
    if t == "BOOL":  # Boolean([0b_____101])
        bits = []
        for chunk in re.findall(r"0b([_01]+)", body):
            b = chunk.replace("_", "")
            bits.extend(int(c) for c in reversed(b))
        return bits
    vals = [v.strip() for v in body.split(",") if v.strip()]
    out = []
    for v in vals:
        if t[0] == "F":
            out.append({"NaN": "nan", "inf": "inf", "-inf": "-inf"}.get(v, v))
        else:
            out.append(int(v))
    return out


COL_RE = re.compile(r"^(?:Column\()?(?:NullableColumn \{ column: )?([A-Za-z0-9]+)\((\[.*?\])\)(?:, validity: \[(.*?)\] \})?\)?$")


def parse_column(data):
    m = COL_RE.match(data.strip())
    if not m or m.group(1) not in TYPES:
        raise Skip(data)
    t = TYPES[m.group(1)]
    vals = parse_values(t, m.group(2)[1:-1])
    valid = None
    if m.group(3) is not None:
        valid = []
        for chunk in re.findall(r"0b([_01]+)", m.group(3)):
            b = chunk.replace("_", "")
            valid.extend(int(c) for c in reversed(b))
        valid = valid[:len(vals)]
    if t == "BOOL":
        vals = vals[:len(valid)] if valid is not None else vals
    return t, vals, valid


def bind(e, names):
    if e[0] == "colname":
        if e[1] not in names:
            raise Skip("unknown column " + e[1])
        return ["col", names.index(e[1])]
    if e[0] == "cast":
        return ["cast", bind(e[1], names), e[2], e[3]]
    if e[0] == "call":
        return ["call", e[1]] + [bind(a, names) for a in e[2:]]
    return e


def cases_of(fname):
    with open(os.path.join(ROOT, fname)) as f:
        lines = f.read().split("\n")
    blocks, cur, start = [], [], 1
    for ln, line in enumerate(lines, 1):
        if line.startswith("ast ") and cur:
            blocks.append((start, cur))
            cur, start = [], ln
        cur.append(line)
    blocks.append((start, cur))
    out = []
    for start, b in blocks:
        checked = [l for l in b if l.startswith("checked expr")]
        if not checked or any(l.startswith("error") for l in b[:3]):
            continue
        text = checked[0].split(":", 1)[1].strip()
        try:
            tree = P(text)
            e = tree.expr()
            if tree.i != len(text):
                raise Skip("trailing " + text[tree.i:])
            cols, names, output = [], [], None
            if any(l.startswith("evaluation (internal)") for l in b):
                k = [i for i, l in enumerate(b) if l.startswith("evaluation (internal)")][0]
                for l in b[k + 4:]:
                    if not l.startswith("|"):
                        break
                    name, data = [x.strip() for x in l.strip("|").split("|", 1)]
                    data = data.rstrip("|").strip()
                    if name == "Output":
                        output = parse_column(data)
                    else:
                        try:
                            t, v, valid = parse_column(data)
                        except Skip:
                            continue  # a column of another type that this expression may not use
                        names.append(name)
                        cols.append({"type": t, "values": v, "valid": valid})
                rows = len(output[1]) if output[0] != "BOOL" or output[2] is None else len(output[1])
            else:
                ot = [l for l in b if l.startswith("output type")][0].split(":", 1)[1]
                ov = [l for l in b if l.startswith("output  ")][0].split(":", 1)[1].strip()
                t, nullable = parse_type(ot)
                if ov == "NULL":
                    output = (t, [0], [0])
                else:
                    v = (ov == "true") if t == "BOOL" else (ov if t[0] == "F" else int(ov))
                    output = (t, [int(v) if t == "BOOL" else v], None)
                rows = 1
            e = bind(e, names)
            n_rows = len(cols[0]["values"]) if cols else rows
            if output[0] == "BOOL":
                output = (output[0], output[1][:n_rows], output[2])
            out.append({"src": f"{fname}:{start}", "sql": b[0].split(":", 1)[1].strip(), "checked": text, "expr": e, "columns": cols,
                        "rows": n_rows, "out_type": output[0], "out_values": output[1], "out_valid": output[2]})
        except Skip:
            continue
    return out


def main():
    cases = cases_of("arithmetic.txt") + cases_of("cast.txt") + cases_of("boolean.txt") + cases_of("comparison.txt")
    base_cols = [{"type": "U32", "values": [10, 20, 30], "valid": None}, {"type": "I64", "values": [2**63 - 1, -2**63, 0], "valid": None}]
    errors = [
        {"src": "arithmetic.txt:681", "sql": "-g", "expr": ["call", "negate", ["col", 1]], "columns": base_cols, "rows": 3, "error": "number overflowed", "row": 1},
        {"src": "arithmetic.txt:1451", "sql": "c div 0", "expr": ["call", "div", ["col", 0], ["lit", 0, "U8"]], "columns": base_cols, "rows": 3, "error": "divided by zero", "row": 0},
        {"src": "arithmetic.txt:1575", "sql": "c % 0", "expr": ["call", "modulo", ["col", 0], ["lit", 0, "U8"]], "columns": base_cols, "rows": 3, "error": "Division by zero", "row": 0},
    ]
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "arithmetic.json"), "w") as f:
        json.dump({"generated_by": "tests/golden/make_arith_golden.py", "cases": cases, "errors": errors}, f, indent=0)
    print(len(cases), "cases;", len(errors), "error cases", file=sys.stderr)
    for c in cases:
        print(c["src"], c["checked"], "->", c["out_type"], c["out_values"], c["out_valid"])


if __name__ == "__main__":
    main()
