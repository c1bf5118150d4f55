#!/usr/bin/env python3
"""Generate bit-packing golden vectors by EXECUTING the reference's own Java source.

There is no JVM in this image, so this script contains a tiny interpreter for the
subset of Java that the machine-generated bodies of

    JavaFastPFOR/src/main/java/me/lemire/integercompression/BitPacking.java
    JavaFastPFOR/src/main/java/me/lemire/integercompression/differential/IntegratedBitPacking.java

use: assignments `out[expr] = expr;` over 32-bit ints with | & << >> >>> + - and
array reads, plus `System.arraycopy(in, inpos, out, outpos, 32)` and
`Arrays.fill(out, outpos, outpos + 32, x)`.  Every fastpack<b>, fastpackwithoutmask<b>,
fastunpack<b>, integratedpack<b>, integratedunpack<b> (b = 0..32) is parsed from the
reference file and run on seeded random inputs; inputs and outputs are written to
tests/golden/bitpacking_golden.json.  The oracle (oracle/codec_oracle.c) and the CUDA
kernels are then checked against those vectors, which pins the bit layout to the
reference's real code rather than to our reading of it.

Run here only (needs /root/reference):  python tests/golden/make_bitpacking_golden.py
"""
from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/JavaFastPFOR/src/main/java/me/lemire/integercompression"
M32 = 0xFFFFFFFF

TOKEN = re.compile(r"\s*(>>>|<<|>>|[()\[\]|&+\-]|\d+|[A-Za-z_]\w*)")


def tokenize(src: str):
    pos, out = 0, []
    src = src.strip()
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos+40]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


class Parser:
    """Java int-expression evaluator (precedence: unary-, + -, shifts, &, |)."""

    def __init__(self, toks, env):
        self.t, self.i, self.env = toks, 0, env

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def eat(self, tok=None):
        cur = self.peek()
        if tok is not None and cur != tok:
            raise SyntaxError(f"expected {tok}, got {cur}")
        self.i += 1
        return cur

    def parse_or(self):
        v = self.parse_and()
        while self.peek() == "|":
            self.eat()
            v = (v | self.parse_and()) & M32
        return v

    def parse_and(self):
        v = self.parse_shift()
        while self.peek() == "&":
            self.eat()
            v = (v & self.parse_shift()) & M32
        return v

    def parse_shift(self):
        v = self.parse_add()
        while self.peek() in ("<<", ">>", ">>>"):
            op = self.eat()
            s = self.parse_add() & 31  # Java masks int shift distances to 5 bits
            if op == "<<":
                v = (v << s) & M32
            elif op == ">>>":
                v = (v & M32) >> s
            else:  # arithmetic
                sv = v - (1 << 32) if v & 0x80000000 else v
                v = (sv >> s) & M32
        return v

    def parse_add(self):
        v = self.parse_unary()
        while self.peek() in ("+", "-"):
            op = self.eat()
            r = self.parse_unary()
            v = (v + r) & M32 if op == "+" else (v - r) & M32
        return v

    def parse_unary(self):
        if self.peek() == "-":
            self.eat()
            return (-self.parse_unary()) & M32
        return self.parse_atom()

    def parse_atom(self):
        tok = self.eat()
        if tok == "(":
            v = self.parse_or()
            self.eat(")")
            return v
        if tok.isdigit():
            return int(tok) & M32
        if self.peek() == "[":
            self.eat("[")
            idx = self.parse_or()
            self.eat("]")
            return int(self.env[tok][idx]) & M32
        return int(self.env[tok]) & M32


def eval_expr(src, env):
    p = Parser(tokenize(src), env)
    v = p.parse_or()
    if p.peek() is not None:
        raise SyntaxError(f"trailing tokens in {src!r}")
    return v


METHOD = re.compile(
    r"protected static void (fastpack|fastpackwithoutmask|fastunpack|integratedpack|integratedunpack)(\d+)\s*\(([^)]*)\)\s*\{",
)


def load_methods(path):
    text = open(path).read()
    methods = {}
    for m in METHOD.finditer(text):
        depth, j = 1, m.end()
        while depth:
            ch = text[j]
            depth += ch == "{"
            depth -= ch == "}"
            j += 1
        body = text[m.end(): j - 1]
        body = re.sub(r"//[^\n]*", "", body)
        methods[(m.group(1), int(m.group(2)))] = [s.strip() for s in body.split(";") if s.strip()]
    return methods


def run_method(stmts, in_arr, n_out, init=0):
    """Execute one generated method body with inpos = outpos = 0 offsets of 3 (to catch offset bugs)."""
    OFF_IN, OFF_OUT = 3, 5
    inbuf = [0] * OFF_IN + [int(x) & M32 for x in in_arr] + [0] * 40
    outbuf = [0] * (OFF_OUT + n_out + 40)
    env = {"in": inbuf, "out": outbuf, "inpos": OFF_IN, "outpos": OFF_OUT, "initoffset": init & M32}
    for st in stmts:
        if st.startswith("System.arraycopy"):
            args = [a.strip() for a in st[st.index("(") + 1: st.rindex(")")].split(",")]
            assert args[0] == "in" and args[2] == "out"
            sp, dp, ln = eval_expr(args[1], env), eval_expr(args[3], env), eval_expr(args[4], env)
            outbuf[dp: dp + ln] = inbuf[sp: sp + ln]
        elif st.startswith("Arrays.fill"):
            args = [a.strip() for a in st[st.index("(") + 1: st.rindex(")")].split(",")]
            assert args[0] == "out"
            a, b, v = eval_expr(args[1], env), eval_expr(args[2], env), eval_expr(args[3], env)
            for i in range(a, b):
                outbuf[i] = v
        else:
            lhs, rhs = st.split("=", 1)
            m = re.fullmatch(r"out\[(.*)\]", lhs.strip(), re.S)
            if not m:
                raise SyntaxError(f"unsupported statement {st[:60]!r}")
            outbuf[eval_expr(m.group(1), env)] = eval_expr(rhs, env)
    assert all(v == 0 for v in outbuf[:OFF_OUT]), "wrote before outpos"
    assert all(v == 0 for v in outbuf[OFF_OUT + n_out:]), "wrote past the expected extent"
    return outbuf[OFF_OUT: OFF_OUT + n_out]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (run in the build container, not on the GPU box)")
    bp = load_methods(os.path.join(REF, "BitPacking.java"))
    ibp = load_methods(os.path.join(REF, "differential", "IntegratedBitPacking.java"))
    rng = np.random.default_rng(0x5EED601D)
    golden = {"source": "interpreted from BitPacking.java / IntegratedBitPacking.java (JavaFastPFOR 0.1.8-SNAPSHOT)",
              "fastpack": [], "fastpackwithoutmask": [], "fastunpack": [], "integratedpack": [], "integratedunpack": []}
    for b in range(33):
        lim = (1 << b)
        for trial in range(3):
            # values that fit in b bits, plus full-range junk for the masked packer
            fit = [int(x) for x in rng.integers(0, lim, 32, dtype=np.uint64)]
            junk = [int(x) for x in rng.integers(0, 1 << 32, 32, dtype=np.uint64)]
            if trial == 0 and b > 0:
                fit[7] = lim - 1
            golden["fastpack"].append({"b": b, "in": junk, "out": run_method(bp[("fastpack", b)], junk, b)})
            golden["fastpack"].append({"b": b, "in": fit, "out": run_method(bp[("fastpack", b)], fit, b)})
            golden["fastpackwithoutmask"].append(
                {"b": b, "in": fit, "out": run_method(bp[("fastpackwithoutmask", b)], fit, b)})
            golden["fastpackwithoutmask"].append(
                {"b": b, "in": junk, "out": run_method(bp[("fastpackwithoutmask", b)], junk, b)})
            words = [int(x) for x in rng.integers(0, 1 << 32, max(b, 1), dtype=np.uint64)][:b]
            golden["fastunpack"].append({"b": b, "in": words, "out": run_method(bp[("fastunpack", b)], words, 32)})
            # sorted list whose gaps fit in b bits (what IntegratedBinaryPacking feeds it), and junk
            init = int(rng.integers(0, 1 << 20))
            gaps = rng.integers(0, lim, 32, dtype=np.uint64)
            sorted_vals = [int((init + int(c)) & M32) for c in np.cumsum(gaps)]
            for vals in (sorted_vals, junk):
                golden["integratedpack"].append(
                    {"b": b, "init": init, "in": vals, "out": run_method(ibp[("integratedpack", b)], vals, b, init)})
            golden["integratedunpack"].append(
                {"b": b, "init": init, "in": words, "out": run_method(ibp[("integratedunpack", b)], words, 32, init)})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bitpacking_golden.json")
    with open(out, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", out, {k: len(v) for k, v in golden.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
