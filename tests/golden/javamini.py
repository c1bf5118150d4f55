#!/usr/bin/env python3
"""javamini -- execute the reference's Java sources without a JVM.

There is no JVM (nor javac, Maven, Ant) in this image, so the reference's codecs cannot be run
as shipped.  This module is a small but real *Java-subset -> Python transpiler*: a tokenizer, a
recursive-descent parser for the part of Java the reference's codec classes are written in
(classes, interfaces, fields, constructors, overloaded methods, int/long/byte/short/char/boolean/
double arithmetic with Java's wrap-around and promotion rules, arrays, for / while / do / if /
switch / break / continue / return / throw / try, the conditional operator, compound assignments,
++/--, casts) and a typed code generator that emits one Python class per Java class.  The few JDK
classes the codecs touch (ByteBuffer / IntBuffer, Arrays, System.arraycopy, Math, Integer, Long,
Double, ArrayList) are provided by a tiny runtime below.

It is TEST INFRASTRUCTURE: it reads the Java files where they lie under /root/reference, never
copies them, and is used only by the golden-vector generators in this directory (which run in the
build container, where /root/reference exists).  The vectors it produces -- the reference's own
code executed on seeded inputs -- are what pins oracle/codec_oracle.c and the CUDA kernels to the
reference (tests/golden/codec_streams_golden.json).

Semantics kept exactly: 32-bit / 64-bit two's-complement wrap on + - * << and unary -, logical vs
arithmetic right shifts with Java's 5- / 6-bit shift-count masking, truncating integer division
and remainder, narrowing casts, binary numeric promotion (int op long -> long), implicit narrowing
in compound assignments, left-to-right evaluation of assignment targets, static overload
resolution by parameter types, dynamic dispatch on the receiver.
"""
from __future__ import annotations

import re
import struct

# ------------------------------------------------------------------------------------------------
# runtime (the JDK surface the codecs use)
# ------------------------------------------------------------------------------------------------


class JavaException(Exception):
    def __init__(self, kind, msg=None):
        super().__init__(f"{kind}: {msg}")
        self.kind, self.msg = kind, msg


def _w32(x):
    return ((x + 0x80000000) & 0xFFFFFFFF) - 0x80000000


def _w64(x):
    return ((x + 0x8000000000000000) & 0xFFFFFFFFFFFFFFFF) - 0x8000000000000000


def _idiv(a, b):
    if b == 0:
        raise JavaException("ArithmeticException", "/ by zero")
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def _irem(a, b):
    if b == 0:
        raise JavaException("ArithmeticException", "/ by zero")
    r = abs(a) % abs(b)
    return r if a >= 0 else -r


def _d2i(x, bits):
    if x != x:
        return 0
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    if x >= hi:
        return hi
    if x <= lo:
        return lo
    return int(x)


def _f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def _setidx(a, i, v):
    if i < 0:
        raise JavaException("ArrayIndexOutOfBoundsException", i)
    a[i] = v
    return v


def _setattr(o, name, v):
    setattr(o, name, v)
    return v


def _idx(a, i):
    """Checked read, used only where the index expression can be negative at run time."""
    if i < 0:
        raise JavaException("ArrayIndexOutOfBoundsException", i)
    return a[i]


def _cat(a, b):
    def s(v):
        if v is None:
            return "null"
        if v is True:
            return "true"
        if v is False:
            return "false"
        if hasattr(v, "toString__"):
            return v.toString__()
        return str(v)
    return s(a) + s(b)


def _nlz32(x):
    x &= 0xFFFFFFFF
    return 32 - x.bit_length()


def _nlz64(x):
    x &= 0xFFFFFFFFFFFFFFFF
    return 64 - x.bit_length()


def _ntz(x, bits):
    x &= (1 << bits) - 1
    if x == 0:
        return bits
    return (x & -x).bit_length() - 1


def _arraycopy(src, sp, dst, dp, n):
    if n < 0 or sp < 0 or dp < 0 or sp + n > len(src) or dp + n > len(dst):
        raise JavaException("ArrayIndexOutOfBoundsException", f"arraycopy {sp}+{n}/{len(src)} -> {dp}+{n}/{len(dst)}")
    dst[dp:dp + n] = src[sp:sp + n]


def _fill(a, *args):
    if len(args) == 1:
        a[:] = [args[0]] * len(a)
    else:
        lo, hi, v = args
        if lo < 0 or hi > len(a) or lo > hi:
            raise JavaException("ArrayIndexOutOfBoundsException", f"fill {lo}..{hi}/{len(a)}")
        a[lo:hi] = [v] * (hi - lo)


def _copyof(a, n, dflt):
    return a[:n] + [dflt] * max(0, n - len(a))


def _copyofrange(a, lo, hi, dflt):
    return a[lo:hi] + [dflt] * max(0, hi - len(a))


class ByteOrder:
    LITTLE_ENDIAN = "<"
    BIG_ENDIAN = ">"

    @staticmethod
    def nativeOrder():
        return "<"


class IntBuffer:
    def __init__(self, bb, start, count, order):
        self.bb, self.start, self.count, self.order, self.pos = bb, start, count, order, 0

    def get(self, dst, off, n):
        if n > self.count - self.pos:
            raise JavaException("BufferUnderflowException")
        base = self.start + 4 * self.pos
        vals = struct.unpack_from(f"{self.order}{n}i", self.bb.data, base)
        if off < 0 or off + n > len(dst):
            raise JavaException("IndexOutOfBoundsException", f"{off}+{n}/{len(dst)}")
        dst[off:off + n] = vals
        self.pos += n
        return self

    def put(self, src, off, n):
        if n > self.count - self.pos:
            raise JavaException("BufferOverflowException")
        if off < 0 or off + n > len(src):
            raise JavaException("IndexOutOfBoundsException", f"{off}+{n}/{len(src)}")
        struct.pack_into(f"{self.order}{n}i", self.bb.data, self.start + 4 * self.pos, *src[off:off + n])
        self.pos += n
        return self


class ByteBuffer:
    def __init__(self, n):
        self.data = bytearray(n)
        self.pos, self.lim, self.cap, self.ord = 0, n, n, ">"

    @staticmethod
    def allocateDirect(n):
        return ByteBuffer(n)

    @staticmethod
    def allocate(n):
        return ByteBuffer(n)

    @staticmethod
    def wrap(arr):
        b = ByteBuffer(len(arr))
        b.data[:] = bytes(v & 0xFF for v in arr)
        return b

    def order(self, o=None):
        if o is None:
            return self.ord
        self.ord = o
        return self

    def put(self, *a):
        if len(a) == 1:
            if self.pos >= self.lim:
                raise JavaException("BufferOverflowException")
            self.data[self.pos] = a[0] & 0xFF
            self.pos += 1
        else:
            self.data[a[0]] = a[1] & 0xFF
        return self

    def get(self, *a):
        if a:
            v = self.data[a[0]]
        else:
            if self.pos >= self.lim:
                raise JavaException("BufferUnderflowException")
            v = self.data[self.pos]
            self.pos += 1
        return v - 256 if v > 127 else v

    def _num(self, fmt, size, *a):
        if len(a) == 0:  # relative get
            v = struct.unpack_from(self.ord + fmt, self.data, self.pos)[0]
            self.pos += size
            return v
        raise NotImplementedError

    def getInt(self):
        return self._num("i", 4)

    def getLong(self):
        return self._num("q", 8)

    def getDouble(self):
        return self._num("d", 8)

    def putInt(self, v):
        struct.pack_into(self.ord + "i", self.data, self.pos, v)
        self.pos += 4
        return self

    def putLong(self, v):
        struct.pack_into(self.ord + "q", self.data, self.pos, v)
        self.pos += 8
        return self

    def putDouble(self, v):
        struct.pack_into(self.ord + "d", self.data, self.pos, v)
        self.pos += 8
        return self

    def array(self):
        return [v - 256 if v > 127 else v for v in self.data]

    def clear(self):
        self.pos, self.lim = 0, self.cap
        return self

    def flip(self):
        self.lim, self.pos = self.pos, 0
        return self

    def rewind(self):
        self.pos = 0
        return self

    def position(self, *a):
        if a:
            self.pos = a[0]
            return self
        return self.pos

    def limit(self, *a):
        if a:
            self.lim = a[0]
            return self
        return self.lim

    def remaining(self):
        return self.lim - self.pos

    def capacity(self):
        return self.cap

    def asIntBuffer(self):
        return IntBuffer(self, self.pos, (self.lim - self.pos) // 4, self.ord)


class ArrayList(list):
    def add(self, v):
        self.append(v)
        return True

    def size(self):
        return len(self)

    def get(self, i):
        return self[i]

    def isEmpty(self):
        return not self

    def clear(self):
        del self[:]


def _dbl2long(x):
    return struct.unpack("<q", struct.pack("<d", x))[0]


def _long2dbl(x):
    return struct.unpack("<d", struct.pack("<q", x))[0]


def _flt2int(x):
    return struct.unpack("<i", struct.pack("<f", x))[0]


def _int2flt(x):
    return struct.unpack("<f", struct.pack("<i", x))[0]


RUNTIME = {
    "JavaException": JavaException, "_w32": _w32, "_w64": _w64, "_idiv": _idiv, "_irem": _irem, "_d2i": _d2i,
    "_f32": _f32, "_setidx": _setidx, "_setattr": _setattr, "_idx": _idx, "_cat": _cat, "_nlz32": _nlz32,
    "_nlz64": _nlz64, "_ntz": _ntz, "_arraycopy": _arraycopy, "_fill": _fill, "_copyof": _copyof,
    "_copyofrange": _copyofrange, "ByteOrder": ByteOrder, "ByteBuffer": ByteBuffer, "IntBuffer": IntBuffer,
    "ArrayList": ArrayList, "_dbl2long": _dbl2long, "_long2dbl": _long2dbl, "_flt2int": _flt2int,
    "_int2flt": _int2flt,
}

# (class, method) -> (python callable text taking the emitted arg strings, return type or callable(argtypes))
INTS = ("byte", "short", "char", "int")


def _promote(ts):
    if "double" in ts:
        return "double"
    if "float" in ts:
        return "float"
    if "long" in ts:
        return "long"
    return "int"


EXTERN_STATIC = {
    ("Math", "min"): (lambda a: f"min({a[0]}, {a[1]})", _promote),
    ("Math", "max"): (lambda a: f"max({a[0]}, {a[1]})", _promote),
    ("Math", "abs"): (lambda a: f"abs({a[0]})", _promote),
    ("Integer", "numberOfLeadingZeros"): (lambda a: f"_nlz32({a[0]})", "int"),
    ("Integer", "numberOfTrailingZeros"): (lambda a: f"_ntz({a[0]}, 32)", "int"),
    ("Integer", "bitCount"): (lambda a: f"bin(({a[0]}) & 0xFFFFFFFF).count('1')", "int"),
    ("Integer", "toString"): (lambda a: f"str({a[0]})", "String"),
    ("Integer", "valueOf"): (lambda a: f"({a[0]})", "int"),
    ("Long", "numberOfLeadingZeros"): (lambda a: f"_nlz64({a[0]})", "int"),
    ("Long", "numberOfTrailingZeros"): (lambda a: f"_ntz({a[0]}, 64)", "int"),
    ("Long", "bitCount"): (lambda a: f"bin(({a[0]}) & 0xFFFFFFFFFFFFFFFF).count('1')", "int"),
    ("Double", "doubleToLongBits"): (lambda a: f"_dbl2long({a[0]})", "long"),
    ("Double", "doubleToRawLongBits"): (lambda a: f"_dbl2long({a[0]})", "long"),
    ("Double", "longBitsToDouble"): (lambda a: f"_long2dbl({a[0]})", "double"),
    ("Float", "floatToIntBits"): (lambda a: f"_flt2int({a[0]})", "int"),
    ("Float", "floatToRawIntBits"): (lambda a: f"_flt2int({a[0]})", "int"),
    ("Float", "intBitsToFloat"): (lambda a: f"_int2flt({a[0]})", "float"),
    ("System", "arraycopy"): (lambda a: f"_arraycopy({', '.join(a)})", "void"),
    ("Arrays", "fill"): (lambda a: f"_fill({', '.join(a)})", "void"),
    ("ByteBuffer", "allocateDirect"): (lambda a: f"ByteBuffer.allocateDirect({a[0]})", "ByteBuffer"),
    ("ByteBuffer", "allocate"): (lambda a: f"ByteBuffer.allocate({a[0]})", "ByteBuffer"),
    ("ByteBuffer", "wrap"): (lambda a: f"ByteBuffer.wrap({a[0]})", "ByteBuffer"),
    ("ByteOrder", "nativeOrder"): (lambda a: "ByteOrder.nativeOrder()", "ByteOrder"),
}
EXTERN_INSTANCE = {  # (static type, method) -> return type ("self" = the receiver's type)
    ("ByteBuffer", "order"): "ByteBuffer", ("ByteBuffer", "put"): "ByteBuffer", ("ByteBuffer", "get"): "byte",
    ("ByteBuffer", "clear"): "ByteBuffer", ("ByteBuffer", "flip"): "ByteBuffer", ("ByteBuffer", "rewind"): "ByteBuffer",
    ("ByteBuffer", "position"): "int", ("ByteBuffer", "limit"): "int", ("ByteBuffer", "remaining"): "int",
    ("ByteBuffer", "capacity"): "int", ("ByteBuffer", "asIntBuffer"): "IntBuffer", ("ByteBuffer", "getInt"): "int",
    ("ByteBuffer", "getLong"): "long", ("ByteBuffer", "getDouble"): "double", ("ByteBuffer", "putInt"): "ByteBuffer",
    ("ByteBuffer", "putLong"): "ByteBuffer", ("ByteBuffer", "putDouble"): "ByteBuffer", ("ByteBuffer", "array"): "byte[]",
    ("IntBuffer", "get"): "IntBuffer", ("IntBuffer", "put"): "IntBuffer",
    ("List", "size"): "int", ("List", "add"): "boolean", ("List", "get"): "elem", ("List", "isEmpty"): "boolean",
    ("List", "clear"): "void",
}
BOXED = {"Integer": "int", "Long": "long", "Double": "double", "Float": "float", "Byte": "byte", "Short": "short",
         "Character": "char", "Boolean": "boolean"}
PRIMS = ("int", "long", "byte", "short", "char", "boolean", "double", "float", "void")

# ------------------------------------------------------------------------------------------------
# tokenizer
# ------------------------------------------------------------------------------------------------

TOKEN_RE = re.compile(r"""
  (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
 |(?P<num>0[xX][0-9a-fA-F_]+[lL]?|0[bB][01_]+[lL]?|(?:\d[\d_]*\.\d+|\d[\d_]*)(?:[eE][+-]?\d+)?[fFdDlL]?)
 |(?P<id>[A-Za-z_$][\w$]*)
 |(?P<str>"(?:\\.|[^"\\])*")
 |(?P<chr>'(?:\\.|[^'\\])+')
 |(?P<op>>>>=|<<=|>>=|>>>|\.\.\.|\+\+|--|&&|\|\||==|!=|<=|>=|\+=|-=|\*=|/=|%=|&=|\|=|\^=|->|::|<<|>>|[-+*/%&|^~!<>=?:;,.(){}\[\]@])
""", re.S | re.X)


class Tok:
    __slots__ = ("kind", "text", "line")

    def __init__(self, kind, text, line):
        self.kind, self.text, self.line = kind, text, line

    def __repr__(self):
        return f"{self.text!r}@{self.line}"


def tokenize(src, fname="<java>"):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"{fname}:{line}: cannot tokenize {src[pos:pos + 30]!r}")
        kind = m.lastgroup
        text = m.group(kind)
        if kind != "ws":
            toks.append(Tok(kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(Tok("eof", "<eof>", line))
    return toks


# ------------------------------------------------------------------------------------------------
# parser -> AST (tuples: kind first, source line last for statements)
# ------------------------------------------------------------------------------------------------

MODIFIERS = {"public", "private", "protected", "static", "final", "abstract", "strictfp", "synchronized", "native",
             "transient", "volatile", "default"}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>="}
BINARY_PREC = [["||"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", ">", "<=", ">=", "instanceof"],
               ["<<", ">>", ">>>"], ["+", "-"], ["*", "/", "%"]]


class Parser:
    def __init__(self, src, fname):
        self.toks, self.i, self.fname = tokenize(src, fname), 0, fname
        self.splits = []  # indices where a '>>' / '>>>' token was split to close nested generics (undo log)

    def undo_splits(self, mark):
        while len(self.splits) > mark:
            idx = self.splits.pop()
            a, b = self.toks[idx], self.toks[idx + 1]
            self.toks[idx:idx + 2] = [Tok("op", a.text + b.text, a.line)]

    # -- token helpers
    @property
    def t(self):
        return self.toks[self.i]

    def la(self, k=1):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def err(self, msg):
        raise SyntaxError(f"{self.fname}:{self.t.line}: {msg} (at {self.t.text!r})")

    def at(self, text):
        return self.t.text == text and self.t.kind in ("op", "id")

    def accept(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            self.err(f"expected {text!r}")

    def ident(self):
        if self.t.kind != "id":
            self.err("expected identifier")
        self.i += 1
        return self.toks[self.i - 1].text

    # -- compilation unit
    def parse_unit(self):
        classes = []
        while self.t.kind != "eof":
            if self.at("package") or self.at("import"):
                while not self.accept(";"):
                    self.i += 1
                continue
            if self.accept(";"):
                continue
            classes.extend(self.parse_type_decl())
        return classes

    def modifiers(self):
        mods = set()
        while True:
            if self.t.kind == "id" and self.t.text in MODIFIERS:
                mods.add(self.ident())
            elif self.at("@") and self.la().text != "interface":
                self.i += 1
                self.qualified()
                if self.at("("):
                    self.skip_parens()
            else:
                return mods

    def skip_parens(self):
        depth = 0
        while True:
            if self.at("("):
                depth += 1
            elif self.at(")"):
                depth -= 1
            self.i += 1
            if depth == 0:
                return

    def qualified(self):
        name = self.ident()
        while self.at(".") and self.la().kind == "id":
            self.i += 1
            name = self.ident()
        return name  # the simple name is all we keep

    def parse_type_decl(self):
        mods = self.modifiers()
        kind = self.ident()
        if kind not in ("class", "interface", "enum"):
            self.err("expected class / interface / enum")
        name = self.ident()
        if self.at("<"):
            self.skip_generics()
        supers, ifaces = [], []
        if self.accept("extends"):
            supers.append(self.parse_type())
            while self.accept(","):
                supers.append(self.parse_type())
        if self.accept("implements"):
            ifaces.append(self.parse_type())
            while self.accept(","):
                ifaces.append(self.parse_type())
        cls = {"name": name, "kind": kind, "mods": mods, "extends": supers if kind == "class" else [],
               "implements": ifaces + (supers if kind == "interface" else []), "fields": [], "methods": [], "ctors": [],
               "static_init": [], "inst_init": [], "line": self.t.line}
        out = [cls]
        self.expect("{")
        if kind == "enum":
            consts = []
            while self.t.kind == "id":
                consts.append(self.ident())
                if self.at("("):
                    self.skip_parens()
                if not self.accept(","):
                    break
            self.accept(";")
            cls["enum_constants"] = consts
        while not self.accept("}"):
            if self.accept(";"):
                continue
            out.extend(self.parse_member(cls))
        return out

    def skip_generics(self):
        depth = 0
        while True:
            tx = self.t.text
            if tx == "<":
                depth += 1
            elif tx == ">":
                depth -= 1
            elif tx == ">>":
                depth -= 2
            elif tx == ">>>":
                depth -= 3
            self.i += 1
            if depth <= 0:
                return

    def parse_member(self, cls):
        line = self.t.line
        mods = self.modifiers()
        if self.at("class") or self.at("interface") or self.at("enum") or self.at("@"):
            if self.at("@"):
                self.i += 1  # @interface
            return self.parse_type_decl_with(mods)
        if self.at("{"):
            body = self.parse_block()
            (cls["static_init"] if "static" in mods else cls["inst_init"]).append(body)
            return []
        if self.at("<"):
            self.skip_generics()
        if self.t.kind == "id" and self.t.text == cls["name"] and self.la().text == "(":
            self.ident()
            params = self.parse_params()
            self.skip_throws()
            body = self.parse_block()
            cls["ctors"].append({"params": params, "body": body, "mods": mods, "line": line})
            return []
        typ = self.parse_type()
        name = self.ident()
        if self.at("("):
            params = self.parse_params()
            while self.accept("["):
                self.expect("]")
                typ += "[]"
            self.skip_throws()
            body = None
            if not self.accept(";"):
                body = self.parse_block()
            cls["methods"].append({"name": name, "params": params, "ret": typ, "body": body, "mods": mods, "line": line})
            return []
        while True:
            ftyp = typ
            while self.accept("["):
                self.expect("]")
                ftyp += "[]"
            init = None
            if self.accept("="):
                init = self.parse_var_init()
            cls["fields"].append({"name": name, "type": ftyp, "init": init, "mods": mods, "line": line})
            if self.accept(","):
                name = self.ident()
                continue
            self.expect(";")
            return []

    def parse_type_decl_with(self, mods):
        decls = self.parse_type_decl()
        decls[0]["mods"] |= mods
        return decls

    def skip_throws(self):
        if self.accept("throws"):
            self.qualified()
            while self.accept(","):
                self.qualified()

    def parse_params(self):
        self.expect("(")
        params = []
        while not self.accept(")"):
            self.modifiers()
            typ = self.parse_type()
            if self.accept("..."):
                typ += "[]"
            name = self.ident()
            while self.accept("["):
                self.expect("]")
                typ += "[]"
            params.append((typ, name))
            self.accept(",")
        return params

    def parse_type(self):
        name = self.qualified()
        name = BOXED.get(name, name)
        if self.at("<"):
            self.i += 1
            args = []
            if not self.at(">"):
                while True:
                    if self.accept("?"):
                        if self.accept("extends") or self.accept("super"):
                            args.append(self.parse_type())
                        else:
                            args.append("Object")
                    else:
                        args.append(self.parse_type())
                    if not self.accept(","):
                        break
            self.close_generic()
            if name in ("List", "ArrayList", "Collection", "Iterable", "LinkedList") and args:
                name = f"List<{args[0]}>"
            elif name in ("List", "ArrayList"):
                name = "List<Object>"
        elif name in ("ArrayList", "LinkedList"):
            name = "List<Object>"
        while self.at("[") and self.la().text == "]":
            self.i += 2
            name += "[]"
        return name

    def close_generic(self):
        tx = self.t.text
        if tx == ">":
            self.i += 1
        elif tx in (">>", ">>>"):
            line = self.t.line
            self.toks[self.i:self.i + 1] = [Tok("op", ">", line), Tok("op", tx[1:], line)]
            self.splits.append(self.i)
            self.i += 1
        else:
            self.err("expected '>'")

    # -- statements
    def parse_block(self):
        line = self.t.line
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.parse_stmt())
        return ("block", stmts, line)

    def is_local_decl(self):
        """Speculate: [final] Type Ident ( = | ; | , | : | [ )"""
        save, mark = self.i, len(self.splits)
        ok = False
        try:
            self.modifiers()
            if self.t.kind != "id":
                return False
            if self.t.text in ("return", "throw", "new", "this", "super", "true", "false", "null"):
                return False
            try:
                self.parse_type()
            except SyntaxError:
                return False
            if self.t.kind != "id":
                return False
            nxt = self.la().text
            ok = nxt in ("=", ";", ",", ":", "[")
            return ok
        finally:
            self.i = save
            if not ok:
                self.undo_splits(mark)  # a '>>' split by close_generic belongs to an expression after all

    def parse_local_decl(self):
        line = self.t.line
        self.modifiers()
        typ = self.parse_type()
        decls = []
        while True:
            name = self.ident()
            vtyp = typ
            while self.accept("["):
                self.expect("]")
                vtyp += "[]"
            init = None
            if self.accept("="):
                init = self.parse_var_init()
            decls.append((vtyp, name, init))
            if not self.accept(","):
                break
        return ("local", decls, line)

    def parse_var_init(self):
        if self.at("{"):
            return self.parse_array_init()
        return self.parse_expr()

    def parse_array_init(self):
        self.expect("{")
        elems = []
        while not self.accept("}"):
            elems.append(self.parse_var_init())
            self.accept(",")
        return ("arrayinit", elems)

    def parse_stmt(self):
        line = self.t.line
        tx = self.t.text if self.t.kind in ("id", "op") else None
        if tx == "{":
            return self.parse_block()
        if tx == ";":
            self.i += 1
            return ("empty", line)
        if tx == "if":
            self.i += 1
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            a = self.parse_stmt()
            b = self.parse_stmt() if self.accept("else") else None
            return ("if", c, a, b, line)
        if tx == "while":
            self.i += 1
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            return ("while", c, self.parse_stmt(), line)
        if tx == "do":
            self.i += 1
            body = self.parse_stmt()
            self.expect("while")
            self.expect("(")
            c = self.parse_expr()
            self.expect(")")
            self.expect(";")
            return ("dowhile", body, c, line)
        if tx == "for":
            self.i += 1
            self.expect("(")
            if self.is_local_decl():
                save = self.i
                decl = self.parse_local_decl()
                if self.accept(":"):
                    it = self.parse_expr()
                    self.expect(")")
                    (vt, vn, _), = decl[1]
                    return ("foreach", vt, vn, it, self.parse_stmt(), line)
                init = [decl]
                del save
            else:
                init = []
                if not self.at(";"):
                    init.append(("expr", self.parse_expr(), line))
                    while self.accept(","):
                        init.append(("expr", self.parse_expr(), line))
            self.expect(";")
            cond = None if self.at(";") else self.parse_expr()
            self.expect(";")
            upd = []
            if not self.at(")"):
                upd.append(self.parse_expr())
                while self.accept(","):
                    upd.append(self.parse_expr())
            self.expect(")")
            return ("for", init, cond, upd, self.parse_stmt(), line)
        if tx == "return":
            self.i += 1
            e = None if self.at(";") else self.parse_expr()
            self.expect(";")
            return ("return", e, line)
        if tx == "break":
            self.i += 1
            if self.t.kind == "id":
                self.err("labelled break is not supported")
            self.expect(";")
            return ("break", line)
        if tx == "continue":
            self.i += 1
            if self.t.kind == "id":
                self.err("labelled continue is not supported")
            self.expect(";")
            return ("continue", line)
        if tx == "throw":
            self.i += 1
            e = self.parse_expr()
            self.expect(";")
            return ("throw", e, line)
        if tx == "switch":
            self.i += 1
            self.expect("(")
            sel = self.parse_expr()
            self.expect(")")
            self.expect("{")
            groups = []  # [(labels, stmts)]
            while not self.accept("}"):
                labels = []
                while self.at("case") or self.at("default"):
                    if self.accept("default"):
                        labels.append(None)
                    else:
                        self.i += 1
                        labels.append(self.parse_ternary())
                    self.expect(":")
                stmts = []
                while not (self.at("case") or self.at("default") or self.at("}")):
                    stmts.append(self.parse_stmt())
                groups.append((labels, stmts))
            return ("switch", sel, groups, line)
        if tx == "try":
            self.i += 1
            body = self.parse_block()
            catches, fin = [], None
            while self.accept("catch"):
                self.expect("(")
                self.modifiers()
                types = [self.parse_type()]
                while self.accept("|"):
                    types.append(self.parse_type())
                var = self.ident()
                self.expect(")")
                catches.append((types, var, self.parse_block()))
            if self.accept("finally"):
                fin = self.parse_block()
            return ("try", body, catches, fin, line)
        if tx == "synchronized" and self.la().text == "(":
            self.i += 1
            self.skip_parens()
            return self.parse_block()
        if self.is_local_decl():
            d = self.parse_local_decl()
            self.expect(";")
            return d
        e = self.parse_expr()
        self.expect(";")
        return ("expr", e, line)

    # -- expressions
    def parse_expr(self):
        lhs = self.parse_ternary()
        if self.t.kind == "op" and self.t.text in ASSIGN_OPS:
            op = self.t.text
            self.i += 1
            rhs = self.parse_expr()
            return ("assign", op, lhs, rhs)
        return lhs

    def parse_ternary(self):
        c = self.parse_binary(0)
        if self.accept("?"):
            a = self.parse_expr()
            self.expect(":")
            b = self.parse_ternary() if not self._lambda_like() else self.parse_expr()
            return ("cond", c, a, b)
        return c

    def _lambda_like(self):
        return False

    def parse_binary(self, level):
        if level == len(BINARY_PREC):
            return self.parse_unary()
        lhs = self.parse_binary(level + 1)
        ops = BINARY_PREC[level]
        while (self.t.kind == "op" and self.t.text in ops) or (self.t.kind == "id" and self.t.text == "instanceof"
                                                               and "instanceof" in ops):
            op = self.t.text
            self.i += 1
            if op == "instanceof":
                lhs = ("instanceof", lhs, self.parse_type())
            else:
                lhs = ("binary", op, lhs, self.parse_binary(level + 1))
        return lhs

    def parse_unary(self):
        if self.t.kind == "op":
            tx = self.t.text
            if tx in ("+", "-", "~", "!"):
                self.i += 1
                e = self.parse_unary()
                if tx == "-" and e[0] == "lit" and e[2] in ("int", "long", "double", "float") and not e[3]:
                    return ("lit", -e[1], e[2], True)
                return ("unary", tx, e)
            if tx in ("++", "--"):
                self.i += 1
                return ("preinc", tx, self.parse_unary())
            if tx == "(":
                cast = self.try_cast()
                if cast is not None:
                    return cast
        return self.parse_postfix(self.parse_primary())

    def try_cast(self):
        save, mark = self.i, len(self.splits)
        self.i += 1
        if self.t.kind == "id":
            is_prim = self.t.text in PRIMS
            try:
                typ = self.parse_type()
            except SyntaxError:
                self.i = save
                self.undo_splits(mark)
                return None
            if self.accept(")"):
                nxt = self.t
                if is_prim or "[]" in typ:
                    return ("cast", typ, self.parse_unary())
                if nxt.kind in ("id", "num", "str", "chr") and nxt.text != "instanceof" or nxt.text in ("(", "!", "~"):
                    return ("cast", typ, self.parse_unary())
        self.i = save
        self.undo_splits(mark)
        return None

    def parse_primary(self):
        t = self.t
        if t.kind == "num":
            self.i += 1
            return self.number(t.text)
        if t.kind == "str":
            self.i += 1
            return ("lit", eval(t.text), "String", False)
        if t.kind == "chr":
            self.i += 1
            return ("lit", ord(eval(t.text)), "char", False)
        if t.kind == "op" and t.text == "(":
            self.i += 1
            e = self.parse_expr()
            self.expect(")")
            return ("paren", e)
        if t.kind == "id":
            tx = t.text
            if tx in ("true", "false"):
                self.i += 1
                return ("lit", tx == "true", "boolean", False)
            if tx == "null":
                self.i += 1
                return ("lit", None, "null", False)
            if tx == "this":
                self.i += 1
                if self.at("("):
                    return ("ctorcall", "this", self.parse_args())
                return ("this",)
            if tx == "super":
                self.i += 1
                if self.at("("):
                    return ("ctorcall", "super", self.parse_args())
                self.expect(".")
                name = self.ident()
                return ("supercall", name, self.parse_args())
            if tx == "new":
                return self.parse_new()
            if tx in PRIMS and self.la().text in ("[", "."):
                # int[].class and the like are not supported
                self.err("unsupported primary")
            self.i += 1
            if self.at("("):
                return ("call", None, tx, self.parse_args())
            return ("name", tx)
        self.err("unexpected token in expression")

    def number(self, text):
        tx = text.replace("_", "")
        typ = "int"
        if tx[:2].lower() == "0x":
            if tx[-1] in "lL":
                typ, tx = "long", tx[:-1]
            v = int(tx, 16)
            v = _w64(v) if typ == "long" else _w32(v)
        elif tx[:2].lower() == "0b":
            if tx[-1] in "lL":
                typ, tx = "long", tx[:-1]
            v = int(tx[2:], 2)
            v = _w64(v) if typ == "long" else _w32(v)
        elif tx[-1] in "lL":
            typ, v = "long", int(tx[:-1])
        elif tx[-1] in "fF":
            typ, v = "float", _f32(float(tx[:-1]))
        elif tx[-1] in "dD":
            typ, v = "double", float(tx[:-1])
        elif "." in tx or "e" in tx.lower():
            typ, v = "double", float(tx)
        elif len(tx) > 1 and tx[0] == "0":
            v = _w32(int(tx, 8))
        else:
            v = int(tx)
        return ("lit", v, typ, False)

    def parse_args(self):
        self.expect("(")
        args = []
        while not self.accept(")"):
            args.append(self.parse_expr())
            self.accept(",")
        return args

    def parse_new(self):
        self.expect("new")
        base = self.qualified()
        base = BOXED.get(base, base)
        if self.at("<"):
            save = self.i
            self.i += 1
            elem = "Object"
            if not self.at(">"):
                elem = self.parse_type()
                while self.accept(","):
                    self.parse_type()
            self.close_generic()
            del save
            if base in ("ArrayList", "LinkedList", "List"):
                base = f"List<{elem}>"
        elif base in ("ArrayList", "LinkedList"):
            base = "List<Object>"
        if self.at("["):
            dims, extra = [], 0
            while self.at("["):
                self.i += 1
                if self.accept("]"):
                    extra += 1
                else:
                    if extra:
                        self.err("dimension after []")
                    dims.append(self.parse_expr())
                    self.expect("]")
            init = None
            if self.at("{"):
                init = self.parse_array_init()
            return ("newarray", base, dims, extra, init)
        args = self.parse_args()
        if self.at("{"):
            self.err("anonymous classes are not supported")
        return ("new", base, args)

    def parse_postfix(self, e):
        while True:
            if self.at("."):
                self.i += 1
                if self.at("<"):
                    self.skip_generics()
                name = self.ident()
                if self.at("("):
                    e = ("call", e, name, self.parse_args())
                else:
                    e = ("field", e, name)
            elif self.at("["):
                self.i += 1
                idx = self.parse_expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.t.kind == "op" and self.t.text in ("++", "--"):
                op = self.t.text
                self.i += 1
                e = ("postinc", op, e)
            else:
                return e


# ------------------------------------------------------------------------------------------------
# code generation
# ------------------------------------------------------------------------------------------------


def _mangle_type(t):
    return re.sub(r"[^\w]", "", t.replace("[]", "A").replace("<", "_").replace(">", ""))


def mangle(name, ptypes):
    return name + "__" + "_".join(_mangle_type(t) for t in ptypes)


PYKW = {"in", "is", "not", "and", "or", "def", "del", "from", "global", "lambda", "pass", "with", "yield", "as",
        "assert", "async", "await", "except", "exec", "print", "raise", "None", "True", "False", "nonlocal", "elif",
        "len", "min", "max", "abs", "str", "int", "range", "bin", "self", "list", "type", "float", "ord", "repr",
        "struct", "re", "object", "id", "input", "next", "iter", "map", "filter", "sum", "any", "all", "set", "dict", "tuple"}


def pyname(n):
    return n + "_" if n in PYKW else n


class Scope:
    def __init__(self, parent=None):
        self.vars, self.parent = {}, parent

    def lookup(self, n):
        s = self
        while s:
            if n in s.vars:
                return s.vars[n]
            s = s.parent
        return None


class Gen:
    def __init__(self, classes):
        self.classes = {c["name"]: c for c in classes}
        self.order = [c["name"] for c in classes]
        self.out = []
        self.tmp = 0
        for c in classes:
            for m in c["methods"]:
                m["ptypes"] = [p[0] for p in m["params"]]
                m["py"] = mangle(m["name"], m["ptypes"])
            for k in c["ctors"]:
                k["ptypes"] = [p[0] for p in k["params"]]
                k["py"] = mangle("_ctor", k["ptypes"])

    # -- class table queries
    def supertypes(self, cname):
        seen, todo = [], [cname]
        while todo:
            n = todo.pop(0)
            if n in seen:
                continue
            seen.append(n)
            c = self.classes.get(n)
            if c:
                todo.extend(c["extends"] + c["implements"])
        return seen

    def find_field(self, cname, fname):
        for n in self.supertypes(cname):
            c = self.classes.get(n)
            if c:
                for f in c["fields"]:
                    if f["name"] == fname:
                        return c, f
                if fname in c.get("enum_constants", ()):
                    return c, {"name": fname, "type": n, "mods": {"static", "final"}, "init": None}
        return None, None

    def assignable(self, src, dst):
        if src == dst:
            return 0
        if dst == "Object":
            return 3
        if src == "null":
            return 1 if dst not in PRIMS else None
        order = ["byte", "short", "int", "long", "float", "double"]
        if src == "char" and dst in ("int", "long", "float", "double"):
            return 1
        if src in order and dst in order and order.index(src) <= order.index(dst):
            return 1
        if src.startswith("List<") and dst.startswith("List<"):
            return 1
        if src in self.classes and dst in self.supertypes(src):
            return 2
        if src not in self.classes and src not in PRIMS and dst not in PRIMS and "[]" not in src and "[]" not in dst:
            return 3  # unknown external type: give it the benefit of the doubt
        return None

    def find_method(self, cname, mname, argtypes, line):
        cands = []
        for n in self.supertypes(cname):
            c = self.classes.get(n)
            if not c:
                continue
            for m in c["methods"]:
                if m["name"] == mname and len(m["params"]) == len(argtypes):
                    score = 0
                    for a, p in zip(argtypes, m["ptypes"]):
                        s = self.assignable(a, p)
                        if s is None:
                            score = None
                            break
                        score += s
                    if score is not None:
                        cands.append((score, len(cands), n, m))
        if not cands:
            return None, None
        cands.sort(key=lambda x: x[:2])
        best = cands[0]
        return best[2], best[3]

    # -- emit helpers
    def emit(self, ind, text):
        self.out.append("    " * ind + text)

    def newtmp(self):
        self.tmp += 1
        return f"_t{self.tmp}"

    # -- program
    def generate(self):
        self.emit(0, "# generated by javamini from the reference's Java sources -- do not edit")
        for name in self.order:
            self.gen_class(self.classes[name])
        self.emit(0, "")
        self.emit(0, "def _static_init():")
        self.emit(1, "pending = []")
        n = 0
        for name in self.order:
            c = self.classes[name]
            for f in c["fields"]:
                if "static" in f["mods"] or c["kind"] == "interface":
                    n += 1
                    fn = f"_sinit_{name}_{f['name']}"
                    self.emit(1, f"def {fn}():")
                    self.cls, self.scope, self.static_ctx, self.loops = c, Scope(), True, []
                    self.ret = "void"
                    if f["init"] is None:
                        self.emit(2, f"{name}.{pyname(f['name'])} = {self.default(f['type'])}")
                    else:
                        code = self.conv(self.init_expr(f["init"], f["type"]), f["type"])
                        self.emit(2, f"{name}.{pyname(f['name'])} = {code}")
                    self.emit(1, f"pending.append({fn})")
            for k, blk in enumerate(c["static_init"]):
                fn = f"_sblock_{name}_{k}"
                self.emit(1, f"def {fn}():")
                self.cls, self.scope, self.static_ctx, self.loops, self.ret = c, Scope(), True, [], "void"
                self.gen_body(blk, 2)
                self.emit(1, f"pending.append({fn})")
            for k, cn in enumerate(c.get("enum_constants", ())):
                self.emit(1, f"{name}.{pyname(cn)} = {name}._enum({cn!r}, {k})")
        self.emit(1, "for _round in range(len(pending) + 1):")
        self.emit(2, "later = []")
        self.emit(2, "for f in pending:")
        self.emit(3, "try:")
        self.emit(4, "f()")
        self.emit(3, "except (AttributeError, NameError, TypeError):")
        self.emit(4, "later.append(f)")
        self.emit(2, "if not later:")
        self.emit(3, "break")
        self.emit(2, "pending = later")
        self.emit(1, "else:")
        self.emit(2, "pending[0]()")
        self.emit(0, "_static_init()")
        return "\n".join(self.out) + "\n"

    def gen_class(self, c):
        self.cls = c
        base = ""
        for s in c["extends"]:
            if s in self.classes:
                base = f"({s})"
        self.emit(0, "")
        self.emit(0, f"class {c['name']}{base}:")
        self.emit(1, f"_java_class = {c['name']!r}")
        if c["kind"] == "enum":
            self.emit(1, "@classmethod")
            self.emit(1, "def _enum(cls, name, ordinal):")
            self.emit(2, "o = cls.__new__(cls); o._name = name; o._ordinal = ordinal; return o")
            self.emit(1, "def toString__(self): return self._name")
            self.emit(1, "def name__(self): return self._name")
            self.emit(1, "def ordinal__(self): return self._ordinal")
        # instance field initialisers
        self.emit(1, "def _init_fields(self):")
        self.static_ctx, self.scope, self.loops, self.ret = False, Scope(), [], "void"
        if base:
            self.emit(2, f"{base[1:-1]}._init_fields(self)")
        any_ = False
        for f in c["fields"]:
            if "static" in f["mods"] or c["kind"] == "interface":
                continue
            any_ = True
            if f["init"] is None:
                self.emit(2, f"self.{pyname(f['name'])} = {self.default(f['type'])}")
            else:
                code = self.conv(self.init_expr(f["init"], f["type"]), f["type"])
                self.emit(2, f"self.{pyname(f['name'])} = {code}")
        for blk in c["inst_init"]:
            any_ = True
            self.gen_body(blk, 2)
        if not any_:
            self.emit(2, "pass")
        # constructors
        ctors = c["ctors"] or [{"params": [], "ptypes": [], "py": "_ctor__", "body": ("block", [], 0), "mods": set(), "line": 0}]
        for k in ctors:
            pnames = [pyname(p[1]) for p in k["params"]]
            self.emit(1, f"def {k['py']}(self{''.join(', ' + p for p in pnames)}):")
            self.static_ctx, self.scope, self.loops, self.ret = False, Scope(), [], "void"
            for (pt, pn) in k["params"]:
                self.scope.vars[pn] = pt
            self.gen_body(k["body"], 2)
            self.emit(1, "@staticmethod")
            self.emit(1, f"def {k['py'].replace('_ctor', '_new', 1)}({', '.join(pnames)}):")
            self.emit(2, f"self = {c['name']}.__new__({c['name']})")
            self.emit(2, "self._init_fields()")
            self.emit(2, f"self.{k['py']}({', '.join(pnames)})")
            self.emit(2, "return self")
        c["_ctors"] = ctors
        # methods
        byname = {}
        for m in c["methods"]:
            byname.setdefault((m["name"], len(m["params"])), []).append(m)
            if m["body"] is None:
                continue
            pnames = [pyname(p[1]) for p in m["params"]]
            static = "static" in m["mods"]
            if static:
                self.emit(1, "@staticmethod")
                self.emit(1, f"def {m['py']}({', '.join(pnames)}):")
            else:
                self.emit(1, f"def {m['py']}(self{''.join(', ' + p for p in pnames)}):")
            self.static_ctx, self.scope, self.loops, self.ret = static, Scope(), [], m["ret"]
            for (pt, pn) in m["params"]:
                self.scope.vars[pn] = pt
            self.gen_body(m["body"], 2)
        # convenience aliases for names that are not overloaded at a given arity (used by the Python harness)
        for (nm, ar), ms in byname.items():
            ms = [m for m in ms if m["body"] is not None]
            if len(ms) == 1 and len([1 for (n2, _a) in byname if n2 == nm]) == 1:
                self.emit(1, f"{pyname(nm)} = {ms[0]['py']}")

    def gen_body(self, blk, ind):
        n0 = len(self.out)
        self.scope = Scope(self.scope)
        for s in blk[1]:
            self.stmt(s, ind)
        self.scope = self.scope.parent
        if len(self.out) == n0:
            self.emit(ind, "pass")

    def default(self, t):
        if t in ("int", "long", "byte", "short", "char"):
            return "0"
        if t in ("double", "float"):
            return "0.0"
        if t == "boolean":
            return "False"
        return "None"

    # -- statements
    def stmt(self, s, ind):
        k = s[0]
        if k == "block":
            self.gen_body(s, ind)
        elif k == "empty":
            self.emit(ind, "pass")
        elif k == "local":
            for (vt, vn, init) in s[1]:
                self.scope.vars[vn] = vt
                if init is None:
                    self.emit(ind, f"{pyname(vn)} = {self.default(vt)}")
                else:
                    self.emit(ind, f"{pyname(vn)} = {self.conv(self.init_expr(init, vt), vt)}")
        elif k == "expr":
            self.expr_stmt(s[1], ind)
        elif k == "if":
            self.emit(ind, f"if {self.cond(s[1])}:")
            self.sub(s[2], ind + 1)
            if s[3] is not None:
                if s[3][0] == "if":
                    # keep chains flat
                    self.emit(ind, "else:")
                    self.sub(s[3], ind + 1)
                else:
                    self.emit(ind, "else:")
                    self.sub(s[3], ind + 1)
        elif k == "while":
            self.emit(ind, f"while {self.cond(s[1])}:")
            self.loops.append(("while", None))
            self.sub(s[2], ind + 1)
            self.loops.pop()
        elif k == "dowhile":
            self.emit(ind, "while True:")
            self.loops.append(("dowhile", s[2]))
            self.sub(s[1], ind + 1)
            self.loops.pop()
            self.emit(ind + 1, f"if not ({self.cond(s[2])}):")
            self.emit(ind + 2, "break")
        elif k == "for":
            self.gen_for(s, ind)
        elif k == "foreach":
            _, vt, vn, it, body, _line = s
            code, t = self.expr(it)
            self.scope = Scope(self.scope)
            self.scope.vars[vn] = vt
            self.emit(ind, f"for {pyname(vn)} in list({code}):")
            self.loops.append(("while", None))
            self.sub(body, ind + 1)
            self.loops.pop()
            self.scope = self.scope.parent
        elif k == "return":
            if s[1] is None:
                self.emit(ind, "return")
            else:
                self.emit(ind, f"return {self.conv(self.expr(s[1]), self.ret)}")
        elif k == "break":
            if self.loops and self.loops[-1][0] == "switch":
                raise NotImplementedError(f"{self.cls['name']}:{s[1]}: break inside a switch case body")
            self.emit(ind, "break")
        elif k == "continue":
            loops = [lp for lp in self.loops if lp[0] != "switch"]
            kind, extra = loops[-1]
            if kind == "for":
                for u in extra:
                    self.expr_stmt(u, ind)
                self.emit(ind, "continue")
            elif kind == "dowhile":
                self.emit(ind, f"if not ({self.cond(extra)}):")
                self.emit(ind + 1, "break")
                self.emit(ind, "continue")
            else:
                self.emit(ind, "continue")
        elif k == "throw":
            e = s[1]
            if e[0] == "new":
                args = [self.expr(a)[0] for a in e[2]]
                self.emit(ind, f"raise JavaException({e[1]!r}{''.join(', ' + a for a in args[:1])})")
            else:
                self.emit(ind, f"raise {self.expr(e)[0]}")
        elif k == "switch":
            self.gen_switch(s, ind)
        elif k == "try":
            _, body, catches, fin, _line = s
            self.emit(ind, "try:")
            self.sub(body, ind + 1)
            for (types, var, blk) in catches:
                self.emit(ind, f"except (JavaException, IndexError, ZeroDivisionError) as {pyname(var)}:")
                self.scope = Scope(self.scope)
                self.scope.vars[var] = types[0]
                self.sub(blk, ind + 1)
                self.scope = self.scope.parent
            if fin is not None:
                self.emit(ind, "finally:")
                self.sub(fin, ind + 1)
        else:
            raise NotImplementedError(k)

    def sub(self, s, ind):
        if s[0] == "block":
            self.gen_body(s, ind)
        else:
            self.scope = Scope(self.scope)
            n0 = len(self.out)
            self.stmt(s, ind)
            if len(self.out) == n0:
                self.emit(ind, "pass")
            self.scope = self.scope.parent

    def gen_switch(self, s, ind):
        _, sel, groups, line = s
        code, t = self.expr(sel)
        tmp = self.newtmp()
        self.emit(ind, f"{tmp} = {code}")
        # every group must end in break / return / throw / continue (no fall-through into the next group's code)
        first = True
        default_group = None
        self.loops.append(("switch", None))
        for gi, (labels, stmts) in enumerate(groups):
            last = gi == len(groups) - 1
            if stmts and not last and stmts[-1][0] not in ("break", "return", "throw", "continue"):
                raise NotImplementedError(f"{self.cls['name']}:{line}: switch fall-through")
            if not stmts and not last:
                raise NotImplementedError(f"{self.cls['name']}:{line}: empty case group")
            body = stmts[:-1] if stmts and stmts[-1][0] == "break" else stmts
            if None in labels:
                default_group = body
                labels = [lb for lb in labels if lb is not None]
                if not labels:
                    continue
            conds = " or ".join(f"{tmp} == {self.case_label(lb, t)}" for lb in labels)
            self.emit(ind, f"{'if' if first else 'elif'} {conds}:")
            first = False
            self.gen_body(("block", body, line), ind + 1)
        if default_group is not None:
            if first:
                self.emit(ind, "if True:")
            else:
                self.emit(ind, "else:")
            self.gen_body(("block", default_group, line), ind + 1)
        self.loops.pop()

    def case_label(self, lb, seltype):
        if lb[0] == "name" and seltype in self.classes and self.classes[seltype]["kind"] == "enum":
            return f"{seltype}.{pyname(lb[1])}"
        return self.expr(lb)[0]

    def assigned_names(self, node, acc):
        """Local names written anywhere inside an AST fragment."""
        if isinstance(node, tuple):
            if node and node[0] == "assign" and node[2][0] == "name":
                acc.add(node[2][1])
            if node and node[0] in ("preinc", "postinc") and node[2][0] == "name":
                acc.add(node[2][1])
            if node and node[0] == "local":
                for (_t, n, _i) in node[1]:
                    acc.add(n)
            for x in node:
                self.assigned_names(x, acc)
        elif isinstance(node, list):
            for x in node:
                self.assigned_names(x, acc)
        return acc

    def pure_local_expr(self, e, banned):
        """Literal / local / static-final arithmetic with no calls: safe as a range() bound."""
        k = e[0]
        if k == "lit":
            return True
        if k == "paren":
            return self.pure_local_expr(e[1], banned)
        if k == "name":
            if e[1] in banned:
                return False
            if self.scope.lookup(e[1]) is not None:
                return True
            c, f = self.find_field(self.cls["name"], e[1])
            return f is not None and "final" in f["mods"] and f["type"] in ("int", "long")
        if k == "binary" and e[1] in ("+", "-", "*", "/", "<<", ">>"):
            return self.pure_local_expr(e[2], banned) and self.pure_local_expr(e[3], banned)
        return False

    def gen_for(self, s, ind):
        _, init, cond, upd, body, line = s
        self.scope = Scope(self.scope)
        # fast path: for (int k = a; k < b; ++k | k++ | k += c) with a loop-invariant bound
        if (len(init) == 1 and init[0][0] == "local" and len(init[0][1]) == 1 and init[0][1][0][2] is not None
                and init[0][1][0][0] == "int" and cond is not None and cond[0] == "binary" and cond[1] in ("<", "<=")
                and cond[2] == ("name", init[0][1][0][1]) and len(upd) == 1):
            var = init[0][1][0][1]
            step = None
            u = upd[0]
            if u[0] in ("preinc", "postinc") and u[1] == "++" and u[2] == ("name", var):
                step = 1
            elif u[0] == "assign" and u[1] == "+=" and u[2] == ("name", var) and u[3][0] == "lit" and u[3][2] == "int" \
                    and u[3][1] > 0:
                step = u[3][1]
            written = self.assigned_names(body, set())
            if step is not None and var not in written and self.pure_local_expr(cond[3], written | {var}):
                self.scope.vars[var] = "int"
                lo = self.conv(self.expr(init[0][1][0][2]), "int")
                hi, ht = self.expr(cond[3])
                if ht in ("int", "short", "byte", "char"):
                    if cond[1] == "<=":
                        hi = f"({hi}) + 1"
                    rng = f"range({lo}, {hi})" if step == 1 else f"range({lo}, {hi}, {step})"
                    self.emit(ind, f"for {pyname(var)} in {rng}:")
                    self.loops.append(("while", None))
                    self.sub(body, ind + 1)
                    self.loops.pop()
                    self.scope = self.scope.parent
                    return
        for st in init:
            self.stmt(st, ind)
        self.emit(ind, f"while {self.cond(cond) if cond is not None else 'True'}:")
        self.loops.append(("for", upd))
        self.sub(body, ind + 1)
        self.loops.pop()
        for u in upd:
            self.expr_stmt(u, ind + 1)
        self.scope = self.scope.parent

    # -- expression statements (fast, statement-shaped forms of assignment / ++ / --)
    def has_effects(self, e):
        if isinstance(e, tuple):
            if e and e[0] in ("assign", "preinc", "postinc", "call", "new", "supercall"):
                return True
            return any(self.has_effects(x) for x in e)
        if isinstance(e, list):
            return any(self.has_effects(x) for x in e)
        return False

    def names_read(self, e, acc):
        if isinstance(e, tuple):
            if e and e[0] == "name":
                acc.add(e[1])
            for x in e:
                self.names_read(x, acc)
        elif isinstance(e, list):
            for x in e:
                self.names_read(x, acc)
        return acc

    def expr_stmt(self, e, ind):
        while e[0] == "paren":
            e = e[1]
        k = e[0]
        if k in ("preinc", "postinc"):
            one = ("lit", 1, "int", False)
            return self.expr_stmt(("assign", "+=" if e[1] == "++" else "-=", e[2], one), ind)
        if k == "assign":
            op, lhs, rhs = e[1], e[2], e[3]
            while lhs[0] == "paren":
                lhs = lhs[1]
            if lhs[0] == "index":
                acode, at = self.expr(lhs[1])
                icode, it = self.expr(lhs[2])
                et = at[:-2] if at.endswith("[]") else "Object"
                need_tmp = self.has_effects(lhs[1]) or self.has_effects(lhs[2])
                if not need_tmp and self.has_effects(rhs):
                    written = self.assigned_names(rhs, set())
                    reads = self.names_read(lhs, set())
                    simple = all(self.scope.lookup(n) is not None for n in reads)
                    if (written & reads) or not simple:
                        need_tmp = True
                if need_tmp:
                    ta, ti = self.newtmp(), self.newtmp()
                    self.emit(ind, f"{ta} = {acode}")
                    self.emit(ind, f"{ti} = {icode}")
                    acode, icode = ta, ti
                target = f"{acode}[{icode}]"
                if op == "=":
                    self.emit(ind, f"{target} = {self.conv(self.expr(rhs), et)}")
                else:
                    val = self.binop(op[:-1], (target, et), self.expr(rhs))
                    self.emit(ind, f"{target} = {self.narrow(val, et)}")
                return
            if lhs[0] in ("name", "field"):
                target, tt = self.lvalue(lhs)
                if op == "=":
                    self.emit(ind, f"{target} = {self.conv(self.expr(rhs), tt)}")
                else:
                    val = self.binop(op[:-1], (target, tt), self.expr(rhs))
                    self.emit(ind, f"{target} = {self.narrow(val, tt)}")
                return
        code, _t = self.expr(e)
        self.emit(ind, code)

    def lvalue(self, lhs):
        """(python target text, java type) of a name / field l-value."""
        if lhs[0] == "name":
            n = lhs[1]
            t = self.scope.lookup(n)
            if t is not None:
                return pyname(n), t
            c, f = self.find_field(self.cls["name"], n)
            if f is None:
                raise NameError(f"{self.cls['name']}: unknown name {n}")
            if "static" in f["mods"] or c["kind"] == "interface":
                return f"{c['name']}.{pyname(n)}", f["type"]
            return f"self.{pyname(n)}", f["type"]
        if lhs[0] == "field":
            code, t = self.expr(lhs)
            return code, t
        raise NotImplementedError(lhs[0])

    # -- expressions: return (python text, java type)
    def cond(self, e):
        code, t = self.expr(e)
        return code

    def init_expr(self, init, typ):
        if init[0] == "arrayinit":
            et = typ[:-2]
            return "[" + ", ".join(self.conv(self.init_expr(x, et), et) for x in init[1]) + "]", typ
        return self.expr(init)

    def conv(self, ct, target):
        """Assignment conversion of (code, type) to `target` (widening / constant narrowing / boxing)."""
        code, t = ct
        if t == target or target in ("Object", "void") or t == "null":
            return code
        if target == "float" and t != "float":
            return f"_f32(float({code}))"
        if target == "double" and t in INTS + ("long",):
            return f"float({code})"
        if target == "double":
            return code
        if target == "long" and t in INTS:
            return code
        if target in ("byte", "short", "char") and t in INTS:
            return self.narrow(ct, target)  # constant narrowing (byte b = 5) or a compound assignment
        return code

    def narrow(self, ct, target):
        """Narrowing (cast) conversion."""
        code, t = ct
        if target == t:
            return code
        if target == "boolean" or t == "boolean":
            return code
        if target in ("double",):
            return f"float({code})" if t not in ("double", "float") else code
        if target == "float":
            return f"_f32(float({code}))"
        if t in ("double", "float"):
            if target == "long":
                return f"_d2i({code}, 64)"
            inner = f"_d2i({code}, 32)"
            return inner if target == "int" else self.narrow((inner, "int"), target)
        if target == "long":
            return code
        if target == "int":
            return f"_w32({code})" if t == "long" else code
        if target == "byte":
            return code if t == "byte" else f"((({code}) + 128 & 255) - 128)"
        if target == "short":
            return code if t in ("byte", "short") else f"((({code}) + 32768 & 65535) - 32768)"
        if target == "char":
            return code if t == "char" else f"(({code}) & 65535)"
        return code

    def wrap(self, code, t):
        if t == "int":
            return f"(({code}) + 2147483648 & 4294967295) - 2147483648"
        if t == "long":
            return f"(({code}) + 9223372036854775808 & 18446744073709551615) - 9223372036854775808"
        return code

    def binop(self, op, l, r):
        (lc, lt), (rc, rt) = l, r
        if op == "+" and ("String" in (lt, rt)):
            return f"_cat({lc}, {rc})", "String"
        if op in ("&&", "||"):
            return f"({lc} {'and' if op == '&&' else 'or'} {rc})", "boolean"
        if op in ("==", "!="):
            ref = (lt not in PRIMS and lt != "null") or (rt not in PRIMS and rt != "null") or "null" in (lt, rt)
            if ref and not (lt in PRIMS or rt in PRIMS):
                return f"({lc} {'is' if op == '==' else 'is not'} {rc})", "boolean"
            return f"({lc} {op} {rc})", "boolean"
        if op in ("<", ">", "<=", ">="):
            return f"({lc} {op} {rc})", "boolean"
        if op in ("&", "|", "^") and lt == "boolean" and rt == "boolean":
            return f"({lc} {op} {rc})", "boolean"
        if op in ("<<", ">>", ">>>"):
            t = "long" if lt == "long" else "int"
            bits, mask = (64, 63) if t == "long" else (32, 31)
            m = re.fullmatch(r"\(?(-?\d+)\)?", rc)
            if m:
                sh = str(int(m.group(1)) & mask)
                const = int(sh)
            else:
                sh, const = f"(({rc}) & {mask})", None
            if op == "<<":
                return "(" + self.wrap(f"({lc}) << {sh}", t) + ")", t
            if op == ">>":
                return f"(({lc}) >> {sh})", t
            full = (1 << bits) - 1
            if const is not None and const > 0:
                return f"((({lc}) & {full}) >> {sh})", t
            return "(" + self.wrap(f"(({lc}) & {full}) >> {sh}", t) + ")", t
        t = _promote((lt, rt))
        if t in ("double", "float"):
            pyop = op
            if op == "%":
                code = f"__import__('math').fmod({lc}, {rc})"
            elif op == "/":
                code = (f"(({lc}) / ({rc}) if ({rc}) != 0 else "
                        f"(float('nan') if ({lc}) == 0 or ({lc}) != ({lc}) else float('inf') * (1 if (({lc}) > 0) == "
                        f"(__import__('math').copysign(1.0, {rc}) > 0) else -1)))")
            else:
                code = f"(({lc}) {pyop} ({rc}))"
            if t == "float":
                code = f"_f32({code})"
            return code, t
        if op in ("&", "|", "^"):
            return f"({lc} {op} {rc})", t
        if op in ("+", "-", "*"):
            return "(" + self.wrap(f"{lc} {op} {rc}", t) + ")", t
        if op == "/":
            return "(" + self.wrap(f"_idiv({lc}, {rc})", t) + ")", t
        if op == "%":
            return f"_irem({lc}, {rc})", t
        raise NotImplementedError(op)

    def static_type_name(self, e):
        """If `e` names a class (known or external), return it."""
        if e[0] == "name" and self.scope.lookup(e[1]) is None:
            n = e[1]
            if n in self.classes:
                _c, f = self.find_field(self.cls["name"], n)
                return n if f is None else None
            if n in ("Math", "Integer", "Long", "Double", "Float", "System", "Arrays", "ByteBuffer", "ByteOrder", "String",
                     "Short", "Byte", "Character"):
                return n
        if e[0] == "field" and e[1][0] == "name" and e[1][1] == "System":
            return "System." + e[2]
        return None

    def expr(self, e):
        k = e[0]
        if k == "lit":
            v, t = e[1], e[2]
            if t == "String":
                return repr(v), t
            if t == "boolean":
                return ("True" if v else "False"), t
            if t == "null":
                return "None", t
            if t in ("double", "float"):
                return repr(float(v)), t
            return (f"({v})" if v < 0 else str(v)), t
        if k == "paren":
            code, t = self.expr(e[1])
            return f"({code})", t
        if k == "this":
            return "self", self.cls["name"]
        if k == "name":
            n = e[1]
            t = self.scope.lookup(n)
            if t is not None:
                return pyname(n), t
            c, f = self.find_field(self.cls["name"], n)
            if f is not None:
                if "static" in f["mods"] or c["kind"] == "interface":
                    return f"{c['name']}.{pyname(n)}", f["type"]
                return f"self.{pyname(n)}", f["type"]
            if n in self.classes:
                return n, "class:" + n
            raise NameError(f"{self.cls['name']}: unknown name {n!r}")
        if k == "field":
            cn = self.static_type_name(e[1])
            if cn is not None:
                if cn in self.classes:
                    c, f = self.find_field(cn, e[2])
                    if f is None:
                        raise NameError(f"{cn}.{e[2]}")
                    return f"{c['name']}.{pyname(e[2])}", f["type"]
                if cn == "Integer" and e[2] in ("MAX_VALUE", "MIN_VALUE", "SIZE"):
                    return {"MAX_VALUE": "2147483647", "MIN_VALUE": "(-2147483648)", "SIZE": "32"}[e[2]], "int"
                if cn == "Long" and e[2] in ("MAX_VALUE", "MIN_VALUE", "SIZE"):
                    return {"MAX_VALUE": "9223372036854775807", "MIN_VALUE": "(-9223372036854775808)", "SIZE": "64"}[e[2]], \
                        ("long" if e[2] != "SIZE" else "int")
                if cn == "ByteOrder":
                    return f"ByteOrder.{e[2]}", "ByteOrder"
                if cn == "System":
                    return f"'System.{e[2]}'", "PrintStream"
                raise NameError(f"{cn}.{e[2]}")
            code, t = self.expr(e[1])
            if e[2] == "length" and t.endswith("[]"):
                return f"len({code})", "int"
            if t in self.classes:
                c, f = self.find_field(t, e[2])
                if f is not None:
                    if "static" in f["mods"]:
                        return f"{c['name']}.{pyname(e[2])}", f["type"]
                    return f"{code}.{pyname(e[2])}", f["type"]
            raise NameError(f"{self.cls['name']}: field {e[2]!r} of type {t!r}")
        if k == "index":
            acode, at = self.expr(e[1])
            icode, it = self.expr(e[2])
            et = at[:-2] if at.endswith("[]") else "Object"
            if self.maybe_negative(e[2]):
                return f"_idx({acode}, {icode})", et
            return f"{acode}[{icode}]", et
        if k == "unary":
            code, t = self.expr(e[2])
            op = e[1]
            if op == "!":
                return f"(not {code})", "boolean"
            pt = _promote((t,))
            if op == "+":
                return code, pt
            if op == "~":
                return f"(~({code}))", pt
            if pt in ("double", "float"):
                return f"(-({code}))", pt
            return "(" + self.wrap(f"-({code})", pt) + ")", pt
        if k == "binary":
            op = e[1]
            return self.binop(op, self.expr(e[2]), self.expr(e[3]))
        if k == "cond":
            c = self.cond(e[1])
            a, b = self.expr(e[2]), self.expr(e[3])
            if a[1] == b[1]:
                t = a[1]
            elif a[1] in PRIMS and b[1] in PRIMS and "boolean" not in (a[1], b[1]):
                # constant int fits the other (narrower) operand's type -> that type; else promotion
                if e[2][0] == "lit" and a[1] == "int" and b[1] in ("byte", "short", "char"):
                    t = b[1]
                elif e[3][0] == "lit" and b[1] == "int" and a[1] in ("byte", "short", "char"):
                    t = a[1]
                else:
                    t = _promote((a[1], b[1]))
            else:
                t = a[1] if a[1] != "null" else b[1]
            return f"({self.conv(a, t)} if {c} else {self.conv(b, t)})", t
        if k == "cast":
            inner = self.expr(e[2])
            tt = e[1]
            if tt in PRIMS:
                return self.narrow(inner, tt), tt
            return inner[0], tt
        if k == "instanceof":
            code, _t = self.expr(e[1])
            return f"(_instanceof({code}, {e[2]!r}))", "boolean"
        if k == "assign":
            return self.assign_expr(e)
        if k in ("preinc", "postinc"):
            return self.incdec_expr(e)
        if k == "new":
            return self.new_expr(e)
        if k == "newarray":
            return self.newarray_expr(e)
        if k == "call":
            return self.call_expr(e)
        if k == "ctorcall":
            args = [self.expr(a) for a in e[2]]
            if e[1] == "this":
                ctor = self.pick_ctor(self.cls, [a[1] for a in args])
                argc = ", ".join(self.conv(a, pt) for a, pt in zip(args, ctor["ptypes"]))
                return f"self.{ctor['py']}({argc})", "void"
            sup = [s for s in self.cls["extends"] if s in self.classes]
            if not sup:
                return "None", "void"
            ctor = self.pick_ctor(self.classes[sup[0]], [a[1] for a in args])
            argc = ", ".join(self.conv(a, pt) for a, pt in zip(args, ctor["ptypes"]))
            return f"{sup[0]}.{ctor['py']}(self{', ' if argc else ''}{argc})", "void"
        if k == "supercall":
            args = [self.expr(a) for a in e[2]]
            sup = [s for s in self.cls["extends"] if s in self.classes]
            owner, m = self.find_method(sup[0], e[1], [a[1] for a in args], 0)
            argc = ", ".join(self.conv(a, pt) for a, pt in zip(args, m["ptypes"]))
            return f"{owner}.{m['py']}(self{', ' if argc else ''}{argc})", m["ret"]
        raise NotImplementedError(k)

    def maybe_negative(self, idx):
        """Index expressions that subtract can go below zero; Python would silently wrap around."""
        if idx[0] == "binary" and idx[1] == "-":
            return True
        if idx[0] == "paren":
            return self.maybe_negative(idx[1])
        if idx[0] == "binary" and idx[1] == "+":
            return self.maybe_negative(idx[2]) or self.maybe_negative(idx[3])
        return False

    def assign_expr(self, e):
        """Assignment in value position."""
        op, lhs, rhs = e[1], e[2], e[3]
        while lhs[0] == "paren":
            lhs = lhs[1]
        if lhs[0] == "index":
            acode, at = self.expr(lhs[1])
            icode, it = self.expr(lhs[2])
            et = at[:-2]
            if op == "=":
                return f"_setidx({acode}, {icode}, {self.conv(self.expr(rhs), et)})", et
            ta, ti = self.newtmp(), self.newtmp()
            val = self.binop(op[:-1], (f"{ta}[{ti}]", et), self.expr(rhs))
            return f"_setidx(({ta} := {acode}), ({ti} := {icode}), {self.narrow(val, et)})", et
        target, tt = self.lvalue(lhs)
        if op == "=":
            val = self.conv(self.expr(rhs), tt)
        else:
            val = self.narrow(self.binop(op[:-1], (target, tt), self.expr(rhs)), tt)
        if "." in target:
            obj, attr = target.rsplit(".", 1)
            return f"_setattr({obj}, {attr!r}, {val})", tt
        return f"({target} := {val})", tt

    def incdec_expr(self, e):
        kind, op, lhs = e
        while lhs[0] == "paren":
            lhs = lhs[1]
        delta = "+" if op == "++" else "-"
        if lhs[0] == "index":
            acode, at = self.expr(lhs[1])
            icode, it = self.expr(lhs[2])
            et = at[:-2]
            ta, ti, tv = self.newtmp(), self.newtmp(), self.newtmp()
            newv = self.narrow(self.binop(delta, (tv, et), ("1", "int")), et)
            if kind == "preinc":
                return f"_setidx(({ta} := {acode}), ({ti} := {icode}), [{tv} := {ta}[{ti}], {newv}][1])", et
            return f"[({ta} := {acode}), ({ti} := {icode}), ({tv} := {ta}[{ti}]), _setidx({ta}, {ti}, {newv})][2]", et
        target, tt = self.lvalue(lhs)
        newv = self.narrow(self.binop(delta, (target, tt), ("1", "int")), tt)
        if "." in target:
            obj, attr = target.rsplit(".", 1)
            if kind == "preinc":
                return f"_setattr({obj}, {attr!r}, {newv})", tt
            tv = self.newtmp()
            return f"[({tv} := {target}), _setattr({obj}, {attr!r}, {newv})][0]", tt
        if kind == "preinc":
            return f"({target} := {newv})", tt
        tv = self.newtmp()
        return f"[({tv} := {target}), ({target} := {newv})][0]", tt

    def pick_ctor(self, c, argtypes):
        ctors = c.get("_ctors") or c["ctors"] or [{"params": [], "ptypes": [], "py": "_ctor__"}]
        best = None
        for k in ctors:
            if len(k["ptypes"]) != len(argtypes):
                continue
            score = 0
            for a, p in zip(argtypes, k["ptypes"]):
                s = self.assignable(a, p)
                if s is None:
                    score = None
                    break
                score += s
            if score is not None and (best is None or score < best[0]):
                best = (score, k)
        if best is None:
            raise TypeError(f"no constructor {c['name']}({argtypes})")
        return best[1]

    def new_expr(self, e):
        _, cname, args = e
        a = [self.expr(x) for x in args]
        if cname in self.classes:
            ctor = self.pick_ctor(self.classes[cname], [x[1] for x in a])
            ptypes = ctor["ptypes"]
            py = ctor["py"] if "py" in ctor else mangle("_ctor", ptypes)
            argc = ", ".join(self.conv(x, pt) for x, pt in zip(a, ptypes))
            return f"{cname}.{py.replace('_ctor', '_new', 1)}({argc})", cname
        if cname.startswith("List<"):
            if a and not a[0][1] in INTS:
                return f"ArrayList({a[0][0]})", cname
            return "ArrayList()", cname
        if cname in ("RuntimeException", "IllegalArgumentException", "ArrayIndexOutOfBoundsException", "IOException",
                     "IllegalStateException", "UnsupportedOperationException", "IndexOutOfBoundsException"):
            return f"JavaException({cname!r}{''.join(', ' + x[0] for x in a[:1])})", cname
        if cname == "String":
            return (a[0][0] if a else "''"), "String"
        raise NotImplementedError(f"new {cname}")

    def newarray_expr(self, e):
        _, base, dims, extra, init = e
        full = base + "[]" * (len(dims) + extra)
        if init is not None:
            return self.init_expr(init, full)

        def build(level):
            if level == len(dims) - 1:
                et = base + "[]" * extra
                if extra:
                    return f"[None] * ({self.expr(dims[level])[0]})"
                return f"[{self.default(et)}] * ({self.expr(dims[level])[0]})"
            return f"[{build(level + 1)} for _ in range({self.expr(dims[level])[0]})]"
        return build(0), full

    def call_expr(self, e):
        _, target, name, args = e
        a = [self.expr(x) for x in args]
        atypes = [x[1] for x in a]
        # implicit this / own static
        if target is None:
            owner, m = self.find_method(self.cls["name"], name, atypes, 0)
            if m is None:
                raise NameError(f"{self.cls['name']}: no method {name}({atypes})")
            argc = ", ".join(self.conv(x, pt) for x, pt in zip(a, m["ptypes"]))
            if "static" in m["mods"]:
                return f"{owner}.{m['py']}({argc})", m["ret"]
            return f"self.{m['py']}({argc})", m["ret"]
        cn = self.static_type_name(target)
        if cn is not None:
            if cn in self.classes:
                owner, m = self.find_method(cn, name, atypes, 0)
                if m is None:
                    raise NameError(f"no static method {cn}.{name}({atypes})")
                argc = ", ".join(self.conv(x, pt) for x, pt in zip(a, m["ptypes"]))
                return f"{owner}.{m['py']}({argc})", m["ret"]
            if cn.startswith("System."):
                return f"None", "void"  # System.out/err.println: diagnostics only
            if cn == "Arrays" and name == "copyOf":
                dflt = self.default(atypes[0][:-2])
                return f"_copyof({a[0][0]}, {a[1][0]}, {dflt})", atypes[0]
            if cn == "Arrays" and name == "copyOfRange":
                dflt = self.default(atypes[0][:-2])
                return f"_copyofrange({a[0][0]}, {a[1][0]}, {a[2][0]}, {dflt})", atypes[0]
            if cn == "String" and name == "valueOf":
                return f"_cat('', {a[0][0]})", "String"
            ent = EXTERN_STATIC.get((cn, name))
            if ent is None:
                raise NotImplementedError(f"{cn}.{name}")
            fn, rt = ent
            return fn([x[0] for x in a]), (rt(atypes) if callable(rt) else rt)
        tcode, tt = self.expr(target)
        if tt in self.classes:
            owner, m = self.find_method(tt, name, atypes, 0)
            if m is None:
                if name == "getClass":
                    return f"{tcode}", "Class"
                if name == "toString" and not a:
                    return f"_cat('', {tcode})", "String"
                raise NameError(f"no method {tt}.{name}({atypes})")
            argc = ", ".join(self.conv(x, pt) for x, pt in zip(a, m["ptypes"]))
            if "static" in m["mods"]:
                return f"{owner}.{m['py']}({argc})", m["ret"]
            return f"{tcode}.{m['py']}({argc})", m["ret"]
        if tt == "Class" and name == "getSimpleName":
            return f"{tcode}._java_class", "String"
        if tt == "PrintStream":
            return "None", "void"
        if tt.startswith("List<"):
            rt = EXTERN_INSTANCE.get(("List", name))
            if rt is None:
                raise NotImplementedError(f"List.{name}")
            if rt == "elem":
                rt = tt[5:-1]
            argc = ", ".join(x[0] for x in a)
            return f"{tcode}.{name}({argc})", rt
        if tt == "String":
            if name == "length":
                return f"len({tcode})", "int"
            if name == "charAt":
                return f"ord({tcode}[{a[0][0]}])", "char"
            if name == "equals":
                return f"({tcode} == {a[0][0]})", "boolean"
            if name == "toString":
                return tcode, "String"
        rt = EXTERN_INSTANCE.get((tt, name))
        if rt is not None:
            argc = ", ".join(x[0] for x in a)
            return f"{tcode}.{name}({argc})", rt
        if name == "toString" and not a:
            return f"_cat('', {tcode})", "String"
        if name in ("printStackTrace", "getMessage"):
            return "None", "String"
        raise NotImplementedError(f"{self.cls['name']}: call {name} on {tt}")


# ------------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------------


def extract_members(path, names, class_name):
    """Build a synthetic class from selected members of a big Java file that, as a whole, is outside the subset
    (MapR-DB / Jackson glue).  `names`: method or field names; the text of each member is taken verbatim, by brace
    matching, from the file where it lies."""
    text = open(path).read()
    parts = []
    for nm in names:
        found = False
        for m in re.finditer(r"^[ \t]*(?:(?:public|private|protected|static|final)\s+)*[\w<>\[\], ]+?\s+" + re.escape(nm)
                             + r"\s*(\(|=|;)", text, re.M):
            start = m.start()
            if m.group(1) == "(":
                j = text.index("{", m.end())
                depth, j = 1, j + 1
                while depth:
                    ch = text[j]
                    depth += ch == "{"
                    depth -= ch == "}"
                    j += 1
                parts.append(text[start:j])
            else:
                parts.append(text[start: text.index(";", m.end() - 1) + 1])
            found = True
        if not found:
            raise KeyError(f"{nm} not found in {path}")
    return f"class {class_name} {{\n" + "\n".join(parts) + "\n}\n"


class Program:
    """Transpile a set of Java sources and exec the result; attribute access yields the Python classes."""

    def __init__(self, files=(), sources=()):
        classes = []
        for f in files:
            classes.extend(Parser(open(f).read(), f).parse_unit())
        for (name, src) in sources:
            classes.extend(Parser(src, name).parse_unit())
        self.gen = Gen(classes)
        self.python = self.gen.generate()
        self.ns = dict(RUNTIME)
        self.ns["_instanceof"] = self._instanceof
        exec(compile(self.python, "<javamini>", "exec"), self.ns)

    def _instanceof(self, obj, tname):
        return obj is not None and tname in self.gen.supertypes(getattr(obj, "_java_class", ""))

    def __getattr__(self, name):
        try:
            return self.ns[name]
        except KeyError:
            raise AttributeError(name)

    def new(self, cname, *args):
        """`new cname(args)` with the constructor chosen by arity (and list-vs-int for overloads of one arity)."""
        c = self.gen.classes[cname]
        ctors = [k for k in c["_ctors"] if len(k["ptypes"]) == len(args)]
        if len(ctors) != 1:
            def ok(k):
                return all((isinstance(a, list)) == ("[]" in p) for a, p in zip(args, k["ptypes"]))
            ctors = [k for k in ctors if ok(k)]
        if len(ctors) != 1:
            raise TypeError(f"ambiguous constructor {cname}/{len(args)}")
        return getattr(self.ns[cname], ctors[0]["py"].replace("_ctor", "_new", 1))(*args)

    def method(self, obj_or_class, name, *ptypes):
        """Bound method by Java name and declared parameter types."""
        return getattr(obj_or_class, mangle(name, ptypes))
