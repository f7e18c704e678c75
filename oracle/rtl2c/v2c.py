#!/usr/bin/env python3
"""Mechanical Verilog -> C++ cycle-model translator (test infrastructure; builds oracle/_ref, never the product).

    python oracle/rtl2c/v2c.py --top sw_pe_array -o oracle/_ref/rtl_model.hpp /root/reference/*.v

Reads the mounted reference RTL where it lies, elaborates the module hierarchy under `--top` (parameters resolved per
instance) and writes ONE generated header: a C++ struct per elaborated module with
    bool comb();    one pass over continuous assigns, combinational always blocks and child instances (true = changed)
    void seq();     every `always @(posedge clk)` block, non-blocking: next values into shadow copies / write queues
    void commit();  shadow copies -> registers, queued memory writes -> memories
A cycle is: settle (comb() until nothing changes) -> seq() -> commit().  Two-state: x/z read as 0; one clock domain.

The translation is purely structural -- IEEE 1364-2005 expression sizing / signedness (5.4, 5.5) is implemented once
in `Gen` and applied to whatever the source says; nothing in here knows what the design computes.  The generated file is
a derivative of the reference sources, so it is written under oracle/_ref/ (git-ignored) and never committed.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from vparse import parse_files  # noqa: E402

ARITH = {"+", "-", "*", "/", "%", "&", "|", "^", "~^", "^~"}
COMPARE = {"==", "!=", "===", "!==", "<", "<=", ">", ">="}
LOGICAL = {"&&", "||"}
SHIFT = {"<<", ">>", "<<<", ">>>", "**"}


def mask(w):
    return (1 << w) - 1


def sx(v, w):
    v &= mask(w)
    return v - (1 << w) if v >> (w - 1) else v


class Sig:
    def __init__(self, name, width, signed, kinds, depth=0, base=0, init=None, lsb=0):
        self.name, self.width, self.signed, self.kinds = name, width, signed, set(kinds)
        self.depth, self.base, self.init, self.lsb = depth, base, init, lsb
        self.seq = False            # written by a clocked block

    @property
    def is_mem(self):
        return self.depth > 0


class Module:
    """One elaborated (module, parameter values) pair."""

    def __init__(self, src, cname, params):
        self.src, self.cname, self.params = src, cname, params
        self.sigs = {}
        self.assigns, self.comb_blocks, self.seq_blocks, self.insts = [], [], [], []
        self.implicit = []

    # ---------------- constant evaluation (parameters, ranges, loop bounds) ----------------
    def selfsize(self, e):
        k = e[0]
        if k == "num":
            return (e[2] if e[2] is not None else 32, e[3])
        if k == "str":
            return (32, False)
        if k == "id":
            if e[1] in self.params:
                return self.selfsize(self.params[e[1]])
            s = self.sig(e[1])
            return (s.width, s.signed)
        if k == "index":
            s = self.sig(e[1])
            return (s.width, s.signed) if s.is_mem else (1, False)
        if k == "range":
            return (self.const(e[2]) - self.const(e[3]) + 1, False)
        if k == "concat":
            return (sum(self.selfsize(x)[0] for x in e[1]), False)
        if k == "repl":
            return (self.const(e[1]) * sum(self.selfsize(x)[0] for x in e[2]), False)
        if k == "un":
            return self.selfsize(e[2]) if e[1] in ("-", "+", "~") else (1, False)
        if k == "bin":
            op = e[1]
            if op in COMPARE or op in LOGICAL:
                return (1, False)
            if op in SHIFT:
                return self.selfsize(e[2])
            (wa, sa), (wb, sb) = self.selfsize(e[2]), self.selfsize(e[3])
            return (max(wa, wb), sa and sb)
        if k == "cond":
            (wa, sa), (wb, sb) = self.selfsize(e[2]), self.selfsize(e[3])
            return (max(wa, wb), sa and sb)
        if k == "call":
            w, _ = self.selfsize(e[2][0])
            return (w, e[1] == "$signed")
        raise NotImplementedError(e)

    def interp(self, e, W, S):
        """Value of a CONSTANT expression in a context of width W / signedness S (same rules as Gen.gen)."""
        k = e[0]
        if k == "num":
            w = e[2] if e[2] is not None else 32
            v = e[1] & mask(w)
            return (sx(v, w) if (S and e[3]) else v) & mask(W)
        if k == "str":
            return 0
        if k == "id":
            if e[1] not in self.params:
                raise ValueError(f"{e[1]} is not a constant in {self.src['name']}")
            return self.interp(self.params[e[1]], W, S)
        if k in ("concat", "repl"):
            items = e[1] if k == "concat" else e[2]
            v = 0
            for x in items:
                w, s = self.selfsize(x)
                v = (v << w) | self.interp(x, w, s)
            if k == "repl":
                n, w1 = self.const(e[1]), sum(self.selfsize(x)[0] for x in items)
                v = sum(v << (i * w1) for i in range(n))
            return v & mask(W)
        if k == "un":
            op = e[1]
            if op in ("-", "+", "~"):
                a = self.interp(e[2], W, S)
                return {"-": -a, "+": a, "~": ~a}[op] & mask(W)
            w, s = self.selfsize(e[2])
            a = self.interp(e[2], w, s)
            return {"!": int(a == 0), "|": int(a != 0), "&": int(a == mask(w)), "^": bin(a).count("1") & 1,
                    "~|": int(a == 0), "~&": int(a != mask(w)), "~^": 1 - (bin(a).count("1") & 1)}[op]
        if k == "bin":
            op = e[1]
            if op in COMPARE:
                (wa, sa), (wb, sb) = self.selfsize(e[2]), self.selfsize(e[3])
                w, s = max(wa, wb), sa and sb
                a, b = self.interp(e[2], w, s), self.interp(e[3], w, s)
                if s:
                    a, b = sx(a, w), sx(b, w)
                return int({"==": a == b, "===": a == b, "!=": a != b, "!==": a != b, "<": a < b, "<=": a <= b,
                            ">": a > b, ">=": a >= b}[op])
            if op in LOGICAL:
                a = self.interp(e[2], *self.selfsize(e[2])) != 0
                b = self.interp(e[3], *self.selfsize(e[3])) != 0
                return int(a and b) if op == "&&" else int(a or b)
            if op in SHIFT:
                a = self.interp(e[2], W, S)
                b = self.interp(e[3], *self.selfsize(e[3]))
                if op in ("<<", "<<<"):
                    return (a << b) & mask(W)
                if op == ">>" or (op == ">>>" and not S):
                    return a >> b
                if op == ">>>":
                    return (sx(a, W) >> b) & mask(W)
                return (sx(a, W) if S else a) ** b & mask(W)                 # '**'
            a, b = self.interp(e[2], W, S), self.interp(e[3], W, S)
            if op in ("/", "%"):
                if b == 0:
                    return 0
                if S:
                    a, b = sx(a, W), sx(b, W)
                q = abs(a) // abs(b) * (1 if (a < 0) == (b < 0) else -1)
                return (q if op == "/" else a - q * b) & mask(W)
            return {"+": a + b, "-": a - b, "*": a * b, "&": a & b, "|": a | b, "^": a ^ b, "~^": ~(a ^ b),
                    "^~": ~(a ^ b)}[op] & mask(W)
        if k == "cond":
            c = self.interp(e[1], *self.selfsize(e[1])) != 0
            return self.interp(e[2] if c else e[3], W, S)
        if k == "call":
            w, s = self.selfsize(e[2][0])
            v = self.interp(e[2][0], w, s)
            return (sx(v, w) if (S and e[1] == "$signed") else v) & mask(W)
        raise NotImplementedError(e)

    def const(self, e):
        """Constant expression as a Python int (signed interpretation when the expression is signed)."""
        w, s = self.selfsize(e)
        w = max(w, 32)
        v = self.interp(e, w, s)
        return sx(v, w) if s else v

    def const_num(self, e):
        """Constant expression folded to a ('num', ...) node keeping its self-determined width and signedness."""
        if e[0] == "str":
            return ("num", 0, 32, False)
        w, s = self.selfsize(e)
        return ("num", self.interp(e, w, s), w, s)

    def sig(self, name):
        if name not in self.sigs:
            self.sigs[name] = Sig(name, 1, False, {"wire"})             # implicit net (Verilog default nettype)
            self.implicit.append(name)
        return self.sigs[name]


def subst(e, var, val):
    """Replace a loop variable by a 32-bit signed constant."""
    if isinstance(e, tuple):
        if e[0] == "id" and e[1] == var:
            return ("num", val & mask(32), 32, True)
        return tuple(subst(x, var, val) for x in e)
    if isinstance(e, list):
        return [subst(x, var, val) for x in e]
    return e


class Design:
    def __init__(self, sources):
        self.sources = sources
        self.modules = {}            # key -> Module
        self.order = []

    def elaborate(self, name, overrides=None):
        src = self.sources[name]
        m = Module(src, "", {})
        overrides = overrides or {}
        for pname, pexpr in src["params"]:
            m.params[pname] = overrides[pname] if pname in overrides else m.const_num(pexpr)
        key = (name, tuple(sorted((k, v[1], v[2], v[3]) for k, v in m.params.items() if k in overrides)))
        if key in self.modules:
            return self.modules[key]
        m.cname = f"M_{name}" + (f"__{sum(1 for k in self.modules if k[0] == name)}" if key[1] else "")
        self.modules[key] = m

        for d in src["decls"]:
            width, lsb = 1, 0
            if d["range"]:
                a, b = m.const(d["range"][0]), m.const(d["range"][1])
                width, lsb = abs(a - b) + 1, min(a, b)
                assert a >= b, f"{name}.{d['name']}: ascending packed range not supported"
            depth = base = 0
            if d["mem"]:
                a, b = m.const(d["mem"][0]), m.const(d["mem"][1])
                depth, base = abs(a - b) + 1, min(a, b)
            if width > 64:
                raise NotImplementedError(f"{name}.{d['name']}: {width} bits > 64")
            if d["name"] in m.sigs:                                      # 'output x; reg x;' style re-declaration
                s = m.sigs[d["name"]]
                s.kinds |= d["kinds"]
                if d["range"]:
                    s.width, s.lsb = width, lsb
                s.signed = s.signed or d["signed"]
                if d["init"] is not None:
                    s.init = d["init"]
                continue
            s = Sig(d["name"], width, d["signed"], d["kinds"], depth, base, None, lsb)
            m.sigs[d["name"]] = s
            if d["init"] is not None:
                if "reg" in d["kinds"] or "integer" in d["kinds"]:
                    s.init = d["init"]
                else:
                    m.assigns.append((("id", d["name"]), d["init"]))   # 'wire x = expr;'
        for p in src["ports"]:
            assert p in m.sigs, f"{name}: port {p} undeclared"

        for lv, rhs in src["assigns"]:
            m.assigns.append((lv, rhs))
        for _, clocked, body in src["always"]:
            body = self.unroll(m, body)
            if clocked:
                m.seq_blocks.append(body)
                for t in self.targets(body):
                    m.sig(t).seq = True
            else:
                m.comb_blocks.append(body)

        for inst in src["insts"]:
            child_over = {pn: m.const_num(pe) for pn, pe in inst["params"]}
            child = self.elaborate(inst["module"], child_over)
            conns = {}
            for pn, pe in inst["ports"]:
                assert pn in child.sigs, f"{name}.{inst['name']}: no port {pn} on {inst['module']}"
                conns[pn] = pe
            m.insts.append((inst["name"], child, conns))
        self.order.append(m)
        return m

    def unroll(self, m, st):
        k = st[0]
        if k == "block":
            out = []
            for s in st[1]:
                u = self.unroll(m, s)
                out += u[1] if u[0] == "block" else [u]
            return ("block", out)
        if k == "if":
            return ("if", st[1], self.unroll(m, st[2]), self.unroll(m, st[3]) if st[3] else None)
        if k == "case":
            return ("case", st[1], [(l, self.unroll(m, s)) for l, s in st[2]], self.unroll(m, st[3]) if st[3] else None)
        if k == "for":
            _, var, start, cond, step, body = st
            out, v, guard = [], m.const(start), 0
            while m.const(subst(cond, var, v)):
                out.append(self.unroll(m, subst(body, var, v)))
                v = m.const(subst(step, var, v))
                guard += 1
                assert guard < 100000
            return ("block", out)
        return st

    def targets(self, st):
        k = st[0]
        if k == "block":
            return [t for s in st[1] for t in self.targets(s)]
        if k == "if":
            return self.targets(st[2]) + (self.targets(st[3]) if st[3] else [])
        if k == "case":
            return [t for _, s in st[2] for t in self.targets(s)] + (self.targets(st[3]) if st[3] else [])
        lv = st[1]
        return [x[1] for x in lv[1]] if lv[0] == "concat" else [lv[1]]


# ---------------------------------------------------------------------------------------------------------------------
class Gen:
    """C++ text for one elaborated module."""

    def __init__(self, m):
        self.m = m
        self.local = {}        # signal name -> C++ lvalue override (temporaries inside a comb always block)

    def ref(self, name):
        return self.local.get(name, f"s_{name}")

    @staticmethod
    def lit(v):
        return f"{v}ULL" if v < 10 else f"0x{v:x}ULL"

    def ext(self, text, w, W, sign):
        if W < w:
            return f"({text} & {self.lit(mask(W))})"
        if W > w and sign:
            return f"(SX({text}, {w}) & {self.lit(mask(W))})" if W < 64 else f"((uint64_t)SX({text}, {w}))"
        return text

    def self_gen(self, e):
        w, s = self.m.selfsize(e)
        return self.gen(e, w, s), w, s

    def reads(self, e, acc):
        k = e[0]
        if k == "id":
            if e[1] not in self.m.params:
                acc.add(e[1])
        elif k in ("index", "range"):
            acc.add(e[1])
            for x in e[2:]:
                self.reads(x, acc)
        elif k in ("concat",):
            for x in e[1]:
                self.reads(x, acc)
        elif k == "repl":
            for x in e[2]:
                self.reads(x, acc)
        elif k == "un":
            self.reads(e[2], acc)
        elif k == "bin":
            self.reads(e[2], acc)
            self.reads(e[3], acc)
        elif k == "cond":
            for x in e[1:]:
                self.reads(x, acc)
        elif k == "call":
            for x in e[2]:
                self.reads(x, acc)
        return acc

    def gen(self, e, W, S):
        m, k = self.m, e[0]
        assert W <= 64, f"{m.src['name']}: expression wider than 64 bits"
        if k in ("num", "str") or (k == "id" and e[1] in m.params):
            return self.lit(m.interp(e, W, S))
        if k == "id":
            s = m.sig(e[1])
            assert not s.is_mem, f"memory {e[1]} used without index"
            return self.ext(self.ref(e[1]), s.width, W, S and s.signed)
        if k == "index":
            s = m.sig(e[1])
            idx, _, _ = self.self_gen(e[2])
            if s.is_mem:
                rd = f"RDMEM(m_{e[1]}, {idx}, {s.base}, {s.depth})"
                return self.ext(rd, s.width, W, S and s.signed)
            return f"BITSEL({self.ref(e[1])}, {idx}, {s.lsb}, {s.width})"
        if k == "range":
            s = m.sig(e[1])
            msb, lsb = m.const(e[2]), m.const(e[3])
            assert msb >= lsb and lsb >= s.lsb and msb < s.lsb + s.width, f"{m.src['name']}: bad part-select {e}"
            w = msb - lsb + 1
            sh = lsb - s.lsb
            base = self.ref(e[1]) if sh == 0 else f"({self.ref(e[1])} >> {sh})"
            return base if w == s.width else f"({base} & {self.lit(mask(w))})"
        if k in ("concat", "repl"):
            items = e[1] if k == "concat" else e[2]
            parts, total = [], 0
            for x in items:
                t, w, _ = self.self_gen(x)
                parts.append((t, w))
                total += w
            n = m.const(e[1]) if k == "repl" else 1
            assert total * n <= 64, f"{m.src['name']}: concatenation wider than 64 bits"
            terms, sh = [], total * n
            for _ in range(n):
                for t, w in parts:
                    sh -= w
                    terms.append(f"({t} << {sh})" if sh else t)
            return self.ext("(" + " | ".join(terms) + ")", total * n, W, False)
        if k == "un":
            op = e[1]
            if op in ("-", "+", "~"):
                a = self.gen(e[2], W, S)
                return a if op == "+" else f"(({'0 - ' if op == '-' else '~'}{a}) & {self.lit(mask(W))})"
            a, w, _ = self.self_gen(e[2])
            full = self.lit(mask(w))
            return {"!": f"(uint64_t)({a} == 0)", "|": f"(uint64_t)({a} != 0)", "&": f"(uint64_t)({a} == {full})",
                    "^": f"(uint64_t)__builtin_parityll({a})", "~|": f"(uint64_t)({a} == 0)",
                    "~&": f"(uint64_t)({a} != {full})", "~^": f"(uint64_t)(1 ^ __builtin_parityll({a}))"}[op]
        if k == "bin":
            op = e[1]
            if op in COMPARE:
                (wa, sa), (wb, sb) = m.selfsize(e[2]), m.selfsize(e[3])
                w, s = max(wa, wb), sa and sb
                a, b = self.gen(e[2], w, s), self.gen(e[3], w, s)
                cop = {"===": "==", "!==": "!="}.get(op, op)
                if s and cop not in ("==", "!="):
                    return f"(uint64_t)(SX({a}, {w}) {cop} SX({b}, {w}))"
                return f"(uint64_t)({a} {cop} {b})"
            if op in LOGICAL:
                a, _, _ = self.self_gen(e[2])
                b, _, _ = self.self_gen(e[3])
                return f"(uint64_t)(({a} != 0) {op} ({b} != 0))"
            if op in SHIFT:
                a = self.gen(e[2], W, S)
                b, _, _ = self.self_gen(e[3])
                if op in ("<<", "<<<"):
                    return f"(SHL({a}, {b}) & {self.lit(mask(W))})"
                if op == ">>" or (op == ">>>" and not S):
                    return f"SHR({a}, {b})"
                if op == ">>>":
                    return f"(ASHR(SX({a}, {W}), {b}) & {self.lit(mask(W))})"
                raise NotImplementedError("non-constant '**'")
            a, b = self.gen(e[2], W, S), self.gen(e[3], W, S)
            if op in ("/", "%"):
                fn = ("SDIV" if op == "/" else "SMOD") if S else ("UDIV" if op == "/" else "UMOD")
                args = f"SX({a}, {W}), SX({b}, {W})" if S else f"{a}, {b}"
                return f"({fn}({args}) & {self.lit(mask(W))})"
            if op in ("~^", "^~"):
                return f"(~({a} ^ {b}) & {self.lit(mask(W))})"
            if op in ("&", "|", "^"):
                return f"({a} {op} {b})"
            return f"(({a} {op} {b}) & {self.lit(mask(W))})"
        if k == "cond":
            c, _, _ = self.self_gen(e[1])
            return f"(({c}) != 0 ? {self.gen(e[2], W, S)} : {self.gen(e[3], W, S)})"
        if k == "call":
            if e[1] not in ("$signed", "$unsigned"):
                raise NotImplementedError(e[1])
            a, w, _ = self.self_gen(e[2][0])
            return self.ext(a, w, W, S and e[1] == "$signed")
        raise NotImplementedError(e)

    def rhs(self, e, lw):
        """RHS of an assignment to an lw-bit target: context width max(lw, self width), truncated to lw."""
        w, s = self.m.selfsize(e)
        W = max(w, lw)
        t = self.gen(e, W, s)
        return t if W == lw else f"({t} & {self.lit(mask(lw))})"

    def lv_width(self, lv):
        m = self.m
        if lv[0] == "id":
            return m.sig(lv[1]).width
        if lv[0] == "index":
            s = m.sig(lv[1])
            return s.width if s.is_mem else 1
        if lv[0] == "range":
            return m.const(lv[2]) - m.const(lv[3]) + 1
        if lv[0] == "concat":
            return sum(self.lv_width(x) for x in lv[1])
        raise NotImplementedError(lv)

    def store(self, lv, val, dest, out, ind):
        """Emit `lv = val` where dest(name) is the C++ variable that receives writes to signal `name`."""
        m, pad = self.m, "    " * ind
        if lv[0] == "id":
            out.append(f"{pad}{dest(lv[1])} = {val};")
        elif lv[0] == "range":
            s = m.sig(lv[1])
            msb, lsb = m.const(lv[2]), m.const(lv[3])
            mk = mask(msb - lsb + 1) << (lsb - s.lsb)
            out.append(f"{pad}{dest(lv[1])} = ({dest(lv[1])} & ~{self.lit(mk)}) | (({val}) << {lsb - s.lsb});")
        elif lv[0] == "index":
            s = m.sig(lv[1])
            assert not s.is_mem
            idx = m.const(lv[2]) - s.lsb
            out.append(f"{pad}{dest(lv[1])} = ({dest(lv[1])} & ~{self.lit(1 << idx)}) | (({val}) << {idx});")
        elif lv[0] == "concat":
            out.append(f"{pad}{{ uint64_t cv = {val};")
            sh = self.lv_width(lv)
            for x in lv[1]:
                w = self.lv_width(x)
                sh -= w
                self.store(x, f"((cv >> {sh}) & {self.lit(mask(w))})", dest, out, ind + 1)
            out.append(f"{pad}}}")
        else:
            raise NotImplementedError(lv)

    def stmt(self, st, out, ind, clocked, dest):
        m, pad = self.m, "    " * ind
        k = st[0]
        if k == "block":
            for s in st[1]:
                self.stmt(s, out, ind, clocked, dest)
        elif k == "if":
            c, _, _ = self.self_gen(st[1])
            out.append(f"{pad}if (({c}) != 0) {{")
            self.stmt(st[2], out, ind + 1, clocked, dest)
            if st[3]:
                out.append(f"{pad}}} else {{")
                self.stmt(st[3], out, ind + 1, clocked, dest)
            out.append(f"{pad}}}")
        elif k == "case":
            ws = [m.selfsize(st[1])] + [m.selfsize(l) for ls, _ in st[2] for l in ls]
            w, s = max(x[0] for x in ws), all(x[1] for x in ws)
            out.append(f"{pad}{{ uint64_t sel = {self.gen(st[1], w, s)};")
            first = True
            for labels, body in st[2]:
                cond = " || ".join(f"sel == {self.gen(l, w, s)}" for l in labels)
                out.append(f"{pad}{'if' if first else '} else if'} ({cond}) {{")
                first = False
                self.stmt(body, out, ind + 1, clocked, dest)
            if st[3]:
                out.append(f"{pad}}} else {{" if not first else f"{pad}{{")
                self.stmt(st[3], out, ind + 1, clocked, dest)
            out.append(f"{pad}}} }}")
        elif k in ("nba", "ba"):
            lv, e = st[1], st[2]
            if clocked and k != "nba":
                raise NotImplementedError(f"{m.src['name']}: blocking assignment in a clocked block")
            if not clocked and k != "ba":
                raise NotImplementedError(f"{m.src['name']}: non-blocking assignment in a combinational block")
            if lv[0] == "index" and m.sig(lv[1]).is_mem:
                s = m.sig(lv[1])
                assert clocked, "memory written from a combinational block"
                idx, _, _ = self.self_gen(lv[2])
                out.append(f"{pad}WRMEM(q_{lv[1]}, nq_{lv[1]}, {idx}, {self.rhs(e, s.width)});")
            else:
                self.store(lv, self.rhs(e, self.lv_width(lv)), dest, out, ind)
        else:
            raise NotImplementedError(st)

    def stmt_reads(self, st, acc):
        k = st[0]
        if k == "block":
            for s in st[1]:
                self.stmt_reads(s, acc)
        elif k == "if":
            self.reads(st[1], acc)
            self.stmt_reads(st[2], acc)
            if st[3]:
                self.stmt_reads(st[3], acc)
        elif k == "case":
            self.reads(st[1], acc)
            for ls, s in st[2]:
                for l in ls:
                    self.reads(l, acc)
                self.stmt_reads(s, acc)
            if st[3]:
                self.stmt_reads(st[3], acc)
        else:
            self.reads(st[2], acc)
            if st[1][0] in ("index", "range"):
                for x in st[1][2:]:
                    self.reads(x, acc)
        return acc


def topo(nodes):
    """nodes: list of (reads:set, writes:set, payload).  Returns (ordered payloads, acyclic?)."""
    writers = {}
    for i, (_, w, _) in enumerate(nodes):
        for s in w:
            writers.setdefault(s, []).append(i)
    deps = [set(j for s in r for j in writers.get(s, []) if j != i) for i, (r, _, _) in enumerate(nodes)]
    order, state, acyclic = [], [0] * len(nodes), True
    for root in range(len(nodes)):
        if state[root]:
            continue
        stack = [(root, iter(sorted(deps[root])))]
        state[root] = 1
        while stack:
            n, it = stack[-1]
            for j in it:
                if state[j] == 0:
                    state[j] = 1
                    stack.append((j, iter(sorted(deps[j]))))
                    break
                if state[j] == 1:
                    acyclic = False
            else:
                state[n] = 2
                order.append(n)
                stack.pop()
    return [nodes[i][2] for i in order], acyclic


def emit_module(design, m, out):
    g = Gen(m)
    src = m.src
    ports = src["ports"]
    L = out.append
    L(f"// ---- module {src['name']}" + (f"  params: " + ", ".join(f"{k}={v[1]}" for k, v in m.params.items())
                                        if m.cname != 'M_' + src['name'] else ""))
    L(f"struct {m.cname} {{")

    # gather seq-written signals before declaring
    nba_mem_counts = {}

    def count_mem_writes(st):
        k = st[0]
        if k == "block":
            for s in st[1]:
                count_mem_writes(s)
        elif k == "if":
            count_mem_writes(st[2])
            if st[3]:
                count_mem_writes(st[3])
        elif k == "case":
            for _, s in st[2]:
                count_mem_writes(s)
            if st[3]:
                count_mem_writes(st[3])
        elif st[1][0] == "index" and m.sig(st[1][1]).is_mem:
            nba_mem_counts[st[1][1]] = nba_mem_counts.get(st[1][1], 0) + 1

    for b in m.seq_blocks:
        count_mem_writes(b)

    # ---- code bodies first (they may create implicit nets)
    comb_nodes = []
    for lv, rhs in m.assigns:
        body = []
        tgt = lv[1] if lv[0] != "concat" else None
        assert tgt, "concat on the left of a continuous assign"
        if lv[0] == "id":
            s = m.sig(tgt)
            body.append(f"        SET(s_{tgt}, {g.rhs(rhs, s.width)});")
        else:
            tmp = []
            g.store(lv, g.rhs(rhs, g.lv_width(lv)), lambda n: "t", tmp, 2)
            body.append(f"        {{ uint64_t t = s_{tgt};")
            body += tmp
            body.append(f"        SET(s_{tgt}, t); }}")
        comb_nodes.append((g.reads(rhs, set()), {tgt}, body))
    for blk in m.comb_blocks:
        tg = sorted(set(design.targets(blk)))
        body = ["        {"]
        for t in tg:
            body.append(f"        uint64_t t_{t} = s_{t};")
            g.local[t] = f"t_{t}"
        g.stmt(blk, body, 2, False, lambda n: f"t_{n}")
        g.local = {}
        for t in tg:
            body.append(f"        SET(s_{t}, t_{t});")
        body.append("        }")
        reads = g.stmt_reads(blk, set()) - set(tg)
        comb_nodes.append((reads, set(tg), body))
    for iname, child, conns in m.insts:
        body, reads, writes = [], set(), set()
        outs = []
        for pn, pe in conns.items():
            ps = child.sigs[pn]
            if "input" in ps.kinds:
                val = g.rhs(pe, ps.width) if pe is not None else "0ULL"
                if pe is not None:
                    g.reads(pe, reads)
                body.append(f"        {{ uint64_t v_ = {val}; if ({iname}.s_{pn} != v_) {{ {iname}.s_{pn} = v_; {iname}.dirty = true; ch = true; }} }}")
            elif "output" in ps.kinds:
                if pe is None:
                    continue
                assert pe[0] in ("id", "range", "index"), f"output port {pn} connected to an expression"
                outs.append((pn, pe, ps))
                writes.add(pe[1])
            else:
                raise NotImplementedError(f"inout port {pn}")
        body.append(f"        ch |= {iname}.comb();")
        for pn, pe, ps in outs:
            lw = g.lv_width(pe)
            val = f"({iname}.s_{pn} & {g.lit(mask(lw))})" if lw < ps.width else f"{iname}.s_{pn}"
            if pe[0] == "id":
                body.append(f"        SET(s_{pe[1]}, {val});")
            else:
                tmp = []
                g.store(pe, val, lambda n: "t", tmp, 2)
                body.append(f"        {{ uint64_t t = s_{pe[1]};")
                body += tmp
                body.append(f"        SET(s_{pe[1]}, t); }}")
        comb_nodes.append((reads, writes, body))
    ordered, acyclic = topo(comb_nodes)

    seq_body = []
    for blk in m.seq_blocks:
        g.stmt(blk, seq_body, 2, True, lambda n: f"n_{n}")

    # ---- declarations
    for s in m.sigs.values():
        if s.is_mem:
            L(f"    uint64_t m_{s.name}[{s.depth}];")
            n = max(1, nba_mem_counts.get(s.name, 0))
            L(f"    MemWrite q_{s.name}[{n}]; int nq_{s.name} = 0;")
        else:
            L(f"    uint64_t s_{s.name} = 0;" + (f" uint64_t n_{s.name} = 0;" if s.seq else ""))
    for iname, child, _ in m.insts:
        L(f"    {child.cname} {iname};")
    L("    bool dirty = true;")

    # ---- constructor: declared initial values
    L(f"    {m.cname}() {{")
    for s in m.sigs.values():
        if s.is_mem:
            L(f"        for (int i = 0; i < {s.depth}; ++i) m_{s.name}[i] = 0;")
        elif s.init is not None:
            L(f"        s_{s.name} = {g.rhs(s.init, s.width)};")
    L("    }")

    # ---- scramble: junk into everything that has no declared initial value (X-independence checks)
    L("    void scramble(uint64_t &rng) {")
    for s in m.sigs.values():
        if s.is_mem:
            L(f"        for (int i = 0; i < {s.depth}; ++i) m_{s.name}[i] = NEXTRAND(rng) & {g.lit(mask(s.width))};")
        elif s.seq and s.init is None:
            L(f"        s_{s.name} = NEXTRAND(rng) & {g.lit(mask(s.width))};")
    for iname, _, _ in m.insts:
        L(f"        {iname}.scramble(rng);")
    L("        dirty = true;")
    L("    }")

    # ---- comb
    L("    bool comb() {")
    L("        if (!dirty) return false;")
    L("        bool ch = false;")
    for body in ordered:
        for line in body:
            L(line)
    L(f"        dirty = {'false' if acyclic and not m.insts else 'ch'};")
    L("        return ch;")
    L("    }")

    # ---- seq / commit
    L("    void seq() {")
    for s in m.sigs.values():
        if s.seq and not s.is_mem:
            L(f"        n_{s.name} = s_{s.name};")
    for line in seq_body:
        L(line)
    for iname, _, _ in m.insts:
        L(f"        {iname}.seq();")
    L("    }")
    L("    void commit() {")
    for s in m.sigs.values():
        if s.is_mem:
            if s.name in nba_mem_counts:
                L(f"        for (int i = 0; i < nq_{s.name}; ++i) if (q_{s.name}[i].a >= {s.base} && q_{s.name}[i].a < {s.base + s.depth}) "
                  f"m_{s.name}[q_{s.name}[i].a - {s.base}] = q_{s.name}[i].v;")
                L(f"        nq_{s.name} = 0;")
        elif s.seq:
            L(f"        s_{s.name} = n_{s.name};")
    for iname, _, _ in m.insts:
        L(f"        {iname}.commit();")
    L("        dirty = true;")
    L("    }")
    L("};")
    L("")
    return acyclic


PRELUDE = r"""// GENERATED by oracle/rtl2c/v2c.py from the reference RTL -- do not edit, do not commit (derivative of /root/reference).
#pragma once
#include <cstdint>
#include <cstdlib>
namespace rtl {
struct MemWrite { uint64_t a, v; };
static inline int64_t SX(uint64_t v, int w) { return w >= 64 ? (int64_t)v : ((int64_t)(v << (64 - w))) >> (64 - w); }
static inline uint64_t SHL(uint64_t a, uint64_t n) { return n >= 64 ? 0 : a << n; }
static inline uint64_t SHR(uint64_t a, uint64_t n) { return n >= 64 ? 0 : a >> n; }
static inline uint64_t ASHR(int64_t a, uint64_t n) { return (uint64_t)(a >> (n >= 63 ? 63 : n)); }
static inline uint64_t UDIV(uint64_t a, uint64_t b) { return b ? a / b : 0; }
static inline uint64_t UMOD(uint64_t a, uint64_t b) { return b ? a % b : 0; }
static inline uint64_t SDIV(int64_t a, int64_t b) { return b ? (uint64_t)(a / b) : 0; }
static inline uint64_t SMOD(int64_t a, int64_t b) { return b ? (uint64_t)(a % b) : 0; }
static inline uint64_t BITSEL(uint64_t v, uint64_t i, int lsb, int w) { return (i >= (uint64_t)lsb && i < (uint64_t)(lsb + w)) ? (v >> (i - lsb)) & 1 : 0; }
static inline uint64_t NEXTRAND(uint64_t &s) { s += 0x9E3779B97F4A7C15ULL; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
#define RDMEM(mem, idx, base, depth) (((idx) >= (uint64_t)(base) && (idx) < (uint64_t)((base) + (depth))) ? mem[(idx) - (base)] : 0ULL)
#define WRMEM(q, nq, idx, val) do { q[nq].a = (idx); q[nq].v = (val); ++nq; } while (0)
#define SET(sig, val) do { uint64_t v_ = (val); if (sig != v_) { sig = v_; ch = true; } } while (0)
template <class T> static inline int settle(T &m) { int n = 0; m.dirty = true; while (m.comb()) { if (++n > 10000) abort(); } return n; }
template <class T> static inline void clock(T &m) { settle(m); m.seq(); m.commit(); }
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", action="append", required=True)
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("files", nargs="+")
    a = ap.parse_args()
    design = Design(parse_files(a.files))
    for t in a.top:
        design.elaborate(t)
    out = [PRELUDE]
    stats = []
    for m in design.order:
        acyclic = emit_module(design, m, out)
        stats.append(f"//   {m.cname}: {len(m.sigs)} signals, {len(m.assigns)} assigns, {len(m.comb_blocks)} comb blocks, "
                     f"{len(m.seq_blocks)} clocked blocks, {len(m.insts)} instances, comb order "
                     f"{'acyclic' if acyclic else 'has feedback through instances (fixpoint)'}"
                     + (f", implicit nets: {', '.join(m.implicit)}" if m.implicit else ""))
    out.append("}  // namespace rtl")
    out.append("// elaboration summary:")
    out += stats
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"v2c: {len(design.order)} elaborated modules -> {a.out}")


if __name__ == "__main__":
    main()
