"""Verilog-2001 subset parser for the mounted RTL (test infrastructure, part of the oracle tooling).

The reference (/root/reference/*.v) is Vivado-HLS output plus a few hand-written template modules.  This parser
covers exactly the constructs those files use: module headers (ANSI and non-ANSI, `#(parameter ...)`), parameter /
input / output / wire / reg / integer declarations (ranges, `signed`, initial values, one memory dimension),
continuous assigns, `always @(posedge clk)` / `always @(a or b ...)` / `always @(*)` with begin/end, if/else, case,
for, blocking and non-blocking assignments, and module instances with named parameter / port connections.
Nothing here is specific to the Smith-Waterman design: the translation stays mechanical (see v2c.py).

AST nodes are plain tuples:  ('num', value, width|None, signed) ('id', name) ('str', text)
  ('index', name, expr) ('range', name, msb_expr, lsb_expr) ('concat', [exprs]) ('repl', count_expr, [exprs])
  ('un', op, a) ('bin', op, a, b) ('cond', c, a, b) ('call', name, [args])
"""
import re

KEYWORDS = {
    "module", "endmodule", "parameter", "localparam", "input", "output", "inout", "wire", "reg", "integer", "signed",
    "assign", "always", "begin", "end", "if", "else", "case", "endcase", "default", "for", "posedge", "negedge", "or",
}

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<num>(?:\d+)?\s*'[sS]?[bBdDhHoO]\s*[0-9a-fA-FxXzZ_?]+)
  | (?P<dec>\d[\d_]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_$]*)
  | (?P<sys>\$[A-Za-z_][A-Za-z0-9_$]*)
  | (?P<str>"[^"]*")
  | (?P<op>===|!==|<<<|>>>|==|!=|<=|>=|&&|\|\||<<|>>|\*\*|~\^|\^~|~&|~\||@\*|[-+*/%&|^~!<>=?:;,.()\[\]{}#@])
""", re.X)


def preprocess(text):
    """Strip comments, attributes and the handful of compiler directives the reference uses."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"@\s*\(\s*\*\s*\)", "@*", text)          # keep '@(*)' apart from '(* attribute *)'
    text = re.sub(r"\(\*.*?\*\)", " ", text, flags=re.S)
    text = re.sub(r"`(timescale|include|define|ifdef|ifndef|endif|else|elsif)[^\n]*", " ", text)
    text = re.sub(r"`[A-Za-z_][A-Za-z0-9_]*", '"macro"', text)  # `GRAM_AUTO etc.: RAM-style strings, no semantics
    return text


def tokenize(text):
    toks, pos = [], 0
    while pos < len(text):
        m = TOKEN_RE.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {text[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        val = m.group(kind)
        if kind == "id" and val in KEYWORDS:
            kind = "kw"
        toks.append((kind, val))
    toks.append(("eof", ""))
    return toks


def parse_number(text):
    """Sized / based literal -> ('num', value, width or None, signed).  x/z digits read as 0 (two-state model)."""
    m = re.match(r"(\d+)?\s*'([sS])?([bBdDhHoO])\s*([0-9a-fA-FxXzZ_?]+)$", text)
    width = int(m.group(1)) if m.group(1) else None
    signed = bool(m.group(2))
    base = {"b": 2, "d": 10, "h": 16, "o": 8}[m.group(3).lower()]
    digits = re.sub(r"[xXzZ?]", "0", m.group(4).replace("_", ""))
    value = int(digits, base)
    if width is not None:
        value &= (1 << width) - 1
    return ("num", value, width, signed)


BINARY_PREC = [            # lowest to highest
    ["||"], ["&&"], ["|", "~|"], ["^", "~^", "^~"], ["&", "~&"], ["==", "!=", "===", "!=="],
    ["<", "<=", ">", ">="], ["<<", ">>", "<<<", ">>>"], ["+", "-"], ["*", "/", "%"], ["**"],
]


class Parser:
    def __init__(self, text):
        self.toks = tokenize(preprocess(text))
        self.i = 0

    # ---- token helpers
    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def at(self, val):
        return self.peek()[1] == val and self.peek()[0] in ("kw", "op")

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            ctx = " ".join(t[1] for t in self.toks[max(0, self.i - 8):self.i + 4])
            raise SyntaxError(f"expected {val!r}, got {self.peek()!r} near: {ctx}")

    def ident(self):
        k, v = self.next()
        if k != "id":
            raise SyntaxError(f"expected identifier, got {(k, v)!r}")
        return v

    # ---- expressions
    def expr(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("cond", c, a, b)
        return c

    def binary(self, level):
        if level == len(BINARY_PREC):
            return self.unary()
        a = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in BINARY_PREC[level]:
            op = self.next()[1]
            b = self.binary(level + 1)
            a = ("bin", op, a, b)
        return a

    def unary(self):
        k, v = self.peek()
        if k == "op" and v in ("-", "+", "~", "!", "&", "|", "^", "~&", "~|", "~^", "^~"):
            self.next()
            return ("un", v, self.unary())
        return self.primary()

    def primary(self):
        k, v = self.next()
        if k == "num":
            return parse_number(v)
        if k == "dec":
            return ("num", int(v.replace("_", "")), None, True)      # plain decimal: signed, 32 bit
        if k == "str":
            return ("str", v[1:-1])
        if k == "sys":
            self.expect("(")
            args = [self.expr()]
            while self.accept(","):
                args.append(self.expr())
            self.expect(")")
            return ("call", v, args)
        if k == "id":
            if self.accept("["):
                a = self.expr()
                if self.accept(":"):
                    b = self.expr()
                    self.expect("]")
                    return ("range", v, a, b)
                self.expect("]")
                return ("index", v, a)
            return ("id", v)
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return e
        if k == "op" and v == "{":
            first = self.expr()
            if self.at("{"):                                          # replication {n{a,b}}
                self.next()
                items = [self.expr()]
                while self.accept(","):
                    items.append(self.expr())
                self.expect("}")
                self.expect("}")
                return ("repl", first, items)
            items = [first]
            while self.accept(","):
                items.append(self.expr())
            self.expect("}")
            return ("concat", items)
        raise SyntaxError(f"unexpected token {(k, v)!r} in expression")

    # ---- declarations
    def opt_range(self):
        if self.accept("["):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            self.expect("]")
            return (a, b)
        return None

    def decl_names(self, mod, kinds, signed, rng):
        """name [= init] [mem range] {, name ...}  -- after the type keywords and packed range."""
        while True:
            name = self.ident()
            mem = self.opt_range()
            init = self.expr() if self.accept("=") else None
            mod["decls"].append(dict(name=name, kinds=set(kinds), signed=signed, range=rng, mem=mem, init=init))
            # another name follows only if the next token after ',' is an identifier (ANSI port lists restart with a keyword)
            if self.at(",") and self.peek(1)[0] == "id":
                self.next()
                continue
            break

    def declaration(self, mod):
        kinds = []
        while self.peek()[0] == "kw" and self.peek()[1] in ("input", "output", "inout", "wire", "reg", "integer"):
            kinds.append(self.next()[1])
        signed = self.accept("signed")
        rng = self.opt_range()
        if "integer" in kinds:
            rng = (("num", 31, None, True), ("num", 0, None, True))
            signed = True
        self.decl_names(mod, kinds, signed, rng)

    def parameter_decl(self, mod, in_header):
        self.next()                                                   # parameter / localparam
        self.accept("signed")
        self.opt_range()
        while True:
            name = self.ident()
            self.expect("=")
            mod["params"].append((name, self.expr()))
            if in_header:
                # '#(parameter A = 1, B = 2)' or '#(parameter A = 1, parameter B = 2)'
                if self.at(",") and self.peek(1)[0] == "id":
                    self.next()
                    continue
                break
            if self.accept(","):
                continue
            break

    # ---- statements
    def lvalue(self):
        if self.at("{"):
            self.next()
            items = [self.lvalue()]
            while self.accept(","):
                items.append(self.lvalue())
            self.expect("}")
            return ("concat", items)
        name = self.ident()
        if self.accept("["):
            a = self.expr()
            if self.accept(":"):
                b = self.expr()
                self.expect("]")
                return ("range", name, a, b)
            self.expect("]")
            return ("index", name, a)
        return ("id", name)

    def statement(self):
        if self.accept("begin"):
            if self.accept(":"):
                self.ident()
            body = []
            while not self.accept("end"):
                body.append(self.statement())
            return ("block", body)
        if self.accept("if"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            t = self.statement()
            e = self.statement() if self.accept("else") else None
            return ("if", c, t, e)
        if self.accept("case"):
            self.expect("(")
            sel = self.expr()
            self.expect(")")
            items, default = [], None
            while not self.accept("endcase"):
                if self.accept("default"):
                    self.accept(":")
                    default = self.statement()
                    continue
                labels = [self.expr()]
                while self.accept(","):
                    labels.append(self.expr())
                self.expect(":")
                items.append((labels, self.statement()))
            return ("case", sel, items, default)
        if self.accept("for"):
            self.expect("(")
            var = self.ident()
            self.expect("=")
            start = self.expr()
            self.expect(";")
            cond = self.expr()
            self.expect(";")
            var2 = self.ident()
            self.expect("=")
            step = self.expr()
            self.expect(")")
            assert var == var2
            return ("for", var, start, cond, step, self.statement())
        if self.accept(";"):
            return ("block", [])
        lv = self.lvalue()
        if self.accept("<="):
            kind = "nba"
        else:
            self.expect("=")
            kind = "ba"
        rhs = self.expr()
        self.expect(";")
        return (kind, lv, rhs)

    def always(self):
        self.expect("always")
        clocked, sens = None, []
        if self.accept("@*"):
            pass
        else:
            self.expect("@")
            self.expect("(")
            while True:
                if self.accept("posedge"):
                    clocked = self.ident()
                elif self.accept("negedge"):
                    raise SyntaxError("negedge not supported")
                else:
                    sens.append(self.expr())
                if self.accept("or") or self.accept(","):
                    continue
                break
            self.expect(")")
        return ("always", clocked, self.statement())

    def instance(self, mod):
        mname = self.ident()
        params = []
        if self.accept("#"):
            self.expect("(")
            params = self.connections()
            self.expect(")")
        iname = self.ident()
        self.expect("(")
        ports = self.connections()
        self.expect(")")
        self.expect(";")
        mod["insts"].append(dict(module=mname, name=iname, params=params, ports=ports))

    def connections(self):
        out = []
        if self.at(")"):
            return out
        while True:
            self.expect(".")
            name = self.ident()
            self.expect("(")
            e = None if self.at(")") else self.expr()
            self.expect(")")
            out.append((name, e))
            if not self.accept(","):
                break
        return out

    # ---- module
    def module(self):
        self.expect("module")
        mod = dict(name=self.ident(), params=[], decls=[], assigns=[], always=[], insts=[], ports=[])
        if self.accept("#"):
            self.expect("(")
            while not self.at(")"):
                if self.at("parameter"):
                    self.parameter_decl(mod, in_header=True)
                else:                                                 # '#(parameter A=1, B=2': bare continuation
                    name = self.ident()
                    self.expect("=")
                    mod["params"].append((name, self.expr()))
                self.accept(",")
            self.expect(")")
        if self.accept("("):
            while not self.at(")"):
                if self.peek()[0] == "kw":                            # ANSI port declaration
                    n0 = len(mod["decls"])
                    self.declaration(mod)
                    mod["ports"] += [d["name"] for d in mod["decls"][n0:]]
                else:
                    mod["ports"].append(self.ident())
                self.accept(",")
            self.expect(")")
        self.expect(";")
        while not self.accept("endmodule"):
            k, v = self.peek()
            if k == "kw" and v in ("parameter", "localparam"):
                self.parameter_decl(mod, in_header=False)
                self.expect(";")
            elif k == "kw" and v in ("input", "output", "inout", "wire", "reg", "integer"):
                self.declaration(mod)
                self.expect(";")
            elif k == "kw" and v == "assign":
                self.next()
                while True:
                    lv = self.lvalue()
                    self.expect("=")
                    mod["assigns"].append((lv, self.expr()))
                    if not self.accept(","):
                        break
                self.expect(";")
            elif k == "kw" and v == "always":
                mod["always"].append(self.always())
            elif k == "id":
                self.instance(mod)
            else:
                raise SyntaxError(f"unexpected {(k, v)!r} in module {mod['name']}")
        return mod

    def source(self):
        mods = []
        while self.peek()[0] != "eof":
            mods.append(self.module())
        return mods


def parse_files(paths):
    mods = {}
    for p in paths:
        with open(p) as f:
            for m in Parser(f.read()).source():
                mods[m["name"]] = m
    return mods
