"""A small interpreter for the Fortran-90 subset the reference's kernels are written in.

TEST INFRASTRUCTURE (generator side only).  Purpose: pin the CPU oracle (oracle/sbd_oracle.cpp, a hand
restatement) against the REFERENCE'S OWN SOURCE TEXT.  The image has no Fortran compiler, so the reference
cannot be built; this module instead executes the unmodified files

    /root/reference/python_wrapper/seabreezediag/sobel.f90                  (get_edges, get_dist)
    /root/reference/python_wrapper/seabreezediag/seabreeze_diag_python.f90  (diag, sigmoid)
    /root/reference/generic/sea_breeze_diag.f90                             (sigmoid only: the rest has UB)

statement by statement with gfortran's default-kind semantics, and tests/golden/make_reference_golden.py
stores the inputs/outputs as fixtures (tests/golden/ref_*.npz) that tests/test_reference_golden.py
replays against the oracle on any machine.  Nothing here is a restatement of the algorithm: the
arithmetic that runs is whatever the .f90 text says.

Semantics implemented (what gfortran does for these sources on x86-64 without -ffast-math):
  * default REAL = IEEE binary32, default INTEGER = 32-bit, LOGICAL; every REAL operation is rounded to
    binary32 (numpy float32 scalar / array arithmetic: one IEEE operation per Fortran operation, no FMA);
  * mixed INTEGER/REAL operands: the integer is converted to REAL first; INTEGER division truncates;
  * x**2 with an integer exponent is repeated multiplication (x*x);
  * operator precedence and left-to-right association of the Fortran standard (a*b*c = (a*b)*c,
    -a*b = -(a*b), a**b**c = a**(b**c));
  * decimal literals are rounded ONCE to binary32 (exact rational comparison, not via binary64);
  * sqrt is the IEEE correctly rounded one; exp, sin, cos, atan2 call the platform libm's
    expf/sinf/cosf/atan2f through ctypes -- exactly what a gfortran build links against;
  * sum() accumulates sequentially in array element (column-major) order in binary32; maxval/minval/
    minloc(…,1) (first minimum); modulo for INTEGER (sign of the divisor) and REAL (fmodf + fix-up,
    libgfortran's definition); min/max/abs/int/real/size/shape/reshape and (/ … /) constructors;
  * arrays are 1-based, column-major, bounds-CHECKED (an out-of-bounds subscript raises, as
    -fbounds-check -- the reference's own build flag, python_wrapper/setup.py:13 -- would);
  * reading a variable that was never assigned raises (this is how the generic variant's
    uninitialised `found`, generic/sea_breeze_diag.f90:139,191, shows up);
  * `!$omp` directives and `!$ ` conditional lines are comments (single-thread order is the definition:
    the OpenMP regions of get_dist and sigmoid are racy); `!f2py` lines are comments.
Not implemented: everything these three files do not use (I/O, derived types, pointers, modules' data …).
"""
import ctypes
import ctypes.util
import re
from decimal import Decimal
from fractions import Fraction

import numpy as np

F32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n, _k in (("expf", 1), ("sinf", 1), ("cosf", 1), ("atan2f", 2), ("fmodf", 2)):
    _f = getattr(_libm, _n)
    _f.restype = ctypes.c_float
    _f.argtypes = [ctypes.c_float] * _k


class FortranError(RuntimeError):
    pass


class _Undef:
    def __repr__(self):
        return "<undefined>"


UNDEF = _Undef()


def dec_to_f32(text):
    """Decimal literal -> nearest binary32 (ties to even), rounded once."""
    t = text.lower().replace("d", "e")
    x = Fraction(Decimal(t))
    f = F32(float(x))
    best = f
    for cand in (np.nextafter(f, F32(-np.inf)), np.nextafter(f, F32(np.inf))):
        if not np.isfinite(cand):
            continue
        dc, db = abs(Fraction(float(cand)) - x), abs(Fraction(float(best)) - x)
        if dc < db or (dc == db and (int(cand.view(np.uint32)) & 1) == 0 and (int(best.view(np.uint32)) & 1) == 1):
            best = cand
    return F32(best)


# ------------------------------------------------------------------------------------------------
# source -> logical lines
# ------------------------------------------------------------------------------------------------
def logical_lines(text):
    out, cur, cur_no = [], "", None
    for no, raw in enumerate(text.splitlines(), 1):
        line = raw
        # comments: '!' outside of character strings (the kernels only have strings in parameter decls)
        res, q = "", None
        for ch in line:
            if q:
                res += ch
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                res += ch
            elif ch == "!":
                break
            else:
                res += ch
        line = res.strip()
        if not line:
            continue
        if cur:
            if line.startswith("&"):
                line = line[1:].lstrip()
            cur += " " + line
        else:
            cur, cur_no = line, no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        out.append((cur_no, cur.lower()))
        cur = ""
    if cur:
        out.append((cur_no, cur.lower()))
    return out


# ------------------------------------------------------------------------------------------------
# tokens / expressions
# ------------------------------------------------------------------------------------------------
_TOK = re.compile(r"""\s*(?:
    (?P<real>(?:\d+\.\d*|\.\d+)(?:[ed][+-]?\d+)?|\d+[ed][+-]?\d+)|
    (?P<int>\d+)|
    (?P<dotop>\.(?:and|or|not|eqv|neqv|true|false|lt|le|gt|ge|eq|ne)\.)|
    (?P<name>[a-z_][a-z0-9_]*)|
    (?P<op>\(/|/\)|\*\*|==|/=|<=|>=|::|[-+*/(),:<>=%])
)""", re.X)


def tokenize(s):
    pos, toks = 0, []
    while pos < len(s):
        if s[pos:].strip() == "":
            break
        m = _TOK.match(s, pos)
        if not m:
            raise FortranError("cannot tokenize %r at %r" % (s, s[pos:pos + 20]))
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    # a real literal directly followed by a dot-operator: "1.and." does not occur in these sources
    return toks


_REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "/="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise FortranError("expected %r, got %r in %r" % (val, self.peek(), self.t))

    # precedence climbing, lowest first
    def expr(self):
        return self.eqv()

    def eqv(self):
        a = self.or_()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.next()[1]
            a = ("eqv" if op == ".eqv." else "neqv", a, self.or_())
        return a

    def or_(self):
        a = self.and_()
        while self.accept(".or."):
            a = ("or", a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.accept(".and."):
            a = ("and", a, self.not_())
        return a

    def not_(self):
        if self.accept(".not."):
            return ("not", self.not_())
        return self.rel()

    def rel(self):
        a = self.add()
        v = self.peek()[1]
        v = _REL.get(v, v)
        if v in ("<", "<=", ">", ">=", "==", "/="):
            self.next()
            return ("rel", v, a, self.add())
        return a

    def add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = self.mul()
            a = ("neg", a) if op == "-" else a
        else:
            a = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = ("bin", op, a, self.mul())
        return a

    def mul(self):
        a = self.pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = ("bin", op, a, self.pow())
        return a

    def pow(self):
        a = self.primary()
        if self.accept("**"):
            # right associative; a unary minus may follow ** only in parentheses in standard Fortran
            return ("bin", "**", a, self.pow())
        return a

    def primary(self):
        k, v = self.next()
        if k == "int":
            return ("ilit", int(v))
        if k == "real":
            return ("rlit", dec_to_f32(v))
        if k == "dotop" and v in (".true.", ".false."):
            return ("llit", v == ".true.")
        if v == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if v == "(/":
            items = [self.expr()]
            while self.accept(","):
                items.append(self.expr())
            self.expect("/)")
            return ("acons", items)
        if k == "name":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    args.append(self.subscript())
                    while self.accept(","):
                        args.append(self.subscript())
                    self.expect(")")
                return ("call", v, args)
            return ("var", v)
        raise FortranError("unexpected token %r in %r" % ((k, v), self.t))

    def subscript(self):
        # expr | [expr] : [expr]
        if self.peek()[1] == ":":
            self.next()
            hi = None if self.peek()[1] in (",", ")") else self.expr()
            return ("slice", None, hi)
        e = self.expr()
        if self.accept(":"):
            hi = None if self.peek()[1] in (",", ")") else self.expr()
            return ("slice", e, hi)
        return e


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] is not None:
        raise FortranError("trailing tokens in expression %r" % s)
    return e


# ------------------------------------------------------------------------------------------------
# values and arithmetic
# ------------------------------------------------------------------------------------------------
def _is_int(v):
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_))


def _is_intarr(v):
    return isinstance(v, np.ndarray) and v.dtype.kind == "i"


def _to_real(v):
    if isinstance(v, np.ndarray):
        return v if v.dtype == np.float32 else v.astype(np.float32)
    if isinstance(v, np.float32):
        return v
    if _is_int(v):
        return F32(int(v))
    raise FortranError("cannot convert %r to REAL" % (v,))


def _integral(v):
    return _is_int(v) or _is_intarr(v)


def arith(op, a, b):
    if a is UNDEF or b is UNDEF:
        raise FortranError("arithmetic on an undefined value")
    if _integral(a) and _integral(b):
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                return (np.trunc(np.asarray(a) / np.asarray(b))).astype(np.int64)
            q = abs(int(a)) // abs(int(b))
            return q if (a >= 0) == (b >= 0) else -q
        if op == "**":
            return int(a) ** int(b)
    if op == "**":
        if _is_int(b):
            x = _to_real(a)
            n = int(b)
            if n < 0:
                raise FortranError("negative integer power not supported")
            if n == 0:
                return F32(1.0) if not isinstance(x, np.ndarray) else np.ones_like(x)
            r = x
            for _ in range(n - 1):   # gfortran expands small integer powers into multiplications (x**2 = x*x)
                r = r * x
            return _r(r)
        raise FortranError("REAL ** REAL is not used by the reference kernels")
    x, y = _to_real(a), _to_real(b)
    if op == "+":
        return _r(x + y)
    if op == "-":
        return _r(x - y)
    if op == "*":
        return _r(x * y)
    if op == "/":
        with np.errstate(divide="ignore", invalid="ignore"):
            return _r(x / y)
    raise FortranError("operator " + op)


def _r(v):
    if isinstance(v, np.ndarray):
        assert v.dtype == np.float32
        return v
    return F32(v)


def relop(op, a, b):
    if a is UNDEF or b is UNDEF:
        raise FortranError("comparison with an undefined value")
    if not (_integral(a) and _integral(b)):
        a, b = _to_real(a), _to_real(b)
    if op == "<":
        return a < b
    if op == "<=":
        return a <= b
    if op == ">":
        return a > b
    if op == ">=":
        return a >= b
    if op == "==":
        return a == b
    return a != b


def _libm1(name):
    fn = getattr(_libm, name)

    def f(x):
        x = _to_real(x)
        if isinstance(x, np.ndarray):
            flat = x.ravel(order="K")
            out = np.array([fn(float(t)) for t in flat], dtype=np.float32).reshape(x.shape, order="F" if x.flags.f_contiguous else "C")
            return out
        return F32(fn(float(x)))
    return f


def _seq_sum(a):
    a = _to_real(np.asarray(a))
    flat = a.ravel(order="F")
    if flat.size == 0:
        return F32(0.0)
    return F32(np.add.accumulate(flat, dtype=np.float32)[-1])   # strictly sequential binary32 adds


def _modulo(a, p):
    if _is_int(a) and _is_int(p):
        return int(a) - int(np.floor(int(a) / int(p))) * int(p) if False else int(a) % int(p)  # sign of the divisor
    a, p = _to_real(a), _to_real(p)
    r = F32(_libm.fmodf(float(a), float(p)))
    if r != 0 and ((r < 0) != (p < 0)):
        r = F32(r + p)
    return r


def _minmax(fn):
    def f(*args):
        if all(_is_int(a) for a in args):
            return fn(int(a) for a in args)
        vals = [_to_real(a) for a in args]
        r = vals[0]
        for v in vals[1:]:
            r = fn(r, v) if not isinstance(r, np.ndarray) else (np.minimum(r, v) if fn is min else np.maximum(r, v))
        return _r(r)
    return f


def _minloc(a, dim=None):
    a = np.asarray(a)
    if a.ndim != 1:
        raise FortranError("minloc: only rank-1 arrays")
    return int(np.argmin(a)) + 1   # first minimum, 1-based


def _int(x):
    if _is_int(x):
        return int(x)
    return int(np.trunc(float(_to_real(x))))


def _real(x):
    return _to_real(x)


def _abs(x):
    if _is_int(x):
        return abs(int(x))
    if _is_intarr(x):
        return np.abs(x)
    return _r(np.abs(_to_real(x)))


def _sqrt(x):
    x = _to_real(x)
    with np.errstate(invalid="ignore"):
        return _r(np.sqrt(x))   # IEEE correctly rounded in binary32


def _atan2(y, x):
    return F32(_libm.atan2f(float(_to_real(y)), float(_to_real(x))))


INTRINSICS = {
    "sqrt": _sqrt, "exp": _libm1("expf"), "sin": _libm1("sinf"), "cos": _libm1("cosf"), "atan2": _atan2,
    "abs": _abs, "int": _int, "real": _real, "modulo": _modulo, "min": _minmax(min), "max": _minmax(max),
    "sum": _seq_sum, "maxval": lambda a: _r(np.max(_to_real(np.asarray(a)))),
    "minval": lambda a: _r(np.min(_to_real(np.asarray(a)))), "minloc": _minloc,
    "shape": lambda a: np.array(a.shape, dtype=np.int64),
    "size": lambda a, d=None: int(a.size) if d is None else int(a.shape[int(d) - 1]),
    "reshape": lambda src, shp: np.asarray(src).reshape(tuple(int(s) for s in shp), order="F").copy(order="F"),
}


# ------------------------------------------------------------------------------------------------
# program units
# ------------------------------------------------------------------------------------------------
class Var:
    __slots__ = ("name", "type", "dims", "param", "init")

    def __init__(self, name, typ, dims, param, init):
        self.name, self.type, self.dims, self.param, self.init = name, typ, dims, param, init


class Sub:
    def __init__(self, name, args):
        self.name, self.args = name, args
        self.vars = {}     # name -> Var (declaration order preserved)
        self.body = []     # nested statements


_DECL = re.compile(r"^(integer|real|logical|character)\b(.*?)::(.*)$")


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    i = 0
    while i < len(s):
        ch = s[i]
        if s.startswith("(/", i):
            depth += 1
            cur += "(/"
            i += 2
            continue
        if s.startswith("/)", i):
            depth -= 1
            cur += "/)"
            i += 2
            continue
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out


def parse_units(text):
    """-> {name: Sub}"""
    lines = logical_lines(text)
    subs, i = {}, 0

    def parse_block(i, enders):
        """statements until a line whose first word(s) are in enders; returns (stmts, index of the ender, ender)"""
        stmts = []
        while i < len(lines):
            no, ln = lines[i]
            key = re.sub(r"\s+", "", ln)
            for e in enders:
                if key == e or key.startswith(e) and e in ("elseif(", "else", "endsubroutine", "enddo", "endif"):
                    if e == "else" and key.startswith("elseif"):
                        continue
                    return stmts, i, e
            # --- block constructs
            m = re.match(r"^do\s+while\s*\((.*)\)$", ln)
            if m:
                body, j, _ = parse_block(i + 1, ("enddo",))
                stmts.append(("dowhile", no, parse_expr(m.group(1)), body))
                i = j + 1
                continue
            m = re.match(r"^do\s+([a-z_][a-z0-9_]*)\s*=\s*(.*)$", ln)
            if m:
                parts = _split_top(m.group(2))
                body, j, _ = parse_block(i + 1, ("enddo",))
                stmts.append(("do", no, m.group(1), [parse_expr(p) for p in parts], body))
                i = j + 1
                continue
            m = re.match(r"^if\s*\((.*)\)\s*then$", ln)
            if m:
                clauses = []
                cond = parse_expr(m.group(1))
                j = i + 1
                while True:
                    body, j, e = parse_block(j, ("elseif(", "else", "endif"))
                    clauses.append((cond, body))
                    if e == "endif":
                        break
                    if e == "else":
                        body, j, e2 = parse_block(j + 1, ("endif",))
                        clauses.append((None, body))
                        break
                    mm = re.match(r"^else\s*if\s*\((.*)\)\s*then$", lines[j][1])
                    cond = parse_expr(mm.group(1))
                    j += 1
                stmts.append(("if", no, clauses))
                i = j + 1
                continue
            m = re.match(r"^if\s*\(", ln)
            if m:
                # one-line logical IF: find the matching parenthesis
                depth, k = 0, ln.index("(")
                for k in range(ln.index("("), len(ln)):
                    if ln[k] == "(":
                        depth += 1
                    elif ln[k] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                cond = parse_expr(ln[ln.index("(") + 1:k])
                stmts.append(("if", no, [(cond, [simple_stmt(no, ln[k + 1:].strip())])]))
                i += 1
                continue
            stmts.append(simple_stmt(no, ln))
            i += 1
        raise FortranError("missing %s" % (enders,))

    def simple_stmt(no, ln):
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*\((.*)\)$", ln)
        if m:
            return ("call", no, m.group(1), [parse_expr(a) for a in _split_top(m.group(2))])
        if ln in ("exit", "cycle", "return"):
            return (ln, no)
        if re.match(r"^allocate\s*\(", ln):
            inner = ln[ln.index("(") + 1:ln.rindex(")")].strip()
            return ("allocate", no, parse_expr(inner))
        # assignment: split at the top-level '='
        depth = 0
        for k, ch in enumerate(ln):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and ln[k - 1] not in "<>=/" and ln[k + 1:k + 2] != "=":
                return ("assign", no, parse_expr(ln[:k]), parse_expr(ln[k + 1:]))
        raise FortranError("line %d: cannot parse statement %r" % (no, ln))

    while i < len(lines):
        no, ln = lines[i]
        m = re.match(r"^subroutine\s+([a-z_][a-z0-9_]*)\s*\((.*)\)$", ln)
        if not m:
            i += 1
            continue
        sub = Sub(m.group(1), [a.strip() for a in m.group(2).split(",")])
        i += 1
        # specification part
        while i < len(lines):
            no, ln = lines[i]
            if ln.startswith("use ") or ln.startswith("implicit ") or ln.startswith("character"):
                i += 1
                continue
            d = _DECL.match(ln)
            if not d:
                break
            typ, attrs, ents = d.group(1), d.group(2), d.group(3)
            param = "parameter" in attrs
            dm = re.search(r"dimension\s*\((.*)\)", attrs)
            common_dims = None
            if dm:
                # the dimension(...) attribute: take the balanced parenthesis content
                start = attrs.index("(", attrs.index("dimension"))
                depth = 0
                for k in range(start, len(attrs)):
                    if attrs[k] == "(":
                        depth += 1
                    elif attrs[k] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                common_dims = _split_top(attrs[start + 1:k])
            for ent in _split_top(ents):
                init = None
                if "=" in ent and not param:
                    name_part, init = ent.split("=", 1)
                elif "=" in ent:
                    name_part, init = ent.split("=", 1)
                else:
                    name_part = ent
                name_part = name_part.strip()
                dims = common_dims
                mm = re.match(r"^([a-z_][a-z0-9_]*)\s*\((.*)\)$", name_part)
                if mm:
                    name_part, dims = mm.group(1), _split_top(mm.group(2))
                sub.vars[name_part] = Var(name_part, typ, dims, param, parse_expr(init.strip()) if init else None)
            i += 1
        sub.body, i, _ = parse_block(i, ("endsubroutine",))
        i += 1
        subs[sub.name] = sub
    return subs


# ------------------------------------------------------------------------------------------------
# execution
# ------------------------------------------------------------------------------------------------
class _Exit(Exception):
    pass


class _Cycle(Exception):
    pass


class _Return(Exception):
    pass


class Machine:
    def __init__(self, *sources, externals=None):
        self.subs = {}
        for text in sources:
            self.subs.update(parse_units(text))
        self.externals = dict(externals or {})
        self.stmt_count = 0

    # -- public: call a subroutine with python values; arrays are numpy float32/int arrays in Fortran
    #    index order (shape == the Fortran shape), passed by reference; returns the final scalar values
    def call(self, name, *actuals):
        sub = self.subs[name]
        if len(actuals) != len(sub.args):
            raise FortranError("%s takes %d arguments, got %d" % (name, len(sub.args), len(actuals)))
        env = {}
        for a, v in zip(sub.args, actuals):
            env[a] = v
        self._enter(sub, env)
        try:
            self._run(sub, sub.body, env)
        except _Return:
            pass
        return {a: env[a] for a in sub.args}

    def _enter(self, sub, env):
        # dummy arguments: check type/shape; locals: allocate; parameters: evaluate (in declaration order)
        for name, var in sub.vars.items():
            if name in sub.args:
                v = env[name]
                if var.dims is not None:
                    if not isinstance(v, np.ndarray):
                        raise FortranError("%s: argument %s must be an array" % (sub.name, name))
                    want_dt = np.float32 if var.type == "real" else np.int64
                    if v.dtype != want_dt:
                        raise FortranError("%s: argument %s has dtype %s" % (sub.name, name, v.dtype))
                    if not all(d.strip() == ":" for d in var.dims):
                        # explicit-shape dummy: evaluated after the scalar arguments are known (below)
                        pass
                else:
                    if var.type == "real":
                        env[name] = _to_real(v)
                    elif var.type == "integer":
                        if not _is_int(v):
                            raise FortranError("%s: argument %s must be an integer" % (sub.name, name))
                        env[name] = int(v)
        for name, var in sub.vars.items():
            if name in sub.args:
                if var.dims is not None and not all(d.strip() == ":" for d in var.dims):
                    shp = tuple(int(self._eval(parse_expr(d), env, sub)) for d in var.dims)
                    if tuple(env[name].shape) != shp:
                        raise FortranError("%s: argument %s has shape %s, declared %s" % (sub.name, name, env[name].shape, shp))
                continue
            if var.param:
                val = self._eval(var.init, env, sub)
                env[name] = _to_real(val) if var.type == "real" else val
            elif var.dims is not None:
                if any(d.strip() == ":" for d in var.dims):
                    env[name] = UNDEF   # allocatable
                else:
                    shp = tuple(int(self._eval(parse_expr(d), env, sub)) for d in var.dims)
                    if var.type == "real":
                        env[name] = np.full(shp, np.nan, dtype=np.float32, order="F")   # "uninitialised"
                    else:
                        env[name] = np.full(shp, -2 ** 40, dtype=np.int64, order="F")
            else:
                env[name] = UNDEF
        for a in sub.args:
            if a not in sub.vars:
                raise FortranError("%s: dummy argument %s is not declared (implicit none)" % (sub.name, a))

    def _store_scalar(self, sub, env, name, val):
        var = sub.vars.get(name)
        if var is None:
            raise FortranError("%s: assignment to undeclared variable %s (implicit none)" % (sub.name, name))
        if var.param:
            raise FortranError("assignment to parameter " + name)
        if var.type == "real":
            env[name] = _to_real(val)
        elif var.type == "integer":
            env[name] = _int(val)
        elif var.type == "logical":
            env[name] = bool(val)
        else:
            raise FortranError("type " + var.type)

    def _index(self, arr, subs, env, sub, name):
        idx = []
        for d, s in enumerate(subs):
            if isinstance(s, tuple) and s[0] == "slice":
                lo = 1 if s[1] is None else int(self._eval(s[1], env, sub))
                hi = arr.shape[d] if s[2] is None else int(self._eval(s[2], env, sub))
                if lo < 1 or hi > arr.shape[d]:
                    raise FortranError("%s: section %d:%d of %s out of bounds (extent %d)" % (sub.name, lo, hi, name, arr.shape[d]))
                idx.append(slice(lo - 1, hi))
            else:
                k = self._eval(s, env, sub)
                if not _is_int(k):
                    raise FortranError("subscript of %s is not an integer" % name)
                k = int(k)
                if k < 1 or k > arr.shape[d]:
                    raise FortranError("%s: subscript %d of %s is out of bounds: %d not in 1..%d"
                                       % (sub.name, d + 1, name, k, arr.shape[d]))
                idx.append(k - 1)
        if len(idx) != arr.ndim:
            raise FortranError("rank mismatch indexing " + name)
        return tuple(idx)

    def _eval(self, e, env, sub):
        k = e[0]
        if k == "ilit" or k == "rlit" or k == "llit":
            return e[1]
        if k == "var":
            if e[1] not in env:
                raise FortranError("%s: undeclared variable %s" % (sub.name, e[1]))
            v = env[e[1]]
            if v is UNDEF:
                raise FortranError("%s: variable %s is used before it is defined" % (sub.name, e[1]))
            return v
        if k == "paren":
            return self._eval(e[1], env, sub)
        if k == "bin":
            return arith(e[1], self._eval(e[2], env, sub), self._eval(e[3], env, sub))
        if k == "neg":
            v = self._eval(e[1], env, sub)
            return -v if _integral(v) else _r(-_to_real(v))
        if k == "rel":
            return relop(e[1], self._eval(e[2], env, sub), self._eval(e[3], env, sub))
        if k == "and":
            return bool(self._eval(e[1], env, sub)) and bool(self._eval(e[2], env, sub))
        if k == "or":
            return bool(self._eval(e[1], env, sub)) or bool(self._eval(e[2], env, sub))
        if k == "not":
            return not bool(self._eval(e[1], env, sub))
        if k == "eqv":
            return bool(self._eval(e[1], env, sub)) == bool(self._eval(e[2], env, sub))
        if k == "neqv":
            return bool(self._eval(e[1], env, sub)) != bool(self._eval(e[2], env, sub))
        if k == "acons":
            vals = [self._eval(x, env, sub) for x in e[1]]
            if all(_is_int(v) for v in vals):
                return np.array(vals, dtype=np.int64)
            return np.array([_to_real(v) for v in vals], dtype=np.float32)
        if k == "call":
            name, args = e[1], e[2]
            if name in env and isinstance(env[name], np.ndarray):
                arr = env[name]
                v = arr[self._index(arr, args, env, sub, name)]
                if isinstance(v, np.ndarray):
                    return v
                if arr.dtype == np.float32:
                    if np.isnan(v) and sub.vars[name].dims is not None and name not in sub.args and not getattr(self, "allow_nan", False):
                        raise FortranError("%s: element of local array %s is used before it is defined" % (sub.name, name))
                    return F32(v)
                if v == -2 ** 40:
                    raise FortranError("%s: element of %s used before it is defined" % (sub.name, name))
                return int(v)
            if name in env and env[name] is UNDEF:
                raise FortranError("%s: array %s is used before it is allocated/defined" % (sub.name, name))
            if name in INTRINSICS:
                vals = [self._eval(a, env, sub) for a in args]
                return INTRINSICS[name](*vals)
            raise FortranError("%s: unknown function or array %s" % (sub.name, name))
        raise FortranError("cannot evaluate %r" % (e,))

    def _run(self, sub, stmts, env):
        for st in stmts:
            self.stmt_count += 1
            k = st[0]
            if k == "assign":
                lhs, rhs = st[2], st[3]
                val = self._eval(rhs, env, sub)
                if lhs[0] == "var":
                    name = lhs[1]
                    cur = env.get(name, None)
                    if isinstance(cur, np.ndarray):     # whole-array assignment
                        if isinstance(val, np.ndarray) and val.shape != cur.shape:
                            raise FortranError("%s: shape mismatch assigning to %s" % (sub.name, name))
                        cur[...] = val if cur.dtype != np.float32 else _to_real(val)
                    else:
                        self._store_scalar(sub, env, name, val)
                else:
                    name = lhs[1]
                    arr = env.get(name)
                    if not isinstance(arr, np.ndarray):
                        raise FortranError("%s: %s is not an array" % (sub.name, name))
                    idx = self._index(arr, lhs[2], env, sub, name)
                    if arr.dtype == np.float32:
                        arr[idx] = _to_real(val)
                    else:
                        arr[idx] = _int(val) if not isinstance(val, np.ndarray) else val
            elif k == "if":
                for cond, body in st[2]:
                    if cond is None or bool(self._eval(cond, env, sub)):
                        self._run(sub, body, env)
                        break
            elif k == "do":
                var = st[2]
                lims = [int(self._eval(x, env, sub)) for x in st[3]]
                lo, hi = lims[0], lims[1]
                step = lims[2] if len(lims) > 2 else 1
                n = max((hi - lo + step) // step, 0)
                v = lo
                try:
                    for _ in range(n):
                        self._store_scalar(sub, env, var, v)
                        try:
                            self._run(sub, st[4], env)
                        except _Cycle:
                            pass
                        v += step
                    self._store_scalar(sub, env, var, v)
                except _Exit:
                    pass
            elif k == "dowhile":
                try:
                    guard = 0
                    while bool(self._eval(st[2], env, sub)):
                        try:
                            self._run(sub, st[3], env)
                        except _Cycle:
                            pass
                        guard += 1
                        if guard > 100000:
                            raise FortranError("%s line %d: do while does not terminate" % (sub.name, st[1]))
                except _Exit:
                    pass
            elif k == "call":
                name = st[2]
                vals = [self._eval(a, env, sub) for a in st[3]]
                if name in self.subs:
                    res = self.call(name, *vals)
                    callee = self.subs[name]
                    for formal, actual in zip(callee.args, st[3]):   # scalar copy-out (by-reference semantics)
                        if actual[0] == "var" and not isinstance(env[actual[1]], np.ndarray):
                            if not sub.vars[actual[1]].param and res[formal] is not UNDEF:
                                self._store_scalar(sub, env, actual[1], res[formal])
                elif name in self.externals:
                    self.externals[name](*vals)
                else:
                    raise FortranError("%s: call to unknown subroutine %s" % (sub.name, name))
            elif k == "allocate":
                e = st[2]
                shp = tuple(int(self._eval(a, env, sub)) for a in e[2])
                var = sub.vars[e[1]]
                env[e[1]] = (np.full(shp, np.nan, dtype=np.float32, order="F") if var.type == "real"
                             else np.full(shp, -2 ** 40, dtype=np.int64, order="F"))
            elif k == "exit":
                raise _Exit()
            elif k == "cycle":
                raise _Cycle()
            elif k == "return":
                raise _Return()
            else:
                raise FortranError("statement kind " + k)
