"""Coefficient front-end: the handful of `fenics` names graphnics users pass as model data
(SURVEY.md 8b "Coefficient inputs"): Constant, Expression("c string", degree=, t=...),
UserExpression subclasses, and UFL-like expressions of SpatialCoordinate / Constant
(tests/test_flow_models.py:94-119, tests/test_time_stepping.py:17-60).

Everything symbolic is a sympy expression underneath; it is compiled to the RPN program of
include/graphnics_b200.h (gnx_program_t) and evaluated ON THE DEVICE by gnx_eval_expr / inside the
kernels that need values.  Python callbacks (UserExpression.eval / eval_cell) are the only thing
evaluated on the host -- exactly as slow as in the reference, and only if the user supplies one.
"""
import ctypes as C
import re

import numpy as np
import sympy as sp

from . import _lib

__all__ = ["Constant", "Expression", "UserExpression", "UFLExpr", "SpatialCoordinate", "as_coefficient",
           "sin", "cos", "tan", "exp", "ln", "sqrt", "pi", "diff", "compile_program", "CompiledCoefficient",
           "conditional", "lt", "gt", "le", "ge", "tanh", "dot", "grad", "inner"]

_X = sp.symbols("x0 x1 x2", real=True)
_TAU = sp.symbols("tau0 tau1 tau2", real=True)
_EDGE = sp.Symbol("edge_attr", real=True)
pi = sp.pi


# ----------------------------------------------------------------------------------------
# symbolic wrappers
# ----------------------------------------------------------------------------------------
def _sym(v):
    if isinstance(v, UFLExpr):
        return v.expr
    if isinstance(v, Constant):
        return v.symbol
    if isinstance(v, Expression):
        return v.as_ufl().expr
    if isinstance(v, (int, float, np.floating, np.integer)):
        return sp.Float(float(v)) if isinstance(v, (float, np.floating)) else sp.Integer(int(v))
    if isinstance(v, sp.Expr):
        return v
    raise TypeError(f"cannot use {type(v).__name__} in a coefficient expression")


def _params(*vs):
    out = {}
    for v in vs:
        if isinstance(v, UFLExpr):
            out.update(v.params)
        elif isinstance(v, Constant):
            out[v.symbol] = v
        elif isinstance(v, Expression):
            out.update(v.as_ufl().params)
    return out


class _Arith:
    def _bin(self, other, fn):
        return UFLExpr(fn(_sym(self), _sym(other)), _params(self, other))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._bin(o, lambda a, b: b ** a)
    def __neg__(self): return UFLExpr(-_sym(self), _params(self))
    def __pos__(self): return self


class UFLExpr(_Arith):
    """A scalar symbolic expression of x0,x1,x2, tangent components and Constants."""

    def __init__(self, expr, params=None, degree=2):
        self.expr = sp.sympify(expr)
        self.params = dict(params or {})     # sympy Symbol -> Constant
        self.degree = degree

    def __repr__(self):
        return f"UFLExpr({self.expr})"


_const_counter = [0]


class Constant(_Arith):
    """dolfin.Constant: a mutable scalar (`assign`) that expressions can depend on."""

    def __init__(self, value=0.0, name=None):
        if isinstance(value, (list, tuple, np.ndarray)):
            self._vec = np.asarray(value, dtype=np.float64)
            value = 0.0
        else:
            self._vec = None
        self._value = float(value)
        _const_counter[0] += 1
        self.symbol = sp.Symbol(name or f"c{_const_counter[0]}", real=True)
        self.degree = 0

    def assign(self, v):
        self._value = float(v._value if isinstance(v, Constant) else v)

    def values(self):
        return np.array([self._value]) if self._vec is None else self._vec

    def __float__(self):
        return self._value

    def __call__(self, *x):
        return self._value

    def __repr__(self):
        return f"Constant({self._value})"


class Expression(_Arith):
    """dolfin.Expression("c-syntax string", degree=d, **params).  Parameters are mutable
    attributes (`f.t = 0.3`, time_stepping.py:207-211)."""

    _FUNCS = {"sin": sp.sin, "cos": sp.cos, "tan": sp.tan, "exp": sp.exp, "log": sp.log, "sqrt": sp.sqrt,
              "pow": lambda a, b: a ** b, "fabs": sp.Abs, "abs": sp.Abs, "tanh": sp.tanh, "sinh": sp.sinh,
              "cosh": sp.cosh, "atan": sp.atan, "asin": sp.asin, "acos": sp.acos, "atan2": sp.atan2,
              "floor": sp.floor, "ceil": sp.ceiling, "fmod": lambda a, b: sp.Mod(a, b), "fmin": sp.Min,
              "fmax": sp.Max, "M_PI": sp.pi, "pi": sp.pi, "DOLFIN_PI": sp.pi, "erf": sp.erf}

    def __init__(self, cppcode, degree=None, element=None, **kwargs):
        if not isinstance(cppcode, str):
            raise TypeError("only scalar c-string Expressions are supported")
        object.__setattr__(self, "_pnames", list(kwargs))
        object.__setattr__(self, "_pconst", {k: Constant(float(v), name=f"p_{k}_{id(self)}") for k, v in kwargs.items()})
        self.cppcode = cppcode
        self.degree = 2 if degree is None else int(degree)
        self._ufl = None

    def __setattr__(self, k, v):
        if k in getattr(self, "_pnames", ()):
            self._pconst[k].assign(float(v))
        else:
            object.__setattr__(self, k, v)

    def __getattr__(self, k):
        if k in self.__dict__.get("_pnames", ()):
            return float(self.__dict__["_pconst"][k])
        raise AttributeError(k)

    def as_ufl(self):
        if self._ufl is None:
            code = re.sub(r"x\s*\[\s*(\d)\s*\]", r"x\1", self.cppcode)
            local = dict(self._FUNCS)
            local.update({f"x{i}": _X[i] for i in range(3)})
            local.update({k: c.symbol for k, c in self._pconst.items()})
            from sympy.parsing.sympy_parser import parse_expr
            expr = parse_expr(code, local_dict=local, evaluate=True)
            self._ufl = UFLExpr(expr, {c.symbol: c for c in self._pconst.values()}, self.degree)
        return self._ufl

    def __call__(self, *x):
        x = np.asarray(x[0] if len(x) == 1 and np.ndim(x[0]) else x, dtype=np.float64)
        return float(as_coefficient(self).eval_host(x.reshape(1, -1))[0])


class UserExpression:
    """dolfin.UserExpression: subclass and override eval(values, x) or eval_cell(values, x, cell).
    Evaluated on the host, point by point (as in the reference)."""

    def __init__(self, *args, degree=2, **kwargs):
        self.degree = degree

    def value_shape(self):
        return ()


class _Cell:
    def __init__(self, index):
        self.index = int(index)


def SpatialCoordinate(mesh):
    gd = mesh.geometry().dim() if hasattr(mesh, "geometry") else int(mesh)
    return tuple(UFLExpr(_X[i]) for i in range(gd))


def _fn(f):
    def g(v):
        if isinstance(v, (int, float)):
            return float(f(v))
        return UFLExpr(f(_sym(v)), _params(v))
    return g


sin, cos, tan, exp, ln, sqrt, tanh = (_fn(f) for f in (sp.sin, sp.cos, sp.tan, sp.exp, sp.log, sp.sqrt, sp.tanh))


def lt(a, b): return UFLExpr(sp.StrictLessThan(_sym(a), _sym(b)), _params(a, b))
def gt(a, b): return UFLExpr(sp.StrictGreaterThan(_sym(a), _sym(b)), _params(a, b))
def le(a, b): return UFLExpr(sp.LessThan(_sym(a), _sym(b)), _params(a, b))
def ge(a, b): return UFLExpr(sp.GreaterThan(_sym(a), _sym(b)), _params(a, b))


def conditional(cond, tv, fv):
    return UFLExpr(sp.Piecewise((_sym(tv), cond.expr), (_sym(fv), True)), _params(cond, tv, fv))


def diff(f, v):
    """ufl.diff(f, v): derivative w.r.t. a Constant (tests/test_time_stepping.py:41) or a coordinate."""
    return UFLExpr(sp.diff(_sym(f), _sym(v)), _params(f))


class _Grad:
    def __init__(self, f):
        self.f = f


def grad(f):
    return _Grad(f)


def dot(a, b):
    """dot(grad(f), tangent): the only use in graphnics (fenics_graph.py:259,266)."""
    if isinstance(a, _Grad):
        f = a.f
        comps = b if isinstance(b, (list, tuple, np.ndarray)) else [UFLExpr(_TAU[i]) for i in range(3)]
        e = 0
        for i, ti in enumerate(list(comps)[:3]):
            e = e + sp.diff(_sym(f), _X[i]) * _sym(ti)
        return UFLExpr(e, _params(f))
    return a * b


inner = dot


def tangent_components(gdim):
    return [UFLExpr(_TAU[i]) for i in range(gdim)]


# ----------------------------------------------------------------------------------------
# sympy -> RPN
# ----------------------------------------------------------------------------------------
_OPS = dict(CONST=0, X=1, PARAM=2, TAU=3, EDGE=4, ADD=10, SUB=11, MUL=12, DIV=13, POW=14, NEG=15, MIN=16, MAX=17,
            SIN=20, COS=21, TAN=22, EXP=23, LOG=24, SQRT=25, ABS=26, TANH=27, SINH=28, COSH=29, ATAN=30, ASIN=31,
            ACOS=32, FLOOR=33, FMOD=34, POWI=35, ATAN2=36, SIGN=37, LT=38, GT=39, LE=40, GE=41, SELECT=42, CEIL=43)
_UNARY = {sp.sin: "SIN", sp.cos: "COS", sp.tan: "TAN", sp.exp: "EXP", sp.log: "LOG", sp.Abs: "ABS",
          sp.tanh: "TANH", sp.sinh: "SINH", sp.cosh: "COSH", sp.atan: "ATAN", sp.asin: "ASIN", sp.acos: "ACOS",
          sp.floor: "FLOOR", sp.ceiling: "CEIL", sp.sign: "SIGN"}
_REL = {sp.StrictLessThan: "LT", sp.StrictGreaterThan: "GT", sp.LessThan: "LE", sp.GreaterThan: "GE"}
PROG_MAX, PROG_PARAMS = 96, 8


class gnx_program_t(C.Structure):
    _fields_ = [("len", C.c_int32), ("reserved", C.c_int32), ("op", C.c_int16 * PROG_MAX),
                ("arg", C.c_int16 * PROG_MAX), ("val", C.c_double * PROG_MAX), ("params", C.c_double * PROG_PARAMS)]


def _emit(e, code, pslots):
    def push(op, arg=0, val=0.0):
        code.append((_OPS[op], int(arg), float(val)))

    if e.is_Number or e in (sp.pi, sp.E) or (e.is_number and not e.free_symbols):
        push("CONST", 0, float(e))
    elif e.is_Symbol:
        if e in _X:
            push("X", _X.index(e))
        elif e in _TAU:
            push("TAU", _TAU.index(e))
        elif e == _EDGE:
            push("EDGE")
        else:
            if e not in pslots:
                if len(pslots) >= PROG_PARAMS:
                    raise ValueError("too many run-time parameters in one coefficient")
                pslots[e] = len(pslots)
            push("PARAM", pslots[e])
    elif e.is_Add:
        args = e.as_ordered_terms()
        _emit(args[0], code, pslots)
        for a in args[1:]:
            _emit(a, code, pslots); push("ADD")
    elif e.is_Mul:
        args = e.as_ordered_factors()
        _emit(args[0], code, pslots)
        for a in args[1:]:
            _emit(a, code, pslots); push("MUL")
    elif e.is_Pow:
        base, ex = e.as_base_exp()
        _emit(base, code, pslots)
        if ex.is_Integer and abs(int(ex)) < 30000:
            push("POWI", int(ex))
        elif ex == sp.Rational(1, 2):
            push("SQRT")
        else:
            _emit(ex, code, pslots); push("POW")
    elif e.func in _UNARY:
        _emit(e.args[0], code, pslots); push(_UNARY[e.func])
    elif e.func in _REL:
        _emit(e.args[0], code, pslots); _emit(e.args[1], code, pslots); push(_REL[e.func])
    elif e.func in (sp.Min, sp.Max):
        _emit(e.args[0], code, pslots)
        for a in e.args[1:]:
            _emit(a, code, pslots); push("MIN" if e.func is sp.Min else "MAX")
    elif e.func is sp.Mod:
        _emit(e.args[0], code, pslots); _emit(e.args[1], code, pslots); push("FMOD")
    elif e.func is sp.atan2:
        _emit(e.args[0], code, pslots); _emit(e.args[1], code, pslots); push("ATAN2")
    elif isinstance(e, sp.Piecewise):
        (v0, c0), rest = e.args[0], e.args[1:]
        if c0 == True:  # noqa: E712
            _emit(v0, code, pslots)
        else:
            _emit(c0, code, pslots); _emit(v0, code, pslots)
            _emit(sp.Piecewise(*rest) if rest else sp.Integer(0), code, pslots)
            push("SELECT")
    else:
        raise NotImplementedError(f"cannot compile {e.func} to the device expression language")


def compile_program(expr):
    """sympy expr -> (code list, {param symbol: slot})"""
    code, pslots = [], {}
    _emit(sp.sympify(expr), code, pslots)
    if len(code) > PROG_MAX:
        # common sub-expression heavy: try a simplified form once
        code, pslots = [], {}
        _emit(sp.simplify(expr), code, pslots)
        if len(code) > PROG_MAX:
            raise ValueError(f"coefficient expression too long for the device interpreter ({len(code)} > {PROG_MAX} ops)")
    return code, pslots


class CompiledCoefficient:
    """A coefficient ready for the device: RPN program + live parameter bindings."""

    def __init__(self, ufl: UFLExpr = None, constant=None, user=None, degree=2):
        self.ufl, self.constant, self.user = ufl, constant, user
        self.degree = degree
        self._code = None
        if ufl is not None:
            self._code, self._pslots = compile_program(ufl.expr)
            self.uses_tangent = any(op == _OPS["TAU"] for op, _, _ in self._code)

    @property
    def is_constant(self):
        return self.constant is not None

    def const_value(self):
        return float(self.constant)

    def program(self):
        """gnx_program_t with the CURRENT parameter values."""
        p = gnx_program_t()
        p.len = len(self._code)
        for i, (op, arg, val) in enumerate(self._code):
            p.op[i], p.arg[i], p.val[i] = op, arg, val
        for sym, slot in self._pslots.items():
            p.params[slot] = float(self.ufl.params[sym])
        return p

    # -- evaluation -------------------------------------------------------------------------
    def eval_device(self, x, pts_per_edge=0, tangents=None, edge_attr=None, cell_index=None, out=None):
        """x: [npts, gdim] float64 CUDA tensor -> [npts] CUDA tensor (`out` is reused when given)."""
        import torch
        from .device import current_stream
        npts, gd = int(x.shape[0]), int(x.shape[1])
        if self.is_constant:
            if out is not None:
                return out.fill_(self.const_value())
            return torch.full((npts,), self.const_value(), dtype=torch.float64, device=x.device)
        if self.user is not None:
            vals = self.eval_host(x.cpu().numpy(), cell_index)
            if out is not None:
                out.copy_(torch.from_numpy(vals))
                return out
            return torch.from_numpy(vals).to(x.device)
        if out is None:
            out = torch.empty(npts, dtype=torch.float64, device=x.device)
        prog = self.program()
        lib = _lib.require_device()
        _lib.check(lib.gnx_eval_expr(C.byref(prog), _lib.ptr(x), gd, npts, int(pts_per_edge), _lib.ptr(tangents),
                                     _lib.ptr(edge_attr), _lib.ptr(out), current_stream()))
        return out

    def eval_host(self, x, cell_index=None, tangents=None):
        """numpy evaluation (point evaluation of tiny sets, UserExpression callbacks)."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self.is_constant:
            return np.full(len(x), self.const_value())
        if self.user is not None:
            out = np.empty(len(x))
            buf = np.zeros(1)
            has_cell = callable(getattr(self.user, "eval_cell", None))
            has_eval = callable(getattr(self.user, "eval", None))
            for i, xi in enumerate(x):
                if has_cell and (cell_index is not None or not has_eval):
                    self.user.eval_cell(buf, xi, _Cell(cell_index[i] if cell_index is not None else 0))
                else:
                    self.user.eval(buf, xi)
                out[i] = buf[0]
            return out
        subs_syms = list(_X) + list(_TAU) + list(self._pslots)
        f = sp.lambdify(subs_syms, self.ufl.expr, "numpy")
        cols = [x[:, i] if i < x.shape[1] else np.zeros(len(x)) for i in range(3)]
        tau = [tangents[:, i] if tangents is not None and i < tangents.shape[1] else np.zeros(len(x)) for i in range(3)]
        pv = [float(self.ufl.params[s]) for s in self._pslots]
        return np.broadcast_to(np.asarray(f(*cols, *tau, *pv), dtype=np.float64), (len(x),)).copy()


def as_coefficient(c):
    """anything a graphnics model accepts as f / g / p_bc / Res -> CompiledCoefficient"""
    if isinstance(c, CompiledCoefficient):
        return c
    if isinstance(c, Constant):
        return CompiledCoefficient(constant=c, degree=0)
    if isinstance(c, (int, float, np.floating, np.integer)):
        return CompiledCoefficient(constant=Constant(float(c)), degree=0)
    if isinstance(c, (Expression, UFLExpr)):
        # compiled once per object; parameter VALUES (Constants, Expression attributes such as .t)
        # are read every time the program is instantiated, so the cache never goes stale
        cc = c.__dict__.get("_compiled")
        if cc is None or cc.degree != c.degree:
            cc = CompiledCoefficient(ufl=c.as_ufl() if isinstance(c, Expression) else c, degree=c.degree)
            c.__dict__["_compiled"] = cc
        return cc
    if isinstance(c, UserExpression):
        return CompiledCoefficient(user=c, degree=getattr(c, "degree", 2))
    if hasattr(c, "as_coefficient"):
        return c.as_coefficient()
    raise TypeError(f"unsupported coefficient type {type(c).__name__}")
