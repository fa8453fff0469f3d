"""Safe evaluator for DeviceFlow rate functions such as ``"math.sin(t)+1"``.

The reference calls ``eval(func_string)`` on request-supplied text (ols_core/deviceflow/non_grpc/strategy.py:342,
ols_core/deviceflow/utils/validate_parameters.py:217) — remote code execution by design.  Here the expression is
parsed once with ``ast`` into a closure over a whitelist: the variable ``t``, numeric literals, arithmetic,
comparisons, conditional expressions and a fixed table of ``math.*`` / ``np.*`` functions.  Anything else raises
``UnsafeExpression`` at compile time.
"""
from __future__ import annotations

import ast
import math
import operator
from typing import Callable, Dict

import numpy as np


class UnsafeExpression(ValueError):
    pass


_FUNCS: Dict[str, Callable] = {
    n: getattr(math, n) for n in (
        "sin", "cos", "tan", "asin", "acos", "atan", "atan2", "sinh", "cosh", "tanh", "exp", "expm1", "log", "log2",
        "log10", "log1p", "sqrt", "pow", "fabs", "floor", "ceil", "fmod", "hypot", "degrees", "radians", "erf",
        "erfc", "gamma", "lgamma", "copysign", "trunc")
}
_FUNCS.update({"abs": abs, "min": min, "max": max, "round": round, "float": float, "int": int})
_NP_ALIASES = {"arcsin": "asin", "arccos": "acos", "arctan": "atan", "arctan2": "atan2", "absolute": "fabs",
               "abs": "abs", "power": "pow", "maximum": "max", "minimum": "min"}
_CONSTS = {"pi": math.pi, "e": math.e, "tau": math.tau, "inf": math.inf}
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.FloorDiv: operator.floordiv, ast.Mod: operator.mod, ast.Pow: operator.pow}
_UNOPS = {ast.UAdd: operator.pos, ast.USub: operator.neg, ast.Not: operator.not_}
_CMPS = {ast.Lt: operator.lt, ast.LtE: operator.le, ast.Gt: operator.gt, ast.GtE: operator.ge, ast.Eq: operator.eq,
         ast.NotEq: operator.ne}
_MAX_NODES = 256


def _resolve_name(node: ast.AST):
    """Name or module attribute -> ('func', f) | ('const', v) | ('var', 't')"""
    if isinstance(node, ast.Name):
        if node.id == "t":
            return "var", "t"
        if node.id in _CONSTS:
            return "const", _CONSTS[node.id]
        if node.id in _FUNCS:
            return "func", _FUNCS[node.id]
        raise UnsafeExpression(f"name {node.id!r} is not allowed")
    if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ("math", "np", "numpy"):
        name = node.attr
        if node.value.id != "math":
            name = _NP_ALIASES.get(name, name)
        if name in _CONSTS:
            return "const", _CONSTS[name]
        if name in _FUNCS:
            return "func", _FUNCS[name]
        raise UnsafeExpression(f"{node.value.id}.{node.attr} is not allowed")
    raise UnsafeExpression("only t, numbers and math./np. functions are allowed")


def compile_expr(src: str) -> Callable[[float], float]:
    """Compile ``src`` into ``f(t) -> float``.  Raises UnsafeExpression for anything outside the whitelist."""
    if not isinstance(src, str) or len(src) > 512:
        raise UnsafeExpression("expression must be a string of at most 512 characters")
    try:
        tree = ast.parse(src.strip(), mode="eval")
    except SyntaxError as e:
        raise UnsafeExpression(f"syntax error: {e.msg}") from None
    if sum(1 for _ in ast.walk(tree)) > _MAX_NODES:
        raise UnsafeExpression("expression too complex")

    def build(n: ast.AST) -> Callable[[float], float]:
        if isinstance(n, ast.Expression):
            return build(n.body)
        if isinstance(n, ast.Constant):
            if isinstance(n.value, bool) or not isinstance(n.value, (int, float)):
                raise UnsafeExpression("only numeric literals are allowed")
            v = float(n.value)                  # float arithmetic only: no unbounded Python big-int intermediates
            return lambda t: v
        if isinstance(n, (ast.Name, ast.Attribute)):
            kind, val = _resolve_name(n)
            if kind == "var":
                return lambda t: t
            if kind == "const":
                return lambda t: val
            raise UnsafeExpression("function used without a call")
        if isinstance(n, ast.BinOp) and type(n.op) in _BINOPS:
            op, a, b = _BINOPS[type(n.op)], build(n.left), build(n.right)
            if isinstance(n.op, ast.Pow):
                def powf(t, a=a, b=b):
                    x, y = float(a(t)), float(b(t))
                    if abs(y) > 64 and abs(x) > 1:
                        raise UnsafeExpression("exponent too large")
                    r = x ** y
                    if isinstance(r, complex):
                        raise UnsafeExpression("complex result")
                    return r
                return powf
            return lambda t, op=op, a=a, b=b: op(float(a(t)), float(b(t)))
        if isinstance(n, ast.UnaryOp) and type(n.op) in _UNOPS:
            op, a = _UNOPS[type(n.op)], build(n.operand)
            return lambda t: op(a(t))
        if isinstance(n, ast.Compare) and len(n.ops) == 1 and type(n.ops[0]) in _CMPS:
            op, a, b = _CMPS[type(n.ops[0])], build(n.left), build(n.comparators[0])
            return lambda t: op(a(t), b(t))
        if isinstance(n, ast.IfExp):
            c, a, b = build(n.test), build(n.body), build(n.orelse)
            return lambda t: a(t) if c(t) else b(t)
        if isinstance(n, ast.BoolOp):
            parts = [build(v) for v in n.values]
            if isinstance(n.op, ast.And):
                return lambda t: all(p(t) for p in parts)
            return lambda t: any(p(t) for p in parts)
        if isinstance(n, ast.Call):
            if n.keywords:
                raise UnsafeExpression("keyword arguments are not allowed")
            kind, fn = _resolve_name(n.func)
            if kind != "func":
                raise UnsafeExpression("only whitelisted functions can be called")
            args = [build(a) for a in n.args]

            def call(t, fn=fn, args=args):
                r = fn(*[a(t) for a in args])
                # int() / round() / floor() return Python ints: bring them back to float so later ** and * stay bounded
                return float(r) if isinstance(r, int) and not isinstance(r, bool) else r
            return call
        raise UnsafeExpression(f"syntax element {type(n).__name__} is not allowed")

    f = build(tree)

    def guarded(t: float):
        """every arithmetic failure of a request-supplied expression surfaces as UnsafeExpression (never a bare
        OverflowError / ZeroDivisionError / ValueError out of the service)"""
        try:
            return f(float(t))
        except UnsafeExpression:
            raise
        except (OverflowError, ZeroDivisionError, ValueError, TypeError) as e:
            raise UnsafeExpression(f"evaluation failed: {type(e).__name__}: {e}") from None
    return guarded


def evaluate(src: str, t: float) -> float:
    return compile_expr(src)(t)


def sample(src: str, ts: np.ndarray) -> np.ndarray:
    f = compile_expr(src)
    return np.array([float(f(float(t))) for t in np.asarray(ts).ravel()], dtype=np.float64).reshape(np.shape(ts))
