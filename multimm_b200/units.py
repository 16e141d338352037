"""Minimal stand-in for the handful of ``openmm.unit`` features the reference's config uses
(config.py:24-49 parses ``"<float> <unit-expr>"`` strings by evaluating the unit expression
against openmm.unit).  Everything is reduced to OpenMM's own unit system — nm, ps, kJ/mol,
rad, K — which is what the engine's C-ABI takes.
"""
from __future__ import annotations

import ast
import math
import operator


class Unit:
    """A product of base dimensions with a scale factor to the MD unit system."""

    __slots__ = ("scale", "dims", "name")

    def __init__(self, scale: float, dims: dict, name: str):
        self.scale = float(scale)
        self.dims = {k: v for k, v in dims.items() if v != 0}
        self.name = name

    def _combine(self, other, sign):
        if isinstance(other, (int, float)):
            return Unit(self.scale * (other if sign > 0 else 1.0 / other), self.dims, self.name)
        dims = dict(self.dims)
        for k, v in other.dims.items():
            dims[k] = dims.get(k, 0) + sign * v
        op = "*" if sign > 0 else "/"
        return Unit(self.scale * (other.scale if sign > 0 else 1.0 / other.scale), dims,
                    f"{self.name}{op}{other.name}")

    def __mul__(self, other):
        return self._combine(other, +1)

    def __truediv__(self, other):
        return self._combine(other, -1)

    def __pow__(self, p):
        return Unit(self.scale ** p, {k: v * p for k, v in self.dims.items()}, f"{self.name}**{p}")

    def get_name(self):
        return self.name

    def __repr__(self):
        return f"Unit({self.name})"


def _u(scale, name, **dims):
    return Unit(scale, dims, name)


_KCAL = 4.184
UNITS = {
    # length (-> nm)
    "nanometer": _u(1.0, "nanometer", L=1), "nanometers": _u(1.0, "nanometer", L=1),
    "angstrom": _u(0.1, "angstrom", L=1), "angstroms": _u(0.1, "angstrom", L=1),
    "picometer": _u(1e-3, "picometer", L=1), "micrometer": _u(1e3, "micrometer", L=1),
    "meter": _u(1e9, "meter", L=1),
    # time (-> ps)
    "femtosecond": _u(1e-3, "femtosecond", T=1), "femtoseconds": _u(1e-3, "femtosecond", T=1),
    "picosecond": _u(1.0, "picosecond", T=1), "picoseconds": _u(1.0, "picosecond", T=1),
    "nanosecond": _u(1e3, "nanosecond", T=1), "nanoseconds": _u(1e3, "nanosecond", T=1),
    # energy (-> kJ/mol)
    "kilojoule_per_mole": _u(1.0, "kilojoule/mole", E=1), "kilojoules_per_mole": _u(1.0, "kilojoule/mole", E=1),
    "kilocalorie_per_mole": _u(_KCAL, "kilocalorie/mole", E=1),
    "kilocalories_per_mole": _u(_KCAL, "kilocalorie/mole", E=1),
    # the names get_name() writes into config_auto.ini ("kilojoule/mole/nanometer**2") must parse
    # back: energy per amount x amount, so that kilojoule/mole == kilojoule_per_mole
    "kilojoule": _u(1.0, "kilojoule", E=1, N=1), "kilojoules": _u(1.0, "kilojoule", E=1, N=1),
    "kilocalorie": _u(_KCAL, "kilocalorie", E=1, N=1), "kilocalories": _u(_KCAL, "kilocalorie", E=1, N=1),
    "mole": _u(1.0, "mole", N=1), "moles": _u(1.0, "mole", N=1),
    # angle (-> rad)
    "radian": _u(1.0, "radian", A=1), "radians": _u(1.0, "radian", A=1),
    "degree": _u(math.pi / 180.0, "degree", A=1), "degrees": _u(math.pi / 180.0, "degree", A=1),
    # temperature
    "kelvin": _u(1.0, "kelvin", K=1), "kelvins": _u(1.0, "kelvin", K=1),
    "dimensionless": _u(1.0, "dimensionless"),
}

_BINOPS = {ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow}


def parse_unit(expr: str) -> Unit:
    """Evaluate a unit expression such as ``kilojoules_per_mole/nanometer**2`` (no eval())."""

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Name):
            if node.id not in UNITS:
                raise ValueError(f"unknown unit {node.id!r}")
            return UNITS[node.id]
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            return -ev(node.operand)
        if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
            return _BINOPS[type(node.op)](ev(node.left), ev(node.right))
        raise ValueError(f"unsupported unit expression: {expr!r}")

    out = ev(ast.parse(expr.strip(), mode="eval"))
    if not isinstance(out, Unit):
        raise ValueError(f"Expression {expr} did not evaluate to a Unit")
    return out


class Quantity:
    """value * unit; ``float(q)`` and ``q.md`` give the value in the MD unit system."""

    __slots__ = ("_value", "unit")

    def __init__(self, value: float, unit: Unit):
        self._value = float(value)
        self.unit = unit

    @property
    def md(self) -> float:
        return self._value * self.unit.scale

    def __float__(self):
        return self.md

    def value_in_unit(self, unit: Unit) -> float:
        if unit.dims != self.unit.dims:
            raise TypeError(f"incompatible units {self.unit.name} and {unit.name}")
        return self._value * self.unit.scale / unit.scale

    # the arithmetic openmm.unit.Quantity offers and the reference uses on config values
    # (model.py:678 forms 1.0 / r0 ** 2 from LE_HARMONIC_BOND_R0)
    def __pow__(self, p):
        return Quantity(self._value ** p, self.unit ** p)

    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value * other._value, self.unit * other.unit)
        return Quantity(self._value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self._value / other._value, self.unit / other.unit)
        return Quantity(self._value / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / self._value, self.unit ** -1)

    def __eq__(self, other):
        return isinstance(other, Quantity) and self.unit.dims == other.unit.dims and math.isclose(self.md, other.md)

    def __repr__(self):
        return f"Quantity({self._value}, {self.unit.name})"


nanometers = nanometer = UNITS["nanometer"]
kelvin = UNITS["kelvin"]


def parse_quantity(val) -> Quantity:
    """Same contract as the reference's parse_quantity (config.py:24-49)."""
    if isinstance(val, Quantity):
        return val
    if not isinstance(val, str) or val.strip() == "":
        raise ValueError("Invalid Quantity format")
    parts = val.strip().split(maxsplit=1)
    if len(parts) != 2:
        raise ValueError(f"Can't recognise Quantity format: {val}")
    value_str, unit_str = parts
    try:
        value = float(value_str)
    except ValueError:
        raise ValueError(f"Invalid float value: {value_str}")
    try:
        unit = parse_unit(unit_str)
    except Exception as e:  # noqa: BLE001
        raise ValueError(f"Can't recognise unit expression {unit_str} in {val}: {e}")
    return Quantity(value, unit)
