"""Unit bookkeeping behind the reference's helpers (xequinet/utils/qc.py:13-148: `unit_conversion`,
`set_default_units`, `get_default_units`).  Everything is expressed in Hartree atomic units from the CODATA 2018
constants the reference uses; unit expressions ("kcal/mol/Angstrom", "eV/Angstrom^3", "e*Angstrom") are evaluated
by a small recursive-descent parser (names, integers, * / ^ and parentheses) instead of `eval`."""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional

from . import keys


def _atomic_unit_table() -> Dict[str, float]:
    c = 299792458.0               # speed of light, m/s (exact)
    mu0 = 4.0e-7 * math.pi        # vacuum permeability as used by the reference
    h = 6.62607015e-34            # Planck constant, J s (exact)
    e = 1.602176634e-19           # elementary charge, C (exact)
    me = 9.1093837015e-31         # electron mass, kg
    na = 6.02214076e23            # Avogadro constant (exact)
    amu = 1.66053906660e-27       # atomic mass unit, kg
    eps0 = 1.0 / mu0 / c**2
    hbar = h / (2.0 * math.pi)
    four_pi_eps0 = 4.0 * math.pi * eps0

    metre = me * e**2 / (four_pi_eps0 * hbar**2)   # 1 m in Bohr
    joule = (four_pi_eps0 * hbar) ** 2 / (me * e**4)  # 1 J in Hartree
    second = me * e**4 / four_pi_eps0**2 / hbar**3  # 1 s in atomic time units
    t: Dict[str, float] = {}

    def put(value: float, *names: str) -> None:
        for n in names:
            t[n] = value

    put(1.0, "AU", "au", "e", "Bohr", "a0", "Hartree", "Ha", "Eh")
    put(na, "mol")
    put(1.0 / e, "Coulomb", "C")
    put(metre, "meter", "m")
    put(metre * 1e-10, "Angstrom", "Ang")
    put(metre * 1e-2, "cm")
    put(metre * 1e-9, "nm")
    put(1.0 / amu, "kg")
    put(1e-3 / amu, "g")
    put(joule, "Joule", "J")
    put(joule * 1e3, "kJoule", "kJ")
    put(joule * e, "eV")
    put(joule * e * 1e-3, "meV")
    put(joule * 4.184, "cal")
    put(joule * 4184.0, "kcal")
    put(me * e / (1e21 * four_pi_eps0 * hbar**2 * c), "Debye", "D")
    put(second, "second", "s")
    put(second * 1e-15, "fs")
    put(second * 1e-12, "ps")
    pascal = joule / metre**3
    put(pascal, "Pascal", "Pa")
    put(pascal * 1e9, "GPa")
    put(pascal * 1e5, "bar")
    put(pascal * 1e8, "kbar")
    put(0.5, "Bohr_magneton", "muB")
    return t


UNITS = _atomic_unit_table()
_TOKEN = re.compile(r"\s*([A-Za-z_][A-Za-z_0-9]*|\d+|\*\*|[*/^()])")


def _tokens(expr: str) -> List[str]:
    out, pos = [], 0
    expr = expr.rstrip()
    while pos < len(expr):
        m = _TOKEN.match(expr, pos)
        if not m:
            raise ValueError(f"Invalid unit {expr}")
        out.append("^" if m.group(1) == "**" else m.group(1))
        pos = m.end()
    return out


def eval_unit(unit: str) -> float:
    """Value of a unit expression in atomic units (utils/qc.py:96-103)."""
    toks = _tokens(unit)
    pos = 0

    def atom() -> float:
        nonlocal pos
        if pos >= len(toks):
            raise ValueError(f"Invalid unit {unit}")
        tok = toks[pos]
        pos += 1
        if tok == "(":
            v = product()
            if pos >= len(toks) or toks[pos] != ")":
                raise ValueError(f"Invalid unit {unit}")
            pos += 1
            return v
        if tok.isdigit():
            return float(tok)
        if tok in UNITS:
            return UNITS[tok]
        raise ValueError(f"Invalid unit {unit}")

    def power() -> float:
        nonlocal pos
        base = atom()
        if pos < len(toks) and toks[pos] == "^":
            pos += 1
            return base ** power()
        return base

    def product() -> float:
        nonlocal pos
        v = power()
        while pos < len(toks) and toks[pos] in "*/":
            op = toks[pos]
            pos += 1
            rhs = power()
            v = v * rhs if op == "*" else v / rhs
        return v

    value = product()
    if pos != len(toks):
        raise ValueError(f"Invalid unit {unit}")
    return value


def check_unit(unit: str) -> bool:
    try:
        eval_unit(unit)
        return True
    except ValueError:
        return False


def unit_conversion(unit_in: Optional[str], unit_out: Optional[str]) -> float:
    """Factor that takes a number in `unit_in` to `unit_out` (utils/qc.py:106-114)."""
    if unit_in is None or unit_out is None or unit_in == unit_out:
        return 1.0
    return eval_unit(unit_in) / eval_unit(unit_out)


DEFAULT_UNITS_MAP: Dict[str, str] = {keys.POSITIONS: "Angstrom"}
_GRAD_PROPERTIES = {keys.FORCES, "base_forces", keys.VIRIAL}
_BASE_PROPERTIES = {"base_energy": keys.TOTAL_ENERGY, "base_forces": keys.FORCES, "base_charges": keys.ATOMIC_CHARGES,
                    "base_dipole": keys.DIPOLE}


def set_default_units(unit_dict: Dict[str, str]) -> None:
    """utils/qc.py:117-144: units of the derived (gradient / base) properties follow from energy, length and charge."""
    for prop, unit in unit_dict.items():
        if prop in _GRAD_PROPERTIES:
            raise ValueError("Please do not set units for gradient properties directly. Set the units for the "
                             "corresponding properties instead.")
        if prop in _BASE_PROPERTIES:
            raise ValueError("Please do not set units for base properties directly. Set the units for the "
                             "corresponding properties instead.")
        if prop == keys.ATOMIC_CHARGES:
            raise ValueError("Please do not set units for atomic charges. Set the charge instead.")
        if not check_unit(unit):
            raise ValueError(f"Invalid unit {unit} for property {prop}")
    DEFAULT_UNITS_MAP.update(unit_dict)
    if keys.TOTAL_ENERGY in DEFAULT_UNITS_MAP:
        energy_unit, pos_unit = DEFAULT_UNITS_MAP[keys.TOTAL_ENERGY], DEFAULT_UNITS_MAP[keys.POSITIONS]
        DEFAULT_UNITS_MAP[keys.FORCES] = f"{energy_unit}/{pos_unit}"
        DEFAULT_UNITS_MAP[keys.VIRIAL] = f"{energy_unit}/{pos_unit}^3"
    if keys.TOTAL_CHARGE in DEFAULT_UNITS_MAP:
        DEFAULT_UNITS_MAP[keys.ATOMIC_CHARGES] = DEFAULT_UNITS_MAP[keys.TOTAL_CHARGE]
    for base_prop, prop in _BASE_PROPERTIES.items():
        if prop in DEFAULT_UNITS_MAP:
            DEFAULT_UNITS_MAP[base_prop] = DEFAULT_UNITS_MAP[prop]


def get_default_units() -> Dict[str, str]:
    return DEFAULT_UNITS_MAP
