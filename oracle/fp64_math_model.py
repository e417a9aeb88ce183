"""Exact-FMA CPU model of the kernels' table-driven fp64 log / exp / expm1 (csrc/fp64_math.cuh), reading the same
tables (csrc/fp64_tables.inc) — every fma is evaluated in rational arithmetic and rounded once, as the GPU does.

TEST INFRASTRUCTURE ONLY.  Used by tests/test_host_side.py to check tables, constants and algorithms against mpmath
without a GPU (Constants(True) models the immediate-constant variant that was measured on the B200 in r02a and dropped:
constants whose low 32 bits are zero; kept as a record of its accuracy).  The functions the reference evaluates through libm here are np.log / np.exp /
pow / cosh / sinh inside pde_rhs (marlpde/LHeureux_model.py:413-520)."""
import math
import os
import re
import struct
from fractions import Fraction as Fr

_INC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                    "integrating-diagenetic-equations-using-python_b200", "csrc", "fp64_tables.inc")
_SRC = open(_INC).read()


def _table(name):
    body = _SRC[_SRC.index(name):]
    body = body[body.index("{") + 1:body.index("};")]
    return [(float.fromhex(a), float.fromhex(b)) for a, b in re.findall(r"\{(\S+), (\S+)\}", body)]


def _const(name):
    return float.fromhex(re.search(name + r" = (\S+);", _SRC).group(1))


LOG_TAB, EXP_TAB = _table("kLogTab"), _table("kExpTab")


def fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))


def _hi_lo(x):
    b = struct.unpack("<Q", struct.pack("<d", x))[0]
    return b >> 32, b & 0xFFFFFFFF


def _from_hi_lo(h, lo):
    return struct.unpack("<d", struct.pack("<Q", ((h & 0xFFFFFFFF) << 32) | (lo & 0xFFFFFFFF)))[0]


class Constants:
    """imm = False: the constants of the default build; True: those of MARLPDE_FP64_IMM=1 (fp64_math.cuh)."""

    def __init__(self, imm=False):
        self.log = [-0.5, 1 / 3, -0.25, 0.2, -1 / 6, 1 / 7]
        self.exp = [0.5, 1 / 6, 1 / 24, 1 / 120, 1 / 720]
        self.ln2hi, self.ln2lo = _const("kLn2Hi"), _const("kLn2Lo")
        self.l64hi, self.l64lo, self.k64 = _const("kLn2_64Hi"), _const("kLn2_64Lo"), _const("k64_Ln2")
        if imm:
            self.log[3], self.log[4] = float.fromhex("0x1.9999ap-3"), float.fromhex("-0x1.55555p-3")
            self.exp[3], self.exp[4] = float.fromhex("0x1.11111p-7"), float.fromhex("0x1.6c16cp-10")
            self.ln2hi, self.ln2lo = float.fromhex("0x1.62e42p-1"), float.fromhex("0x1.fdf473de6af28p-22")
            self.l64hi, self.l64lo = float.fromhex("0x1.62e42p-7"), float.fromhex("0x1.fdf473de6af28p-28")
            self.k64 = float.fromhex("0x1.71547p+6")
            for v in (self.ln2hi, self.l64hi, self.k64, self.log[0], self.log[2], self.log[3], self.log[4],
                      self.exp[0], self.exp[3], self.exp[4]):
                assert _hi_lo(v)[1] == 0          # encodable as an instruction immediate


def log(c, x):
    """fm::log / fm::log_nb for a positive normal finite x."""
    h, lo = _hi_lo(x)
    e, j = (h >> 20) - 1023, (h >> 13) & 127
    m = _from_hi_lo((h & 0xFFFFF) | 0x3FF00000, lo)
    ic, L = LOG_TAB[j]
    r = fma(m, ic, -1.0)
    p = fma(c.log[5], r, c.log[4])
    for k in (3, 2, 1, 0):
        p = fma(p, r, c.log[k])
    l1p = fma(r * r, p, r)
    return fma(float(e), c.ln2hi, L) + fma(float(e), c.ln2lo, l1p)


def _exp_core(c, x, extra):
    magic = 6755399441055744.0
    tn = fma(x, c.k64, magic)
    n = struct.unpack("<i", struct.pack("<I", _hi_lo(tn)[1]))[0]
    nf = tn - magic
    r = fma(nf, -c.l64lo, fma(nf, -c.l64hi, x))
    p = fma(c.exp[4], r, c.exp[3]) if extra else c.exp[3]
    for i in (2, 1, 0):
        p = fma(p, r, c.exp[i])
    return fma(r * r, p, r), EXP_TAB[n & 63], n >> 6


def exp(c, x):
    """fm::exp / fm::exp_nb for |x| < 690."""
    p, T, k = _exp_core(c, x, False)
    return math.ldexp(fma(T[0], p, T[0]), k)


def expm1(c, x):
    """fm::expm1 / fm::expm1_nb for |x| < 600."""
    p, T, k = _exp_core(c, x, True)
    sc = math.ldexp(1.0, k)
    s, sl = T[0] * sc, T[1] * sc
    return fma(s, p, s - 1.0) + fma(sl, p, sl)
