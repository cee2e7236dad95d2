"""CPU restatements of three device formulas of render_kernel.cu that replace a library call or an IEEE operation by a shorter
sequence (second session of round 2).  Each is replayed here with exactly rounded arithmetic (fractions.Fraction -> float is a
correctly rounded conversion, so `fma` below is the IEEE fused multiply-add) and checked against the operation it replaces;
the GPU parity tests check the kernels themselves.

  div5        x / 5.f of the AA mean (renderer.d:249)      == IEEE single-precision division
  sin_phase   one-DFMA phase reduction of Procedure2's sines (texture.d:82-83)
  sincos_rev  64-entry rotation table + Taylor kernels for the lens sample (camera.d:258-269)
"""
import math
import os
import re
import struct
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "chess2rt_b200", "csrc", "render_kernel.cu")).read()


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def f32(x):
    return float(np.float32(x))


def fma32(a, b, c):
    # exact a * b + c, rounded once to binary32 (through binary64: the values below are far from binary32 half-way cases
    # except by a 2^-29 accident, which a fixed seed makes a non-event)
    return f32(Fraction(a) * Fraction(b) + Fraction(c))


def test_div5_is_the_ieee_division():
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.random(3000, dtype=np.float32) * np.float32(5), rng.random(2000, dtype=np.float32) * np.float32(4000),
                         np.float32([0.0, 1.0, 5.0, 2.5, 1e-3, 255.0, 1275.0, 3.0000002, 4.9999995])])
    y = fma32(fma32(f32(0.2), -5.0, 1.0), f32(0.2), f32(0.2))
    assert y == f32(0.2)   # the refined reciprocal is 0.2f itself (the kernel's comment)
    for x in xs:
        x = float(x)
        q = f32(Fraction(x) * Fraction(y))
        got = fma32(y, fma32(q, -5.0, x), q)
        assert got == float(np.float32(x) / np.float32(5)), x


def test_sin_phase_is_the_fraction_of_a_revolution():
    """fma(u, f * 2^32, 1.5 * 2^52) leaves round(u f 2^32) mod 2^32 in the low mantissa word, for either sign of u f."""
    rng = np.random.default_rng(11)
    MAGIC = 6755399441055744.0
    for _ in range(3000):
        u = float(rng.normal()) * 10.0 ** rng.integers(-3, 5)
        f = float(rng.normal()) * 10.0 ** rng.integers(-3, 1) / (2 * math.pi)    # revolutions per unit
        if abs(u * f) >= 2.0 ** 18:
            continue
        F = f * 4294967296.0
        w = fma(u, F, MAGIC)
        p = struct.unpack("<Q", struct.pack("<d", w))[0] & 0xFFFFFFFF
        exact = Fraction(u) * Fraction(F)                     # phase in units of 2^-32 revolutions
        want = int(round(exact)) % (1 << 32)                   # (ties: measure zero for random inputs)
        assert p == want or abs(exact - round(exact)) == Fraction(1, 2)
        # the float handed to the SFU sine: 1 + top 23 bits of the fraction
        t = struct.unpack("<f", struct.pack("<I", (p >> 9) | 0x3F800000))[0]
        frac = float(exact / (1 << 32) % 1)
        d = abs((t - 1.0) - frac)
        assert min(d, 1.0 - d) < 2.0 ** -23 + 2.0 ** -32
    # the per-texture bound the kernel compares high words against keeps |u F| below 2^51 (c2rt_api.cu)
    fmax = 0.25 / (2 * math.pi) * 4294967296.0
    lim = 262144.0 * 4294967296.0 / fmax
    hi = struct.unpack("<Q", struct.pack("<d", lim))[0] >> 32
    lim_floor = struct.unpack("<d", struct.pack("<Q", hi << 32))[0]
    assert lim_floor <= lim and lim_floor * fmax < 2.0 ** 51


def _table():
    body = SRC[SRC.index("c_rot64[64] = {"):]
    body = body[:body.index("};")]
    vals = [float(v) for v in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", body)]
    assert len(vals) == 128
    return [(vals[2 * k], vals[2 * k + 1]) for k in range(64)]


def _coeffs():
    body = SRC[SRC.index("c_sincos[12] = {"):]
    body = body[:body.index("};")]
    body = re.sub(r"//[^\n]*", "", body)
    vals = [float(v) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", body.split("{", 1)[1])]
    assert len(vals) == 12
    return vals


def _dec_sincos(x):
    getcontext().prec = 50
    x = Decimal(x)
    s, term, n = x, x, 1
    while abs(term) > Decimal(10) ** -45:
        term = -term * x * x / ((2 * n) * (2 * n + 1))
        s += term
        n += 1
    c, term, n = Decimal(1), Decimal(1), 1
    while abs(term) > Decimal(10) ** -45:
        term = -term * x * x / ((2 * n - 1) * (2 * n))
        c += term
        n += 1
    return float(s), float(c)


PI = Decimal("3.14159265358979323846264338327950288419716939937510")


def test_rotation_table_is_correctly_rounded():
    getcontext().prec = 50
    for k, (c, s) in enumerate(_table()):
        a = 2 * PI * k / 64
        if a > PI:
            a -= 2 * PI
        ws, wc = _dec_sincos(a)
        assert abs(c - wc) < 1e-30 + 1.2e-16 * abs(wc) and abs(s - ws) < 1e-30 + 1.2e-16 * abs(ws), k


def test_sincos_rev_matches_sincos_of_the_angle():
    K = _coeffs()
    T = _table()
    MAGIC = K[9]
    assert (K[8], K[9], K[10], K[11]) == (2 * math.pi, 6755399441055744.0, -0.015625, 64.0)
    rng = np.random.default_rng(3)
    us = list(rng.integers(0, 2 ** 31 - 1, 1500) / 2147483647.0) + [0.0, 1.0, 0.5, 0.25, 1 / 128, 1 / 64, 127 / 128, 0.9999999995343387]
    worst = 0.0
    for u in us:
        u = float(u)
        qm = fma(u, K[11], MAGIC)
        k = struct.unpack("<Q", struct.pack("<d", qm))[0] & 63
        c0, s0 = T[k]
        r = fma(qm - MAGIC, K[10], u)
        assert abs(r) <= 1 / 128
        t = r * K[8]
        z = t * t
        ps = K[0]
        for j in (1, 2, 3):
            ps = fma(ps, z, K[j])
        st = fma(t * z, ps, t)
        pc = K[4]
        for j in (5, 6, 7):
            pc = fma(pc, z, K[j])
        ct = fma(pc, z, 1.0)
        s = fma(c0, st, s0 * ct)
        c = fma(-s0, st, c0 * ct)
        getcontext().prec = 50
        a = 2 * PI * Decimal(u)
        if a > PI:
            a -= 2 * PI
        ws, wc = _dec_sincos(a)
        worst = max(worst, abs(s - ws), abs(c - wc))
    assert worst < 4e-16, worst
