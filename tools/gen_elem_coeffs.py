#!/usr/bin/env python
"""Generate the polynomial coefficients used by include/rpgo_elem.h.

The PCM hot path needs acos / sin / cos / tan (SE(3) Logmap) and atan2 (SE(2)
Logmap).  glibc's libm and CUDA's libdevice differ in the last ulp, which would
break bit-exact parity between the CPU oracle and the sm_100a kernels, so both
sides use one deterministic implementation made only of IEEE-754 +,-,*,/,sqrt,fma.
This script derives near-minimax (Chebyshev) coefficients with mpmath and prints
them as C hex-float literals.  Run:  python tools/gen_elem_coeffs.py
"""
import mpmath as mp

mp.mp.dps = 60


def hexf(x):
    return float(x).hex()


def fit(f, a, b, n):
    # chebyfit returns coefficients highest power first
    c = mp.chebyfit(f, [a, b], n)
    return [float(v) for v in c][::-1]  # lowest power first


def f_sin(z):
    z = mp.mpf(z)
    if z < mp.mpf('1e-30'):
        return -mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.sin(r) / r - 1) / z


def f_cos(z):
    z = mp.mpf(z)
    if z < mp.mpf('1e-20'):
        return mp.mpf(1) / 24
    r = mp.sqrt(z)
    return (mp.cos(r) - 1 + z / 2) / (z * z)


def f_asin(z):
    z = mp.mpf(z)
    if z < mp.mpf('1e-30'):
        return mp.mpf(1) / 6
    r = mp.sqrt(z)
    return (mp.asin(r) / r - 1) / z


def f_atan(z):
    z = mp.mpf(z)
    if z < mp.mpf('1e-30'):
        return -mp.mpf(1) / 3
    r = mp.sqrt(z)
    return (mp.atan(r) / r - 1) / z


def emit(name, coeffs):
    print("static const double %s[%d] = {" % (name, len(coeffs)))
    for c in coeffs:
        print("    %s,  /* %.17g */" % (hexf(c), c))
    print("};")


if __name__ == "__main__":
    zmax = (mp.pi / 4 + mp.mpf('0.01')) ** 2
    emit("RPGO_SIN_C", fit(f_sin, 0, zmax, 8))
    emit("RPGO_COS_C", fit(f_cos, 0, zmax, 8))
    emit("RPGO_ASIN_C", fit(f_asin, 0, mp.mpf('0.2501'), 17))
    emit("RPGO_ATAN_C", fit(f_atan, 0, 1, 26))
    pio2 = mp.pi / 2
    p1 = float(pio2)
    p2 = float(pio2 - mp.mpf(p1))
    p3 = float(pio2 - mp.mpf(p1) - mp.mpf(p2))
    print("PIO2_1", hexf(p1), "PIO2_2", hexf(p2), "PIO2_3", hexf(p3))
    pi1 = float(mp.pi)
    pi2 = float(mp.pi - mp.mpf(pi1))
    print("PI_1", hexf(pi1), "PI_2", hexf(pi2))
    pio4 = float(mp.pi / 4)
    print("PIO4_1", hexf(pio4), "PIO4_2", hexf(mp.pi / 4 - mp.mpf(pio4)))
    print("TWO_OVER_PI", hexf(2 / mp.pi))
