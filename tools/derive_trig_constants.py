"""Derives the constants of the sin / cos / tan / cot sequences (device/gdv_device_lib.cuh and
oracle/gdv_oracle.cc hold the output): 1280 bits of 2/pi for the integer argument reduction, pi/2 as
a double-double, and the Taylor coefficients (-1)^k / (2k+1)!, (-1)^k / (2k)! as round-to-nearest
doubles.  pi comes from Machin's formula in integer arithmetic; nothing is read from a library."""
from fractions import Fraction
import struct

BITS = 1280
GUARD = 128


def arctan_inv(n, scale):   # atan(1/n) * scale, integer series
    total, term, k = 0, scale // n, 0
    n2 = n * n
    while term:
        total += term // (2 * k + 1) if k % 2 == 0 else -(term // (2 * k + 1))
        term //= n2
        k += 1
    return total


scale = 1 << (BITS + GUARD)
pi_scaled = 4 * (4 * arctan_inv(5, scale) - arctan_inv(239, scale))          # pi * 2^(BITS+GUARD)
two_over_pi = (2 << (2 * (BITS + GUARD))) // pi_scaled                        # (2/pi) * 2^(BITS+GUARD)
two_over_pi >>= GUARD                                                         # 1280 fraction bits
words = [(two_over_pi >> (BITS - 64 * (k + 1))) & ((1 << 64) - 1) for k in range(BITS // 64)]
print("2/pi, 64 bits per word, most significant first:")
for k in range(0, len(words), 4):
    print("  " + ", ".join("0x%016xull" % w for w in words[k:k + 4]) + ",")


def rn(fr):  # Fraction -> nearest double (Python's int / int division is correctly rounded)
    return fr.numerator / fr.denominator


pi_half = Fraction(pi_scaled, 2 * scale)
hi = rn(pi_half)
lo = rn(pi_half - Fraction(hi))
print("pi/2 hi = %r (%s)  lo = %r (%s)" % (hi, struct.pack(">d", hi).hex(), lo, struct.pack(">d", lo).hex()))
print("pi/4 as double = %r" % rn(pi_half / 2))
fact = 1
coef_s, coef_c = [], []
for n in range(2, 19):
    fact *= n
    if n % 2 == 1:
        coef_s.append(rn(Fraction((-1) ** (n // 2), fact)))
    elif n >= 4:
        coef_c.append(rn(Fraction((-1) ** (n // 2), fact)))
print("S1..S8 =", ", ".join(repr(c) for c in coef_s))
print("C1..C7 =", ", ".join(repr(c) for c in coef_c))

# ln 2 in Q1.127 for power(): sum 1 / (k 2^k) in integers with 200 guard bits
G = 200
total, k = 0, 1
while (1 << (127 + G)) // (k << k):
    total += (1 << (127 + G)) // (k << k)
    k += 1
print("ln 2 * 2^127 = 0x%032x" % (total >> G))
print("log2(e) * 2^126 = 0x%032x" % ((1 << (126 + 127 + G)) // total))

# atan(2^-i) in Q2.126 for the CORDIC of atan / atan2 / asin / acos (i = 0..42; below that 2^-i itself),
# and pi in Q2.126
def atan_pow2(i, scale):   # atan(2^-i) * scale; i = 0: pi / 4
    if i == 0:
        return pi_scaled * scale // (4 << (BITS + GUARD))
    return arctan_inv(1 << i, scale)


S126 = 1 << (126 + 64)
print("atan(2^-i) * 2^126:")
for i in range(43):
    v = atan_pow2(i, S126) >> 64
    print("  ((u128)0x%016xull << 64) | 0x%016xull,  // i = %d" % (v >> 64, v & ((1 << 64) - 1), i))
pi_q = (pi_scaled >> (BITS + GUARD - 126 - 8)) >> 8
print("pi * 2^126 = ((u128)0x%016xull << 64) | 0x%016xull" % (pi_q >> 64, pi_q & ((1 << 64) - 1)))
for name, fr in (("pi", Fraction(pi_scaled, scale)), ("pi/2", Fraction(pi_scaled, 2 * scale)), ("pi/4", Fraction(pi_scaled, 4 * scale)),
                 ("3pi/4", Fraction(3 * pi_scaled, 4 * scale))):
    print("RN(%s) = %r" % (name, rn(fr)))
