"""ORACLE (test infrastructure only): numpy restatement of the device draw generator (csrc/philox.cuh).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123) with
counter = (patient lo, patient hi, column // 2, stream) and key = (seed lo, seed hi); stream 0 -> Box-Muller noise
pair (x 0.01), 1 recovery, 2 chemo, 3 radio uniforms; the top 52 bits of a 64-bit word fill the mantissa of a double
in [1,2) (the recovery stream adds 2^-53: bin centres).  The block function is pinned to the Random123 known-answer vectors in tests/test_oracle.py.  The reference
itself draws from numpy's global MT19937 stream (cancer_simulation.py:275-279); this generator only exists in
throughput mode, so parity with the reference goes through b200i_philox_draws -> b200i_sim_factual.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised block function; all arguments broadcastable unsigned integers < 2**32; returns 4 uint64 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _mant12(lo, hi):
    bits = (np.uint64(0x3FF) << np.uint64(52)) | ((hi << np.uint64(32) | lo) >> np.uint64(12))
    return bits.view(np.float64)


def draw_factual(n, T, seed, patient_base=0):
    """dict(noise, recovery, chemo, radio) of (n, T) float64 arrays, T even."""
    assert T % 2 == 0
    gp = (np.arange(n, dtype=np.uint64) + np.uint64(patient_base))[:, None]
    tp = np.arange(T // 2, dtype=np.uint64)[None, :]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = {}
    for name, s in (('noise', 0), ('recovery', 1), ('chemo', 2), ('radio', 3)):
        x, y, z, w = philox4x32_10(gp & MASK, gp >> np.uint64(32), tp, np.uint64(s), k0, k1)
        a, b = _mant12(x, y), _mant12(z, w)
        if s == 0:
            rad = 0.01 * np.sqrt(-2.0 * np.log(2.0 - a))
            ang = np.pi * (2.0 * (b - 1.0))
            ev, od = rad * np.cos(ang), rad * np.sin(ang)
        elif s == 1:      # recovery: bin centres, in (0,1)
            ev, od = (a - 1.0) + 2.0 ** -53, (b - 1.0) + 2.0 ** -53
        else:
            ev, od = a - 1.0, b - 1.0
        arr = np.empty((n, T))
        arr[:, 0::2], arr[:, 1::2] = ev, od
        out[name] = arr
    return out
