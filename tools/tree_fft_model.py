"""Lane-level numpy model of the warp-resident negacyclic tree FFT used by the blind-rotation kernel (design check).

Mirrors experimental-tfhe_b200/csrc/tree_fft.cuh step by step: T = M/16 lanes, 16 points per lane, pass A (depths 0-3,
lane-uniform twiddles), shared-memory transpose, pass B (depths 4-7), then LOGM-8 shuffle stages.  Verifies
  * forward o backward == M * identity,
  * pointwise products give the negacyclic convolution of the folded real polynomials.
"""
import sys
import numpy as np


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def tw(d, nu):
    """twiddle of node (depth d, index nu): exp(2 pi i (1 + 4 rev_d(nu)) / 2^(d+3))"""
    return np.exp(2j * np.pi * (1 + 4 * bitrev(nu, d)) / 2 ** (d + 3))


def pass16(v, node0_of_lane, depth0, inverse):
    """v: [T][16]; 4 depths depth0..depth0+3; node index at depth0 for each lane = node0_of_lane (array[T])."""
    T = v.shape[0]
    rs = range(4) if not inverse else range(3, -1, -1)
    for r in rs:
        span = 8 >> r
        for lo in range(16):
            if (lo // span) % 2:
                continue
            hi = lo + span
            sigma = lo // (2 * span)
            w = np.array([tw(depth0 + r, (int(node0_of_lane[t]) << r) + sigma) for t in range(T)])
            if not inverse:
                a, b = v[:, lo].copy(), v[:, hi].copy()
                v[:, lo] = a + w * b
                v[:, hi] = a - w * b
            else:
                a, b = v[:, lo].copy(), v[:, hi].copy()
                v[:, lo] = a + b
                v[:, hi] = (a - b) * np.conj(w)


def cswap(v, mask, h):
    """lane(h=0).v[8+k] <-> lane(h=1).v[k] with partner lane ^ mask"""
    T = v.shape[0]
    out = v.copy()
    for t in range(T):
        pt = t ^ mask
        for k in range(8):
            if h[t] == 0:
                out[t, 8 + k] = v[pt, k]
            else:
                out[t, k] = v[pt, 8 + k]
    return out


def forward(z, logM):
    M = 1 << logM
    T = M // 16
    P = T // 16                      # lanes per depth-4 node
    NS = logM - 8
    lanes = np.arange(T)
    v = np.array([[z[t + T * m] for m in range(16)] for t in range(T)], dtype=complex)
    pass16(v, np.zeros(T, int), 0, False)
    # transpose: element (node m, coef t) -> lane t' = P*b + p holds (b, p + P*u)
    buf = {}
    for t in range(T):
        for m in range(16):
            buf[(m, t)] = v[t, m]
    b = lanes // P
    p = lanes % P
    v = np.array([[buf[(int(b[t]), int(p[t]) + P * u)] for u in range(16)] for t in range(T)], dtype=complex)
    pass16(v, b, 4, False)
    node = 16 * b                     # node index prefix at depth 8 is 16b + u
    prev = None
    for s in range(NS):
        mask = P >> (s + 1)
        h = (p >> (NS - 1 - s)) & 1
        v = cswap(v, mask, h)
        for k in range(8):
            if s == 0:
                nu = 16 * b + 8 * h + k
            else:
                nu = 2 * prev[:, k] + h
            w = np.array([tw(8 + s, int(nu[t])) for t in range(T)])
            a, c = v[:, k].copy(), v[:, 8 + k].copy()
            v[:, k] = a + w * c
            v[:, 8 + k] = a - w * c
        if s == 0:
            prev = np.array([[16 * b[t] + 8 * h[t] + k for k in range(8)] for t in range(T)])
        else:
            prev = np.array([[2 * prev[t, k] + h[t] for k in range(8)] for t in range(T)])
    return v


def backward(v, logM):
    M = 1 << logM
    T = M // 16
    P = T // 16
    NS = logM - 8
    lanes = np.arange(T)
    b = lanes // P
    p = lanes % P
    v = v.copy()
    prevs = []
    prev = None
    for s in range(NS):
        h = (p >> (NS - 1 - s)) & 1
        if s == 0:
            prev = np.array([[16 * b[t] + 8 * h[t] + k for k in range(8)] for t in range(T)])
        else:
            prev = np.array([[2 * prev[t, k] + h[t] for k in range(8)] for t in range(T)])
        prevs.append(prev)
    for s in range(NS - 1, -1, -1):
        mask = P >> (s + 1)
        h = (p >> (NS - 1 - s)) & 1
        for k in range(8):
            w = np.array([tw(8 + s, int(prevs[s][t, k])) for t in range(T)])
            a, c = v[:, k].copy(), v[:, 8 + k].copy()
            v[:, k] = a + c
            v[:, 8 + k] = (a - c) * np.conj(w)
        v = cswap(v, mask, h)
    pass16(v, b, 4, True)
    buf = {}
    for t in range(T):
        for u in range(16):
            buf[(int(b[t]), int(p[t]) + P * u)] = v[t, u]
    v = np.array([[buf[(m, t)] for m in range(16)] for t in range(T)], dtype=complex)
    pass16(v, np.zeros(T, int), 0, True)
    z = np.zeros(M, complex)
    for t in range(T):
        for m in range(16):
            z[t + T * m] = v[t, m]
    return z


# ---------------------------------------------------------------------------------------------------------------
# Select-free variant (tree_fft.cuh as shipped): the lane exchange of a shuffle stage always sends the ODD registers
# v[2m+1] and receives into them, whatever the lane's side h.  Lanes with h = 1 run the stage BEFORE the exchange with
# negated twiddles (their two outputs land swapped) and the stage AFTER it with conjugated twiddles on (hi, lo) instead
# of (lo, hi); their results then carry a known unit factor g per slot, the same for every transform, so pointwise
# multiply-accumulates against TRUE key spectra need no correction (a and the accumulator carry the same g).
# ---------------------------------------------------------------------------------------------------------------
def tables2(logM):
    """per-lane twiddles: tb_sign[T] (depth-7 sign), c8[T][8], c9[T][8] (logM = 10), g[T][16]"""
    M = 1 << logM
    T = M // 16
    P = T // 16
    NS = logM - 8
    tb_sign = np.ones(T)
    c8 = np.zeros((T, 8), complex)
    c9 = np.zeros((T, 8), complex)
    g = np.ones((T, 16), complex)
    for t in range(T):
        b, p = t // P, t % P
        h8 = (p >> (NS - 1)) & 1
        h9 = p & 1 if NS > 1 else 0
        if h8:
            tb_sign[t] = -1.0
        for m in range(8):
            A = 16 * b + 2 * m + h8                       # depth-8 node this lane finishes
            w8 = tw(8, A)
            c = np.conj(w8) if h8 else w8
            if NS > 1 and h9:
                c = -c
            c8[t, m] = c
            g8 = (np.conj(w8), -np.conj(w8)) if h8 else (1.0, 1.0)     # factors on (plus, minus) of node A
            if NS == 1:
                g[t, 2 * m], g[t, 2 * m + 1] = g8
            else:
                D = 2 * A + h9                              # depth-9 node: plus child (h9 = 0) or minus child (h9 = 1)
                w9 = tw(9, D)
                c9[t, m] = np.conj(w9) if h9 else w9
                base = g8[h9]
                g9 = (np.conj(w9), -np.conj(w9)) if h9 else (1.0, 1.0)
                g[t, 2 * m], g[t, 2 * m + 1] = base * g9[0], base * g9[1]
    return tb_sign, c8, c9, g


def odd_swap(v, mask):
    out = v.copy()
    for t in range(v.shape[0]):
        out[t, 1::2] = v[t ^ mask, 1::2]
    return out


def pass16_signed(v, node0_of_lane, depth0, inverse, sign7):
    """pass16 with the depth0+3 twiddles multiplied by sign7[t]"""
    T = v.shape[0]
    rs = range(4) if not inverse else range(3, -1, -1)
    for r in rs:
        span = 8 >> r
        for lo in range(16):
            if (lo // span) % 2:
                continue
            hi = lo + span
            sigma = lo // (2 * span)
            w = np.array([tw(depth0 + r, (int(node0_of_lane[t]) << r) + sigma) * (sign7[t] if r == 3 else 1.0) for t in range(T)])
            a, b = v[:, lo].copy(), v[:, hi].copy()
            if not inverse:
                v[:, lo] = a + w * b
                v[:, hi] = a - w * b
            else:
                v[:, lo] = a + b
                v[:, hi] = (a - b) * np.conj(w)


def forward2(z, logM):
    M = 1 << logM
    T = M // 16
    P = T // 16
    NS = logM - 8
    lanes = np.arange(T)
    tb_sign, c8, c9, g = tables2(logM)
    v = np.array([[z[t + T * m] for m in range(16)] for t in range(T)], dtype=complex)
    pass16(v, np.zeros(T, int), 0, False)
    buf = {}
    for t in range(T):
        for m in range(16):
            buf[(m, t)] = v[t, m]
    b = lanes // P
    p = lanes % P
    v = np.array([[buf[(int(b[t]), int(p[t]) + P * u)] for u in range(16)] for t in range(T)], dtype=complex)
    pass16_signed(v, b, 4, False, tb_sign)
    for s, c in enumerate((c8, c9)[:NS]):
        v = odd_swap(v, P >> (s + 1))
        for m in range(8):
            x, y = v[:, 2 * m].copy(), v[:, 2 * m + 1].copy()
            v[:, 2 * m] = x + c[:, m] * y
            v[:, 2 * m + 1] = x - c[:, m] * y
    return v


def backward2(v, logM):
    M = 1 << logM
    T = M // 16
    P = T // 16
    NS = logM - 8
    lanes = np.arange(T)
    tb_sign, c8, c9, g = tables2(logM)
    b = lanes // P
    p = lanes % P
    v = v.copy()
    for s in range(NS - 1, -1, -1):
        c = (c8, c9)[s]
        for m in range(8):
            x, y = v[:, 2 * m].copy(), v[:, 2 * m + 1].copy()
            v[:, 2 * m] = x + y
            v[:, 2 * m + 1] = (x - y) * np.conj(c[:, m])
        v = odd_swap(v, P >> (s + 1))
    pass16_signed(v, b, 4, True, tb_sign)
    buf = {}
    for t in range(T):
        for u in range(16):
            buf[(int(b[t]), int(p[t]) + P * u)] = v[t, u]
    v = np.array([[buf[(m, t)] for m in range(16)] for t in range(T)], dtype=complex)
    pass16(v, np.zeros(T, int), 0, True)
    z = np.zeros(M, complex)
    for t in range(T):
        for m in range(16):
            z[t + T * m] = v[t, m]
    return z


def negacyclic(a, b):
    N = len(a)
    full = np.convolve(a, b)
    res = full[:N].copy()
    res[: N - 1] -= full[N:]
    return res


def main():
    rng = np.random.default_rng(0)
    for logM in (9, 10):
        M = 1 << logM
        N = 2 * M
        a = rng.integers(-512, 512, N).astype(float)
        b = rng.integers(-1000, 1000, N).astype(float)
        za = a[:M] + 1j * a[M:]
        zb = b[:M] + 1j * b[M:]
        fa, fb = forward(za, logM), forward(zb, logM)
        back = backward(fa, logM) / M
        assert np.abs(back - za).max() < 1e-8, "round trip failed"
        prod = backward(fa * fb, logM) / M
        c = np.concatenate([prod.real, prod.imag])
        ref = negacyclic(a, b)
        err = np.abs(c - ref).max()
        print(f"logM={logM}: round trip ok, negacyclic product max err {err:.3e}")
        assert err < 1e-4
    for logM in (9, 10):
        M = 1 << logM
        N = 2 * M
        a = rng.integers(-512, 512, N).astype(float)
        b = rng.integers(-1000, 1000, N).astype(float)
        za = a[:M] + 1j * a[M:]
        zb = b[:M] + 1j * b[M:]
        g = tables2(logM)[3]
        fa_c, fb_true = forward2(za, logM), forward2(zb, logM) / g
        # the true spectrum is the same multiset of values as the select-based transform produces
        ref_set = np.sort_complex(np.round(forward(zb, logM).ravel(), 6))
        assert np.abs(np.sort_complex(np.round(fb_true.ravel(), 6)) - ref_set).max() < 1e-5, "spectrum values differ"
        assert np.abs(backward2(fa_c, logM) / M - za).max() < 1e-8, "round trip (select-free) failed"
        prod = backward2(fa_c * fb_true, logM) / M
        err = np.abs(np.concatenate([prod.real, prod.imag]) - negacyclic(a, b)).max()
        print(f"logM={logM} select-free: round trip ok, negacyclic product max err {err:.3e}")
        assert err < 1e-4
    print("tree FFT model OK")


if __name__ == "__main__":
    sys.exit(main())
