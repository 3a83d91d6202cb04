"""Numpy model of the warp-level 8 x 8 x R3 FFT used by csrc/warp_fft.cuh.

Design aid (not product, not oracle): mirrors the register/lane/shared-memory index
maps of the CUDA code so they can be checked against numpy.fft on the CPU, and counts
shared-memory bank conflicts of the two exchanges.
N = 64*R3, team of TL = N/40 lanes, 40 complex values per lane.
"""
import sys
import numpy as np


def conflicts(addrs):
    """addrs: [TL] 8-byte word addresses of one warp-wide LDS/STS.64.
    Returns wavefronts needed (half-warp granularity, 16 banks of 8 bytes)."""
    wf = 0
    for h in range(0, len(addrs), 16):
        half = addrs[h:h + 16]
        banks = {}
        for a in half:
            banks.setdefault(a % 16, set()).add(a)
        wf += max(len(s) for s in banks.values())
    return wf


def model(N, S1pad=2, S2=None, verbose=True):
    R3 = N // 64
    TL = N // 40
    NQ = 64 // TL                      # (k1,k2) pairs per lane in pass 3
    if S2 is None:
        S2 = R3 + 1
    S1 = N // 8 + S1pad
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    w = lambda M, e: np.exp(2j * np.pi * (e % M) / M)
    v = np.zeros((TL, 40), complex)
    # load
    for t in range(TL):
        for j in range(5):
            for n1 in range(8):
                v[t, j * 8 + n1] = x[n1 * (N // 8) + t + TL * j]
    # pass 1: radix-8 over n1 + twiddle w_N^{m k1}
    for t in range(TL):
        for j in range(5):
            m = t + TL * j
            blk = v[t, j * 8:(j + 1) * 8].copy()
            for k1 in range(8):
                v[t, j * 8 + k1] = sum(blk[n1] * w(8, n1 * k1) for n1 in range(8)) * w(N, m * k1)
    # exchange 1
    sm = np.zeros(8 * S1 + 64, complex)
    wf_w1 = wf_r1 = 0
    for j in range(5):
        for k1 in range(8):
            addrs = [k1 * S1 + t + TL * j for t in range(TL)]
            wf_w1 += conflicts(addrs)
            for t in range(TL):
                sm[addrs[t]] = v[t, j * 8 + k1]
    for jp in range(5):
        for n2 in range(8):
            addrs = []
            for t in range(TL):
                p = t + TL * jp
                n3, k1 = divmod(p, 8)
                addrs.append(k1 * S1 + n2 * R3 + n3)
            wf_r1 += conflicts(addrs)
            for t in range(TL):
                v[t, jp * 8 + n2] = sm[addrs[t]]
    # pass 2: radix-8 over n2 + twiddle w_{N/8}^{n3 k2}
    for t in range(TL):
        for jp in range(5):
            p = t + TL * jp
            n3, k1 = divmod(p, 8)
            blk = v[t, jp * 8:(jp + 1) * 8].copy()
            for k2 in range(8):
                v[t, jp * 8 + k2] = sum(blk[n2] * w(8, n2 * k2) for n2 in range(8)) * w(N // 8, n3 * k2)
    # exchange 2
    sm = np.zeros(64 * R3 + 64, complex)
    wf_w2 = wf_r2 = 0
    for jp in range(5):
        for k2 in range(8):
            addrs = []
            for t in range(TL):
                p = t + TL * jp
                n3, k1 = divmod(p, 8)
                addrs.append(64 * n3 + ((k1 + 8 * k2) ^ (8 * (n3 & 1))))
            wf_w2 += conflicts(addrs)
            for t in range(TL):
                sm[addrs[t]] = v[t, jp * 8 + k2]
    for u in range(NQ):
        for n3 in range(R3):
            addrs = [64 * n3 + ((t + TL * u) ^ (8 * (n3 & 1))) for t in range(TL)]
            wf_r2 += conflicts(addrs)
            for t in range(TL):
                v[t, u * R3 + n3] = sm[addrs[t]]
    # pass 3: radix-R3 over n3
    X = np.zeros(N, complex)
    for t in range(TL):
        for u in range(NQ):
            q = t + TL * u
            blk = v[t, u * R3:(u + 1) * R3].copy()
            for k3 in range(R3):
                X[q + 64 * k3] = sum(blk[n3] * w(R3, n3 * k3) for n3 in range(R3))
    ref = np.fft.ifft(x) * N
    err = np.abs(X - ref).max() / np.abs(ref).max()
    ideal = lambda n: n * max(1, TL // 16)
    if verbose:
        print(f'N={N} R3={R3} TL={TL} S1={S1} err={err:.2e} '
              f'wavefronts w1={wf_w1}/{ideal(40)} r1={wf_r1}/{ideal(40)} '
              f'w2={wf_w2}/{ideal(40)} r2={wf_r2}/{ideal(40)}')
    return err, (wf_w1, wf_r1, wf_w2, wf_r2)


if __name__ == '__main__':
    model(1280)
