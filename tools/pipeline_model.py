"""Numpy model of the two-pass real-even transforms in csrc/psfr_passes.cu / psfr_hot.cu
(index shifts, signs, Hermitian extension, transposed orientation), checked against the
oracle's psd_to_psf at a small grid size.  Design aid, not product code."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import psfr_oracle as orc

N = 64
NH = N // 2
rng = np.random.default_rng(1)


def line_fft(x):            # sign +i, unnormalised (what warp_fft computes)
    return np.fft.ifft(x) * N


def pass_rows_even(P):
    """LoadEvenRows + StoreTransposedPair: Bt[y][a] for a in [0, NH]."""
    idx = (-np.arange(N)) % N
    E = 0.5 * (P + P[idx][:, idx])
    Bt = np.zeros((N, NH + 2), complex)
    for a in range(NH + 1):
        Bt[:, a] = line_fft(E[a])
    return Bt


def pass_cols_hermitian(Bt, nrows_out, last_valid, scale):
    """LoadHermitianPair + StoreRealRows: out[o][b], o = output row, b = (x + N/2) % N."""
    out = np.zeros((nrows_out, N))
    for o in range(last_valid + 1):
        col = Bt[(o + NH) % N]
        V = np.zeros(N, complex)
        V[:NH + 1] = col[:NH + 1]
        V[NH + 1:] = np.conj(col[1:NH][::-1])        # V[n] = conj(col[N-n])
        Z = line_fft(V)
        assert np.abs(Z.imag).max() < 1e-9 * np.abs(Z).max()
        for b in range(N):
            out[o, b] = scale * (-1) ** (o + b) * Z[(b + NH) % N].real
    return out


def structure_function(P):
    Bt = pass_rows_even(P)
    raw = pass_cols_hermitian(Bt, NH + 2, NH, 2.0 / 16 ** 2)
    D = raw[NH, NH] - raw
    D[NH + 1] = 0
    return D                                        # Dt[alpha][beta] = Dphi_c[beta][alpha]


def telescope_otf_half():
    pup = orc.pupil_mask(N / 4, N / 2, 0.14)
    T = orc.telescope_otf(pup, N)                   # centred, symmetric
    Th = np.zeros((NH + 2, N))
    Th[:NH + 1] = T.T[:NH + 1]
    return Th, pup


def full_psf(D, T, c):
    O = np.exp(-c * D) * T                          # [NH+2][N] transposed half-plane
    Bt = np.zeros((N, NH + 2), complex)
    for a in range(NH + 1):
        Bt[:, a] = line_fft(O[a])
    return pass_cols_hermitian(Bt, N, N - 1, 1.0)


P = rng.uniform(0, 1, (N, N)) * 1e3
P[NH - 5:NH + 5, NH - 5:NH + 5] += 1e5
D = structure_function(P)
ref = orc.structure_function_unit(P, L=16)
print('D_unit err', np.abs(D[:NH + 1] - ref.T[:NH + 1]).max() / ref.max())
T, pup = telescope_otf_half()
lb = 700e-9
c = 0.5 * (2 * np.pi / 700.) ** 2
# scale the PSD so that exp(-c D) is not degenerate
psf = full_psf(D * 1e-4, T, c)
# oracle on the same scaled structure function: rebuild through its formula
otf = np.exp(-c * ref * 1e-4) * orc.telescope_otf(pup, N)
pr = np.real(np.fft.fftshift(np.fft.ifft2(np.fft.fftshift(otf))))
pr /= pr.sum()
print('PSF err', np.abs(psf - pr).max() / pr.max(), 'sum', psf.sum())
