"""Prototype of the mixed-radix Stockham autosort FFT used by csrc/local_energy.cu (index formulas, twiddle tables,
shared-memory bank-conflict model).  Natural order in, natural order out, radices 8/4/2, passes with Ns = product of
the radices already done:  butterfly j reads x[j + r n/R], multiplies by exp(-+2 pi i (j % Ns) r / (Ns R)), does an
R-point DFT and writes y[(j // Ns) Ns R + (j % Ns) + r Ns]."""
import numpy as np


def radices(lg):
    r = [8] * (lg // 3)
    if lg % 3 == 1: r.append(2)
    if lg % 3 == 2: r.append(4)
    return r


def fft_stockham(x, inverse=False):
    n = len(x); lg = n.bit_length() - 1
    sgn = 1.0 if inverse else -1.0
    Ns = 1
    x = x.astype(np.complex128)
    for R in radices(lg):
        y = np.empty_like(x)
        for j in range(n // R):
            v = np.array([x[j + r * (n // R)] * np.exp(sgn * 2j * np.pi * (j % Ns) * r / (Ns * R)) for r in range(R)])
            v = np.array([sum(v[r] * np.exp(sgn * 2j * np.pi * r * k / R) for r in range(R)) for k in range(R)])
            d = (j // Ns) * Ns * R + (j % Ns)
            for r in range(R):
                y[d + r * Ns] = v[r]
        x = y; Ns *= R
    return x


def wavefronts(idx):
    """64-bit shared-memory accesses of one warp: two half-warps, each needs max multiplicity of (word-pair bank)"""
    tot = 0
    for h in range(0, 32, 16):
        a = sorted(set(idx[h:h + 16]))
        banks = {}
        for i in a: banks[i % 16] = banks.get(i % 16, 0) + 1
        tot += max(banks.values())
    return tot


def conflict_model(n, pad):
    lg = n.bit_length() - 1
    Ns = 1
    for R in radices(lg):
        nb = n // R
        rd = wr = 0; cnt = 0
        for w0 in range(0, max(nb, 32), 32):
            lanes = [(w0 + l) % nb + ((w0 + l) // nb) * 0 for l in range(32)]
            for r in range(R):
                rd += wavefronts([pad(j + r * nb) for j in lanes])
                wr += wavefronts([pad((j // Ns) * Ns * R + (j % Ns) + r * Ns) for j in lanes])
                cnt += 2
        print(f"  n={n} R={R} Ns={Ns}: read {rd / cnt:.2f}x write {wr / cnt:.2f}x of the minimum")
        Ns *= R


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for lg in range(3, 12):
        n = 1 << lg
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        e1 = np.abs(fft_stockham(x) - np.fft.fft(x)).max()
        e2 = np.abs(fft_stockham(x, True) - np.fft.ifft(x) * n).max()
        print(n, radices(lg), e1, e2)
        if lg > 8: break
    for name, pad in (("none", lambda i: i), ("i + i/8", lambda i: i + (i >> 3)), ("i + i/16", lambda i: i + (i >> 4)),
                      ("i + i/32", lambda i: i + (i >> 5)), ("i+i/8+i/64", lambda i: i + (i >> 3) + (i >> 6))):
        print("pad", name)
        for n in (512, 1024, 256):
            conflict_model(n, pad)


def seqfast_conflicts():
    """le_cols: first pass writes / last pass reads with the sequence index fastest across lanes (global accesses of a
    tile of columns are then contiguous).  Which pitch (H + pad) keeps those shared-memory accesses conflict-free?"""
    swz = lambda i: i ^ ((i >> 3) & 15)
    for n in (64, 256, 512, 1024, 2048):
        lg = n.bit_length() - 1
        spp = 2048 // n
        rs = radices(lg)
        R0, RL = rs[0], rs[-1]
        NsL = n // RL
        for pad in (0, 1, 2, 4, 6, 8, 10):
            pitch = n + pad
            w = r_ = cnt = 0
            for w0 in range(0, 2048 // R0, 32):
                lanes = [((w0 + l) & (spp - 1), (w0 + l) // spp) for l in range(32)]
                for r in range(R0):      # first pass (Ns = 1) writes d + r, d = j * R0
                    w += wavefronts([q * pitch + swz(j * R0 + r) for q, j in lanes]); cnt += 2
            cnt2 = 0
            for w0 in range(0, 2048 // RL, 32):
                lanes = [((w0 + l) & (spp - 1), (w0 + l) // spp) for l in range(32)]
                for r in range(RL):      # last pass reads j + r n / RL
                    r_ += wavefronts([q * pitch + swz(j + r * NsL) for q, j in lanes]); cnt2 += 2
            print(f"  n={n} spp={spp} pad={pad}: first-pass writes {w / cnt:.2f}x  last-pass reads {r_ / cnt2:.2f}x")


if __name__ == "__main__":
    seqfast_conflicts()
