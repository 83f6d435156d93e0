"""Prototype of the warp-level 128/64-point complex FFT index maps used by the
CUDA kernels (ns_fft.cuh): checks the maths against numpy and counts shared
memory bank conflicts for 64-bit accesses (half-warp phases, 32 banks x 4 B)."""
import numpy as np

def conflicts(idx_by_lane, active=None):
    """idx in float2 units; returns wavefronts needed (ideal 2 for 32 lanes)."""
    tot = 0
    for half in range(2):
        lanes = [l for l in range(16 * half, 16 * half + 16) if active is None or active[l]]
        banks = {}
        for l in lanes:
            b = (2 * idx_by_lane[l]) % 32
            banks.setdefault(b, set()).add(idx_by_lane[l])
        tot += max([len(v) for v in banks.values()] + [0])
    return tot

def fft_warp(z, sign, N, pad1, pad2, report):
    """z: N complex in natural order. returns Z natural order. emulates lanes."""
    L = N // 4                      # active lanes (32 or 16)
    W = lambda n, d: np.exp(sign * 2j * np.pi * n / d)
    w4 = W(1, 4)
    def r4(v):
        return [sum(v[n] * w4 ** (n * k) for n in range(4)) for k in range(4)]
    P1 = lambda k1, n0: k1 * (L + pad1) + n0
    S1 = np.zeros(4 * (L + pad1) + 64, complex)
    # pass 1
    st = [[0] * 32 for _ in range(4)]
    for lane in range(L):
        v = [z[L * n1 + lane] for n1 in range(4)]
        A = r4(v)
        for k1 in range(4):
            S1[P1(k1, lane)] = A[k1] * W(lane * k1, N)
            st[k1][lane] = P1(k1, lane)
    act = [l < L for l in range(32)]
    for k1 in range(4): report('p1 store', conflicts(st[k1], act))
    # pass 2: lane = (k1, m0); sub-DFT length L over n0 = (L/4) m1 + m0
    M = L // 4                      # 8 or 4
    P2 = lambda k1, j1, m0: (k1 * 4 + j1) * (M + pad2) + m0
    S2 = np.zeros(16 * (M + pad2) + 64, complex)
    ld = [[0] * 32 for _ in range(4)]; st = [[0] * 32 for _ in range(4)]
    for lane in range(L):
        k1, m0 = lane // M, lane % M
        v = [S1[P1(k1, M * m1 + m0)] for m1 in range(4)]
        for m1 in range(4): ld[m1][lane] = P1(k1, M * m1 + m0)
        C = r4(v)
        for j1 in range(4):
            S2[P2(k1, j1, m0)] = C[j1] * W(m0 * j1, L)
            st[j1][lane] = P2(k1, j1, m0)
    for r in range(4): report('p2 load', conflicts(ld[r], act)); report('p2 store', conflicts(st[r], act))
    Z = np.zeros(N, complex)
    ld = [[0] * 32 for _ in range(4)]
    outidx = [[0] * 32 for _ in range(4)]
    if N == 128:
        # pass 3: lane=(k1,j1,p0); radix-4 over p1 (m0 = 2 p1 + p0), twiddle W8^{p0 q1}, then radix-2 with lane^1
        D = np.zeros((32, 4), complex)
        for lane in range(32):
            k1, j1, p0 = lane >> 3, (lane >> 1) & 3, lane & 1
            v = [S2[P2(k1, j1, 2 * p1 + p0)] for p1 in range(4)]
            for p1 in range(4): ld[p1][lane] = P2(k1, j1, 2 * p1 + p0)
            d = r4(v)
            for q1 in range(4): D[lane, q1] = d[q1] * W(p0 * q1, 8)
        for lane in range(32):
            k1, j1, p0 = lane >> 3, (lane >> 1) & 3, lane & 1
            for q1 in range(4):
                other = D[lane ^ 1, q1]
                val = D[lane, q1] + other if p0 == 0 else other - D[lane, q1]
                k = k1 + 4 * j1 + 16 * q1 + 64 * p0
                Z[k] = val; outidx[q1][lane] = k
    else:
        for lane in range(16):
            k1, j1 = lane >> 2, lane & 3
            v = [S2[P2(k1, j1, m0)] for m0 in range(4)]
            for m0 in range(4): ld[m0][lane] = P2(k1, j1, m0)
            d = r4(v)
            for q1 in range(4):
                k = k1 + 4 * j1 + 16 * q1
                Z[k] = d[q1]; outidx[q1][lane] = k
    for r in range(4): report('p3 load', conflicts(ld[r], act))
    return Z, outidx

if __name__ == '__main__':
    rs = np.random.RandomState(1)
    for N in (128, 64):
        z = rs.randn(N) + 1j * rs.randn(N)
        best = None
        for pad1 in range(0, 9):
            for pad2 in range(0, 9):
                rep = {}
                def report(k, v): rep[k] = max(rep.get(k, 0), v)
                Z, outidx = fft_warp(z, +1, N, pad1, pad2, report)
                ref = np.fft.ifft(z) * N
                assert np.abs(Z - ref).max() < 1e-9
                score = sum(rep.values())
                if best is None or score < best[0]: best = (score, pad1, pad2, dict(rep))
        print(N, best)
        # output store patterns: Zs[k] natural with pad function k + (k>>5)*padz
        Z, outidx = fft_warp(z, +1, N, best[1], best[2], lambda k, v: None)
        for padz in range(0, 5):
            act = [l < N // 4 for l in range(32)]
            print(' padz', padz, [conflicts([k + (k >> 4) * padz for k in outidx[q]], act) for q in range(4)])
