"""The lane -> state-index algebra of the factored Jacobian kernels (csrc/kernels_factoredj.cuh), restated in numpy and checked on the
CPU: with the DMMA m8n8k4 fragment layouts (A: lane holds A[lane >> 2][lane & 3]; B: B[lane & 3][lane >> 2]; D: D[lane >> 2][2 (lane & 3) + {0, 1}])
the chain step  e'(a', r) = sum_a F[a][a'] e(a, r)  and the accumulate  acc[a][b] += sum_r e(a, r) s(b, r)  as the kernels index them
reproduce the dense embedded operation for every target-qubit combination at d = 64 and d = 256, every output element is written, and the
shared-memory swizzle keeps the 16 lanes of a half-warp on 16 different 8-byte banks for the fragment loads.  (The CUDA kernels are
tested on the GPU: tests/test_gpu_synthetic.py::test_factored_jacobian_matches_oracle, tests/test_gpu_parity.py::test_factored_jacobian_golden_c3.)"""
import itertools

import numpy as np
import pytest


def insert2(v, sh):
    return ((v >> sh) << (sh + 2)) | (v & ((1 << sh) - 1))


def sw(i):                                                   # fj_sw
    d2, d3 = (i >> 4) & 3, (i >> 6) & 3
    c3 = (d3 >> 1) | (((d3 ^ (d3 >> 1)) & 1) << 1)
    return i ^ (d2 ^ c3) ^ ((d2 ^ d3) << 2)


def idx1(a, r, sh):
    return sw(insert2(r, sh) | (a << sh))


def idx2(a, r, lo, hi, s0, s1):
    return sw(insert2(insert2(r, lo), hi) | ((a >> 2) << s0) | ((a & 3) << s1))


def dense_embed(small, shifts, D):
    idx = np.arange(D); tm = 0
    for s in shifts:
        tm |= 3 << s
    t = np.zeros(D, int)
    for s in shifts:
        t = (t << 2) | ((idx >> s) & 3)
    rest = idx & ~tm
    F = np.where(rest[:, None] == rest[None, :], small[t[:, None], t[None, :]], 0.0)
    return F, t, rest


LANES = np.arange(32); LG = LANES >> 2; LT = LANES & 3


def _cases():
    for D, nq in ((64, 3), (256, 4)):
        for tg in itertools.permutations(range(nq), 2):
            yield D, nq, tg
        for q in range(nq):
            yield D, nq, (q,)


@pytest.mark.parametrize("D,nq,targets", list(_cases()))
def test_fragment_indices_reproduce_the_embedded_factor(D, nq, targets):
    rng = np.random.default_rng(D + 10 * len(targets) + sum(targets))
    shifts = [2 * (nq - 1 - q) for q in targets]
    ds = 4 ** len(targets)
    small = rng.standard_normal((ds, ds)); e = rng.standard_normal(D); s = rng.standard_normal(D)
    F, t, rest = dense_embed(small, shifts, D)
    acc_ref = np.zeros((ds, ds))
    np.add.at(acc_ref, (t[:, None].repeat(D, 1), t[None, :].repeat(D, 0)), np.outer(e, s) * (rest[:, None] == rest[None, :]))
    e_ref = F.T @ e
    cur = np.zeros(D); sb = np.zeros(D)
    cur[[sw(i) for i in range(D)]] = e; sb[[sw(i) for i in range(D)]] = s
    nxt = np.full(D, np.nan)
    half_warp_banks = []                                     # (index list) of every fragment load, per half-warp
    if len(targets) == 2:
        R = D // 16; s0, s1 = shifts; lo, hi = min(s0, s1), max(s0, s1)
        acc = np.zeros((16, 16))
        for kk in range(R // 4):
            A = [np.zeros((8, 4)) for _ in range(2)]; B = [np.zeros((4, 8)) for _ in range(2)]
            for mt in range(2):
                ix = [idx2(LG[l] + 8 * mt, 4 * kk + LT[l], lo, hi, s0, s1) for l in range(32)]
                half_warp_banks += [ix[:16], ix[16:]]
                for l in range(32):
                    A[mt][LG[l], LT[l]] = cur[ix[l]]; B[mt][LT[l], LG[l]] = sb[ix[l]]
            for mt in range(2):
                for nt in range(2):
                    acc[8 * mt:8 * mt + 8, 8 * nt:8 * nt + 8] += A[mt] @ B[nt]
        Afr = [[np.array([[small[4 * kk + k, 8 * mt + m] for k in range(4)] for m in range(8)]) for kk in range(4)] for mt in range(2)]
        for nt in range((R + 7) // 8):
            o = [np.zeros((8, 8)), np.zeros((8, 8))]
            for kk in range(4):
                ix = [idx2(4 * kk + LT[l], (8 * nt + LG[l]) & (R - 1), lo, hi, s0, s1) for l in range(32)]
                half_warp_banks += [ix[:16], ix[16:]] if R >= 8 else [ix[:16]]
                Bf = np.zeros((4, 8))
                for l in range(32):
                    Bf[LT[l], LG[l]] = cur[ix[l]]
                for mt in range(2):
                    o[mt] += Afr[mt][kk] @ Bf
            for l in range(32):
                rO = 8 * nt + 2 * LT[l]
                if rO < R:
                    for mt in range(2):
                        for j in range(2):
                            nxt[idx2(LG[l] + 8 * mt, rO + j, lo, hi, s0, s1)] = o[mt][LG[l], 2 * LT[l] + j]
    else:
        R = D // 4; sh = shifts[0]
        cc = np.zeros((8, 8))
        for kk in range(R // 4):
            ix = [idx1(LG[l] & 3, 4 * kk + LT[l], sh) for l in range(32)]
            half_warp_banks += [ix[:16]]
            A = np.zeros((8, 4)); B = np.zeros((4, 8))
            for l in range(32):
                A[LG[l], LT[l]] = cur[ix[l]]; B[LT[l], LG[l]] = sb[ix[l]]
            cc += A @ B
        acc = cc[:4, :4]
        Af = np.array([[small[k, m & 3] for k in range(4)] for m in range(8)])
        for nt in range(R // 8):
            ix = [idx1(LT[l], 8 * nt + LG[l], sh) for l in range(32)]
            half_warp_banks += [ix[:16], ix[16:]]
            Bf = np.zeros((4, 8))
            for l in range(32):
                Bf[LT[l], LG[l]] = cur[ix[l]]
            o = Af @ Bf
            for l in range(32):
                if LG[l] < 4:
                    for j in range(2):
                        nxt[idx1(LG[l], 8 * nt + 2 * LT[l] + j, sh)] = o[LG[l], 2 * LT[l] + j]
    e_new = nxt[[sw(i) for i in range(D)]]
    assert not np.isnan(e_new).any()                         # every element of the new backward vector is written exactly where it is read
    assert np.max(np.abs(acc - acc_ref)) <= 1e-13 and np.max(np.abs(e_new - e_ref)) <= 1e-13
    for ix in half_warp_banks:                               # 16 lanes -> 16 different 8-byte banks (16 banks of 8 bytes per 128-byte row)
        assert len({i % 16 for i in ix}) == 16, (targets, ix)
