// kernels_levelj.cuh -- analytic Jacobian for large superoperators (d = 64, 256): level-batched forward AND backward
// sweeps on the FP64 tensor cores, then a sparse contraction with the member-derivative map.  This is the dprobs
// path of BASELINE configs 3 (3 qubits, d = 64, 50 k random circuits) and 5 (4 qubits, d = 256).
//
// Same mathematics as the d = 16 kernels (== MatrixForwardSimulator._dprobs_from_rho_e,
// pygsti/forwardsims/matrixforwardsim.py:1059-1139):
//     s_0 = rho, s_{k+1} = G_{ops[k]} s_k;      e^(L) = E_j, e^(k) = G_{ops[k]}^T e^(k+1)
//     J[el, p] = sum_k sum_{(i,j) in nz(dG_{ops[k]}/dp)} v e^(k+1)[i] s_k[j]  + sum_i dRho_i/dp e^(0)[i] + sum_i dE_i/dp s_L[i]
// but organised for matrices that do not fit in shared memory:
//   phase F  k_level_gemm_rows<D> (forward):  one launch per depth level; the circuits whose layer at that level is
//            gate g form a dense product (tiles of 32 rows) exactly as in kernels_level.cuh, except that EVERY level is
//            kept: state table FS[fbase[c] + k][D].
//   phase B  the same kernel on the transposed gates: rows are (circuit, outcome) pairs, level m handles step
//            k = L-1-m of every circuit with L > m: table BH[bbase[c] + o (L+1) + k][D].
//   phase C  k_level_accum<D>: one CTA per (circuit, tile of PT parameters).  Gate by gate, the circuit's steps that
//            apply that gate are staged in shared memory (s_k and e^(k+1) of every outcome); each thread owns a
//            (parameter, outcome) pair and walks the non-zeros of that parameter's column of D inside the gate's block.
//            Parameters shared by several gates (e.g. one `Gxpi2` on three qubits) accumulate in shared memory, so every
//            Jacobian entry is written exactly once, in a fixed order (deterministic), coalesced along p, with the
//            objective-function row scale applied in the same store.
// The W matrix (n_elements x n_ops d^2) of the correctness-first generic path is never formed.
#pragma once
#include "common.cuh"
#include "kernels_level.cuh"

#define LJ_PT_MIN 256      // parameters per accumulate tile: chosen per atom (engine.cu, set_derivs), a power of two in [MIN, MAX]
#define LJ_PT_MAX 2048
#define LJ_THREADS 128

struct LevelJDev {
    const uint32_t* fbase;     // [n_circ] first FS row of the circuit (rows fbase + 0..L)
    const uint32_t* bbase;     // [n_circ] first BH row (rows bbase + o*(L+1) + 0..L)
    const uint16_t* bperm;     // [n_prop] per circuit: step indices sorted by gate
    const uint16_t* bcnt;      // [n_circ][n_ops] bucket sizes
    const uint32_t* ti_ptr;    // [n_tiles][n_ops + 2] item ranges per (tile, gate); slot n_ops = SPAM rows
    const uint4* items;        // (p_local, nz lo, nz hi, 0) into crow / cval
    const int32_t* crow;       // CSC rows of D (W-space index)
    const double* cval;
    double* FS; double* BH;
    int n_tiles, n_params, no_max, ts, pt;   // pt: parameters per accumulate tile
};

// FS[fbase[c]] = rho[prep_c];  BH[bbase[c] + o (L+1) + L] = E[eff_o]
template <int D>
__global__ void __launch_bounds__(128)
k_levelj_init(AtomDev a, ModelDev m, LevelJDev lj)
{
    const double* rho = m.M + m.off_rho;
    const double* E = m.M + m.off_eff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * 4 + warp; c < a.n_circ; c += gridDim.x * 4) {
        const uint32_t L = a.circ_ptr[c + 1] - a.circ_ptr[c];
        double* f = lj.FS + (size_t)lj.fbase[c] * D;
        const double* r = rho + (size_t)a.circ_prep[c] * D;
        for (int i = lane; i < D; i += 32) f[i] = r[i];
        const int q0 = a.out_ptr[c], nout = a.out_ptr[c + 1] - q0;
        for (int o = 0; o < nout; ++o) {
            double* h = lj.BH + ((size_t)lj.bbase[c] + (size_t)o * (L + 1) + L) * D;
            const double* e = E + (size_t)a.out_eff[q0 + o] * D;
            for (int i = lane; i < D; i += 32) h[i] = e[i];
        }
    }
}

// One level of a sweep: T[rows[e] + delta] = Gm_g . T[rows[e]]  for the 32 rows of a tile (Gm = G forward, G^T backward).
// dynamic smem: 32 * (D + 4) doubles
template <int D>
__global__ void __launch_bounds__(128)
k_level_gemm_rows(const double* __restrict__ Gm, const LevelTile* __restrict__ tiles, const uint32_t* __restrict__ rows,
                  double* __restrict__ T, int delta)
{
    constexpr int LDS_ = D + 4;
    constexpr int NT = 2;                       // a CTA owns 64 output components (blockIdx.y selects which)
    extern __shared__ __align__(16) double st[];
    const LevelTile tl = tiles[blockIdx.x];
    const uint32_t* rw = rows + tl.first;
    const int cnt = tl.count;
    const double* Gg = Gm + (size_t)tl.gate * D * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < 32 * (D / 2); idx += blockDim.x) {
        const int r = idx / (D / 2), j2 = idx - r * (D / 2);
        double2 v = make_double2(0.0, 0.0);
        if (r < cnt) v = *reinterpret_cast<const double2*>(T + (size_t)rw[r] * D + 2 * j2);
        *reinterpret_cast<double2*>(st + r * LDS_ + 2 * j2) = v;
    }
    __syncthreads();
    const int mrow = lane >> 2, q = lane & 3;
    const int n0 = blockIdx.y * 64 + warp * 16;
    double acc[4][NT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const double* bp = Gg + (size_t)(n0 + mrow) * D + q;
    const double* ap = st + mrow * LDS_ + q;
    level_gemm_core<D, NT>(ap, bp, acc);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int r = mt * 8 + mrow;
        if (r < cnt) {
            double* op = T + (size_t)((int64_t)rw[r] + delta) * D + n0 + 2 * q;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(op + nt * 8) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
    }
}

// probabilities from the state table: p_el = E_e . s_L  (one warp per circuit)
template <int D>
__global__ void __launch_bounds__(128)
k_levelj_probs(AtomDev a, ModelDev m, LevelJDev lj, double* __restrict__ out)
{
    const double* E = m.M + m.off_eff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * 4 + warp; c < a.n_circ; c += gridDim.x * 4) {
        const uint32_t L = a.circ_ptr[c + 1] - a.circ_ptr[c];
        const double* s = lj.FS + ((size_t)lj.fbase[c] + L) * D;
        for (int qo = a.out_ptr[c]; qo < a.out_ptr[c + 1]; ++qo) {
            const double* e = E + (size_t)a.out_eff[qo] * D;
            double part = 0.0;
#pragma unroll
            for (int i = lane; i < D; i += 32) part += e[i] * s[i];
#pragma unroll
            for (int mk = 16; mk > 0; mk >>= 1) part += shfl_xor_f64(part, mk);
            if (lane == 0) out[a.out_el[qo]] = part;
        }
    }
}

// phase C.  dynamic smem (doubles): no_max * LJ_PT  (accumulators)  +  ts * D  (states)  +  no_max * (ts * D + 1)  (adjoints)
template <int D>
__global__ void __launch_bounds__(LJ_THREADS)
k_level_accum(AtomDev a, ModelDev m, LevelJDev lj, double* __restrict__ J, int64_t ld, const double* __restrict__ row_scale)
{
    extern __shared__ __align__(16) double sml[];
    const int TS = lj.ts;
    const int LJ_PT = lj.pt;
    const int ES = TS * D + 1;                       // outcome stride of the adjoint stage (odd: consecutive outcomes -> different banks)
    double* Jacc = sml;                              // [no_max][LJ_PT]
    double* s_st = Jacc + (size_t)lj.no_max * LJ_PT; // [TS][D]
    double* e_st = s_st + (size_t)TS * D;            // [no_max][ES]
    const int c = blockIdx.x / lj.n_tiles, tile = blockIdx.x - c * lj.n_tiles;
    const int tid = threadIdx.x;
    const int q0 = a.out_ptr[c], nout = a.out_ptr[c + 1] - q0;
    const uint32_t p0 = a.circ_ptr[c], Lc = a.circ_ptr[c + 1] - p0;
    const size_t fb = lj.fbase[c], bb = lj.bbase[c];
    const int tile_p0 = tile * LJ_PT;
    const int tile_np = (lj.n_params - tile_p0 < LJ_PT) ? lj.n_params - tile_p0 : LJ_PT;
    const uint32_t* tp = lj.ti_ptr + (size_t)tile * (a.n_ops + 2);
    const uint16_t* cn = lj.bcnt + (size_t)c * a.n_ops;
    const uint16_t* perm = lj.bperm + p0;

    for (int idx = tid; idx < nout * LJ_PT; idx += LJ_THREADS) Jacc[idx] = 0.0;
    uint32_t tb = 0;
    for (int g = 0; g < a.n_ops; ++g) {
        const int cnt = cn[g];
        const uint32_t it0 = tp[g], it1 = tp[g + 1];
        if (cnt > 0 && it1 > it0) {
            const int gbase = g * D * D;
            const int nwork = (int)(it1 - it0) * nout;
            for (int s0 = 0; s0 < cnt; s0 += TS) {
                const int ns = (cnt - s0 < TS) ? cnt - s0 : TS;
                __syncthreads();                                  // consumers of the previous stage are done
                for (int idx = tid; idx < (1 + nout) * ns * D; idx += LJ_THREADS) {
                    const int row = idx / D, col = idx - row * D;
                    if (row < ns) {
                        const uint32_t k = perm[tb + s0 + row];
                        s_st[row * D + col] = lj.FS[(fb + k) * D + col];
                    } else {
                        const int r2 = row - ns, o = r2 / ns, ts = r2 - o * ns;
                        const uint32_t k = perm[tb + s0 + ts];
                        e_st[o * ES + ts * D + col] = lj.BH[(bb + (size_t)o * (Lc + 1) + k + 1) * D + col];
                    }
                }
                __syncthreads();
                for (int w = tid; w < nwork; w += LJ_THREADS) {
                    const int it = w / nout, o = w - it * nout;
                    const uint4 item = __ldg(lj.items + it0 + it);
                    const double* eo = e_st + o * ES;
                    double acc = 0.0;
                    for (uint32_t t = item.y; t < item.z; ++t) {
                        const int wl = __ldg(lj.crow + t) - gbase;
                        const int i = wl / D, j = wl - i * D;
                        double tacc = 0.0;
                        for (int ts = 0; ts < ns; ++ts) tacc = fma(eo[ts * D + i], s_st[ts * D + j], tacc);
                        acc = fma(__ldg(lj.cval + t), tacc, acc);
                    }
                    Jacc[o * LJ_PT + item.x] += acc;
                }
            }
        }
        tb += cnt;
    }
    __syncthreads();
    {   // state-preparation and effect rows of D
        const uint32_t it0 = tp[a.n_ops], it1 = tp[a.n_ops + 1];
        const int nwork = (int)(it1 - it0) * nout;
        const int prep = a.circ_prep[c];
        const double* sL = lj.FS + (fb + Lc) * D;
        for (int w = tid; w < nwork; w += LJ_THREADS) {
            const int it = w / nout, o = w - it * nout;
            const uint4 item = __ldg(lj.items + it0 + it);
            const int eff = a.out_eff[q0 + o];
            const double* e0 = lj.BH + (bb + (size_t)o * (Lc + 1)) * D;
            double acc = 0.0;
            for (uint32_t t = item.y; t < item.z; ++t) {
                const int64_t wl = __ldg(lj.crow + t);
                if (wl < m.off_eff) {
                    const int r = (int)((wl - m.off_rho) / D), i = (int)((wl - m.off_rho) - (int64_t)r * D);
                    if (r == prep) acc = fma(__ldg(lj.cval + t), e0[i], acc);
                } else {
                    const int r = (int)((wl - m.off_eff) / D), i = (int)((wl - m.off_eff) - (int64_t)r * D);
                    if (r == eff) acc = fma(__ldg(lj.cval + t), sL[i], acc);
                }
            }
            Jacc[o * LJ_PT + item.x] += acc;
        }
    }
    __syncthreads();
    for (int o = 0; o < nout; ++o) {
        const int64_t el = a.out_el[q0 + o];
        const double sc = row_scale ? __ldg(row_scale + el) : 1.0;
        double* Jr = J + el * ld + tile_p0;
        for (int p = tid; p < tile_np; p += LJ_THREADS) __stcs(Jr + p, Jacc[o * LJ_PT + p] * sc);
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase C, version 2 (default): the outer products are formed on the FP64 tensor cores.
//   W_g^{(o)}[i][j] = sum_{t in bucket(c, g)} e^{(o)}_{t+1}[i] s_t[j]   (M = i, N = j, K = the bucket's steps, DMMA m8n8k4)
// is accumulated in registers for one 64 x 64 sub-block of the d x d matrix at a time (d = 64: the whole matrix), ONLY for
// the 8 x 8 tiles that contain a non-zero of dG_g/dtheta for a parameter of this CTA's tile (64-bit tile mask built on
// the host: an embedded 1-qubit gate needs 8-32 of the 64 tiles, a fully parameterised gate only the rows of its tile),
// parked in shared memory, and contracted with the sparse derivative map by one thread per parameter.
// A / B fragments are gathered straight from the adjoint / state tables (8-byte loads through L1: the four warps share
// the state rows); buckets are padded to a multiple of 4 steps with an all-zero table row.
// One CTA per (circuit, tile of LJ_PT parameters); per gate, per outcome: DMMA -> smem tile -> sparse contraction.
// dynamic smem (doubles): no_max * LJ_PT (accumulators) + 64 * LJ_LDW (W tile) + LJ_KMAX ints (bucket step list)
// ------------------------------------------------------------------------------------------------------------
#define LJ_LDW 66
#define LJ_KMAX 256

struct LevelJ2Dev {
    const uint32_t* ti_ptr2;   // [n_tiles][n_ops][nsb + 1] item ranges per (tile, gate, 64x64 sub-block)
    const uint64_t* mask2;     // [n_tiles][n_ops][nsb] needed 8x8 tiles of the sub-block (bit 8*mt + nt)
    const uint4* items2;       // (p_local, nz lo, nz hi, 0) into nz_ij / nz_v
    const uint16_t* nz_ij;     // (i_local << 8) | j_local inside the sub-block
    const double* nz_v;
    uint32_t zrow_f, zrow_b;   // all-zero rows of FS / BH
    int nsb;                   // sub-blocks per matrix = (D / 64)^2
};

// (forcing 128 registers / 4 CTAs per SM was measured slower, 88.6 vs 84.9 ms at BASELINE config 3: the spills cost more than the warps gain)
template <int D>
__global__ void __launch_bounds__(LJ_THREADS)
k_level_accum2(AtomDev a, ModelDev m, LevelJDev lj, LevelJ2Dev l2, double* __restrict__ J, int64_t ld,
               const double* __restrict__ row_scale)
{
    constexpr int SBD = D / 64;                       // sub-blocks per dimension
    const int LJ_PT = lj.pt;
    extern __shared__ __align__(16) double sml[];
    double* Jacc = sml;                               // [no_max][LJ_PT]
    double* Wt = Jacc + (size_t)lj.no_max * LJ_PT;    // [64][LJ_LDW]
    int* kidx = reinterpret_cast<int*>(Wt + 64 * LJ_LDW);   // [LJ_KMAX] step index, or -1 = padding
    const int c = blockIdx.x / lj.n_tiles, tile = blockIdx.x - c * lj.n_tiles;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mrow = lane >> 2, q = lane & 3;
    const int q0 = a.out_ptr[c], nout = a.out_ptr[c + 1] - q0;
    const uint32_t p0 = a.circ_ptr[c], Lc = a.circ_ptr[c + 1] - p0;
    const size_t fb = lj.fbase[c], bb = lj.bbase[c];
    const int tile_p0 = tile * LJ_PT;
    const int tile_np = (lj.n_params - tile_p0 < LJ_PT) ? lj.n_params - tile_p0 : LJ_PT;
    const uint16_t* cn = lj.bcnt + (size_t)c * a.n_ops;
    const uint16_t* perm = lj.bperm + p0;
    const uint32_t* tp2 = l2.ti_ptr2 + (size_t)tile * a.n_ops * (l2.nsb + 1);
    const uint64_t* mk2 = l2.mask2 + (size_t)tile * a.n_ops * l2.nsb;

    for (int idx = tid; idx < nout * LJ_PT; idx += LJ_THREADS) Jacc[idx] = 0.0;
    uint32_t tb = 0;
    for (int g = 0; g < a.n_ops; ++g) {
        const int cnt = cn[g];
        const uint32_t* tpg = tp2 + (size_t)g * (l2.nsb + 1);
        if (cnt > 0 && tpg[l2.nsb] > tpg[0]) {
            for (int s0 = 0; s0 < cnt; s0 += LJ_KMAX) {           // (buckets longer than LJ_KMAX steps: several passes)
                const int ns = (cnt - s0 < LJ_KMAX) ? cnt - s0 : LJ_KMAX;
                const int ns4 = (ns + 3) & ~3;
                __syncthreads();
                for (int t = tid; t < ns4; t += LJ_THREADS) kidx[t] = (t < ns) ? (int)perm[tb + s0 + t] : -1;
                __syncthreads();
                for (int sb = 0; sb < l2.nsb; ++sb) {
                    const uint32_t it0 = tpg[sb], it1 = tpg[sb + 1];
                    if (it1 == it0) continue;
                    const uint64_t mask = mk2[(size_t)g * l2.nsb + sb];
                    const unsigned wm = (unsigned)(mask >> (16 * warp)) & 0xffffu;     // this warp: m-tiles 2w, 2w+1
                    const unsigned need_n = (wm | (wm >> 8)) & 0xffu;
                    const int ib = (sb / SBD) * 64, jb = (sb % SBD) * 64;
                    for (int o = 0; o < nout; ++o) {
                        double acc[2][8][2];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
                        if (wm) {
                            // fragments of the next 4 steps are fetched while the current ones are multiplied
                            const size_t brow0 = bb + (size_t)o * (Lc + 1) + 1;
                            double a0, a1, b[8], na0, na1, nb[8];
                            auto fetch = [&](int k0, double& fa0, double& fa1, double (&fb_)[8]) {
                                const int k = kidx[k0 + q];
                                const double* er = lj.BH + (k >= 0 ? (brow0 + k) : (size_t)l2.zrow_b) * D + ib + warp * 16 + mrow;
                                const double* sr = lj.FS + (k >= 0 ? (fb + k) : (size_t)l2.zrow_f) * D + jb + mrow;
                                fa0 = __ldg(er); fa1 = __ldg(er + 8);
#pragma unroll
                                for (int nt = 0; nt < 8; ++nt) fb_[nt] = (need_n & (1u << nt)) ? __ldg(sr + nt * 8) : 0.0;
                            };
                            fetch(0, a0, a1, b);
                            for (int k0 = 0; k0 < ns4; k0 += 4) {
                                if (k0 + 4 < ns4) fetch(k0 + 4, na0, na1, nb);
#pragma unroll
                                for (int nt = 0; nt < 8; ++nt) {
                                    if (wm & (1u << nt)) dmma884(acc[0][nt][0], acc[0][nt][1], a0, b[nt]);          // warp-uniform
                                    if (wm & (1u << (8 + nt))) dmma884(acc[1][nt][0], acc[1][nt][1], a1, b[nt]);
                                }
                                a0 = na0; a1 = na1;
#pragma unroll
                                for (int nt = 0; nt < 8; ++nt) b[nt] = nb[nt];
                            }
                        }
                        __syncthreads();                                        // previous contraction has read Wt
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt)
                                if (wm & (1u << (8 * mt + nt)))
                                    *reinterpret_cast<double2*>(Wt + (warp * 16 + mt * 8 + mrow) * LJ_LDW + nt * 8 + 2 * q) =
                                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                        __syncthreads();
                        // four adjacent lanes share one parameter (non-zeros lo + sub, lo + sub + 4, ...); a warp takes 8 per round
                        const uint32_t n_it = it1 - it0;
                        for (uint32_t base = warp * 8; base < n_it; base += (LJ_THREADS / 32) * 8) {     // warp-uniform trip count
                            const bool ok = base + (lane >> 2) < n_it;
                            double s = 0.0; uint32_t pl = 0;
                            if (ok) {
                                const uint4 item = __ldg(l2.items2 + it0 + base + (lane >> 2));
                                pl = item.x;
                                for (uint32_t t = item.y + (lane & 3); t < item.z; t += 4) {
                                    const unsigned ij = __ldg(l2.nz_ij + t);
                                    s = fma(__ldg(l2.nz_v + t), Wt[(ij >> 8) * LJ_LDW + (ij & 0xffu)], s);
                                }
                            }
                            s += shfl_xor_f64(s, 1); s += shfl_xor_f64(s, 2);
                            if (ok && (lane & 3) == 0) Jacc[o * LJ_PT + pl] += s;
                        }
                    }
                }
            }
        }
        tb += cnt;
    }
    __syncthreads();
    {   // state-preparation and effect rows of D (same as version 1)
        const uint32_t* tp = lj.ti_ptr + (size_t)tile * (a.n_ops + 2);
        const uint32_t it0 = tp[a.n_ops], it1 = tp[a.n_ops + 1];
        const int nwork = (int)(it1 - it0) * nout;
        const int prep = a.circ_prep[c];
        const double* sL = lj.FS + (fb + Lc) * D;
        for (int w = tid; w < nwork; w += LJ_THREADS) {
            const int it = w / nout, o = w - it * nout;
            const uint4 item = __ldg(lj.items + it0 + it);
            const int eff = a.out_eff[q0 + o];
            const double* e0 = lj.BH + (bb + (size_t)o * (Lc + 1)) * D;
            double acc = 0.0;
            for (uint32_t t = item.y; t < item.z; ++t) {
                const int64_t wl = __ldg(lj.crow + t);
                if (wl < m.off_eff) {
                    const int r = (int)((wl - m.off_rho) / D), i = (int)((wl - m.off_rho) - (int64_t)r * D);
                    if (r == prep) acc = fma(__ldg(lj.cval + t), e0[i], acc);
                } else {
                    const int r = (int)((wl - m.off_eff) / D), i = (int)((wl - m.off_eff) - (int64_t)r * D);
                    if (r == eff) acc = fma(__ldg(lj.cval + t), sL[i], acc);
                }
            }
            Jacc[o * LJ_PT + item.x] += acc;
        }
    }
    __syncthreads();
    for (int o = 0; o < nout; ++o) {
        const int64_t el = a.out_el[q0 + o];
        const double sc = row_scale ? __ldg(row_scale + el) : 1.0;
        double* Jr = J + el * ld + tile_p0;
        for (int p = tid; p < tile_np; p += LJ_THREADS) __stcs(Jr + p, Jacc[o * LJ_PT + p] * sc);
    }
}

// ------------------------------------------------------------------------------------------------------------
// phase C, version 3 (round 2, default): no CTA barriers, one WARP per (circuit, outcome).
// Version 2 ran 10 gates x 8 outcomes = 80 serial phases per CTA, each a few DMMA between two __syncthreads and a
// dependent gather (ncu: DMMA waiting on the long scoreboard 24 % of the samples, barrier 17 %, DMMA pipe 25 %; 66 of the
// 85 ms of BASELINE config 3).  Here a warp owns one outcome of one circuit for the whole kernel:
//   * its adjoint rows e^(k+1) (the 26 GB table BH at config 3) are read exactly once, half a row (32 components = one M
//     half of the 64 x 64 block) per pass; the state rows s_k are shared by the warps of the CTA through L1;
//   * per (gate, 64 x 64 sub-block, M half): W[i][j] = sum_k e^(k+1)[i] s_k[j] on DMMA (M = 32, N = 64, K = the bucket's
//     steps, only the 8 x 8 tiles the derivative map needs: 32-bit tile mask from the host), software-pipelined one group
//     of 4 steps ahead, parked in a WARP-PRIVATE shared tile and contracted with the sparse map by the same warp (one lane
//     per parameter) into the warp's own accumulator row: only __syncwarp, no atomics, fixed order (deterministic);
//   * the finished row is stored coalesced with the objective-function row scale applied.
// One CTA = 4 warps = 4 outcomes of one circuit; all parameters at once (no parameter tiles).
// dynamic smem (doubles): LJ3_WARPS * np_pad (accumulator rows) + LJ3_WARPS * 32 * LJ3_LDW (W tiles)
// ------------------------------------------------------------------------------------------------------------
#define LJ3_WARPS 4
#define LJ3_LDW 72       // 72 = 8 mod 32 doubles... row stride chosen so that the 128-bit accumulator stores of a quarter-warp hit 8 distinct bank groups

struct LevelJ3Dev {
    const uint2* tp3;          // [n_ops * nsb * 2] per (gate, sub-block, M half): (first entry, iterations) of its lane-major list
    const uint32_t* mask3;     // [n_ops * nsb * 2] needed 8x8 tiles of the half block (bit 8*mt + nt, mt < 4)
    // Non-zeros of dG/dtheta inside the half block as LANE-MAJOR lists: entry (first + 32 k + lane) is the k-th non-zero of
    // `lane`.  Every parameter of the block lives in exactly one lane (no cross-lane conflicts, fixed summation order), the
    // lanes are balanced by non-zero count and padded with parameter -1.  The three loads of an iteration are coalesced and
    // independent of everything the loop computes, so they pipeline (round-2 profile of the item-list version: 25 % of the
    // kernel's stall samples sat on the dependent item -> non-zero -> W chain).
    const uint16_t* nz_ij;     // (i_local << 8) | j_local, i_local < 32, j_local < 64
    const int32_t* nz_p;       // parameter (column of J) or -1
    const double* nz_v;
    uint32_t zrow_f, zrow_b;   // all-zero rows of FS / BH
    int nsb, np_pad, n_og;     // sub-blocks per matrix; padded accumulator row length; outcome groups per circuit
};

struct PeerOut { int n; double* J[B200_PEERS_MAX]; };     // fused exchange: the same rows also go to the peers' arrays

template <int D>
__global__ void __launch_bounds__(LJ3_WARPS * 32, 2)
k_level_accum3(AtomDev a, ModelDev m, LevelJDev lj, LevelJ3Dev l3, double* __restrict__ J, int64_t ld,
               const double* __restrict__ row_scale, PeerOut peers)
{
    constexpr int SBD = D / 64;
    extern __shared__ __align__(16) double sml[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mrow = lane >> 2, q = lane & 3;
    const int c = blockIdx.x / l3.n_og, og = blockIdx.x - c * l3.n_og;
    const int q0 = a.out_ptr[c], nout = a.out_ptr[c + 1] - q0;
    const int o = og * LJ3_WARPS + warp;
    if (o >= nout) return;                              // (no CTA-wide barrier anywhere below)
    double* Jacc = sml + (size_t)warp * l3.np_pad;
    double* Wt = sml + (size_t)LJ3_WARPS * l3.np_pad + (size_t)warp * (32 * LJ3_LDW);
    const uint32_t p0 = a.circ_ptr[c], Lc = a.circ_ptr[c + 1] - p0;
    const size_t fb = lj.fbase[c];
    const size_t brow0 = (size_t)lj.bbase[c] + (size_t)o * (Lc + 1) + 1;       // row of e^(k+1) for step k = brow0 + k
    const uint16_t* cn = lj.bcnt + (size_t)c * a.n_ops;
    const uint16_t* perm = lj.bperm + p0;

    for (int p = lane; p < lj.n_params; p += 32) Jacc[p] = 0.0;
    __syncwarp();

    // The passes (gate, sub-block, M half) of this warp as a flat sequence, so that the first fragments of pass n + 1 can be
    // requested before pass n is parked and contracted (one exposed round trip per warp instead of one per pass).
    int g = -1, sb = 0, mh = 0, cnt = 0;
    uint32_t tb = 0, it0 = 0, it1 = 0;
    unsigned mask = 0;
    auto advance = [&]() -> bool {
        for (;;) {
            bool next_gate = (g < 0) || (cnt == 0);
            if (!next_gate) {
                if (mh == 0) mh = 1; else { mh = 0; ++sb; }
                if (sb >= l3.nsb) next_gate = true;
            }
            if (next_gate) {
                if (g >= 0) tb += cnt;
                ++g;
                if (g >= a.n_ops) return false;
                cnt = cn[g]; sb = 0; mh = 0;
                {   // pull the adjoint rows of the NEXT gate's bucket (HBM: each is read by this warp only) into L2 while this
                    // gate is processed: their later gathers then cost an L2 hit instead of a DRAM round trip
                    const int gn = g + 1;
                    if (gn < a.n_ops) {
                        const int cn_next = cn[gn];
                        for (int t = lane; t < cn_next; t += 32) {
                            const int kk = (int)__ldg(perm + tb + cnt + t);
                            const double* er = lj.BH + (brow0 + kk) * D;
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(er), "r"(D * 8) : "memory");
                        }
                    }
                }
                if (cnt == 0) continue;
            }
            const int key = (g * l3.nsb + sb) * 2 + mh;
            const uint2 tp = __ldg(l3.tp3 + key);
            it0 = tp.x; it1 = tp.y;                     // first entry, iterations
            if (it1 == 0) continue;
            mask = __ldg(l3.mask3 + key);
            return true;
        }
    };
    // fragments of the group of 4 steps starting at k0 of the CURRENT iterator position.  A (adjoint rows: streamed from HBM / L2,
    // read by this warp only) is requested TWO groups ahead, B (state rows: shared by the warps of the CTA, L1 hits) one group
    // ahead -- the profile of the one-group-ahead version had 20 % of its stall samples on the first DMMA of a group.
    auto step_of = [&](int k0) -> int { const int t = k0 + q; return (t < cnt) ? (int)__ldg(perm + tb + t) : -1; };
    auto fetchA = [&](int k0, double (&xa)[4]) {
        const int k = step_of(k0);
        const int ib = (sb / SBD) * 64 + mh * 32;
        const double* er = lj.BH + (k >= 0 ? (brow0 + k) : (size_t)l3.zrow_b) * D + ib + mrow;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) xa[mt] = ((mask >> (8 * mt)) & 0xffu) ? __ldg(er + 8 * mt) : 0.0;
    };
    auto fetchB = [&](int k0, double (&xb)[8]) {
        const int k = step_of(k0);
        const unsigned need_n = (mask | (mask >> 8) | (mask >> 16) | (mask >> 24)) & 0xffu;
        const int jb = (sb % SBD) * 64;
        const double* sr = lj.FS + (k >= 0 ? (fb + k) : (size_t)l3.zrow_f) * D + jb + mrow;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) xb[nt] = (need_n & (1u << nt)) ? __ldg(sr + 8 * nt) : 0.0;
    };
    double fa[4], fa1[4], fbv[8];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) fa1[mt] = 0.0;
    bool have = advance();
    if (have) { fetchA(0, fa); fetchB(0, fbv); if (cnt > 4) fetchA(4, fa1); }
    while (have) {
        const unsigned cmask = mask;
        const uint32_t cit0 = it0, cit1 = it1;
        const int ns4 = (cnt + 3) & ~3;
        double acc[4][8][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
        for (int k0 = 0; k0 < ns4; k0 += 4) {
            double fa2[4], nb[8];
            if (k0 + 8 < ns4) fetchA(k0 + 8, fa2);
            if (k0 + 4 < ns4) fetchB(k0 + 4, nb);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    if (cmask & (1u << (8 * mt + nt))) dmma884(acc[mt][nt][0], acc[mt][nt][1], fa[mt], fbv[nt]);   // warp-uniform
            if (k0 + 4 < ns4) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) fa[mt] = fa1[mt];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) fbv[nt] = nb[nt];
            }
            if (k0 + 8 < ns4) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) fa1[mt] = fa2[mt];
            }
        }
        have = advance();                               // iterator now at the NEXT pass: request its first fragments
        if (have) { fetchA(0, fa); fetchB(0, fbv); if (cnt > 4) fetchA(4, fa1); }
        __syncwarp();                                   // the previous contraction has read Wt
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
                if (cmask & (1u << (8 * mt + nt)))
                    *reinterpret_cast<double2*>(Wt + (mt * 8 + mrow) * LJ3_LDW + nt * 8 + 2 * q) =
                        make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        __syncwarp();
        {
            double sacc = 0.0; int pp = -1;
#pragma unroll 4
            for (uint32_t k = 0; k < cit1; ++k) {
                const uint32_t e = cit0 + k * 32 + lane;
                const unsigned ij = __ldg(l3.nz_ij + e);
                const int pn = __ldg(l3.nz_p + e);
                const double v = __ldg(l3.nz_v + e);
                const double w = Wt[(ij >> 8) * LJ3_LDW + (ij & 0xffu)];
                if (pn != pp) { if (pp >= 0) Jacc[pp] += sacc; sacc = 0.0; pp = pn; }
                sacc = fma(v, w, sacc);
            }
            if (pp >= 0) Jacc[pp] += sacc;
        }
    }
    __syncwarp();
    {   // state-preparation and effect rows of D (item lists per parameter tile, as in versions 1 / 2)
        const int prep = a.circ_prep[c];
        const int eff = a.out_eff[q0 + o];
        const double* sL = lj.FS + (fb + Lc) * D;
        const double* e0 = lj.BH + (brow0 - 1) * D;
        for (int tile = 0; tile < lj.n_tiles; ++tile) {
            const uint32_t* tp = lj.ti_ptr + (size_t)tile * (a.n_ops + 2);
            const uint32_t it0 = tp[a.n_ops], it1 = tp[a.n_ops + 1];
            for (uint32_t it = it0 + lane; it < it1; it += 32) {
                const uint4 item = __ldg(lj.items + it);
                double acc = 0.0;
                for (uint32_t t = item.y; t < item.z; ++t) {
                    const int64_t wl = __ldg(lj.crow + t);
                    if (wl < m.off_eff) {
                        const int r = (int)((wl - m.off_rho) / D), i = (int)((wl - m.off_rho) - (int64_t)r * D);
                        if (r == prep) acc = fma(__ldg(lj.cval + t), e0[i], acc);
                    } else {
                        const int r = (int)((wl - m.off_eff) / D), i = (int)((wl - m.off_eff) - (int64_t)r * D);
                        if (r == eff) acc = fma(__ldg(lj.cval + t), sL[i], acc);
                    }
                }
                Jacc[tile * lj.pt + item.x] += acc;
            }
        }
    }
    __syncwarp();
    const int64_t el = a.out_el[q0 + o];
    const double sc = row_scale ? __ldg(row_scale + el) : 1.0;
    double* Jr = J + el * ld;
    for (int p = lane; p < lj.n_params; p += 32) __stcs(Jr + p, Jacc[p] * sc);
    for (int r = 0; r < peers.n; ++r) {
        double* Jp = peers.J[r] + el * ld;
        for (int p = lane; p < lj.n_params; p += 32) __stcs(Jp + p, Jacc[p] * sc);
    }
}
