/*
 * fwdsim_oracle.c -- CPU ORACLE (plain C restatement).  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may load this
 * library; nothing under pygsti_b200/ links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_cpu.py checks it against golden vectors produced by
 * running the reference itself (tests/golden/make_golden.py) and against oracle/_ref (the reference's
 * own C++ reps compiled from /root/reference).
 *
 * Restated reference code (pyGSTi @ 6822f14):
 *   oracle_mapfill_probs ....... dm_mapfill_probs, pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:194-287
 *                                with OpCRep_Dense::acton (pygsti/evotypes/densitymx/opcreps.cpp:40-54),
 *                                StateCRep::copy_from (statecreps.cpp:52-56) and
 *                                EffectCRep_Dense::probability (effectcreps.cpp:39-45)
 *   oracle_dprobs_fd ........... mapfill_dprobs_atom, pyx:290-383 (base pass + one full table pass per
 *                                parameter, (probs2 - probs)/eps); the model update
 *                                model.set_parameter_values (pygsti/models/model.py:1223-1310) is restated
 *                                for members LINEAR in their parameters: M += eps * dM/dtheta_p in place
 *   oracle_dprobs_analytic ..... MatrixForwardSimulator._dprobs_from_rho_e,
 *                                pygsti/forwardsims/matrixforwardsim.py:1059-1139 (dp_dOps + dp_drhos + dp_dEs)
 *                                evaluated with forward/backward vectors instead of product caches
 * Build: gcc -O3 -fPIC -shared -fopenmp -o oracle/liboracle.so oracle/fwdsim_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int dim, n_ops, n_rho, n_eff;
    int64_t n_rows, n_elements;
    int32_t cache_size;
    const int32_t *row_ptr, *row_ops, *row_istart, *row_prep, *row_icache;
    const int32_t *out_ptr, *out_eff, *out_el;
} oracle_atom;

/* out[i] = sum_j G[i][j] v[j]  -- opcreps.cpp:40-54 (same loop order, no FMA contraction requested) */
static void dense_acton(const double* G, const double* v, double* out, int d) {
    for (int i = 0; i < d; i++) {
        double acc = 0.0;
        const double* row = G + (int64_t)i * d;
        for (int j = 0; j < d; j++) acc += row[j] * v[j];
        out[i] = acc;
    }
}

/* pyx:224-283.  cache: [cache_size][d] scratch provided by the caller (create_rhocache, pyx:178). */
static void mapfill_probs_core(const oracle_atom* a, const double* G, const double* rho, const double* E,
                               double* cache, double* tmp /* 2*d */, double* out) {
    const int d = a->dim;
    for (int64_t k = 0; k < a->n_rows; k++) {
        const double* init = (a->row_istart[k] < 0) ? rho + (int64_t)a->row_prep[k] * d
                                                    : cache + (int64_t)a->row_istart[k] * d;
        double* p1 = tmp; double* p2 = tmp + d;
        memcpy(p1, init, sizeof(double) * d);                          /* copy_from */
        for (int32_t l = a->row_ptr[k]; l < a->row_ptr[k + 1]; l++) {
            dense_acton(G + (int64_t)a->row_ops[l] * d * d, p1, p2, d);
            double* t = p1; p1 = p2; p2 = t;                           /* swap prop1 <-> prop2 */
        }
        for (int32_t j = a->out_ptr[k]; j < a->out_ptr[k + 1]; j++) {
            const double* e = E + (int64_t)a->out_eff[j] * d;
            double p = 0.0;
            for (int i = 0; i < d; i++) p += e[i] * p1[i];              /* effectcreps.cpp:39-45 */
            out[a->out_el[j]] = p;
        }
        if (a->row_icache[k] >= 0) memcpy(cache + (int64_t)a->row_icache[k] * d, p1, sizeof(double) * d);
    }
}

int oracle_mapfill_probs(const oracle_atom* a, const double* G, const double* rho, const double* E, double* out) {
    const int d = a->dim;
    double* cache = (double*)calloc((size_t)(a->cache_size > 0 ? a->cache_size : 1) * d, sizeof(double));
    double* tmp = (double*)malloc(sizeof(double) * 2 * d);
    if (!cache || !tmp) { free(cache); free(tmp); return -1; }
    mapfill_probs_core(a, G, rho, E, cache, tmp, out);
    free(cache); free(tmp);
    return 0;
}

/* D in CSC: column p holds entries [cptr[p], cptr[p+1]) of (row w in W space, value).
 * out[el*ld + (p - p_lo)] for p in [p_lo, p_hi).  n_threads <= 0 -> all available. */
int oracle_dprobs_fd(const oracle_atom* a, const double* G, const double* rho, const double* E,
                     const int32_t* cptr, const int32_t* crow, const double* cval,
                     int p_lo, int p_hi, double eps, double* out, int64_t ld, double* probs_out, int n_threads) {
    const int d = a->dim;
    const int64_t nG = (int64_t)a->n_ops * d * d, nR = (int64_t)a->n_rho * d, nE_ = (int64_t)a->n_eff * d;
    const int64_t nW = nG + nR + nE_;
    double* probs = (double*)malloc(sizeof(double) * (a->n_elements > 0 ? a->n_elements : 1));
    if (!probs) return -1;
    if (oracle_mapfill_probs(a, G, rho, E, probs)) { free(probs); return -1; }
    if (probs_out) memcpy(probs_out, probs, sizeof(double) * a->n_elements);
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        double* M = (double*)malloc(sizeof(double) * nW);
        double* cache = (double*)calloc((size_t)(a->cache_size > 0 ? a->cache_size : 1) * d, sizeof(double));
        double* tmp = (double*)malloc(sizeof(double) * 2 * d);
        double* probs2 = (double*)malloc(sizeof(double) * (a->n_elements > 0 ? a->n_elements : 1));
        if (!M || !cache || !tmp || !probs2) {
#pragma omp atomic write
            err = 1;
        } else {
            memcpy(M, G, sizeof(double) * nG); memcpy(M + nG, rho, sizeof(double) * nR);
            memcpy(M + nG + nR, E, sizeof(double) * nE_);
#pragma omp for schedule(dynamic, 1)
            for (int p = p_lo; p < p_hi; p++) {
                for (int32_t t = cptr[p]; t < cptr[p + 1]; t++) M[crow[t]] += eps * cval[t];   /* set_parameter_value */
                mapfill_probs_core(a, M, M + nG, M + nG + nR, cache, tmp, probs2);
                for (int64_t el = 0; el < a->n_elements; el++)
                    out[el * ld + (p - p_lo)] = (probs2[el] - probs[el]) / eps;                /* pyx:374 */
                /* restore: the reference re-sets the previous parameter to its original value (pyx:371,381) */
                for (int32_t t = cptr[p]; t < cptr[p + 1]; t++) {
                    int64_t w = crow[t];
                    M[w] = (w < nG) ? G[w] : (w < nG + nR) ? rho[w - nG] : E[w - nG - nR];
                }
            }
        }
        free(M); free(cache); free(tmp); free(probs2);
    }
    free(probs);
    return err ? -1 : 0;
}

/* Analytic Jacobian: for every table row (circuit) forward states s_k, then per outcome the backward
 * vectors e_k; W-space gradient accumulated in wrow then contracted with D (CSC). */
int oracle_dprobs_analytic(const oracle_atom* a, const double* G, const double* rho, const double* E,
                           const int32_t* cptr, const int32_t* crow, const double* cval, int n_params,
                           double* out, int64_t ld, double* probs_out) {
    const int d = a->dim;
    const int64_t nG = (int64_t)a->n_ops * d * d, nR = (int64_t)a->n_rho * d, nE_ = (int64_t)a->n_eff * d;
    const int64_t nW = nG + nR + nE_;
    /* expanded sequences: cache slot -> (prep, ops) of the producing row */
    int64_t* slot_row = (int64_t*)malloc(sizeof(int64_t) * (a->cache_size > 0 ? a->cache_size : 1));
    int64_t* xptr = (int64_t*)calloc((size_t)a->n_rows + 1, sizeof(int64_t));
    int32_t* xprep = (int32_t*)malloc(sizeof(int32_t) * (a->n_rows > 0 ? a->n_rows : 1));
    if (!slot_row || !xptr || !xprep) return -1;
    for (int64_t k = 0; k < a->n_rows; k++) {
        int64_t len = a->row_ptr[k + 1] - a->row_ptr[k];
        if (a->row_istart[k] >= 0) { int64_t s = slot_row[a->row_istart[k]]; len += xptr[s + 1] - xptr[s]; xprep[k] = xprep[s]; }
        else xprep[k] = a->row_prep[k];
        xptr[k + 1] = xptr[k] + len;
        if (a->row_icache[k] >= 0) slot_row[a->row_icache[k]] = k;
    }
    int32_t* xops = (int32_t*)malloc(sizeof(int32_t) * (xptr[a->n_rows] > 0 ? xptr[a->n_rows] : 1));
    int64_t maxL = 0;
    for (int64_t k = 0; k < a->n_rows; k++) if (xptr[k + 1] - xptr[k] > maxL) maxL = xptr[k + 1] - xptr[k];
    /* second pass with slot tracking in evaluation order */
    for (int32_t s = 0; s < a->cache_size; s++) slot_row[s] = -1;
    for (int64_t k = 0; k < a->n_rows; k++) {
        int32_t* dst = xops + xptr[k];
        if (a->row_istart[k] >= 0) {
            int64_t s = slot_row[a->row_istart[k]];
            int64_t n = xptr[s + 1] - xptr[s];
            memcpy(dst, xops + xptr[s], sizeof(int32_t) * n); dst += n;
        }
        for (int32_t l = a->row_ptr[k]; l < a->row_ptr[k + 1]; l++) *dst++ = a->row_ops[l];
        if (a->row_icache[k] >= 0) slot_row[a->row_icache[k]] = k;
    }
    double* st = (double*)malloc(sizeof(double) * (maxL + 1) * d);
    double* wrow = (double*)malloc(sizeof(double) * nW);
    double* e = (double*)malloc(sizeof(double) * 2 * d);
    for (int64_t k = 0; k < a->n_rows; k++) {
        const int32_t* seq = xops + xptr[k];
        const int64_t L = xptr[k + 1] - xptr[k];
        memcpy(st, rho + (int64_t)xprep[k] * d, sizeof(double) * d);
        for (int64_t m = 0; m < L; m++) dense_acton(G + (int64_t)seq[m] * d * d, st + m * d, st + (m + 1) * d, d);
        const double* sL = st + L * d;
        for (int32_t j = a->out_ptr[k]; j < a->out_ptr[k + 1]; j++) {
            const int64_t el = a->out_el[j]; const int ei = a->out_eff[j];
            memset(wrow, 0, sizeof(double) * nW);
            double* ec = e; double* en = e + d;
            memcpy(ec, E + (int64_t)ei * d, sizeof(double) * d);
            double p = 0.0;
            for (int i = 0; i < d; i++) { p += ec[i] * sL[i]; wrow[nG + nR + (int64_t)ei * d + i] += sL[i]; }
            if (probs_out) probs_out[el] = p;
            for (int64_t m = L - 1; m >= 0; m--) {
                const int g = seq[m];
                const double* Gg = G + (int64_t)g * d * d;
                const double* s = st + m * d;
                double* wg = wrow + (int64_t)g * d * d;
                for (int i = 0; i < d; i++) for (int jj = 0; jj < d; jj++) wg[i * d + jj] += ec[i] * s[jj];
                for (int jj = 0; jj < d; jj++) {            /* adjoint_acton, opcreps.cpp:56-68 */
                    double acc = 0.0;
                    for (int i = 0; i < d; i++) acc += Gg[(int64_t)i * d + jj] * ec[i];
                    en[jj] = acc;
                }
                double* t = ec; ec = en; en = t;
            }
            for (int i = 0; i < d; i++) wrow[nG + (int64_t)xprep[k] * d + i] += ec[i];
            for (int pcol = 0; pcol < n_params; pcol++) {
                double acc = 0.0;
                for (int32_t t = cptr[pcol]; t < cptr[pcol + 1]; t++) acc += cval[t] * wrow[crow[t]];
                out[el * ld + pcol] = acc;
            }
        }
    }
    free(st); free(wrow); free(e); free(xops); free(xprep); free(xptr); free(slot_row);
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
