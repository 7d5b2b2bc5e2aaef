// ref_driver.cpp -- drives the REFERENCE's own C++ reps (compiled from /root/reference, never copied)
// through the reference's table-interpreter loop.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// Links against  pygsti/evotypes/densitymx/{statecreps,opcreps,effectcreps}.cpp  of the reference:
//   CReps_densitymx::OpCRep_Dense::acton          (opcreps.cpp:40-54)
//   CReps_densitymx::EffectCRep_Dense::probability (effectcreps.cpp:39-45)
//   CReps_densitymx::StateCRep                     (statecreps.cpp:20-57)
// The loop below is a C++ restatement of the Cython `dm_mapfill_probs`
// (pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:194-287: same pointer-swap cache management) and of the
// finite-difference driver `mapfill_dprobs_atom` (pyx:290-383), because Cython sources cannot be linked
// without Python.  The reps read the gate arrays IN PLACE (opreps.pyx:86-89), so -- exactly like the
// reference -- a parameter perturbation is a write into the model buffer between passes.
#include "statecreps.h"
#include "opcreps.h"
#include "effectcreps.h"
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace CReps_densitymx;

struct ref_atom {
    int dim, n_ops, n_rho, n_eff;
    int64_t n_rows, n_elements;
    int32_t cache_size;
    const int32_t *row_ptr, *row_ops, *row_istart, *row_prep, *row_icache;
    const int32_t *out_ptr, *out_eff, *out_el;
};

struct RefModel {
    std::vector<double> M;                       // [G | rho | E], owned, perturbed in place
    std::vector<OpCRep*> ops; std::vector<StateCRep*> rhos; std::vector<EffectCRep*> effs;
    RefModel(const ref_atom* a, const double* G, const double* rho, const double* E) {
        const INT d = a->dim;
        size_t nG = (size_t)a->n_ops * d * d, nR = (size_t)a->n_rho * d, nE = (size_t)a->n_eff * d;
        M.resize(nG + nR + nE);
        memcpy(M.data(), G, nG * 8); memcpy(M.data() + nG, rho, nR * 8); memcpy(M.data() + nG + nR, E, nE * 8);
        for (int g = 0; g < a->n_ops; g++) ops.push_back(new OpCRep_Dense(M.data() + (size_t)g * d * d, d));
        for (int r = 0; r < a->n_rho; r++) rhos.push_back(new StateCRep(M.data() + nG + (size_t)r * d, d, false));
        for (int e = 0; e < a->n_eff; e++) effs.push_back(new EffectCRep_Dense(M.data() + nG + nR + (size_t)e * d, d));
    }
    ~RefModel() { for (auto p : ops) delete p; for (auto p : rhos) delete p; for (auto p : effs) delete p; }
};

// pyx:194-287
static void dm_mapfill_probs(const ref_atom* a, RefModel& m, std::vector<StateCRep*>& rho_cache, double* out) {
    const INT dim = a->dim;
    StateCRep *init_state, *prop1, *tprop, *final_state;
    StateCRep* prop2 = new StateCRep(dim);
    StateCRep* shelved = new StateCRep(dim);
    for (int64_t k = 0; k < a->n_rows; k++) {
        const int32_t istart = a->row_istart[k], icache = a->row_icache[k];
        init_state = (istart == -1) ? m.rhos[a->row_prep[k]] : rho_cache[istart];
        prop1 = (icache == -1) ? shelved : rho_cache[icache];
        prop1->copy_from(init_state);
        for (int32_t l = a->row_ptr[k]; l < a->row_ptr[k + 1]; l++) {
            m.ops[a->row_ops[l]]->acton(prop1, prop2);
            tprop = prop1; prop1 = prop2; prop2 = tprop;
        }
        final_state = prop1;
        StateCRep* precomp_state = prop2; INT precomp_id = 0;
        for (int32_t j = a->out_ptr[k]; j < a->out_ptr[k + 1]; j++)
            out[a->out_el[j]] = m.effs[a->out_eff[j]]->probability_using_cache(final_state, precomp_state, precomp_id);
        if (icache != -1) rho_cache[icache] = final_state; else shelved = final_state;
    }
    delete prop2; delete shelved;
}

static std::vector<StateCRep*> create_rhocache(int32_t n, INT dim) {   // pyx:113-119
    std::vector<StateCRep*> c((size_t)n);
    for (auto& p : c) p = new StateCRep(dim);
    return c;
}

extern "C" int ref_mapfill_probs(const ref_atom* a, const double* G, const double* rho, const double* E, double* out) {
    RefModel m(a, G, rho, E);
    std::vector<StateCRep*> cache = create_rhocache(a->cache_size, a->dim);
    dm_mapfill_probs(a, m, cache, out);
    for (auto p : cache) delete p;
    return 0;
}

// pyx:290-383 with set_parameter_value restated for members linear in their parameters (CSC column of D)
extern "C" int ref_dprobs_fd(const ref_atom* a, const double* G, const double* rho, const double* E,
                             const int32_t* cptr, const int32_t* crow, const double* cval,
                             int p_lo, int p_hi, double eps, double* out, int64_t ld, double* probs_out, int n_threads) {
    std::vector<double> probs((size_t)a->n_elements);
    ref_mapfill_probs(a, G, rho, E, probs.data());
    if (probs_out) memcpy(probs_out, probs.data(), probs.size() * 8);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        RefModel m(a, G, rho, E);
        std::vector<StateCRep*> cache = create_rhocache(a->cache_size, a->dim);
        std::vector<double> probs2((size_t)a->n_elements);
        std::vector<double> saved;
#pragma omp for schedule(dynamic, 1)
        for (int p = p_lo; p < p_hi; p++) {
            saved.clear();
            for (int32_t t = cptr[p]; t < cptr[p + 1]; t++) { saved.push_back(m.M[crow[t]]); m.M[crow[t]] += eps * cval[t]; }
            dm_mapfill_probs(a, m, cache, probs2.data());
            for (int64_t el = 0; el < a->n_elements; el++) out[el * ld + (p - p_lo)] = (probs2[el] - probs[el]) / eps;
            for (int32_t t = cptr[p], i = 0; t < cptr[p + 1]; t++, i++) m.M[crow[t]] = saved[i];
        }
        for (auto p : cache) delete p;
    }
    return 0;
}

extern "C" int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
