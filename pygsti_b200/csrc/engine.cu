// engine.cu -- host side of libb200fwdsim.so: contexts, atoms, the C ABI of include/b200_fwdsim.h.
//
// There is no CPU compute path in this library: every fill entry point launches CUDA kernels or fails.
#include "../../include/b200_fwdsim.h"
#include "common.cuh"
#include "kernels_generic.cuh"
#include "kernels_d16_trie.cuh"
#include "kernels_level.cuh"
#include "kernels_levelj.cuh"
#include "kernels_jtj.cuh"
#include "kernels_gemm.cuh"
#include "kernels_factored.cuh"
#include "kernels_factoredj.cuh"
#include "kernels_ozaki.cuh"
#include <cstdlib>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <climits>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    int code_ = (e_ == cudaErrorMemoryAllocation) ? B200_E_NOMEM : B200_E_CUDA; \
    return fail(code_, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    int64_t launches = 0;
    DevBuf out_buf;      // device staging of host-buffer results
    DevBuf probs_buf;
    DevBuf w_buf;        // W matrix of the general path
    DevBuf scratch;      // forward-state scratch of the generic kernels
    DevBuf lvl_states;   // ping-pong state buffers of the level-batched dense path
    DevBuf lj_fs, lj_bh; // level-batched Jacobian path: state table / adjoint table (all levels kept)
    cudaStream_t aux = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // backward sweep runs beside the forward sweep
    DevBuf scale_buf, f_buf, jtj_buf, jtf_buf;   // fused objective Jacobian / J^T J
    DevBuf atb_part, atb_part_f;                 // partial tiles of the A^T B reductions (k_atb_dmma)
    DevBuf fj_fs, fj_counter;                    // factored Jacobian: forward states of every factor step, work counter
    DevBuf oz_S, oz_misc, oz_part;               // Ozaki J^T J: int8 digit slices, (column maxima | exponents | tile list), partial tiles
    DevBuf hb[24];                               // scratch of the Hessian-block path (kept between calls: cudaMalloc / cudaFree per
                                                 // rectangle cost more than the kernels once peer access is enabled)
    DevBuf fd_models, fd_gt, fd_probs;
    DevBuf lind[20];                                     // b200_lindblad_members: inputs, intermediates, outputs
    int jtj_mode = -1;                                   // b200_ctx_set_jtj_mode
    std::vector<int2> oz_tiles; int oz_tiles_np = -1;    // lower-triangle tile list of the Ozaki SYRK (host copy outlives the async upload)
    bool phase_timing = false;                           // b200_ctx_phase_timing: events around the d16 trie phases
    std::vector<cudaEvent_t> phase_events;               // 4 per call: start, after prepare, after chains, after accumulate
};

struct b200_atom {
    b200_ctx* ctx = nullptr;
    int dim = 0, n_ops = 0, n_rho = 0, n_eff = 0;
    int64_t n_rows = 0, n_elements = 0, n_prop_table = 0, n_prop_expanded = 0;
    int max_depth = 0;
    int64_t n_w = 0, off_rho = 0, off_eff = 0;
    // device tables
    DevBuf circ_ptr, circ_ops, circ_prep, out_ptr, out_eff, out_el;
    DevBuf bperm, bcnt;                // per circuit: step indices sorted by gate (bucket order), bucket sizes [n_ops]
    // level-batched dense path (d >= 64)
    bool has_levels = false;
    DevBuf lvl_circ, lvl_tiles;
    std::vector<uint32_t> lvl_tile_ptr;     // [max_depth+1] tile offsets per level (host)
    // level-batched Jacobian path (d >= 64): row maps of the state / adjoint tables, backward-sweep tiles, D tile plan
    bool has_lj = false;
    DevBuf lj_fbase, lj_bbase, lj_frow, lj_brow, lj_btiles, lj_ti_ptr, lj_items;
    DevBuf lj2_ti_ptr, lj2_mask, lj2_items, lj2_ij, lj2_v;       // DMMA accumulate (k_level_accum2): per (tile, gate, sub-block)
    DevBuf lj3_tp, lj3_mask, lj3_items, lj3_ij, lj3_v;           // warp-per-outcome accumulate (k_level_accum3): per (gate, sub-block, M half)
    std::vector<uint32_t> lj_btile_ptr;     // [max_depth+1]
    uint64_t lj_rows_f = 0, lj_rows_b = 0;
    int lj_no_max = 0, lj_n_tiles = 0, lj_pt = LJ_PT_MIN;   // lj_pt: parameters per accumulate tile (chosen in set_derivs)
    // trie path (prefix + suffix sharing)
    bool has_trie = false;
    DevBuf tf_meta, tf_op, tb_meta, tb_op, t_fn, t_bn, t_fend, t_bend, tf_par, tb_par;   // (t*_par: nodes some chain waits for)
    uint32_t n_fpar = 0, n_bpar = 0;
    DevBuf t_SH, t_counters, t_units, t_uidx;      // t_SH: both value tables in ONE allocation, [H | S] (one L2 persisting window)
    size_t t_S_bytes = 0, t_H_bytes = 0;
    double* tH() { return t_SH.as<double>(); }
    double* tS() { return reinterpret_cast<double*>(reinterpret_cast<char*>(t_SH.p) + t_H_bytes); }
    int n_units = 0;
    int n_fchains = 0, n_bchains = 0; uint32_t n_fnodes = 0, n_bnodes = 0;
    // model
    bool has_model = false;
    DevBuf M, Gt;
    bool has_factored = false;             // gates also held as factor programs (b200_atom_set_model_factored)
    DevBuf fac_ptr, fac_rec, fac_mats; int fac_n_mats = 0, fac_n = 0; bool fac_probs_ok = false;   // (programs small enough for shared memory)
    std::vector<FactorRec> h_fac; std::vector<int32_t> h_fptr;     // host copy of the factor programs
    DevBuf fac_ffo; int fac_n_frag = 0;                             // DMMA fragment offsets of the factors (k_probs_fdmma)
    // factor-space derivative map (b200_atom_set_derivs_factored): the Jacobian straight from the factor programs (kernels_factoredj.cuh)
    bool has_fderivs = false; int32_t fj_n_params = 0; int fj_n_acc = 0, fj_n_frag = 0; uint64_t fj_rows = 0;
    std::vector<FactorRec> fj_fac;                                 // the factor structure the map was built for
    DevBuf fj_base, fj_out_circ, fj_fao, fj_ffo, fj_cptr, fj_ccode, fj_cval, fj_slots, fj_steps; int fj_unit = 0; int fj_slot_fao[FJ64_REG_SLOTS] = {-1, -1, -1, -1};
    // derivative map
    bool has_derivs = false;
    int32_t n_params = 0;
    bool unit_perm = false;
    DevBuf cptr, crow, cval;               // CSC of D
    bool has_dense = false; int64_t ldD = 0;   // general path: D also DENSE [n_w x ldD] + per 128-column tile the 16-row K chunks with non-zeros
    DevBuf Dd, kt_ptr, kt_idx;
    DevBuf colmap, spam_col, spam_w;       // fused-path maps (unit partial permutation)
    int n_spam = 0;
    DevBuf id_colmap, id_spam_col, id_spam_w;  // identity maps: d16 kernel writing W for the general path
    int id_n_spam = 0;
    std::vector<int32_t> h_cptr, h_crow; std::vector<double> h_cval;  // host copy (hessian / fd)
    // affine model update (b200_atom_bind_params): CSR of D by W-space row, M_const, parameter staging
    bool has_bind = false; int32_t bind_n_params = 0;
    DevBuf aff_rptr, aff_rcol, aff_rval, aff_const, aff_theta;
};

static AtomDev atom_dev(b200_atom* a) {
    AtomDev d;
    d.dim = a->dim; d.n_ops = a->n_ops; d.n_rho = a->n_rho; d.n_eff = a->n_eff;
    d.n_circ = (int)a->n_rows; d.max_depth = a->max_depth; d.n_elements = a->n_elements;
    d.circ_ptr = a->circ_ptr.as<uint32_t>(); d.circ_ops = a->circ_ops.as<int32_t>();
    d.circ_prep = a->circ_prep.as<int32_t>(); d.out_ptr = a->out_ptr.as<int32_t>();
    d.out_eff = a->out_eff.as<int32_t>(); d.out_el = a->out_el.as<int32_t>();
    return d;
}
static ModelDev model_dev(b200_atom* a) {
    ModelDev m;
    m.M = a->M.as<double>(); m.Gt = a->Gt.as<double>();
    m.n_w = a->n_w; m.off_rho = a->off_rho; m.off_eff = a->off_eff;
    return m;
}

// ------------------------------------------------------------------------------------------------
extern "C" int b200_version(void) { return 100; }
extern "C" const char* b200_last_error(void) { return g_err.c_str(); }

extern "C" int b200_device_count(int* n_out) {
    if (!n_out) return fail(B200_E_INVALID, "n_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *n_out = 0; return fail(B200_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *n_out = n;
    return B200_OK;
}

extern "C" int b200_ctx_create(int device, void* stream, b200_ctx** out) {
    if (!out) return fail(B200_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(B200_E_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(B200_E_INVALID, "device %d out of range [0,%d)", device, n);
    CU(cudaSetDevice(device));
    b200_ctx* c = new b200_ctx();
    c->device = device;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("B200_JTJ")) c->jtj_mode = !strcmp(e, "dmma") ? 0 : !strcmp(e, "ozaki7") ? 7 : !strcmp(e, "ozaki") ? 8 : -1;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    *out = c;
    return B200_OK;
}

static void phase_clear(b200_ctx* c);
extern "C" int b200_ctx_destroy(b200_ctx* c) {
    if (!c) return B200_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->out_buf.release(); c->probs_buf.release(); c->w_buf.release(); c->scratch.release(); c->lvl_states.release();
    c->lj_fs.release(); c->lj_bh.release();
    if (c->aux) { cudaStreamSynchronize(c->aux); cudaStreamDestroy(c->aux); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    phase_clear(c);
    for (DevBuf& b : c->lind) b.release();
    c->scale_buf.release(); c->f_buf.release(); c->jtj_buf.release(); c->jtf_buf.release();
    c->atb_part.release(); c->atb_part_f.release();
    for (DevBuf& b : c->hb) b.release();
    c->oz_S.release(); c->oz_misc.release(); c->oz_part.release();
    c->fj_fs.release(); c->fj_counter.release();
    c->fd_models.release(); c->fd_gt.release(); c->fd_probs.release();
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B200_OK;
}

extern "C" int b200_ctx_sync(b200_ctx* c) {
    if (!c) return fail(B200_E_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

static int phase_mark(b200_ctx* c) {          // record one phase boundary on the launching stream (timing mode only)
    if (!c->phase_timing) return B200_OK;
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    c->phase_events.push_back(e);
    CU(cudaEventRecord(e, c->stream));
    return B200_OK;
}
static void phase_clear(b200_ctx* c) {
    for (cudaEvent_t e : c->phase_events) cudaEventDestroy(e);
    c->phase_events.clear();
}
extern "C" int b200_ctx_set_jtj_mode(b200_ctx* c, int mode) {
    if (!c) return fail(B200_E_INVALID, "NULL ctx");
    if (mode != -1 && mode != 0 && mode != 7 && mode != 8) return fail(B200_E_INVALID, "jtj mode must be -1, 0, 7 or 8");
    c->jtj_mode = mode;
    return B200_OK;
}

extern "C" int b200_ctx_phase_timing(b200_ctx* c, int on) {
    if (!c) return fail(B200_E_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    phase_clear(c);
    c->phase_timing = on != 0;
    return B200_OK;
}
extern "C" int b200_ctx_phase_ms(b200_ctx* c, double ms_out[3], int64_t* n_calls_out) {
    if (!c || !ms_out || !n_calls_out) return fail(B200_E_INVALID, "NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    ms_out[0] = ms_out[1] = ms_out[2] = 0.0;
    const size_t n = c->phase_events.size() / 4;
    for (size_t k = 0; k < n; ++k)
        for (int ph = 0; ph < 3; ++ph) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->phase_events[4 * k + ph], c->phase_events[4 * k + ph + 1]));
            ms_out[ph] += ms;
        }
    *n_calls_out = (int64_t)n;
    phase_clear(c);
    return B200_OK;
}

extern "C" int b200_ctx_launch_count(b200_ctx* c, int64_t* n_out) {
    if (!c || !n_out) return fail(B200_E_INVALID, "NULL argument");
    *n_out = c->launches;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// atom upload: validate the prefix table, expand it into independent circuits (host), upload.
// ------------------------------------------------------------------------------------------------
template <class T>
static int upload_vec(DevBuf& b, const std::vector<T>& v, cudaStream_t s) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CU(b.ensure(bytes));
    if (!v.empty()) CU(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return B200_OK;
}

#include "kernels_lindblad.cuh"
#include "trie_host.h"   // TrieHost, build_trie: prefix / suffix tries cut into chains (heavy-path decomposition)

extern "C" int b200_atom_upload(b200_ctx* ctx, int dim, int n_ops, int n_rho, int n_eff,
                                int64_t n_rows, const int32_t* row_ptr, const int32_t* row_ops,
                                const int32_t* row_istart, const int32_t* row_prep, const int32_t* row_icache,
                                int32_t cache_size,
                                const int32_t* out_ptr, const int32_t* out_eff, const int32_t* out_el,
                                int64_t n_elements, b200_atom** out)
{
    if (!ctx || !out) return fail(B200_E_INVALID, "NULL ctx/out");
    *out = nullptr;
    if (dim != 4 && dim != 16 && dim != 64 && dim != 256)
        return fail(B200_E_UNSUPPORTED, "dim=%d: kernels are built for d = 4, 16, 64, 256 (1-4 qubits)", dim);
    if (n_rows < 0 || n_elements < 0 || n_ops < 0 || n_rho <= 0 || n_eff <= 0 || cache_size < 0)
        return fail(B200_E_INVALID, "negative/zero size argument");
    if (n_rows > 0 && (!row_ptr || !row_istart || !row_prep || !row_icache || !out_ptr))
        return fail(B200_E_INVALID, "NULL table pointer");
    if (n_rows >= (int64_t)1 << 31) return fail(B200_E_UNSUPPORTED, "too many rows");
    CU(cudaSetDevice(ctx->device));

    // ---- expand: full op sequence of every row; cache slot -> producing row -------------------
    std::vector<int64_t> slot_row((size_t)cache_size, -1);
    std::vector<uint64_t> xptr((size_t)n_rows + 1, 0);
    std::vector<int32_t> xprep((size_t)n_rows);
    // first pass: lengths
    for (int64_t k = 0; k < n_rows; ++k) {
        if (row_ptr[k + 1] < row_ptr[k]) return fail(B200_E_INVALID, "row_ptr not monotone at row %lld", (long long)k);
        int64_t rem = row_ptr[k + 1] - row_ptr[k];
        int64_t len;
        if (row_istart[k] < 0) {
            if (row_prep[k] < 0 || row_prep[k] >= n_rho) return fail(B200_E_INVALID, "row %lld: bad prep index %d", (long long)k, row_prep[k]);
            len = rem; xprep[k] = row_prep[k];
        } else {
            if (row_istart[k] >= cache_size || slot_row[row_istart[k]] < 0)
                return fail(B200_E_INVALID, "row %lld starts from cache slot %d which no earlier row wrote", (long long)k, row_istart[k]);
            int64_t src = slot_row[row_istart[k]];
            len = (int64_t)(xptr[src + 1] - xptr[src]) + rem; xprep[k] = xprep[src];
        }
        xptr[k + 1] = xptr[k] + (uint64_t)len;
        if (row_icache[k] >= 0) {
            if (row_icache[k] >= cache_size) return fail(B200_E_INVALID, "row %lld: cache slot %d >= cache_size", (long long)k, row_icache[k]);
            slot_row[row_icache[k]] = k;
        }
        if (out_ptr[k + 1] < out_ptr[k]) return fail(B200_E_INVALID, "out_ptr not monotone");
    }
    if (xptr[n_rows] >= ((uint64_t)1 << 32)) return fail(B200_E_UNSUPPORTED, "expanded layout has >= 2^32 propagations");
    std::vector<int32_t> xops((size_t)xptr[n_rows]);
    std::fill(slot_row.begin(), slot_row.end(), -1);
    int max_depth = 0;
    for (int64_t k = 0; k < n_rows; ++k) {
        int32_t* dst = xops.data() + xptr[k];
        if (row_istart[k] >= 0) {
            int64_t src = slot_row[row_istart[k]];
            size_t n = (size_t)(xptr[src + 1] - xptr[src]);
            if (n) memcpy(dst, xops.data() + xptr[src], n * sizeof(int32_t));
            dst += n;
        }
        for (int32_t t = row_ptr[k]; t < row_ptr[k + 1]; ++t) {
            if (row_ops[t] < 0 || row_ops[t] >= n_ops) return fail(B200_E_INVALID, "row %lld: op index %d out of range", (long long)k, row_ops[t]);
            *dst++ = row_ops[t];
        }
        if (row_icache[k] >= 0) slot_row[row_icache[k]] = k;
        max_depth = std::max<int>(max_depth, (int)(xptr[k + 1] - xptr[k]));
    }
    int64_t n_out = n_rows ? out_ptr[n_rows] : 0;
    for (int64_t t = 0; t < n_out; ++t) {
        if (out_eff[t] < 0 || out_eff[t] >= n_eff) return fail(B200_E_INVALID, "effect index %d out of range", out_eff[t]);
        if (out_el[t] < 0 || out_el[t] >= n_elements) return fail(B200_E_INVALID, "element index %d out of range", out_el[t]);
    }

    // ---- order circuits longest-first (tail of the persistent grid = short circuits) -----------
    std::vector<int64_t> order((size_t)n_rows);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
        return (xptr[x + 1] - xptr[x]) > (xptr[y + 1] - xptr[y]); });
    std::vector<uint32_t> cptr((size_t)n_rows + 1, 0);
    std::vector<int32_t> cops(xops.size()), cprep((size_t)n_rows), coptr((size_t)n_rows + 1, 0), coeff((size_t)n_out), coel((size_t)n_out);
    for (int64_t i = 0; i < n_rows; ++i) {
        int64_t k = order[i];
        size_t n = (size_t)(xptr[k + 1] - xptr[k]);
        if (n) memcpy(cops.data() + cptr[i], xops.data() + xptr[k], n * sizeof(int32_t));
        cptr[i + 1] = cptr[i] + (uint32_t)n;
        cprep[i] = xprep[k];
        int32_t no = out_ptr[k + 1] - out_ptr[k];
        memcpy(coeff.data() + coptr[i], out_eff + out_ptr[k], (size_t)no * sizeof(int32_t));
        memcpy(coel.data() + coptr[i], out_el + out_ptr[k], (size_t)no * sizeof(int32_t));
        coptr[i + 1] = coptr[i] + no;
    }

    // per circuit: the steps sorted by gate ("gate buckets") -- the order the accumulate kernels (d = 16 trie path, d >= 64
    // level path) walk them in
    std::vector<uint16_t> bperm, bcnt;
    if ((dim == 16 || dim >= 64) && max_depth < 65535) {
        bperm.resize(cops.size()); bcnt.assign((size_t)n_rows * std::max(n_ops, 1), 0);
        std::vector<uint32_t> off((size_t)std::max(n_ops, 1) + 1);
        for (int64_t i = 0; i < n_rows; ++i) {
            const uint32_t b0 = cptr[i], L = cptr[i + 1] - b0;
            uint16_t* cn = bcnt.data() + (size_t)i * std::max(n_ops, 1);
            for (uint32_t k = 0; k < L; ++k) cn[cops[b0 + k]]++;
            off[0] = 0;
            for (int g = 0; g < n_ops; ++g) off[g + 1] = off[g] + cn[g];
            for (uint32_t k = 0; k < L; ++k) bperm[b0 + off[cops[b0 + k]]++] = (uint16_t)k;
        }
    }

    b200_atom* a = new b200_atom();
    a->ctx = ctx; a->dim = dim; a->n_ops = n_ops; a->n_rho = n_rho; a->n_eff = n_eff;
    a->n_rows = n_rows; a->n_elements = n_elements;
    a->n_prop_table = n_rows ? row_ptr[n_rows] : 0;
    a->n_prop_expanded = (int64_t)xptr[n_rows];
    a->max_depth = max_depth;
    a->off_rho = (int64_t)n_ops * dim * dim;
    a->off_eff = a->off_rho + (int64_t)n_rho * dim;
    a->n_w = a->off_eff + (int64_t)n_eff * dim;
    int rc;
    if ((rc = upload_vec(a->circ_ptr, cptr, ctx->stream)) || (rc = upload_vec(a->circ_ops, cops, ctx->stream)) ||
        (rc = upload_vec(a->circ_prep, cprep, ctx->stream)) || (rc = upload_vec(a->out_ptr, coptr, ctx->stream)) ||
        (rc = upload_vec(a->out_eff, coeff, ctx->stream)) || (rc = upload_vec(a->out_el, coel, ctx->stream)) ||
        (rc = upload_vec(a->bperm, bperm, ctx->stream)) ||
        (rc = upload_vec(a->bcnt, bcnt, ctx->stream))) {
        b200_atom_free(ctx, a); return rc;
    }
    // level-batched dense path (d >= 64): per depth level, circuits grouped by the gate of that level, in tiles of 32
    if (dim >= 64 && n_rows > 0 && n_ops <= 65535) {
        std::vector<uint32_t> lc; lc.reserve(cops.size());
        std::vector<LevelTile> tiles;
        a->lvl_tile_ptr.assign((size_t)max_depth + 1, 0);
        std::vector<std::vector<uint32_t>> by_gate((size_t)n_ops);
        // circuits are stored longest first: those with depth > k are a prefix [0, n_k)
        int64_t n_k = n_rows;
        for (int k = 0; k < max_depth; ++k) {
            while (n_k > 0 && (int)(cptr[n_k] - cptr[n_k - 1]) <= k) --n_k;
            for (auto& v : by_gate) v.clear();
            for (int64_t i = 0; i < n_k; ++i) by_gate[cops[cptr[i] + k]].push_back((uint32_t)i);
            a->lvl_tile_ptr[k] = (uint32_t)tiles.size();
            for (int g = 0; g < n_ops; ++g) {
                const auto& v = by_gate[g];
                for (size_t o = 0; o < v.size(); o += 32) {
                    LevelTile tl; tl.first = (uint32_t)lc.size() + (uint32_t)o; tl.count = (uint16_t)std::min<size_t>(32, v.size() - o);
                    tl.gate = (uint16_t)g; tiles.push_back(tl);
                }
                lc.insert(lc.end(), v.begin(), v.end());
            }
        }
        a->lvl_tile_ptr[max_depth] = (uint32_t)tiles.size();
        if ((rc = upload_vec(a->lvl_circ, lc, ctx->stream)) || (rc = upload_vec(a->lvl_tiles, tiles, ctx->stream))) { b200_atom_free(ctx, a); return rc; }
        CU(cudaStreamSynchronize(ctx->stream));
        a->has_levels = true;
        // Jacobian tables: every level is kept.  FS rows fbase[i] + k (k = 0..L), BH rows bbase[i] + o (L+1) + k.
        uint64_t rows_f = 0, rows_b = 0; int no_max = 0;
        std::vector<uint32_t> fbase((size_t)n_rows), bbase((size_t)n_rows);
        for (int64_t i = 0; i < n_rows; ++i) {
            const uint64_t L = cptr[i + 1] - cptr[i], no = (uint64_t)(coptr[i + 1] - coptr[i]);
            fbase[i] = (uint32_t)rows_f; bbase[i] = (uint32_t)rows_b;
            rows_f += L + 1; rows_b += no * (L + 1);
            no_max = std::max(no_max, (int)no);
        }
        if (rows_f < ((uint64_t)1 << 32) - 64 && rows_b < ((uint64_t)1 << 32) - 64 && !bperm.empty() && no_max >= 1) {
            // forward sweep reuses the level tiles; frow[e] = FS row of the state entering level k for lvl_circ[e]
            std::vector<uint32_t> frow(lc.size());
            {
                size_t e = 0; int64_t nk = n_rows;
                for (int k = 0; k < max_depth; ++k) {
                    while (nk > 0 && (int)(cptr[nk] - cptr[nk - 1]) <= k) --nk;
                    for (int64_t j = 0; j < nk; ++j, ++e) frow[e] = fbase[lc[e]] + (uint32_t)k;
                }
            }
            // backward sweep: level m handles step k = L-1-m of every circuit with L > m; rows are (circuit, outcome)
            std::vector<uint32_t> brow; brow.reserve((size_t)std::min<uint64_t>(rows_b, (uint64_t)1 << 31));
            std::vector<LevelTile> btiles;
            a->lj_btile_ptr.assign((size_t)max_depth + 1, 0);
            std::vector<std::vector<uint32_t>> rows_by_gate((size_t)n_ops);
            int64_t nm = n_rows;
            for (int mlev = 0; mlev < max_depth; ++mlev) {
                while (nm > 0 && (int)(cptr[nm] - cptr[nm - 1]) <= mlev) --nm;
                for (auto& v : rows_by_gate) v.clear();
                for (int64_t i = 0; i < nm; ++i) {
                    const uint32_t L = cptr[i + 1] - cptr[i];
                    const int g = cops[cptr[i] + (L - 1 - (uint32_t)mlev)];
                    const uint32_t no = (uint32_t)(coptr[i + 1] - coptr[i]);
                    for (uint32_t o = 0; o < no; ++o) rows_by_gate[g].push_back(bbase[i] + o * (L + 1) + (L - (uint32_t)mlev));
                }
                a->lj_btile_ptr[mlev] = (uint32_t)btiles.size();
                for (int g = 0; g < n_ops; ++g) {
                    const auto& v = rows_by_gate[g];
                    for (size_t o = 0; o < v.size(); o += 32) {
                        LevelTile tl; tl.first = (uint32_t)(brow.size() + o); tl.count = (uint16_t)std::min<size_t>(32, v.size() - o);
                        tl.gate = (uint16_t)g; btiles.push_back(tl);
                    }
                    brow.insert(brow.end(), v.begin(), v.end());
                }
            }
            a->lj_btile_ptr[max_depth] = (uint32_t)btiles.size();
            if ((rc = upload_vec(a->lj_fbase, fbase, ctx->stream)) || (rc = upload_vec(a->lj_bbase, bbase, ctx->stream)) ||
                (rc = upload_vec(a->lj_frow, frow, ctx->stream)) || (rc = upload_vec(a->lj_brow, brow, ctx->stream)) ||
                (rc = upload_vec(a->lj_btiles, btiles, ctx->stream))) { b200_atom_free(ctx, a); return rc; }
            CU(cudaStreamSynchronize(ctx->stream));
            a->lj_rows_f = rows_f; a->lj_rows_b = rows_b; a->lj_no_max = no_max;
            a->has_lj = brow.size() < ((size_t)1 << 32);
        }
    }
    // trie path tables (d = 16, <= 8 effects, <= 255 ops): prefix trie of (prep, ops), suffix trie of reversed ops
    if (dim == 16 && n_eff <= 8 && n_ops <= 255 && n_ops >= 1 && !bperm.empty() && xptr[n_rows] < ((uint64_t)1 << 31)) {
        TrieHost TF, TB;
        std::vector<int32_t> zero_root((size_t)n_rows, 0);
        build_trie(n_rows, cprep, cptr, cops, false, TF);
        build_trie(n_rows, zero_root, cptr, cops, true, TB);
        if (TF.node_op.size() < ((size_t)1 << 31) && TB.node_op.size() < ((size_t)1 << 31)) {
            std::vector<uint32_t> fn(cops.size()), bn(cops.size()), fend((size_t)n_rows), bend((size_t)n_rows);
            for (int64_t i = 0; i < n_rows; ++i) {
                const uint32_t b0 = cptr[i], L = cptr[i + 1] - b0;
                for (uint32_t t = 0; t < L; ++t) {
                    const uint32_t k = bperm[b0 + t];                       // step index in gate-bucket order
                    fn[b0 + t] = TF.depth_node[TF.dptr[i] + k];             // s_k      : prefix of length k
                    bn[b0 + t] = TB.depth_node[TB.dptr[i] + (L - 1 - k)];   // e_k      : suffix of length L-1-k
                }
                fend[i] = TF.depth_node[TF.dptr[i] + L];
                bend[i] = TB.depth_node[TB.dptr[i] + L];
            }
            // phase-B units: (circuit, outcome group of <= 4 effects, gate), in suffix-lexicographic circuit order
            // (consecutive units then gather from the same region of the large backward table H, and the small
            // forward table S stays L2 resident).  `uidx` is one contiguous stream in unit order of
            // (S element offset, H element offset incl. the group's effect base); buckets padded to 4 steps with the
            // offsets of the all-zero rows.
            const uint32_t zf_node = (uint32_t)TF.node_op.size(), zb_node = (uint32_t)TB.node_op.size();   // all-zero rows
            const uint32_t ne16 = (uint32_t)n_eff * 16u;
            const int gsz = AT_NO;                                         // outcomes per phase-B unit
            std::vector<UnitRec> units; std::vector<uint2> uidx;
            bool trie_ok = ((uint64_t)(zb_node + 2) * ne16 < ((uint64_t)1 << 32)) && ((uint64_t)(zf_node + 2) * 16 < ((uint64_t)1 << 32));
            // unit order: suffix-lexicographic (consecutive units gather from the same region of H)
            const std::vector<int64_t>& uorder = TB.sorted;
            for (int64_t si = 0; si < n_rows; ++si) {
                const int64_t i = uorder[si];
                const uint32_t b0 = cptr[i];
                const uint16_t* cn = bcnt.data() + (size_t)i * n_ops;
                for (int eb = 0; eb < n_eff; eb += gsz) {
                    int32_t gel[4] = {-1, -1, -1, -1};
                    bool any = false;
                    for (int32_t oq = coptr[i]; oq < coptr[i + 1]; ++oq) {
                        const int e = coeff[oq];
                        if (e >= eb && e < eb + gsz) {
                            if (gel[e - eb] >= 0) trie_ok = false;         // same effect twice in one circuit: not representable
                            gel[e - eb] = coel[oq]; any = true;
                        }
                    }
                    if (!any) continue;
                    if ((uint32_t)cprep[i] > UNIT_MAX_PREP || eb > 7) trie_ok = false;      // (packed into the unit record)
                    uint32_t tbase = 0;
                    for (int g = 0; g < n_ops; ++g) {
                        UnitRec un; memset(&un, 0, sizeof un);
                        for (int o = 0; o < 4; ++o) un.el[o] = gel[o];
                        const uint32_t ng = (cn[g] + 3u) / 4u;
                        if (ng > UNIT_MAX_GROUPS) trie_ok = false;
                        un.off = (uint32_t)uidx.size();
                        un.g_ng = (uint32_t)g | ((ng & UNIT_MAX_GROUPS) << 8) | ((uint32_t)(eb & 7) << 22) | (((uint32_t)cprep[i] & UNIT_MAX_PREP) << 25);
                        un.f_end = fend[i]; un.b_end = bend[i];
                        for (uint32_t t = 0; t < ng * 4u; ++t) {
                            const bool ok = t < cn[g];
                            uint2 e;
                            e.x = (ok ? fn[b0 + tbase + t] : zf_node) * 16u;
                            e.y = (ok ? bn[b0 + tbase + t] : zb_node) * ne16 + (uint32_t)eb * 16u;
                            uidx.push_back(e);
                        }
                        tbase += cn[g];
                        units.push_back(un);
                    }
                }
            }
            if (uidx.size() >= ((size_t)1 << 32) - 64) trie_ok = false;
            for (int pad = 0; pad < 16; ++pad) uidx.push_back(make_uint2(zf_node * 16u, zb_node * ne16));   // prefetch slack
            if (uidx.empty()) uidx.push_back(make_uint2(zf_node, zb_node));
            a->n_units = (int)units.size();
            if ((rc = upload_vec(a->t_units, units, ctx->stream)) || (rc = upload_vec(a->t_uidx, uidx, ctx->stream))) { b200_atom_free(ctx, a); return rc; }
            a->n_fchains = (int)TF.chain_first.size(); a->n_bchains = (int)TB.chain_first.size();
            a->n_fnodes = (uint32_t)TF.node_op.size(); a->n_bnodes = (uint32_t)TB.node_op.size();
            std::vector<unsigned> zc(4, 0u);
            auto pack_meta = [](const TrieHost& T) {
                std::vector<int4> mt(T.chain_first.size());
                for (size_t i = 0; i < mt.size(); ++i) mt[i] = make_int4(T.chain_parent[i], (int)T.chain_first[i], (int)T.chain_len[i], 0);
                return mt; };
            const std::vector<int4> fmeta = pack_meta(TF), bmeta = pack_meta(TB);
            auto parents = [](const TrieHost& T) {                          // distinct parent nodes of the chain heads
                std::vector<uint32_t> v;
                for (int32_t pnode : T.chain_parent) if (pnode >= 0) v.push_back((uint32_t)pnode);
                std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
                return v; };
            const std::vector<uint32_t> fpar = parents(TF), bpar = parents(TB);
            a->n_fpar = (uint32_t)fpar.size(); a->n_bpar = (uint32_t)bpar.size();
            if ((rc = upload_vec(a->tf_par, fpar, ctx->stream)) || (rc = upload_vec(a->tb_par, bpar, ctx->stream))) { b200_atom_free(ctx, a); return rc; }
            TF.node_op.resize(TF.node_op.size() + 32, 0); TB.node_op.resize(TB.node_op.size() + 32, 0);   // 32-wide op fetches
            if ((rc = upload_vec(a->tf_meta, fmeta, ctx->stream)) || (rc = upload_vec(a->tf_op, TF.node_op, ctx->stream)) ||
                (rc = upload_vec(a->tb_meta, bmeta, ctx->stream)) || (rc = upload_vec(a->tb_op, TB.node_op, ctx->stream)) ||
                (rc = upload_vec(a->t_fn, fn, ctx->stream)) || (rc = upload_vec(a->t_bn, bn, ctx->stream)) ||
                (rc = upload_vec(a->t_fend, fend, ctx->stream)) || (rc = upload_vec(a->t_bend, bend, ctx->stream)) ||
                (rc = upload_vec(a->t_counters, zc, ctx->stream))) { b200_atom_free(ctx, a); return rc; }
            a->t_S_bytes = ((size_t)a->n_fnodes + 1) * 128;                                              // incl. the zero rows
            a->t_H_bytes = (((size_t)a->n_bnodes + 2) * n_eff * 128 + 255) & ~(size_t)255;
            if (a->t_SH.ensure(a->t_S_bytes + a->t_H_bytes) != cudaSuccess) { b200_atom_free(ctx, a); return fail(B200_E_NOMEM, "trie tables: out of device memory"); }
            CU(cudaMemsetAsync(a->t_SH.p, 0, a->t_S_bytes + a->t_H_bytes, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            a->has_trie = trie_ok;
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
    *out = a;
    return B200_OK;
}

extern "C" int b200_atom_free(b200_ctx* ctx, b200_atom* a) {
    if (!a) return B200_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    DevBuf* bufs[] = {&a->circ_ptr, &a->circ_ops, &a->circ_prep, &a->out_ptr, &a->out_eff, &a->out_el, &a->M, &a->Gt, &a->fac_ptr, &a->fac_rec, &a->fac_mats,
                      &a->bperm, &a->bcnt, &a->lvl_circ, &a->lvl_tiles,
                      &a->lj_fbase, &a->lj_bbase, &a->lj_frow, &a->lj_brow, &a->lj_btiles, &a->lj_ti_ptr, &a->lj_items,
                      &a->lj2_ti_ptr, &a->lj2_mask, &a->lj2_items, &a->lj2_ij, &a->lj2_v, &a->lj3_tp, &a->lj3_mask, &a->lj3_items, &a->lj3_ij, &a->lj3_v, &a->tf_meta, &a->tf_op, &a->tb_meta, &a->tb_op, &a->t_fn, &a->t_bn, &a->t_fend, &a->t_bend, &a->tf_par, &a->tb_par, &a->t_SH,
                      &a->t_counters, &a->t_units, &a->t_uidx,
                      &a->cptr, &a->crow, &a->cval, &a->Dd, &a->kt_ptr, &a->kt_idx, &a->colmap, &a->spam_col, &a->spam_w,
                      &a->aff_rptr, &a->aff_rcol, &a->aff_rval, &a->aff_const, &a->aff_theta,
                      &a->id_colmap, &a->id_spam_col, &a->id_spam_w,
                      &a->fj_base, &a->fj_out_circ, &a->fj_fao, &a->fj_ffo, &a->fj_cptr, &a->fj_ccode, &a->fj_cval, &a->fj_slots, &a->fj_steps, &a->fac_ffo};
    for (DevBuf* b : bufs) b->release();
    delete a;
    return B200_OK;
}

extern "C" int b200_atom_info(b200_atom* a, int64_t info[8]) {
    if (!a || !info) return fail(B200_E_INVALID, "NULL argument");
    info[0] = a->n_rows; info[1] = a->n_elements; info[2] = a->n_prop_table; info[3] = a->n_prop_expanded;
    info[4] = a->max_depth; info[5] = a->n_w; info[6] = a->has_derivs ? a->n_params : -1;
    info[7] = (a->has_derivs && a->unit_perm) ? 1 : 0;
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// model / derivative upload
// ------------------------------------------------------------------------------------------------
extern "C" int b200_atom_set_model(b200_ctx* ctx, b200_atom* a, const double* G, const double* rho, const double* E) {
    if (!ctx || !a || !rho || !E || (a->n_ops > 0 && !G)) return fail(B200_E_INVALID, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    const int d = a->dim;
    CU(a->M.ensure((size_t)a->n_w * sizeof(double)));
    CU(a->Gt.ensure(std::max<size_t>((size_t)a->off_rho * sizeof(double), 16)));
    double* M = a->M.as<double>();
    if (a->n_ops) CU(cudaMemcpyAsync(M, G, (size_t)a->off_rho * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(M + a->off_rho, rho, (size_t)a->n_rho * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(M + a->off_eff, E, (size_t)a->n_eff * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (a->n_ops) {
        dim3 grid((unsigned)std::min<int64_t>((a->off_rho + 255) / 256, 1024), 1);
        k_transpose_gates<<<grid, 256, 0, ctx->stream>>>(M, a->n_w, a->n_ops, d, a->Gt.as<double>());
        ctx->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->stream));   // caller may free/modify its host arrays on return
    a->has_model = true; a->has_factored = false;
    return B200_OK;
}

template <int D>
static int launch_factored_dense(b200_ctx* c, b200_atom* a, const FactoredDev& fd) {
    const size_t smem = (size_t)FAC_WARPS * 2 * D * 8;
    CU(cudaFuncSetAttribute(k_factored_to_dense<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)a->n_ops * D + FAC_WARPS - 1) / FAC_WARPS, (int64_t)c->sm_count * 8));
    k_factored_to_dense<D><<<grid, FAC_WARPS * 32, smem, c->stream>>>(a->n_ops, fd, a->M.as<double>(), a->Gt.as<double>());
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}
static FactoredDev factored_dev(b200_atom* a) {
    FactoredDev fd; fd.op_fptr = a->fac_ptr.as<int32_t>(); fd.fac = a->fac_rec.as<FactorRec>(); fd.mats = a->fac_mats.as<double>();
    return fd;
}

extern "C" int b200_atom_set_model_factored(b200_ctx* ctx, b200_atom* a, int32_t n_factors, const int32_t* op_fptr,
                                            const int32_t* f_nq, const int32_t* f_targets, const int64_t* f_moff,
                                            const double* mats, int64_t n_mats, const double* rho, const double* E) {
    if (!ctx || !a || !op_fptr || !rho || !E || n_factors < 0 || (n_factors > 0 && (!f_nq || !f_targets || !f_moff || !mats)))
        return fail(B200_E_INVALID, "NULL / negative argument");
    const int d = a->dim;
    int nq = 0; while ((1 << (2 * nq)) < d) ++nq;
    if ((1 << (2 * nq)) != d || (d != 16 && d != 64 && d != 256)) return fail(B200_E_UNSUPPORTED, "factored models need dim = 16, 64 or 256");
    if (op_fptr[0] != 0 || op_fptr[a->n_ops] != n_factors) return fail(B200_E_INVALID, "op_fptr must run from 0 to n_factors");
    std::vector<FactorRec> recs((size_t)std::max(n_factors, 1));
    for (int g = 0; g < a->n_ops; ++g) if (op_fptr[g + 1] < op_fptr[g]) return fail(B200_E_INVALID, "op_fptr not monotone");
    for (int f = 0; f < n_factors; ++f) {
        const int k = f_nq[f];
        if (k < 1 || k > 2) return fail(B200_E_UNSUPPORTED, "factor %d acts on %d qubits (1 or 2 supported)", f, k);
        const int64_t ds = (int64_t)1 << (2 * k);
        if (f_moff[f] < 0 || f_moff[f] + ds * ds > n_mats || f_moff[f] >= ((int64_t)1 << 31)) return fail(B200_E_INVALID, "factor %d: matrix offset out of range", f);
        FactorRec r; r.nq = k; r.shift[0] = r.shift[1] = 0; r.moff = (int32_t)f_moff[f];
        for (int t = 0; t < k; ++t) {
            const int q = f_targets[4 * f + t];
            if (q < 0 || q >= nq || (t == 1 && q == f_targets[4 * f])) return fail(B200_E_INVALID, "factor %d: bad target qubit %d", f, q);
            r.shift[t] = 2 * (nq - 1 - q);
        }
        recs[f] = r;
    }
    CU(cudaSetDevice(ctx->device));
    std::vector<int32_t> fptr(op_fptr, op_fptr + a->n_ops + 1);
    std::vector<double> mv(mats, mats + std::max<int64_t>(n_mats, 0)); if (mv.empty()) mv.push_back(0.0);
    int rc;
    if ((rc = upload_vec(a->fac_ptr, fptr, ctx->stream)) || (rc = upload_vec(a->fac_rec, recs, ctx->stream)) ||
        (rc = upload_vec(a->fac_mats, mv, ctx->stream))) return rc;
    CU(a->M.ensure((size_t)a->n_w * sizeof(double)));
    CU(a->Gt.ensure(std::max<size_t>((size_t)a->off_rho * sizeof(double), 16)));
    double* M = a->M.as<double>();
    CU(cudaMemcpyAsync(M + a->off_rho, rho, (size_t)a->n_rho * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(M + a->off_eff, E, (size_t)a->n_eff * d * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (a->n_ops) {
        const FactoredDev fd = factored_dev(a);
        rc = d == 16 ? launch_factored_dense<16>(ctx, a, fd) : d == 64 ? launch_factored_dense<64>(ctx, a, fd) : launch_factored_dense<256>(ctx, a, fd);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    a->has_model = true; a->has_factored = true; a->fac_n_mats = (int)std::min<int64_t>(n_mats, INT_MAX); a->fac_n = n_factors;
    a->h_fac.assign(recs.begin(), recs.begin() + n_factors); a->h_fptr = fptr;
    {
        std::vector<int32_t> ffo((size_t)std::max(n_factors, 1), 0); int nf = 0;
        for (int f = 0; f < n_factors; ++f) { ffo[f] = nf; nf += recs[f].nq == 2 ? 256 : 32; }
        if ((rc = upload_vec(a->fac_ffo, ffo, ctx->stream))) return rc;
        CU(cudaStreamSynchronize(ctx->stream));
        a->fac_n_frag = nf;
    }
    a->fac_probs_ok = n_mats <= FAC_MATS_MAX && n_factors <= FAC_RECS_MAX && a->n_ops <= 4096;
    return B200_OK;
}

static bool same_factor_structure(const std::vector<FactorRec>& x, const std::vector<FactorRec>& y) {
    if (x.size() != y.size()) return false;
    for (size_t i = 0; i < x.size(); ++i)
        if (x[i].nq != y[i].nq || x[i].shift[0] != y[i].shift[0] || x[i].shift[1] != y[i].shift[1] || x[i].moff != y[i].moff) return false;
    return true;
}

// Derivative map in FACTOR space: rows index [factor matrices (the `mats` array of b200_atom_set_model_factored) | rho | E].
extern "C" int b200_atom_set_derivs_factored(b200_ctx* ctx, b200_atom* a, int64_t n_wf, int32_t n_params, int64_t nnz,
                                             const int32_t* rows, const int32_t* cols, const double* vals) {
    if (!ctx || !a || n_params < 0 || nnz < 0 || (nnz > 0 && (!rows || !cols || !vals))) return fail(B200_E_INVALID, "NULL / negative argument");
    if (!a->has_factored) return fail(B200_E_STATE, "b200_atom_set_model_factored has not been called");
    const int d = a->dim;
    if (d != 64 && d != 256) return fail(B200_E_UNSUPPORTED, "factor-space derivatives need dim = 64 or 256");
    const int64_t n_mats = a->fac_n_mats;
    if (n_wf != n_mats + (int64_t)(a->n_rho + a->n_eff) * d) return fail(B200_E_INVALID, "n_wf=%lld, expected %lld", (long long)n_wf, (long long)(n_mats + (int64_t)(a->n_rho + a->n_eff) * d));
    if (a->n_eff >= (1 << 13) || a->n_rho >= (1 << 13)) return fail(B200_E_UNSUPPORTED, "too many preps / effects");
    if (a->has_derivs && a->n_params != n_params) return fail(B200_E_INVALID, "factor-space map has %d parameters, the dense map %d", n_params, a->n_params);
    const int n_fac = a->fac_n;
    // accumulator / fragment offsets, owners of every matrix element (factors may share a matrix: their accumulators add up)
    std::vector<int32_t> fao((size_t)std::max(n_fac, 1)), ffo((size_t)std::max(n_fac, 1)), first((size_t)std::max<int64_t>(n_mats, 1), -1), next((size_t)std::max(n_fac, 1), -1);
    int n_acc = 0, n_frag = 0;
    for (int f = 0; f < n_fac; ++f) {
        const FactorRec& r = a->h_fac[f];
        const int ds2 = r.nq == 2 ? 256 : 16;
        fao[f] = n_acc; ffo[f] = n_frag; n_acc += (r.nq == 1 && d == 64) ? 32 : ds2; n_frag += r.nq == 2 ? 256 : 32;
        const int head = first[r.moff];
        if (head >= 0 && (a->h_fac[head].nq != r.nq)) return fail(B200_E_UNSUPPORTED, "factors of different size share a matrix");
        for (int i = 0; i < ds2; ++i) {
            if (first[r.moff + i] != head) return fail(B200_E_UNSUPPORTED, "factor matrices overlap partially");
        }
        next[f] = head;
        for (int i = 0; i < ds2; ++i) first[r.moff + i] = f;
    }
    if (n_acc >= (1 << 16) || n_frag >= (1 << 16)) return fail(B200_E_UNSUPPORTED, "factor programs too large for the factored Jacobian");
    // d = 64: the accumulators of the first FJ64_REG_SLOTS two-qubit factors live in registers; a 1-qubit factor keeps two partial
    // 4 x 4 blocks (one per half of the rest index) that the contraction adds
    std::vector<int32_t> slots((size_t)std::max(n_fac, 1), 0xFF);
    for (int s2 = 0; s2 < FJ64_REG_SLOTS; ++s2) a->fj_slot_fao[s2] = -1;
    if (d == 64) {
        int ns = 0;
        for (int f = 0; f < n_fac && ns < FJ64_REG_SLOTS; ++f) if (a->h_fac[f].nq == 2) { slots[f] = ns; a->fj_slot_fao[ns] = fao[f]; ++ns; }
    }
    const int dup1 = d == 64 ? 2 : 1;
    // CSC with element codes
    std::vector<int32_t> cptr((size_t)n_params + 1, 0);
    auto n_owner = [&](int64_t w) { int n = 0; for (int f = first[w]; f >= 0; f = next[f]) n += a->h_fac[f].nq == 1 ? dup1 : 1; return n; };
    for (int64_t t = 0; t < nnz; ++t) {
        const int64_t w = rows[t]; const int p = cols[t];
        if (w < 0 || w >= n_wf || p < 0 || p >= n_params) return fail(B200_E_INVALID, "entry %lld out of range", (long long)t);
        cptr[p + 1] += w < n_mats ? n_owner(w) : 1;
    }
    for (int p = 0; p < n_params; ++p) cptr[p + 1] += cptr[p];
    const int64_t tot = cptr[n_params];
    std::vector<uint32_t> ccode((size_t)std::max<int64_t>(tot, 1)); std::vector<double> cval((size_t)std::max<int64_t>(tot, 1));
    std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
    for (int64_t t = 0; t < nnz; ++t) {
        const int64_t w = rows[t]; const int p = cols[t];
        if (w < n_mats) {
            for (int f = first[w]; f >= 0; f = next[f]) {
                const FactorRec& r = a->h_fac[f];
                const int e = (int)(w - r.moff);
                uint32_t code;
                if (r.nq == 2) { const int ra = e >> 4, cb = e & 15; code = (uint32_t)(fao[f] + ((((ra >> 3) * 2 + (cb >> 3)) * 32 + (((ra & 7) << 2) | ((cb & 7) >> 1))) * 2) + (cb & 1)); }
                else {
                    code = (uint32_t)(fao[f] + e);
                    if (dup1 == 2) { ccode[fill[p]] = code + 16u; cval[fill[p]++] = vals[t]; }
                }
                ccode[fill[p]] = code; cval[fill[p]++] = vals[t];
            }
        } else {
            const int64_t v = w - n_mats;
            const bool is_rho = v < (int64_t)a->n_rho * d;
            const int64_t u = is_rho ? v : v - (int64_t)a->n_rho * d;
            ccode[fill[p]] = ((is_rho ? 1u : 2u) << 30) | ((uint32_t)(u / d) << 16) | (uint32_t)(u % d);
            cval[fill[p]++] = vals[t];
        }
    }
    bool unit = true;                                   // all values +-1 (full / TP members): the sign travels in bit 29 of the code
    for (int64_t q = 0; q < tot && unit; ++q) unit = cval[q] == 1.0 || cval[q] == -1.0;
    if (unit) for (int64_t q = 0; q < tot; ++q) if (cval[q] < 0.0) ccode[q] |= 0x20000000u;
    a->fj_unit = unit ? 1 : 0;
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_vec(a->fj_fao, fao, ctx->stream)) || (rc = upload_vec(a->fj_ffo, ffo, ctx->stream)) || (rc = upload_vec(a->fj_slots, slots, ctx->stream)) ||
        (rc = upload_vec(a->fj_cptr, cptr, ctx->stream)) || (rc = upload_vec(a->fj_ccode, ccode, ctx->stream)) ||
        (rc = upload_vec(a->fj_cval, cval, ctx->stream))) return rc;
    // rows of the forward-state table: one per factor step of every circuit + the initial state
    const int64_t n_out = a->n_elements;
    CU(a->fj_base.ensure(((size_t)a->n_rows + 1) * 4 + 16));
    CU(a->fj_out_circ.ensure(std::max<size_t>((size_t)n_out * 4, 16)));
    std::vector<uint32_t> cnt((size_t)a->n_rows + 1, 0);
    if (a->n_rows > 0) {
        k_fj_count<<<(unsigned)std::min<int64_t>((a->n_rows + 255) / 256, 4096), 256, 0, ctx->stream>>>(atom_dev(a), a->fac_ptr.as<int32_t>(), a->fj_base.as<uint32_t>(), a->fj_out_circ.as<int32_t>());
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(cnt.data(), a->fj_base.p, (size_t)a->n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    uint64_t run = 0;
    for (int64_t c = 0; c < a->n_rows; ++c) { const uint32_t n = cnt[c]; cnt[c] = (uint32_t)run; run += n; }
    cnt[a->n_rows] = (uint32_t)run;
    if (run >= ((uint64_t)1 << 32)) return fail(B200_E_UNSUPPORTED, "forward-state table too large");
    CU(cudaMemcpyAsync(a->fj_base.p, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(a->fj_steps.ensure(std::max<size_t>((size_t)run * 2, 16)));
    if (n_fac >= (1 << 16)) return fail(B200_E_UNSUPPORTED, "too many factors");
    if (a->n_rows > 0) {
        k_fj_steps<<<(unsigned)std::min<int64_t>((a->n_rows + 255) / 256, 4096), 256, 0, ctx->stream>>>(atom_dev(a), a->fac_ptr.as<int32_t>(), a->fj_base.as<uint32_t>(), a->fj_steps.as<uint16_t>());
        ctx->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(ctx->stream));
    a->fj_rows = run; a->fj_n_acc = n_acc; a->fj_n_frag = n_frag; a->fj_n_params = n_params; a->fj_fac = a->h_fac;
    a->has_fderivs = true;
    if (!a->has_derivs) a->n_params = n_params;
    return B200_OK;
}

// M[w] = M_const[w] + sum_t rval[t] * theta[rcol[t]]   (one thread per W-space element, rows summed in CSR order)
__global__ void k_model_affine(int64_t n_w, const int32_t* __restrict__ rptr, const int32_t* __restrict__ rcol,
                               const double* __restrict__ rval, const double* __restrict__ Mconst,
                               const double* __restrict__ theta, double* __restrict__ M)
{
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_w; w += (int64_t)gridDim.x * blockDim.x) {
        double acc = Mconst[w];
        for (int32_t t = rptr[w]; t < rptr[w + 1]; ++t) acc = fma(rval[t], theta[rcol[t]], acc);
        M[w] = acc;
    }
}

extern "C" int b200_lindblad_members(b200_ctx* c, int d, int n_eg, const int32_t* eg_ncoeff, const int32_t* eg_npar,
                                     const double* B_re, const double* B_im, const double* c_re, const double* c_im,
                                     const double* dc_re, const double* dc_im,
                                     int n_mem, const int32_t* m_kind, const int32_t* m_eg, const double* stat,
                                     double* val_out, double* dval_out)
{
    if (!c || d <= 0 || n_eg <= 0 || n_mem < 0 || !eg_ncoeff || !eg_npar || !B_re || !B_im || !c_re || !c_im || !dc_re || !dc_im)
        return fail(B200_E_INVALID, "NULL / non-positive argument");
    if (n_mem > 0 && (!m_kind || !m_eg || !stat || !val_out || !dval_out)) return fail(B200_E_INVALID, "NULL member argument");
    CU(cudaSetDevice(c->device));
    const long long n = (long long)d * d;
    std::vector<int> ncoeff(n_eg), npar(n_eg);
    std::vector<long long> boff(n_eg), doff(n_eg), poff((size_t)n_eg + 1, 0);
    long long nc_tot = 0, dc_tot = 0;
    for (int e = 0; e < n_eg; ++e) {
        if (eg_ncoeff[e] < 0 || eg_npar[e] < 0) return fail(B200_E_INVALID, "negative generator size");
        ncoeff[e] = eg_ncoeff[e]; npar[e] = eg_npar[e];
        boff[e] = nc_tot; doff[e] = dc_tot;
        nc_tot += eg_ncoeff[e]; dc_tot += (long long)eg_ncoeff[e] * eg_npar[e];
        poff[e + 1] = poff[e] + eg_npar[e];
    }
    const long long np_tot = poff[n_eg], n_rows = np_tot + n_eg;
    std::vector<int> kind(n_mem), meg(n_mem);
    std::vector<long long> soff(n_mem), vptr((size_t)n_mem + 1, 0), dptr((size_t)n_mem + 1, 0);
    long long s_tot = 0;
    for (int m = 0; m < n_mem; ++m) {
        if (m_kind[m] < 0 || m_kind[m] > 2 || m_eg[m] < 0 || m_eg[m] >= n_eg) return fail(B200_E_INVALID, "member %d: bad kind / generator", m);
        kind[m] = m_kind[m]; meg[m] = m_eg[m];
        const long long size = m_kind[m] == 0 ? n : d;
        soff[m] = s_tot; s_tot += size;
        vptr[m + 1] = vptr[m] + size; dptr[m + 1] = dptr[m] + size * eg_npar[m_eg[m]];
    }
    enum { I_NC, I_NP, I_BO, I_DO, I_PO, I_BR, I_BI, I_CR, I_CI, I_DR, I_DI, I_L, I_DL, I_E, I_DE, I_WK, I_MK, I_ME, I_ST, I_OUT };
    DevBuf* b = c->lind;
    int rc;
    if ((rc = upload_vec(b[I_NC], ncoeff, c->stream)) || (rc = upload_vec(b[I_NP], npar, c->stream)) ||
        (rc = upload_vec(b[I_BO], boff, c->stream)) || (rc = upload_vec(b[I_DO], doff, c->stream)) ||
        (rc = upload_vec(b[I_PO], poff, c->stream)) || (rc = upload_vec(b[I_MK], kind, c->stream)) ||
        (rc = upload_vec(b[I_ME], meg, c->stream))) return rc;
    auto up = [&](DevBuf& buf, const double* src, long long count) -> int {
        CU(buf.ensure(std::max<size_t>((size_t)count * 8, 16)));
        if (count > 0) CU(cudaMemcpyAsync(buf.p, src, (size_t)count * 8, cudaMemcpyHostToDevice, c->stream));
        return B200_OK;
    };
    if ((rc = up(b[I_BR], B_re, nc_tot * n)) || (rc = up(b[I_BI], B_im, nc_tot * n)) || (rc = up(b[I_CR], c_re, nc_tot)) ||
        (rc = up(b[I_CI], c_im, nc_tot)) || (rc = up(b[I_DR], dc_re, dc_tot)) || (rc = up(b[I_DI], dc_im, dc_tot)) ||
        (rc = up(b[I_ST], stat, s_tot))) return rc;
    CU(b[I_L].ensure((size_t)n_eg * n * 8)); CU(b[I_E].ensure((size_t)n_eg * n * 8));
    CU(b[I_DL].ensure(std::max<size_t>((size_t)np_tot * n * 8, 16))); CU(b[I_DE].ensure(std::max<size_t>((size_t)np_tot * n * 8, 16)));
    CU(b[I_WK].ensure(std::max((size_t)n_rows * 7 * n * 8, ((size_t)n_eg * 81 * n + (size_t)n_eg + 16) * 8)));
    // outputs + the two prefix arrays of the members in one buffer: [val | dval | vptr | dptr]
    const long long n_val = vptr[n_mem], n_dval = dptr[n_mem];
    CU(b[I_OUT].ensure(std::max<size_t>((size_t)(n_val + n_dval) * 8 + (size_t)2 * (n_mem + 1) * 8, 16)));
    double* d_val = b[I_OUT].as<double>(); double* d_dval = d_val + n_val;
    long long* d_vptr = reinterpret_cast<long long*>(d_dval + n_dval); long long* d_dptr = d_vptr + (n_mem + 1);
    CU(cudaMemcpyAsync(d_vptr, vptr.data(), (size_t)(n_mem + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_dptr, dptr.data(), (size_t)(n_mem + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    // soff travels at the tail of the static buffer's sibling: a small separate upload through the generic helper
    DevBuf soff_buf;
    if ((rc = upload_vec(soff_buf, soff, c->stream))) { soff_buf.release(); return rc; }
    LindDev a;
    a.d = d; a.n_eg = n_eg; a.n_mem = n_mem;
    a.eg_ncoeff = b[I_NC].as<int>(); a.eg_npar = b[I_NP].as<int>(); a.eg_boff = b[I_BO].as<long long>();
    a.eg_doff = b[I_DO].as<long long>(); a.eg_poff = b[I_PO].as<long long>();
    a.B_re = b[I_BR].as<double>(); a.B_im = b[I_BI].as<double>(); a.c_re = b[I_CR].as<double>(); a.c_im = b[I_CI].as<double>();
    a.dc_re = b[I_DR].as<double>(); a.dc_im = b[I_DI].as<double>();
    a.L = b[I_L].as<double>(); a.dL = b[I_DL].as<double>(); a.E = b[I_E].as<double>(); a.dE = b[I_DE].as<double>(); a.work = b[I_WK].as<double>();
    a.m_kind = b[I_MK].as<int>(); a.m_eg = b[I_ME].as<int>(); a.m_soff = soff_buf.as<long long>();
    a.stat = b[I_ST].as<double>(); a.val = d_val; a.dval = d_dval;
    const int g1 = (int)std::max<long long>(1, std::min<long long>((n_rows * n + 255) / 256, (long long)c->sm_count * 16));
    k_lind_errgen<<<g1, 256, 0, c->stream>>>(a, n_rows);
    if (d == 16) {              // warp-per-row DMMA recursion (the work area of the per-thread kernel holds T_k, E_q, X, s: 81 d^2 + 1 per generator)
        const long long n_par_rows = n_rows - n_eg;
        k_lind_gen16<<<(unsigned)n_eg, 32, 0, c->stream>>>(a);
        if (n_par_rows > 0) k_lind_dexp16<<<(unsigned)((n_par_rows + LB16_WARPS - 1) / LB16_WARPS), LB16_WARPS * 32, 0, c->stream>>>(a, n_par_rows);
        c->launches++;
    } else {
        k_lind_expm<<<(unsigned)((n_rows + 63) / 64), 64, 0, c->stream>>>(a, n_rows);
    }
    c->launches += 2;
    if (n_mem > 0) {
        const int g3 = (int)std::max<long long>(1, std::min<long long>((n_val + n_dval + 255) / 256, (long long)c->sm_count * 16));
        k_lind_compose<<<g3, 256, 0, c->stream>>>(a, n_val, n_dval, d_vptr, d_dptr);
        c->launches++;
        CU(cudaMemcpyAsync(val_out, d_val, (size_t)n_val * 8, cudaMemcpyDeviceToHost, c->stream));
        if (n_dval > 0) CU(cudaMemcpyAsync(dval_out, d_dval, (size_t)n_dval * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(c->stream);
    soff_buf.release();
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(B200_E_CUDA, "b200_lindblad_members: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return B200_OK;
}

extern "C" int b200_atom_get_model(b200_ctx* ctx, b200_atom* a, int64_t n_w, double* M_out) {
    if (!ctx || !a || !M_out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (n_w != a->n_w) return fail(B200_E_INVALID, "n_w=%lld does not match the atom's W space (%lld)", (long long)n_w, (long long)a->n_w);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(M_out, a->M.p, (size_t)a->n_w * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

extern "C" int b200_atom_bind_params(b200_ctx* ctx, b200_atom* a, int32_t n_params, const double* theta0) {
    if (!ctx || !a || (n_params > 0 && !theta0)) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (!a->has_derivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (n_params != a->n_params)
        return fail(B200_E_STATE, "the derivative map on the device has %d columns, not the %d parameters of theta0 "
                                  "(upload it for ALL parameters before binding)", a->n_params, n_params);
    CU(cudaSetDevice(ctx->device));
    const int64_t n_w = a->n_w;
    // CSC (by parameter) -> CSR (by W-space row); entries of a row in increasing parameter order
    std::vector<int32_t> rptr((size_t)n_w + 1, 0);
    for (int32_t r : a->h_crow) rptr[(size_t)r + 1]++;
    for (int64_t w = 0; w < n_w; ++w) rptr[w + 1] += rptr[w];
    std::vector<int32_t> rcol(a->h_crow.size()); std::vector<double> rval(a->h_crow.size());
    { std::vector<int32_t> pos(rptr.begin(), rptr.end() - 1);
      for (int32_t p = 0; p < n_params; ++p)
          for (int32_t t = a->h_cptr[p]; t < a->h_cptr[p + 1]; ++t) { const int32_t q = pos[a->h_crow[t]]++; rcol[q] = p; rval[q] = a->h_cval[t]; } }
    // M_const = M - D theta0 (same summation order as the kernel)
    std::vector<double> Mh((size_t)n_w);
    CU(cudaMemcpyAsync(Mh.data(), a->M.p, (size_t)n_w * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int64_t w = 0; w < n_w; ++w) {
        double lin = 0.0;
        for (int32_t t = rptr[w]; t < rptr[w + 1]; ++t) lin = fma(rval[t], theta0[rcol[t]], lin);
        Mh[w] -= lin;
    }
    int rc;
    if ((rc = upload_vec(a->aff_rptr, rptr, ctx->stream)) || (rc = upload_vec(a->aff_rcol, rcol, ctx->stream)) ||
        (rc = upload_vec(a->aff_rval, rval, ctx->stream)) || (rc = upload_vec(a->aff_const, Mh, ctx->stream))) return rc;
    CU(a->aff_theta.ensure(std::max<size_t>((size_t)n_params * 8, 16)));
    CU(cudaStreamSynchronize(ctx->stream));
    a->has_bind = true; a->bind_n_params = n_params;
    return B200_OK;
}

extern "C" int b200_atom_set_params_dev(b200_ctx* ctx, b200_atom* a, int32_t n_params, const double* d_theta) {
    if (!ctx || !a || (n_params > 0 && !d_theta)) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_bind) return fail(B200_E_STATE, "b200_atom_bind_params has not been called");
    if (n_params != a->bind_n_params) return fail(B200_E_INVALID, "n_params=%d, bound with %d", n_params, a->bind_n_params);
    CU(cudaSetDevice(ctx->device));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_w + 255) / 256, (int64_t)ctx->sm_count * 8));
    a->has_factored = false;                  // M is about to change: the factor programs no longer describe it
    k_model_affine<<<grid, 256, 0, ctx->stream>>>(a->n_w, a->aff_rptr.as<int32_t>(), a->aff_rcol.as<int32_t>(), a->aff_rval.as<double>(),
                                                  a->aff_const.as<double>(), d_theta, a->M.as<double>());
    ctx->launches++;
    if (a->n_ops) {
        dim3 gt((unsigned)std::min<int64_t>((a->off_rho + 255) / 256, 1024), 1);
        k_transpose_gates<<<gt, 256, 0, ctx->stream>>>(a->M.as<double>(), a->n_w, a->n_ops, a->dim, a->Gt.as<double>());
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_atom_set_params(b200_ctx* ctx, b200_atom* a, int32_t n_params, const double* theta) {
    if (!ctx || !a || (n_params > 0 && !theta)) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_bind) return fail(B200_E_STATE, "b200_atom_bind_params has not been called");
    if (n_params != a->bind_n_params) return fail(B200_E_INVALID, "n_params=%d, bound with %d", n_params, a->bind_n_params);
    CU(cudaSetDevice(ctx->device));
    if (n_params > 0) CU(cudaMemcpyAsync(a->aff_theta.p, theta, (size_t)n_params * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200_atom_set_params_dev(ctx, a, n_params, a->aff_theta.as<double>());
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));   // caller may modify theta on return
    return B200_OK;
}

// Per tile of GM_TN columns: the sorted list of GM_KC-row chunks of the right-hand side that contain a non-zero (host side).
// `col_rows(c, fn)` calls fn(row) for every non-zero row of column c.
template <class F>
static void build_kt_lists(int n_cols, F col_rows, std::vector<int32_t>& kt_ptr, std::vector<int32_t>& kt_idx) {
    const int n_tiles = std::max(1, (n_cols + GM_TN - 1) / GM_TN);
    kt_ptr.assign((size_t)n_tiles + 1, 0); kt_idx.clear();
    std::vector<int32_t> tmp;
    for (int tl = 0; tl < n_tiles; ++tl) {
        tmp.clear();
        for (int c = tl * GM_TN; c < std::min(n_cols, (tl + 1) * GM_TN); ++c) col_rows(c, [&](int64_t r) { tmp.push_back((int32_t)(r / GM_KC)); });
        std::sort(tmp.begin(), tmp.end()); tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        kt_idx.insert(kt_idx.end(), tmp.begin(), tmp.end());
        kt_ptr[tl + 1] = (int32_t)kt_idx.size();
    }
    if (kt_idx.empty()) kt_idx.push_back(0);
}

extern "C" int b200_atom_set_derivs(b200_ctx* ctx, b200_atom* a, int64_t n_w, int32_t n_params,
                                    int64_t nnz, const int32_t* rows, const int32_t* cols, const double* vals)
{
    if (!ctx || !a) return fail(B200_E_INVALID, "NULL argument");
    if (n_w != a->n_w) return fail(B200_E_INVALID, "n_w=%lld does not match the atom's W space (%lld)", (long long)n_w, (long long)a->n_w);
    if (n_params < 0 || nnz < 0 || (nnz > 0 && (!rows || !cols || !vals))) return fail(B200_E_INVALID, "bad derivative map");
    a->has_bind = false;                      // a parameter binding refers to the map it was made with
    if (nnz >= ((int64_t)1 << 31)) return fail(B200_E_UNSUPPORTED, "derivative map too large");
    CU(cudaSetDevice(ctx->device));
    // COO -> CSC with duplicates summed
    std::vector<int64_t> idx((size_t)nnz);
    std::iota(idx.begin(), idx.end(), 0);
    for (int64_t t = 0; t < nnz; ++t) {
        if (rows[t] < 0 || rows[t] >= n_w) return fail(B200_E_INVALID, "D row %d out of range", rows[t]);
        if (cols[t] < 0 || cols[t] >= n_params) return fail(B200_E_INVALID, "D col %d out of range", cols[t]);
    }
    std::sort(idx.begin(), idx.end(), [&](int64_t x, int64_t y) {
        return cols[x] != cols[y] ? cols[x] < cols[y] : rows[x] < rows[y]; });
    std::vector<int32_t> cptr((size_t)n_params + 1, 0), crow; std::vector<double> cval;
    crow.reserve((size_t)nnz); cval.reserve((size_t)nnz);
    std::vector<int32_t> ccol; ccol.reserve((size_t)nnz);
    for (int64_t t = 0; t < nnz; ++t) {
        int64_t i = idx[t];
        if (!crow.empty() && ccol.back() == cols[i] && crow.back() == rows[i]) cval.back() += vals[i];
        else { crow.push_back(rows[i]); ccol.push_back(cols[i]); cval.push_back(vals[i]); }
    }
    for (size_t t = 0; t < ccol.size(); ++t) cptr[ccol[t] + 1]++;
    for (int p = 0; p < n_params; ++p) cptr[p + 1] += cptr[p];

    // unit partial permutation?  every column <= 1 entry, every row <= 1 entry, all values == 1.0
    bool unit = true;
    std::vector<int32_t> colmap((size_t)n_w, -1);
    for (int p = 0; p < n_params && unit; ++p) {
        int n = cptr[p + 1] - cptr[p];
        if (n > 1) unit = false;
        else if (n == 1) {
            int t = cptr[p];
            if (cval[t] != 1.0 || colmap[crow[t]] != -1) unit = false;
            else colmap[crow[t]] = p;
        }
    }
    a->n_params = n_params; a->unit_perm = unit;
    a->h_cptr = cptr; a->h_crow = crow; a->h_cval = cval;
    int rc;
    if ((rc = upload_vec(a->cptr, cptr, ctx->stream)) || (rc = upload_vec(a->crow, crow, ctx->stream)) ||
        (rc = upload_vec(a->cval, cval, ctx->stream))) return rc;
    if (unit) {
        // columns not fed by a gate element: fed by a rho/effect element (w >= off_rho) or by nothing
        std::vector<int32_t> spam_col, spam_w;
        std::vector<char> fed((size_t)n_params, 0);
        for (int64_t w = 0; w < a->off_rho; ++w) if (colmap[w] >= 0) fed[colmap[w]] = 1;
        std::vector<int32_t> src((size_t)n_params, -1);
        for (int64_t w = a->off_rho; w < n_w; ++w) if (colmap[w] >= 0) src[colmap[w]] = (int32_t)w;
        for (int p = 0; p < n_params; ++p) if (!fed[p]) { spam_col.push_back(p); spam_w.push_back(src[p]); }
        a->n_spam = (int)spam_col.size();
        if ((rc = upload_vec(a->colmap, colmap, ctx->stream)) || (rc = upload_vec(a->spam_col, spam_col, ctx->stream)) ||
            (rc = upload_vec(a->spam_w, spam_w, ctx->stream))) return rc;
    }
    a->has_dense = false;
    if (!unit && !a->has_lj && n_params > 0) {
        // general path: J = W . D as a DMMA GEMM -- D dense on the device + K-chunk lists (kernels_gemm.cuh)
        const int64_t ldD = (n_params + 1) & ~1;
        if ((double)n_w * (double)ldD * 8.0 <= 2.0e9) {
            std::vector<int32_t> ktp, kti;
            build_kt_lists(n_params, [&](int col, auto fn) { for (int t = cptr[col]; t < cptr[col + 1]; ++t) fn(crow[t]); }, ktp, kti);
            if ((rc = upload_vec(a->kt_ptr, ktp, ctx->stream)) || (rc = upload_vec(a->kt_idx, kti, ctx->stream))) return rc;
            CU(a->Dd.ensure((size_t)n_w * ldD * 8));
            CU(cudaMemsetAsync(a->Dd.p, 0, (size_t)n_w * ldD * 8, ctx->stream));
            k_csc_to_dense<<<(unsigned)std::min(n_params, 4096), 128, 0, ctx->stream>>>(a->cptr.as<int32_t>(), a->crow.as<int32_t>(), a->cval.as<double>(),
                                                                                     nullptr, n_params, a->Dd.as<double>(), ldD);
            ctx->launches++;
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(ctx->stream));
            a->ldD = ldD; a->has_dense = true;
        }
    }
    if (a->has_lj) {
        // tile plan of the level-batched Jacobian path: per (tile of LJ_PT parameters, W-space block) the columns that
        // have non-zeros in that block, as (p_local, lo, hi) ranges of the CSC arrays (rows are sorted within a column)
        const int nb = a->n_ops + 1;                       // gate blocks + one block for the prep / effect rows
        // Parameters per tile (dev knob B200_LJ_PT).  Wider tiles were measured SLOWER on BASELINE config 3 (Np = 775: one
        // 1024-wide tile per circuit 95 ms, four 256-wide tiles 85 ms): the accumulate kernel is bound by the latency of its
        // serial (gate, outcome) phases, and narrow tiles spread a circuit's gates over more concurrent CTAs.
        int pt = LJ_PT_MIN;
        { const char* e = getenv("B200_LJ_PT"); if (e && atoi(e) >= 32) pt = std::min(LJ_PT_MAX, (atoi(e) + 31) & ~31); }
        a->lj_pt = pt;
        const int LJ_PT = pt;
        const int n_tiles = std::max(1, (n_params + LJ_PT - 1) / LJ_PT);
        const int64_t dd = (int64_t)a->dim * a->dim;
        std::vector<std::vector<uint4>> lists((size_t)n_tiles * nb);
        for (int p = 0; p < n_params; ++p) {
            int t = cptr[p];
            while (t < cptr[p + 1]) {
                const int b = (crow[t] < a->off_rho) ? (int)(crow[t] / dd) : a->n_ops;
                int t1 = t + 1;
                while (t1 < cptr[p + 1] && ((crow[t1] < a->off_rho) ? (int)(crow[t1] / dd) : a->n_ops) == b) ++t1;
                lists[(size_t)(p / LJ_PT) * nb + b].push_back(make_uint4((unsigned)(p % LJ_PT), (unsigned)t, (unsigned)t1, 0u));
                t = t1;
            }
        }
        std::vector<uint32_t> ti_ptr((size_t)n_tiles * (nb + 1));
        std::vector<uint4> items;
        for (int tl = 0; tl < n_tiles; ++tl) {
            for (int b = 0; b < nb; ++b) {
                ti_ptr[(size_t)tl * (nb + 1) + b] = (uint32_t)items.size();
                const auto& v = lists[(size_t)tl * nb + b];
                items.insert(items.end(), v.begin(), v.end());
            }
            ti_ptr[(size_t)tl * (nb + 1) + nb] = (uint32_t)items.size();
        }
        if (items.empty()) items.push_back(make_uint4(0, 0, 0, 0));
        a->lj_n_tiles = n_tiles;
        if ((rc = upload_vec(a->lj_ti_ptr, ti_ptr, ctx->stream)) || (rc = upload_vec(a->lj_items, items, ctx->stream))) return rc;
        // plan of the DMMA accumulate: gate-block non-zeros regrouped by (parameter tile, gate, 64 x 64 sub-block), with the
        // 8 x 8 tiles each group touches
        {
            const int d = a->dim, sbd = d / 64, nsb = sbd * sbd;
            struct Rec { uint64_t key; uint32_t p_local; uint16_t ij; double v; };
            std::vector<Rec> recs;
            for (int p = 0; p < n_params; ++p)
                for (int t = cptr[p]; t < cptr[p + 1]; ++t) {
                    if (crow[t] >= a->off_rho) break;                      // rows are sorted: the rest is prep / effect rows
                    const int g = (int)(crow[t] / dd); const int wl = (int)(crow[t] - (int64_t)g * dd);
                    const int i = wl / d, j = wl - i * d;
                    const int sb = (i / 64) * sbd + (j / 64);
                    Rec r; r.key = ((uint64_t)(p / LJ_PT) * a->n_ops + g) * nsb + sb; r.p_local = (uint32_t)(p % LJ_PT);
                    r.ij = (uint16_t)(((i & 63) << 8) | (j & 63)); r.v = cval[t];
                    recs.push_back(r);
                }
            std::stable_sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) {
                return x.key != y.key ? x.key < y.key : x.p_local < y.p_local; });
            const size_t n_keys = (size_t)n_tiles * a->n_ops * nsb;
            std::vector<uint32_t> tp2((size_t)n_tiles * a->n_ops * (nsb + 1), 0);
            std::vector<uint64_t> mask2(std::max<size_t>(n_keys, 1), 0);
            std::vector<uint4> items2; std::vector<uint16_t> nz_ij(recs.size()); std::vector<double> nz_v(recs.size());
            size_t r = 0;
            for (size_t key = 0; key < n_keys; ++key) {
                const size_t tg = key / nsb, sb = key % nsb;               // tg = tile * n_ops + g
                tp2[tg * (nsb + 1) + sb] = (uint32_t)items2.size();
                while (r < recs.size() && recs[r].key == key) {
                    uint4 it = make_uint4(recs[r].p_local, (unsigned)r, (unsigned)r, 0u);
                    while (r < recs.size() && recs[r].key == key && recs[r].p_local == it.x) {
                        nz_ij[r] = recs[r].ij; nz_v[r] = recs[r].v;
                        mask2[key] |= (uint64_t)1 << (8 * ((recs[r].ij >> 8) / 8) + (recs[r].ij & 0xff) / 8);
                        ++r;
                    }
                    it.z = (unsigned)r;
                    items2.push_back(it);
                }
                if (sb == (size_t)nsb - 1) tp2[tg * (nsb + 1) + nsb] = (uint32_t)items2.size();
            }
            if (items2.empty()) items2.push_back(make_uint4(0, 0, 0, 0));
            if (nz_ij.empty()) { nz_ij.push_back(0); nz_v.push_back(0.0); }
            if ((rc = upload_vec(a->lj2_ti_ptr, tp2, ctx->stream)) || (rc = upload_vec(a->lj2_mask, mask2, ctx->stream)) ||
                (rc = upload_vec(a->lj2_items, items2, ctx->stream)) || (rc = upload_vec(a->lj2_ij, nz_ij, ctx->stream)) ||
                (rc = upload_vec(a->lj2_v, nz_v, ctx->stream))) return rc;
            CU(cudaStreamSynchronize(ctx->stream));
        }
        // plan of the warp-per-outcome accumulate (k_level_accum3): gate-block non-zeros regrouped by (gate, 64 x 64 sub-block,
        // M half of 32 rows), ALL parameters together, with the 8 x 8 tiles (4 x 8 per half) each group touches
        {
            const int d = a->dim, sbd = d / 64, nsb = sbd * sbd;
            struct Rec3 { uint32_t key; uint32_t p; uint16_t ij; double v; };
            std::vector<Rec3> recs;
            for (int p = 0; p < n_params; ++p)
                for (int t = cptr[p]; t < cptr[p + 1]; ++t) {
                    if (crow[t] >= a->off_rho) break;
                    const int g = (int)(crow[t] / dd); const int wl = (int)(crow[t] - (int64_t)g * dd);
                    const int i = wl / d, j = wl - i * d;
                    const int sb = (i / 64) * sbd + (j / 64), mh = (i & 63) / 32;
                    Rec3 r; r.key = (uint32_t)((g * nsb + sb) * 2 + mh); r.p = (uint32_t)p;
                    r.ij = (uint16_t)(((i & 31) << 8) | (j & 63)); r.v = cval[t];
                    recs.push_back(r);
                }
            std::stable_sort(recs.begin(), recs.end(), [](const Rec3& x, const Rec3& y) { return x.key != y.key ? x.key < y.key : x.p < y.p; });
            const size_t n_keys = (size_t)a->n_ops * nsb * 2;
            std::vector<uint2> tp3(std::max<size_t>(n_keys, 1), make_uint2(0u, 0u));
            std::vector<uint32_t> mask3(std::max<size_t>(n_keys, 1), 0);
            std::vector<uint16_t> ij3; std::vector<int32_t> p3; std::vector<double> v3;
            size_t r = 0;
            for (size_t key = 0; key < n_keys; ++key) {
                // parameters of this (gate, sub-block, half), each with its run of non-zeros [lo, hi) in recs
                std::vector<std::pair<size_t, size_t>> runs;
                while (r < recs.size() && recs[r].key == key) {
                    const size_t lo = r; const uint32_t pcur = recs[r].p;
                    while (r < recs.size() && recs[r].key == key && recs[r].p == pcur) {
                        mask3[key] |= 1u << (8 * ((recs[r].ij >> 8) / 8) + (recs[r].ij & 0xff) / 8);
                        ++r;
                    }
                    runs.emplace_back(lo, r);
                }
                if (runs.empty()) continue;
                // whole parameters to lanes: longest run first onto the least loaded lane (ties: lowest lane) -- deterministic
                std::vector<size_t> order(runs.size());
                std::iota(order.begin(), order.end(), 0);
                std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return runs[x].second - runs[x].first > runs[y].second - runs[y].first; });
                std::vector<std::vector<size_t>> lane_runs(32);
                std::vector<size_t> load(32, 0);
                for (size_t o : order) {
                    const int l = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                    lane_runs[l].push_back(o); load[l] += runs[o].second - runs[o].first;
                }
                const size_t n_it = *std::max_element(load.begin(), load.end());
                const size_t base = ij3.size();
                ij3.resize(base + n_it * 32, 0); p3.resize(base + n_it * 32, -1); v3.resize(base + n_it * 32, 0.0);
                for (int l = 0; l < 32; ++l) {
                    size_t k = 0;
                    for (size_t o : lane_runs[l])
                        for (size_t q = runs[o].first; q < runs[o].second; ++q, ++k) {
                            ij3[base + k * 32 + l] = recs[q].ij; p3[base + k * 32 + l] = (int32_t)recs[q].p; v3[base + k * 32 + l] = recs[q].v;
                        }
                }
                tp3[key] = make_uint2((unsigned)base, (unsigned)n_it);
            }
            if (ij3.empty()) { ij3.push_back(0); p3.push_back(-1); v3.push_back(0.0); }
            if (ij3.size() >= ((size_t)1 << 32)) return fail(B200_E_UNSUPPORTED, "derivative map too large for the level-batched Jacobian plan");
            if ((rc = upload_vec(a->lj3_tp, tp3, ctx->stream)) || (rc = upload_vec(a->lj3_mask, mask3, ctx->stream)) ||
                (rc = upload_vec(a->lj3_items, p3, ctx->stream)) || (rc = upload_vec(a->lj3_ij, ij3, ctx->stream)) ||
                (rc = upload_vec(a->lj3_v, v3, ctx->stream))) return rc;
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    a->has_derivs = true;
    a->has_fderivs = false;      // (a factor-space map of the same parameter block is set AFTER the dense one)
    return B200_OK;
}

// C (+)= A . B  (kernels_gemm.cuh); `batch` entries of A / C share B
static int ab_device(b200_ctx* c, const double* A, int64_t lda, int64_t strideA, const double* B, int64_t ldb,
                     double* C, int64_t ldc, int64_t strideC, int64_t M, int N, int K,
                     const int32_t* kt_ptr, const int32_t* kt_idx, const double* row_scale, bool accumulate, int batch) {
    if (M <= 0 || N <= 0 || batch <= 0) return B200_OK;
    if ((lda & 1) || (ldb & 1) || (strideA & 1) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
        return fail(B200_E_INVALID, "A . B: operands must be 16-byte aligned with even row strides");
    AbArgs p;
    p.A = A; p.lda = lda; p.strideA = strideA; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.strideC = strideC;
    p.M = M; p.N = N; p.K = K; p.kt_ptr = kt_ptr; p.kt_idx = kt_idx; p.row_scale = row_scale; p.accumulate = accumulate ? 1 : 0;
    const size_t smem = (size_t)GM_ST * GM_STAGE_DOUBLES * 8;
    CU(cudaFuncSetAttribute(k_ab_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t gm = (M + GM_TM - 1) / GM_TM;
    if (gm >= ((int64_t)1 << 31) || batch > 65535) return fail(B200_E_UNSUPPORTED, "A . B: grid too large");
    dim3 grid((unsigned)gm, (unsigned)((N + GM_TN - 1) / GM_TN), (unsigned)batch);
    k_ab_dmma<<<grid, 256, smem, c->stream>>>(p);
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

// identity maps so the d16 kernel can emit W (general path)
static int ensure_identity_maps(b200_ctx* ctx, b200_atom* a) {
    if (a->id_colmap.p) return B200_OK;
    std::vector<int32_t> cm((size_t)a->n_w), sc, sw;
    for (int64_t w = 0; w < a->n_w; ++w) cm[w] = (int32_t)w;
    for (int64_t w = a->off_rho; w < a->n_w; ++w) { sc.push_back((int32_t)w); sw.push_back((int32_t)w); }
    a->id_n_spam = (int)sc.size();
    int rc;
    if ((rc = upload_vec(a->id_colmap, cm, ctx->stream)) || (rc = upload_vec(a->id_spam_col, sc, ctx->stream)) ||
        (rc = upload_vec(a->id_spam_w, sw, ctx->stream))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static int grid_for(b200_ctx* c, int64_t n_units, int per_sm) {
    int64_t g = (int64_t)c->sm_count * per_sm;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, n_units));
}

template <int D>
static int launch_probs_generic(b200_ctx* c, b200_atom* a, const double* M, const double* Gt, int nb,
                                double* out, int64_t el_stride, int64_t batch_stride) {
    if (a->n_rows == 0) return B200_OK;
    size_t smem = (size_t)GEN_WARPS * 2 * D * sizeof(double);
    int per_sm = nb > 1 ? 2 : 8;
    int gx = grid_for(c, (a->n_rows + GEN_WARPS - 1) / GEN_WARPS, per_sm);
    dim3 grid(gx, nb);
    k_probs_generic<D><<<grid, GEN_WARPS * 32, smem, c->stream>>>(atom_dev(a), M, Gt, a->n_w, out, el_stride, batch_stride);
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}
static int launch_probs(b200_ctx* c, b200_atom* a, const double* M, const double* Gt, int nb,
                        double* out, int64_t el_stride, int64_t batch_stride) {
    switch (a->dim) {
        case 4: return launch_probs_generic<4>(c, a, M, Gt, nb, out, el_stride, batch_stride);
        case 16: return launch_probs_generic<16>(c, a, M, Gt, nb, out, el_stride, batch_stride);
        case 64: return launch_probs_generic<64>(c, a, M, Gt, nb, out, el_stride, batch_stride);
        case 256: return launch_probs_generic<256>(c, a, M, Gt, nb, out, el_stride, batch_stride);
    }
    return fail(B200_E_UNSUPPORTED, "dim %d", a->dim);
}

template <int D>
static int launch_w_generic(b200_ctx* c, b200_atom* a, double* W, int64_t ldw, double* probs) {
    if (a->n_rows == 0) return B200_OK;
    int gx = grid_for(c, (a->n_rows + GEN_WARPS - 1) / GEN_WARPS, 8);
    size_t need = (size_t)gx * GEN_WARPS * (size_t)(a->max_depth + 1) * D * sizeof(double);
    CU(c->scratch.ensure(need));
    size_t smem = (size_t)GEN_WARPS * 2 * D * sizeof(double);
    k_w_generic<D><<<gx, GEN_WARPS * 32, smem, c->stream>>>(atom_dev(a), model_dev(a), W, ldw, probs, c->scratch.as<double>());
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

// d = 16 Jacobian: the trie path (prefix + suffix sharing).  B200_D16_MODE=generic (test knob) forces the generic kernels.
static bool d16_generic_forced() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200_D16_MODE"); v = (e && !strcmp(e, "generic")) ? 1 : 0; }
    return v == 1;
}
static bool d16_ok(b200_ctx* c, b200_atom* a) {
    return a->dim == 16 && a->has_trie && a->n_ops >= 1 && a->n_ops <= 255 && !d16_generic_forced() &&
           (size_t)a->n_ops * 4096 + 4096 <= c->smem_optin;
}
static int launch_d16(b200_ctx* c, b200_atom* a, const D16Args& args) {
    if (a->n_rows == 0) return B200_OK;
    TrieDev t;
    t.f_meta = a->tf_meta.as<int4>();
    t.f_op = a->tf_op.as<uint8_t>(); t.n_fchains = a->n_fchains; t.n_fnodes = a->n_fnodes;
    t.b_meta = a->tb_meta.as<int4>();
    t.b_op = a->tb_op.as<uint8_t>(); t.n_bchains = a->n_bchains; t.n_bnodes = a->n_bnodes;
    t.fn_b = a->t_fn.as<uint32_t>(); t.bn_b = a->t_bn.as<uint32_t>(); t.f_end = a->t_fend.as<uint32_t>(); t.b_end = a->t_bend.as<uint32_t>();
    t.bcnt = a->bcnt.as<uint16_t>();
    t.S = a->tS(); t.H = a->tH();
    t.counters = a->t_counters.as<unsigned>();
    { int rcP = phase_mark(c); if (rcP) return rcP; }
    k_trie_prepare<<<c->sm_count * 8, 256, 0, c->stream>>>(a->tS(), a->tf_par.as<uint32_t>(), a->n_fpar, a->tH(), a->tb_par.as<uint32_t>(), a->n_bpar,
                                                         (uint32_t)a->n_eff * 16u, a->t_counters.as<unsigned>());
    { int rcP = phase_mark(c); if (rcP) return rcP; }
    const size_t smemA = (size_t)a->n_ops * 256 * 8 + (size_t)TRIE_WARPS * 32 * 8;
    const size_t smemB = (size_t)AT_WARPS * 5 * 16 * 8 + (size_t)a->n_ops * 4 * 32 * 8 + (size_t)2 * D16_SPAM_MAX * 4 + 16;
    CU(cudaFuncSetAttribute(k_trie_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    // 3 chain CTAs per SM and role (even blocks = forward trie, odd = backward trie), 100 ns poll interval of a waiting chain
    k_trie_chains<<<2 * c->sm_count * 3, TRIE_WARPS * 32, smemA, c->stream>>>(atom_dev(a), model_dev(a), t, 0, 100u);
    { int rcP = phase_mark(c); if (rcP) return rcP; }
    const int gB = grid_for(c, ((int64_t)a->n_units + AT_WARPS * AT_CHUNK - 1) / (AT_WARPS * AT_CHUNK), 2);
    if (args.n_peers > 0) {
        const size_t smemP = smemB + 128 + (size_t)AT_WARPS * AT_NO * 256 * 8 + (size_t)a->n_ops * 4;      // + bulk-store staging + per-gate block bases
        CU(cudaFuncSetAttribute(k_accum_trie_d16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
        k_accum_trie_d16<true><<<gB, AT_WARPS * 32, smemP, c->stream>>>(atom_dev(a), model_dev(a), t, args, a->t_units.as<UnitRec>(), a->n_units,
                                                                       a->t_uidx.as<uint2>(), a->t_counters.as<unsigned>() + 2, AT_CHUNK);
    } else {
        CU(cudaFuncSetAttribute(k_accum_trie_d16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
        k_accum_trie_d16<false><<<gB, AT_WARPS * 32, smemB, c->stream>>>(atom_dev(a), model_dev(a), t, args, a->t_units.as<UnitRec>(), a->n_units,
                                                                        a->t_uidx.as<uint2>(), a->t_counters.as<unsigned>() + 2, AT_CHUNK);
    }
    { int rcP = phase_mark(c); if (rcP) return rcP; }
    c->launches += 3;
    CU(cudaGetLastError());
    return B200_OK;
}

// W[el][w] for the whole atom (general path)
static int compute_w(b200_ctx* c, b200_atom* a, double* W, double* probs) {
    if (d16_ok(c, a)) {
        int rc = ensure_identity_maps(c, a);
        if (rc) return rc;
        D16Args args;
        args.colmap = a->id_colmap.as<int32_t>(); args.spam_col = a->id_spam_col.as<int32_t>();
        args.spam_w = a->id_spam_w.as<int32_t>(); args.n_spam = a->id_n_spam;
        args.J = W; args.ld = a->n_w; args.probs = probs; args.row_scale = nullptr; args.n_peers = 0;
        return launch_d16(c, a, args);
    }
    CU(cudaMemsetAsync(W, 0, (size_t)a->n_elements * a->n_w * sizeof(double), c->stream));
    switch (a->dim) {
        case 4: return launch_w_generic<4>(c, a, W, a->n_w, probs);
        case 16: return launch_w_generic<16>(c, a, W, a->n_w, probs);
        case 64: return launch_w_generic<64>(c, a, W, a->n_w, probs);
        case 256: return launch_w_generic<256>(c, a, W, a->n_w, probs);
    }
    return fail(B200_E_UNSUPPORTED, "dim %d", a->dim);
}

template <int D>
static int launch_probs_level(b200_ctx* c, b200_atom* a, double* d_out) {
    const size_t nS = (size_t)a->n_rows * D;
    CU(c->lvl_states.ensure(2 * nS * sizeof(double)));
    double* S0 = c->lvl_states.as<double>(); double* S1 = S0 + nS;
    const double* M = a->M.as<double>();
    int g0 = (int)std::min<int64_t>((int64_t)(nS + 127) / 128, (int64_t)c->sm_count * 8);
    k_level_init<D><<<g0, 128, 0, c->stream>>>(atom_dev(a), M + a->off_rho, S0);
    const size_t smem = (size_t)32 * (D + 4) * sizeof(double);
    CU(cudaFuncSetAttribute(k_level_gemm<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int k = 0; k < a->max_depth; ++k) {
        const uint32_t t0 = a->lvl_tile_ptr[k], t1 = a->lvl_tile_ptr[k + 1];
        if (t1 == t0) continue;
        k_level_gemm<D><<<dim3(t1 - t0, D / 64), 128, smem, c->stream>>>(M, a->lvl_tiles.as<LevelTile>() + t0, a->lvl_circ.as<uint32_t>(),
                                                           (k & 1) ? S1 : S0, (k & 1) ? S0 : S1);
        c->launches++;
    }
    int gp = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + 3) / 4, (int64_t)c->sm_count * 8));
    k_level_probs<D><<<gp, 128, 0, c->stream>>>(atom_dev(a), M + a->off_eff, S0, S1, d_out, 1);
    c->launches += 2;
    CU(cudaGetLastError());
    return B200_OK;
}

static int launch_probs_trie(b200_ctx* c, b200_atom* a, double* d_out) {
    TrieDev t; memset(&t, 0, sizeof t);
    t.f_meta = a->tf_meta.as<int4>();
    t.f_op = a->tf_op.as<uint8_t>(); t.n_fchains = a->n_fchains; t.n_fnodes = a->n_fnodes;
    t.S = a->tS(); t.counters = a->t_counters.as<unsigned>();
    k_trie_prepare<<<c->sm_count * 4, 256, 0, c->stream>>>(a->tS(), a->tf_par.as<uint32_t>(), a->n_fpar, nullptr, nullptr, 0u, 0u, a->t_counters.as<unsigned>());
    c->launches++;
    const size_t smemA = (size_t)a->n_ops * 256 * 8 + (size_t)TRIE_WARPS * 32 * 8;
    CU(cudaFuncSetAttribute(k_trie_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA));
    k_trie_chains<<<c->sm_count * 4, TRIE_WARPS * 32, smemA, c->stream>>>(atom_dev(a), model_dev(a), t, 1, 100u);
    int gp = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + 15) / 16, (int64_t)c->sm_count * 8));
    k_probs_trie_d16<<<gp, 256, 0, c->stream>>>(atom_dev(a), model_dev(a), a->t_fend.as<uint32_t>(), a->tS(), d_out, 1);
    c->launches += 2;
    CU(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_fill_probs_dev(b200_ctx* c, b200_atom* a, double* d_out) {
    if (!c || !a || !d_out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    CU(cudaSetDevice(c->device));
    if (a->n_rows > 0 && d16_ok(c, a)) return launch_probs_trie(c, a, d_out);
    if (a->has_factored && a->fac_probs_ok && a->n_rows > 0 && a->dim >= 64 && !getenv("B200_NO_FACTORED")) {
        // gates as factor programs (Embedded / Composed reps): no dense d x d products at all
        const FactoredDev fd = factored_dev(a);
        const double* M = a->M.as<double>();
        {   // the factor chain on the FP64 tensor cores (k_probs_fdmma) when its fragment image and index table fit in shared memory
            static const bool scalar = getenv("B200_FAC_SCALAR") != nullptr;
            const size_t smem = (size_t)((a->fac_n_frag + 1) & ~1) * 8 + (size_t)a->fac_n * 1024 + (size_t)(((a->n_ops + 1 + a->fac_n + 3) >> 2) * 2) * 8 + (size_t)8 * a->dim * 8;
            if (!scalar && smem <= std::min<size_t>(c->smem_optin, (size_t)110 * 1024)) {
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + 7) / 8, (int64_t)c->sm_count * 2));
                if (a->dim == 64) {
                    CU(cudaFuncSetAttribute(k_probs_fdmma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_probs_fdmma<64><<<grid, 256, smem, c->stream>>>(atom_dev(a), fd, a->fac_ffo.as<int32_t>(), a->fac_n_frag, a->fac_n, M + a->off_rho, M + a->off_eff, d_out, 1);
                } else {
                    CU(cudaFuncSetAttribute(k_probs_fdmma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    k_probs_fdmma<256><<<grid, 256, smem, c->stream>>>(atom_dev(a), fd, a->fac_ffo.as<int32_t>(), a->fac_n_frag, a->fac_n, M + a->off_rho, M + a->off_eff, d_out, 1);
                }
                c->launches++;
                CU(cudaGetLastError());
                return B200_OK;
            }
        }
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + FAC_WARPS - 1) / FAC_WARPS, (int64_t)c->sm_count * 4));
        const int n_ms = a->fac_n_mats, n_fac = a->fac_n;
        const size_t tail = ((size_t)((n_ms + 1) & ~1)) * 8 + (size_t)n_fac * sizeof(FactorRec) + ((size_t)a->n_ops + 1) * 4;
        if (a->dim == 64) {
            const size_t smem = (size_t)FAC_WARPS * 2 * 64 * 8 + tail;
            CU(cudaFuncSetAttribute(k_probs_factored<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_probs_factored<64><<<grid, FAC_WARPS * 32, smem, c->stream>>>(atom_dev(a), fd, n_ms, n_fac, M + a->off_rho, M + a->off_eff, d_out, 1);
        } else {
            const size_t smem = (size_t)FAC_WARPS * 2 * 256 * 8 + tail;
            CU(cudaFuncSetAttribute(k_probs_factored<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_probs_factored<256><<<grid, FAC_WARPS * 32, smem, c->stream>>>(atom_dev(a), fd, n_ms, n_fac, M + a->off_rho, M + a->off_eff, d_out, 1);
        }
        c->launches++;
        CU(cudaGetLastError());
        return B200_OK;
    }
    if (a->has_levels && a->n_rows > 0 && !getenv("B200_NO_LEVELS")) {
        if (a->dim == 64) return launch_probs_level<64>(c, a, d_out);
        if (a->dim == 256) return launch_probs_level<256>(c, a, d_out);
    }
    return launch_probs(c, a, a->M.as<double>(), a->Gt.as<double>(), 1, d_out, 1, 0);
}

// ------------------------------------------------------------------------------------------------
// level-batched Jacobian (d >= 64): forward sweep || backward sweep (two streams), then the sparse contraction
// ------------------------------------------------------------------------------------------------
static int levelj_ts(b200_ctx* c, b200_atom* a, size_t* smem_out) {
    const size_t acc = (size_t)a->lj_no_max * a->lj_pt * 8;
    const size_t per_ts = (size_t)(1 + a->lj_no_max) * a->dim * 8;
    int ts = 8;
    while (ts > 1 && acc + ts * per_ts + (size_t)a->lj_no_max * 8 > std::max<size_t>((size_t)64 * 1024, acc + per_ts + (size_t)a->lj_no_max * 8)) --ts;
    const size_t smem = acc + ts * per_ts + (size_t)a->lj_no_max * 8 + 16;
    if (smem_out) *smem_out = smem;
    return (smem + 1024 <= c->smem_optin) ? ts : 0;
}
static bool levelj_ok(b200_ctx* c, b200_atom* a) {
    return a->has_lj && a->has_levels && (a->dim == 64 || a->dim == 256) && !getenv("B200_NO_LEVELJ") &&
           (int64_t)a->n_rows * std::max(a->lj_n_tiles, 1) < ((int64_t)1 << 31) && levelj_ts(c, a, nullptr) > 0;
}
template <int D>
static int launch_levelj(b200_ctx* c, b200_atom* a, double* d_out, int64_t ld, double* d_probs, const double* d_scale,
                         const PeerOut& peers, bool* peers_done) {
    if (a->n_rows == 0) return B200_OK;
    CU(c->lj_fs.ensure(((size_t)a->lj_rows_f + 1) * D * 8));      // + one all-zero row each (bucket padding of the DMMA accumulate)
    CU(c->lj_bh.ensure(((size_t)a->lj_rows_b + 1) * D * 8));
    CU(cudaMemsetAsync(c->lj_fs.as<double>() + (size_t)a->lj_rows_f * D, 0, (size_t)D * 8, c->stream));
    CU(cudaMemsetAsync(c->lj_bh.as<double>() + (size_t)a->lj_rows_b * D, 0, (size_t)D * 8, c->stream));
    if (!c->aux) {
        CU(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    LevelJDev lj;
    lj.fbase = a->lj_fbase.as<uint32_t>(); lj.bbase = a->lj_bbase.as<uint32_t>();
    lj.bperm = a->bperm.as<uint16_t>(); lj.bcnt = a->bcnt.as<uint16_t>();
    lj.ti_ptr = a->lj_ti_ptr.as<uint32_t>(); lj.items = a->lj_items.as<uint4>();
    lj.crow = a->crow.as<int32_t>(); lj.cval = a->cval.as<double>();
    lj.FS = c->lj_fs.as<double>(); lj.BH = c->lj_bh.as<double>();
    lj.n_tiles = a->lj_n_tiles; lj.n_params = a->n_params; lj.no_max = a->lj_no_max; lj.pt = a->lj_pt;
    size_t smemC = 0;
    lj.ts = levelj_ts(c, a, &smemC);
    const AtomDev ad = atom_dev(a); const ModelDev md = model_dev(a);
    int gi = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + 3) / 4, (int64_t)c->sm_count * 8));
    k_levelj_init<D><<<gi, 128, 0, c->stream>>>(ad, md, lj);
    c->launches++;
    const size_t smem = (size_t)32 * (D + 4) * sizeof(double);
    CU(cudaFuncSetAttribute(k_level_gemm_rows<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool serial = getenv("B200_LJ_SERIAL") != nullptr;
    cudaStream_t sb = serial ? c->stream : c->aux;
    if (!serial) { CU(cudaEventRecord(c->ev_fork, c->stream)); CU(cudaStreamWaitEvent(sb, c->ev_fork, 0)); }
    for (int k = 0; k < a->max_depth; ++k) {          // forward sweep: FS[row + 1] = G FS[row]
        const uint32_t t0 = a->lvl_tile_ptr[k], t1 = a->lvl_tile_ptr[k + 1];
        if (t1 > t0) {
            k_level_gemm_rows<D><<<dim3(t1 - t0, D / 64), 128, smem, c->stream>>>(md.M, a->lvl_tiles.as<LevelTile>() + t0, a->lj_frow.as<uint32_t>(), lj.FS, +1);
            c->launches++;
        }
        const uint32_t b0 = a->lj_btile_ptr[k], b1 = a->lj_btile_ptr[k + 1];
        if (b1 > b0) {                                 // backward sweep (other stream): BH[row - 1] = G^T BH[row]
            k_level_gemm_rows<D><<<dim3(b1 - b0, D / 64), 128, smem, sb>>>(md.Gt, a->lj_btiles.as<LevelTile>() + b0, a->lj_brow.as<uint32_t>(), lj.BH, -1);
            c->launches++;
        }
    }
    if (!serial) { CU(cudaEventRecord(c->ev_join, sb)); CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0)); }
    if (d_probs) {
        k_levelj_probs<D><<<gi, 128, 0, c->stream>>>(ad, md, lj, d_probs);
        c->launches++;
    }
    static const bool accum_v1 = getenv("B200_LJ_ACCUM_V1") != nullptr;      // test knobs: scalar shared-memory contraction (v1),
    static const bool accum_v2 = getenv("B200_LJ_ACCUM_V2") != nullptr;      // CTA-per-(circuit, parameter tile) DMMA accumulate (v2)
    const size_t smemC2 = ((size_t)a->lj_no_max * a->lj_pt + 64 * LJ_LDW) * 8 + LJ_KMAX * 4 + 16;
    const int np_pad = (a->n_params + 7) & ~7;
    const size_t smemC3 = ((size_t)LJ3_WARPS * np_pad + (size_t)LJ3_WARPS * 32 * LJ3_LDW) * 8;
    const int n_og = (a->lj_no_max + LJ3_WARPS - 1) / LJ3_WARPS;
    if (!accum_v1 && !accum_v2 && smemC3 + 1024 <= c->smem_optin && (int64_t)a->n_rows * n_og < ((int64_t)1 << 31)) {
        LevelJ3Dev l3;
        l3.tp3 = a->lj3_tp.as<uint2>(); l3.mask3 = a->lj3_mask.as<uint32_t>(); l3.nz_p = a->lj3_items.as<int32_t>();
        l3.nz_ij = a->lj3_ij.as<uint16_t>(); l3.nz_v = a->lj3_v.as<double>();
        l3.zrow_f = (uint32_t)a->lj_rows_f; l3.zrow_b = (uint32_t)a->lj_rows_b; l3.nsb = (D / 64) * (D / 64); l3.np_pad = np_pad; l3.n_og = n_og;
        CU(cudaFuncSetAttribute(k_level_accum3<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC3));
        k_level_accum3<D><<<(unsigned)(a->n_rows * n_og), LJ3_WARPS * 32, smemC3, c->stream>>>(ad, md, lj, l3, d_out, ld, d_scale, peers);
        if (peers_done) *peers_done = true;
    } else
    if (accum_v1 || smemC2 + 1024 > c->smem_optin) {
        CU(cudaFuncSetAttribute(k_level_accum<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC));
        k_level_accum<D><<<(unsigned)(a->n_rows * a->lj_n_tiles), LJ_THREADS, smemC, c->stream>>>(ad, md, lj, d_out, ld, d_scale);
    } else {
        LevelJ2Dev l2;
        l2.ti_ptr2 = a->lj2_ti_ptr.as<uint32_t>(); l2.mask2 = a->lj2_mask.as<uint64_t>(); l2.items2 = a->lj2_items.as<uint4>();
        l2.nz_ij = a->lj2_ij.as<uint16_t>(); l2.nz_v = a->lj2_v.as<double>();
        l2.zrow_f = (uint32_t)a->lj_rows_f; l2.zrow_b = (uint32_t)a->lj_rows_b; l2.nsb = (D / 64) * (D / 64);
        CU(cudaFuncSetAttribute(k_level_accum2<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemC2));
        k_level_accum2<D><<<(unsigned)(a->n_rows * a->lj_n_tiles), LJ_THREADS, smemC2, c->stream>>>(ad, md, lj, l2, d_out, ld, d_scale);
    }
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

// destinations of the fused exchange (b200_fill_dprobs_bcast_dev): the same slot inside the arrays of the peer GPUs
struct PeerSpec { int n = 0; double* J[B200_PEERS_MAX]; double* P[B200_PEERS_MAX]; bool j_done = false, p_done = false; };

// Jacobian into a device buffer; d_scale (device, [n_elements]) or nullptr.  `ps` (optional): kernels with a fused peer epilogue
// also store into the peers' arrays and mark ps->j_done / ps->p_done.
// ------------------------------------------------------------------------------------------------
// factored Jacobian (gates as factor programs + factor-space derivative map): kernels_factoredj.cuh
// ------------------------------------------------------------------------------------------------
static size_t fj_smem(b200_atom* a, int warps) {
    if (a->dim == 64)        // k_fj64_backward: fragment image | index table | fptr | per warp: 3 vectors + accumulators
        return (size_t)((a->fj_n_frag + 1) & ~1) * 8 + (size_t)a->fac_n * 32 * 16 + ((size_t)a->n_ops + 1) * 4 + 16 + (size_t)warps * (3 * 64 + (size_t)a->fj_n_acc) * 8;
    return (size_t)((a->fj_n_frag + 1) & ~1) * 8 + (size_t)a->fac_n * sizeof(FactorRec) + ((size_t)a->n_ops + 1 + 2 * (size_t)a->fac_n) * 4 + 16 +
           (size_t)warps * (3 * (size_t)a->dim + a->fj_n_acc) * 8;
}
static int fj_warps(b200_ctx* c, b200_atom* a) {
    for (int w = 8; w >= 1; w >>= 1) if (fj_smem(a, w) <= std::min<size_t>(c->smem_optin, (size_t)110 * 1024) || (w == 1 && fj_smem(a, 1) <= c->smem_optin)) return w;
    return 0;
}
static bool factoredj_ok(b200_ctx* c, b200_atom* a) {
    static const bool off = getenv("B200_NO_FACTOREDJ") != nullptr;
    return !off && a->has_factored && a->has_fderivs && a->fac_probs_ok && (a->dim == 64 || a->dim == 256) && a->n_rows > 0 &&
           a->fj_n_params == a->n_params && same_factor_structure(a->fj_fac, a->h_fac) && fj_warps(c, a) > 0;
}
template <int D>
static int launch_factoredj(b200_ctx* c, b200_atom* a, double* d_out, int64_t ld, double* d_probs, const double* d_scale, PeerSpec* ps) {
    CU(c->fj_fs.ensure((size_t)a->fj_rows * D * 8));
    CU(c->fj_counter.ensure(16));
    const FactoredDev fd = factored_dev(a);
    const double* M = a->M.as<double>();
    FjDev fj;
    fj.base = a->fj_base.as<uint32_t>(); fj.out_circ = a->fj_out_circ.as<int32_t>(); fj.step_fac = a->fj_steps.as<uint16_t>(); fj.fao = a->fj_fao.as<int32_t>(); fj.ffo = a->fj_ffo.as<int32_t>();
    fj.n_acc = a->fj_n_acc; fj.n_frag = a->fj_n_frag; fj.cptr = a->fj_cptr.as<int32_t>(); fj.ccode = a->fj_ccode.as<uint32_t>(); fj.cval = a->fj_cval.as<double>();
    fj.n_params = a->fj_n_params; fj.unit = a->fj_unit;
    if (D == 64) {
        const size_t smem = (size_t)((a->fj_n_frag + 1) & ~1) * 8 + (size_t)a->fac_n * 32 * 16 + ((size_t)a->n_ops + 1) * 4 + 16 + (size_t)8 * 2 * 64 * 8;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + 7) / 8, (int64_t)c->sm_count * 6));
        CU(cudaFuncSetAttribute(k_fj64_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fj64_forward<<<grid, 256, smem, c->stream>>>(atom_dev(a), fd, fj, a->fac_n, M + a->off_rho, M + a->off_eff, c->fj_fs.as<double>(), d_probs);
    } else {
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((a->n_rows + FAC_WARPS - 1) / FAC_WARPS, (int64_t)c->sm_count * 4));
        const size_t smem = (size_t)FAC_WARPS * 2 * D * 8 + ((size_t)((a->fac_n_mats + 1) & ~1)) * 8 + (size_t)a->fac_n * sizeof(FactorRec) + ((size_t)a->n_ops + 1) * 4;
        CU(cudaFuncSetAttribute(k_fj_forward<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fj_forward<D><<<grid, FAC_WARPS * 32, smem, c->stream>>>(atom_dev(a), fd, a->fac_n_mats, a->fac_n, a->fj_base.as<uint32_t>(), M + a->off_rho, M + a->off_eff,
                                                                c->fj_fs.as<double>(), d_probs);
    }
    CU(cudaMemsetAsync(c->fj_counter.p, 0, 4, c->stream));
    const int warps = fj_warps(c, a);
    const size_t smem = fj_smem(a, warps);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, ((size_t)227 * 1024) / (smem + 1024)));
    const int64_t n_items = a->n_elements;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_items + warps - 1) / warps, (int64_t)c->sm_count * per_sm));
    if (D == 64) {
        Fj64Slots sl; for (int q = 0; q < FJ64_REG_SLOTS; ++q) sl.fao[q] = a->fj_slot_fao[q];
        FjPeers fp; fp.n = 0;
        if (ps && ps->n > 0) { fp.n = ps->n; for (int r = 0; r < ps->n; ++r) fp.J[r] = ps->J[r]; ps->j_done = true; }      // rows also stored into the peers' arrays
        if (fp.n > 0) {
            CU(cudaFuncSetAttribute(k_fj64_backward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fj64_backward<true><<<grid, warps * 32, smem, c->stream>>>(atom_dev(a), fd, fj, a->fac_n, a->fj_slots.as<int32_t>(), sl, M + a->off_eff, c->fj_fs.as<double>(),
                                                                       d_out, ld, d_scale, c->fj_counter.as<unsigned>(), (int)n_items, fp);
        } else {
            CU(cudaFuncSetAttribute(k_fj64_backward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fj64_backward<false><<<grid, warps * 32, smem, c->stream>>>(atom_dev(a), fd, fj, a->fac_n, a->fj_slots.as<int32_t>(), sl, M + a->off_eff, c->fj_fs.as<double>(),
                                                                        d_out, ld, d_scale, c->fj_counter.as<unsigned>(), (int)n_items, fp);
        }
    } else {
        CU(cudaFuncSetAttribute(k_fj_backward<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fj_backward<D><<<grid, warps * 32, smem, c->stream>>>(atom_dev(a), fd, fj, a->fac_n, M + a->off_eff, c->fj_fs.as<double>(), d_out, ld, d_scale,
                                                             c->fj_counter.as<unsigned>(), (int)n_items);
    }
    c->launches += 2;
    CU(cudaGetLastError());
    return B200_OK;
}

static int fill_dprobs_device(b200_ctx* c, b200_atom* a, double* d_out, int64_t ld, double* d_probs, const double* d_scale,
                              PeerSpec* ps = nullptr) {
    if (!c || !a || !d_out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (!a->has_derivs && !a->has_fderivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (ld < a->n_params) return fail(B200_E_INVALID, "ld=%lld < n_params=%d", (long long)ld, a->n_params);
    CU(cudaSetDevice(c->device));
    if (a->n_elements == 0 || a->n_params == 0) {
        if (d_probs && a->n_elements) return b200_fill_probs_dev(c, a, d_probs);
        return B200_OK;
    }
    if (a->unit_perm && d16_ok(c, a)) {
        D16Args args;
        args.colmap = a->colmap.as<int32_t>(); args.spam_col = a->spam_col.as<int32_t>();
        args.spam_w = a->spam_w.as<int32_t>(); args.n_spam = a->n_spam;
        args.J = d_out; args.ld = ld; args.probs = d_probs;
        args.row_scale = d_scale;                  // applied in the store epilogue
        args.n_peers = 0;
        if (ps && ps->n > 0) {
            args.n_peers = ps->n;
            for (int r = 0; r < ps->n; ++r) { args.peerJ[r] = ps->J[r]; args.peerP[r] = d_probs ? ps->P[r] : nullptr; }
            ps->j_done = true; ps->p_done = true;
        }
        return launch_d16(c, a, args);
    }
    if (factoredj_ok(c, a)) {
        return a->dim == 64 ? launch_factoredj<64>(c, a, d_out, ld, d_probs, d_scale, ps) : launch_factoredj<256>(c, a, d_out, ld, d_probs, d_scale, ps);
    }
    if (!a->has_derivs) return fail(B200_E_STATE, "only a factor-space derivative map is set and the factored Jacobian path is unavailable for this atom");
    if (levelj_ok(c, a)) {
        PeerOut po; po.n = 0;
        if (ps) { po.n = ps->n; for (int r = 0; r < ps->n; ++r) po.J[r] = ps->J[r]; }
        bool done = false;
        int rc = a->dim == 64 ? launch_levelj<64>(c, a, d_out, ld, d_probs, d_scale, po, &done)
                              : launch_levelj<256>(c, a, d_out, ld, d_probs, d_scale, po, &done);
        if (ps && done) ps->j_done = true;
        return rc;
    }
    // general path: W then J = W . D
    CU(c->w_buf.ensure((size_t)a->n_elements * a->n_w * sizeof(double)));
    int rc = compute_w(c, a, c->w_buf.as<double>(), d_probs);
    if (rc) return rc;
    if (a->has_dense)
        return ab_device(c, c->w_buf.as<double>(), a->n_w, 0, a->Dd.as<double>(), a->ldD, d_out, ld, 0, a->n_elements, a->n_params, (int)a->n_w,
                         a->kt_ptr.as<int32_t>(), a->kt_idx.as<int32_t>(), d_scale, false, 1);
    dim3 grid((a->n_params + 127) / 128, (unsigned)std::min<int64_t>(a->n_elements, 65535));
    k_contract_csc<<<grid, 128, 0, c->stream>>>(c->w_buf.as<double>(), a->n_w, a->n_elements, a->n_params,
                                                 a->cptr.as<int32_t>(), a->crow.as<int32_t>(), a->cval.as<double>(), d_out, ld, d_scale);
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_fill_dprobs_dev(b200_ctx* c, b200_atom* a, double* d_out, int64_t ld, double* d_probs) {
    return fill_dprobs_device(c, a, d_out, ld, d_probs, nullptr);
}

// ---- peer memory: fused fill + exchange (one process per GPU) ---------------------------------------------------------
extern "C" int b200_peer_alloc(b200_ctx* c, int64_t bytes, void** d_ptr_out, unsigned char handle_out[64]) {
    if (!c || !d_ptr_out || !handle_out || bytes <= 0) return fail(B200_E_INVALID, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CU(cudaSetDevice(c->device));
    *d_ptr_out = nullptr;
    void* p = nullptr;
    CU(cudaMalloc(&p, (size_t)bytes));                  // (a plain cudaMalloc allocation: the kind CUDA IPC can export)
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return fail(B200_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    *d_ptr_out = p;
    return B200_OK;
}
extern "C" int b200_peer_open(b200_ctx* c, const unsigned char handle[64], void** d_ptr_out) {
    if (!c || !handle || !d_ptr_out) return fail(B200_E_INVALID, "bad argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
    *d_ptr_out = nullptr;
    CU(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return B200_OK;
}
extern "C" int b200_peer_close(b200_ctx* c, void* d_ptr) {
    if (!c || !d_ptr) return B200_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaIpcCloseMemHandle(d_ptr));
    return B200_OK;
}
extern "C" int b200_peer_free(b200_ctx* c, void* d_ptr) {
    if (!c || !d_ptr) return B200_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(d_ptr));
    return B200_OK;
}

extern "C" int b200_fill_dprobs_bcast_dev(b200_ctx* c, b200_atom* a, double* d_out, int64_t ld, double* d_probs,
                                          int n_peers, double* const* d_out_peers, double* const* d_probs_peers) {
    if (!c || !a || !d_out) return fail(B200_E_INVALID, "NULL argument");
    if (n_peers < 0 || n_peers > B200_MAX_PEERS) return fail(B200_E_INVALID, "n_peers must be in [0, %d]", B200_MAX_PEERS);
    if (n_peers > 0 && !d_out_peers) return fail(B200_E_INVALID, "d_out_peers is NULL");
    PeerSpec ps; ps.n = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!d_out_peers[r]) return fail(B200_E_INVALID, "d_out_peers[%d] is NULL", r);
        ps.J[r] = d_out_peers[r]; ps.P[r] = d_probs_peers ? d_probs_peers[r] : nullptr;
    }
    int rc = fill_dprobs_device(c, a, d_out, ld, d_probs, nullptr, &ps);
    if (rc) return rc;
    // kernels without a fused peer epilogue: forward the finished slot with asynchronous peer copies on the same stream
    for (int r = 0; r < n_peers; ++r) {
        if (!ps.j_done && a->n_elements > 0 && a->n_params > 0)
            CU(cudaMemcpy2DAsync(ps.J[r], (size_t)ld * 8, d_out, (size_t)ld * 8, (size_t)a->n_params * 8, (size_t)a->n_elements, cudaMemcpyDefault, c->stream));
        if (!ps.p_done && d_probs && ps.P[r] && a->n_elements > 0)
            CU(cudaMemcpyAsync(ps.P[r], d_probs, (size_t)a->n_elements * 8, cudaMemcpyDefault, c->stream));
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points
// ------------------------------------------------------------------------------------------------
// device [n_rows x width] (ld = width) -> host with row stride `hstride` doubles (a contiguous destination is one 1-D copy).
// One copy stream: 1..4 concurrent streams all landed at 53-56 GB/s (PCIe-bound) in round 1.
static int copy_out_2d(b200_ctx* c, const double* d_src, int64_t width, int64_t n_rows, double* h_dst, int64_t hstride) {
    if (n_rows == 0 || width == 0) return B200_OK;
    if (hstride == width) { CU(cudaMemcpyAsync(h_dst, d_src, (size_t)n_rows * width * 8, cudaMemcpyDeviceToHost, c->stream)); }
    else CU(cudaMemcpy2DAsync(h_dst, (size_t)hstride * 8, d_src, (size_t)width * 8, (size_t)width * 8, (size_t)n_rows, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

extern "C" int b200_fill_probs(b200_ctx* c, b200_atom* a, double* out, int64_t out_stride) {
    if (!c || !a || !out) return fail(B200_E_INVALID, "NULL argument");
    if (out_stride < 1) return fail(B200_E_INVALID, "out_stride must be >= 1");
    CU(cudaSetDevice(c->device));
    CU(c->probs_buf.ensure(std::max<size_t>((size_t)a->n_elements * 8, 16)));
    int rc = b200_fill_probs_dev(c, a, c->probs_buf.as<double>());
    if (rc) return rc;
    return copy_out_2d(c, c->probs_buf.as<double>(), 1, a->n_elements, out, out_stride);
}

extern "C" int b200_fill_dprobs(b200_ctx* c, b200_atom* a, double* out, int64_t row_stride,
                                double* probs_out, int64_t probs_stride) {
    if (!c || !a || !out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_derivs && !a->has_fderivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (row_stride < a->n_params) return fail(B200_E_INVALID, "row_stride < n_params");
    if (probs_out && probs_stride < 1) return fail(B200_E_INVALID, "probs_stride must be >= 1");
    CU(cudaSetDevice(c->device));
    CU(c->out_buf.ensure(std::max<size_t>((size_t)a->n_elements * a->n_params * 8, 16)));
    double* d_probs = nullptr;
    if (probs_out) { CU(c->probs_buf.ensure(std::max<size_t>((size_t)a->n_elements * 8, 16))); d_probs = c->probs_buf.as<double>(); }
    int rc = b200_fill_dprobs_dev(c, a, c->out_buf.as<double>(), a->n_params, d_probs);
    if (rc) return rc;
    rc = copy_out_2d(c, c->out_buf.as<double>(), a->n_params, a->n_elements, out, row_stride);
    if (rc) return rc;
    if (probs_out) return copy_out_2d(c, d_probs, 1, a->n_elements, probs_out, probs_stride);
    return B200_OK;
}

static int upload_scale(b200_ctx* c, b200_atom* a, const double* row_scale, const double** d_scale) {
    *d_scale = nullptr;
    if (!row_scale || a->n_elements == 0) return B200_OK;
    CU(c->scale_buf.ensure((size_t)a->n_elements * 8));
    CU(cudaMemcpyAsync(c->scale_buf.p, row_scale, (size_t)a->n_elements * 8, cudaMemcpyHostToDevice, c->stream));
    *d_scale = c->scale_buf.as<double>();
    return B200_OK;
}

extern "C" int b200_fill_dprobs_scaled(b200_ctx* c, b200_atom* a, const double* row_scale, double* out, int64_t row_stride,
                                       double* probs_out, int64_t probs_stride) {
    if (!c || !a || !out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_derivs && !a->has_fderivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (row_stride < a->n_params) return fail(B200_E_INVALID, "row_stride < n_params");
    if (probs_out && probs_stride < 1) return fail(B200_E_INVALID, "probs_stride must be >= 1");
    CU(cudaSetDevice(c->device));
    CU(c->out_buf.ensure(std::max<size_t>((size_t)a->n_elements * a->n_params * 8, 16)));
    double* d_probs = nullptr;
    if (probs_out) { CU(c->probs_buf.ensure(std::max<size_t>((size_t)a->n_elements * 8, 16))); d_probs = c->probs_buf.as<double>(); }
    const double* d_scale = nullptr;
    int rc = upload_scale(c, a, row_scale, &d_scale);
    if (rc) return rc;
    rc = fill_dprobs_device(c, a, c->out_buf.as<double>(), a->n_params, d_probs, d_scale);
    if (rc) return rc;
    rc = copy_out_2d(c, c->out_buf.as<double>(), a->n_params, a->n_elements, out, row_stride);
    if (rc) return rc;
    if (probs_out) return copy_out_2d(c, d_probs, 1, a->n_elements, probs_out, probs_stride);
    return B200_OK;
}

// C [na x nb] (+)= A^T B over nE rows on the FP64 tensor cores (kernels_jtj.cuh); tri: A == B, symmetric result.  atf = A^T f.
static int atb_device(b200_ctx* c, const double* A, int64_t lda, int na, const double* B, int64_t ldb, int nb, int64_t nE,
                      bool tri, const double* f, double* C, int64_t ldc, double* atf, bool accumulate) {
    if (na <= 0 || nb <= 0) return B200_OK;
    if ((lda & 1) || (ldb & 1) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15))
        return fail(B200_E_INVALID, "A^T B: operands must be 16-byte aligned with even row strides");
    AtbArgs p;
    p.A = A; p.lda = lda; p.na = na; p.B = B; p.ldb = ldb; p.nb = nb; p.nE = nE; p.tri = tri ? 1 : 0;
    p.n_bi = (na + JT_T - 1) / JT_T; p.n_bj = (nb + JT_T - 1) / JT_T;
    p.n_tiles = tri ? p.n_bi * (p.n_bi + 1) / 2 : p.n_bi * p.n_bj;
    // K slices (partial tiles are summed in slice order afterwards): one CTA per (slice, tile) and one CTA per SM, so the
    // slice count is chosen to fill whole waves of SMs -- 9 slices x 66 tiles = 594 CTAs on 148 SMs ran 5 waves for 4.01 waves
    // of work (measured: 0.57 of the DMMA peak with the pipe 84 % busy while active).  Search 3..10 waves, >= 256 rows a slice.
    int64_t S = 1;
    {
        const int64_t s_max = std::max<int64_t>(1, std::min<int64_t>(256, nE / 256));
        double best = -1.0;
        for (int64_t cand = 1; cand <= s_max; ++cand) {
            const int64_t ctas = cand * p.n_tiles, waves = (ctas + c->sm_count - 1) / c->sm_count;
            if (waves > 10 && best > 0) break;
            double eff = (double)ctas / (double)(waves * c->sm_count);
            if (waves < 3) eff *= 0.9;                       // few long CTAs: the tail of the last wave is not hidden
            if (eff > best + 1e-9) { best = eff; S = cand; }
        }
    }
    p.rows_per_slice = std::max<int64_t>(JT_KC, ((nE + S - 1) / S + JT_KC - 1) / JT_KC * JT_KC);
    p.n_slices = (int)std::max<int64_t>(1, (nE + p.rows_per_slice - 1) / p.rows_per_slice);
    p.f = (atf && f) ? f : nullptr;
    CU(c->atb_part.ensure((size_t)p.n_slices * p.n_tiles * JT_T * JT_T * 8));
    CU(c->atb_part_f.ensure((size_t)p.n_slices * p.n_bi * JT_T * 8));
    p.part = c->atb_part.as<double>(); p.part_f = c->atb_part_f.as<double>();
    const size_t smem = (size_t)JT_ST * JT_STAGE_DOUBLES * 8;
    CU(cudaFuncSetAttribute(k_atb_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_atb_dmma<<<(unsigned)(p.n_slices * p.n_tiles), 256, smem, c->stream>>>(p);
    k_atb_reduce<<<(unsigned)p.n_tiles, 256, 0, c->stream>>>(p, C, ldc, atf, accumulate ? 1 : 0);
    c->launches += 2;
    CU(cudaGetLastError());
    return B200_OK;
}

// J^T J on the 5th-generation tensor cores (tcgen05.mma kind::i8 + TMEM) through the Ozaki splitting -- kernels_ozaki.cuh.
// J [nE x ldj] on the device; result d_jtj [Np x Np] (row-major, full symmetric).  T = 8 digits (62 bits) or 7 (55 bits).
template <int T>
static int jtj_ozaki(b200_ctx* c, const double* J, int64_t ldj, int64_t nE, int Np, const double* d_f, double* d_jtj, double* d_jtf) {
    const int n_bi = (Np + OZ_TM - 1) / OZ_TM, n_bj = (Np + OZ_TN - 1) / OZ_TN;
    const int64_t P_pad = (int64_t)n_bi * OZ_TM;
    std::vector<int2>& tiles = c->oz_tiles;
    if (c->oz_tiles_np != Np) {
        CU(cudaStreamSynchronize(c->stream));       // an earlier asynchronous upload may still read the old list
        tiles.clear();
        for (int bi = 0; bi < n_bi; ++bi)
            for (int bj = 0; bj < n_bj; ++bj) if ((int64_t)bj * OZ_TN <= (int64_t)bi * OZ_TM + OZ_TM - 1) tiles.push_back(make_int2(bi, bj));
        c->oz_tiles_np = Np;
    }
    const int n_tiles = (int)tiles.size();
    const int64_t stages_total = (nE + OZ_KS - 1) / OZ_KS;
    // K slices: short enough that every level's int32 sum stays exact, and as many more as fill whole waves of SMs (one CTA per SM)
    const int64_t ksl_min = (stages_total + OZ_MAX_STAGES_PER_SLICE - 1) / OZ_MAX_STAGES_PER_SLICE;
    int64_t ksl = ksl_min; double best = -1.0;
    for (int64_t cand = ksl_min; cand <= ksl_min + 24 && cand <= std::max<int64_t>(1, stages_total / 64); ++cand) {
        const int64_t ctas = cand * n_tiles, waves = (ctas + c->sm_count - 1) / c->sm_count;
        const double eff = (double)ctas / (double)(waves * c->sm_count);
        if (eff > best + 0.01) { best = eff; ksl = cand; }
    }
    const int64_t sps = (stages_total + ksl - 1) / ksl;
    const int64_t n_stages = ksl * sps;
    if (sps > OZ_MAX_STAGES_PER_SLICE) return fail(B200_E_UNSUPPORTED, "Ozaki J^T J: K slice too long");
    // column statistics: one pass over J for the exponents and J^T f
    const int gy = (int)std::max<int64_t>(1, std::min<int64_t>(128, (nE + 15) / 16));      // block rows of the statistics pass
    const int ns = 4 * gy;                                                                  // partial slots (row lanes)
    CU(c->oz_S.ensure((size_t)n_stages * T * P_pad * OZ_KS));
    const size_t misc_bytes = (size_t)2 * ns * Np * 8 + (size_t)P_pad * 4 + (size_t)n_tiles * sizeof(int2) + 64;
    CU(c->oz_misc.ensure(misc_bytes));
    CU(c->oz_part.ensure((size_t)ksl * n_tiles * OZ_TM * OZ_TN * 8));
    double* d_psum = c->oz_misc.as<double>();
    double* d_pmax = d_psum + (size_t)ns * Np;
    int* d_expo = reinterpret_cast<int*>(d_pmax + (size_t)ns * Np);
    int2* d_tiles = reinterpret_cast<int2*>(d_expo + P_pad);            // P_pad is a multiple of 128: 8-byte aligned
    CU(cudaMemsetAsync(d_expo, 0, (size_t)P_pad * 4, c->stream));
    CU(cudaMemcpyAsync(d_tiles, tiles.data(), (size_t)n_tiles * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    k_oz_colstats<<<dim3((unsigned)((Np + 63) / 64), (unsigned)gy), 256, 0, c->stream>>>(J, ldj, nE, Np, d_jtf ? d_f : nullptr, d_psum, d_pmax);
    k_oz_colstats_reduce<<<(unsigned)((Np + 255) / 256), 256, 0, c->stream>>>(d_psum, d_pmax, Np, ns, d_jtf, d_expo);
    k_oz_slice<T><<<dim3((unsigned)(P_pad / 64), (unsigned)((n_stages + 1) / 2)), 256, 0, c->stream>>>(J, ldj, nE, Np, d_expo, c->oz_S.as<int8_t>(), P_pad, n_stages);
    OzArgs p;
    p.S = c->oz_S.as<int8_t>(); p.P_pad = P_pad; p.tiles = d_tiles; p.n_tiles = n_tiles;
    p.n_kslices = (int)ksl; p.stages_per_slice = (int)sps; p.part = c->oz_part.as<double>();
    const size_t smem = (size_t)OZ_NST * T * (OZ_TM + OZ_TN) * OZ_KS;
    CU(cudaFuncSetAttribute(k_oz_syrk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_oz_syrk<T><<<(unsigned)(ksl * n_tiles), 128, smem, c->stream>>>(p);
    k_oz_reduce<<<(unsigned)n_tiles, 256, 0, c->stream>>>(p, d_expo, Np, d_jtj, Np);
    c->launches += 5;
    CU(cudaGetLastError());
    return B200_OK;
}

// J^T J (full symmetric, row-major) and J^T f into DEVICE buffers; everything asynchronous on the ctx stream
static int jtj_device(b200_ctx* c, b200_atom* a, const double* d_scale, const double* d_f, double* d_jtj, double* d_jtf) {
    const int Np = a->n_params; const int64_t nE = a->n_elements;
    if (Np == 0) return B200_OK;
    const int64_t ldj = (Np + 1) & ~1;                  // even row stride: 16-byte aligned cp.async rows
    CU(c->out_buf.ensure(std::max<size_t>((size_t)nE * ldj * 8, 16)));
    int rc = fill_dprobs_device(c, a, c->out_buf.as<double>(), ldj, nullptr, d_scale);
    if (rc) return rc;
    if (nE == 0) {
        CU(cudaMemsetAsync(d_jtj, 0, (size_t)Np * Np * 8, c->stream));
        if (d_jtf) CU(cudaMemsetAsync(d_jtf, 0, (size_t)Np * 8, c->stream));
        return B200_OK;
    }
    int mode = c->jtj_mode;
    if (mode < 0) mode = (nE >= 4096) ? 8 : 0;
    if (mode == 8) return jtj_ozaki<8>(c, c->out_buf.as<double>(), ldj, nE, Np, d_f, d_jtj, d_jtf);
    if (mode == 7) return jtj_ozaki<7>(c, c->out_buf.as<double>(), ldj, nE, Np, d_f, d_jtj, d_jtf);
    return atb_device(c, c->out_buf.as<double>(), ldj, Np, c->out_buf.as<double>(), ldj, Np, nE, true, d_f, d_jtj, Np, d_jtf, false);
}

extern "C" int b200_jtj(b200_ctx* c, b200_atom* a, const double* row_scale, const double* f, double* jtj_out, double* jtf_out) {
    if (!c || !a || !jtj_out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_derivs && !a->has_fderivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (jtf_out && !f) return fail(B200_E_INVALID, "jtf_out requires f");
    CU(cudaSetDevice(c->device));
    const int Np = a->n_params; const int64_t nE = a->n_elements;
    if (Np == 0) return B200_OK;
    CU(c->jtj_buf.ensure((size_t)Np * Np * 8));
    const double* d_scale = nullptr;
    int rc = upload_scale(c, a, row_scale, &d_scale);
    if (rc) return rc;
    double* d_f = nullptr; double* d_jtf = nullptr;
    if (jtf_out) {
        CU(c->f_buf.ensure(std::max<size_t>((size_t)nE * 8, 16)));
        CU(c->jtf_buf.ensure((size_t)Np * 8));
        if (nE > 0) CU(cudaMemcpyAsync(c->f_buf.p, f, (size_t)nE * 8, cudaMemcpyHostToDevice, c->stream));
        d_f = c->f_buf.as<double>(); d_jtf = c->jtf_buf.as<double>();
    }
    rc = jtj_device(c, a, d_scale, d_f, c->jtj_buf.as<double>(), d_jtf);
    if (rc) return rc;
    if (jtf_out) CU(cudaMemcpyAsync(jtf_out, c->jtf_buf.p, (size_t)Np * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(jtj_out, c->jtj_buf.p, (size_t)Np * Np * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return B200_OK;
}

extern "C" int b200_jtj_dev(b200_ctx* c, b200_atom* a, const double* d_row_scale, const double* d_f, double* d_jtj, double* d_jtf) {
    if (!c || !a || !d_jtj) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (!a->has_derivs && !a->has_fderivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (d_jtf && !d_f) return fail(B200_E_INVALID, "d_jtf requires d_f");
    CU(cudaSetDevice(c->device));
    return jtj_device(c, a, d_row_scale, d_f, d_jtj, d_jtf);
}

extern "C" int b200_fill_dprobs_fd(b200_ctx* c, b200_atom* a, double eps, double* out, int64_t row_stride,
                                   double* probs_out, int64_t probs_stride) {
    if (!c || !a || !out) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (!a->has_derivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    if (row_stride < a->n_params) return fail(B200_E_INVALID, "row_stride < n_params");
    if (eps == 0.0) return fail(B200_E_INVALID, "eps must be non-zero");
    CU(cudaSetDevice(c->device));
    const int64_t nE = a->n_elements; const int Np = a->n_params;
    CU(c->out_buf.ensure(std::max<size_t>((size_t)nE * Np * 8, 16)));
    CU(c->probs_buf.ensure(std::max<size_t>((size_t)nE * 8, 16)));
    double* P0 = c->probs_buf.as<double>();
    int rc = launch_probs(c, a, a->M.as<double>(), a->Gt.as<double>(), 1, P0, 1, 0);
    if (rc) return rc;
    // batches of perturbed models
    const int B = (int)std::max<int64_t>(1, std::min<int64_t>(Np, std::min<int64_t>(256, ((int64_t)256 << 20) / std::max<int64_t>(1, nE * 8))));
    CU(c->fd_models.ensure((size_t)B * a->n_w * 8));
    CU(c->fd_gt.ensure(std::max<size_t>((size_t)B * a->off_rho * 8, 16)));
    CU(c->fd_probs.ensure(std::max<size_t>((size_t)B * nE * 8, 16)));
    for (int p0 = 0; p0 < Np; p0 += B) {
        const int nb = std::min(B, Np - p0);
        dim3 g1((unsigned)std::min<int64_t>((a->n_w + 255) / 256, 256), nb);
        k_perturb_models<<<g1, 256, 0, c->stream>>>(a->M.as<double>(), a->n_w, c->fd_models.as<double>());
        dim3 g2(4, nb);
        k_perturb_apply<<<g2, 128, 0, c->stream>>>(a->n_w, p0, eps, a->cptr.as<int32_t>(), a->crow.as<int32_t>(),
                                                   a->cval.as<double>(), c->fd_models.as<double>());
        c->launches += 2;
        if (a->n_ops) {
            dim3 g3((unsigned)std::min<int64_t>((a->off_rho + 255) / 256, 256), nb);
            k_transpose_gates<<<g3, 256, 0, c->stream>>>(c->fd_models.as<double>(), a->n_w, a->n_ops, a->dim, c->fd_gt.as<double>());
            c->launches++;
        }
        CU(cudaGetLastError());
        rc = launch_probs(c, a, c->fd_models.as<double>(), c->fd_gt.as<double>(), nb, c->fd_probs.as<double>(), 1, nE);
        if (rc) return rc;
        if (nE) {
            int gx = (int)std::min<int64_t>((nE + 3) / 4, 4096);
            k_fd_finish<<<gx, 128, 0, c->stream>>>(c->fd_probs.as<double>(), P0, nE, nb, p0, eps, c->out_buf.as<double>(), Np);
            c->launches++;
            CU(cudaGetLastError());
        }
    }
    rc = copy_out_2d(c, c->out_buf.as<double>(), Np, nE, out, row_stride);
    if (rc) return rc;
    if (probs_out) return copy_out_2d(c, P0, 1, nE, probs_out, probs_stride);
    return B200_OK;
}

template <int D>
static int launch_w_tangent(b200_ctx* c, b200_atom* a, int nb, const double* dMb, const double* dGtb, double* Wb, const unsigned char* need) {
    int gx = grid_for(c, (a->n_rows + GEN_WARPS - 1) / GEN_WARPS, 4);
    const size_t scratch_bytes = (size_t)nb * gx * GEN_WARPS * 2 * (size_t)(a->max_depth + 1) * D * sizeof(double);
    CU(c->scratch.ensure(scratch_bytes));
    size_t smem = (size_t)GEN_WARPS * 4 * D * sizeof(double);
    dim3 grid(gx, nb);
    k_w_tangent_generic<D><<<grid, GEN_WARPS * 32, smem, c->stream>>>(atom_dev(a), model_dev(a), dMb, dGtb, Wb, a->n_w,
                                                                       c->scratch.as<double>(), need);
    c->launches++;
    CU(cudaGetLastError());
    return B200_OK;
}

static int fill_hprobs_impl(b200_ctx* c, b200_atom* a, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                           int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b, const double* h_vals,
                           double* out, const double* w_h = nullptr, const double* w_d = nullptr, double* red_out = nullptr);

extern "C" int b200_fill_hprobs_linear(b200_ctx* c, b200_atom* a, int32_t n1, const int32_t* p1,
                                       int32_t n2, const int32_t* p2, double* out) {
    return fill_hprobs_impl(c, a, n1, p1, n2, p2, 0, nullptr, nullptr, nullptr, nullptr, out);
}

extern "C" int b200_fill_hprobs(b200_ctx* c, b200_atom* a, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                                int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b,
                                const double* h_vals, double* out) {
    if (nnz2 < 0 || (nnz2 > 0 && (!h_rows || !h_a || !h_b || !h_vals))) return fail(B200_E_INVALID, "bad second-derivative map");
    return fill_hprobs_impl(c, a, n1, p1, n2, p2, nnz2, h_rows, h_a, h_b, h_vals, out);
}

extern "C" int b200_hessian_block(b200_ctx* c, b200_atom* a, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                                  int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b,
                                  const double* h_vals, const double* w_h, const double* w_d, double* out) {
    if (!w_h || !w_d || !out) return fail(B200_E_INVALID, "NULL argument");
    if (nnz2 < 0 || (nnz2 > 0 && (!h_rows || !h_a || !h_b || !h_vals))) return fail(B200_E_INVALID, "bad second-derivative map");
    return fill_hprobs_impl(c, a, n1, p1, n2, p2, nnz2, h_rows, h_a, h_b, h_vals, nullptr, w_h, w_d, out);
}

// column gather with optional row weights: out[el][k] = (w ? w[el] : 1) * J[el][cols[k]]   (out row stride ldo)
__global__ void __launch_bounds__(256)
k_gather_cols(const double* __restrict__ J, int64_t ld, int64_t n_el, int n, const int32_t* __restrict__ cols,
              const double* __restrict__ w, double* __restrict__ out, int64_t ldo)
{
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_el * n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t el = idx / n; const int k = (int)(idx - el * n);
        const double v = J[el * ld + cols[k]];
        out[el * ldo + k] = w ? v * w[el] : v;
    }
}

static int fill_hprobs_impl(b200_ctx* c, b200_atom* a, int32_t n1, const int32_t* p1, int32_t n2, const int32_t* p2,
                           int64_t nnz2, const int32_t* h_rows, const int32_t* h_a, const int32_t* h_b, const double* h_vals,
                           double* out, const double* w_h, const double* w_d, double* red_out) {
    // scratch kept in the context between calls
    DevBuf& d_p1 = c->hb[0];
    DevBuf& d_p2 = c->hb[1];
    DevBuf& d_out = c->hb[2];
    DevBuf& d_mrow = c->hb[3];
    DevBuf& d_mcol = c->hb[4];
    DevBuf& d_mval = c->hb[5];
    DevBuf& d_ktp = c->hb[6];
    DevBuf& d_kti = c->hb[7];
    DevBuf& d_D2 = c->hb[8];
    DevBuf& d_ukey = c->hb[9];
    DevBuf& d_kptr = c->hb[10];
    DevBuf& d_krow = c->hb[11];
    DevBuf& d_kval = c->hb[12];
    DevBuf& d_Dsel = c->hb[13];
    DevBuf& d_ktp2 = c->hb[14];
    DevBuf& d_kti2 = c->hb[15];
    DevBuf& d_need = c->hb[16];
    DevBuf& d_wh = c->hb[17];
    DevBuf& d_wd = c->hb[18];
    DevBuf& d_red = c->hb[19];
    DevBuf& d_j1 = c->hb[20];
    DevBuf& d_j2 = c->hb[21];
    DevBuf& d_cs = c->hb[22];

    if (!c || !a || (!out && !red_out) || (n1 > 0 && !p1) || (n2 > 0 && !p2)) return fail(B200_E_INVALID, "NULL argument");
    if (!a->has_model) return fail(B200_E_STATE, "b200_atom_set_model has not been called");
    if (!a->has_derivs) return fail(B200_E_STATE, "b200_atom_set_derivs has not been called");
    for (int i = 0; i < n1; ++i) if (p1[i] < 0 || p1[i] >= a->n_params) return fail(B200_E_INVALID, "p1[%d] out of range", i);
    for (int i = 0; i < n2; ++i) if (p2[i] < 0 || p2[i] >= a->n_params) return fail(B200_E_INVALID, "p2[%d] out of range", i);
    CU(cudaSetDevice(c->device));
    const int64_t nE = a->n_elements;
    if (nE == 0 || n1 == 0 || n2 == 0) return B200_OK;
    std::vector<int32_t> v1(p1, p1 + n1), v2(p2, p2 + n2);
    int rc;
    if ((rc = upload_vec(d_p1, v1, c->stream)) || (rc = upload_vec(d_p2, v2, c->stream))) return rc;
    CU(d_out.ensure((size_t)nE * n1 * n2 * 8));
    const int64_t per = nE * a->n_w * 8;
    int accumulate = 0;
    if (nnz2 > 0) {
        // second-derivative term first: out = sum_w W[el, w] d2M_w/dp1 dp2, entries grouped by (a, b)
        if (nnz2 >= ((int64_t)1 << 31)) return fail(B200_E_UNSUPPORTED, "second-derivative map too large");
        for (int64_t t = 0; t < nnz2; ++t) {
            if (h_rows[t] < 0 || h_rows[t] >= a->n_w) return fail(B200_E_INVALID, "D2 row %d out of range", h_rows[t]);
            if (h_a[t] < 0 || h_a[t] >= n1 || h_b[t] < 0 || h_b[t] >= n2) return fail(B200_E_INVALID, "D2 index out of range");
        }
        // order the entries by key = a * n2 + b, then by W row: counting sort over the n1 n2 keys + a short sort inside each key
        // (a comparison sort of all 5e5 entries of a CPTPLND rectangle was 40 of the 93 ms of a rectangle)
        auto keyof = [&](int64_t t) { return (int64_t)h_a[t] * n2 + h_b[t]; };
        std::vector<int64_t> idx((size_t)nnz2);
        {
            std::vector<int64_t> cnt((size_t)n1 * n2 + 1, 0);
            for (int64_t t = 0; t < nnz2; ++t) cnt[(size_t)keyof(t) + 1]++;
            for (size_t k = 0; k + 1 < cnt.size(); ++k) cnt[k + 1] += cnt[k];
            std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
            for (int64_t t = 0; t < nnz2; ++t) idx[(size_t)pos[(size_t)keyof(t)]++] = t;
            for (size_t k = 0; k + 1 < cnt.size(); ++k)
                if (cnt[k + 1] - cnt[k] > 1 && !std::is_sorted(idx.begin() + cnt[k], idx.begin() + cnt[k + 1], [&](int64_t x, int64_t y) { return h_rows[x] < h_rows[y]; }))
                    std::stable_sort(idx.begin() + cnt[k], idx.begin() + cnt[k + 1], [&](int64_t x, int64_t y) { return h_rows[x] < h_rows[y]; });
        }
        std::vector<int64_t> ukey; std::vector<int32_t> kptr, krow((size_t)nnz2); std::vector<double> kval((size_t)nnz2);
        for (int64_t t = 0; t < nnz2; ++t) {
            const int64_t i = idx[t];
            if (ukey.empty() || ukey.back() != keyof(i)) { ukey.push_back(keyof(i)); kptr.push_back((int32_t)t); }
            krow[t] = h_rows[i]; kval[t] = h_vals[i];
        }
        kptr.push_back((int32_t)nnz2);
        CU(c->w_buf.ensure((size_t)per));
        rc = compute_w(c, a, c->w_buf.as<double>(), nullptr);
        if (rc) return rc;
        const int64_t n12 = (int64_t)n1 * n2, ld12 = (n12 + 1) & ~(int64_t)1;
        if ((double)a->n_w * (double)ld12 * 8.0 <= 2.0e9 && n12 < ((int64_t)1 << 31)) {
            // out = W . D2 as a DMMA GEMM: D2 dense [n_w x n1 n2] (duplicates merged on the host) + K-chunk lists per column tile
            std::vector<int32_t> mrow; std::vector<int64_t> mcol; std::vector<double> mval;
            mrow.reserve((size_t)nnz2); mcol.reserve((size_t)nnz2); mval.reserve((size_t)nnz2);
            for (size_t k = 0; k < ukey.size(); ++k)
                for (int32_t t = kptr[k]; t < kptr[k + 1]; ++t) {
                    if (t > kptr[k] && krow[t] == mrow.back()) { mval.back() += kval[t]; continue; }
                    mrow.push_back(krow[t]); mcol.push_back(ukey[k]); mval.push_back(kval[t]);
                }
            std::vector<int32_t> ktp, kti;
            {   // entries are sorted by key = column: walk them tile by tile
                size_t pos = 0;
                build_kt_lists((int)n12, [&](int col, auto fn) {
                    while (pos < mcol.size() && mcol[pos] < col) ++pos;
                    for (size_t q = pos; q < mcol.size() && mcol[q] == col; ++q) fn(mrow[q]); }, ktp, kti);
            }
            if ((rc = upload_vec(d_mrow, mrow, c->stream)) || (rc = upload_vec(d_mcol, mcol, c->stream)) || (rc = upload_vec(d_mval, mval, c->stream)) ||
                (rc = upload_vec(d_ktp, ktp, c->stream)) || (rc = upload_vec(d_kti, kti, c->stream))) return rc;
            CU(d_D2.ensure((size_t)a->n_w * ld12 * 8));
            CU(cudaMemsetAsync(d_D2.p, 0, (size_t)a->n_w * ld12 * 8, c->stream));
            k_coo_to_dense<<<(unsigned)std::min<int64_t>(((int64_t)mrow.size() + 255) / 256 + 1, 4096), 256, 0, c->stream>>>(
                (int64_t)mrow.size(), d_mrow.as<int32_t>(), d_mcol.as<int64_t>(), d_mval.as<double>(), d_D2.as<double>(), ld12);
            c->launches++;
            rc = ab_device(c, c->w_buf.as<double>(), a->n_w, 0, d_D2.as<double>(), ld12, d_out.as<double>(), n12, 0, nE, (int)n12, (int)a->n_w,
                           d_ktp.as<int32_t>(), d_kti.as<int32_t>(), nullptr, false, 1);
            if (rc) return rc;
            CU(cudaStreamSynchronize(c->stream));
        } else {
            if ((rc = upload_vec(d_ukey, ukey, c->stream)) || (rc = upload_vec(d_kptr, kptr, c->stream)) ||
                (rc = upload_vec(d_krow, krow, c->stream)) || (rc = upload_vec(d_kval, kval, c->stream))) return rc;
            CU(cudaMemsetAsync(d_out.p, 0, (size_t)nE * n1 * n2 * 8, c->stream));
            const int nk = (int)ukey.size();
            dim3 gk((nk + 127) / 128, (unsigned)std::min<int64_t>(nE, 65535));
            k_hess_d2<<<gk, 128, 0, c->stream>>>(c->w_buf.as<double>(), a->n_w, nE, n1, n2, nk, d_ukey.as<int64_t>(), d_kptr.as<int32_t>(),
                                                d_krow.as<int32_t>(), d_kval.as<double>(), d_out.as<double>());
            c->launches++;
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
        }
        accumulate = 1;
    }
    // first-order part: out[:, a, :] (+)= Wp_a . D[:, p2] per tangent direction a -- dense D[:, p2] + K-chunk lists for the DMMA GEMM
    const int64_t ld2 = (n2 + 1) & ~1;
    const bool dense2 = (double)a->n_w * (double)ld2 * 8.0 <= 2.0e9;
    if (dense2) {
        std::vector<int32_t> ktp, kti;
        build_kt_lists(n2, [&](int col, auto fn) { for (int t = a->h_cptr[p2[col]]; t < a->h_cptr[p2[col] + 1]; ++t) fn(a->h_crow[t]); }, ktp, kti);
        if ((rc = upload_vec(d_ktp2, ktp, c->stream)) || (rc = upload_vec(d_kti2, kti, c->stream))) return rc;
        CU(d_Dsel.ensure((size_t)a->n_w * ld2 * 8));
        CU(cudaMemsetAsync(d_Dsel.p, 0, (size_t)a->n_w * ld2 * 8, c->stream));
        k_csc_to_dense<<<(unsigned)std::min(n2, 4096), 128, 0, c->stream>>>(a->cptr.as<int32_t>(), a->crow.as<int32_t>(), a->cval.as<double>(),
                                                                           d_p2.as<int32_t>(), n2, d_Dsel.as<double>(), ld2);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));          // the host K-chunk lists go out of scope
    }
    // blocks of the W row the contraction reads (rows of D[:, p2] with a non-zero): gates / state preparations / effects
    std::vector<unsigned char> need((size_t)a->n_ops + 2, 0);
    {
        const int64_t dd = (int64_t)a->dim * a->dim;
        for (int b2 = 0; b2 < n2; ++b2)
            for (int t = a->h_cptr[p2[b2]]; t < a->h_cptr[p2[b2] + 1]; ++t) {
                const int64_t w = a->h_crow[t];
                need[w < a->off_rho ? (size_t)(w / dd) : (w < a->off_eff ? (size_t)a->n_ops : (size_t)a->n_ops + 1)] = 1;
            }
        if ((rc = upload_vec(d_need, need, c->stream))) return rc;
    }
    // batch of tangent directions bounded by ~8 GB of W scratch
    const int B = (int)std::max<int64_t>(1, std::min<int64_t>(n1, ((int64_t)8 << 30) / std::max<int64_t>(per, 1)));
    CU(c->w_buf.ensure((size_t)B * per));
    CU(c->fd_models.ensure((size_t)B * a->n_w * 8));
    CU(c->fd_gt.ensure(std::max<size_t>((size_t)B * a->off_rho * 8, 16)));
    for (int a0 = 0; a0 < n1; a0 += B) {
        const int nb = std::min(B, n1 - a0);
        CU(cudaMemsetAsync(c->fd_models.p, 0, (size_t)nb * a->n_w * 8, c->stream));
        {   // zero only the blocks that are accumulated (and read)
            const int64_t dd = (int64_t)a->dim * a->dim;
            auto zero_cols = [&](int64_t col0, int64_t ncols) -> int {
                CU(cudaMemset2DAsync(c->w_buf.as<double>() + col0, (size_t)a->n_w * 8, 0, (size_t)ncols * 8, (size_t)nb * nE, c->stream));
                return B200_OK; };
            for (int g = 0; g < a->n_ops; ++g) if (need[g] && (rc = zero_cols(g * dd, dd))) return rc;
            if (need[a->n_ops] && (rc = zero_cols(a->off_rho, a->off_eff - a->off_rho))) return rc;
            if (need[a->n_ops + 1] && (rc = zero_cols(a->off_eff, a->n_w - a->off_eff))) return rc;
        }
        dim3 g1(4, nb);
        k_tangent_models<<<g1, 128, 0, c->stream>>>(a->n_w, d_p1.as<int32_t>(), a0, a->cptr.as<int32_t>(),
                                                    a->crow.as<int32_t>(), a->cval.as<double>(), c->fd_models.as<double>());
        c->launches++;
        if (a->n_ops) {
            dim3 g3((unsigned)std::min<int64_t>((a->off_rho + 255) / 256, 256), nb);
            k_transpose_gates<<<g3, 256, 0, c->stream>>>(c->fd_models.as<double>(), a->n_w, a->n_ops, a->dim, c->fd_gt.as<double>());
            c->launches++;
        }
        CU(cudaGetLastError());
        switch (a->dim) {
            case 4: rc = launch_w_tangent<4>(c, a, nb, c->fd_models.as<double>(), c->fd_gt.as<double>(), c->w_buf.as<double>(), d_need.as<unsigned char>()); break;
            case 16: rc = launch_w_tangent<16>(c, a, nb, c->fd_models.as<double>(), c->fd_gt.as<double>(), c->w_buf.as<double>(), d_need.as<unsigned char>()); break;
            case 64: rc = launch_w_tangent<64>(c, a, nb, c->fd_models.as<double>(), c->fd_gt.as<double>(), c->w_buf.as<double>(), d_need.as<unsigned char>()); break;
            case 256: rc = launch_w_tangent<256>(c, a, nb, c->fd_models.as<double>(), c->fd_gt.as<double>(), c->w_buf.as<double>(), d_need.as<unsigned char>()); break;
            default: rc = fail(B200_E_UNSUPPORTED, "dim %d", a->dim);
        }
        if (rc) return rc;
        if (dense2) {
            rc = ab_device(c, c->w_buf.as<double>(), a->n_w, nE * a->n_w, d_Dsel.as<double>(), ld2, d_out.as<double>() + (size_t)a0 * n2,
                           (int64_t)n1 * n2, n2, nE, n2, (int)a->n_w, d_ktp2.as<int32_t>(), d_kti2.as<int32_t>(), nullptr, accumulate != 0, nb);
            if (rc) return rc;
        } else {
            dim3 g4((n2 + 127) / 128, (unsigned)std::min<int64_t>(nE * nb, 65535));
            k_contract_hess<<<g4, 128, 0, c->stream>>>(c->w_buf.as<double>(), a->n_w, nE, nb, a0, n1, n2, d_p2.as<int32_t>(),
                                                       a->cptr.as<int32_t>(), a->crow.as<int32_t>(), a->cval.as<double>(),
                                                       d_out.as<double>(), accumulate);
            c->launches++;
            CU(cudaGetLastError());
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    if (red_out) {
        // MLE Hessian block (objectivefns.py:4914-4990 `_hessian_from_block` without omitted-outcome rows):
        //   red[a][b] = sum_el w_h[el] H[el][a][b] + w_d[el] J[el][p1[a]] J[el][p2[b]]      -- only n1 x n2 doubles leave the device
        const int64_t ld1 = (n1 + 1) & ~1, ld2 = (n2 + 1) & ~1, n12 = (int64_t)n1 * n2;
        CU(d_wh.ensure((size_t)nE * 8)); CU(d_wd.ensure((size_t)nE * 8)); CU(d_red.ensure((size_t)n12 * 8));
        CU(d_j1.ensure((size_t)nE * ld1 * 8)); CU(d_j2.ensure((size_t)nE * ld2 * 8));
        CU(cudaMemcpyAsync(d_wh.p, w_h, (size_t)nE * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_wd.p, w_d, (size_t)nE * 8, cudaMemcpyHostToDevice, c->stream));
        CU(c->out_buf.ensure(std::max<size_t>((size_t)nE * a->n_params * 8, 16)));
        rc = fill_dprobs_device(c, a, c->out_buf.as<double>(), a->n_params, nullptr, nullptr);
        if (rc) return rc;
        const int gg = (int)std::min<int64_t>((nE * std::max(n1, n2) + 255) / 256, (int64_t)c->sm_count * 16);
        k_gather_cols<<<gg, 256, 0, c->stream>>>(c->out_buf.as<double>(), a->n_params, nE, n1, d_p1.as<int32_t>(), d_wd.as<double>(), d_j1.as<double>(), ld1);
        k_gather_cols<<<gg, 256, 0, c->stream>>>(c->out_buf.as<double>(), a->n_params, nE, n2, d_p2.as<int32_t>(), nullptr, d_j2.as<double>(), ld2);
        c->launches += 2;
        CU(cudaGetLastError());
        // red = sum_el w_h[el] H[el]  (weighted column sums of the [nE x n1 n2] block, two deterministic passes) ...
        const int ns = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)c->sm_count * 8 * 256 + n12 - 1) / n12, (nE + 63) / 64));
        const int64_t rps = (nE + ns - 1) / ns;
        CU(d_cs.ensure((size_t)ns * n12 * 8));
        k_wcolsum<<<dim3((unsigned)((n12 + 255) / 256), (unsigned)ns), 256, 0, c->stream>>>(d_out.as<double>(), n12, nE, d_wh.as<double>(), rps, d_cs.as<double>());
        k_wcolsum_reduce<<<(unsigned)((n12 + 255) / 256), 256, 0, c->stream>>>(d_cs.as<double>(), n12, ns, d_red.as<double>());
        c->launches += 2;
        CU(cudaGetLastError());
        // ... += (w_d J1)^T J2 on the FP64 tensor cores
        rc = atb_device(c, d_j1.as<double>(), ld1, n1, d_j2.as<double>(), ld2, n2, nE, false, nullptr, d_red.as<double>(), n2, nullptr, true);
        if (rc) return rc;
        CU(cudaMemcpyAsync(red_out, d_red.p, (size_t)n1 * n2 * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        CU(cudaMemcpyAsync(out, d_out.p, (size_t)nE * n1 * n2 * 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------
// pinned host memory
// ------------------------------------------------------------------------------------------------
static std::mutex g_reg_mu;
static std::map<void*, size_t> g_registered;

extern "C" int b200_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return fail(B200_E_INVALID, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA, "cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e)); }
    return B200_OK;
}
extern "C" int b200_host_free(void* p) {
    if (!p) return B200_OK;
    CU(cudaFreeHost(p));
    return B200_OK;
}
extern "C" int b200_host_register(void* p, int64_t bytes) {
    if (!p || bytes <= 0) return fail(B200_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registered.find(p);
    if (it != g_registered.end()) {
        if (it->second >= (size_t)bytes) return B200_OK;
        cudaHostUnregister(p); g_registered.erase(it);
    }
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200_E_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    g_registered[p] = (size_t)bytes;
    return B200_OK;
}
extern "C" int b200_host_unregister(void* p) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_registered.find(p);
    if (it == g_registered.end()) return B200_OK;
    cudaHostUnregister(p);
    g_registered.erase(it);
    return B200_OK;
}
