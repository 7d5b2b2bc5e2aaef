"""
CPU ORACLE (numpy) -- TEST INFRASTRUCTURE ONLY.

A plain restatement of the reference's algorithm for the forward-simulation hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm may import
this module; nothing under ``pygsti_b200/`` does (the product path fails loudly without its CUDA
library and never falls back to this code).

Parity status: PINNED.  ``tests/test_oracle_cpu.py`` checks every function here against golden
vectors produced by running the reference itself (Cython ``MapForwardSimulator`` for probs and
finite-difference dprobs, ``MatrixForwardSimulator`` for analytic dprobs / hprobs) -- see
``tests/golden/make_golden.py``.

Reference lines restated:
  * prefix-table interpreter ......... pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:194-287
                                       (numpy twin: mapforwardsim_calc_generic.py:26-78)
  * dense op action  out = G v ....... pygsti/evotypes/densitymx/opcreps.cpp:40-54
                                       (pygsti/evotypes/densitymx_slow/opreps.py:88-92)
  * dense effect  p = E . v .......... pygsti/evotypes/densitymx/effectcreps.cpp:39-45
  * forward-difference Jacobian ...... pyx:290-383  (eps = 1e-7, (p2 - p)/eps, one table pass per parameter)
  * FD-of-FD Hessian ................. pygsti/forwardsims/mapforwardsim.py:394-438
  * analytic Jacobian ................ pygsti/forwardsims/matrixforwardsim.py:1059-1139
                                       (dp_dOps + dp_drhos + dp_dEs), product rule :729-792
"""
import numpy as np


# --------------------------------------------------------------------------------------
#  table helpers
# --------------------------------------------------------------------------------------
def expand_rows(t):
    """Full (prep index, op-sequence) of every table row, following the cache links
    (prefixtable.py:705-741 semantics: ``iStart`` indexes the *cache*, not rows)."""
    n_rows = t.row_dest.shape[0]
    cache_prep = {}
    cache_ops = {}
    preps = np.empty(n_rows, np.int32)
    seqs = []
    for k in range(n_rows):
        rem = t.row_ops[t.row_ptr[k]:t.row_ptr[k + 1]]
        if t.row_istart[k] < 0:
            prep = int(t.row_prep[k]); seq = rem
        else:
            s = int(t.row_istart[k])
            prep = cache_prep[s]; seq = np.concatenate([cache_ops[s], rem])
        if t.row_icache[k] >= 0:
            cache_prep[int(t.row_icache[k])] = prep
            cache_ops[int(t.row_icache[k])] = seq
        preps[k] = prep
        seqs.append(np.asarray(seq, dtype=np.int32))
    return preps, seqs


# --------------------------------------------------------------------------------------
#  probs : the reference interpreter, row by row, with the state cache
# --------------------------------------------------------------------------------------
def mapfill_probs(t, G, rho, E, out=None):
    """pyx:224-283: copy init state, acton chain, effect dots, store into cache."""
    d = t.dim
    if out is None:
        out = np.full(t.n_elements, np.nan)
    cache = np.zeros((max(t.cache_size, 1), d))
    for k in range(t.row_dest.shape[0]):
        v = rho[t.row_prep[k]] if t.row_istart[k] < 0 else cache[t.row_istart[k]]
        for g in t.row_ops[t.row_ptr[k]:t.row_ptr[k + 1]]:
            v = G[g] @ v                                   # opcreps.cpp:40-54
        for j in range(t.out_ptr[k], t.out_ptr[k + 1]):
            out[t.out_el[j]] = E[t.out_eff[j]] @ v         # effectcreps.cpp:39-45
        if t.row_icache[k] >= 0:
            cache[t.row_icache[k]] = v
    return out


# --------------------------------------------------------------------------------------
#  W : derivative of every outcome probability w.r.t. every dense member element
# --------------------------------------------------------------------------------------
def n_w(t):
    d = t.dim
    return t.n_ops * d * d + t.n_rho * d + t.n_eff * d


def w_matrix(t, G, rho, E):
    """W[el, w] = d p_el / d (member element w), by the adjoint recursion of SURVEY App. B.2,
    which equals matrixforwardsim.py:1059-1139 with every member fully parameterised."""
    d = t.dim
    W = np.zeros((t.n_elements, n_w(t)))
    off_rho = t.n_ops * d * d
    off_eff = off_rho + t.n_rho * d
    preps, seqs = expand_rows(t)
    for k in range(t.row_dest.shape[0]):
        seq = seqs[k]
        L = len(seq)
        s = np.empty((L + 1, d))
        s[0] = rho[preps[k]]
        for m in range(L):
            s[m + 1] = G[seq[m]] @ s[m]
        for j in range(t.out_ptr[k], t.out_ptr[k + 1]):
            el = t.out_el[j]; ei = t.out_eff[j]
            W[el, off_eff + ei * d: off_eff + (ei + 1) * d] += s[L]
            e = E[ei].copy()
            for m in range(L - 1, -1, -1):
                g = seq[m]
                W[el, g * d * d:(g + 1) * d * d] += np.outer(e, s[m]).ravel()
                e = G[g].T @ e
            W[el, off_rho + preps[k] * d: off_rho + (preps[k] + 1) * d] += e
    return W


def dense_D(D):
    M = np.zeros((D.n_w, D.n_params))
    np.add.at(M, (D.rows, D.cols), D.vals)
    return M


def dprobs_analytic(t, G, rho, E, D):
    """J = W . D   (chain rule through the dense member elements)."""
    return w_matrix(t, G, rho, E) @ dense_D(D)


def dprobs_product_rule(t, G, rho, E, D):
    """Independent small-case check in the Matrix simulator's own form:
    dProd/dtheta = sum_k G_L..G_{k+1} (dG_k/dtheta) G_{k-1}..G_1   (matrixforwardsim.py:729-792),
    then  e . dProd . rho + e . Prod . drho + dE . Prod . rho   (:1059-1139).  O(L^2) -- tiny inputs only."""
    d = t.dim
    Dm = dense_D(D)
    Np = D.n_params
    off_rho = t.n_ops * d * d
    off_eff = off_rho + t.n_rho * d
    dG = Dm[:off_rho].reshape(t.n_ops, d, d, Np)
    drho = Dm[off_rho:off_eff].reshape(t.n_rho, d, Np)
    dE = Dm[off_eff:].reshape(t.n_eff, d, Np)
    J = np.zeros((t.n_elements, Np))
    preps, seqs = expand_rows(t)
    for k in range(t.row_dest.shape[0]):
        seq = seqs[k]
        prod = np.eye(d)
        dprod = np.zeros((Np, d, d))
        for g in seq:                                 # prod <- G prod ; dprod <- dG prod + G dprod
            dprod = np.einsum('ijp,jk->pik', dG[g], prod) + np.einsum('ij,pjk->pik', G[g], dprod)
            prod = G[g] @ prod
        r = rho[preps[k]]
        for j in range(t.out_ptr[k], t.out_ptr[k + 1]):
            el = t.out_el[j]; ei = t.out_eff[j]
            J[el] = (np.einsum('i,pij,j->p', E[ei], dprod, r)
                     + (E[ei] @ prod) @ drho[preps[k]]
                     + (prod @ r) @ dE[ei])
    return J


# --------------------------------------------------------------------------------------
#  reference-semantics finite differences (exact for members that are LINEAR in their parameters:
#  full / TP / static members -- set_parameter_values then writes M + eps*dM/dtheta_p in place)
# --------------------------------------------------------------------------------------
def _perturbed(t, G, rho, E, Dm, p, eps):
    d = t.dim
    off_rho = t.n_ops * d * d
    off_eff = off_rho + t.n_rho * d
    col = Dm[:, p] * eps
    return (G + col[:off_rho].reshape(G.shape), rho + col[off_rho:off_eff].reshape(rho.shape),
            E + col[off_eff:].reshape(E.shape))


def dprobs_fd_linear(t, G, rho, E, D, eps=1e-7, cols=None):
    """pyx:349-378: base pass, then one full table pass per parameter, (probs2 - probs)/eps."""
    Dm = dense_D(D)
    cols = range(D.n_params) if cols is None else cols
    p0 = mapfill_probs(t, G, rho, E)
    J = np.zeros((t.n_elements, len(cols)))
    for c, p in enumerate(cols):
        G2, r2, E2 = _perturbed(t, G, rho, E, Dm, p, eps)
        J[:, c] = (mapfill_probs(t, G2, r2, E2) - p0) / eps
    return J


def hprobs_linear(t, G, rho, E, D):
    """Exact Hessian for members linear in their parameters (d2M = 0):
    H[el,p,q] = sum_{w,w'} D[w,p] D[w',q] d2 p_el / dw dw'.  Computed as the directional derivative of
    the analytic Jacobian along each parameter (forward-over-reverse), O(Np) Jacobians -- tiny inputs only."""
    Dm = dense_D(D)
    Np = D.n_params
    d = t.dim
    off_rho = t.n_ops * d * d
    off_eff = off_rho + t.n_rho * d
    H = np.zeros((t.n_elements, Np, Np))
    preps, seqs = expand_rows(t)
    dG = Dm[:off_rho].reshape(t.n_ops, d, d, Np)
    drho = Dm[off_rho:off_eff].reshape(t.n_rho, d, Np)
    dE = Dm[off_eff:].reshape(t.n_eff, d, Np)
    for k in range(t.row_dest.shape[0]):
        seq = seqs[k]; L = len(seq); r = preps[k]
        # states and their tangents ds[m][:, p]
        s = [rho[r]]; ds = [drho[r]]
        for m in range(L):
            g = seq[m]
            ds.append(G[g] @ ds[m] + np.einsum('ijp,j->ip', dG[g], s[m]))
            s.append(G[g] @ s[m])
        for j in range(t.out_ptr[k], t.out_ptr[k + 1]):
            el = t.out_el[j]; ei = t.out_eff[j]
            e = E[ei].copy(); de = dE[ei].copy()          # de[:, p] = d e / d theta_p
            Hel = np.einsum('iq,ip->pq', dE[ei], ds[L])    # d/dp of  s_L . dE[:,q]
            for m in range(L - 1, -1, -1):
                g = seq[m]
                # J_q += e^T dG_q s_m  ->  d/dp: de_p^T dG_q s_m + e^T dG_q ds_m,p
                Hel += np.einsum('ip,ijq,j->pq', de, dG[g], s[m]) + np.einsum('i,ijq,jp->pq', e, dG[g], ds[m])
                de = G[g].T @ de + np.einsum('ijp,i->jp', dG[g], e)
                e = G[g].T @ e
            Hel += np.einsum('ip,iq->pq', de, drho[r])     # J_q += e_0 . drho[:,q]
            H[el] = Hel
    return H


def hprobs_general(t, G, rho, E, D, H2=None, p1=None, p2=None):
    """Hessian rectangle for arbitrary members:  H[el, a, b] = (linear-member part)[el, p1[a], p2[b]]
    + sum_w W[el, w] d2M_w/dp1[a]dp2[b]   (== MatrixForwardSimulator._hprobs_from_rho_e, matrixforwardsim.py:1141-1287,
    where the second term is the `_hoperation` / SPAM-hessian contribution).  H2: (rows, a, b, vals) COO or None."""
    Np = D.n_params
    p1 = np.arange(Np) if p1 is None else np.asarray(p1)
    p2 = np.arange(Np) if p2 is None else np.asarray(p2)
    H = hprobs_linear(t, G, rho, E, D)[:, p1][:, :, p2]
    if H2 is not None and len(H2[0]):
        W = w_matrix(t, G, rho, E)
        rows, a, b, vals = H2
        np.add.at(H, (slice(None), a, b), W[:, rows] * vals[None, :])
    return H


# --------------------------------------------------------------------------------------
#  factored gates: layer operations as products of small operations embedded on 1-2 qubits
#  (OpCRep_Composed of OpCRep_Embedded, pygsti/evotypes/densitymx/opcreps.cpp:242-276, 93-158; the derivative of an
#  EmbeddedOp is its embedded operation's own derivative, pygsti/modelmembers/operations/embeddedop.py deriv_wrt_params)
# --------------------------------------------------------------------------------------
def _embed(fm, f, d):
    """dense d x d matrix of factor f and, per small-matrix entry (a, b), the (row, col) index pairs it occupies."""
    nq = fm.n_qubits; k = int(fm.f_nq[f]); ds = 4 ** k
    small = fm.mats[fm.f_moff[f]:fm.f_moff[f] + ds * ds].reshape(ds, ds)
    shifts = [2 * (nq - 1 - int(q)) for q in fm.f_targets[f, :k]]
    idx = np.arange(d)
    tmask = 0
    for s in shifts:
        tmask |= 3 << s
    t = np.zeros(d, dtype=np.int64)
    for s in shifts:
        t = (t << 2) | ((idx >> s) & 3)
    rest = idx & ~tmask
    F = np.zeros((d, d))
    for i in range(d):
        for j in range(d):
            if rest[i] == rest[j]:
                F[i, j] = small[t[i], t[j]]
    return F, t, rest


def w_matrix_factored(t, fm):
    """Wf[el, w] = d p_el / d (factor-space element w), w over [fm.mats | rho | E]: for a factor F = embed(g) between the
    forward state s (before it) and the backward vector e (after it), d p / d g[a][b] = sum_{i ~ a, j ~ b, rest(i) = rest(j)} e_i s_j."""
    d = t.dim
    n_mats = fm.mats.size
    Wf = np.zeros((t.n_elements, n_mats + (t.n_rho + t.n_eff) * d))
    off_rho = n_mats; off_eff = n_mats + t.n_rho * d
    n_fac = fm.f_nq.shape[0]
    emb = [_embed(fm, f, d) for f in range(n_fac)]
    preps, seqs = expand_rows(t)
    for k in range(t.row_dest.shape[0]):
        steps = [f for g in seqs[k] for f in range(fm.op_fptr[g], fm.op_fptr[g + 1])]      # factors in order of application
        s = [fm.rho[preps[k]]]
        for f in steps:
            s.append(emb[f][0] @ s[-1])
        for j in range(t.out_ptr[k], t.out_ptr[k + 1]):
            el = t.out_el[j]; ei = t.out_eff[j]
            Wf[el, off_eff + ei * d: off_eff + (ei + 1) * d] += s[-1]
            e = fm.E[ei].copy()
            for m in range(len(steps) - 1, -1, -1):
                f = steps[m]; F, tt, rest = emb[f]
                ds = 4 ** int(fm.f_nq[f])
                outer = np.outer(e, s[m]) * (rest[:, None] == rest[None, :])
                blk = np.zeros((ds, ds))
                np.add.at(blk, (tt[:, None].repeat(d, 1), tt[None, :].repeat(d, 0)), outer)
                Wf[el, fm.f_moff[f]:fm.f_moff[f] + ds * ds] += blk.ravel()
                e = F.T @ e
            Wf[el, off_rho + preps[k] * d: off_rho + (preps[k] + 1) * d] += e
    return Wf


def dprobs_factored(t, fm, Df):
    """J = Wf . Df   (chain rule through the entries of the small embedded operations)."""
    return w_matrix_factored(t, fm) @ dense_D(Df)
