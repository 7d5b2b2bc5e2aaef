"""
ORACLE (test infrastructure only -- never imported by pygsti_b200/): numpy restatement of the host-side member update for
Lindblad-parameterised operations, the second half of SURVEY.md 8f rank 3 (on-device model update; product side:
b200_lindblad_members, pygsti_b200/csrc/kernels_lindblad.cuh).

What the reference does per parameter-vector update of a `CPTPLND` / `H+S` / `GLND` gate  G = exp(L(theta)) . G_target
(`ComposedOp([static target, ExpErrorgenOp(LindbladErrorgen)])`):

  * coefficients            c(theta)  -- `LindbladCoefficientBlock.from_vector`  (lindbladcoefficients.py:897-911), and their
                            Jacobian  dc/dtheta  -- `deriv_wrt_params` (:913-926); cheap, parameterisation-specific: KEPT on the host
  * error generator         L = Re sum_i c_i B_i          -- `LindbladErrorgen._update_rep`, dense branch (lindbladerrorgen.py:700-708)
                            dL/dtheta_p = Re sum_i dc_i/dtheta_p B_i   -- `superop_deriv_wrt_params` (lindbladcoefficients.py:943-...)
  * exponential             E = expm(L)                    -- `ExpErrorgenOp._update_rep` (experrorgenop.py:114-125, scipy.linalg.expm)
                            dE/dtheta_p = Dexp(L)[dL/dtheta_p]  -- `ExpErrorgenOp.deriv_wrt_params` (:213-262, `_d_exp_x` series)
  * composition             G = E . G_target, dG = dE . G_target   -- `ComposedOp.to_dense / deriv_wrt_params` (composedop.py)

At BASELINE config 4 the three lower items cost 0.27 s (11 expm) + 0.8 s (einsum + checks) per update on the host against ~10 ms of GPU
work per Jacobian.  The device version takes (B, c, dc/dtheta, static parts) and produces the members and their derivatives; this file
states the arithmetic it has to reproduce (Frechet derivative by the block-triangular exponential, an algorithm independent of the reference's
commutator series) and `tests/test_oracle_cpu.py::test_lindblad_oracle_*` pins it against the reference run in this container.
"""
import numpy as np
import scipy.linalg as la


def errorgen_from_coefficients(c, B):
    """L = Re sum_i c_i B_i ;  c: complex [n_coeff], B: complex [n_coeff, d, d]  (lindbladerrorgen.py:700-708)."""
    return np.real(np.tensordot(c, B, axes=(0, 0)))


def errorgen_derivs(dc, B):
    """dL[p] = Re sum_i dc[i, p] B_i ;  dc: complex [n_coeff, n_params]  ->  [n_params, d, d]."""
    return np.real(np.tensordot(dc.T, B, axes=(1, 0)))


def expm_and_frechet(L, dL):
    """E = expm(L) and dE[p] = Dexp(L)[dL[p]] through exp([[L, dL_p], [0, L]]) = [[E, dE_p], [0, E]]."""
    d = L.shape[0]
    E = la.expm(L)
    dE = np.empty((dL.shape[0], d, d))
    blk = np.zeros((2 * d, 2 * d))
    blk[:d, :d] = L
    blk[d:, d:] = L
    for p in range(dL.shape[0]):
        blk[:d, d:] = dL[p]
        dE[p] = la.expm(blk)[:d, d:]
    return E, dE


def composed_gate(c, dc, B, G_target):
    """Dense G = expm(L) . G_target and dG/dtheta [d*d, n_params] (row-major flattening, as `deriv_wrt_params` returns)."""
    L = errorgen_from_coefficients(c, B)
    dL = errorgen_derivs(dc, B)
    E, dE = expm_and_frechet(L, dL)
    G = E @ G_target
    dG = np.einsum('pij,jk->pik', dE, G_target).reshape(dL.shape[0], -1).T
    return G, dG


# ---- extraction of (B, c, dc/dtheta) from a reference LindbladErrorgen (host side; pyGSTi objects in, numpy out) -------------------
def lindblad_inputs(errorgen):
    """(B, c, dc) of a dense-rep `LindbladErrorgen`: the combined term superoperators, the concatenated block data and its Jacobian
    w.r.t. the errorgen's own parameter vector (blocks own consecutive parameter ranges, lindbladerrorgen.py:882-900)."""
    B = np.asarray(errorgen.combined_lindblad_term_superops)
    c = np.concatenate([np.asarray(blk.block_data).ravel() for blk in errorgen.coefficient_blocks])
    n_par = errorgen.num_params
    dc = np.zeros((c.size, n_par), complex)
    row = col = 0
    for blk in errorgen.coefficient_blocks:
        # at the errorgen's OWN parameter values: a block's to_vector() re-derives them from block_data (Cholesky factor), which
        # can land on an equivalent but different point (lindbladerrorgen.py:1371 passes self.paramvals too)
        J = np.asarray(blk.deriv_wrt_params(np.asarray(errorgen.paramvals[col:col + blk.num_params])))
        J = J.reshape(-1, blk.num_params) if blk.num_params else np.zeros((np.asarray(blk.block_data).size, 0))   # ('other' blocks: [n, n, n, n])
        dc[row:row + J.shape[0], col:col + J.shape[1]] = J
        row += J.shape[0]
        col += J.shape[1]
    assert row == c.size and col == n_par
    return B, c, dc


def composed_state(c, dc, B, rho0):
    """`ComposedState` (static state followed by exp(L)): rho = E rho0, d rho / d theta = dE rho0  -> ([d], [d, n_params])
    (pygsti/modelmembers/states/composedstate.py: to_dense, deriv_wrt_params)."""
    E, dE = expm_and_frechet(errorgen_from_coefficients(c, B), errorgen_derivs(dc, B))
    return E @ rho0, np.einsum('pij,j->ip', dE, rho0)


def composed_effect(c, dc, B, e0):
    """`ComposedPOVMEffect` (exp(L) acts on the state before the static effect): e = E^T e0, de/dtheta = dE^T e0
    (pygsti/modelmembers/povms/composedeffect.py:117-160, 284-311)."""
    E, dE = expm_and_frechet(errorgen_from_coefficients(c, B), errorgen_derivs(dc, B))
    return E.T @ e0, np.einsum('pji,j->ip', dE, e0)
