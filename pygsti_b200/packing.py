"""
Host-side packing: pyGSTi layout atoms and models -> the flat arrays the C ABI takes.

Nothing here imports pyGSTi; the functions are duck-typed against the reference objects:

* ``pack_atom(atom)``  reads exactly the members the reference's Cython conversion code reads
  (``layout_atom.table.contents``, ``rho_labels``, ``op_labels``, ``full_effect_labels``,
  ``elbl_indices_by_expcircuit``, ``elindices_by_expcircuit``; reference
  ``pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:55-101,163-181``) and produces int32 CSR
  tables in the reference's own row format ``[iDest, iStart, iCache, prep?, ops...]`` (pyx:55-77).
* ``pack_model(model, atom)``  pulls the dense superoperators / superkets through
  ``model._circuit_layer_operator(lbl, typ).to_dense('HilbertSchmidt')`` (same access path as
  pyx:164-167; dense arrays as in ``matrixforwardsim.py:700,1039-1045``).
* ``pack_derivs(model, atom, param_slice)``  builds the sparse "member-element -> parameter" matrix D
  from ``member.deriv_wrt_params()`` and ``member.gpindices`` (same inputs as
  ``matrixforwardsim.py:126-170, 1111-1136``).

Index conventions ("W space", one column per dense member element):
    op g, element (i, j)   ->  g*d*d + i*d + j          (row-major flattening, matrixforwardsim.py:114-124)
    prep r, element i      ->  n_ops*d*d + r*d + i
    effect e, element i    ->  n_ops*d*d + n_rho*d + e*d + i
"""
import weakref
from dataclasses import dataclass, field
import numpy as np


@dataclass
class AtomTables:
    """Integer tables of one layout atom, in the reference's prefix-table row format."""
    dim: int
    n_ops: int
    n_rho: int
    n_eff: int
    n_elements: int
    cache_size: int
    row_dest: np.ndarray      # int32 [n_rows]   expanded-circuit index (iDest)
    row_istart: np.ndarray    # int32 [n_rows]   cache slot to start from, -1 = start from a prep
    row_icache: np.ndarray    # int32 [n_rows]   cache slot to store the final state in, -1 = none
    row_prep: np.ndarray      # int32 [n_rows]   prep index when row_istart == -1, else -1
    row_ptr: np.ndarray       # int32 [n_rows+1] CSR pointer into row_ops
    row_ops: np.ndarray       # int32 [nnz]      op indices of the row's remainder (prep label removed)
    out_ptr: np.ndarray       # int32 [n_rows+1] CSR pointer (per ROW, in row order) into out_eff/out_el
    out_eff: np.ndarray       # int32 [n_out]    effect index of each outcome
    out_el: np.ndarray        # int32 [n_out]    atom-local element index of each outcome

    @property
    def n_rows(self):
        return int(self.row_dest.shape[0])

    def num_state_propagations(self):
        """Same count as the reference's ``PrefixTable.num_state_propagations`` (prefixtable.py:106)."""
        return int(self.row_ops.shape[0])

    def to_dict(self, prefix=""):
        d = {prefix + "meta": np.array([self.dim, self.n_ops, self.n_rho, self.n_eff, self.n_elements,
                                        self.cache_size], dtype=np.int64)}
        for k in ("row_dest", "row_istart", "row_icache", "row_prep", "row_ptr", "row_ops",
                  "out_ptr", "out_eff", "out_el"):
            d[prefix + k] = getattr(self, k)
        return d

    @classmethod
    def from_dict(cls, d, prefix=""):
        m = d[prefix + "meta"]
        kw = {k: np.ascontiguousarray(d[prefix + k], dtype=np.int32)
              for k in ("row_dest", "row_istart", "row_icache", "row_prep", "row_ptr", "row_ops",
                        "out_ptr", "out_eff", "out_el")}
        return cls(dim=int(m[0]), n_ops=int(m[1]), n_rho=int(m[2]), n_eff=int(m[3]),
                   n_elements=int(m[4]), cache_size=int(m[5]), **kw)


@dataclass
class ModelTensors:
    G: np.ndarray      # f64 [n_ops, d, d] row-major superoperators
    rho: np.ndarray    # f64 [n_rho, d]
    E: np.ndarray      # f64 [n_eff, d]


@dataclass
class DerivMap:
    """Sparse D[w, p] = d(member element w)/d(parameter p), COO with duplicates already summed."""
    n_w: int
    n_params: int
    rows: np.ndarray   # int32 [nnz]  W-space index
    cols: np.ndarray   # int32 [nnz]  parameter column (relative to the requested slice)
    vals: np.ndarray   # f64   [nnz]


def atom_labels(atom):
    """(rho_labels, op_labels, effect_labels) in the exact order the reference enumerates them
    (pyx:163-167: ``enumerate(layout_atom.rho_labels)``, ``enumerate(layout_atom.op_labels)`` and
    iteration over the ``full_effect_labels`` set object, whose order defines ``elabel_lookup``,
    maplayout.py:104-105)."""
    return list(atom.rho_labels), list(atom.op_labels), list(atom.full_effect_labels)


def pack_atom(atom, dim):
    rho_labels, op_labels, eff_labels = atom_labels(atom)
    rho_lookup = {lbl: i for i, lbl in enumerate(rho_labels)}
    op_lookup = {lbl: i for i, lbl in enumerate(op_labels)}
    contents = atom.table.contents
    n_rows = len(contents)
    row_dest = np.empty(n_rows, np.int32)
    row_istart = np.empty(n_rows, np.int32)
    row_icache = np.empty(n_rows, np.int32)
    row_prep = np.full(n_rows, -1, np.int32)
    row_ptr = np.zeros(n_rows + 1, np.int64)
    ops = []
    out_ptr = np.zeros(n_rows + 1, np.int64)
    out_eff = []
    out_el = []
    elbl = atom.elbl_indices_by_expcircuit
    elind = atom.elindices_by_expcircuit
    for k, (i_dest, i_start, remainder, i_cache) in enumerate(contents):
        row_dest[k] = i_dest
        row_icache[k] = -1 if i_cache is None else i_cache
        if i_start is None:
            row_istart[k] = -1
            row_prep[k] = rho_lookup[remainder[0]]
            rem = remainder[1:]
        else:
            row_istart[k] = i_start
            rem = remainder
        ops.extend(op_lookup[gl] for gl in rem)
        row_ptr[k + 1] = len(ops)
        out_eff.extend(elbl[i_dest])
        out_el.extend(elind[i_dest])
        out_ptr[k + 1] = len(out_eff)
    if row_ptr[-1] >= 2**31 or out_ptr[-1] >= 2**31:
        raise MemoryError("layout atom too large for int32 tables")
    return AtomTables(
        dim=int(dim), n_ops=len(op_labels), n_rho=len(rho_labels), n_eff=len(eff_labels),
        n_elements=int(atom.num_elements), cache_size=int(atom.cache_size),
        row_dest=row_dest, row_istart=row_istart, row_icache=row_icache, row_prep=row_prep,
        row_ptr=row_ptr.astype(np.int32), row_ops=np.asarray(ops, dtype=np.int32),
        out_ptr=out_ptr.astype(np.int32), out_eff=np.asarray(out_eff, dtype=np.int32),
        out_el=np.asarray(out_el, dtype=np.int32))


def _members(model, atom):
    rho_labels, op_labels, eff_labels = atom_labels(atom)
    M = model._circuit_layer_operator
    ops = [M(l, 'op') for l in op_labels]
    rhos = [M(l, 'prep') for l in rho_labels]
    effs = [M(l, 'povm') for l in eff_labels]
    return ops, rhos, effs


def pack_model(model, atom, dim):
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    G = np.empty((len(ops), d, d), dtype=np.float64)
    for i, op in enumerate(ops):
        G[i] = np.asarray(op.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(d, d)
    rho = np.empty((len(rhos), d), dtype=np.float64)
    for i, r in enumerate(rhos):
        rho[i] = np.asarray(r.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(d)
    E = np.empty((len(effs), d), dtype=np.float64)
    for i, e in enumerate(effs):
        E[i] = np.asarray(e.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(d)
    return ModelTensors(G=G, rho=rho, E=E)


def _gp_array(gpindices):
    if gpindices is None:
        return np.zeros(0, dtype=np.int64)
    if isinstance(gpindices, slice):
        if gpindices.start is None or gpindices.stop is None:
            return np.zeros(0, dtype=np.int64)
        return np.arange(gpindices.start, gpindices.stop, gpindices.step or 1, dtype=np.int64)
    return np.asarray(gpindices, dtype=np.int64)


def param_slice_to_array(param_slice, num_params):
    if param_slice is None:
        return np.arange(num_params, dtype=np.int64)
    return _gp_array(param_slice)


def pack_derivs(model, atom, dim, param_indices=None):
    """Sparse D restricted to ``param_indices`` (columns renumbered 0..len-1 in the given order).

    With a parameter interposer (``model._param_interposer``; matrixforwardsim.py:140-152, 1089-1105)
    D is first assembled w.r.t. *op* parameters and then multiplied by d(op_params)/d(model_params).
    """
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    n_w = len(ops) * d * d + len(rhos) * d + len(effs) * d
    interposer = getattr(model, '_param_interposer', None)
    n_model_params = int(model.num_params)
    n_op_params = int(interposer.num_op_params) if interposer is not None else n_model_params

    rows, cols, vals = [], [], []
    off = 0
    for group, size in ((ops, d * d), (rhos, d), (effs, d)):
        for m in group:
            gp = _gp_array(m.gpindices)
            if gp.size:
                dM = np.asarray(m.deriv_wrt_params(), dtype=np.float64).reshape(size, gp.size)
                r, c = np.nonzero(dM)
                rows.append(off + r)
                cols.append(gp[c])
                vals.append(dM[r, c])
            off += size
    if rows:
        rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    else:
        rows = np.zeros(0, np.int64); cols = np.zeros(0, np.int64); vals = np.zeros(0, np.float64)

    import scipy.sparse as sps
    D = sps.coo_matrix((vals, (rows, cols)), shape=(n_w, n_op_params)).tocsr()  # sums duplicates
    if interposer is not None:
        D = (D @ sps.csr_matrix(interposer.deriv_op_params_wrt_model_params())).tocsr()
    pidx = param_slice_to_array(param_indices, n_model_params)
    if not (pidx.size == n_model_params and np.array_equal(pidx, np.arange(n_model_params))):
        D = D[:, pidx]
    D = D.tocoo()
    keep = D.data != 0.0
    return DerivMap(n_w=n_w, n_params=int(pidx.size),
                    rows=D.row[keep].astype(np.int32), cols=D.col[keep].astype(np.int32),
                    vals=D.data[keep].astype(np.float64))


class HessMap:
    """Sparse second derivatives of the dense member elements for one Hessian rectangle:
    entries (w, a, b, val) = d2 M_w / d theta_{p1[a]} d theta_{p2[b]}  (zero for members linear in their parameters)."""

    def __init__(self, n_w, n1, n2, rows, a, b, vals):
        self.n_w, self.n1, self.n2 = int(n_w), int(n1), int(n2)
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.a = np.ascontiguousarray(a, dtype=np.int32)
        self.b = np.ascontiguousarray(b, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)

    @property
    def nnz(self):
        return int(self.rows.size)


# A member's `hessian_wrt_params` costs the same whether a rectangle or the whole (n_p x n_p) Hessian is asked for (9.9 s for a
# 64 x 64 rectangle of a 2-qubit CPTPLND gate, Np = 240), and the MLE Hessian walks over ~16 rectangles inside every member's
# parameter block (`_iter_atom_hprobs_by_rectangle`, distforwardsim.py:304-340): the full member Hessian is computed once per
# parameter vector and the rectangles are sliced from it.  Bounded cache (bytes), keyed by the member object and its parameters.
_HESS_CACHE = {}
_HESS_CACHE_MAX_BYTES = 2 << 30


def _member_hessian(m, size, n_p):
    key = id(m)
    pv = np.asarray(m.to_vector(), dtype=np.float64).tobytes()
    ent = _HESS_CACHE.get(key)
    if ent is not None and ent[0] == pv and ent[1]() is m:
        return ent[2]
    H = np.asarray(m.hessian_wrt_params())
    if np.iscomplexobj(H):
        H = H.real
    H = np.ascontiguousarray(H, dtype=np.float64).reshape(size, n_p, n_p)
    _HESS_CACHE[key] = (pv, weakref.ref(m), H)
    total = sum(e[2].nbytes for e in _HESS_CACHE.values())
    for k in list(_HESS_CACHE):                       # oldest first (insertion order) until under the cap
        if total <= _HESS_CACHE_MAX_BYTES or k == key:
            continue
        total -= _HESS_CACHE.pop(k)[2].nbytes
    return H


def pack_hessians(model, atom, dim, param_indices1=None, param_indices2=None):
    """Second derivatives of every member with ``has_nonzero_hessian()`` through its own ``hessian_wrt_params`` -- the
    quantity MatrixForwardSimulator._hoperation / _hprobs_from_rho_e read (matrixforwardsim.py:172-218, 1196-1237) --
    restricted to the rectangle (param_indices1 x param_indices2).  Parameter interposers are not supported here
    (the caller falls back to the reference driver)."""
    if getattr(model, '_param_interposer', None) is not None:
        raise NotImplementedError("pack_hessians: parameter interposer")
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    n_w = len(ops) * d * d + len(rhos) * d + len(effs) * d
    n_model_params = int(model.num_params)
    p1 = param_slice_to_array(param_indices1, n_model_params)
    p2 = param_slice_to_array(param_indices2, n_model_params)
    pos1 = np.full(n_model_params, -1, np.int64); pos1[p1] = np.arange(p1.size)
    pos2 = np.full(n_model_params, -1, np.int64); pos2[p2] = np.arange(p2.size)
    rows, aa, bb, vals = [], [], [], []
    off = 0
    for group, size in ((ops, d * d), (rhos, d), (effs, d)):
        for m in group:
            gp = _gp_array(m.gpindices)
            if gp.size and m.has_nonzero_hessian():
                l1 = np.flatnonzero(pos1[gp] >= 0); l2 = np.flatnonzero(pos2[gp] >= 0)
                if l1.size and l2.size:
                    H = _member_hessian(m, size, gp.size)[:, l1][:, :, l2]
                    r, i, j = np.nonzero(H)
                    rows.append(off + r); aa.append(pos1[gp[l1[i]]]); bb.append(pos2[gp[l2[j]]]); vals.append(H[r, i, j])
            off += size
    cat = lambda x, dt: np.concatenate(x).astype(dt) if x else np.zeros(0, dt)
    return HessMap(n_w, p1.size, p2.size, cat(rows, np.int32), cat(aa, np.int32), cat(bb, np.int32), cat(vals, np.float64))


# ---------------------------------------------------------------------------------------------------------------------
# Lindblad-parameterised members (second part -- host side of the on-device model update, `device_lindblad`).
# A `CPTPLND` / `H+S` / `GLND` member is  exp(L(theta))  composed with a static object, L = Re sum_i c_i(theta) B_i.
# What stays on the host is cheap and parameterisation-specific (c, dc/dtheta: LindbladCoefficientBlock.from_vector /
# deriv_wrt_params, lindbladcoefficients.py:897-926); the device takes over the dense algebra that dominates an update on the
# host (the 240-term contractions, expm, its Frechet derivative: b200_lindblad_members, csrc/kernels_lindblad.cuh).  This packer extracts exactly
# the inputs of that algebra; tests/test_plugin_cpu.py assembles M and D from them with the oracle and compares with
# pack_model / pack_derivs.
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class LindbladMember:
    kind: str            # 'op' | 'rho' | 'eff'
    w_offset: int        # first W-space index of the member (d*d entries for an op, d for a state / effect)
    static: np.ndarray   # op: static target matrix [d, d] (G = exp(L) target); rho: static state [d]; eff: static effect [d]
    errgen: int          # index into LindbladInputs.errgens (effects of one POVM share theirs)
    gpindices: np.ndarray  # int64 [n_p]: model parameter of each errorgen parameter


@dataclass
class LindbladErrgen:
    B_re: np.ndarray     # [n_coeff, d, d]   constant term superoperators (real / imaginary parts)
    B_im: np.ndarray
    c: np.ndarray        # complex [n_coeff]          coefficients at the current parameters
    dc: np.ndarray       # complex [n_coeff, n_p]     their Jacobian w.r.t. the errorgen's parameters


@dataclass
class LindbladInputs:
    members: list        # LindbladMember, in W-space order
    errgens: list        # LindbladErrgen
    host_members: list   # (kind, w_offset, member) of everything else (packed densely on the host as before)


def _errgen_inputs(errorgen):
    B = np.asarray(errorgen.combined_lindblad_term_superops)
    c = np.concatenate([np.asarray(blk.block_data).ravel() for blk in errorgen.coefficient_blocks]).astype(complex)
    n_par = int(errorgen.num_params)
    dc = np.zeros((c.size, n_par), complex)
    row = col = 0
    for blk in errorgen.coefficient_blocks:
        nb = int(blk.num_params)
        size = int(np.asarray(blk.block_data).size)
        if nb:
            # at the errorgen's own parameter values (a block's to_vector() may return an equivalent but different point)
            J = np.asarray(blk.deriv_wrt_params(np.asarray(errorgen.paramvals[col:col + nb])))
            dc[row:row + size, col:col + nb] = J.reshape(size, nb)
        row += size
        col += nb
    return LindbladErrgen(B_re=np.ascontiguousarray(B.real), B_im=np.ascontiguousarray(B.imag), c=c, dc=dc)


def pack_lindblad(model, atom, dim):
    """Split the atom's members into Lindblad members (inputs of the device algebra) and host members."""
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    members, errgens, host = [], [], []
    seen = {}

    def eg_index(errorgen):
        key = id(errorgen)
        if key not in seen:
            seen[key] = len(errgens)
            errgens.append(_errgen_inputs(errorgen))
        return seen[key]

    def dense_rep(errorgen):
        return getattr(errorgen, "_rep_type", None) == "dense" or hasattr(errorgen, "combined_lindblad_term_superops")

    off = 0
    for kind, group, size in (("op", ops, d * d), ("rho", rhos, d), ("eff", effs, d)):
        for m in group:
            name = type(m).__name__
            lm = None
            if kind == "op" and name == "ComposedOp":
                f = list(m.factorops)
                if len(f) == 2 and type(f[1]).__name__ == "ExpErrorgenOp" and f[0].num_params == 0 \
                        and type(f[1].errorgen).__name__ == "LindbladErrorgen" and dense_rep(f[1].errorgen):
                    lm = LindbladMember(kind, off, np.asarray(f[0].to_dense('HilbertSchmidt'), float).reshape(d, d),
                                        eg_index(f[1].errorgen), _gp_array(m.gpindices))
            elif kind == "rho" and name == "ComposedState":
                em = m.error_map
                if type(em).__name__ == "ExpErrorgenOp" and m.state_vec.num_params == 0 \
                        and type(em.errorgen).__name__ == "LindbladErrorgen" and dense_rep(em.errorgen):
                    lm = LindbladMember(kind, off, np.asarray(m.state_vec.to_dense('HilbertSchmidt'), float).reshape(d),
                                        eg_index(em.errorgen), _gp_array(m.gpindices))
            elif kind == "eff" and name == "ComposedPOVMEffect":
                em = m.error_map
                if type(em).__name__ == "ExpErrorgenOp" and m.effect_vec.num_params == 0 \
                        and type(em.errorgen).__name__ == "LindbladErrorgen" and dense_rep(em.errorgen):
                    lm = LindbladMember(kind, off, np.asarray(m.effect_vec.to_dense('HilbertSchmidt'), float).reshape(d),
                                        eg_index(em.errorgen), _gp_array(m.gpindices))
            if lm is not None and lm.gpindices.size == errgens[lm.errgen].dc.shape[1]:
                members.append(lm)
            else:
                host.append((kind, off, m))
            off += size
    return LindbladInputs(members=members, errgens=errgens, host_members=host)


def assemble_lindblad(li, outs, model, atom, dim, param_indices=None):
    """(ModelTensors, DerivMap) of the atom from the Lindblad members' dense values / derivatives ``outs`` -- the result of
    ``engine.Context.lindblad_members(dim, li.errgens, li.members)`` -- plus the remaining members packed on the host as in
    ``pack_model`` / ``pack_derivs``.  Same output as those two functions (no parameter interposer)."""
    if getattr(model, '_param_interposer', None) is not None:
        raise ValueError("assemble_lindblad does not support parameter interposers")
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    n_w = len(ops) * d * d + len(rhos) * d + len(effs) * d
    n_model_params = int(model.num_params)
    M = np.empty(n_w, dtype=np.float64)
    rows, cols, vals = [], [], []
    for lm, (val, dval) in zip(li.members, outs):
        size = val.size
        M[lm.w_offset:lm.w_offset + size] = val
        r, c = np.nonzero(dval)
        rows.append(lm.w_offset + r); cols.append(lm.gpindices[c]); vals.append(dval[r, c])
    for kind, off, m in li.host_members:
        size = d * d if kind == "op" else d
        M[off:off + size] = np.asarray(m.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(size)
        gp = _gp_array(m.gpindices)
        if gp.size:
            dM = np.asarray(m.deriv_wrt_params(), dtype=np.float64).reshape(size, gp.size)
            r, c = np.nonzero(dM)
            rows.append(off + r); cols.append(gp[c]); vals.append(dM[r, c])
    no, nr = len(ops) * d * d, len(rhos) * d
    mt = ModelTensors(G=M[:no].reshape(len(ops), d, d).copy(), rho=M[no:no + nr].reshape(len(rhos), d).copy(),
                      E=M[no + nr:].reshape(len(effs), d).copy())
    if rows:
        rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    else:
        rows = np.zeros(0, np.int64); cols = np.zeros(0, np.int64); vals = np.zeros(0, np.float64)
    import scipy.sparse as sps
    D = sps.coo_matrix((vals, (rows, cols)), shape=(n_w, n_model_params)).tocsr()   # sums duplicates (shared generators)
    pidx = param_slice_to_array(param_indices, n_model_params)
    if not (pidx.size == n_model_params and np.array_equal(pidx, np.arange(n_model_params))):
        D = D[:, pidx]
    D = D.tocoo()
    keep = D.data != 0.0
    return mt, DerivMap(n_w=n_w, n_params=int(pidx.size), rows=D.row[keep].astype(np.int32), cols=D.col[keep].astype(np.int32),
                        vals=D.data[keep].astype(np.float64))


# ----------------------------------------------------------------------------------------------------------------------
# Factored (non-dense) gate representation: layer operations that pyGSTi builds as products of small operations embedded
# on 1-2 qubits (`ComposedOp` of `EmbeddedOp`; pygsti/modelmembers/operations/{composedop,embeddedop}.py -- the reference's
# OpCRep_Composed / OpCRep_Embedded, pygsti/evotypes/densitymx/opcreps.cpp:93-158, 242-276) are handed to the engine as
# "factor programs": per layer label a list of (small dense superoperator, target qubits).  The engine applies the factors to
# the state directly (4 or 16 multiply-adds per component instead of d = 256) and builds the dense d x d matrices it needs
# for the derivative paths ON THE DEVICE -- the host never calls to_dense on a d = 256 operation (0.01-0.2 s per label).
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class FactoredModel:
    n_qubits: int
    op_fptr: np.ndarray      # int32 [n_ops + 1]  factors of op g: [op_fptr[g], op_fptr[g+1]), applied in this order
    f_nq: np.ndarray         # int32 [n_factors]  number of target qubits (1 or 2)
    f_targets: np.ndarray    # int32 [n_factors, 4]  qubit positions in state-space order (0 = most significant), -1 padded
    f_moff: np.ndarray       # int64 [n_factors]  offset of the factor's (4^nq x 4^nq) row-major matrix in `mats`
    mats: np.ndarray         # float64
    rho: np.ndarray
    E: np.ndarray


def _flatten_factors(op, qubit_labels, out):
    """Append (targets, small dense matrix) of `op` to `out`; False if `op` is not a product of embedded <= 2-qubit ops."""
    name = type(op).__name__
    if name == "ComposedOp":
        return all(_flatten_factors(f, qubit_labels, out) for f in op.factorops)
    if name == "EmbeddedOp":
        try:
            targets = [qubit_labels.index(t) for t in op.target_labels]
        except ValueError:
            return False
        small = np.asarray(op.embedded_op.to_dense('HilbertSchmidt'), dtype=np.float64)
        k = len(targets)
        if k < 1 or k > 2 or small.shape != (4 ** k, 4 ** k) or len(set(targets)) != k:
            return False
        out.append((targets, small, op.embedded_op))
        return True
    return False


def pack_model_factored(model, atom, dim):
    """FactoredModel of the atom's layer operations, or None when some operation is not a product of embedded 1-2 qubit
    operations on a pure n-qubit space (the caller then packs dense matrices as before)."""
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    nq = int(round(np.log(d) / np.log(4)))
    if 4 ** nq != d or nq < 2 or not ops:
        return None
    try:
        labels = list(model.state_space.sole_tensor_product_block_labels)
    except Exception:
        return None
    if len(labels) != nq:
        return None
    fptr, f_nq, f_targets, f_moff, mats = [0], [], [], [], []
    off = 0
    for op in ops:
        fac = []
        if not _flatten_factors(op, labels, fac):
            return None
        for targets, small, _ in fac:
            f_nq.append(len(targets)); f_targets.append(list(targets) + [-1] * (4 - len(targets)))
            f_moff.append(off); mats.append(small.ravel()); off += small.size
        fptr.append(len(f_nq))
    rho = np.empty((len(rhos), d)); E = np.empty((len(effs), d))
    for i, r in enumerate(rhos):
        rho[i] = np.asarray(r.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(d)
    for i, e in enumerate(effs):
        E[i] = np.asarray(e.to_dense('HilbertSchmidt'), dtype=np.float64).reshape(d)
    return FactoredModel(n_qubits=nq, op_fptr=np.asarray(fptr, np.int32), f_nq=np.asarray(f_nq, np.int32),
                         f_targets=np.asarray(f_targets, np.int32).reshape(-1, 4), f_moff=np.asarray(f_moff, np.int64),
                         mats=(np.concatenate(mats) if mats else np.zeros(0)), rho=rho, E=E)


def pack_derivs_factored(model, atom, dim, fm, param_indices=None):
    """Derivative map in FACTOR space: rows index [fm.mats | rho | E] -- the entries of the small embedded operations themselves
    (``embedded_op.deriv_wrt_params()``: 16 x 12 for a 1-qubit full TP gate) instead of the d^2 entries of every dense layer
    operation -- for ``engine.Atom.set_derivs_factored`` (the Jacobian kernels of csrc/kernels_factoredj.cuh).  ``fm`` is the
    FactoredModel of the same (model, atom).  None with a parameter interposer (the caller keeps the dense map)."""
    if getattr(model, '_param_interposer', None) is not None or fm is None:
        return None
    ops, rhos, effs = _members(model, atom)
    d = int(dim)
    labels = list(model.state_space.sole_tensor_product_block_labels)
    n_mats = int(fm.mats.size)
    n_wf = n_mats + (len(rhos) + len(effs)) * d
    n_model_params = int(model.num_params)
    rows, cols, vals = [], [], []
    f = 0
    for op in ops:
        fac = []
        if not _flatten_factors(op, labels, fac):
            return None
        for targets, small, eop in fac:
            gp = _gp_array(eop.gpindices)
            if gp.size:
                dM = np.asarray(eop.deriv_wrt_params(), dtype=np.float64).reshape(small.size, gp.size)
                r, c = np.nonzero(dM)
                rows.append(int(fm.f_moff[f]) + r); cols.append(gp[c]); vals.append(dM[r, c])
            f += 1
    off = n_mats
    for group in (rhos, effs):
        for m in group:
            gp = _gp_array(m.gpindices)
            if gp.size:
                dM = np.asarray(m.deriv_wrt_params(), dtype=np.float64).reshape(d, gp.size)
                r, c = np.nonzero(dM)
                rows.append(off + r); cols.append(gp[c]); vals.append(dM[r, c])
            off += d
    if rows:
        rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
    else:
        rows = np.zeros(0, np.int64); cols = np.zeros(0, np.int64); vals = np.zeros(0, np.float64)
    import scipy.sparse as sps
    D = sps.coo_matrix((vals, (rows, cols)), shape=(n_wf, n_model_params)).tocsr()
    pidx = param_slice_to_array(param_indices, n_model_params)
    if not (pidx.size == n_model_params and np.array_equal(pidx, np.arange(n_model_params))):
        D = D[:, pidx]
    D = D.tocoo()
    keep = D.data != 0.0
    return DerivMap(n_w=n_wf, n_params=int(pidx.size), rows=D.row[keep].astype(np.int32), cols=D.col[keep].astype(np.int32),
                    vals=D.data[keep].astype(np.float64))


def factored_to_dense(fm, dim):
    """numpy restatement of what the device does with a FactoredModel (test / oracle helper): dense G [n_ops, d, d]."""
    d = int(dim); nq = fm.n_qubits
    n_ops = fm.op_fptr.shape[0] - 1
    G = np.empty((n_ops, d, d))
    idx = np.arange(d)
    for g in range(n_ops):
        M = np.eye(d)
        for f in range(fm.op_fptr[g], fm.op_fptr[g + 1]):
            k = int(fm.f_nq[f]); ds = 4 ** k
            small = fm.mats[fm.f_moff[f]:fm.f_moff[f] + ds * ds].reshape(ds, ds)
            shifts = [2 * (nq - 1 - int(q)) for q in fm.f_targets[f, :k]]
            tmask = 0
            for s in shifts:
                tmask |= 3 << s
            t = np.zeros(d, dtype=np.int64)
            for s in shifts:
                t = (t << 2) | ((idx >> s) & 3)
            rest = idx & ~tmask
            F = np.zeros((d, d))
            for tp in range(ds):
                col = rest.copy()
                for a, s in enumerate(shifts):
                    col |= ((tp >> (2 * (k - 1 - a))) & 3) << s
                F[idx, col] += small[t, tp]
            M = F @ M                                  # factors act in order: the first factor is applied first
        G[g] = M
    return G
