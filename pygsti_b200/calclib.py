"""
``pygsti_b200.calclib`` -- a drop-in *calclib* module for pyGSTi's Map forward simulator.

pyGSTi's ``SimpleMapForwardSimulator._set_evotype`` imports
``pygsti.forwardsims.mapforwardsim_calc_<evotype>`` and calls three functions of it
(reference: pygsti/forwardsims/mapforwardsim.py:93-102, :372-391).  This module exports the same three
functions with the same signatures and argument meaning as the Cython original
(pygsti/forwardsims/mapforwardsim_calc_densitymx.pyx:40, :149, :290), implemented on the B200 engine:

    propagate_staterep(staterep, operationreps)
    mapfill_probs_atom(fwdsim, array_to_fill, dest_indices, layout_atom, resource_alloc)
    mapfill_dprobs_atom(fwdsim, array_to_fill, dest_indices, dest_param_indices, layout_atom,
                        param_indices, resource_alloc, eps)

Differences, by design:
  * the Jacobian is ANALYTIC by default (parity target: the reference's MatrixForwardSimulator, <= 1e-10 abs);
    ``fwdsim.derivative_mode = 'fd'`` reproduces the reference's forward differences (pyx:349-378)
    with the given ``eps`` on the device;
  * layout tables are converted once per atom and cached on the device (the reference converts on every
    call, pyx:163-181); the model tensors are re-read on every call, as the reference does;
  * there is NO CPU fallback: without the CUDA library / a GPU these functions raise.
"""
import os
import weakref

import numpy as np

from . import engine, packing

_CONTEXTS = {}
_ATOM_CACHE = weakref.WeakKeyDictionary()   # layout atom -> {device: uploaded engine atom}; never pickled

# members whose derivative w.r.t. their own parameters does not depend on the parameter values
# (the dense array is linear in the parameters), so the derivative map can be cached per layout atom
_LINEAR_MEMBERS = frozenset([
    "FullArbitraryOp", "FullTPOp", "StaticArbitraryOp", "StaticUnitaryOp", "StaticCliffordOp", "StaticStandardOp",
    "FullState", "TPState", "StaticState", "FullPOVMEffect", "StaticPOVMEffect", "ComplementPOVMEffect",
    "ConjugatedStatePOVMEffect"])


def get_context(device=None):
    """Process-wide engine context for a GPU (default: LOCAL_RANK's GPU, else 0)."""
    if device is None:
        device = int(os.environ.get("B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        n = engine.device_count()
        if n > 0:
            device %= n
    if device not in _CONTEXTS:
        _CONTEXTS[device] = engine.Context(device)
    return _CONTEXTS[device]


def _device_for(fwdsim, layout_atom):
    dev = getattr(layout_atom, "_b200_device", None)   # single-process multi-GPU: atoms are tagged
    if dev is not None:
        return dev
    return getattr(fwdsim, "_b200_device", None)


def _engine_atom(fwdsim, layout_atom):
    """Upload (once) and return the device-resident copy of a layout atom."""
    ctx = get_context(_device_for(fwdsim, layout_atom))
    cache = _ATOM_CACHE.get(layout_atom)
    if cache is None:
        cache = _ATOM_CACHE[layout_atom] = {}
    ent = cache.get(ctx.device)
    if ent is None:
        dim = fwdsim.model.dim
        tables = packing.pack_atom(layout_atom, dim)
        ent = {"atom": ctx.upload_atom(tables), "tables": tables, "deriv_key": None}
        cache[ctx.device] = ent
    return ctx, ent


_REGISTERED = {}


def _pin_destination(arr):
    """Page-lock the caller's destination buffer once (cudaHostRegister, cached by base address) so that results
    are DMA'd straight into it at PCIe rate.  pyGSTi's objective functions re-use their `probs` / `jac` arrays
    across optimizer iterations (objectivefns.py:4548, 4609), so the one-time pinning cost is amortised.
    Small arrays and foreign memory (e.g. MPI shared memory that refuses registration) are left alone."""
    if os.environ.get("B200_NO_HOST_REGISTER") or arr.nbytes < (1 << 25):
        return
    base = arr
    while isinstance(getattr(base, "base", None), np.ndarray):
        base = base.base
    if not isinstance(base, np.ndarray) or not base.flags.c_contiguous:
        return
    key = base.ctypes.data
    if _REGISTERED.get(key, 0) >= base.nbytes:
        return
    try:
        from . import _lib
        import ctypes as C
        lib = _lib.load()
        rc = lib.b200_host_register(C.c_void_p(key), int(base.nbytes))
        _REGISTERED[key] = base.nbytes if rc == 0 else (1 << 62)     # do not retry a buffer that cannot be pinned
        if rc != 0:
            import warnings
            warnings.warn("B200 engine: cudaHostRegister refused the destination array; device-to-host copies of this array "
                          "run through pageable memory (~3x slower)", RuntimeWarning)
        if rc == 0:   # un-pin when the owner array is garbage collected (before its pages go back to the allocator)
            def _unpin(k=key):
                _REGISTERED.pop(k, None)
                lib.b200_host_unregister(C.c_void_p(k))
            weakref.finalize(base, _unpin)
        if len(_REGISTERED) > 64:                                    # keep the table small: forget the oldest entries
            for k in list(_REGISTERED)[:16]:
                if _REGISTERED[k] < (1 << 62):
                    _lib.load().b200_host_unregister(C.c_void_p(k))
                del _REGISTERED[k]
    except Exception as e:                                           # pageable destination: correct, but ~3x slower D2H
        import warnings
        warnings.warn("B200 engine: could not page-lock the destination array (%s: %s); the copy will be staged" %
                      (type(e).__name__, e), RuntimeWarning)


def _to_index_array(idx, n):
    if idx is None:
        return None
    if isinstance(idx, slice):
        start, stop, step = idx.indices(n)
        return np.arange(start, stop, step, dtype=np.int64)
    return np.asarray(idx, dtype=np.int64)


def _is_full_range(idx, n):
    if idx is None:
        return True
    if isinstance(idx, slice):
        return idx.indices(n) == (0, n, 1)
    idx = np.asarray(idx)
    return idx.shape == (n,) and np.array_equal(idx, np.arange(n))


def _contiguous_block(idx, n):
    """(start, stop) if idx selects a contiguous ascending block of range(n), else None."""
    if idx is None:
        return (0, n)
    if isinstance(idx, slice):
        start, stop, step = idx.indices(n)
        return (start, max(start, stop)) if step == 1 else None
    idx = np.asarray(idx)
    if idx.size == 0:
        return (0, 0)
    if np.array_equal(idx, np.arange(idx[0], idx[0] + idx.size)):
        return (int(idx[0]), int(idx[0]) + idx.size)
    return None


# --------------------------------------------------------------------------------------------------
def propagate_staterep(staterep, operationreps):
    """Same contract as pyx:40-47 (used only by the single-circuit reference path); host-side rep objects
    are the reference's own, so this simply chains their ``acton``."""
    ret = staterep
    for oprep in operationreps:
        ret = oprep.acton(ret)
    return ret


def mapfill_probs_atom(fwdsim, array_to_fill, dest_indices, layout_atom, resource_alloc):
    """array_to_fill[dest_indices] = outcome probabilities of every element of ``layout_atom``
    (replaces pyx:149-190 + dm_mapfill_probs pyx:194-287)."""
    shared_mem_leader = resource_alloc.is_host_leader if (resource_alloc is not None) else True
    if not shared_mem_leader:
        return  # same guard as pyx:158,183: only the host leader writes shared memory
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    atom = ent["atom"]
    _upload_model(fwdsim, layout_atom, ent)
    nE = layout_atom.num_elements
    blk = _contiguous_block(dest_indices, array_to_fill.shape[0])
    if blk is not None and blk[1] - blk[0] == nE and array_to_fill.dtype == np.float64:
        atom.fill_probs(array_to_fill[blk[0]:blk[1]])
    else:
        tmp = np.empty(nE)
        atom.fill_probs(tmp)
        array_to_fill[_to_index_array(dest_indices, array_to_fill.shape[0])] = tmp


def _bound_to(ent, model):
    ref = ent.get("bound")
    return bool(ref) and ref() is model


def _upload_model(fwdsim, layout_atom, ent):
    """Bring the device copy of the model tensors up to date with ``fwdsim.model``.
    Default: pack the members' dense matrices on the host and upload them (replaces the rep lookups of pyx:164-167).
    With ``device_model_update`` and a parameter binding (all members affine, full derivative map resident): upload only
    the parameter vector; the engine evaluates M = M_const + D theta on the device."""
    model = fwdsim.model
    atom = ent["atom"]
    ent["fm"] = None
    if getattr(fwdsim, "device_model_update", False) and _bound_to(ent, model):
        atom.set_params(model.to_vector())
        return
    if model.dim >= 64 and getattr(fwdsim, "factored_reps", True):
        # dim > 64: pyGSTi's own reps are Composed / Embedded (evotype.py:97); hand the factor programs to the engine instead of
        # densifying every layer label on the host (packing.pack_model_factored returns None for anything else)
        fm = packing.pack_model_factored(model, layout_atom, model.dim)
        if fm is not None:
            atom.set_model_factored(fm)
            ent["fm"] = fm
            return
    atom.set_model(packing.pack_model(model, layout_atom, model.dim))


def _maybe_bind(fwdsim, layout_atom, ent, pidx):
    """After a host upload of the model AND of the full derivative map: bind the parameter vector so that later fills can
    update the model on the device."""
    model = fwdsim.model
    if not getattr(fwdsim, "device_model_update", False) or _bound_to(ent, model):
        return
    if pidx.size != model.num_params or not np.array_equal(pidx, np.arange(model.num_params)):
        return
    if not all_members_linear(fwdsim, layout_atom):
        return
    ent["atom"].bind_params(model.to_vector())
    ent["bound"] = weakref.ref(model)       # M_const belongs to THIS model object (its static members, its structure)


def _lindblad_model_and_derivs(fwdsim, layout_atom, ent, param_indices, ctx):
    """``device_lindblad``: for models with Lindblad-parameterised members (CPTPLND, H+S, GLND) the dense algebra of the model
    update -- error generators, exponentials, Frechet derivatives, composition with the static parts -- runs on the device
    (``b200_lindblad_members``) instead of in ``to_dense`` / ``deriv_wrt_params`` of every member; coefficients and their Jacobian
    stay on the host.  Returns False (nothing uploaded) when the atom has no such member or the model has a parameter interposer."""
    model = fwdsim.model
    if getattr(model, "_param_interposer", None) is not None:
        return False
    li = packing.pack_lindblad(model, layout_atom, model.dim)
    if not li.members:
        return False
    outs = ctx.lindblad_members(model.dim, li.errgens, li.members)
    mt, D = packing.assemble_lindblad(li, outs, model, layout_atom, model.dim, param_indices)
    atom = ent["atom"]
    atom.set_model(mt)
    atom.set_derivs(D)
    ent["deriv_key"] = None          # the map depends on the parameters: never cached
    ent["bound"] = None
    return True


def all_members_linear(fwdsim, layout_atom):
    model = fwdsim.model
    ops, rhos, effs = packing._members(model, layout_atom)
    return all(type(m).__name__ in _LINEAR_MEMBERS for m in ops + rhos + effs) \
        and getattr(model, "_param_interposer", None) is None


def mapfill_hprobs_atom_linear(fwdsim, array_to_fill, dest_param_indices1, dest_param_indices2, layout_atom,
                               param_indices1, param_indices2, resource_alloc):
    """array_to_fill[:, dest1, dest2] = d2 p / d theta_{p1} d theta_{p2} for members linear in their parameters
    (replaces MapForwardSimulator._mapfill_hprobs_atom, mapforwardsim.py:394-438)."""
    shared_mem_leader = resource_alloc.is_host_leader if (resource_alloc is not None) else True
    model = fwdsim.model
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    atom = ent["atom"]
    _upload_model(fwdsim, layout_atom, ent)
    _maybe_bind(fwdsim, layout_atom, ent, _deriv_map(fwdsim, layout_atom, ent, None))          # full derivative map; blocks select its columns
    if not shared_mem_leader:
        return
    p1 = packing.param_slice_to_array(param_indices1, model.num_params)
    p2 = packing.param_slice_to_array(param_indices2, model.num_params)
    nE = layout_atom.num_elements
    if nE == 0 or p1.size == 0 or p2.size == 0:
        return
    tmp = np.empty((nE, p1.size, p2.size))
    atom.fill_hprobs_linear(p1, p2, tmp)
    d1 = np.arange(p1.size) if dest_param_indices1 is None else _to_index_array(dest_param_indices1, array_to_fill.shape[1])
    d2 = np.arange(p2.size) if dest_param_indices2 is None else _to_index_array(dest_param_indices2, array_to_fill.shape[2])
    array_to_fill[np.ix_(np.arange(nE), d1, d2)] = tmp


def mapfill_hprobs_atom_analytic(fwdsim, array_to_fill, dest_param_indices1, dest_param_indices2, layout_atom,
                                 param_indices1, param_indices2, resource_alloc):
    """Fully analytic Hessian rectangle for arbitrary members (CPTPLND, H+S, ...): the members' own
    ``hessian_wrt_params`` supply the second-derivative term (as in MatrixForwardSimulator._hprobs_from_rho_e,
    matrixforwardsim.py:1141-1287).  Returns False -- nothing written -- when a member cannot provide it, so that the
    caller can fall back to the reference's finite-difference driver (mapforwardsim.py:394-438)."""
    model = fwdsim.model
    p1 = packing.param_slice_to_array(param_indices1, model.num_params)
    p2 = packing.param_slice_to_array(param_indices2, model.num_params)
    try:
        hess = packing.pack_hessians(model, layout_atom, model.dim, p1, p2)
    except (NotImplementedError, ValueError, AssertionError):
        return False
    shared_mem_leader = resource_alloc.is_host_leader if (resource_alloc is not None) else True
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    atom = ent["atom"]
    _upload_model(fwdsim, layout_atom, ent)
    _maybe_bind(fwdsim, layout_atom, ent, _deriv_map(fwdsim, layout_atom, ent, None))
    if not shared_mem_leader:
        return True
    nE = layout_atom.num_elements
    if nE == 0 or p1.size == 0 or p2.size == 0:
        return True
    tmp = np.empty((nE, p1.size, p2.size))
    atom.fill_hprobs(p1, p2, tmp, hess)
    d1 = np.arange(p1.size) if dest_param_indices1 is None else _to_index_array(dest_param_indices1, array_to_fill.shape[1])
    d2 = np.arange(p2.size) if dest_param_indices2 is None else _to_index_array(dest_param_indices2, array_to_fill.shape[2])
    array_to_fill[np.ix_(np.arange(nE), d1, d2)] = tmp
    return True


def atom_hessian_block(fwdsim, layout_atom, param_indices1, param_indices2, w_h, w_d):
    """One (n1 x n2) block of the MLE Hessian of this atom's elements, reduced on the device
    (`_hessian_from_block`, objectivefns.py:4914-4990).  Returns None when a member cannot provide analytic second
    derivatives (the caller then uses the reference path)."""
    model = fwdsim.model
    p1 = packing.param_slice_to_array(param_indices1, model.num_params)
    p2 = packing.param_slice_to_array(param_indices2, model.num_params)
    hess = None
    if not all_members_linear(fwdsim, layout_atom):
        try:
            hess = packing.pack_hessians(model, layout_atom, model.dim, p1, p2)
        except (NotImplementedError, ValueError, AssertionError):
            return None
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    atom = ent["atom"]
    _upload_model(fwdsim, layout_atom, ent)
    _maybe_bind(fwdsim, layout_atom, ent, _deriv_map(fwdsim, layout_atom, ent, None))
    if layout_atom.num_elements == 0 or p1.size == 0 or p2.size == 0:
        return np.zeros((p1.size, p2.size))
    return atom.hessian_block(p1, p2, w_h, w_d, hess)


def _deriv_map(fwdsim, layout_atom, ent, param_indices):
    """Upload the derivative map for this parameter block (cached when every member is linear in its
    parameters, in which case D does not depend on the current parameter vector)."""
    model = fwdsim.model
    pidx = packing.param_slice_to_array(param_indices, model.num_params)
    ops, rhos, effs = packing._members(model, layout_atom)
    linear = all(type(m).__name__ in _LINEAR_MEMBERS for m in ops + rhos + effs) \
        and getattr(model, "_param_interposer", None) is None
    # D is constant only for THIS model's members and parameter allocation: the key names them (a layout re-used with another
    # or re-parameterised model of equal size must not hit a stale map)
    key = (pidx.tobytes(), model.num_params, id(model),
           tuple((id(m), getattr(m.gpindices, "start", None), getattr(m.gpindices, "stop", None)) if isinstance(m.gpindices, slice)
                 else (id(m), bytes(np.asarray(m.gpindices).tobytes())) for m in ops + rhos + effs)) if linear else None
    if key is not None and ent["deriv_key"] == key:
        return pidx
    D = packing.pack_derivs(model, layout_atom, model.dim, pidx)
    ent["atom"].set_derivs(D)
    if ent.get("fm") is not None and model.dim in (64, 256):
        # the model went up as factor programs: hand over the derivative map in factor space as well -- the Jacobian is then
        # evaluated factor by factor (csrc/kernels_factoredj.cuh) instead of through dense d x d sweeps
        Df = packing.pack_derivs_factored(model, layout_atom, model.dim, ent["fm"], pidx)
        if Df is not None:
            ent["atom"].set_derivs_factored(Df)
    ent["deriv_key"] = key
    ent["bound"] = None             # a parameter binding refers to the map it was made with
    return pidx


def prepare_dprobs_atom(fwdsim, layout_atom, param_indices):
    """Host half of a Jacobian fill: bring the device copy of the model tensors and of the derivative map for this parameter block
    up to date (host packing + small uploads).  Returns (engine atom, parameter index array).  Runs on the calling thread only:
    it walks pyGSTi model objects."""
    model = fwdsim.model
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    if getattr(fwdsim, "device_lindblad", True) and _lindblad_model_and_derivs(fwdsim, layout_atom, ent, param_indices, ctx):
        pidx = packing.param_slice_to_array(param_indices, model.num_params)
    else:
        _upload_model(fwdsim, layout_atom, ent)
        pidx = _deriv_map(fwdsim, layout_atom, ent, param_indices)
        _maybe_bind(fwdsim, layout_atom, ent, pidx)
    return ent["atom"], pidx


def run_dprobs_atom(fwdsim, atom, pidx, array_to_fill, dest_indices, dest_param_indices, layout_atom, eps, row_scale=None,
                    pr_array_to_fill=None):
    """Device half of a Jacobian fill: kernels + device-to-host copy into ``array_to_fill[dest_indices, dest_param_indices]``
    (and, if given, the probabilities into ``pr_array_to_fill``, a view of exactly this atom's elements).  Touches only the
    engine (the C calls release the GIL): atoms that live on different GPUs can run from different threads."""
    nE = layout_atom.num_elements
    nP = int(pidx.size)
    mode = getattr(fwdsim, "derivative_mode", "analytic")
    pr = pr_array_to_fill if (pr_array_to_fill is not None and pr_array_to_fill.dtype == np.float64
                              and pr_array_to_fill.ndim == 1 and pr_array_to_fill.shape[0] == nE) else None
    if mode == "fd":
        if row_scale is not None:
            raise ValueError("row_scale is only supported with derivative_mode='analytic'")
        fill = (lambda out: atom.fill_dprobs_fd(out, eps=eps, probs_out=pr))
    else:
        fill = (lambda out: atom.fill_dprobs(out, probs_out=pr, row_scale=row_scale))

    rblk = _contiguous_block(dest_indices, array_to_fill.shape[0])
    if dest_param_indices is None:
        cblk = (0, nP)
    else:
        cblk = _contiguous_block(dest_param_indices, array_to_fill.shape[1])
    direct = (rblk is not None and cblk is not None and rblk[1] - rblk[0] == nE and cblk[1] - cblk[0] == nP
              and array_to_fill.dtype == np.float64
              and (array_to_fill.shape[1] <= 1 or array_to_fill.strides[1] == 8))
    if nE == 0:
        return
    if nP == 0:
        if pr is not None:
            atom.fill_probs(pr)
        return
    if direct:
        _pin_destination(array_to_fill)
        fill(array_to_fill[rblk[0]:rblk[1], cblk[0]:cblk[1]])
    else:
        tmp = np.empty((nE, nP))
        fill(tmp)
        r = _to_index_array(dest_indices, array_to_fill.shape[0])
        c = np.arange(nP) if dest_param_indices is None else _to_index_array(dest_param_indices, array_to_fill.shape[1])
        array_to_fill[np.ix_(r, c)] = tmp
    if pr_array_to_fill is not None and pr is None:          # unusual destination: separate fill
        tmp = np.empty(nE); atom.fill_probs(tmp); pr_array_to_fill[...] = tmp


def mapfill_dprobs_atom(fwdsim, array_to_fill, dest_indices, dest_param_indices, layout_atom, param_indices,
                        resource_alloc, eps, row_scale=None):
    """array_to_fill[dest_indices, dest_param_indices] = d(probabilities)/d(params[param_indices])
    (replaces pyx:290-383).  ``eps`` is used only in 'fd' mode.  ``row_scale`` (extension, length = number of
    elements of the atom): every Jacobian row is multiplied by its entry on the device (objective-function fill)."""
    shared_mem_leader = resource_alloc.is_host_leader if (resource_alloc is not None) else True
    atom, pidx = prepare_dprobs_atom(fwdsim, layout_atom, param_indices)
    if not shared_mem_leader:
        return
    run_dprobs_atom(fwdsim, atom, pidx, array_to_fill, dest_indices, dest_param_indices, layout_atom, eps, row_scale)


def atom_jtj(fwdsim, layout_atom, row_scale=None, f=None):
    """(J^T J, J^T f) of one atom with J = diag(row_scale) . dprobs kept on the device."""
    model = fwdsim.model
    ctx, ent = _engine_atom(fwdsim, layout_atom)
    atom = ent["atom"]
    _upload_model(fwdsim, layout_atom, ent)
    _maybe_bind(fwdsim, layout_atom, ent, _deriv_map(fwdsim, layout_atom, ent, None))
    return atom.jtj(row_scale, f)


# The reference calclib also exports six time-dependent functions (pyx:400-586).  They are OUT OF SCOPE for
# this engine; calling them fails loudly instead of silently running on the CPU.
def _td_unsupported(*args, **kwargs):
    raise NotImplementedError("time-dependent mapfill_TD* functions are not implemented by the B200 engine; "
                              "use pyGSTi's MapForwardSimulator for time-dependent objective terms")


mapfill_TDchi2_terms = mapfill_TDloglpp_terms = mapfill_TDterms = _td_unsupported
mapfill_TDdchi2_terms = mapfill_TDdloglpp_terms = mapfill_TDdterms = _td_unsupported
