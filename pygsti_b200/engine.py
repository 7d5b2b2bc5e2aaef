"""Thin object layer over the C ABI: Context (one per GPU) and Atom (device-resident layout atom).

Only numpy + ctypes; arrays are handed to the library as raw pointers + strides.  All compute happens in
libb200fwdsim.so on the GPU; a missing library or device raises (no fallback).
"""
import ctypes as C
import numpy as np

from . import _lib
from .packing import AtomTables, DerivMap, ModelTensors


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    lib = _lib.load()
    n = C.c_int(0)
    rc = lib.b200_device_count(C.byref(n))
    if rc != 0:
        return 0
    return n.value


class Context:
    """One engine context per GPU (mirrors one MPI rank's ResourceAllocation in the reference).  A context and its atoms share
    scratch buffers: calls on one context must not overlap (use one context per thread / per GPU)."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        h = C.c_void_p(0)
        _lib.check(self._lib.b200_ctx_create(int(device), C.c_void_p(int(stream) if stream else 0), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_ctx_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _lib.check(self._lib.b200_ctx_sync(self._h))

    @property
    def launch_count(self):
        n = C.c_int64(0)
        _lib.check(self._lib.b200_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def set_jtj_mode(self, mode):
        """-1 automatic (default), 0 FP64 DMMA SYRK, 8 / 7 tcgen05 Ozaki with 8 / 7 digits (b200_ctx_set_jtj_mode)."""
        _lib.check(self._lib.b200_ctx_set_jtj_mode(self._h, int(mode)))

    def phase_timing(self, on=True):
        """Bracket the phases of the d = 16 Jacobian path with CUDA events on the launching stream (bench.py roofline)."""
        _lib.check(self._lib.b200_ctx_phase_timing(self._h, 1 if on else 0))

    def phase_ms(self):
        """(prepare_ms, chains_ms, accumulate_ms) summed over the calls since the last read, and the number of calls."""
        ms = (C.c_double * 3)()
        n = C.c_int64(0)
        _lib.check(self._lib.b200_ctx_phase_ms(self._h, ms, C.byref(n)))
        return (ms[0], ms[1], ms[2]), n.value

    def lindblad_members(self, d, errgens, members):
        """Dense Lindblad members and their parameter derivatives on the device (``b200_lindblad_members``).

        errgens : sequence of objects with ``B_re, B_im`` [n_coeff, d, d], ``c`` complex [n_coeff], ``dc`` complex [n_coeff, n_par]
        members : sequence of objects with ``kind`` ('op' | 'rho' | 'eff'), ``errgen`` (index) and ``static`` ([d, d] or [d])
        (``packing.LindbladErrgen`` / ``packing.LindbladMember``).  Returns ``[(value, dvalue [size, n_par]), ...]`` per member."""
        d = int(d)
        ncoeff = np.array([e.c.size for e in errgens], np.int32)
        npar = np.array([e.dc.shape[1] for e in errgens], np.int32)
        cat = lambda xs: np.ascontiguousarray(np.concatenate([np.ravel(x) for x in xs]) if len(xs) else np.zeros(0), np.float64)
        B_re, B_im = cat([e.B_re for e in errgens]), cat([e.B_im for e in errgens])
        c_re, c_im = cat([e.c.real for e in errgens]), cat([e.c.imag for e in errgens])
        dc_re, dc_im = cat([e.dc.real for e in errgens]), cat([e.dc.imag for e in errgens])
        kinds = {"op": 0, "rho": 1, "eff": 2}
        m_kind = np.array([kinds[m.kind] for m in members], np.int32)
        m_eg = np.array([m.errgen for m in members], np.int32)
        stat = cat([m.static for m in members])
        sizes = [d * d if k == 0 else d for k in m_kind]
        val = np.empty(int(sum(sizes)))
        dsz = [s * int(npar[g]) for s, g in zip(sizes, m_eg)]
        dval = np.empty(int(sum(dsz)))
        _lib.check(self._lib.b200_lindblad_members(self._h, d, len(errgens), _ptr(ncoeff), _ptr(npar), _ptr(B_re), _ptr(B_im),
                                                   _ptr(c_re), _ptr(c_im), _ptr(dc_re), _ptr(dc_im), len(members),
                                                   _ptr(m_kind), _ptr(m_eg), _ptr(stat), _ptr(val), _ptr(dval)))
        out, vo, do = [], 0, 0
        for s, g, ds in zip(sizes, m_eg, dsz):
            out.append((val[vo:vo + s].copy(), dval[do:do + ds].reshape(s, int(npar[g])).copy()))
            vo += s; do += ds
        return out

    def upload_atom(self, t: AtomTables):
        return Atom(self, t)

    # ---- peer memory (one process per GPU): arrays that the kernels of OTHER ranks store into over NVLink ----------------
    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte CUDA IPC handle) of a new device buffer that peers can map (``b200_peer_alloc``)."""
        p = C.c_void_p(0)
        h = (C.c_ubyte * 64)()
        _lib.check(self._lib.b200_peer_alloc(self._h, int(nbytes), C.byref(p), C.cast(h, C.c_void_p)))
        return int(p.value), bytes(h)

    def peer_open(self, handle):
        """Map a peer's buffer from its IPC handle; returns the device pointer valid in THIS process."""
        p = C.c_void_p(0)
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        _lib.check(self._lib.b200_peer_open(self._h, C.cast(h, C.c_void_p), C.byref(p)))
        return int(p.value)

    def peer_close(self, ptr):
        _lib.check(self._lib.b200_peer_close(self._h, C.c_void_p(int(ptr))))

    def peer_free(self, ptr):
        _lib.check(self._lib.b200_peer_free(self._h, C.c_void_p(int(ptr))))


class Atom:
    def __init__(self, ctx: Context, t: AtomTables):
        self.ctx = ctx
        self._lib = ctx._lib
        self.dim = t.dim
        self.n_elements = t.n_elements
        self.n_ops, self.n_rho, self.n_eff = t.n_ops, t.n_rho, t.n_eff
        self.n_w = t.n_ops * t.dim * t.dim + (t.n_rho + t.n_eff) * t.dim
        self.n_params = None
        arrs = [_i32(x) for x in (t.row_ptr, t.row_ops, t.row_istart, t.row_prep, t.row_icache,
                                  t.out_ptr, t.out_eff, t.out_el)]
        h = C.c_void_p(0)
        _lib.check(self._lib.b200_atom_upload(
            ctx._h, t.dim, t.n_ops, t.n_rho, t.n_eff, t.n_rows,
            _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]), _ptr(arrs[4]),
            int(t.cache_size), _ptr(arrs[5]), _ptr(arrs[6]), _ptr(arrs[7]), int(t.n_elements), C.byref(h)))
        self._h = h

    def free(self):
        if getattr(self, "_h", None) is not None and self._h.value and self.ctx._h.value:
            self._lib.b200_atom_free(self.ctx._h, self._h)
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def info(self):
        buf = (C.c_int64 * 8)()
        _lib.check(self._lib.b200_atom_info(self._h, buf))
        keys = ("n_rows", "n_elements", "n_prop_table", "n_prop_expanded", "max_depth", "n_w", "n_params",
                "fused_path")
        return dict(zip(keys, list(buf)))

    # ---- uploads -------------------------------------------------------------------------------
    def set_model(self, G, rho=None, E=None):
        if isinstance(G, ModelTensors):
            G, rho, E = G.G, G.rho, G.E
        d = self.dim
        G = _f64(G).reshape(self.n_ops, d, d)
        rho = _f64(rho).reshape(self.n_rho, d)
        E = _f64(E).reshape(self.n_eff, d)
        _lib.check(self._lib.b200_atom_set_model(self.ctx._h, self._h, _ptr(G), _ptr(rho), _ptr(E)))

    def set_model_factored(self, fm):
        """Upload a ``packing.FactoredModel`` (gates as products of small operations embedded on 1-2 qubits): probabilities then
        apply the factors directly and the dense matrices of the derivative paths are built on the device."""
        d = self.dim
        fptr, fnq = _i32(fm.op_fptr), _i32(fm.f_nq)
        ftg = _i32(np.asarray(fm.f_targets).reshape(-1, 4))
        moff = np.ascontiguousarray(fm.f_moff, dtype=np.int64)
        mats = _f64(fm.mats)
        rho = _f64(fm.rho).reshape(self.n_rho, d); E = _f64(fm.E).reshape(self.n_eff, d)
        if fptr.shape[0] != self.n_ops + 1:
            raise ValueError("op_fptr must have n_ops + 1 entries")
        _lib.check(self._lib.b200_atom_set_model_factored(self.ctx._h, self._h, int(fnq.shape[0]), _ptr(fptr), _ptr(fnq), _ptr(ftg),
                                                          _ptr(moff), _ptr(mats), int(mats.shape[0]), _ptr(rho), _ptr(E)))

    def set_derivs(self, D: DerivMap):
        rows, cols, vals = _i32(D.rows), _i32(D.cols), _f64(D.vals)
        _lib.check(self._lib.b200_atom_set_derivs(self.ctx._h, self._h, int(D.n_w), int(D.n_params),
                                                  int(rows.shape[0]), _ptr(rows), _ptr(cols), _ptr(vals)))
        self.n_params = int(D.n_params)

    def set_derivs_factored(self, Df: DerivMap):
        """Derivative map in factor space (``packing.pack_derivs_factored``) of the atom's factored model: the Jacobian of
        d = 64 / 256 atoms is then evaluated from the factor programs themselves.  Call after ``set_model_factored`` and,
        when a dense map of the same parameter block is also wanted (hprobs, FD mode), after ``set_derivs``."""
        rows, cols, vals = _i32(Df.rows), _i32(Df.cols), _f64(Df.vals)
        _lib.check(self._lib.b200_atom_set_derivs_factored(self.ctx._h, self._h, int(Df.n_w), int(Df.n_params),
                                                           int(rows.shape[0]), _ptr(rows), _ptr(cols), _ptr(vals)))
        self.n_params = int(Df.n_params)

    # ---- on-device model update for members affine in their parameters -------
    def bind_params(self, theta0):
        """Fix M_const = M - D theta0 from the model tensors on the device and the parameter vector they came from
        (needs ``set_model`` and ``set_derivs`` over ALL parameters)."""
        th = _f64(theta0).ravel()
        _lib.check(self._lib.b200_atom_bind_params(self.ctx._h, self._h, int(th.shape[0]), _ptr(th)))

    def set_params(self, theta):
        """M = M_const + D theta on the device: replaces model.from_vector + to_dense + set_model for affine members."""
        th = _f64(theta).ravel()
        _lib.check(self._lib.b200_atom_set_params(self.ctx._h, self._h, int(th.shape[0]), _ptr(th)))

    def set_params_dev(self, d_theta_ptr, n_params):
        _lib.check(self._lib.b200_atom_set_params_dev(self.ctx._h, self._h, int(n_params), C.c_void_p(int(d_theta_ptr))))

    def get_model(self):
        """The W-space vector M = [G | rho | E] currently on the device."""
        n_w = self.n_ops * self.dim * self.dim + (self.n_rho + self.n_eff) * self.dim
        out = np.empty(n_w)
        _lib.check(self._lib.b200_atom_get_model(self.ctx._h, self._h, int(n_w), _ptr(out)))
        return out

    # ---- host-buffer fills ---------------------------------------------------------------------
    @staticmethod
    def _vec_stride(a, n):
        if a.dtype != np.float64 or a.ndim != 1 or a.shape[0] != n:
            raise ValueError("expected float64 vector of length %d" % n)
        if n > 1 and (a.strides[0] % 8 or a.strides[0] <= 0):
            raise ValueError("unsupported stride")
        return (a.strides[0] // 8) if n > 1 else 1

    def fill_probs(self, out):
        st = self._vec_stride(out, self.n_elements)
        _lib.check(self._lib.b200_fill_probs(self.ctx._h, self._h, _ptr(out), st))
        return out

    def _mat_stride(self, a):
        if a.dtype != np.float64 or a.ndim != 2 or a.shape != (self.n_elements, self.n_params):
            raise ValueError("expected float64 array of shape (%d, %d), got %s" %
                             (self.n_elements, self.n_params, a.shape))
        if a.shape[1] > 1 and a.strides[1] != 8:
            raise ValueError("Jacobian destination must be contiguous along the parameter axis")
        if a.shape[0] > 1 and (a.strides[0] % 8 or a.strides[0] < 8 * a.shape[1]):
            raise ValueError("unsupported row stride")
        return (a.strides[0] // 8) if a.shape[0] > 1 else max(a.shape[1], 1)

    def fill_dprobs(self, out, probs_out=None, row_scale=None):
        """Jacobian into `out`; with `row_scale` (length n_elements) row el is multiplied by row_scale[el] on the
        device (the objective functions' `dprobs *= dg[:, None]`, objectivefns.py:4609-4616, fused)."""
        rs = self._mat_stride(out)
        ps = self._vec_stride(probs_out, self.n_elements) if probs_out is not None else 1
        if row_scale is None:
            _lib.check(self._lib.b200_fill_dprobs(self.ctx._h, self._h, _ptr(out), rs, _ptr(probs_out), ps))
        else:
            sc = _f64(row_scale).reshape(self.n_elements)
            _lib.check(self._lib.b200_fill_dprobs_scaled(self.ctx._h, self._h, _ptr(sc), _ptr(out), rs, _ptr(probs_out), ps))
        return out

    def jtj(self, row_scale=None, f=None):
        """(J^T J, J^T f) with J = diag(row_scale) . dprobs kept on the device (distlayout.py:1220-1359 fill_jtj/fill_jtf)."""
        n = self.n_params
        sc = _f64(row_scale).reshape(self.n_elements) if row_scale is not None else None
        fv = _f64(f).reshape(self.n_elements) if f is not None else None
        JTJ = np.empty((n, n))
        JTf = np.empty(n) if fv is not None else None
        _lib.check(self._lib.b200_jtj(self.ctx._h, self._h, _ptr(sc), _ptr(fv), _ptr(JTJ), _ptr(JTf)))
        return JTJ, JTf

    def jtj_dev(self, d_jtj, d_row_scale=0, d_f=0, d_jtf=0):
        """Device-pointer variant of :meth:`jtj` (asynchronous on the ctx stream): partial J^T J [n_params^2] / J^T f of
        this atom's element shard, ready for an all-reduce over ranks (``pygsti_b200.dist.allreduce_jtj``)."""
        _lib.check(self._lib.b200_jtj_dev(self.ctx._h, self._h, C.c_void_p(d_row_scale or None), C.c_void_p(d_f or None),
                                          C.c_void_p(d_jtj), C.c_void_p(d_jtf or None)))

    def fill_dprobs_fd(self, out, eps=1e-7, probs_out=None):
        rs = self._mat_stride(out)
        ps = self._vec_stride(probs_out, self.n_elements) if probs_out is not None else 1
        _lib.check(self._lib.b200_fill_dprobs_fd(self.ctx._h, self._h, float(eps), _ptr(out), rs,
                                                 _ptr(probs_out), ps))
        return out

    def fill_hprobs_linear(self, p1, p2, out):
        p1, p2 = _i32(p1), _i32(p2)
        if out.dtype != np.float64 or not out.flags.c_contiguous or \
                out.shape != (self.n_elements, p1.shape[0], p2.shape[0]):
            raise ValueError("expected C-contiguous float64 array (n_elements, n1, n2)")
        _lib.check(self._lib.b200_fill_hprobs_linear(self.ctx._h, self._h, p1.shape[0], _ptr(p1),
                                                     p2.shape[0], _ptr(p2), _ptr(out)))
        return out

    def fill_hprobs(self, p1, p2, out, hess=None):
        """Analytic Hessian rectangle for arbitrary members; ``hess`` is a ``packing.HessMap`` with the members' second
        derivatives for this rectangle (None / empty: members linear in their parameters)."""
        if hess is None or hess.nnz == 0:
            return self.fill_hprobs_linear(p1, p2, out)
        p1, p2 = _i32(p1), _i32(p2)
        if out.dtype != np.float64 or not out.flags.c_contiguous or \
                out.shape != (self.n_elements, p1.shape[0], p2.shape[0]):
            raise ValueError("expected C-contiguous float64 array (n_elements, n1, n2)")
        if hess.n1 != p1.shape[0] or hess.n2 != p2.shape[0] or hess.n_w != self.n_w:
            raise ValueError("second-derivative map does not match the rectangle")
        _lib.check(self._lib.b200_fill_hprobs(self.ctx._h, self._h, p1.shape[0], _ptr(p1), p2.shape[0], _ptr(p2),
                                              hess.nnz, _ptr(hess.rows), _ptr(hess.a), _ptr(hess.b), _ptr(hess.vals),
                                              _ptr(out)))
        return out

    def hessian_block(self, p1, p2, w_h, w_d, hess=None):
        """sum_el w_h[el] hprobs[el, p1, p2] + w_d[el] dprobs[el, p1] dprobs[el, p2]  -> (n1, n2); the per-element arrays
        stay on the device (`_hessian_from_block`, objectivefns.py:4914-4990)."""
        p1, p2 = _i32(p1), _i32(p2)
        wh = _f64(w_h).reshape(self.n_elements); wd = _f64(w_d).reshape(self.n_elements)
        out = np.zeros((p1.shape[0], p2.shape[0]))
        if hess is not None and (hess.n1 != p1.shape[0] or hess.n2 != p2.shape[0] or hess.n_w != self.n_w):
            raise ValueError("second-derivative map does not match the rectangle")
        nnz = hess.nnz if hess is not None else 0
        _lib.check(self._lib.b200_hessian_block(
            self.ctx._h, self._h, p1.shape[0], _ptr(p1), p2.shape[0], _ptr(p2), nnz,
            _ptr(hess.rows) if nnz else None, _ptr(hess.a) if nnz else None, _ptr(hess.b) if nnz else None,
            _ptr(hess.vals) if nnz else None, _ptr(wh), _ptr(wd), _ptr(out)))
        return out

    # ---- device-buffer fills (raw device pointers, e.g. torch.Tensor.data_ptr()) -----------------
    def fill_probs_dev(self, d_out_ptr):
        _lib.check(self._lib.b200_fill_probs_dev(self.ctx._h, self._h, C.c_void_p(int(d_out_ptr))))

    def fill_dprobs_dev(self, d_out_ptr, ld, d_probs_ptr=0):
        _lib.check(self._lib.b200_fill_dprobs_dev(self.ctx._h, self._h, C.c_void_p(int(d_out_ptr)), int(ld),
                                                  C.c_void_p(int(d_probs_ptr))))


def _atom_fill_dprobs_bcast_dev(self, d_out_ptr, ld, d_probs_ptr, peer_out_ptrs, peer_probs_ptrs=None):
    """Jacobian fill FUSED with the exchange (``b200_fill_dprobs_bcast_dev``): rows are stored into this rank's slot at
    ``d_out_ptr`` and, at the same time, into the same slot of every peer's array (``peer_out_ptrs``).  Asynchronous."""
    n = len(peer_out_ptrs)
    arr_t = C.c_void_p * max(n, 1)
    jo = arr_t(*[C.c_void_p(int(x)) for x in peer_out_ptrs])
    po = arr_t(*[C.c_void_p(int(x)) for x in peer_probs_ptrs]) if peer_probs_ptrs else None
    _lib.check(self._lib.b200_fill_dprobs_bcast_dev(self.ctx._h, self._h, C.c_void_p(int(d_out_ptr)), int(ld),
                                                    C.c_void_p(int(d_probs_ptr)), n, C.cast(jo, C.c_void_p),
                                                    C.cast(po, C.c_void_p) if po is not None else None))


Atom.fill_dprobs_bcast_dev = _atom_fill_dprobs_bcast_dev


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by cudaHostAlloc memory; the pages are released (cudaFreeHost) when the array and every view of it
    have been garbage collected."""
    import weakref
    lib = _lib.load()
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p(0)
    _lib.check(lib.b200_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)       # numpy keeps `buf` alive as the base of the array
    weakref.finalize(buf, lib.b200_host_free, C.c_void_p(p.value))
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
