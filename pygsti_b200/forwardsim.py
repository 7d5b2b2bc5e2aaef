"""
``B200ForwardSimulator`` -- the drop-in pyGSTi forward simulator backed by the B200 CUDA engine.

    import pygsti
    from pygsti_b200.forwardsim import B200ForwardSimulator
    model.sim = B200ForwardSimulator()          # instead of 'map' / MapForwardSimulator()
    model.probabilities(circuit); model.sim.bulk_fill_probs(...); bulk_fill_dprobs(...); bulk_fill_hprobs(...)

It subclasses the reference's ``MapForwardSimulator`` (pygsti/forwardsims/mapforwardsim.py:127) and keeps
its layout machinery (``create_layout`` -> ``MapCOPALayout``, atoms, parameter blocks, MPI resource
allocation) untouched; only the three atom-level fills are replaced
(``_bulk_fill_probs_atom`` / ``_bulk_fill_dprobs_atom`` / ``_bulk_fill_hprobs_atom``,
mapforwardsim.py:372-391), by routing ``self.calclib`` to ``pygsti_b200.calclib``.

Requires pyGSTi to be importable (it is the host application); the engine itself does not.
"""
import numpy as _np

from pygsti.forwardsims.mapforwardsim import MapForwardSimulator as _MapForwardSimulator

from . import calclib as _b200_calclib


class B200ForwardSimulator(_MapForwardSimulator):
    """
    Parameters (in addition to those of ``MapForwardSimulator``, mapforwardsim.py:166-172)
    ----------
    derivative_mode : 'analytic' (default) or 'fd'
        'analytic': adjoint Jacobian, equal to the reference's MatrixForwardSimulator to ~1e-14.
        'fd': the reference Map simulator's forward differences (pyx:349-378) evaluated on the GPU with
        ``derivative_eps``.  Members are perturbed to ``M + eps * dM/dtheta_p``: the reference's own values for members linear
        in their parameters (full / TP / static), first-order-linearised (difference O(eps * d2M)) for CPTPLND / H+S members.
        The fused objective-function fills (``pygsti_b200.objective``) delegate to the stock methods in this mode.
    device : int or None
        CUDA device of this process (default: ``LOCAL_RANK`` modulo the device count, else 0).  With
        ``num_atoms > 1`` and no MPI communicator, atoms are spread round-robin over ``devices``.
    devices : sequence of int or None
        GPUs to spread a single process's atoms over (default: just ``device``).
    fused_objective : bool
        True (default): a stock pyGSTi objective function / Levenberg-Marquardt run on this simulator uses the fused row scaling
        and the on-device J^T J / J^T f (``pygsti_b200.objective.install_hooks``); False: pyGSTi's own host-side passes.
    device_lindblad : bool
        True (default since round 2): for Lindblad-parameterised members (CPTPLND / H+S / GLND gates, preps, effects) the dense
        algebra of a model update -- error generators, exponentials, Frechet derivatives, composition with the static parts --
        runs on the device (``b200_lindblad_members``: 1.2 ms of kernels at BASELINE config 4 instead of 1.0-1.2 s of
        ``to_dense`` / ``deriv_wrt_params`` on the host); models without such members, or with a parameter interposer, take the
        host packing path as before.
    analytic_hessian : bool
        True (default): `bulk_fill_hprobs` is fully analytic on the device for every member that provides
        `hessian_wrt_params`.  False: members not linear in their parameters go through the reference's own
        finite-difference driver (mapforwardsim.py:394-438) on top of the analytic device Jacobian.
    """

    def __init__(self, model=None, max_cache_size=None, num_atoms=None, processor_grid=None, param_blk_sizes=None,
                 derivative_eps=1e-7, hessian_eps=1e-5, derivative_mode='analytic', device=None, devices=None,
                 analytic_hessian=True, device_model_update=False, host_prefix_cache=False, device_lindblad=True,
                 fused_objective=True):
        if derivative_mode not in ('analytic', 'fd'):
            raise ValueError("derivative_mode must be 'analytic' or 'fd'")
        self.derivative_mode = derivative_mode
        self.analytic_hessian = bool(analytic_hessian)   # False: non-linear members use the reference's FD driver
        self.device_model_update = bool(device_model_update)
        self.host_prefix_cache = bool(host_prefix_cache)
        self.device_lindblad = bool(device_lindblad)
        self.fused_objective = bool(fused_objective)
        if self.fused_objective:                       # stock dterms / dlsvec / LM normal equations reach the fused kernels
            from . import objective as _objective
            _objective.install_hooks()
        if max_cache_size is None and not self.host_prefix_cache:
            max_cache_size = 0                             # SURVEY 8f rank 4 (layout construction): no host-side cache plan
        self._b200_device = device
        self._b200_devices = tuple(devices) if devices is not None else None
        super().__init__(model, max_cache_size, num_atoms, processor_grid, param_blk_sizes,
                         derivative_eps, hessian_eps)

    # ---- plug-in plumbing -----------------------------------------------------------------------
    def _set_evotype(self, evotype):
        # The reference picks mapforwardsim_calc_<evotype> here (mapforwardsim.py:93-102).  We accept the
        # dense-superoperator evotypes ('densitymx' and its numpy twin 'densitymx_slow': both expose
        # to_dense('HilbertSchmidt')) and route every atom fill to the B200 engine.
        if evotype is not None:
            name = getattr(evotype, 'name', str(evotype))
            if not name.startswith('densitymx'):
                raise ValueError("B200ForwardSimulator supports the 'densitymx' evotype only (got %r)" % name)
            self.calclib = _b200_calclib
        else:
            self.calclib = None

    def _to_nice_serialization(self):
        state = super()._to_nice_serialization()
        state.update({'derivative_mode': self.derivative_mode, 'device': self._b200_device,
                      'analytic_hessian': self.analytic_hessian, 'device_model_update': self.device_model_update,
                      'host_prefix_cache': self.host_prefix_cache, 'device_lindblad': self.device_lindblad,
                      'fused_objective': self.fused_objective})
        return state

    @classmethod
    def _from_nice_serialization(cls, state):
        return cls(None, state['max_cache_size'],
                   derivative_eps=state.get('derivative_epsilon', 1e-7),
                   hessian_eps=state.get('hessian_epsilon', 1e-5),
                   derivative_mode=state.get('derivative_mode', 'analytic'),
                   device=state.get('device', None), analytic_hessian=state.get('analytic_hessian', True),
                   device_model_update=state.get('device_model_update', False),
                   host_prefix_cache=state.get('host_prefix_cache', False),
                   device_lindblad=state.get('device_lindblad', True), fused_objective=state.get('fused_objective', True))

    def copy(self, keep_model_attached=True):
        # MapForwardSimulator.copy hard-codes its own class (mapforwardsim.py:190-204) -> must override,
        # otherwise `model.copy()` / a "stolen" simulator silently degrades to the CPU simulator.
        out = B200ForwardSimulator(self.model, self._max_cache_size, self._num_atoms, self._processor_grid,
                                   self._pblk_sizes, self.derivative_eps, self.hessian_eps,
                                   self.derivative_mode, self._b200_device, self._b200_devices, self.analytic_hessian,
                                   self.device_model_update, self.host_prefix_cache, self.device_lindblad, self.fused_objective)
        if not keep_model_attached:
            out.model = None
        return out

    # ---- layout: tag atoms with a device (single-process multi-GPU) ------------------------------
    def create_layout(self, *args, **kwargs):
        layout = super().create_layout(*args, **kwargs)
        devs = self._b200_devices
        if devs:
            for i, atom in enumerate(layout.atoms):
                atom._b200_device = devs[i % len(devs)]
        return layout

    # ---- single-process multi-GPU: atoms on different GPUs run CONCURRENTLY --------------------------------------------
    def _device_groups(self, layout):
        """[[atoms of GPU 0], [atoms of GPU 1], ...] when this process drives several GPUs and nothing is distributed over MPI
        (the reference then loops over the local atoms one after another, distforwardsim.py:98-99, 123-144); else None."""
        devs = self._b200_devices
        if not devs or len(set(devs)) < 2 or len(layout.atoms) < 2:
            return None
        from .objective import _distributed
        if _distributed(layout):
            return None
        groups = {}
        for atom in layout.atoms:
            groups.setdefault(getattr(atom, "_b200_device", devs[0]), []).append(atom)
        return list(groups.values()) if len(groups) > 1 else None

    @staticmethod
    def _run_groups(groups, work):
        """One host thread per GPU; a GPU's atoms run in order on its own engine context (the C calls release the GIL, so the
        kernels and the device-to-host copies of different GPUs overlap).  Exceptions are re-raised on the caller."""
        from concurrent.futures import ThreadPoolExecutor

        def run(group):
            for item in group:
                work(item)
        with ThreadPoolExecutor(max_workers=len(groups)) as ex:
            for fut in [ex.submit(run, g) for g in groups]:
                fut.result()

    def _bulk_fill_probs(self, array_to_fill, layout):
        """Replaces distforwardsim.py:92-103 when atoms live on several GPUs of this process: same result, atoms of
        different GPUs filled concurrently into their disjoint ``atom.element_slice`` rows."""
        groups = self._device_groups(layout)
        if groups is None:
            return super()._bulk_fill_probs(array_to_fill, layout)
        ralloc = layout.resource_alloc('atom-processing')
        ralloc.host_comm_barrier()
        prepared = {}
        for atom in layout.atoms:                       # host packing + uploads: serial (pyGSTi objects are not thread-safe)
            ctx, ent = _b200_calclib._engine_atom(self, atom)
            _b200_calclib._upload_model(self, atom, ent)
            prepared[id(atom)] = ent["atom"]

        def work(atom):
            dst = array_to_fill[atom.element_slice]
            if dst.dtype == _np.float64 and dst.ndim == 1 and dst.shape[0] == atom.num_elements:
                prepared[id(atom)].fill_probs(dst)
            else:
                tmp = _np.empty(atom.num_elements); prepared[id(atom)].fill_probs(tmp); dst[...] = tmp
        self._run_groups(groups, work)
        ralloc.host_comm_barrier()

    def _bulk_fill_dprobs(self, array_to_fill, layout, pr_array_to_fill):
        """Replaces distforwardsim.py:110-146 in the same situation (no parameter blocks requested): every atom's Jacobian rows
        and probabilities are produced on its GPU and copied into the caller's arrays while the other GPUs do the same."""
        groups = self._device_groups(layout)
        if groups is None or layout.param_dimension_blk_sizes[0] is not None:
            return super()._bulk_fill_dprobs(array_to_fill, layout, pr_array_to_fill)
        ralloc = layout.resource_alloc('atom-processing')
        ralloc.host_comm_barrier()
        prepared = {}
        for atom in layout.atoms:
            prepared[id(atom)] = _b200_calclib.prepare_dprobs_atom(self, atom, layout.global_param_slice)

        def work(atom):
            eng_atom, pidx = prepared[id(atom)]
            pr = None if pr_array_to_fill is None else pr_array_to_fill[atom.element_slice]
            _b200_calclib.run_dprobs_atom(self, eng_atom, pidx, array_to_fill[atom.element_slice, :], None, None, atom,
                                          self.derivative_eps, pr_array_to_fill=pr)
        self._run_groups(groups, work)
        ralloc.host_comm_barrier()

    # ---- hessian ------------------------------------------------------------------------------------
    def _bulk_fill_hprobs_atom(self, array_to_fill, dest_param_slice1, dest_param_slice2, layout_atom,
                               param_slice1, param_slice2, resource_alloc):
        """Replaces mapforwardsim.py:384-391.  Members linear in their parameters (full / TP / static):
        fully analytic second order on the device (== MatrixForwardSimulator to ~1e-13).  Other members (CPTPLND,
        H+S, ...): the same plus the members' own `hessian_wrt_params` second-derivative term, also on the device
        (`b200_fill_hprobs`).  Only if a member cannot provide that (parameter interposer, no analytic Hessian) does
        the reference's own driver `_mapfill_hprobs_atom` (mapforwardsim.py:394-438: finite differences along the
        first parameter axis, `hessian_eps`) run on top of the ANALYTIC device Jacobian."""
        if self.derivative_mode == 'analytic' and _b200_calclib.all_members_linear(self, layout_atom):
            _b200_calclib.mapfill_hprobs_atom_linear(self, array_to_fill, dest_param_slice1, dest_param_slice2,
                                                     layout_atom, param_slice1, param_slice2, resource_alloc)
            return
        if self.derivative_mode == 'analytic' and self.analytic_hessian and \
                _b200_calclib.mapfill_hprobs_atom_analytic(self, array_to_fill, dest_param_slice1, dest_param_slice2,
                                                           layout_atom, param_slice1, param_slice2, resource_alloc):
            return
        super()._bulk_fill_hprobs_atom(array_to_fill, dest_param_slice1, dest_param_slice2, layout_atom,
                                       param_slice1, param_slice2, resource_alloc)

    # ---- extensions for the objective-function Jacobian fill --------------------
    def bulk_fill_dprobs_scaled(self, array_to_fill, layout, row_scale, pr_array_to_fill=None):
        """``bulk_fill_dprobs`` with every row multiplied by ``row_scale[el]`` on the device: the
        `dprobs *= dg_probs[:, None]` / `jac *= p5over_lsvec[:, None]` passes of the objective functions
        (objectivefns.py:4609-4616, 4644-4649) fused into the kernel epilogue."""
        self._require_local(layout, "bulk_fill_dprobs_scaled")
        if pr_array_to_fill is not None:
            self.bulk_fill_probs(pr_array_to_fill, layout)
        ralloc = layout.resource_alloc('param-processing')
        for atom in layout.atoms:
            sl = atom.element_slice
            _b200_calclib.mapfill_dprobs_atom(self, array_to_fill[sl, :], slice(0, atom.num_elements), None, atom,
                                              layout.global_param_slice, ralloc, self.derivative_eps,
                                              row_scale=_np.ascontiguousarray(row_scale[sl]))
        return array_to_fill

    def bulk_jtj(self, layout, row_scale=None, f=None):
        """(J^T J, J^T f) for J = diag(row_scale) . dprobs without the Jacobian leaving the device (what
        `fill_jtj` / `fill_jtf` hand the Levenberg-Marquardt step, distlayout.py:1220-1359)."""
        self._require_local(layout, "bulk_jtj")
        JTJ, JTf = None, None
        for atom in layout.atoms:
            sl = atom.element_slice
            a, b = _b200_calclib.atom_jtj(self, atom, None if row_scale is None else row_scale[sl],
                                          None if f is None else f[sl])
            JTJ = a if JTJ is None else JTJ + a
            if b is not None:
                JTf = b if JTf is None else JTf + b
        return JTJ, JTf

    def bulk_hessian_block(self, layout, param_slice1, param_slice2, w_h, w_d):
        """sum_el w_h[el] hprobs[el, s1, s2] + w_d[el] dprobs[el, s1] dprobs[el, s2] over all local atoms, each reduced on
        the device: the per-rectangle work of the MLE Hessian (`_construct_hessian` / `_hessian_from_block`,
        objectivefns.py:1576-1693, 4914-4990) without moving (nE x B1 x B2) arrays to the host.  Returns None if a member
        has no analytic second derivative."""
        self._require_local(layout, "bulk_hessian_block")
        total = None
        for atom in layout.atoms:
            sl = atom.element_slice
            blk = _b200_calclib.atom_hessian_block(self, atom, param_slice1, param_slice2, w_h[sl], w_d[sl])
            if blk is None:
                return None
            total = blk if total is None else total + blk
        return total

    @staticmethod
    def _require_local(layout, what):
        """The fused reductions sum over the atoms of THIS process only.  With a layout split over MPI processors the reference
        all-reduces over atom processors (fill_jtj / _gather_hessian, distlayout.py:1220-1359, objectivefns.py:1698-1737); use the
        stock objective-function methods there (pygsti_b200.objective delegates automatically) or, with one process per GPU
        under torch.distributed, pygsti_b200.dist.allreduce_jtj."""
        from .objective import _distributed
        if _distributed(layout):
            raise NotImplementedError("%s: layout is distributed over MPI processors; use the objective function's own "
                                      "dterms/dlsvec/hessian (they run on this simulator) or pygsti_b200.dist" % what)

    def __getstate__(self):
        state = super().__getstate__()
        state.pop('calclib', None)
        return state
