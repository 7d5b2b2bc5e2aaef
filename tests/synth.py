"""Synthetic layouts / models for edge-case tests (no pyGSTi needed): random dense models and hand-made prefix tables."""
import numpy as np

from pygsti_b200.packing import AtomTables, DerivMap


def make_tables(dim, n_ops, n_rho, n_eff, circuits, use_cache=True):
    """circuits: list of (prep, [ops], [effect indices]).  With use_cache, circuits that extend an earlier circuit's
    full sequence start from its cached final state (exercises the iStart/iCache links of the reference format)."""
    row_dest, row_istart, row_icache, row_prep, row_ptr, row_ops = [], [], [], [], [0], []
    out_ptr, out_eff, out_el = [0], [], []
    cache = {}
    n_el = 0
    for k, (prep, ops, effs) in enumerate(circuits):
        key = (prep, tuple(ops))
        start, rem = -1, list(ops)
        if use_cache:
            for cut in range(len(ops), 0, -1):
                kk = (prep, tuple(ops[:cut]))
                if kk in cache:
                    start, rem = cache[kk], list(ops[cut:])
                    break
        row_dest.append(k); row_istart.append(start); row_prep.append(prep if start < 0 else -1)
        slot = -1
        if use_cache and key not in cache and len(ops) > 0:
            slot = len(cache); cache[key] = slot
        row_icache.append(slot)
        row_ops.extend(rem); row_ptr.append(len(row_ops))
        for e in effs:
            out_eff.append(e); out_el.append(n_el); n_el += 1
        out_ptr.append(len(out_eff))
    i32 = lambda x: np.asarray(x, dtype=np.int32)
    return AtomTables(dim=dim, n_ops=n_ops, n_rho=n_rho, n_eff=n_eff, n_elements=n_el, cache_size=len(cache),
                      row_dest=i32(row_dest), row_istart=i32(row_istart), row_icache=i32(row_icache),
                      row_prep=i32(row_prep), row_ptr=i32(row_ptr), row_ops=i32(row_ops),
                      out_ptr=i32(out_ptr), out_eff=i32(out_eff), out_el=i32(out_el))


def random_model(dim, n_ops, n_rho, n_eff, seed=0):
    rng = np.random.default_rng(seed)
    G = np.eye(dim)[None] * 0.9 + 0.3 / np.sqrt(dim) * rng.standard_normal((n_ops, dim, dim))
    rho = rng.standard_normal((n_rho, dim)) / np.sqrt(dim)
    E = rng.standard_normal((n_eff, dim)) / np.sqrt(dim)
    return G, rho, E


def random_circuits(n, max_depth, n_ops, n_rho, n_eff, seed=0, subsets=True):
    rng = np.random.default_rng(seed)
    out = [(0, [], list(range(n_eff)))]                      # the empty circuit
    for _ in range(n - 1):
        L = int(rng.integers(0, max_depth + 1))
        ops = [int(x) for x in rng.integers(0, n_ops, size=L)]
        if subsets and rng.random() < 0.3:
            effs = sorted(int(x) for x in rng.choice(n_eff, size=int(rng.integers(1, n_eff + 1)), replace=False))
        else:
            effs = list(range(n_eff))
        out.append((int(rng.integers(0, n_rho)), ops, effs))
    # shared prefixes / suffixes and an exact duplicate
    base = [int(x) for x in rng.integers(0, n_ops, size=max(2, max_depth // 2))]
    for j in range(4):
        out.append((0, base + [j % n_ops], list(range(n_eff))))
        out.append((j % n_rho, [j % n_ops] + base, list(range(n_eff))))
    out.append(out[-1])
    return out


def full_derivs(t):
    """Every dense member element is its own parameter (unit permutation: the fused Jacobian path)."""
    d = t.dim
    n_w = t.n_ops * d * d + (t.n_rho + t.n_eff) * d
    idx = np.arange(n_w, dtype=np.int32)
    return DerivMap(n_w, n_w, idx, idx.copy(), np.ones(n_w))


def random_derivs(t, n_params, density=0.02, seed=0):
    d = t.dim
    n_w = t.n_ops * d * d + (t.n_rho + t.n_eff) * d
    rng = np.random.default_rng(seed)
    nnz = max(n_params, int(density * n_w * n_params))
    rows = rng.integers(0, n_w, size=nnz).astype(np.int32)
    cols = rng.integers(0, n_params, size=nnz).astype(np.int32)
    vals = rng.standard_normal(nnz)
    return DerivMap(n_w, n_params, rows, cols, vals)
