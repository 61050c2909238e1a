"""The step after the hot path: solve the assembled phi-FEM system on the GPU.

The reference solves with PETSc KSP "preonly" + MUMPS LU and lets MUMPS detect the null pivots of the rows
no active cell touches (ICNTL(24) = 1; reference demo/strong-dirichlet/flower/main.py:138-157,
demo/weak-dirichlet/flower/main.py:161-184).  Here the CSR operator stays in HBM and is solved by
Jacobi-preconditioned BiCGStab: phi-FEM rows scale like phi^2, so diagonal scaling is what makes a Krylov method
converge (78 / 116 / 194 iterations at n = 32 / 64 / 128 on the disc problem of the tests).  Rows whose diagonal
entry is zero or absent -- dofs outside Omega_h, and the p dofs of the dual method away from the cut cells -- are
the null pivots: their unknowns are set to zero and they are left out of the iteration.

The whole iteration is hand-written (csrc/solve.cu, `phifem_bicgstab_iterate`): the unknowns are compacted to the active
rows (3.4 M of the 8.6 M box-mode rows at config E: the vectors then fit the L2), the matrix is read in place through
the row list and a column array remapped once per solve, and an iteration is two products fused with their dot
products, three fused vector updates and three one-block scalar updates -- every scalar stays in device memory,
the dot products are reduced in a fixed order (bitwise reproducible), the host reads |r|^2 back every `check_every`
iterations.  (Round 1 ran ~25 torch kernels per iteration over full-length vectors: 1.6 ms per iteration at config E.)
"""
import torch

from . import _lib


def spmv(A, x, out=None):
    """y = A x for a `CSRMatrix` on the device (csrc/solve.cu, phifem_csr_spmv)."""
    if out is None:
        out = torch.empty(A.shape[0], dtype=torch.float64, device=x.device)
    if A.data.numel() == 0:          # empty pattern (no active cell): nothing to point the kernel at
        return out.zero_()
    _lib.check(_lib.load().phifem_csr_spmv(A.shape[0], _lib.ptr(A.indptr), _lib.ptr(A.indices),
                                           _lib.ptr(A.data), _lib.ptr(x), _lib.ptr(out), _lib.stream()))
    return out


def diagonal(A):
    """Diagonal of a CSR matrix (zero where the pattern has no diagonal entry)."""
    n = A.shape[0]
    counts = (A.indptr[1:] - A.indptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n, device=A.data.device), counts)
    on = A.indices.long() == rows
    d = torch.zeros(n, dtype=torch.float64, device=A.data.device)
    d[rows[on]] = A.data[on]
    return d


class SolveInfo:
    def __init__(self, iterations, residual, converged, n_active):
        self.iterations, self.residual, self.converged, self.n_active = iterations, residual, converged, n_active

    def __repr__(self):
        return "SolveInfo(iterations=%d, residual=%.3e, converged=%s, n_active=%d)" % (
            self.iterations, self.residual, self.converged, self.n_active)


def bicgstab(A, b, rtol=1e-10, maxiter=10000, check_every=10, x0=None, pivot_tol=1e-14):
    """Solve A x = b on the rows with a non-zero diagonal (x = 0 elsewhere).  Returns (x, SolveInfo);
    `residual` is |b - A x| / |b| over the active rows.  pivot_tol: a row is a null pivot when |a_rr| <= pivot_tol * max_c |a_rc|.
    Callers must check `info.converged` (the demos raise when it is False).

    The iteration runs on the compact system of the active rows through the fused kernels of csrc/solve.cu
    (`phifem_bicgstab_iterate`: five kernels and three scalar updates per iteration, scalars in device memory); the host
    reads |r|^2 back every `check_every` iterations."""
    import ctypes
    if not b.is_cuda:
        raise RuntimeError("phifem_b200.solve: tensors must live on a CUDA device (no CPU fallback)")
    lib = _lib.load()
    dev, n = b.device, A.shape[0]
    if A.indptr.dtype != torch.int32 or A.indices.dtype != torch.int32:
        from .assemble import CSRMatrix
        A = CSRMatrix(A.indptr.to(torch.int32), A.indices.to(torch.int32), A.data, A.shape)
    # null pivots (MUMPS ICNTL(24)): rows without a diagonal entry or with one that is zero RELATIVE to the row -- a
    # diagonal of 1e-300 beside entries of order one would otherwise turn the Jacobi scaling into 1e300
    d = torch.empty(n, dtype=torch.float64, device=dev)
    rmax = torch.empty(n, dtype=torch.float64, device=dev)
    if A.data.numel():
        _lib.check(lib.phifem_csr_row_scan(n, _lib.ptr(A.indptr), _lib.ptr(A.indices), _lib.ptr(A.data), _lib.ptr(d),
                                           _lib.ptr(rmax), _lib.stream()))
    else:
        d.zero_()
        rmax.zero_()
    active = d.abs() > pivot_tol * rmax
    rows_a = torch.nonzero(active).reshape(-1)
    n_act = int(rows_a.numel())
    x_full = torch.zeros_like(b)
    if n_act == 0:
        return x_full, SolveInfo(0, 0.0, True, 0)
    bc = b[rows_a].contiguous()
    bnorm = float(torch.linalg.vector_norm(bc))
    if bnorm == 0.0:
        return x_full, SolveInfo(0, 0.0, True, n_act)
    # compact numbering: active row k of the matrix is unknown k; columns of inactive rows point at the zero slot n_act
    cmap = torch.full((n,), n_act, dtype=torch.int32, device=dev)
    cmap[rows_a] = torch.arange(n_act, dtype=torch.int32, device=dev)
    cols = torch.empty_like(A.indices)
    _lib.check(lib.phifem_remap_columns(A.indices.numel(), _lib.ptr(A.indices), _lib.ptr(cmap), _lib.ptr(cols),
                                        _lib.stream()))
    rows32 = rows_a.to(torch.int32).contiguous()
    nnz = int(A.data.numel())
    minv = (1.0 / d[rows_a]).contiguous()
    spmv_grid = (n_act * 8 + 255) // 256
    partials = torch.zeros(2 * (spmv_grid + 8 * 148 + 1), dtype=torch.float64, device=dev)
    vec = lambda extra=0: torch.zeros(n_act + extra, dtype=torch.float64, device=dev)     # noqa: E731
    x, p, v, s, t, y, z = vec(1), vec(), vec(), vec(), vec(), vec(1), vec(1)
    st = _lib.stream()
    P = _lib.ptr

    def product(src, dst):      # dst = A src on the compact system (src has the trailing zero slot)
        _lib.check(lib.phifem_csr_spmv_rows(n_act, nnz, P(rows32), P(A.indptr), P(cols), P(A.data), P(src), P(dst),
                                            P(partials), st))

    if x0 is not None:
        x[:n_act] = x0.to(dev, dtype=torch.float64)[rows_a]
        product(x, t)
        r = bc - t
    else:
        r = bc.clone()
    rhat = r.clone()
    rr = torch.dot(r, r)
    partials[0] = rr
    partials[1] = rr
    state = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=dev)
    n_partials = ctypes.c_int32(1)
    res, it, converged = float(torch.sqrt(rr)) / bnorm, 0, False
    if res <= rtol:
        converged = True
    while it < maxiter and not converged:
        k = min(check_every, maxiter - it)
        _lib.check(lib.phifem_bicgstab_iterate(n_act, nnz, P(rows32), P(A.indptr), P(cols), P(A.data), P(minv), P(rhat),
                                               P(x), P(r), P(p), P(v), P(s), P(t), P(y), P(z), P(state), P(partials),
                                               ctypes.byref(n_partials), k, st))
        it += k
        rr_now = float(state[4])
        res = (rr_now ** 0.5) / bnorm if rr_now == rr_now and rr_now >= 0.0 else float("nan")
        if not (res == res):            # NaN: breakdown
            break
        if res <= rtol:
            converged = True
    # true residual of the returned iterate
    product(x, t)
    res = float(torch.linalg.vector_norm(bc - t)) / bnorm
    x_full[rows_a] = x[:n_act]
    return x_full, SolveInfo(it, res, converged and res <= 10 * rtol, n_act)
