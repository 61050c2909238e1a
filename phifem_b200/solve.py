"""The step after the hot path: solve the assembled phi-FEM system on the GPU.

The reference solves with PETSc KSP "preonly" + MUMPS LU and lets MUMPS detect the null pivots of the rows
no active cell touches (ICNTL(24) = 1; reference demo/strong-dirichlet/flower/main.py:138-157,
demo/weak-dirichlet/flower/main.py:161-184).  Here the CSR operator stays in HBM and is solved by
Jacobi-preconditioned BiCGStab: phi-FEM rows scale like phi^2, so diagonal scaling is what makes a Krylov method
converge (78 / 116 / 194 iterations at n = 32 / 64 / 128 on the disc problem of the tests).  Rows whose diagonal
entry is zero or absent -- dofs outside Omega_h, and the p dofs of the dual method away from the cut cells -- are
the null pivots: their unknowns are set to zero and they are left out of the iteration.

The matrix-vector product is the hand-written kernel of csrc/solve.cu; the vector updates are torch ops with
device-resident scalars (no host synchronisation inside an iteration; the residual is read back every
`check_every` iterations).
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
    Callers must check `info.converged` (the demos raise when it is False)."""
    if not b.is_cuda:
        raise RuntimeError("phifem_b200.solve: tensors must live on a CUDA device (no CPU fallback)")
    d = diagonal(A)
    # null pivots (MUMPS ICNTL(24)): rows without a diagonal entry or with one that is zero RELATIVE to the row -- a
    # diagonal of 1e-300 beside entries of order one would otherwise turn the Jacobi scaling into 1e300
    counts = (A.indptr[1:] - A.indptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(A.shape[0], device=b.device), counts)
    rmax = torch.zeros_like(d).scatter_reduce_(0, rows, A.data.abs(), reduce="amax", include_self=True)
    del rows
    active = d.abs() > pivot_tol * rmax
    minv = torch.where(active, 1.0 / torch.where(active, d, torch.ones_like(d)), torch.zeros_like(d))
    mask = active.to(torch.float64)
    bm = b * mask
    bnorm = float(torch.linalg.vector_norm(bm))
    x = torch.zeros_like(b) if x0 is None else (x0 * mask)
    if bnorm == 0.0:
        return x, SolveInfo(0, 0.0, True, int(active.sum()))
    r = bm - spmv(A, x) * mask if x0 is not None else bm.clone()
    rhat = r.clone()
    rho = alpha = omega = torch.ones((), dtype=torch.float64, device=b.device)
    v = torch.zeros_like(b)
    p = torch.zeros_like(b)
    t = torch.empty_like(b)
    res, it, converged = 1.0, 0, False
    tiny = 1e-300
    while it < maxiter:
        rho_new = torch.dot(rhat, r)
        beta = (rho_new / (rho + tiny)) * (alpha / (omega + tiny))
        p = r + beta * (p - omega * v)
        y = p * minv
        spmv(A, y, out=v)
        v *= mask
        alpha = rho_new / (torch.dot(rhat, v) + tiny)
        s = r - alpha * v
        z = s * minv
        spmv(A, z, out=t)
        t *= mask
        omega = torch.dot(t, s) / (torch.dot(t, t) + tiny)
        x = x + alpha * y + omega * z
        r = s - omega * t
        rho = rho_new
        it += 1
        if it % check_every == 0 or it == maxiter:
            res = float(torch.linalg.vector_norm(r)) / bnorm
            if not (res == res):            # NaN: breakdown
                break
            if res <= rtol:
                converged = True
                break
    # true residual of the returned iterate
    res = float(torch.linalg.vector_norm(bm - spmv(A, x) * mask)) / bnorm
    return x, SolveInfo(it, res, converged and res <= 10 * rtol, int(active.sum()))
