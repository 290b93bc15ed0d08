"""fenics_ii / cbc.block façade: ii_assemble, ii_convert, apply_bc, LUSolver, ii_Function.

The reference idiom (tests/test_flow_models.py:47-52, demo/Tree Profiling.ipynb:88-96)

    A, b = map(ii_assemble, (a, L)); A, b = map(ii_convert, (A, b))
    qp = ii_Function(W); LUSolver(A, "mumps").solve(qp.vector(), b)

runs unchanged: `a`/`L` are opaque form handles produced by the model classes, `ii_assemble`
launches the assembly kernels, `ii_convert` is the identity (the matrix is assembled directly in
monolithic CSR), and `LUSolver.solve` runs the Schur-complement solver on the GPU.
"""
import numpy as np

from .function import ii_Function, Vector

__all__ = ["ii_assemble", "ii_convert", "apply_bc", "LUSolver", "ii_Function", "AssembledMatrix",
           "AssembledVector"]


class AssembledMatrix:
    """Monolithic CSR on the device + what the structured solver needs (operator coefficients)."""

    def __init__(self, system, vals, coeffs, kind, bcs=None):
        self.system, self.vals, self.coeffs, self.kind = system, vals, coeffs, kind
        self.bcs = bcs

    @property
    def shape(self):
        return (self.system.N, self.system.N)

    def tocsr(self):
        """scipy CSR copy on the host (tests / inspection)."""
        import scipy.sparse as sp
        rp, ci = self.system.build_pattern()
        return sp.csr_matrix((self.vals.cpu().numpy(), ci.cpu().numpy(), rp.cpu().numpy()), shape=self.shape)

    def array(self):
        return self.tocsr().toarray()

    def __mul__(self, x):
        t = x.tensor() if hasattr(x, "tensor") and callable(x.tensor) else (x.tensor if isinstance(x, Vector) else x)
        return AssembledVector(self.system, self.system.spmv(self.vals, t))


class AssembledVector:
    def __init__(self, system, data):
        self.system, self.data = system, data

    def get_local(self):
        return self.data.cpu().numpy()

    def array(self):
        return self.get_local()

    @property
    def tensor(self):
        return self.data


def ii_assemble(form):
    """xii.ii_assemble: form handle -> assembled device object."""
    if isinstance(form, (AssembledMatrix, AssembledVector)):
        return form
    if hasattr(form, "assemble"):
        return form.assemble()
    raise TypeError("ii_assemble expects a form handle returned by a graphnics_b200 model "
                    f"(a_form()/L_form()/mass_matrix_q()/mass_vector_q()), got {type(form).__name__}")


def ii_convert(obj, *a, **k):
    """xii.ii_convert: block -> monolithic.  Already monolithic here."""
    return obj


def apply_bc(A, b, bcs):
    """xii.apply_bc(A, b, bcs) (flow_models.py:117): symmetric Dirichlet elimination."""
    flat = [bc for row in (bcs or []) for bc in (row if isinstance(row, (list, tuple)) else [row])]
    if not flat:
        return A, b
    return A.system.apply_bc(A, b, flat)


class LUSolver:
    """dolfin.LUSolver(A, "mumps") stand-in.  `solve(x, b)` writes the solution into x."""

    def __init__(self, A, method="mumps", rtol=1e-13):
        self.A, self.method, self.rtol = A, method, rtol
        self.info = None

    def solve(self, x, b):
        xt = x.tensor if isinstance(x, Vector) else x
        bt = b.tensor if isinstance(b, (AssembledVector, Vector)) else b
        _, self.info = self.A.system.solve(self.A.coeffs, bt, x=xt, rtol=self.rtol, bcs=self.A.bcs) \
            if self.A.bcs is not None else self.A.system.solve(self.A.coeffs, bt, x=xt, rtol=self.rtol)
        if isinstance(x, Vector):
            x.touch()
        return self.info
