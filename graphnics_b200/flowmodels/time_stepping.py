"""Time-dependent models and the theta-scheme loop, drop-in for
graphnics/flowmodels/time_stepping.py:17-221.

    (D + theta dt A) x^{n+1} = D x^n + theta dt L^{n+1} - (1-theta) dt A x^n + (1-theta) dt L^n

The reference re-assembles, re-converts and re-factorises (fresh MUMPS LU) every step
(:179-192).  Here the operator D + theta*dt*A is never formed: the structured solver takes its edge
coefficients (a_e = Ainv + theta*dt*Res_e, sigma = theta*dt), so a step is: right-hand-side
kernels + one fused linear combination + eliminate / Schur-CG / back-substitute + one mass apply.
With `reassemble_lhs=False` nothing but the right-hand side is touched.

Reference quirks kept on purpose: the loop runs t_steps-1 times and returns
t_steps states; the loop variable is np.linspace(dt, T, t_steps-1) while the scheme uses dt;
`t.assign(dt)` before the loop does not move the `.t` attribute of Expression data; forms (and,
for the mixed model, the L2 projections inside L_form) are rebuilt at the END of each iteration,
so projected data lag the inlet term by one step.
"""
import os

import numpy as np

from ..coefficients import Constant
from ..function import ii_Function
from ..la import ii_assemble, apply_bc, AssembledVector
from .flow_models import HydraulicNetwork, MixedHydraulicNetwork, _MixedMassForm, _MassVectorForm

__all__ = ["time_stepping_schemes", "TimeDepHydraulicNetwork", "TimeDepMixedHydraulicNetwork",
           "time_stepping_stokes"]

# time_stepping.py:17
time_stepping_schemes = {"IE": {"b1": 0, "b2": 1}, "CN": {"b1": 0.5, "b2": 0.5}}


class TimeDepHydraulicNetwork(HydraulicNetwork):
    """time_stepping.py:20-64"""

    def __init__(self, G, f=Constant(0), g=Constant(0), p_bc=Constant(0), Res=Constant(1), Ainv=Constant(1)):
        self.Ainv = Ainv
        super().__init__(G, f=f, g=g, p_bc=p_bc, Res=Res)

    def mass_matrix_q(self, coeff):
        from ..engine_primal import PrimalMassForm
        return PrimalMassForm(self, coeff)

    def mass_vector_q(self, u, coeff):
        return _MassVectorForm(self, u, coeff)


class TimeDepMixedHydraulicNetwork(MixedHydraulicNetwork):
    """time_stepping.py:68-123.  Like the reference, the `Res` argument is accepted and dropped
    (:70-73): resistance comes from the edge attribute "Res"."""

    def __init__(self, G, f=Constant(0), g=Constant(0), p_bc=Constant(0), Res=Constant(1), Ainv=Constant(1)):
        self.Ainv = Ainv
        super().__init__(G, f=f, g=g, p_bc=p_bc)

    def mass_matrix_q(self, coeff):
        return _MixedMassForm(self, coeff)

    def mass_vector_q(self, u, coeff):
        return _MassVectorForm(self, u, coeff)


def _cell_means(dn, coef):
    """L2 projection onto DG0 of the global mesh = cell means (global cell order), Simpson on every cell
    (midpoint / trapezoid for degree <= 1 data)"""
    import torch
    c = dn.cells.long()
    xa, xb = dn.coords[c[:, 0]], dn.coords[c[:, 1]]
    pts = torch.stack([xa, 0.5 * (xa + xb), xb], dim=1).reshape(-1, dn.gdim).contiguous()
    ci = np.repeat(np.arange(dn.num_cells), 3) if coef.user is not None else None
    v = coef.eval_device(pts, pts_per_edge=3 * dn.m, tangents=dn.tangents, cell_index=ci).reshape(-1, 3)
    if coef.degree <= 1:
        return v[:, 1] if coef.degree == 0 else 0.5 * (v[:, 0] + v[:, 2])
    return (v[:, 0] + 4.0 * v[:, 1] + v[:, 2]) / 6.0


def _as_state(model, qp0):
    """initial state as a monolithic device vector in W layout (time_stepping.py:147-165): None (zero), an
    ii_Function on model.W, or -- like the reference, which projects component by component (:162-165) -- a
    list with one entry per space of W: Functions on W[i] are copied, Constants / numbers / Expressions / UFL
    are projected (per-edge CG1: consistent-mass L2 projection, the kernel L_form uses; DG0: cell means; Real: the
    value of a Constant; CG1 on the global mesh of the primal model: nodal interpolation)."""
    import torch
    from ..coefficients import as_coefficient
    from ..function import Function
    W = model.W
    if qp0 is None:
        return ii_Function(W)
    if isinstance(qp0, ii_Function):
        return qp0.copy()
    if not isinstance(qp0, (list, tuple)) or len(qp0) != len(W):
        raise TypeError("qp0 must be None, an ii_Function on model.W or a list with one entry per space of model.W")
    sys_ = model._system()
    dn = sys_.dn
    out = ii_Function(W)
    x = out.tensor()
    off = W.offsets() if hasattr(W, "offsets") else np.concatenate([[0], np.cumsum([V.dim() for V in W])])

    def put(i, comp, make):
        if isinstance(comp, Function):
            x[int(off[i]):int(off[i + 1])].copy_(comp.vector().tensor)
        else:
            x[int(off[i]):int(off[i + 1])].copy_(make(as_coefficient(comp)))

    if hasattr(W, "role"):                                   # mixed model: [q_0 .. q_{E-1} | p | lambda_0 ..]
        E, m = dn.E, dn.m
        whole = {}                                           # one projection per distinct object, every edge's slice from it

        def q_block(i):
            def make(cc, i=i):
                key = id(qp0[i])
                if key not in whole:
                    whole[key] = sys_.project_coefficient(cc)
                return whole[key][i * (m + 1):(i + 1) * (m + 1)]
            return make
        if all(qp0[i] is qp0[0] for i in range(E)) and not isinstance(qp0[0], Function):
            x[: E * (m + 1)].copy_(sys_.project_coefficient(as_coefficient(qp0[0])))
        else:
            for i in range(E):
                put(i, qp0[i], q_block(i))
        put(E, qp0[E], lambda cc: _cell_means(dn, cc))
        for k in range(dn.B):
            comp = qp0[E + 1 + k]
            if isinstance(comp, Function):
                put(E + 1 + k, comp, None)
            else:
                cc = as_coefficient(comp)
                if not cc.is_constant:
                    raise NotImplementedError("a Real-space component of qp0 must be a Constant, a number or a Function")
                x[int(off[E + 1 + k])] = cc.const_value()
    else:                                                    # primal model: [DG0(mesh) | CG1(mesh)]
        put(0, qp0[0], lambda cc: _cell_means(dn, cc))
        put(1, qp0[1], lambda cc: cc.eval_device(dn.coords))
    out.touch() if hasattr(out, "touch") else None
    return out


def time_stepping_stokes(model, t=Constant(0), t_steps=10, T=1, qp0=None, t_step_scheme="IE",
                         reassemble_lhs=True, progress=False, warm_start=True, check_every=0):
    """Time stepping for  1/A d/dt q + a(U, V) = L(V; f, g)  (time_stepping.py:126-221).
    Returns the list of `t_steps` solutions (ii_Function), initial state first.

    Beyond the reference: the whole history lives in ONE device allocation (the returned functions
    are views into it), steps are enqueued without host synchronisation, and the Krylov part of the
    solve (only present when the network has cycles) starts from the previous step's multipliers.
    `check_every=k` reads the solver status back every k steps (0: once, after the last step)."""
    import torch
    sys_ = model._system()
    dt = T / t_steps                                                  # :150
    cn, cn1 = time_stepping_schemes[t_step_scheme].values()           # :151
    ainv = model.Ainv

    L = model.L_form()                                                # :157
    qp0 = _as_state(model, qp0)                                       # :147-165
    hist = torch.empty((t_steps, sys_.N), dtype=torch.float64, device=qp0.tensor().device)
    hist[0].copy_(qp0.tensor())
    qps = [ii_Function(model.W, data=hist[0])]                        # :167
    x0 = hist[0]

    a_n = model.a_form().assemble() if cn != 0 else None              # An (:170) -- only CN multiplies by it
    Ln = L.assemble().tensor                                          # :170
    DDn = sys_.mass_apply(model._mass_coeff(ainv), x0)                # Dn assembled from qp0 (:155,170)
    t.assign(dt)                                                      # :172
    co_eff = model._coeffs(sigma=cn1 * dt, mass=ainv)                 # DDn1 + cn1*dt*An1 (:174,184)
    bcs = model.get_bc()
    a_n1 = a_n
    N = sys_.N
    b = x0.new_empty(N)
    ax0 = x0.new_empty(N) if cn != 0 else None

    steps = np.linspace(dt, T, t_steps - 1)                           # :176
    if progress:
        from tqdm import tqdm
        steps = tqdm(steps)
    one_call = sys_.dn.part is None and os.environ.get("GNX_TIMESTEP_ONE_CALL", "1") != "0"
    Ln_bufs = [x0.new_empty(N), x0.new_empty(N)] if one_call else None
    for k, t_val in enumerate(steps):
        if reassemble_lhs:                                            # :180-181
            co_eff = model._coeffs(sigma=cn1 * dt, mass=ainv)
            if cn != 0:
                a_n1 = model.a_form().assemble()
        sol = ii_Function(model.W, data=hist[k + 1])                  # :194-198
        last = k == len(steps) - 1
        check = last or (check_every > 0 and (k + 1) % check_every == 0)
        if one_call:
            # :179-192 and :216 behind ONE library call (gnx_timestep_*): L^{n+1}, the scheme's right-hand side, the
            # solve and the mass term of the next step
            Ln1 = Ln_bufs[k & 1]
            info = sys_.timestep(model, L, co_eff, cn1 * dt, cn * dt, a_n, x0, Ln, model._mass_coeff(ainv), DDn, Ln1, b, ax0,
                                 sol.tensor(), bcs, warm_start=warm_start and k > 0, sync=check)
        else:
            Ln1 = L.assemble().tensor                                 # :179
            if cn != 0:                                               # :185
                sys_.spmv(a_n.vals, x0, ax0)
            sys_.lincomb([1.0, cn1 * dt, -dt * cn, dt * cn], [DDn, Ln1, ax0, Ln if cn != 0 else None], out=b)
            _, info = sys_.solve_bc(co_eff, b, sol.tensor(), bcs, warm_start=warm_start and k > 0, sync=check)   # :187-192
        model.solve_info = info
        qps.append(sol)
        x0 = sol.tensor()                                             # :201
        t.assign(t_val + dt)                                          # :204
        for func in (model.p_bc, model.g, model.f):                   # :207-211
            try:
                func.t = t_val + dt
            except AttributeError:
                pass
        a_n, Ln = a_n1, Ln1                                           # :214-215
        if not one_call:
            sys_.mass_apply(model._mass_coeff(ainv), x0, out=DDn)     # :216
        L = model.L_form()                                            # :218-219
    # steps are enqueued without waiting for the solver's status; a step of a cyclic network whose Krylov part did not
    # converge leaves its mark on the device and is reported here, once, instead of passing silently
    bad = sys_.unconverged()
    if bad:
        from .._lib import GnxError
        raise GnxError(f"time_stepping_stokes: the Schur-complement CG did not converge in {bad} of {t_steps - 1} steps "
                       f"(check_every=k reports the first one earlier)")
    return qps
