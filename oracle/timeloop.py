"""Oracle, time loop: literal restatement of `time_stepping_stokes`
(graphnics/flowmodels/time_stepping.py:126-221), quirks included.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Two clocks exist in the reference and they tick differently:
  * "const": the UFL `Constant` t handed to time_stepping_stokes (`t.assign(...)`, :172,:204);
  * "attr" : the `.t` attribute of `Expression(..., t=0)` coefficients (:207-211), which is NOT
    touched by `t.assign(dt)` at :172 -- so Expression data lag one step behind UFL data.
For the mixed model f, g and the outlet p_bc are L2-projected when `L_form()` is CALLED
(flow_models.py:315-317), i.e. frozen at form-build time; the inlet p_bc term and everything
in the primal model is evaluated when the form is ASSEMBLED.
"""
import numpy as np

from . import fem

__all__ = ["TimeCoef", "time_stepping_schemes", "time_stepping"]

# time_stepping.py:17
time_stepping_schemes = {"IE": {"b1": 0, "b2": 1}, "CN": {"b1": 0.5, "b2": 0.5}}


class TimeCoef:
    """fn(x[...,gdim], t) -> [...]; clock in {"const","attr",None}; FEniCS `degree`."""

    def __init__(self, fn, degree=2, clock="const"):
        self.fn, self.degree, self.clock = fn, degree, clock

    def at(self, clocks):
        t = clocks.get(self.clock, 0.0)
        return fem.Coef(lambda x, t=t: self.fn(x, t), self.degree)


def _at(c, clocks):
    return c.at(clocks) if isinstance(c, TimeCoef) else fem.as_coef(c)


def time_stepping(kind, om, f=0.0, g=0.0, p_bc=0.0, Res=1.0, Ainv=1.0, t0=0.0,
                  t_steps=10, T=1.0, x0=None, scheme="IE", project=True):
    """kind in {"mixed","primal"}.  Returns the list of `t_steps` monolithic vectors
    (time_stepping.py:167,198,221)."""
    mixed = kind == "mixed"
    lay = fem.mixed_layout(om) if mixed else fem.primal_layout(om)
    N = lay["N"]
    x0 = np.zeros(N) if x0 is None else np.asarray(x0, dtype=np.float64).copy()

    dt = T / t_steps                                            # :150
    cn, cn1 = time_stepping_schemes[scheme].values()            # :151
    clocks = {"const": t0, "attr": t0}

    if mixed:
        A_mat = fem.assemble_mixed(om, Res)                     # a_form (Res from edge attrs)
        D = fem.assemble_mass_q_mixed(om, Ainv)                 # :154  (model.Ainv, global)
    else:
        A_mat = fem.assemble_primal(om, Res)
        D = fem.assemble_mass_q_primal(om, Ainv)

    def build_L():
        """model.L_form() -- returns an `assemble()` closure (frozen vs. live parts)."""
        if mixed:
            frozen = dict(clocks)
            def assemble():
                # projected data frozen at build time, inlet p_bc live at assembly time
                b_frozen = fem.assemble_mixed_rhs(om, _at(f, frozen), _at(g, frozen),
                                                  _at(p_bc, frozen), project=project)
                if isinstance(p_bc, TimeCoef) and frozen != clocks:
                    live = fem.assemble_mixed_rhs(om, 0.0, 0.0, _at(p_bc, clocks), project=False)
                    stale = fem.assemble_mixed_rhs(om, 0.0, 0.0, _at(p_bc, frozen), project=False)
                    e_in, l_in = np.nonzero(om.vf == fem.BOUN_IN)
                    idx = lay["q_off"][e_in] + l_in
                    b_frozen[idx] += live[idx] - stale[idx]
                return b_frozen
            return assemble
        return lambda: fem.assemble_primal_rhs(om, _at(f, clocks), _at(g, clocks))

    L = build_L()                                               # :157
    qps = [x0.copy()]                                           # :162-167
    An, Ln, DDn = A_mat, L(), D @ x0                            # :170
    clocks["const"] = dt                                        # :172
    An1, Ln1, DDn1 = A_mat, L(), D                              # :174

    for t_val in np.linspace(dt, T, t_steps - 1):               # :176
        Ln1 = L()                                               # :179
        A = (DDn1 + cn1 * dt * An1).tocsr()                     # :184
        b = DDn + cn1 * dt * Ln1 - dt * cn * (An @ x0) + dt * cn * Ln   # :185
        if not mixed:                                           # :188 apply_bc(model.get_bc())
            dofs, vals = fem.primal_bc(om, _at(p_bc, clocks))
            A, b = fem.apply_bc(A, b, dofs, vals)
        x1 = fem.lu_solve(A, b)                                 # :190-192
        qps.append(x1.copy())                                   # :195-198
        x0 = x1                                                 # :201
        clocks["const"] = t_val + dt                            # :204
        clocks["attr"] = t_val + dt                             # :207-211
        An, Ln, DDn = An1, Ln1, DDn1 @ x0                       # :214-216
        L = build_L()                                           # :218-219
    return qps
