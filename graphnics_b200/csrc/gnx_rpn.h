// RPN interpreter for coefficient expressions (device + TEST-ONLY host build).
// The reference takes f, g, p_bc as dolfin Constant / Expression("c string", degree, t=..) / UFL
// (graphnics/flowmodels/flow_models.py:155, tests/test_flow_models.py:94-119); the host compiles
// them to a gnx_program_t (graphnics_b200/coefficients.py) and every kernel that needs values
// interprets the program, one evaluation point per thread.  All threads run the same program, so
// op/arg/val are uniform loads (constant bank when the program is a kernel parameter).
#ifndef GNX_RPN_H
#define GNX_RPN_H
#include <math.h>
#include "gnx_index.h"

constexpr int GNX_RPN_STACK = 16;

GNX_HD double gnx_rpn_eval(const gnx_program_t& P, const double* x, int gd, const double* tau,
                           double edge_attr) {
  double st[GNX_RPN_STACK];
  int sp = 0;
  for (int i = 0; i < P.len; ++i) {
    const int op = P.op[i];
    const int a = P.arg[i];
    if (op < 10) {
      double v = 0.0;
      switch (op) {
        case GNX_OP_CONST: v = P.val[i]; break;
        case GNX_OP_X: v = (a < gd) ? x[a] : 0.0; break;
        case GNX_OP_PARAM: v = P.params[a]; break;
        case GNX_OP_TAU: v = (tau && a < gd) ? tau[a] : 0.0; break;
        case GNX_OP_EDGE: v = edge_attr; break;
      }
      if (sp < GNX_RPN_STACK) st[sp++] = v;
    } else if (op == GNX_OP_SELECT) {
      const double fv = st[--sp], tv = st[--sp], c = st[sp - 1];
      st[sp - 1] = (c != 0.0) ? tv : fv;
    } else if ((op >= GNX_OP_ADD && op <= GNX_OP_POW) || op == GNX_OP_MIN || op == GNX_OP_MAX ||
               op == GNX_OP_FMOD || op == GNX_OP_ATAN2 || (op >= GNX_OP_LT && op <= GNX_OP_GE)) {
      const double r = st[--sp], l = st[sp - 1];
      double v = 0.0;
      switch (op) {
        case GNX_OP_ADD: v = l + r; break;
        case GNX_OP_SUB: v = l - r; break;
        case GNX_OP_MUL: v = l * r; break;
        case GNX_OP_DIV: v = l / r; break;
        case GNX_OP_POW: v = pow(l, r); break;
        case GNX_OP_MIN: v = fmin(l, r); break;
        case GNX_OP_MAX: v = fmax(l, r); break;
        case GNX_OP_FMOD: v = fmod(l, r); break;
        case GNX_OP_ATAN2: v = atan2(l, r); break;
        case GNX_OP_LT: v = l < r; break;
        case GNX_OP_GT: v = l > r; break;
        case GNX_OP_LE: v = l <= r; break;
        case GNX_OP_GE: v = l >= r; break;
      }
      st[sp - 1] = v;
    } else {
      const double u = st[sp - 1];
      double v = u;
      switch (op) {
        case GNX_OP_NEG: v = -u; break;
        case GNX_OP_SIN: v = sin(u); break;
        case GNX_OP_COS: v = cos(u); break;
        case GNX_OP_TAN: v = tan(u); break;
        case GNX_OP_EXP: v = exp(u); break;
        case GNX_OP_LOG: v = log(u); break;
        case GNX_OP_SQRT: v = sqrt(u); break;
        case GNX_OP_ABS: v = fabs(u); break;
        case GNX_OP_TANH: v = tanh(u); break;
        case GNX_OP_SINH: v = sinh(u); break;
        case GNX_OP_COSH: v = cosh(u); break;
        case GNX_OP_ATAN: v = atan(u); break;
        case GNX_OP_ASIN: v = asin(u); break;
        case GNX_OP_ACOS: v = acos(u); break;
        case GNX_OP_FLOOR: v = floor(u); break;
        case GNX_OP_CEIL: v = ceil(u); break;
        case GNX_OP_SIGN: v = (u > 0.0) - (u < 0.0); break;
        case GNX_OP_POWI: {
          int e = a < 0 ? -a : a;
          double b = u; v = 1.0;
          while (e) { if (e & 1) v *= b; b *= b; e >>= 1; }
          if (a < 0) v = 1.0 / v;
        } break;
      }
      st[sp - 1] = v;
    }
  }
  return sp > 0 ? st[sp - 1] : 0.0;
}
#endif  // GNX_RPN_H
