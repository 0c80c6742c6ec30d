/*
 * siafd_oracle.cc -- TEST INFRASTRUCTURE ONLY (see siafd_oracle.h).
 *
 * Plain-array CPU restatement of PISM v1.2.1 stressbalance::SIAFD::update() and the
 * pieces of rheology / EnthalpyConverter / IceGrid / BedSmoother / Mask it calls.
 * Same multi-pass structure as the reference (gradient -> diffusivity with stored
 * delta columns -> flux -> I -> u,v), same operation order inside expressions, same
 * libm calls (exp, pow, sqrt, floor), same loop bounds including ghost rings.
 * Build WITHOUT -ffast-math and WITHOUT FMA contraction (-ffp-contract=off): the
 * reference's stock x86-64 build has neither.
 *
 * File:line citations refer to the reference tree (juliusgarbe/pism, v1.2.1).
 */
#include "siafd_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------
// EnthalpyConverter  (src/util/EnthalpyConverter.cc)
// ---------------------------------------------------------------------------------------
struct Converter {
  double p_air, g, beta, rho_i, c_i, c_w, L0, T_melting, T_0;

  explicit Converter(const orc_params &p)
      : p_air(p.ec_p_air), g(p.ec_g), beta(p.ec_beta), rho_i(p.ec_rho_i), c_i(p.ec_c_i), c_w(p.ec_c_w),
        L0(p.ec_L), T_melting(p.ec_T_melting), T_0(p.ec_T_0) {}

  // EnthalpyConverter.cc:137-143 (scalar form clamps depth < 0)
  double pressure(double depth) const {
    if (depth >= 0.0) {
      return p_air + rho_i * g * depth;
    }
    return p_air;
  }
  // EnthalpyConverter.cc:146-152 (column form, NO clamp)
  void pressure(const std::vector<double> &depth, unsigned int ks, std::vector<double> &result) const {
    for (unsigned int k = 0; k <= ks; ++k) {
      result[k] = p_air + rho_i * g * depth[k];
    }
  }
  // :158-160
  double melting_temperature(double P) const { return T_melting - beta * P; }
  // :365-367
  double L(double T_pm) const { return L0 + (c_w - c_i) * (T_pm - 273.15); }
  // :378-380
  double enthalpy_cts(double P) const { return c_i * (melting_temperature(P) - T_0); }
  // :383-385
  double enthalpy_cold(double T) const { return c_i * (T - T_0); }
  // :388-390
  double temperature_cold(double E) const { return (E / c_i) + T_0; }
  // :180-188
  double temperature(double E, double P) const {
    if (E < enthalpy_cts(P)) {
      return temperature_cold(E);
    }
    return melting_temperature(P);
  }
  // :196-198
  double pressure_adjusted_temperature(double E, double P) const {
    return temperature(E, P) - melting_temperature(P) + T_melting;
  }
  // :214-223
  double water_fraction(double E, double P) const {
    double E_s = enthalpy_cts(P);
    if (E <= E_s) {
      return 0.0;
    }
    return (E - E_s) / L(melting_temperature(P));
  }
  // :243-253
  double enthalpy(double T, double omega, double P) const {
    const double T_m = melting_temperature(P);
    if (T < T_m) {
      return enthalpy_cold(T);
    }
    return enthalpy_cts(P) + omega * L(T_m);
  }
  // :277-285
  double enthalpy_permissive(double T, double omega, double P) const {
    const double T_m = melting_temperature(P);
    if (T < T_m) {
      return enthalpy(T, 0.0, P);
    }
    return enthalpy(T_m, std::max(0.0, std::min(omega, 1.0)), P);
  }
};

// ---------------------------------------------------------------------------------------
// Flow laws  (src/rheology/*.cc).  The class tree and the virtual dispatch per level are
// kept on purpose: they are part of what the CPU baseline costs.
// ---------------------------------------------------------------------------------------
class Law {
public:
  Law(const orc_params &p, const Converter &ec)
      : m_ec(ec), m_n(p.fl_n), m_A_cold(p.fl_A_cold), m_A_warm(p.fl_A_warm), m_Q_cold(p.fl_Q_cold),
        m_Q_warm(p.fl_Q_warm), m_crit_temp(p.fl_T_crit), m_R(p.fl_R), m_rho(p.fl_rho), m_g(p.fl_g) {
    // FlowLaw.cc:45: beta_CC_grad = beta * rho * g
    m_beta_CC_grad = p.fl_beta * m_rho * m_g;
  }
  virtual ~Law() {}

  // FlowLaw.cc:97-100
  double flow(double stress, double E, double pressure, double gs) const {
    return this->flow_impl(stress, E, pressure, gs);
  }
  // FlowLaw.cc:107-119
  void flow_n(const double *stress, const double *E, const double *pressure, const double *gs, unsigned int n,
              double *result) const {
    for (unsigned int k = 0; k < n; ++k) {
      result[k] = this->flow(stress[k], E[k], pressure[k], gs[k]);
    }
  }
  double softness(double E, double p) const { return this->softness_impl(E, p); }
  double exponent() const { return m_n; }

protected:
  // FlowLaw.cc:102-105
  virtual double flow_impl(double stress, double E, double pressure, double /*gs*/) const {
    return softness(E, pressure) * pow(stress, m_n - 1);
  }
  virtual double softness_impl(double E, double p) const = 0;

  // FlowLaw.cc:89-94
  double softness_paterson_budd(double T_pa) const {
    const double A = T_pa < m_crit_temp ? m_A_cold : m_A_warm;
    const double Q = T_pa < m_crit_temp ? m_Q_cold : m_Q_warm;
    return A * exp(-Q / (m_R * T_pa));
  }

  const Converter &m_ec;
  double m_n, m_A_cold, m_A_warm, m_Q_cold, m_Q_warm, m_crit_temp, m_R, m_rho, m_g, m_beta_CC_grad;
};

// PatersonBudd.cc:41-59
class LawPB : public Law {
public:
  using Law::Law;

protected:
  double softness_impl(double E, double pressure) const override {
    double T_pa = m_ec.pressure_adjusted_temperature(E, pressure);
    return softness_from_temp(T_pa);
  }
  double flow_impl(double stress, double E, double pressure, double gs) const override {
    double temp = m_ec.temperature(E, pressure);
    return flow_from_temp(stress, temp, pressure, gs);
  }
  virtual double flow_from_temp(double stress, double temp, double pressure, double /*gs*/) const {
    const double T_pa = temp + (m_beta_CC_grad / (m_rho * m_g)) * pressure;
    return softness_from_temp(T_pa) * pow(stress, m_n - 1);
  }
  virtual double softness_from_temp(double T_pa) const { return softness_paterson_budd(T_pa); }
};

// PatersonBuddCold.cc:43-51 ("arr")
class LawArr : public LawPB {
public:
  using LawPB::LawPB;

protected:
  double softness_from_temp(double T_pa) const override { return m_A_cold * exp(-m_Q_cold / (m_R * T_pa)); }
  double flow_from_temp(double stress, double temp, double, double) const override {
    return softness_from_temp(temp) * pow(stress, m_n - 1);
  }
};

// PatersonBuddWarm.cc:42-50 ("arrwarm")
class LawArrWarm : public LawPB {
public:
  using LawPB::LawPB;

protected:
  double softness_from_temp(double T_pa) const override { return m_A_warm * exp(-m_Q_warm / (m_R * T_pa)); }
  double flow_from_temp(double stress, double temp, double, double) const override {
    return softness_from_temp(temp) * pow(stress, m_n - 1);
  }
};

// IsothermalGlen.cc:37-43
class LawIso : public LawPB {
public:
  LawIso(const orc_params &p, const Converter &ec) : LawPB(p, ec), m_softness_A(p.iso_softness_A) {}

protected:
  double flow_impl(double stress, double, double, double) const override {
    return m_softness_A * pow(stress, m_n - 1);
  }
  double softness_impl(double, double) const override { return m_softness_A; }
  double m_softness_A;
};

// Hooke.cc:46-49 (flow goes through LawPB::flow_impl / flow_from_temp)
class LawHooke : public LawPB {
public:
  LawHooke(const orc_params &p, const Converter &ec)
      : LawPB(p, ec), m_Q(p.hooke_Q), m_A(p.hooke_A), m_C(p.hooke_C), m_K(p.hooke_K), m_Tr(p.hooke_Tr) {}

protected:
  double softness_from_temp(double T_pa) const override {
    return m_A * exp(-m_Q / (m_R * T_pa) + 3.0 * m_C * pow(m_Tr - T_pa, -m_K));
  }
  double m_Q, m_A, m_C, m_K, m_Tr;
};

// GPBLD.cc:49-61
class LawGPBLD : public Law {
public:
  LawGPBLD(const orc_params &p, const Converter &ec)
      : Law(p, ec), m_T_0(p.gpbld_T_0), m_water_frac_coeff(p.gpbld_water_frac_coeff),
        m_water_frac_observed_limit(p.gpbld_water_frac_limit) {}

protected:
  double softness_impl(double enthalpy, double pressure) const override {
    const double E_s = m_ec.enthalpy_cts(pressure);
    if (enthalpy < E_s) {
      double T_pa = m_ec.pressure_adjusted_temperature(enthalpy, pressure);
      return softness_paterson_budd(T_pa);
    }
    double omega = m_ec.water_fraction(enthalpy, pressure);
    omega = std::min(omega, m_water_frac_observed_limit);
    return softness_paterson_budd(m_T_0) * (1.0 + m_water_frac_coeff * omega);
  }
  double m_T_0, m_water_frac_coeff, m_water_frac_observed_limit;
};

// GoldsbyKohlstedt.cc:31-72, :114-150
class LawGK : public Law {
public:
  using Law::Law;

protected:
  double softness_impl(double, double) const override { return NAN; /* reference throws: :102-108 */ }
  double flow_impl(double stress, double E, double pressure, double grainsize) const override {
    double temp = m_ec.temperature(E, pressure);
    return flow_from_temp(stress, temp, pressure, grainsize);
  }
  double flow_from_temp(double stress, double temp, double pressure, double gs) const {
    const double V_act_vol = -13.e-6, disl_crit_temp = 258.0, disl_A_cold = 4.0e-19, disl_A_warm = 6.0e4,
                 disl_n = 4.0, disl_Q_cold = 60.e3, disl_Q_warm = 180.e3, gbs_crit_temp = 255.0,
                 gbs_A_cold = 6.1811e-14, gbs_A_warm = 4.7547e15, gbs_n = 1.8, gbs_Q_cold = 49.e3,
                 gbs_Q_warm = 192.e3, p_grain_sz_exp = 1.4, basal_A = 2.1896e-7, basal_n = 2.4, basal_Q = 60.e3,
                 diff_crit_temp = 258.0, diff_V_m = 1.97e-5, diff_D_0v = 9.10e-4, diff_Q_v = 59.4e3,
                 diff_D_0b = 5.8e-4, diff_Q_b = 49.e3, diff_delta = 9.04e-10;
    double eps_diff, eps_disl, eps_basal, eps_gbs, diff_D_b;

    if (fabs(stress) < 1e-10) {
      return 0;
    }
    const double T = temp + (m_beta_CC_grad / (m_rho * m_g)) * pressure;
    const double pV = pressure * V_act_vol;
    const double RT = m_R * T;
    const double diff_D_v = diff_D_0v * exp(-diff_Q_v / RT);
    diff_D_b = diff_D_0b * exp(-diff_Q_b / RT);
    if (T > diff_crit_temp) {
      diff_D_b *= 1000;
    }
    eps_diff = 42 * diff_V_m * (diff_D_v + M_PI * diff_delta * diff_D_b / gs) / (RT * (gs * gs));
    if (T > disl_crit_temp) {
      eps_disl = disl_A_warm * pow(stress, disl_n - 1) * exp(-(disl_Q_warm + pV) / RT);
    } else {
      eps_disl = disl_A_cold * pow(stress, disl_n - 1) * exp(-(disl_Q_cold + pV) / RT);
    }
    eps_basal = basal_A * pow(stress, basal_n - 1) * exp(-(basal_Q + pV) / RT);
    if (T > gbs_crit_temp) {
      eps_gbs = gbs_A_warm * (pow(stress, gbs_n - 1) / pow(gs, p_grain_sz_exp)) * exp(-(gbs_Q_warm + pV) / RT);
    } else {
      eps_gbs = gbs_A_cold * (pow(stress, gbs_n - 1) / pow(gs, p_grain_sz_exp)) * exp(-(gbs_Q_cold + pV) / RT);
    }
    return eps_diff + eps_disl + (eps_basal * eps_gbs) / (eps_basal + eps_gbs);
  }
};

// FlowLawFactory.cc:71-90
std::unique_ptr<Law> make_law(const orc_params &p, const Converter &ec) {
  switch (p.flow_law) {
  case ORC_FLOW_ISOTHERMAL_GLEN:
    return std::unique_ptr<Law>(new LawIso(p, ec));
  case ORC_FLOW_PB:
    return std::unique_ptr<Law>(new LawPB(p, ec));
  case ORC_FLOW_GPBLD:
    return std::unique_ptr<Law>(new LawGPBLD(p, ec));
  case ORC_FLOW_HOOKE:
    return std::unique_ptr<Law>(new LawHooke(p, ec));
  case ORC_FLOW_ARR:
    return std::unique_ptr<Law>(new LawArr(p, ec));
  case ORC_FLOW_ARRWARM:
    return std::unique_ptr<Law>(new LawArrWarm(p, ec));
  case ORC_FLOW_GK:
    return std::unique_ptr<Law>(new LawGK(p, ec));
  default:
    return nullptr;
  }
}

// grain_size_vostok.cc:28-60: linear interpolation (gsl_interp_linear) in a 22-point table,
// argument clamped to the table range.
const int kVostokN = 22;
const double kVostokAge[kVostokN] = {0.0000e+00, 5.0000e+01, 1.0000e+02, 1.2500e+02, 1.5000e+02, 1.5800e+02,
                                     1.6500e+02, 1.7000e+02, 1.8000e+02, 1.8800e+02, 2.0000e+02, 2.2500e+02,
                                     2.4500e+02, 2.6000e+02, 3.0000e+02, 3.2000e+02, 3.5000e+02, 4.0000e+02,
                                     5.0000e+02, 6.0000e+02, 8.0000e+02, 1.0000e+04};
const double kVostokGs[kVostokN] = {1.8000e-03, 2.2000e-03, 3.0000e-03, 4.0000e-03, 4.3000e-03, 3.0000e-03,
                                    3.0000e-03, 4.6000e-03, 3.4000e-03, 3.3000e-03, 5.9000e-03, 6.2000e-03,
                                    5.4000e-03, 6.8000e-03, 3.5000e-03, 6.0000e-03, 8.0000e-03, 8.3000e-03,
                                    3.6000e-03, 3.8000e-03, 9.5000e-03, 1.0000e-02};

// GSL gsl_interp_bsearch(x_array, x, 0, len-1): largest i in [0, len-2] with xa[i] <= x (0 if none).
int gsl_style_find(const double *xa, int len, double x) {
  int ilo = 0, ihi = len - 1;
  while (ihi > ilo + 1) {
    int i = (ihi + ilo) / 2;
    if (xa[i] > x) {
      ihi = i;
    } else {
      ilo = i;
    }
  }
  return ilo;
}

double vostok(double age_years) {
  double a = age_years / 1000.0;
  a = std::max(a, kVostokAge[0]);
  a = std::min(a, kVostokAge[kVostokN - 1]);
  int i = gsl_style_find(kVostokAge, kVostokN, a);
  // gsl linear interpolation: y_lo + (x - x_lo) / (x_hi - x_lo) * (y_hi - y_lo)
  double x_lo = kVostokAge[i], x_hi = kVostokAge[i + 1], y_lo = kVostokGs[i], y_hi = kVostokGs[i + 1];
  double dx = x_hi - x_lo;
  return y_lo + (a - x_lo) / dx * (y_hi - y_lo);
}

// ---------------------------------------------------------------------------------------
// Local ghosted array views ([j][i][dof], src/util/IceModelVec_inline.hh:28-40)
// ---------------------------------------------------------------------------------------
struct View {
  double *a;
  int xs, ys, xm, ym, w, dof;
  View(const double *ptr, const orc_params &p, int width, int ndof)
      : a(const_cast<double *>(ptr)), xs(p.xs), ys(p.ys), xm(p.xm), ym(p.ym), w(width), dof(ndof) {}
  inline long idx(int i, int j) const { return ((long)(j - (ys - w)) * (xm + 2 * w) + (i - (xs - w))) * dof; }
  inline double &operator()(int i, int j) const { return a[idx(i, j)]; }
  inline double &operator()(int i, int j, int k) const { return a[idx(i, j) + k]; }
  inline double *column(int i, int j) const { return a + idx(i, j); }
  long size() const { return (long)(xm + 2 * w) * (ym + 2 * w) * dof; }
  void fill(double v) const { std::fill(a, a + size(), v); }
};

// Mask.hh:37-66 with IceModelVec_inline.hh:95-101 decoding
struct MaskView {
  View v;
  MaskView(const double *ptr, const orc_params &p) : v(ptr, p, p.w_geom, 1) {}
  int as_int(int i, int j) const { return static_cast<int>(floor(v(i, j) + 0.5)); }
  static bool ocean(int M) { return M >= 3; }
  static bool grounded(int M) { return not ocean(M); }
  static bool icy(int M) { return (M == 2) || (M == 3); }
  static bool ice_free(int M) { return not icy(M); }
  bool grounded(int i, int j) const { return grounded(as_int(i, j)); }
  bool icy(int i, int j) const { return icy(as_int(i, j)); }
  bool ice_free(int i, int j) const { return ice_free(as_int(i, j)); }
  bool floating_ice(int i, int j) const {
    int M = as_int(i, j);
    return icy(M) && ocean(M);
  }
  bool ice_free_ocean(int i, int j) const {
    int M = as_int(i, j);
    return ocean(M) && ice_free(M);
  }
};

// IceGrid.cc:427-440
int k_below_height(const double *z, int Mz, double height, int *status) {
  if (height < 0.0 - 1.0e-6) {
    *status = ORC_ERR_HEIGHT_BELOW_BASE;
    return 0;
  }
  if (height > z[Mz - 1] + 1.0e-6) {
    *status = ORC_ERR_HEIGHT_ABOVE_TOP;
    return 0;
  }
  return gsl_style_find(z, Mz, height);
}

// ---------------------------------------------------------------------------------------
// BedSmoother  (src/stressbalance/sia/BedSmoother.cc)
// ---------------------------------------------------------------------------------------
int smoothed_thk(const orc_params &p, const orc_fields &f, double *out) {
  // BedSmoother.cc:284-327
  const int G = p.w_geom;
  View result(out, p, G, 1), thk(f.thickness, p, G, 1), usurf(f.surface, p, G, 1), maxtl(f.maxtl, p, G, 1),
      topgsmooth(f.topgsmooth, p, G, 1);
  MaskView mask(f.mask, p);
  for (int j = p.ys - G; j < p.ys + p.ym + G; ++j) {
    for (int i = p.xs - G; i < p.xs + p.xm + G; ++i) {
      if (thk(i, j) < 0.0) {
        return ORC_ERR_NEGATIVE_THICKNESS;
      } else if (thk(i, j) == 0.0) {
        result(i, j) = 0.0;
      } else if (maxtl(i, j) >= thk(i, j)) {
        result(i, j) = thk(i, j);
      } else {
        if (mask.grounded(i, j)) {
          const double thks_try = usurf(i, j) - topgsmooth(i, j);
          result(i, j) = (thks_try > 0.0) ? thks_try : 0.0;
        } else {
          result(i, j) = thk(i, j);
        }
      }
    }
  }
  return ORC_OK;
}

int theta(const orc_params &p, const orc_fields &f, double *out) {
  // BedSmoother.cc:351-404
  const int G = p.w_geom;
  View result(out, p, G, 1);
  if (not f.smoother_active) {
    result.fill(1.0);
    return ORC_OK;
  }
  View usurf(f.surface, p, G, 1), maxtl(f.maxtl, p, G, 1), topgsmooth(f.topgsmooth, p, G, 1), C2(f.C2, p, G, 1),
      C3(f.C3, p, G, 1), C4(f.C4, p, G, 1);
  const double theta_min = p.theta_min, theta_max = 1.0;
  for (int j = p.ys - G; j < p.ys + p.ym + G; ++j) {
    for (int i = p.xs - G; i < p.xs + p.xm + G; ++i) {
      const double H = usurf(i, j) - topgsmooth(i, j);
      if (H > maxtl(i, j)) {
        const double Hinv = 1.0 / std::max(H, 1.0);
        double omega = 1.0 + Hinv * Hinv * (C2(i, j) + Hinv * (C3(i, j) + Hinv * C4(i, j)));
        if (omega <= 0) {
          return ORC_ERR_OMEGA_NEGATIVE;
        }
        if (omega < 0.001) {
          omega = 0.001;
        }
        result(i, j) = pow(omega, -p.fl_n);
      } else {
        result(i, j) = 0.00;
      }
      // clip(x, a, b) = min(max(a, x), b): src/util/pism_utilities.hh:91-93
      result(i, j) = std::min(std::max(theta_min, result(i, j)), theta_max);
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------
// Surface gradients  (SIAFD.cc:224-500)
// ---------------------------------------------------------------------------------------
void gradient_mahaffy(const orc_params &p, orc_fields &f) {
  // SIAFD.cc:298-324
  const double dx = p.dx, dy = p.dy;
  View h(f.surface, p, p.w_geom, 1), h_x(f.h_x, p, p.w_stag, 2), h_y(f.h_y, p, p.w_stag, 2);
  for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
    for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
      h_x(i, j, 0) = (h(i + 1, j) - h(i, j)) / dx;
      h_y(i, j, 0) = (+h(i + 1, j + 1) + h(i, j + 1) - h(i + 1, j - 1) - h(i, j - 1)) / (4.0 * dy);
      h_y(i, j, 1) = (h(i, j + 1) - h(i, j)) / dy;
      h_x(i, j, 1) = (+h(i + 1, j + 1) + h(i + 1, j) - h(i - 1, j + 1) - h(i - 1, j)) / (4.0 * dx);
    }
  }
}

void gradient_eta(const orc_params &p, orc_fields &f) {
  // SIAFD.cc:224-293
  const double n = p.fl_n, etapow = (2.0 * n + 2.0) / n, invpow = 1.0 / etapow,
               dinvpow = (-n - 2.0) / (2.0 * n + 2.0);
  const double dx = p.dx, dy = p.dy;
  const int G = p.w_geom;
  View eta(f.work2d_0, p, G, 1), H(f.thickness, p, G, 1), b(f.bed, p, G, 1), h_x(f.h_x, p, p.w_stag, 2),
      h_y(f.h_y, p, p.w_stag, 2);
  for (int j = p.ys - G; j < p.ys + p.ym + G; ++j) {
    for (int i = p.xs - G; i < p.xs + p.xm + G; ++i) {
      eta(i, j) = pow(H(i, j), etapow);
    }
  }
  for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
    for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
      // box stencil names: src/util/iceModelVec.hh BoxStencil (ij, n, nw, w, sw, s, se, e, ne)
      const double e_ij = eta(i, j), e_e = eta(i + 1, j), e_w = eta(i - 1, j), e_n = eta(i, j + 1),
                   e_s = eta(i, j - 1), e_ne = eta(i + 1, j + 1), e_nw = eta(i - 1, j + 1),
                   e_se = eta(i + 1, j - 1);
      const double b_ij = b(i, j), b_e = b(i + 1, j), b_w = b(i - 1, j), b_n = b(i, j + 1), b_s = b(i, j - 1),
                   b_ne = b(i + 1, j + 1), b_nw = b(i - 1, j + 1), b_se = b(i + 1, j - 1);
      {
        double mean_eta = 0.5 * (e_e + e_ij);
        if (mean_eta > 0.0) {
          double factor = invpow * pow(mean_eta, dinvpow);
          h_x(i, j, 0) = factor * (e_e - e_ij) / dx;
          h_y(i, j, 0) = factor * (e_ne + e_n - e_se - e_s) / (4.0 * dy);
        } else {
          h_x(i, j, 0) = 0.0;
          h_y(i, j, 0) = 0.0;
        }
        h_x(i, j, 0) += (b_e - b_ij) / dx;
        h_y(i, j, 0) += (b_ne + b_n - b_se - b_s) / (4.0 * dy);
      }
      {
        double mean_eta = 0.5 * (e_n + e_ij);
        if (mean_eta > 0.0) {
          double factor = invpow * pow(mean_eta, dinvpow);
          h_x(i, j, 1) = factor * (e_ne + e_e - e_nw - e_w) / (4.0 * dx);
          h_y(i, j, 1) = factor * (e_n - e_ij) / dy;
        } else {
          h_x(i, j, 1) = 0.0;
          h_y(i, j, 1) = 0.0;
        }
        h_x(i, j, 1) += (b_ne + b_e - b_nw - b_w) / (4.0 * dx);
        h_y(i, j, 1) += (b_n - b_ij) / dy;
      }
    }
  }
}

void gradient_haseloff(const orc_params &p, orc_fields &f) {
  // SIAFD.cc:373-496 (the ghost exchange at :498-499 is the caller's job)
  const double dx = p.dx, dy = p.dy;
  const int G = p.w_geom;
  View h(f.surface, p, G, 1), w_i(f.work2d_0, p, G, 1), w_j(f.work2d_1, p, G, 1), h_x(f.h_x, p, p.w_stag, 2),
      h_y(f.h_y, p, p.w_stag, 2);
  MaskView mask(f.mask, p);

  for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
    for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
      // x-derivative, i-offset
      if ((mask.floating_ice(i, j) && mask.ice_free_ocean(i + 1, j)) ||
          (mask.ice_free_ocean(i, j) && mask.floating_ice(i + 1, j))) {
        h_x(i, j, 0) = 0;
        w_i(i, j) = 0;
      } else if ((mask.icy(i, j) && mask.ice_free(i + 1, j) && h(i + 1, j) > h(i, j)) ||
                 (mask.ice_free(i, j) && mask.icy(i + 1, j) && h(i, j) > h(i + 1, j))) {
        h_x(i, j, 0) = 0.0;
        w_i(i, j) = 0;
      } else {
        h_x(i, j, 0) = (h(i + 1, j) - h(i, j)) / dx;
        w_i(i, j) = 1;
      }
      // y-derivative, j-offset
      if ((mask.floating_ice(i, j) && mask.ice_free_ocean(i, j + 1)) ||
          (mask.ice_free_ocean(i, j) && mask.floating_ice(i, j + 1))) {
        h_y(i, j, 1) = 0.0;
        w_j(i, j) = 0.0;
      } else if ((mask.icy(i, j) && mask.ice_free(i, j + 1) && h(i, j + 1) > h(i, j)) ||
                 (mask.ice_free(i, j) && mask.icy(i, j + 1) && h(i, j) > h(i, j + 1))) {
        h_y(i, j, 1) = 0.0;
        w_j(i, j) = 0.0;
      } else {
        h_y(i, j, 1) = (h(i, j + 1) - h(i, j)) / dy;
        w_j(i, j) = 1.0;
      }
    }
  }

  for (int j = p.ys; j < p.ys + p.ym; ++j) {
    for (int i = p.xs; i < p.xs + p.xm; ++i) {
      // x-derivative, j-offset
      if (w_j(i, j) > 0) {
        double W = w_i(i, j) + w_i(i - 1, j) + w_i(i - 1, j + 1) + w_i(i, j + 1);
        if (W > 0) {
          h_x(i, j, 1) = 1.0 / W * (h_x(i, j, 0) + h_x(i - 1, j, 0) + h_x(i - 1, j + 1, 0) + h_x(i, j + 1, 0));
        } else {
          h_x(i, j, 1) = 0.0;
        }
      } else {
        if (mask.icy(i, j)) {
          double W = w_i(i, j) + w_i(i - 1, j);
          if (W > 0) {
            h_x(i, j, 1) = 1.0 / W * (h_x(i, j, 0) + h_x(i - 1, j, 0));
          } else {
            h_x(i, j, 1) = 0.0;
          }
        } else {
          double W = w_i(i, j + 1) + w_i(i - 1, j + 1);
          if (W > 0) {
            h_x(i, j, 1) = 1.0 / W * (h_x(i - 1, j + 1, 0) + h_x(i, j + 1, 0));
          } else {
            h_x(i, j, 1) = 0.0;
          }
        }
      }
      // y-derivative, i-offset
      if (w_i(i, j) > 0) {
        double W = w_j(i, j) + w_j(i, j - 1) + w_j(i + 1, j - 1) + w_j(i + 1, j);
        if (W > 0) {
          h_y(i, j, 0) = 1.0 / W * (h_y(i, j, 1) + h_y(i, j - 1, 1) + h_y(i + 1, j - 1, 1) + h_y(i + 1, j, 1));
        } else {
          h_y(i, j, 0) = 0.0;
        }
      } else {
        if (mask.icy(i, j)) {
          double W = w_j(i, j) + w_j(i, j - 1);
          if (W > 0) {
            h_y(i, j, 0) = 1.0 / W * (h_y(i, j, 1) + h_y(i, j - 1, 1));
          } else {
            h_y(i, j, 0) = 0.0;
          }
        } else {
          double W = w_j(i + 1, j - 1) + w_j(i + 1, j);
          if (W > 0) {
            h_y(i, j, 0) = 1.0 / W * (h_y(i + 1, j - 1, 1) + h_y(i + 1, j, 1));
          } else {
            h_y(i, j, 0) = 0.0;
          }
        }
      }
    }
  }
}

// SIAFD.cc:951-961
bool interglacial(const orc_params &p, double accumulation_time) {
  if (accumulation_time < p.eemian_start) {
    return false;
  } else if (accumulation_time < p.eemian_end) {
    return true;
  } else if (accumulation_time < p.holocene_start) {
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// compute_diffusivity  (SIAFD.cc:543-770)
// ---------------------------------------------------------------------------------------
int compute_diffusivity(const orc_params &p, orc_fields &f, bool full_update, const Law &law,
                        const Converter &ec) {
  const int G = p.w_geom;
  View thk_smooth(f.work2d_0, p, G, 1), theta_v(f.work2d_1, p, G, 1), h_x(f.h_x, p, p.w_stag, 2),
      h_y(f.h_y, p, p.w_stag, 2), result(f.D, p, p.w_stag, 2), enthalpy(f.enthalpy, p, p.w_3d_in, p.Mz),
      age(f.age, p, p.w_3d_in, p.Mz);
  View delta[2] = {View(f.delta_0, p, p.w_stag, p.Mz), View(f.delta_1, p, p.w_stag, p.Mz)};

  result.fill(0.0); // :561

  const double current_time = f.current_time, enhancement_factor = p.fl_e,
               enhancement_factor_interglacial = p.fl_e_interglacial, D_limit = p.D_limit;
  const bool compute_grain_size_using_age = p.grain_size_age_coupling != 0, e_age_coupling = p.e_age_coupling != 0,
             limit_diffusivity = p.limit_diffusivity != 0, use_age = compute_grain_size_using_age or e_age_coupling;

  if (use_age and f.age == nullptr) {
    return ORC_ERR_BAD_CONFIG;
  }

  int status = theta(p, f, f.work2d_1); // :580
  if (status != ORC_OK) {
    return status;
  }
  status = smoothed_thk(p, f, f.work2d_0); // :582
  if (status != ORC_OK) {
    return status;
  }

  const double *z = p.z;
  const int Mx = p.Mx, My = p.My, Mz = p.Mz;

  std::vector<double> depth(Mz), stress(Mz), pressure(Mz), E(Mz), flow(Mz);
  std::vector<double> delta_ij(Mz);
  std::vector<double> A(Mz), ice_grain_size(Mz, p.grain_size);
  std::vector<double> e_factor(Mz, enhancement_factor);

  double D_max = 0.0;
  int high_diffusivity_counter = 0;
  for (int o = 0; o < 2; o++) {
    for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
      for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
        const int oi = 1 - o, oj = o;

        const double thk = 0.5 * (thk_smooth(i, j) + thk_smooth(i + oi, j + oj));

        if (thk == 0.0) { // :631-637
          result(i, j, o) = 0.0;
          if (full_update) {
            std::fill(delta[o].column(i, j), delta[o].column(i, j) + Mz, 0.0);
          }
          continue;
        }

        const int ks = k_below_height(z, Mz, thk, &status);
        if (status != ORC_OK) {
          return status;
        }

        for (int k = 0; k <= ks; ++k) {
          depth[k] = thk - z[k];
        }

        ec.pressure(depth, ks, pressure);

        if (use_age) {
          const double *age_ij = age.column(i, j), *age_offset = age.column(i + oi, j + oj);
          for (int k = 0; k <= ks; ++k) {
            A[k] = 0.5 * (age_ij[k] + age_offset[k]);
          }
          if (compute_grain_size_using_age) {
            for (int k = 0; k <= ks; ++k) {
              ice_grain_size[k] = vostok(A[k] * p.years_per_second);
            }
          }
          if (e_age_coupling) {
            for (int k = 0; k <= ks; ++k) {
              const double accumulation_time = current_time - A[k];
              if (interglacial(p, accumulation_time)) {
                e_factor[k] = enhancement_factor_interglacial;
              } else {
                e_factor[k] = enhancement_factor;
              }
            }
          }
        }

        {
          const double *E_ij = enthalpy.column(i, j), *E_offset = enthalpy.column(i + oi, j + oj);
          for (int k = 0; k <= ks; ++k) {
            E[k] = 0.5 * (E_ij[k] + E_offset[k]);
          }
        }

        const double hx = h_x(i, j, o), hy = h_y(i, j, o);
        const double alpha = sqrt(hx * hx + hy * hy); // PetscSqr(x) = x*x
        for (int k = 0; k <= ks; ++k) {
          stress[k] = alpha * pressure[k];
        }

        law.flow_n(&stress[0], &E[0], &pressure[0], &ice_grain_size[0], ks + 1, &flow[0]);

        const double theta_local = 0.5 * (theta_v(i, j) + theta_v(i + oi, j + oj));
        for (int k = 0; k <= ks; ++k) {
          delta_ij[k] = e_factor[k] * theta_local * 2.0 * pressure[k] * flow[k];
        }

        double D = 0.0;
        {
          for (int k = 1; k <= ks; ++k) {
            const double dz = z[k] - z[k - 1];
            D += 0.5 * dz * ((depth[k] + dz) * delta_ij[k - 1] + depth[k] * delta_ij[k]);
          }
          const double dz = thk - z[ks];
          D += 0.5 * dz * dz * delta_ij[ks];
        }

        if (i < 0 || i >= Mx - 1 || j < 0 || j >= My - 1) { // :719-722
          D = 0.0;
        }

        if (limit_diffusivity and D >= D_limit) {
          D = D_limit;
          high_diffusivity_counter += 1;
        }

        D_max = std::max(D_max, D);

        result(i, j, o) = D;

        if (full_update) {
          for (int k = ks + 1; k < Mz; ++k) {
            delta_ij[k] = 0.0;
          }
          std::memcpy(delta[o].column(i, j), &delta_ij[0], sizeof(double) * Mz);
        }
      }
    }
  }

  f.D_max = D_max; // GlobalMax over ranks is the caller's job (:748)
  f.high_diffusivity_counter = high_diffusivity_counter;

  if (f.D_max > D_limit) { // :752-760 (per rank here; callers reduce first when multi-patch)
    return ORC_ERR_DIFFUSIVITY;
  }
  return ORC_OK;
}

// SIAFD.cc:772-793
void compute_diffusive_flux(const orc_params &p, orc_fields &f) {
  View h_x(f.h_x, p, p.w_stag, 2), h_y(f.h_y, p, p.w_stag, 2), D(f.D, p, p.w_stag, 2), Q(f.Q, p, p.w_stag, 2);
  for (int o = 0; o < 2; o++) {
    for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
      for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
        const double slope = (o == 0) ? h_x(i, j, o) : h_y(i, j, o);
        Q(i, j, o) = -D(i, j, o) * slope;
      }
    }
  }
}

// SIAFD.cc:807-870
int compute_I(const orc_params &p, orc_fields &f) {
  const int G = p.w_geom, Mz = p.Mz;
  View thk_smooth(f.work2d_0, p, G, 1);
  View I[2] = {View(f.I_0, p, p.w_stag, Mz), View(f.I_1, p, p.w_stag, Mz)};
  View delta[2] = {View(f.delta_0, p, p.w_stag, Mz), View(f.delta_1, p, p.w_stag, Mz)};

  int status = smoothed_thk(p, f, f.work2d_0);
  if (status != ORC_OK) {
    return status;
  }

  std::vector<double> dz(Mz);
  for (int k = 1; k < Mz; ++k) {
    dz[k] = p.z[k] - p.z[k - 1];
  }

  for (int o = 0; o < 2; ++o) {
    for (int j = p.ys - 1; j < p.ys + p.ym + 1; ++j) {
      for (int i = p.xs - 1; i < p.xs + p.xm + 1; ++i) {
        const int oi = 1 - o, oj = o;
        const double thk = 0.5 * (thk_smooth(i, j) + thk_smooth(i + oi, j + oj));
        const double *delta_ij = delta[o].column(i, j);
        double *I_ij = I[o].column(i, j);

        const int ks = k_below_height(p.z, Mz, thk, &status);
        if (status != ORC_OK) {
          return status;
        }

        I_ij[0] = 0.0;
        double I_current = 0.0;
        for (int k = 1; k <= ks; ++k) {
          I_current += 0.5 * dz[k] * (delta_ij[k - 1] + delta_ij[k]);
          I_ij[k] = I_current;
        }
        for (int k = ks + 1; k < Mz; ++k) {
          I_ij[k] = I_current;
        }
      }
    }
  }
  return ORC_OK;
}

// SIAFD.cc:890-948 (ghost exchange at :946-947 is the caller's job)
int compute_3d_horizontal_velocity(const orc_params &p, orc_fields &f) {
  int status = compute_I(p, f);
  if (status != ORC_OK) {
    return status;
  }
  const int Mz = p.Mz;
  View I[2] = {View(f.I_0, p, p.w_stag, Mz), View(f.I_1, p, p.w_stag, Mz)};
  View h_x(f.h_x, p, p.w_stag, 2), h_y(f.h_y, p, p.w_stag, 2), u_out(f.u, p, p.w_uv, Mz),
      v_out(f.v, p, p.w_uv, Mz), sliding(f.sliding, p, p.w_sliding, 2);

  for (int j = p.ys; j < p.ys + p.ym; ++j) {
    for (int i = p.xs; i < p.xs + p.xm; ++i) {
      const double *I_e = I[0].column(i, j), *I_w = I[0].column(i - 1, j), *I_n = I[1].column(i, j),
                   *I_s = I[1].column(i, j - 1);
      const double h_x_w = h_x(i - 1, j, 0), h_x_e = h_x(i, j, 0), h_x_n = h_x(i, j, 1), h_x_s = h_x(i, j - 1, 1);
      const double h_y_w = h_y(i - 1, j, 0), h_y_e = h_y(i, j, 0), h_y_n = h_y(i, j, 1), h_y_s = h_y(i, j - 1, 1);
      const double sliding_velocity_u = sliding(i, j, 0), sliding_velocity_v = sliding(i, j, 1);
      double *u_ij = u_out.column(i, j), *v_ij = v_out.column(i, j);
      for (int k = 0; k < Mz; ++k) {
        u_ij[k] = sliding_velocity_u - 0.25 * (I_e[k] * h_x_e + I_w[k] * h_x_w + I_n[k] * h_x_n + I_s[k] * h_x_s);
      }
      for (int k = 0; k < Mz; ++k) {
        v_ij[k] = sliding_velocity_v - 0.25 * (I_e[k] * h_y_e + I_w[k] * h_y_w + I_n[k] * h_y_n + I_s[k] * h_y_s);
      }
    }
  }
  return ORC_OK;
}

int check_params(const orc_params &p) {
  if (p.Mz < 2 || p.w_geom < 2 || p.w_3d_in < 2 || p.w_stag < 1 || p.w_uv < 1 || p.xm < 1 || p.ym < 1) {
    return ORC_ERR_BAD_CONFIG;
  }
  if (p.gradient_method < 0 || p.gradient_method > 2 || p.flow_law < 0 || p.flow_law > 6) {
    return ORC_ERR_BAD_CONFIG;
  }
  // SIAFD.cc:69-76
  if (p.grain_size_age_coupling && p.flow_law != ORC_FLOW_GK) {
    return ORC_ERR_BAD_CONFIG;
  }
  return ORC_OK;
}

} // namespace

// =========================================================================================
// C interface
// =========================================================================================
extern "C" {

void orc_default_params(orc_params *p) {
  std::memset(p, 0, sizeof(*p));
  p->w_geom = 2;
  p->w_3d_in = 2;
  p->w_stag = 1;
  p->w_uv = 1;
  p->w_sliding = 1;
  // src/pism_config.cdl (line numbers in SURVEY.md section 5.6)
  p->ec_p_air = 0.0;
  p->ec_g = 9.81;
  p->ec_beta = 7.9e-8;
  p->ec_rho_i = 910.0;
  p->ec_c_i = 2009.0;
  p->ec_c_w = 4170.0;
  p->ec_L = 3.34e5;
  p->ec_T_melting = 273.15;
  p->ec_T_0 = 223.15;
  p->flow_law = ORC_FLOW_GPBLD;
  p->fl_n = 3.0;
  p->fl_e = 1.0;
  p->fl_e_interglacial = 1.0;
  p->fl_A_cold = 3.61e-13;
  p->fl_A_warm = 1.73e3;
  p->fl_Q_cold = 6.0e4;
  p->fl_Q_warm = 13.9e4;
  p->fl_T_crit = 263.15;
  p->fl_R = 8.31441;
  p->fl_rho = 910.0;
  p->fl_g = 9.81;
  p->fl_beta = 7.9e-8;
  p->fl_T_melting = 273.15;
  p->gpbld_T_0 = 273.15;
  p->gpbld_water_frac_coeff = 181.25;
  p->gpbld_water_frac_limit = 0.01;
  p->iso_softness_A = 3.1689e-24;
  p->hooke_Q = 7.88e4;
  p->hooke_A = 4.42165e-9;
  p->hooke_C = 0.16612;
  p->hooke_K = 1.17;
  p->hooke_Tr = 273.39;
  p->grain_size = 1.0e-3; // 1 mm in metres
  p->gradient_method = ORC_GRAD_HASELOFF;
  p->limit_diffusivity = 0;
  p->D_limit = 100.0;
  const double secpera = 365.242198781 * 86400.0; // UDUNITS-2 "year"
  p->eemian_start = -132000.0 * secpera;
  p->eemian_end = -114500.0 * secpera;
  p->holocene_start = -11000.0 * secpera;
  p->years_per_second = 1.0 / secpera;
  p->smoother_range = 5.0e3;
  p->theta_min = 0.0;
  p->sea_water_density = 1028.0;
  p->ice_free_thickness = 0.01;
  p->dry_simulation = 0;
}

double orc_flow(const orc_params *p, double stress, double E, double pressure, double gs) {
  Converter ec(*p);
  std::unique_ptr<Law> law = make_law(*p, ec);
  if (not law) {
    return NAN;
  }
  return law->flow(stress, E, pressure, gs);
}

double orc_ec_pressure(const orc_params *p, double depth) { return Converter(*p).pressure(depth); }
double orc_ec_melting_temperature(const orc_params *p, double P) { return Converter(*p).melting_temperature(P); }
double orc_ec_enthalpy_cts(const orc_params *p, double P) { return Converter(*p).enthalpy_cts(P); }
double orc_ec_temperature(const orc_params *p, double E, double P) { return Converter(*p).temperature(E, P); }
double orc_ec_pressure_adjusted_temperature(const orc_params *p, double E, double P) {
  return Converter(*p).pressure_adjusted_temperature(E, P);
}
double orc_ec_water_fraction(const orc_params *p, double E, double P) { return Converter(*p).water_fraction(E, P); }
double orc_ec_enthalpy(const orc_params *p, double T, double omega, double P) {
  return Converter(*p).enthalpy(T, omega, P);
}
double orc_ec_enthalpy_permissive(const orc_params *p, double T, double omega, double P) {
  return Converter(*p).enthalpy_permissive(T, omega, P);
}
double orc_grain_size_vostok(double age_years) { return vostok(age_years); }

// IceGrid.cc:381-424
void orc_vertical_levels(double Lz, int Mz, int quadratic, double lambda, double *z) {
  if (not quadratic) {
    double dz = Lz / ((double)Mz - 1);
    for (int k = 0; k < Mz - 1; k++) {
      z[k] = dz * ((double)k);
    }
    z[Mz - 1] = Lz;
  } else {
    for (int k = 0; k < Mz - 1; k++) {
      const double zeta = ((double)k) / ((double)Mz - 1);
      z[k] = Lz * ((zeta / lambda) * (1.0 + (lambda - 1.0) * zeta));
    }
    z[Mz - 1] = Lz;
  }
}

int orc_k_below_height(const double *z, int Mz, double height, int *status) {
  int s = ORC_OK;
  int k = k_below_height(z, Mz, height, &s);
  if (status) {
    *status = s;
  }
  return k;
}

// IceGrid.cc:443-484
int orc_compute_nprocs(int Mx_, int My_, int size_, int *Nx_out, int *Ny_out) {
  unsigned int Mx = Mx_, My = My_, size = size_, Nx, Ny;
  if (My <= 0) {
    return ORC_ERR_BAD_CONFIG;
  }
  Nx = (unsigned int)(0.5 + sqrt(((double)Mx) * ((double)size) / ((double)My)));
  Ny = 0;
  if (Nx == 0) {
    Nx = 1;
  }
  while (Nx > 0) {
    Ny = size / Nx;
    if (Nx * Ny == (unsigned int)size) {
      break;
    }
    Nx--;
  }
  if (Mx > My and Nx < Ny) {
    int tmp = Nx;
    Nx = Ny;
    Ny = tmp;
  }
  if ((Mx / Nx) < 2 || (My / Ny) < 2) {
    return ORC_ERR_BAD_CONFIG;
  }
  *Nx_out = Nx;
  *Ny_out = Ny;
  return ORC_OK;
}

// IceGrid.cc:489-499
void orc_ownership_ranges(int M, int N, int *out) {
  for (int i = 0; i < N; i++) {
    out[i] = M / N + ((M % N) > i);
  }
}

// Mask.hh:96-133
void orc_geometry_compute(const orc_params *p, int n, const double *sea_level, const double *bed,
                          const double *thickness, double *mask_out, double *surface_out) {
  const double alpha = 1 - p->ec_rho_i / p->sea_water_density; // Mask.hh:72 (constants.ice.density)
  for (int q = 0; q < n; ++q) {
    const double hgrounded = bed[q] + thickness[q];
    const double hfloating = sea_level[q] + alpha * thickness[q];
    const bool is_floating = (hfloating > hgrounded), ice_free = (thickness[q] <= p->ice_free_thickness);
    int mask_result;
    double surface_result;
    if (is_floating && (not p->dry_simulation)) {
      surface_result = hfloating;
      mask_result = ice_free ? 4 : 3;
    } else {
      surface_result = hgrounded;
      mask_result = ice_free ? 0 : 2;
    }
    if (surface_out) {
      surface_out[q] = surface_result;
    }
    if (mask_out) {
      mask_out[q] = mask_result;
    }
  }
}

// BedSmoother.cc:99-267.  smoothing_range <= 0: copy of topg, Nx = Ny = -1 (:101-109); the
// other four fields stay zero (never written; PETSc Vecs start zeroed).
int orc_preprocess_bed(const orc_params *p, const double *topg, double *topgsmooth, double *maxtl, double *C2,
                       double *C3, double *C4, int *Nx_out, int *Ny_out) {
  const int Mx = p->Mx, My = p->My;
  const long N = (long)Mx * My;
  if (p->smoother_range <= 0.0) {
    std::memcpy(topgsmooth, topg, sizeof(double) * N);
    std::fill(maxtl, maxtl + N, 0.0);
    std::fill(C2, C2 + N, 0.0);
    std::fill(C3, C3 + N, 0.0);
    std::fill(C4, C4 + N, 0.0);
    *Nx_out = -1;
    *Ny_out = -1;
    return ORC_OK;
  }
  int Nx = static_cast<int>(ceil(p->smoother_range / p->dx));
  int Ny = static_cast<int>(ceil(p->smoother_range / p->dy));
  if (Nx < 1) {
    Nx = 1;
  }
  if (Ny < 1) {
    Ny = 1;
  }
  if (Nx >= Mx || Ny >= My) {
    return ORC_ERR_BAD_CONFIG;
  }
  *Nx_out = Nx;
  *Ny_out = Ny;
#define B0(i, j) topg[(long)(j) * Mx + (i)]
  // smooth_the_bed_on_proc0, :157-191
  for (int j = 0; j < My; j++) {
    for (int i = 0; i < Mx; i++) {
      double sum = 0.0, count = 0.0;
      for (int r = -Nx; r <= Nx; r++) {
        for (int s = -Ny; s <= Ny; s++) {
          if ((i + r >= 0) and (i + r < Mx) and (j + s >= 0) and (j + s < My)) {
            sum += B0(i + r, j + s);
            count += 1.0;
          }
        }
      }
      topgsmooth[(long)j * Mx + i] = sum / count;
    }
  }
  // compute_coefficients_on_proc0, :194-267
  for (int j = 0; j < My; j++) {
    for (int i = 0; i < Mx; i++) {
      double topgs = topgsmooth[(long)j * Mx + i], maxtltemp = 0.0, sum2 = 0.0, sum3 = 0.0, sum4 = 0.0, count = 0.0;
      for (int r = -Nx; r <= Nx; r++) {
        for (int s = -Ny; s <= Ny; s++) {
          if ((i + r >= 0) && (i + r < Mx) && (j + s >= 0) && (j + s < My)) {
            const double tl = B0(i + r, j + s) - topgs;
            maxtltemp = std::max(maxtltemp, tl);
            const double tl2 = tl * tl;
            sum2 += tl2;
            sum3 += tl2 * tl;
            sum4 += tl2 * tl2;
            count += 1.0;
          }
        }
      }
      maxtl[(long)j * Mx + i] = maxtltemp;
      C2[(long)j * Mx + i] = sum2 / count;
      C3[(long)j * Mx + i] = sum3 / count;
      C4[(long)j * Mx + i] = sum4 / count;
    }
  }
#undef B0
  const double n = p->fl_n, k = (n + 2) / n, s2 = k * (2 * n + 2) / (2 * n), s3 = s2 * (3 * n + 2) / (3 * n),
               s4 = s3 * (4 * n + 2) / (4 * n);
  for (long q = 0; q < N; ++q) { // VecScale
    C2[q] = s2 * C2[q];
    C3[q] = s3 * C3[q];
    C4[q] = s4 * C4[q];
  }
  return ORC_OK;
}

int orc_theta(const orc_params *p, const orc_fields *f, double *theta_out) { return theta(*p, *f, theta_out); }
int orc_smoothed_thk(const orc_params *p, const orc_fields *f, double *result) {
  return smoothed_thk(*p, *f, result);
}

// DMDA periodic ghosts on a single rank: ghost (i, j) <- owned ((i mod Mx), (j mod My)).
void orc_wrap_ghosts(int Mx, int My, int w, int dof, double *a) {
  const int nx = Mx + 2 * w;
  for (int jj = 0; jj < My + 2 * w; ++jj) {
    for (int ii = 0; ii < nx; ++ii) {
      const int i = ii - w, j = jj - w;
      if (i >= 0 && i < Mx && j >= 0 && j < My) {
        continue;
      }
      const int si = ((i % Mx) + Mx) % Mx, sj = ((j % My) + My) % My;
      const double *src = a + ((long)(sj + w) * nx + (si + w)) * dof;
      double *dst = a + ((long)jj * nx + ii) * dof;
      std::memcpy(dst, src, sizeof(double) * dof);
    }
  }
}

int orc_siafd_gradient(const orc_params *p, orc_fields *f) {
  int status = check_params(*p);
  if (status != ORC_OK) {
    return status;
  }
  // SIAFD.cc:197-220
  switch (p->gradient_method) {
  case ORC_GRAD_ETA:
    gradient_eta(*p, *f);
    break;
  case ORC_GRAD_HASELOFF:
    gradient_haseloff(*p, *f);
    break;
  case ORC_GRAD_MAHAFFY:
    gradient_mahaffy(*p, *f);
    break;
  default:
    return ORC_ERR_BAD_CONFIG;
  }
  return ORC_OK;
}

int orc_siafd_flux_velocity(const orc_params *p, orc_fields *f, int full) {
  int status = check_params(*p);
  if (status != ORC_OK) {
    return status;
  }
  Converter ec(*p);
  std::unique_ptr<Law> law = make_law(*p, ec);
  status = compute_diffusivity(*p, *f, full != 0, *law, ec);
  if (status != ORC_OK) {
    return status;
  }
  compute_diffusive_flux(*p, *f);
  if (full) {
    status = compute_3d_horizontal_velocity(*p, *f);
  }
  return status;
}

int orc_siafd_update_single(const orc_params *p, orc_fields *f, int full) {
  if (p->xs != 0 || p->ys != 0 || p->xm != p->Mx || p->ym != p->My) {
    return ORC_ERR_BAD_CONFIG;
  }
  int status = orc_siafd_gradient(p, f);
  if (status != ORC_OK) {
    return status;
  }
  if (p->gradient_method == ORC_GRAD_HASELOFF) { // SIAFD.cc:498-499
    orc_wrap_ghosts(p->Mx, p->My, p->w_stag, 2, f->h_x);
    orc_wrap_ghosts(p->Mx, p->My, p->w_stag, 2, f->h_y);
  }
  status = orc_siafd_flux_velocity(p, f, full);
  if (status != ORC_OK) {
    return status;
  }
  if (full) { // SIAFD.cc:946-947
    orc_wrap_ghosts(p->Mx, p->My, p->w_uv, p->Mz, f->u);
    orc_wrap_ghosts(p->Mx, p->My, p->w_uv, p->Mz, f->v);
  }
  return ORC_OK;
}

int orc_siafd_update_many(int n, const orc_params *p, orc_fields *f, int full, int nthreads) {
  int worst = ORC_OK;
#ifdef _OPENMP
  if (nthreads > 0) {
    omp_set_num_threads(nthreads);
  }
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int q = 0; q < n; ++q) {
    int status = orc_siafd_update_single(&p[q], &f[q], full);
    if (status != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
      worst = status;
    }
  }
  (void)nthreads;
  return worst;
}

// ---------------------------------------------------------------------------------------
// SIAFD::update on the n patches of ONE domain (PISM's DMDA decomposition, one patch per MPI rank in the reference;
// here the patches share the address space and an OpenMP thread takes the place of a rank).  Same passes as the
// reference; its two ghost updates (SIAFD.cc:498-499 h_x, h_y; :946-947 u, v; DMLocalToLocal on the periodic BOX
// stencil) are copies from the owning patch, and D_max is the maximum over the patches (:748).  Used by bench.py's CPU
// arm ("one domain split over all cores") and by the tests of the decomposed path.
// ---------------------------------------------------------------------------------------
namespace {
struct PatchMap {
  std::vector<int> xo, yo; // global column / row -> index of its 1D range
  std::vector<int> rank;   // [iy * nx + ix] -> patch
  int nx = 0, ny = 0;
};
bool build_patch_map(int n, const orc_params *p, PatchMap &M) {
  const int Mx = p[0].Mx, My = p[0].My;
  M.xo.assign(Mx, -1), M.yo.assign(My, -1);
  std::vector<int> xs_list, ys_list;
  for (int q = 0; q < n; ++q) {
    if (std::find(xs_list.begin(), xs_list.end(), p[q].xs) == xs_list.end()) xs_list.push_back(p[q].xs);
    if (std::find(ys_list.begin(), ys_list.end(), p[q].ys) == ys_list.end()) ys_list.push_back(p[q].ys);
  }
  std::sort(xs_list.begin(), xs_list.end()), std::sort(ys_list.begin(), ys_list.end());
  M.nx = (int)xs_list.size(), M.ny = (int)ys_list.size();
  if (M.nx * M.ny != n) return false;
  M.rank.assign(n, -1);
  for (int q = 0; q < n; ++q) {
    const int ix = (int)(std::find(xs_list.begin(), xs_list.end(), p[q].xs) - xs_list.begin());
    const int iy = (int)(std::find(ys_list.begin(), ys_list.end(), p[q].ys) - ys_list.begin());
    M.rank[iy * M.nx + ix] = q;
    for (int i = p[q].xs; i < p[q].xs + p[q].xm; ++i) M.xo[i] = ix;
    for (int j = p[q].ys; j < p[q].ys + p[q].ym; ++j) M.yo[j] = iy;
  }
  for (int v : M.xo) if (v < 0) return false;
  for (int v : M.yo) if (v < 0) return false;
  for (int v : M.rank) if (v < 0) return false;
  return true;
}
// ghosts of width w of patch q's array a[q] (ghost width W, dof values per cell) from the owners' arrays
void pull_ghosts(int q, const orc_params *p, const PatchMap &M, double *const *a, int W, int w, int dof) {
  const orc_params &P = p[q];
  const long rowc = P.xm + 2 * W;
  for (int j = P.ys - w; j < P.ys + P.ym + w; ++j) {
    const int gj = ((j % P.My) + P.My) % P.My;
    const bool jown = j >= P.ys && j < P.ys + P.ym;
    for (int i = P.xs - w; i < P.xs + P.xm + w; ++i) {
      if (jown && i >= P.xs && i < P.xs + P.xm) {
        i = P.xs + P.xm - 1; // skip the owned run
        continue;
      }
      const int gi = ((i % P.Mx) + P.Mx) % P.Mx;
      const int r = M.rank[M.yo[gj] * M.nx + M.xo[gi]];
      const orc_params &R = p[r];
      const double *src = a[r] + ((long)(gj - (R.ys - W)) * (R.xm + 2 * W) + (gi - (R.xs - W))) * dof;
      double *dst = a[q] + ((long)(j - (P.ys - W)) * rowc + (i - (P.xs - W))) * dof;
      std::memcpy(dst, src, sizeof(double) * dof);
    }
  }
}
} // namespace

int orc_siafd_update_decomposed(int n, const orc_params *p, orc_fields *f, int full, int nthreads) {
  PatchMap M;
  if (n < 1 || !build_patch_map(n, p, M)) return ORC_ERR_BAD_CONFIG;
  int worst = ORC_OK;
  std::vector<double *> hx(n), hy(n), u(n), v(n);
  for (int q = 0; q < n; ++q) hx[q] = f[q].h_x, hy[q] = f[q].h_y, u[q] = f[q].u, v[q] = f[q].v;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int q = 0; q < n; ++q) {
      const int st = orc_siafd_gradient(&p[q], &f[q]);
      if (st != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
        worst = st;
      }
    }
    if (p[0].gradient_method == ORC_GRAD_HASELOFF) { // SIAFD.cc:498-499
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
      for (int q = 0; q < n; ++q) {
        pull_ghosts(q, p, M, hx.data(), p[q].w_stag, p[q].w_stag, 2);
        pull_ghosts(q, p, M, hy.data(), p[q].w_stag, p[q].w_stag, 2);
      }
    }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int q = 0; q < n; ++q) {
      const int st = orc_siafd_flux_velocity(&p[q], &f[q], full);
      if (st != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
        worst = st;
      }
    }
    if (full) { // SIAFD.cc:946-947
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
      for (int q = 0; q < n; ++q) {
        pull_ghosts(q, p, M, u.data(), p[q].w_uv, p[q].w_uv, p[q].Mz);
        pull_ghosts(q, p, M, v.data(), p[q].w_uv, p[q].w_uv, p[q].Mz);
      }
    }
  }
  double dmax = 0.0; // SIAFD.cc:748
  for (int q = 0; q < n; ++q) dmax = std::max(dmax, f[q].D_max);
  for (int q = 0; q < n; ++q) f[q].D_max = dmax;
  return worst;
}

// ---------------------------------------------------------------------------------------
// IceModelVec3D::getValZ (src/util/iceModelVec3.cc:153-182), IceModelVec3::getSurfaceValues (:226-240) and
// ::getHorSlice (:209-223).  a: 3D field of ghost width wa; heights: 2D field of ghost width wh (the ice thickness
// for getSurfaceValues) or NULL for the constant height z0; out: owned points only ([ym][xm]).
// ---------------------------------------------------------------------------------------
static double get_val_z(const double *column, const double *z, int Mz, double height) {
  if (height >= z[Mz - 1]) {
    return column[Mz - 1];
  } else if (height <= z[0]) {
    return column[0];
  }
  const int mcurr = gsl_style_find(z, Mz, height); // gsl_interp_accel_find
  const double incr = (height - z[mcurr]) / (z[mcurr + 1] - z[mcurr]);
  const double valm = column[mcurr];
  return valm + incr * (column[mcurr + 1] - valm);
}

void orc_value_at_height(const orc_params *p, const double *a, int wa, const double *heights, int wh, double z0,
                         double *out) {
  const int Mz = p->Mz;
  const long nxa = p->xm + 2 * wa, nxh = p->xm + 2 * wh;
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const double *column = a + ((long)(j - (p->ys - wa)) * nxa + (i - (p->xs - wa))) * Mz;
      const double height = heights ? heights[(long)(j - (p->ys - wh)) * nxh + (i - (p->xs - wh))] : z0;
      out[(long)(j - p->ys) * p->xm + (i - p->xs)] = get_val_z(column, p->z, Mz, height);
    }
  }
}

// ---------------------------------------------------------------------------------------
// StressBalance::compute_vertical_velocity  (src/stressbalance/StressBalance.cc:283-424)
// the next consumer of u, v.  mask: 2D, ghost width w_geom; u, v: 3D, ghost
// width w_uv with valid ghosts; basal_melt_rate: owned points only ([ym][xm]) or NULL; w: 3D,
// owned points only (WITHOUT_GHOSTS, StressBalance.cc:142).
// ---------------------------------------------------------------------------------------
int orc_vertical_velocity(const orc_params *p, const double *mask, const double *u, const double *v,
                          const double *basal_melt_rate, int use_upstream_fd, double *w) {
  const int Mz = p->Mz, wg = p->w_geom, wuv = p->w_uv;
  const long nxg = p->xm + 2 * wg, nxu = p->xm + 2 * wuv;
  const double dx = p->dx, dy = p->dy;
  std::vector<double> u_x_plus_v_y(Mz);
  auto M = [&](int i, int j) { return (int)floor(mask[(long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg))] + 0.5); };
  auto icy = [&](int i, int j) { const int m = M(i, j); return m == 2 || m == 3; };
  auto ice_free = [&](int i, int j) { return !icy(i, j); };
  auto col = [&](const double *a, int i, int j) {
    return a + ((long)(j - (p->ys - wuv)) * nxu + (i - (p->xs - wuv))) * Mz;
  };
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      double *w_ij = w + ((long)(j - p->ys) * p->xm + (i - p->xs)) * Mz;
      const double *u_w = col(u, i - 1, j), *u_ij = col(u, i, j), *u_e = col(u, i + 1, j);
      const double *v_s = col(v, i, j - 1), *v_ij = col(v, i, j), *v_n = col(v, i, j + 1);
      double west = 1.0, east = 1.0, south = 1.0, north = 1.0;
      double D_x = 0, D_y = 0;
      {
        if (use_upstream_fd) {
          const double uw = 0.5 * (u_w[0] + u_ij[0]), ue = 0.5 * (u_ij[0] + u_e[0]);
          if (uw > 0.0 and ue >= 0.0) {
            west = 1.0;
            east = 0.0;
          } else if (uw <= 0.0 and ue < 0.0) {
            west = 0.0;
            east = 1.0;
          } else {
            west = 1.0;
            east = 1.0;
          }
        }
        if ((icy(i, j) and ice_free(i + 1, j)) or (ice_free(i, j) and icy(i + 1, j))) {
          east = 0;
        }
        if ((icy(i, j) and ice_free(i - 1, j)) or (ice_free(i, j) and icy(i - 1, j))) {
          west = 0;
        }
        if (east + west > 0) {
          D_x = 1.0 / (dx * (east + west));
        } else {
          D_x = 0.0;
        }
      }
      {
        if (use_upstream_fd) {
          const double vs = 0.5 * (v_s[0] + v_ij[0]), vn = 0.5 * (v_ij[0] + v_n[0]);
          if (vs > 0.0 and vn >= 0.0) {
            south = 1.0;
            north = 0.0;
          } else if (vs <= 0.0 and vn < 0.0) {
            south = 0.0;
            north = 1.0;
          } else {
            south = 1.0;
            north = 1.0;
          }
        }
        if ((icy(i, j) and ice_free(i, j + 1)) or (ice_free(i, j) and icy(i, j + 1))) {
          north = 0;
        }
        if ((icy(i, j) and ice_free(i, j - 1)) or (ice_free(i, j) and icy(i, j - 1))) {
          south = 0;
        }
        if (north + south > 0) {
          D_y = 1.0 / (dy * (north + south));
        } else {
          D_y = 0.0;
        }
      }
      for (int k = 0; k < Mz; ++k) {
        double u_x = D_x * (west * (u_ij[k] - u_w[k]) + east * (u_e[k] - u_ij[k])),
               v_y = D_y * (south * (v_ij[k] - v_s[k]) + north * (v_n[k] - v_ij[k]));
        u_x_plus_v_y[k] = u_x + v_y;
      }
      if (basal_melt_rate != NULL) {
        w_ij[0] = -basal_melt_rate[(long)(j - p->ys) * p->xm + (i - p->xs)];
      } else {
        w_ij[0] = 0.0;
      }
      for (int k = 1; k < Mz; ++k) {
        const double dz = p->z[k] - p->z[k - 1];
        w_ij[k] = w_ij[k - 1] - (0.5 * dz) * (u_x_plus_v_y[k] + u_x_plus_v_y[k - 1]);
      }
    }
  }
  return ORC_OK;
}


// ---------------------------------------------------------------------------------------
// StressBalance::compute_volumetric_strain_heating  (src/stressbalance/StressBalance.cc:426-642)
// SURVEY.md 8(f) N3.  p describes the flow law the reference takes here -- the SHALLOW stress balance's
// (`stress_balance.ssa.` prefix even with ZeroSliding, ShallowStressBalance.cc / StressBalance.cc:508): its id,
// Glen exponent fl_n and enhancement factor fl_e.  thickness, mask: 2D w_geom; enthalpy: 3D w_3d_in; u, v: 3D w_uv
// with valid ghosts; Sigma: 3D owned only (WITHOUT_GHOSTS, StressBalance.cc:150).
// ---------------------------------------------------------------------------------------
int orc_strain_heating(const orc_params *p, const double *thickness, const double *mask, const double *enthalpy,
                       const double *u, const double *v, double *Sigma_out) {
  const Converter ec(*p);
  std::unique_ptr<Law> law = make_law(*p, ec);
  if (!law) {
    return ORC_ERR_BAD_CONFIG;
  }
  const int Mz = p->Mz, wg = p->w_geom, wuv = p->w_uv, we = p->w_3d_in;
  const long nxg = p->xm + 2 * wg, nxu = p->xm + 2 * wuv, nxe = p->xm + 2 * we;
  const double *z = p->z;
  auto M = [&](int i, int j) { return (int)floor(mask[(long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg))] + 0.5); };
  auto icy = [&](int i, int j) { const int m = M(i, j); return m == 2 || m == 3; };
  auto ice_free = [&](int i, int j) { return !icy(i, j); };
  auto col = [&](const double *a, int i, int j) {
    return a + ((long)(j - (p->ys - wuv)) * nxu + (i - (p->xs - wuv))) * Mz;
  };
  auto D2 = [](double u_x, double u_y, double u_z, double v_x, double v_y, double v_z) { // :446-448
    return 0.5 * ((u_x + v_y) * (u_x + v_y) + u_x * u_x + v_y * v_y +
                  0.5 * ((u_y + v_x) * (u_y + v_x) + u_z * u_z + v_z * v_z));
  };
  const double enhancement_factor = p->fl_e, n = law->exponent(), exponent = 0.5 * (1.0 / n + 1.0),
               e_to_a_power = pow(enhancement_factor, -1.0 / n), hardness_power = -1.0 / n;
  std::vector<double> depth(Mz), pressure(Mz), hardness(Mz);
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const double H = thickness[(long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg))];
      int status = ORC_OK;
      const int ks = k_below_height(z, Mz, H, &status);
      if (status != ORC_OK) {
        return status;
      }
      double west = 1, east = 1, south = 1, north = 1, D_x = 0, D_y = 0;
      {
        if ((icy(i, j) and ice_free(i + 1, j)) or (ice_free(i, j) and icy(i + 1, j))) {
          east = 0;
        }
        if ((icy(i, j) and ice_free(i - 1, j)) or (ice_free(i, j) and icy(i - 1, j))) {
          west = 0;
        }
        if (east + west > 0) {
          D_x = 1.0 / (p->dx * (east + west));
        } else {
          D_x = 0.0;
        }
      }
      {
        if ((icy(i, j) and ice_free(i, j + 1)) or (ice_free(i, j) and icy(i, j + 1))) {
          north = 0;
        }
        if ((icy(i, j) and ice_free(i, j - 1)) or (ice_free(i, j) and icy(i, j - 1))) {
          south = 0;
        }
        if (north + south > 0) {
          D_y = 1.0 / (p->dy * (north + south));
        } else {
          D_y = 0.0;
        }
      }
      const double *u_ij = col(u, i, j), *u_w = col(u, i - 1, j), *u_e = col(u, i + 1, j), *u_s = col(u, i, j - 1),
                   *u_n = col(u, i, j + 1);
      const double *v_ij = col(v, i, j), *v_w = col(v, i - 1, j), *v_e = col(v, i + 1, j), *v_s = col(v, i, j - 1),
                   *v_n = col(v, i, j + 1);
      const double *E_ij = enthalpy + ((long)(j - (p->ys - we)) * nxe + (i - (p->xs - we))) * Mz;
      double *Sigma = Sigma_out + ((long)(j - p->ys) * p->xm + (i - p->xs)) * Mz;
      for (int k = 0; k <= ks; ++k) {
        depth[k] = H - z[k];
      }
      ec.pressure(depth, ks, pressure);
      for (int k = 0; k <= ks; ++k) { // FlowLaw::hardness_n -> hardness_impl, FlowLaw.cc:135-144
        hardness[k] = pow(law->softness(E_ij[k], pressure[k]), hardness_power);
      }
      for (int k = 0; k <= ks; ++k) {
        double dz;
        double u_z = 0.0, v_z = 0.0, u_x = D_x * (west * (u_ij[k] - u_w[k]) + east * (u_e[k] - u_ij[k])),
               u_y = D_y * (south * (u_ij[k] - u_s[k]) + north * (u_n[k] - u_ij[k])),
               v_x = D_x * (west * (v_ij[k] - v_w[k]) + east * (v_e[k] - v_ij[k])),
               v_y = D_y * (south * (v_ij[k] - v_s[k]) + north * (v_n[k] - v_ij[k]));
        if (k > 0) {
          dz = z[k + 1] - z[k - 1];
          u_z = (u_ij[k + 1] - u_ij[k - 1]) / dz;
          v_z = (v_ij[k + 1] - v_ij[k - 1]) / dz;
        } else {
          dz = z[1] - z[0];
          u_z = (u_ij[1] - u_ij[0]) / dz;
          v_z = (v_ij[1] - v_ij[0]) / dz;
        }
        Sigma[k] = 2.0 * e_to_a_power * hardness[k] * pow(D2(u_x, u_y, u_z, v_x, v_y, v_z), exponent);
      }
      for (int k = ks + 1; k < Mz; ++k) {
        Sigma[k] = 0.0;
      }
    }
  }
  return ORC_OK;
}

} // extern "C"
