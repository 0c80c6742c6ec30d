// siafd_device.cuh -- device-side parameter block, enthalpy converter and flow laws.
//
// Device restatement of the scalar physics SIAFD::compute_diffusivity calls per level
// (reference: juliusgarbe/pism v1.2.1).  Arithmetic notes (DESIGN.md "Arithmetic"):
// nvcc contracts a*b+c into FMA and uses libdevice exp/pow/sqrt, so results differ from the
// reference's glibc build in the last 1-2 ulp per operation -- far inside the 1e-10 bar.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace siafd {

enum : int { LAW_ISO = 0, LAW_PB = 1, LAW_GPBLD = 2, LAW_HOOKE = 3, LAW_ARR = 4, LAW_ARRWARM = 5, LAW_GK = 6 };
enum : int { GRAD_HASELOFF = 0, GRAD_MAHAFFY = 1, GRAD_ETA = 2 };

// error bits raised by kernels (mapped to SIAFD_B200_ERR_* on the host)
enum : unsigned { EB_NEG_THK = 1u, EB_OMEGA = 2u, EB_BELOW = 4u, EB_ABOVE = 8u, EB_COMM = 16u };

// Passed by value to every kernel (__grid_constant__).
struct DP {
  int Mx, My, Mz;
  int xs, xm, ys, ym;
  int wg, we, wst, wuv, wsl; // ghost widths: geometry, 3D inputs, staggered, u/v, sliding
  double dx, dy;
  double inv_dx, inv_dy; // correctly rounded reciprocals, for div_rn_by()
  // EnthalpyConverter (util/EnthalpyConverter.cc:55-69)
  double p_air, rg /* rho_i * g */, ec_beta, c_i, inv_c_i, c_w, L0, T_melting, T_0;
  // FlowLaw (rheology/FlowLaw.cc:33-58)
  int law, n_is_3;
  double n, nm1, e, e_inter;
  double A_cold, A_warm, Q_cold, Q_warm, T_crit, R;
  double QoR_cold, QoR_warm; // Q / R, for the lean Arrhenius evaluation (siafd_math.cuh)
  // the same factor in the form exp(lnA - (Q/R) / T) (siafd_slab.cu): ln A; 0.5 / c_i; and the cold-ice test
  // E < E_cts(p) as (E_ij + E_offset) < cts2_a - cts2_b p, cts2_a = 2 c_i (T_melting - T_0), cts2_b = 2 c_i beta
  double lnA_cold, lnA_warm, hic, cts2_a, cts2_b;
  double lnA2_cold, lnA2_warm, QoR2_cold, QoR2_warm; // ln A and Q / R times 16 / ln2 (exp2_tab16, siafd_math.cuh)
  double beta_ratio; // m_beta_CC_grad / (m_rho * m_g), rheology/PatersonBudd.cc:57
  double gp_T0, gp_coeff, gp_limit, gp_softness_T0; // rheology/GPBLD.cc:49-61
  double iso_A;
  double hk_Q, hk_A, hk_C, hk_K, hk_Tr;
  double grain_size;
  // SIAFD (sia/SIAFD.cc:555-570)
  int limit_diffusivity, gs_age, e_age, use_age;
  double D_limit, eemian_start, eemian_end, holocene_start, years_per_second, current_time;
  // BedSmoother (sia/BedSmoother.cc)
  int smoother_active, grad; // grad: GRAD_* (sia/SIAFD.cc:197-220)
  double theta_min;
  // GeometryCalculator (util/Mask.hh:71-79)
  double gc_alpha, gc_icefree;
  int gc_dry, pad2;
  // heaviest-first order of the fused kernel's row segments (siafd_slab.cu): rows per segment and number of segments
  // the 2D pass weighs (0: off)
  int seg_rows, seg_n;
};

// ---- local ghosted array indexing ([j][i][dof], util/IceModelVec_inline.hh:28-40) ----------
__host__ __device__ inline long idx2(const DP &P, int i, int j, int w) {
  return (long)(j - (P.ys - w)) * (P.xm + 2 * w) + (i - (P.xs - w));
}

// a / d, correctly rounded, for a divisor whose correctly rounded reciprocal inv_d = RN(1 / d) is at hand
// (Markstein: q = RN(a inv_d); r = a - d q exactly, by FMA; RN(q + r inv_d) = RN(a / d) in the absence of
// overflow / underflow).  Three FP64 instructions instead of the ~35 of the IEEE division sequence; the
// gradient kernels use it for the divisions by dx and dy that the tests pin bit for bit.
__device__ __forceinline__ double div_rn_by(double a, double d, double inv_d) {
  const double q = __dmul_rn(a, inv_d);
  const double r = __fma_rn(-q, d, a);
  return __fma_rn(r, inv_d, q);
}

// ---- mask predicates (util/Mask.hh:37-66; decoding util/IceModelVec_inline.hh:95-101) -------
__device__ inline int mask_int(double m) { return (int)floor(m + 0.5); }
__device__ inline bool m_ocean(int M) { return M >= 3; }
__device__ inline bool m_grounded(int M) { return !(M >= 3); }
__device__ inline bool m_icy(int M) { return M == 2 || M == 3; }
__device__ inline bool m_ice_free(int M) { return !(M == 2 || M == 3); }
__device__ inline bool m_floating_ice(int M) { return m_icy(M) && m_ocean(M); }
__device__ inline bool m_ice_free_ocean(int M) { return m_ocean(M) && m_ice_free(M); }

// ---- EnthalpyConverter -----------------------------------------------------------------
// util/EnthalpyConverter.cc:158-160
__device__ inline double ec_melting_temperature(const DP &P, double p) { return P.T_melting - P.ec_beta * p; }
// :378-380
__device__ inline double ec_enthalpy_cts(const DP &P, double p) {
  return P.c_i * (ec_melting_temperature(P, p) - P.T_0);
}
// :365-367
__device__ inline double ec_L(const DP &P, double T_pm) { return P.L0 + (P.c_w - P.c_i) * (T_pm - 273.15); }
// :180-188 with temperature_cold :388-390.  E / c_i is evaluated as E * (1/c_i): <= 1 ulp.
__device__ inline double ec_temperature(const DP &P, double E, double p) {
  const double T_m = ec_melting_temperature(P, p);
  const double E_cts = P.c_i * (T_m - P.T_0);
  return (E < E_cts) ? (E * P.inv_c_i + P.T_0) : T_m;
}

// rheology/FlowLaw.cc:89-94
__device__ inline double softness_paterson_budd(const DP &P, double T_pa) {
  const bool cold = T_pa < P.T_crit;
  const double A = cold ? P.A_cold : P.A_warm;
  const double Q = cold ? P.Q_cold : P.Q_warm;
  return A * exp(-Q / (P.R * T_pa));
}

// pow(stress, n-1) of rheology/FlowLaw.cc:104 (n = 3 is x*x: correctly rounded, within 1 ulp of glibc pow)
__device__ inline double stress_power(const DP &P, double stress) {
  return P.n_is_3 ? stress * stress : pow(stress, P.nm1);
}

// rheology/GoldsbyKohlstedt.cc:114-150
__device__ inline double gk_flow(const DP &P, double stress, double temp, double pressure, double gs) {
  const double V_act_vol = -13.e-6, disl_crit_temp = 258.0, disl_A_cold = 4.0e-19, disl_A_warm = 6.0e4, disl_n = 4.0,
               disl_Q_cold = 60.e3, disl_Q_warm = 180.e3, gbs_crit_temp = 255.0, gbs_A_cold = 6.1811e-14,
               gbs_A_warm = 4.7547e15, gbs_n = 1.8, gbs_Q_cold = 49.e3, gbs_Q_warm = 192.e3, p_grain_sz_exp = 1.4,
               basal_A = 2.1896e-7, basal_n = 2.4, basal_Q = 60.e3, diff_crit_temp = 258.0, diff_V_m = 1.97e-5,
               diff_D_0v = 9.10e-4, diff_Q_v = 59.4e3, diff_D_0b = 5.8e-4, diff_Q_b = 49.e3, diff_delta = 9.04e-10;
  if (fabs(stress) < 1e-10) {
    return 0.0;
  }
  const double T = temp + P.beta_ratio * pressure;
  const double pV = pressure * V_act_vol;
  const double RT = P.R * T;
  const double diff_D_v = diff_D_0v * exp(-diff_Q_v / RT);
  double diff_D_b = diff_D_0b * exp(-diff_Q_b / RT);
  if (T > diff_crit_temp) {
    diff_D_b *= 1000;
  }
  const double eps_diff =
      42 * diff_V_m * (diff_D_v + 3.14159265358979323846 * diff_delta * diff_D_b / gs) / (RT * (gs * gs));
  double eps_disl;
  if (T > disl_crit_temp) {
    eps_disl = disl_A_warm * pow(stress, disl_n - 1) * exp(-(disl_Q_warm + pV) / RT);
  } else {
    eps_disl = disl_A_cold * pow(stress, disl_n - 1) * exp(-(disl_Q_cold + pV) / RT);
  }
  const double eps_basal = basal_A * pow(stress, basal_n - 1) * exp(-(basal_Q + pV) / RT);
  double eps_gbs;
  if (T > gbs_crit_temp) {
    eps_gbs = gbs_A_warm * (pow(stress, gbs_n - 1) / pow(gs, p_grain_sz_exp)) * exp(-(gbs_Q_warm + pV) / RT);
  } else {
    eps_gbs = gbs_A_cold * (pow(stress, gbs_n - 1) / pow(gs, p_grain_sz_exp)) * exp(-(gbs_Q_cold + pV) / RT);
  }
  return eps_diff + eps_disl + (eps_basal * eps_gbs) / (eps_basal + eps_gbs);
}

// FlowLaw::flow(stress, E, p, gs) (rheology/FlowLaw.cc:97-105 and the per-law overrides).
template <int LAW> __device__ inline double flow_eval(const DP &P, double stress, double E, double p, double gs) {
  if (LAW == LAW_ISO) {
    // rheology/IsothermalGlen.cc:37-39
    return P.iso_A * stress_power(P, stress);
  } else if (LAW == LAW_ARR) {
    // rheology/PatersonBuddCold.cc:43-51 via PatersonBudd.cc:47-51: NOT pressure-adjusted
    const double T = ec_temperature(P, E, p);
    return (P.A_cold * exp(-P.Q_cold / (P.R * T))) * stress_power(P, stress);
  } else if (LAW == LAW_ARRWARM) {
    // rheology/PatersonBuddWarm.cc:42-50
    const double T = ec_temperature(P, E, p);
    return (P.A_warm * exp(-P.Q_warm / (P.R * T))) * stress_power(P, stress);
  } else if (LAW == LAW_PB) {
    // rheology/PatersonBudd.cc:47-59
    const double T = ec_temperature(P, E, p);
    const double T_pa = T + P.beta_ratio * p;
    return softness_paterson_budd(P, T_pa) * stress_power(P, stress);
  } else if (LAW == LAW_HOOKE) {
    // rheology/Hooke.cc:46-49 through PatersonBudd::flow_impl / flow_from_temp
    const double T = ec_temperature(P, E, p);
    const double T_pa = T + P.beta_ratio * p;
    const double A = P.hk_A * exp(-P.hk_Q / (P.R * T_pa) + 3.0 * P.hk_C * pow(P.hk_Tr - T_pa, -P.hk_K));
    return A * stress_power(P, stress);
  } else if (LAW == LAW_GPBLD) {
    // rheology/GPBLD.cc:49-61 times pow(stress, n-1) (FlowLaw.cc:102-105)
    const double T_m = ec_melting_temperature(P, p);
    const double E_s = P.c_i * (T_m - P.T_0);
    double softness;
    if (E < E_s) {
      // EnthalpyConverter::pressure_adjusted_temperature, util/EnthalpyConverter.cc:196-198
      const double T_pa = (E * P.inv_c_i + P.T_0) - T_m + P.T_melting;
      softness = softness_paterson_budd(P, T_pa);
    } else {
      // water_fraction, util/EnthalpyConverter.cc:214-223 (E == E_s gives 0 either way)
      double omega = (E - E_s) / ec_L(P, T_m);
      omega = fmin(omega, P.gp_limit);
      softness = P.gp_softness_T0 * (1.0 + P.gp_coeff * omega);
    }
    return softness * stress_power(P, stress);
  } else {
    // rheology/GoldsbyKohlstedt.cc:73-77
    const double T = ec_temperature(P, E, p);
    return gk_flow(P, stress, T, p, gs);
  }
}

// rheology/grain_size_vostok.cc:28-60 (clamped piecewise-linear table)
__device__ inline double grain_size_vostok(double age_years) {
  const double AGE[22] = {0.0000e+00, 5.0000e+01, 1.0000e+02, 1.2500e+02, 1.5000e+02, 1.5800e+02, 1.6500e+02, 1.7000e+02,
                          1.8000e+02, 1.8800e+02, 2.0000e+02, 2.2500e+02, 2.4500e+02, 2.6000e+02, 3.0000e+02, 3.2000e+02,
                          3.5000e+02, 4.0000e+02, 5.0000e+02, 6.0000e+02, 8.0000e+02, 1.0000e+04};
  const double GS[22] = {1.8000e-03, 2.2000e-03, 3.0000e-03, 4.0000e-03, 4.3000e-03, 3.0000e-03, 3.0000e-03, 4.6000e-03,
                         3.4000e-03, 3.3000e-03, 5.9000e-03, 6.2000e-03, 5.4000e-03, 6.8000e-03, 3.5000e-03, 6.0000e-03,
                         8.0000e-03, 8.3000e-03, 3.6000e-03, 3.8000e-03, 9.5000e-03, 1.0000e-02};
  double a = age_years / 1000.0;
  a = fmax(a, AGE[0]);
  a = fmin(a, AGE[21]);
  int ilo = 0, ihi = 21;
  while (ihi > ilo + 1) {
    int i = (ihi + ilo) / 2;
    if (AGE[i] > a) {
      ihi = i;
    } else {
      ilo = i;
    }
  }
  return GS[ilo] + (a - AGE[ilo]) / (AGE[ilo + 1] - AGE[ilo]) * (GS[ilo + 1] - GS[ilo]);
}

// sia/SIAFD.cc:951-961
__device__ inline bool interglacial(const DP &P, double accumulation_time) {
  if (accumulation_time < P.eemian_start) {
    return false;
  } else if (accumulation_time < P.eemian_end) {
    return true;
  } else if (accumulation_time < P.holocene_start) {
    return false;
  }
  return true;
}

} // namespace siafd
