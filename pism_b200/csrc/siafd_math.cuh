// siafd_math.cuh -- PTX helpers (mbarrier, bulk copy, cp.async) and the lean FP64 flow-law math shared by the
// fused kernels.  Reference arithmetic: rheology/FlowLaw.cc:89-105, GPBLD.cc:49-61, PatersonBudd*.cc,
// util/EnthalpyConverter.cc:137-223 of juliusgarbe/pism v1.2.1.
#pragma once
#include "siafd_kernels.cuh"

namespace siafd {

#define FULLMASK 0xffffffffu

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async8(unsigned smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// mbarrier + 1-D bulk copy (TMA engine; rows are contiguous runs of doubles, so no tensor map)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra.uni WAIT_DONE;\n"
               "bra.uni WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(unsigned smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk store (a row leaves through the copy engine), its group bookkeeping, and the fence that makes
// this thread's generic-proxy writes to shared memory visible to the copy engine
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// lean FP64 math for the Arrhenius factor A * exp(-Q / (R T))
// ---------------------------------------------------------------------------------------------
// 1 / a for a in the normal range: MUFU.RCP64H seed + two Newton steps (error <= ~1 ulp)
__device__ __forceinline__ double rcp_fast(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// Constants of exp_fast live in constant memory so that DFMA takes them as c[bank][offset] operands
// (literal doubles cost two UMOVs each inside the level loop, measured in profiles/).
static __constant__ double EXPC[16] = {
    1.4426950408889634e+0,   // [0]  log2(e)
    -6.9314718055994529e-1,  // [1]  -ln2 (high part)
    -2.3190468138462996e-17, // [2]  -ln2 (low part)
    1.6059043836821613e-10,  // [3]  1/13!
    2.0876756987868100e-09,  // [4]  1/12!
    2.5052108385441720e-08,  // [5]  1/11!
    2.7557319223985888e-07,  // [6]  1/10!
    2.7557319223985893e-06,  // [7]  1/9!
    2.4801587301587302e-05,  // [8]  1/8!
    1.9841269841269841e-04,  // [9]  1/7!
    1.3888888888888889e-03,  // [10] 1/6!
    8.3333333333333332e-03,  // [11] 1/5!
    4.1666666666666664e-02,  // [12] 1/4!
    1.6666666666666666e-01,  // [13] 1/3!
    6755399441055744.0,      // [14] 1.5 * 2^52
    0.0};

// exp(x) for |x| < 700 (no overflow / underflow / NaN handling): Cody-Waite reduction by ln 2, degree-13
// Taylor polynomial on [-ln2/2, ln2/2] (truncation 4e-18), exponent patched in.  Error <= ~1 ulp.
__device__ __forceinline__ double exp_fast(double x) {
  double t = fma(x, EXPC[0], EXPC[14]);
  const int n = __double2loint(t);
  t -= EXPC[14];
  double r = fma(t, EXPC[1], x);
  r = fma(t, EXPC[2], r);
  double p = EXPC[3];
  p = fma(p, r, EXPC[4]);
  p = fma(p, r, EXPC[5]);
  p = fma(p, r, EXPC[6]);
  p = fma(p, r, EXPC[7]);
  p = fma(p, r, EXPC[8]);
  p = fma(p, r, EXPC[9]);
  p = fma(p, r, EXPC[10]);
  p = fma(p, r, EXPC[11]);
  p = fma(p, r, EXPC[12]);
  p = fma(p, r, EXPC[13]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// exp(x) for |x| < 700 with a 16-entry table: x = (16 m + j) ln2 / 16 + r, |r| <= ln2 / 32, so that a degree-6
// polynomial suffices (truncation 4.4e-16 relative): exp(x) = 2^m * 2^(j/16) * p(r).  11 FP64 operations against
// 17 of exp_fast.  `tab` is EXPT staged in shared memory (lanes index it independently).  Error <= ~2 ulp.
static __constant__ double EXPT[16] = {
    1.0,                1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721,  1.241857812073484,
    1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407, 1.6104903319492543,
    1.681792830507429,  1.7562521603732995, 1.8340080864093424, 1.9152065613971474};
static __constant__ double EXPD[10] = {
    23.083120654223414,     // [0] 16 / ln2
    -0.043321698784993146,  // [1] -ln2 / 16 (high part, 42 bits)
    -3.436201886692732e-15, // [2] -ln2 / 16 (low part)
    0.001388888888888889,   // [3] 1/6!
    0.008333333333333333,   // [4] 1/5!
    0.041666666666666664,   // [5] 1/4!
    0.16666666666666666,    // [6] 1/3!
    6755399441055744.0,     // [7] 1.5 * 2^52
    0.0, 0.0};
__device__ __forceinline__ double exp_tab(double x, const double *tab) {
  double t = fma(x, EXPD[0], EXPD[7]);
  const int n = __double2loint(t);
  t -= EXPD[7];
  double r = fma(t, EXPD[1], x);
  r = fma(t, EXPD[2], r);
  double p = EXPD[3];
  p = fma(p, r, EXPD[4]);
  p = fma(p, r, EXPD[5]);
  p = fma(p, r, EXPD[6]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= tab[n & 15];
  return __hiloint2double(__double2hiint(p) + ((n >> 4) << 20), __double2loint(p));
}

// The same factor with the argument already in units of ln2 / 16: exp2_tab16(y) = 2^(y / 16).  The caller folds
// 16 / ln2 into ln A and Q / R (DP::lnA2_*, QoR2_*), so the reduction is three exact additions (t = y + 1.5 2^52,
// j = low word of t, r = y - (t - 1.5 2^52), |r| <= 1/2) instead of a multiply and a two-constant Cody-Waite
// step, and the polynomial takes r with (ln2 / 16)^k / k! as coefficients (degree 6: truncation 4.4e-16).
// 10 FP64 operations; error <= ~2 ulp.
static __constant__ double EXP2C[8] = {
    0.043321698784996581839,    // [0] ln2/16
    0.00093838479280898715755,  // [1] (ln2/16)^2 / 2
    0.000013550807779497456043, // [2] (ln2/16)^3 / 3!
    1.4676100322919429263e-7,   // [3] (ln2/16)^4 / 4!
    1.2715871950558131622e-9,   // [4] (ln2/16)^5 / 5!
    9.1812195738444387641e-12,  // [5] (ln2/16)^6 / 6!
    6755399441055744.0,         // [6] 1.5 * 2^52
    0.0};
__device__ __forceinline__ double exp2_tab16(double y, const double *tab) {
  double t = y + EXP2C[6];
  const int n = __double2loint(t);
  t -= EXP2C[6];
  const double r = y - t;
  double p = EXP2C[5];
  p = fma(p, r, EXP2C[4]);
  p = fma(p, r, EXP2C[3]);
  p = fma(p, r, EXP2C[2]);
  p = fma(p, r, EXP2C[1]);
  p = fma(p, r, EXP2C[0]);
  p = fma(p, r, 1.0);
  p *= tab[n & 15];
  return __hiloint2double(__double2hiint(p) + (n & ~15) * 65536, __double2loint(p));
}

// 1 / a with one third-order step on the MUFU.RCP64H seed s (20 mantissa bits): e = 1 - a s, 1 / a = s (1 + e + e^2 +
// ...), truncated after e^2: relative error e^3 <= 2^-60.  Three DFMA against the four of rcp_fast.
__device__ __forceinline__ double rcp_cubic(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  const double f = fma(e, e, e);
  return fma(r, f, r);
}

// A * exp(-Q / (R * T)) of FlowLaw::softness_paterson_budd (rheology/FlowLaw.cc:89-94) and the arr / arrwarm
// variants (PatersonBuddCold.cc:43-46, PatersonBuddWarm.cc:42-45), T in (150 K, 400 K)
__device__ __forceinline__ double arrhenius(double A, double Q_over_R, double T) {
  return A * exp_fast(-Q_over_R * rcp_fast(T));
}

// Lean restatement of flow_eval<LAW> (siafd_device.cuh) for NV independent (stress, E, p) triples at once.
// Straight-line code (no branch on the common path), so that the compiler interleaves the NV dependent
// FP64 chains: the kernel is bound by the latency of those chains, not by FP64 throughput (profiles/).
// Returns flow = softness(E, p) * stress^(n-1); inputs may be garbage for masked-out lanes (finite in,
// result discarded by the caller).
template <int LAW, int NV>
__device__ __forceinline__ void flow_lean_v(const DP &P, const double (&stress)[NV], const double (&E)[NV],
                                            const double (&p)[NV], const double (&gs)[NV], double (&out)[NV]) {
  double s2[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    s2[j] = P.n_is_3 ? stress[j] * stress[j] : pow(stress[j], P.nm1); // pow(stress, n-1), FlowLaw.cc:104
  }
  if (LAW == LAW_ISO) {
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = P.iso_A * s2[j];
  } else if (LAW == LAW_ARR || LAW == LAW_ARRWARM || LAW == LAW_PB) {
    double T[NV], A[NV], QoR[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const double T_m = fma(-P.ec_beta, p[j], P.T_melting);
      const double E_cts = P.c_i * (T_m - P.T_0);
      T[j] = (E[j] < E_cts) ? fma(E[j], P.inv_c_i, P.T_0) : T_m; // EnthalpyConverter::temperature, :180-188
      if (LAW == LAW_ARR) {
        A[j] = P.A_cold, QoR[j] = P.QoR_cold;
      } else if (LAW == LAW_ARRWARM) {
        A[j] = P.A_warm, QoR[j] = P.QoR_warm;
      } else {
        T[j] = fma(P.beta_ratio, p[j], T[j]); // rheology/PatersonBudd.cc:57
        const bool cold = T[j] < P.T_crit;
        A[j] = cold ? P.A_cold : P.A_warm, QoR[j] = cold ? P.QoR_cold : P.QoR_warm;
      }
    }
    double x[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) x[j] = -QoR[j] * rcp_fast(T[j]);
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = A[j] * exp_fast(x[j]) * s2[j];
  } else if (LAW == LAW_GPBLD) {
    // rheology/GPBLD.cc:49-61.  Cold branch evaluated for every lane; the (rare) temperate lanes are patched.
    double T_m[NV], E_s[NV], T_pa[NV], A[NV], QoR[NV], x[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      T_m[j] = fma(-P.ec_beta, p[j], P.T_melting);
      E_s[j] = P.c_i * (T_m[j] - P.T_0);
      T_pa[j] = fma(E[j], P.inv_c_i, P.T_0) - T_m[j] + P.T_melting; // EnthalpyConverter.cc:196-198
      const bool cold = T_pa[j] < P.T_crit;
      A[j] = cold ? P.A_cold : P.A_warm, QoR[j] = cold ? P.QoR_cold : P.QoR_warm;
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) x[j] = -QoR[j] * rcp_fast(T_pa[j]);
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = A[j] * exp_fast(x[j]);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (!(E[j] < E_s[j])) { // temperate ice
        const double Lm = fma(P.c_w - P.c_i, T_m[j] - 273.15, P.L0); // EnthalpyConverter::L, :365-367
        const double omega = fmin((E[j] - E_s[j]) * rcp_fast(Lm), P.gp_limit);
        out[j] = P.gp_softness_T0 * fma(P.gp_coeff, omega, 1.0);
      }
      out[j] *= s2[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = flow_eval<LAW>(P, stress[j], E[j], p[j], gs[j]); // hooke, gk
  }
}


} // namespace siafd
