// fastmath.cuh — latency-optimised fp64 exp and division for the scalar that sits on the sequential
// critical path of the logistic solvers (LogisticLoss gradient):
//     c(u) = −μ·y / (1 + exp(y·u))
// Every thread of the cluster evaluates this scalar once per step, between the end of the dot-product
// exchange and the start of the fused update, so its LATENCY is paid in full on every step (profile:
// exp() + __ddiv_rn ≈ 430 of 1060 cycles per SVRG step at d = 1024).  The CUDA math library optimises
// for throughput (Horner polynomial, division subroutine with a slow-path branch); here the dependent
// chain is what matters:
//   * exp: Cody–Waite reduction with the magic-number rounding, degree-13 polynomial split as
//     1 + r + r²·Q(r) with Q evaluated by Estrin's scheme (4 dependent FMA levels instead of 11),
//     2^k applied by an integer add on the exponent field.  ≤ 1.5 ulp (scripts/micro/logistic_micro.cu).
//   * division: MUFU.RCP64H seed (≥ 20 bits), one cubic step folded into the quotient, one
//     exact-remainder correction (Markstein) — five dependent FMAs, no branch, no subroutine.
// The result differs from the reference's libm/Julia `exp` by at most a few ulp — the same order as the
// CUDA library's own exp — far inside the parity tolerance (1e-8 objective, 1e-6 iterate).
#pragma once

__device__ __forceinline__ double fast_exp_core(double t, int kmin, int kmax) {
    // k = round(t / ln2) through the 1.5·2^52 shift: the integer lands in the low word of kd
    const double kd = fma(t, 1.4426950408889634074, 6755399441055744.0);
    const double kf = kd - 6755399441055744.0;
    int k = __double2loint(kd);
    k = max(kmin, min(k, kmax));  // integer pipe, off the fp64 chain: keeps 2^k representable for |t| > 700
    double r = fma(kf, -6.93147180559945290e-01, t);
    r = fma(kf, -2.31904681384629956e-17, r);  // |r| ≤ ln2/2
    const double r2 = r * r;
    const double r4 = r2 * r2;
    // Q(r) = Σ_{j=0..11} r^j / (j+2)!   (Taylor; the truncation error r^14/14! < 5e-18)
    const double p0 = fma(r, 1.0 / 6.0, 0.5);
    const double p1 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double p2 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double p3 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double p4 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
    const double p5 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    const double q0 = fma(r2, p1, p0);
    const double q1 = fma(r2, p3, p2);
    const double q2 = fma(r2, p5, p4);
    const double r8 = r4 * r4;
    const double s0 = fma(r4, q1, q0);
    const double Q = fma(r8, q2, s0);
    const double one_r = 1.0 + r;
    const double P = fma(r2, Q, one_r);  // e^r ∈ [0.70, 1.42]
    return __hiloint2double(__double2hiint(P) + (k << 20), __double2loint(P));
}
// exp(t) for |t| ≤ 700 (outside: 2^k saturates at k = ±1000, i.e. a finite huge / tiny positive number)
__device__ __forceinline__ double fast_exp(double t) { return fast_exp_core(t, -1000, 1000); }

// num / den for a finite den ≥ 1 (normal 1/den): correctly rounded except in rare double-rounding cases (≤ 1 ulp)
__device__ __forceinline__ double fast_div_pos(double num, double den) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));  // MUFU.RCP64H: relative error ≤ 2^-20
    const double e0 = fma(-den, r0, 1.0);
    const double q0 = num * r0;
    const double t = fma(e0, e0, e0);         // e + e²:  1/(1−e) = 1 + e + e² + O(e³),  e³ ≤ 2^-60
    const double q1 = fma(q0, t, q0);
    const double rem = fma(-q1, den, num);    // exact remainder
    return fma(rem, r0, q1);
}

// LogisticLoss coefficient  c = −μ y / (1 + exp(y u))   (∇f_i(x) = c·a_i, u = a_i·x)
__device__ __forceinline__ double logistic_coef_fast(double u, double y, double mu) {
    const double e = fast_exp(__dmul_rn(y, u));
    return fast_div_pos(__dmul_rn(-mu, y), __dadd_rn(1.0, e));
}
