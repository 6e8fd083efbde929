// dsim_packed.cuh — two environments per lane on Blackwell's packed FP32 pipe.
//
// A three-register FFMA / FMUL / FADD issues every other cycle per scheduler on sm_100 (register read ports): scalar FP32
// code tops out at HALF the FP32 peak, and the step kernel's physics sits right at that ceiling while a wave of pages is
// being computed (profiles/r02c_step_kernel.txt: FMA pipe 34 % of issue peak = 67 % of what scalar FP32 can get; four
// warps per scheduler x 886 FP32 instructions x 2 cycles = the 3.7 us the first pages take).  The packed forms
// (add / sub / mul / fma .f32x2, SASS FADD2 / FMUL2 / FFMA2) carry two independent IEEE operations per lane and instruction
// through the same ports.  F2 holds the same quantity of TWO envs (lo: env of page A, hi: env of page B, same lane), so the
// whole templated physics (substep<T>, V3<T>, ...) instantiates for T = F2 unchanged.  nvcc does not contract or vectorise
// across inline asm, so products stay unevaluated (F2Mul) until they meet their addend: a * b + c, c - a * b,
// a * b - c * d become one FFMA2 like the scalar code's FFMA.
// Included by dsim_device.cuh (after the scalar helpers and V3, before the physics that uses it); not a stand-alone header.
#pragma once

namespace dsim {

struct F2Mul;
struct F2 {
    unsigned long long v;
    DSIM_DEV F2() {}
    DSIM_DEV F2(float s) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(s)); }
    DSIM_DEV F2(double s) { const float f = (float)s; asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f)); }
    DSIM_DEV F2(int s) { const float f = (float)s; asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(f)); }
    DSIM_DEV F2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); }
    DSIM_DEV F2(const F2Mul &m);
    DSIM_DEV float lo() const { return __uint_as_float((unsigned)(v & 0xffffffffULL)); }
    DSIM_DEV float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
};
struct F2Mul { F2 a, b; };                                         // a * b, not evaluated yet

DSIM_DEV F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DSIM_DEV F2 sub2(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DSIM_DEV F2 mul2(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DSIM_DEV F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
DSIM_DEV F2 neg2(F2 a) { F2 r; r.v = a.v ^ 0x8000000080000000ULL; return r; }
DSIM_DEV F2::F2(const F2Mul &m) { v = mul2(m.a, m.b).v; }

// ---- sums and differences: every product that meets an addend becomes one FFMA2
DSIM_DEV F2 operator+(F2 a, F2 b) { return add2(a, b); }
DSIM_DEV F2 operator-(F2 a, F2 b) { return sub2(a, b); }
DSIM_DEV F2 operator-(F2 a) { return neg2(a); }
DSIM_DEV F2 operator+(F2Mul m, F2 c) { return fma2(m.a, m.b, c); }
DSIM_DEV F2 operator+(F2 c, F2Mul m) { return fma2(m.a, m.b, c); }
DSIM_DEV F2 operator+(F2Mul m, F2Mul n) { return fma2(m.a, m.b, mul2(n.a, n.b)); }
DSIM_DEV F2 operator-(F2Mul m, F2 c) { return fma2(m.a, m.b, neg2(c)); }
DSIM_DEV F2 operator-(F2 c, F2Mul m) { return fma2(neg2(m.a), m.b, c); }
DSIM_DEV F2 operator-(F2Mul m, F2Mul n) { return fma2(neg2(n.a), n.b, mul2(m.a, m.b)); }
DSIM_DEV F2Mul operator-(F2Mul m) { return F2Mul{neg2(m.a), m.b}; }
// ---- products
DSIM_DEV F2Mul operator*(F2 a, F2 b) { return F2Mul{a, b}; }
DSIM_DEV F2Mul operator*(F2Mul m, F2 c) { return F2Mul{mul2(m.a, m.b), c}; }
DSIM_DEV F2Mul operator*(F2 c, F2Mul m) { return F2Mul{c, mul2(m.a, m.b)}; }
DSIM_DEV F2Mul operator*(F2Mul m, F2Mul n) { return F2Mul{mul2(m.a, m.b), mul2(n.a, n.b)}; }
DSIM_DEV F2 &operator+=(F2 &a, F2 b) { a = add2(a, b); return a; }
DSIM_DEV F2 &operator+=(F2 &a, F2Mul m) { a = fma2(m.a, m.b, a); return a; }
DSIM_DEV F2 &operator-=(F2 &a, F2 b) { a = sub2(a, b); return a; }
DSIM_DEV F2 &operator-=(F2 &a, F2Mul m) { a = fma2(neg2(m.a), m.b, a); return a; }

// ---- V3 of unevaluated products: what the generic mk(s * a.x, s * a.y, s * a.z) returns; converts where a V3<F2> is expected
template <> struct V3<F2Mul> {
    F2Mul x, y, z;
    DSIM_DEV operator V3<F2>() const { V3<F2> r; r.x = F2(x); r.y = F2(y); r.z = F2(z); return r; }
};
// mixed argument lists (a product next to an evaluated value): evaluate
template <typename A, typename B, typename C,
          typename std::enable_if<(std::is_same<A, F2Mul>::value || std::is_same<B, F2Mul>::value || std::is_same<C, F2Mul>::value) &&
                                  !(std::is_same<A, F2Mul>::value && std::is_same<B, F2Mul>::value && std::is_same<C, F2Mul>::value), int>::type = 0>
DSIM_DEV V3<F2> mk(A x, B y, C z) { V3<F2> r; r.x = F2(x); r.y = F2(y); r.z = F2(z); return r; }
DSIM_DEV V3<F2> operator*(F2Mul s, V3<F2> a) { const F2 t(s); return mk(t * a.x, t * a.y, t * a.z); }
DSIM_DEV V3<F2> operator+(V3<F2> a, V3<F2Mul> b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DSIM_DEV V3<F2> operator+(V3<F2Mul> a, V3<F2> b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DSIM_DEV V3<F2> operator+(V3<F2Mul> a, V3<F2Mul> b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
DSIM_DEV V3<F2> operator-(V3<F2> a, V3<F2Mul> b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DSIM_DEV V3<F2> operator-(V3<F2Mul> a, V3<F2> b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
DSIM_DEV V3<F2> operator-(V3<F2Mul> a, V3<F2Mul> b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }

// ---- scalar helpers, per half (the MUFU ops have no packed form)
DSIM_DEV F2 sqrt_(F2 x) { return F2(sqrt_(x.lo()), sqrt_(x.hi())); }
DSIM_DEV F2 rcp_(F2 x) { return F2(rcp_(x.lo()), rcp_(x.hi())); }
DSIM_DEV F2 rsqrt_(F2 x) { return F2(rsqrt_(x.lo()), rsqrt_(x.hi())); }
DSIM_DEV F2 abs_(F2 x) { F2 r; r.v = x.v & 0x7fffffff7fffffffULL; return r; }
DSIM_DEV F2 sqrt_(F2Mul x) { return sqrt_(F2(x)); }
DSIM_DEV F2 rcp_(F2Mul x) { return rcp_(F2(x)); }
DSIM_DEV F2 rsqrt_(F2Mul x) { return rsqrt_(F2(x)); }
DSIM_DEV F2 abs_(F2Mul x) { return abs_(F2(x)); }
DSIM_DEV F2 max_(F2 a, F2 b) { return F2(fmaxf(a.lo(), b.lo()), fmaxf(a.hi(), b.hi())); }
DSIM_DEV F2 min_(F2 a, F2 b) { return F2(fminf(a.lo(), b.lo()), fminf(a.hi(), b.hi())); }
DSIM_DEV F2 max_(F2 a, F2Mul b) { return max_(a, F2(b)); }
DSIM_DEV F2 clamp_(F2 x, F2 lo, F2 hi) { return min_(max_(x, lo), hi); }
// hinge angles: the branch-free Cody-Waite / minimax scheme of the scalar version, packed; the quadrant fix-ups are integer
// operations on the two halves
DSIM_DEV void sincos_hinge(F2 a, F2 *s, F2 *c) {
    const F2 magic(12582912.0f);
    const F2 t = fma2(a, F2(0.636619772f), magic);
    const int q0 = __float_as_int(t.lo()), q1 = __float_as_int(t.hi());
    const F2 j = sub2(t, magic);
    F2 r = fma2(j, F2(-1.57079601e+00f), a);
    r = fma2(j, F2(-3.13916473e-07f), r);
    r = fma2(j, F2(-5.39030253e-15f), r);
    const F2 r2 = mul2(r, r);
    F2 sp = fma2(r2, F2(-1.9515295891e-4f), F2(8.3321608736e-3f)); sp = fma2(sp, r2, F2(-1.6666654611e-1f));
    const F2 sn = fma2(mul2(sp, r2), r, r);
    F2 cp = fma2(r2, F2(2.443315711809948e-5f), F2(-1.388731625493765e-3f)); cp = fma2(cp, r2, F2(4.166664568298827e-2f)); cp = fma2(cp, r2, F2(-0.5f));
    const F2 cs = fma2(cp, r2, F2(1.0f));
    const float sn0 = sn.lo(), sn1 = sn.hi(), cs0 = cs.lo(), cs1 = cs.hi();
    const float ss0 = (q0 & 1) ? cs0 : sn0, cc0 = (q0 & 1) ? sn0 : cs0, ss1 = (q1 & 1) ? cs1 : sn1, cc1 = (q1 & 1) ? sn1 : cs1;
    *s = F2(__int_as_float(__float_as_int(ss0) ^ ((q0 & 2) << 30)), __int_as_float(__float_as_int(ss1) ^ ((q1 & 2) << 30)));
    *c = F2(__int_as_float(__float_as_int(cc0) ^ (((q0 + 1) & 2) << 30)), __int_as_float(__float_as_int(cc1) ^ (((q1 + 1) & 2) << 30)));
}

}  // namespace dsim
