// fp64 log / sincos(pi x) / sincos(x) / sqrt for the toy kernels' Box-Muller noise and the double-well force.
//
// Why not the CUDA math library: its polynomial coefficients are 64-bit immediates that the compiler
// materialises with pairs of UMOV on every evaluation -- 17 % of all instructions executed by
// toy_fused_kernel were UMOV (profiles/README.md) -- and it carries special-case paths (denormals, huge
// arguments, NaN) that cannot occur here.  These versions read their coefficients from the constant bank
// (operands of DFMA, no extra instruction), assume normal, bounded inputs, and are accurate to ~1 ulp
// (tests/test_toy_gpu.py::test_fastmath_accuracy).  Algorithms: the classic kernels of fdlibm
// (k_sin, k_cos, e_log), reciprocal / rsqrt from the hardware seed plus two Newton steps.
#pragma once
#include <cuda_runtime.h>

#include "fastmath64_tables.h"

namespace fm {

static __constant__ double kLg[7] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                     2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                     1.479819860511658591e-01};
static __constant__ double kS[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kC[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
static __constant__ double kMisc[10] = {
    6.93147180369123816490e-01,  // 0 ln2_hi
    1.90821492927058770002e-10,  // 1 ln2_lo
    3.141592653589793116,        // 2 pi_hi
    1.2246467991473532e-16,      // 3 pi_lo
    6.36619772367581382433e-01,  // 4 2/pi
    1.57079632673412561417e+00,  // 5 pio2_1  (first 33 bits of pi/2)
    6.07710050650619224932e-11,  // 6 pio2_1t (pi/2 - pio2_1)
    6755399441055744.0,          // 7 1.5 * 2^52: adding it rounds to the nearest integer, left in the low word
    1.4142135623730951,          // 8 sqrt(2)
    0.0};

// literals that do not fit a DFMA immediate (the compiler would rebuild them with two moves per use)
static __constant__ double kLit[6] = {1.0 / 120.0, -1.0 / 6.0, 1.0 / 24.0,
                                      -1.0 + 1.16415321826934814453125e-10,  // 3  2^-33 - 1
                                      -2.0 + 2.3283064365386962890625e-10,   // 4  2^-32 - 2
                                      0.0};

__device__ __forceinline__ double rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(fma(-x, y, 1.0), y, y);
    return fma(fma(-x, y, 1.0), y, y);
}

// sqrt of a positive normal number: seed (2^-23) -> one Newton step on 1/sqrt (2^-46) -> one Heron step on sqrt
__device__ __forceinline__ double sqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    const double r = x * y;
    return fma(fma(-r, r, x), 0.5 * y, r);
}

// (r + 0.5) * 2^-32 (or 2^-31) without an integer-to-double conversion: r becomes the top 32 fraction
// bits of a double with exponent 2^32, i.e. 2^32 + r exactly; one fma removes the offset (all exact)
__device__ __forceinline__ double uniform_from_u32(unsigned r, bool times_two)
{
    const double x = __hiloint2double((int)(0x41f00000u | (r >> 12)), (int)(r << 20));  // 2^32 + r
    return times_two ? fma(x, 4.656612873077392578125e-10, kLit[4])    // (r + 0.5) 2^-31
                     : fma(x, 2.3283064365386962890625e-10, kLit[3]);  // (r + 0.5) 2^-32
}

// natural logarithm of a positive NORMAL double (fdlibm e_log.c without the special cases)
__device__ __forceinline__ double log_pos(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    double m = __hiloint2double(hi, lo);  // [1, 2)
    const bool big = m > kMisc[8];
    m = big ? 0.5 * m : m;                // [sqrt(1/2), sqrt(2))
    e += big ? 1 : 0;
    const double f = m - 1.0;
    const double s = f * rcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, kLg[5], kLg[3]), kLg[1]);
    const double t2 = z * fma(w, fma(w, fma(w, kLg[6], kLg[4]), kLg[2]), kLg[0]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double k = (double)e;
    return fma(k, kMisc[0], -((hfsq - fma(s, hfsq + R, k * kMisc[1])) - f));
}

// sin and cos of t, |t| <= pi/4 (fdlibm k_sin.c / k_cos.c)
__device__ __forceinline__ void sincos_kernel(double t, double& s, double& c)
{
    const double z = t * t;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, kS[5], kS[4]), kS[3]), kS[2]), kS[1]), kS[0]);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, kC[5], kC[4]), kC[3]), kC[2]), kC[1]), kC[0]);
    s = fma(t * z, ps, t);
    c = fma(z * z, pc, fma(z, -0.5, 1.0));
}

// quadrant n: (sin, cos)(t + n pi/2) from (sin, cos)(t); sign flips are exact bit operations
__device__ __forceinline__ void rotate_quadrant(int n, double& s, double& c)
{
    const double s0 = (n & 1) ? c : s, c0 = (n & 1) ? s : c;
    const int ss = (int)(((unsigned)n & 2u) << 30), cs = (int)((((unsigned)n + 1u) & 2u) << 30);  // sign bits
    s = __hiloint2double(__double2hiint(s0) ^ ss, __double2loint(s0));
    c = __hiloint2double(__double2hiint(c0) ^ cs, __double2loint(c0));
}

// sin(pi x), cos(pi x) for 0 <= x < 4
__device__ __forceinline__ void sincospi_small(double x, double& s, double& c)
{
    const double t = fma(x, 2.0, kMisc[7]);
    const int n = __double2loint(t);
    const double r = fma(t - kMisc[7], -0.5, x);  // exact, |r| <= 1/4
    sincos_kernel(fma(r, kMisc[3], r * kMisc[2]), s, c);
    rotate_quadrant(n, s, c);
}

// sin(x), cos(x) for |x| < ~1e5 (two-term Cody-Waite reduction)
__device__ __forceinline__ void sincos_medium(double x, double& s, double& c)
{
    const double t = fma(x, kMisc[4], kMisc[7]);
    const int n = __double2loint(t);
    const double nd = t - kMisc[7];
    double r = fma(-nd, kMisc[5], x);
    r = fma(-nd, kMisc[6], r);
    sincos_kernel(r, s, c);
    rotate_quadrant(n, s, c);
}

// sin(x) alone for |x| < ~1e5: one polynomial, its coefficients picked by the quadrant's parity
// (sin: t + t z S(z); cos: 1 - z/2 + z^2 C(z); both are base + w P(z))
__device__ __forceinline__ double sin_medium(double x)
{
    const double t0 = fma(x, kMisc[4], kMisc[7]);
    const int n = __double2loint(t0);
    const double nd = t0 - kMisc[7];
    double t = fma(-nd, kMisc[5], x);
    t = fma(-nd, kMisc[6], t);
    const bool odd = n & 1;
    const double z = t * t;
    double p = odd ? kC[5] : kS[5];
    p = fma(z, p, odd ? kC[4] : kS[4]);
    p = fma(z, p, odd ? kC[3] : kS[3]);
    p = fma(z, p, odd ? kC[2] : kS[2]);
    p = fma(z, p, odd ? kC[1] : kS[1]);
    p = fma(z, p, odd ? kC[0] : kS[0]);
    const double base = odd ? fma(z, -0.5, 1.0) : t;
    const double w = z * (odd ? z : t);
    const double r = fma(w, p, base);
    return __hiloint2double(__double2hiint(r) ^ (int)(((unsigned)n & 2u) << 30), __double2loint(r));
}

// ------------------------------------------------------------------------------------------
// Table-driven versions for the Box-Muller pair (the way glibc evaluates log and sincos): the argument's
// leading bits pick a table entry, a short polynomial covers the remainder.  11 fp64 instructions for
// -2 log(u) instead of 28, 11 for sincos instead of 23.  The tables (18 KB, fastmath64_tables.h) are staged
// in shared memory by every CTA that uses them (load_tables, then a block barrier).
struct Tables {
    double2 lg[128];    // {1 / c_i, -2 log c_i}
    double2 tr[1024];   // {sin, cos}(2 pi j / 1024)
    static constexpr int kTrigBits = 10;
    __device__ __forceinline__ double2 log_entry(int i) const { return lg[i]; }
    __device__ __forceinline__ double2 trig_entry(unsigned j) const { return tr[j]; }
};

// The same tables laid out so that a random gather never has a bank conflict: a 16-byte shared-memory load is served one
// quarter-warp at a time (8 lanes x 16 B = the 128-byte width of the 32 banks), so every table is stored 8 times,
// interleaved -- entry i, copy c at index 8 i + c -- and lane l reads only copy l & 7: the 8 lanes of a quarter-warp
// always fall into 8 different 16-byte bank groups, whatever their indices.  (With one copy a random gather of the 1024-entry
// table averaged 9.3 wavefronts per request instead of 4, and the shared-memory pipe was 60 % busy: profiles/README.md.)
// The trig table keeps every second angle (512 entries, |delta| <= pi / 512: cos delta - 1 truncated after delta^4 is off
// by at most 7.4e-17), 64 KB; the log table 16 KB.  A TabView is a thread's pair of base pointers (copy l & 7 selected).
struct TabView {
    const double2* lg;
    const double2* tr;
    static constexpr int kTrigBits = 9;
    static constexpr int kCopies = 8;
    static constexpr int kBytes = (128 + 512) * 8 * (int)sizeof(double2);
    __device__ __forceinline__ double2 log_entry(int i) const { return lg[i * kCopies]; }
    __device__ __forceinline__ double2 trig_entry(unsigned j) const { return tr[j * kCopies]; }
};

// stage the replicated tables into `smem` (kBytes, 16-byte aligned) and return this thread's view.  Two phases: the 640
// distinct entries are read once from global memory (coalesced 16-byte loads) into `scratch` (>= 10 KB of shared memory the
// caller does not need yet), then every thread writes replicated entries from there (a shared-memory broadcast read, a
// conflict-free consecutive write).  Ends with a block barrier; `scratch` is free again afterwards.
__device__ __forceinline__ TabView load_tables_replicated(void* smem, void* scratch)
{
    double2* lg = reinterpret_cast<double2*>(smem);
    double2* tr = lg + 128 * TabView::kCopies;
    double2* tmp = reinterpret_cast<double2*>(scratch);
    for (int i = threadIdx.x; i < 128 + 512; i += blockDim.x)
        tmp[i] = i < 128 ? *reinterpret_cast<const double2*>(kLogTab[i])
                         : *reinterpret_cast<const double2*>(kTrigTab[2 * (i - 128)]);  // every second angle of the 1024-entry table
    __syncthreads();
    for (int i = threadIdx.x; i < (128 + 512) * TabView::kCopies; i += blockDim.x) lg[i] = tmp[i / TabView::kCopies];
    __syncthreads();
    const int c = threadIdx.x & (TabView::kCopies - 1);
    return TabView{lg + c, tr + c};
}

__device__ __forceinline__ void load_tables(Tables& T)
{
    for (int i = threadIdx.x; i < 128; i += blockDim.x) T.lg[i] = make_double2(kLogTab[i][0], kLogTab[i][1]);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) T.tr[i] = make_double2(kTrigTab[i][0], kTrigTab[i][1]);
}

static __constant__ double kL2[12] = {
    -1.386294361119890618834464,   // 0  -2 ln 2
    -2.0 / 7.0, 1.0 / 3.0, -2.0 / 5.0, 0.5, -2.0 / 3.0,  // 1..5  -2 log(1 + r) = r (-2 + r (1 + r (-2/3 + r (1/2 + ...))))
    1.462918079267159681051338e-09,   // 6  2 pi / 2^32
    -9.203883995854807744728685e-03,  // 7  -(2^22 + 2^21 - 1/2) 2 pi / 2^32
    162.9746617261008238273,          // 8  1024 / (2 pi)
    0x1.921fb54400000p-8,             // 9  2 pi / 1024, first 33 bits
    2.373867385353981485e-13,         // 10 2 pi / 1024 - [9]
    0.0};

// per trig-table size (index = kTrigBits - 9): offset of sincos2pi_u32_tab, N / (2 pi), 2 pi / N (first 33 bits, rest, rounded)
static __constant__ double kTrigC[2][5] = {
    {-1.840776872316865512303721e-02, 81.48733086305041191366849, 0x1.921fb54400000p-7, 4.747734770707962695e-13,
     0.0122718463030851298377447},
    {-9.203883995854807744728685e-03, 162.974661726100823827337, 0x1.921fb54400000p-8, 2.373867385353981347e-13,
     0.00613592315154256491887235}};

// -2 log(x) for a positive NORMAL double.  x = 2^k z with z in [0.6855, 1.3711); the interval of z (128 of them,
// centred on c_i, so that z = 1 is a centre and |r| <= 1/256) gives r = z / c_i - 1 with one rounding.
template <class TAB>
__device__ __forceinline__ double neg2log_tab(const TAB& T, double x)
{
    const int hi = __double2hiint(x);
    const int tmp = hi + 0x1000 - 0x3fe60000;
    const int k = tmp >> 20;
    const double z = __hiloint2double(hi - (tmp & (int)0xfff00000), __double2loint(x));
    const double2 e = T.log_entry((tmp >> 13) & 127);
    const double r = fma(z, e.x, -1.0);
    const double w = fma((double)k, kL2[0], e.y);
    double q = fma(r, kL2[1], kL2[2]);
    q = fma(r, q, kL2[3]);
    q = fma(r, q, kL2[4]);
    q = fma(r, q, kL2[5]);
    q = fma(r, q, 1.0);
    return fma(r, fma(r, q, -2.0), w);
}

// sin d and cos d - 1 for |d| <= pi / 1024 (truncation errors d^7 / 5040 and d^6 / 720 are below 2e-18)
__device__ __forceinline__ void sincos_tiny(double d, double& sd, double& cm)
{
    const double d2 = d * d;
    sd = fma(d2 * d, fma(d2, kLit[0], kLit[1]), d);
    cm = d2 * fma(d2, kLit[2], -0.5);
}

// sin and cos of 2 pi (r + 1/2) / 2^32 for a 32-bit word r: the nearest table angle 2 pi j / N comes from the
// top bits of r + 2^(31 - B), the remainder |d| <= pi / N from the low 32 - B bits (N = 2^B table entries)
template <class TAB>
__device__ __forceinline__ void sincos2pi_u32_tab(const TAB& T, unsigned r, double& s, double& c)
{
    constexpr int B = TAB::kTrigBits, E = 32 - B;
    const unsigned rr = r + (1u << (E - 1));  // wraps with the angle
    const double2 e = T.trig_entry(rr >> E);
    const unsigned lowbits = rr & ((1u << E) - 1u);
    const double X = __hiloint2double((int)(((1023u + E) << 20) | (lowbits >> (E - 20))), (int)(lowbits << (52 - E)));  // 2^E + low bits
    double sd, cm;
    sincos_tiny(fma(X, kL2[6], kTrigC[B - 9][0]), sd, cm);
    s = fma(e.x, cm, fma(e.y, sd, e.x));
    c = fma(e.y, cm, fma(-e.x, sd, e.y));
}

// sin(x) for |x| < ~1e5 from the same table: n = round(x N / 2 pi), two-term Cody-Waite remainder
template <class TAB>
__device__ __forceinline__ double sin_tab(const TAB& T, double x)
{
    constexpr int B = TAB::kTrigBits;
    const double t0 = fma(x, kTrigC[B - 9][1], kMisc[7]);
    const double2 e = T.trig_entry((unsigned)__double2loint(t0) & ((1u << B) - 1u));
    const double nd = t0 - kMisc[7];
    double d = fma(-nd, kTrigC[B - 9][2], x);
    d = fma(-nd, kTrigC[B - 9][3], d);
    double sd, cm;
    sincos_tiny(d, sd, cm);
    return fma(e.x, cm, fma(e.y, sd, e.x));
}

// cos(x) likewise
template <class TAB>
__device__ __forceinline__ double cos_tab(const TAB& T, double x)
{
    constexpr int B = TAB::kTrigBits;
    const double t0 = fma(x, kTrigC[B - 9][1], kMisc[7]);
    const double2 e = T.trig_entry((unsigned)__double2loint(t0) & ((1u << B) - 1u));
    const double nd = t0 - kMisc[7];
    double d = fma(-nd, kTrigC[B - 9][2], x);
    d = fma(-nd, kTrigC[B - 9][3], d);
    double sd, cm;
    sincos_tiny(d, sd, cm);
    return fma(e.y, cm, fma(-e.x, sd, e.y));
}

// sin(x) for |x| < 32 with ONE Cody-Waite term, the full-precision 2 pi / N: its representation error (1.1e-16 relative) times
// n 2 pi / N ~ |x| stays below 4e-15, a few ulp of the argument x itself (the double well evaluates sin(5 x + 5), whose
// argument already carries that rounding).  Saves one fp64 instruction per force evaluation.
template <class TAB>
__device__ __forceinline__ double sin_tab1(const TAB& T, double x)
{
    constexpr int B = TAB::kTrigBits;
    const double t0 = fma(x, kTrigC[B - 9][1], kMisc[7]);
    const double2 e = T.trig_entry((unsigned)__double2loint(t0) & ((1u << B) - 1u));
    const double d = fma(kMisc[7] - t0, kTrigC[B - 9][4], x);
    double sd, cm;
    sincos_tiny(d, sd, cm);
    return fma(e.x, cm, fma(e.y, sd, e.x));
}

// sqrt of a positive normal number in five instructions: with e = 1 - x y^2 for the hardware seed y (|e| < 2^-22),
// sqrt x = x y (1 + e/2 + 3 e^2/8 + O(e^3)); `scale` multiplies the result (the O substep's b sigma)
__device__ __forceinline__ double sqrt_pos_scaled(double x, double scale)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double q = e * fma(e, 0.375, 0.5);
    const double ts = t * scale;
    return fma(ts, q, ts);
}

}  // namespace fm
