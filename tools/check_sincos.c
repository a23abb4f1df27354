/* Host check of the arithmetic of rb_sincos_fast (rigidbody_rs_b200/csrc/rb_dyn.cuh): the same operations with C99 fma(),
 * against long double sinl/cosl, over uniformly and logarithmically spaced arguments in (-1e5, 1e5).
 *   gcc -O2 -ffp-contract=off -o /tmp/check_sincos tools/check_sincos.c -lm && /tmp/check_sincos
 * Prints the largest error in ulps of the result; tests/test_oracle.py runs it and holds it to 1.5 ulp. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
static void fast(double x, double* sn, double* cs) {
    const int k = (int)rint(x * 0.6366197723675814);
    const double kd = (double)k;
    double r = fma(-kd, 1.5707963267948966, x);
    r = fma(-kd, 6.123233995736766e-17, r);
    r = fma(-kd, -1.4973849048591698e-33, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s0 = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double hz = 0.5 * z, w = 1.0 - hz;
    const double c0 = w + ((1.0 - w) - hz + z * (z * pc));
    const double a = (k & 1) ? c0 : s0, b = (k & 1) ? s0 : c0;
    *sn = (k & 2) ? -a : a;
    *cs = ((k + 1) & 2) ? -b : b;
}
static double ulps(double got, long double want) {
    const double w = (double)want;
    const double u = nextafter(fabs(w), INFINITY) - fabs(w);
    return (double)(fabsl((long double)got - want) / (u > 0 ? u : 4.9e-324));
}
int main(void) {
    double worst_s = 0, worst_c = 0, at_s = 0, at_c = 0, worst_abs = 0;
    unsigned long long st = 0x9E3779B97F4A7C15ULL;
    for (long i = 0; i < 4000000; ++i) {
        st = st * 6364136223846793005ULL + 1442695040888963407ULL;
        const double u = (double)(st >> 11) / 9007199254740992.0;
        double x;
        switch (i & 3) {
            case 0: x = (2 * u - 1) * 1.0e5; break;
            case 1: x = (2 * u - 1) * 10.0; break;
            case 2: x = exp((u * 2 - 1) * 11.5) * ((st & 1024) ? 1 : -1); break;                  /* 1e-5 .. 1e5 */
            default: x = rint((2 * u - 1) * 6.0e4) * 1.5707963267948966 + (u - 0.5) * 1e-6; break; /* near multiples of pi/2 */
        }
        if (!(fabs(x) < 1.0e5)) continue;
        double s, c;
        fast(x, &s, &c);
        const double es = ulps(s, sinl((long double)x)), ec = ulps(c, cosl((long double)x));
        if (es > worst_s) { worst_s = es; at_s = x; }
        if (ec > worst_c) { worst_c = ec; at_c = x; }
        const double ea = fmax(fabs(s - (double)sinl((long double)x)), fabs(c - (double)cosl((long double)x)));
        if (ea > worst_abs) worst_abs = ea;
    }
    printf("{\"max_ulp_sin\": %.3f, \"at_sin\": %.17g, \"max_ulp_cos\": %.3f, \"at_cos\": %.17g, \"max_abs\": %.3e}\n", worst_s, at_s, worst_c, at_c, worst_abs);
    return 0;
}
