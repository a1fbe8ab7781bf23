// CPU check of csrc/b200_glibc_math.cuh: the restated exp / expm1 / log / pow must return the bits of
// the host's glibc functions (the arithmetic of the oracle, cpp_standalone) for every argument.
// Build: g++ -O2 -ffp-contract=off -mfma -I<dir with b200_libm_tables.h> -I<csrc>.
// Usage: glibc_math_test [arguments per distribution, default 2000000].  Prints one line per
// function and distribution, then "OK" or "FAIL".
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "b200_glibc_math.cuh"

static uint64_t s[2] = {0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull};
static inline uint64_t rnd() {                              // xorshift128+
    uint64_t a = s[0];
    const uint64_t b = s[1];
    s[0] = b;
    a ^= a << 23;
    s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
    return s[1] + b;
}
static inline double uni(double lo, double hi) { return lo + (hi - lo) * ((rnd() >> 11) * 0x1p-53); }
static inline double anybits() { return b200g_dbl(rnd()); }
static inline bool same(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::isnan(a) && std::isnan(b);
    return b200g_bits(a) == b200g_bits(b);
}

static long total_bad = 0;
template <typename Gen, typename Mine, typename Ref>
static void sweep1(const char* name, long n, Gen gen, Mine mine, Ref ref) {
    long bad = 0;
    double first = 0;
    for (long i = 0; i < n; ++i) {
        const double x = gen();
        if (!same(mine(x), ref(x))) { if (!bad) first = x; ++bad; }
    }
    printf("%-28s %10ld arguments, %ld differ", name, n, bad);
    if (bad) printf(" (first: %a -> %a, glibc %a)", first, mine(first), ref(first));
    printf("\n");
    total_bad += bad;
}
template <typename Gen, typename Mine, typename Ref>
static void sweep2(const char* name, long n, Gen gen, Mine mine, Ref ref) {
    long bad = 0;
    double fx = 0, fy = 0;
    for (long i = 0; i < n; ++i) {
        double x, y;
        gen(x, y);
        if (!same(mine(x, y), ref(x, y))) { if (!bad) { fx = x; fy = y; } ++bad; }
    }
    printf("%-28s %10ld arguments, %ld differ", name, n, bad);
    if (bad) printf(" (first: %a, %a -> %a, glibc %a)", fx, fy, mine(fx, fy), ref(fx, fy));
    printf("\n");
    total_bad += bad;
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    auto my_exp = [](double x) { return b200g::exp(x); };
    auto my_expm1 = [](double x) { return b200g::expm1(x); };
    auto my_log = [](double x) { return b200g::log(x); };
    auto my_pow = [](double x, double y) { return b200g::pow(x, y); };
    // volatile function pointers: the compiler must call libm, not fold or substitute
    double (*volatile ref_exp)(double) = ::exp;
    double (*volatile ref_expm1)(double) = ::expm1;
    double (*volatile ref_log)(double) = ::log;
    double (*volatile ref_pow)(double, double) = ::pow;
    auto r_exp = [&](double x) { return ref_exp(x); };
    auto r_expm1 = [&](double x) { return ref_expm1(x); };
    auto r_log = [&](double x) { return ref_log(x); };
    auto r_pow = [&](double x, double y) { return ref_pow(x, y); };

    sweep1("exp  [-20, 5] (HH rates)", n, [] { return uni(-20, 5); }, my_exp, r_exp);
    sweep1("exp  [-1, 1]", n, [] { return uni(-1, 1); }, my_exp, r_exp);
    sweep1("exp  [-760, 720]", n, [] { return uni(-760, 720); }, my_exp, r_exp);
    sweep1("exp  [-746, -707] subnormal", n / 4, [] { return uni(-746, -707); }, my_exp, r_exp);
    sweep1("exp  [709, 710] overflow", n / 4, [] { return uni(709, 710); }, my_exp, r_exp);
    sweep1("exp  any bit pattern", n, [] { return anybits(); }, my_exp, r_exp);
    sweep1("exp  tiny", n / 4, [] { return uni(-1, 1) * std::ldexp(1.0, -(int)(rnd() % 80)); }, my_exp, r_exp);

    sweep1("expm1 [-20, 5] (HH rates)", n, [] { return uni(-20, 5); }, my_expm1, r_expm1);
    sweep1("expm1 [-1.5, 1.5]", n, [] { return uni(-1.5, 1.5); }, my_expm1, r_expm1);
    sweep1("expm1 [-60, 720]", n, [] { return uni(-60, 720); }, my_expm1, r_expm1);
    sweep1("expm1 [35, 45] (k ~ 56)", n / 4, [] { return uni(35, 45); }, my_expm1, r_expm1);
    sweep1("expm1 [10, 16] (k ~ 20)", n / 4, [] { return uni(10, 16); }, my_expm1, r_expm1);
    sweep1("expm1 any bit pattern", n, [] { return anybits(); }, my_expm1, r_expm1);
    sweep1("expm1 tiny", n / 4, [] { return uni(-1, 1) * std::ldexp(1.0, -(int)(rnd() % 80)); }, my_expm1, r_expm1);

    sweep1("log  (0, 100]", n, [] { return uni(0, 100); }, my_log, r_log);
    sweep1("log  [0.9, 1.1] (near one)", n, [] { return uni(0.9, 1.1); }, my_log, r_log);
    sweep1("log  [0.93, 0.94], [1.06, 1.07]", n / 4, [] { return (rnd() & 1) ? uni(0.93, 0.94) : uni(1.06, 1.07); }, my_log, r_log);
    sweep1("log  whole range", n, [] { return std::ldexp(uni(0.5, 1), (int)(rnd() % 2098) - 1074); }, my_log, r_log);
    sweep1("log  any bit pattern", n, [] { return anybits(); }, my_log, r_log);
    sweep1("log  1 +- tiny", n / 4, [] { return 1.0 + uni(-1, 1) * std::ldexp(1.0, -(int)(rnd() % 53)); }, my_log, r_log);

    {
        double (*volatile ref_tanh)(double) = ::tanh;
        double (*volatile ref_sinh)(double) = ::sinh;
        double (*volatile ref_cosh)(double) = ::cosh;
        auto r_tanh = [&](double x) { return ref_tanh(x); };
        auto r_sinh = [&](double x) { return ref_sinh(x); };
        auto r_cosh = [&](double x) { return ref_cosh(x); };
        auto my_tanh = [](double x) { return b200g::tanh(x); };
        auto my_sinh = [](double x) { return b200g::sinh(x); };
        auto my_cosh = [](double x) { return b200g::cosh(x); };
        auto wide = [] { return uni(-1, 1) * std::ldexp(1.0, (int)(rnd() % 70) - 60); };     // 2^-60 .. 2^9
        sweep1("tanh [-25, 25]", n, [] { return uni(-25, 25); }, my_tanh, r_tanh);
        sweep1("tanh [-1.2, 1.2]", n, [] { return uni(-1.2, 1.2); }, my_tanh, r_tanh);
        sweep1("tanh all magnitudes", n / 2, wide, my_tanh, r_tanh);
        sweep1("tanh any bit pattern", n / 2, [] { return anybits(); }, my_tanh, r_tanh);
        sweep1("sinh [-25, 25]", n, [] { return uni(-25, 25); }, my_sinh, r_sinh);
        sweep1("sinh [-1.2, 1.2]", n, [] { return uni(-1.2, 1.2); }, my_sinh, r_sinh);
        sweep1("sinh [-712, 712]", n / 2, [] { return uni(-712, 712); }, my_sinh, r_sinh);
        sweep1("sinh all magnitudes", n / 2, wide, my_sinh, r_sinh);
        sweep1("sinh any bit pattern", n / 2, [] { return anybits(); }, my_sinh, r_sinh);
        sweep1("cosh [-25, 25]", n, [] { return uni(-25, 25); }, my_cosh, r_cosh);
        sweep1("cosh [-0.5, 0.5]", n, [] { return uni(-0.5, 0.5); }, my_cosh, r_cosh);
        sweep1("cosh [-712, 712]", n / 2, [] { return uni(-712, 712); }, my_cosh, r_cosh);
        sweep1("cosh all magnitudes", n / 2, wide, my_cosh, r_cosh);
        sweep1("cosh any bit pattern", n / 2, [] { return anybits(); }, my_cosh, r_cosh);
    }

    {
        double (*volatile ref_sin)(double) = ::sin;
        double (*volatile ref_cos)(double) = ::cos;
        auto r_sin = [&](double x) { return ref_sin(x); };
        auto r_cos = [&](double x) { return ref_cos(x); };
        auto my_sin = [](double x) { return b200g::sin(x); };
        auto my_cos = [](double x) { return b200g::cos(x); };
        auto mags = [] { return uni(-1, 1) * std::ldexp(1.0, (int)(rnd() % 90) - 62); };     // 2^-62 .. 2^27
        sweep1("sin  [-0.9, 0.9]", n, [] { return uni(-0.9, 0.9); }, my_sin, r_sin);
        sweep1("sin  [-2.5, 2.5]", n, [] { return uni(-2.5, 2.5); }, my_sin, r_sin);
        sweep1("sin  [-100, 100]", n, [] { return uni(-100, 100); }, my_sin, r_sin);
        sweep1("sin  [-1e8, 1e8]", n, [] { return uni(-1.05e8, 1.05e8); }, my_sin, r_sin);
        sweep1("sin  all magnitudes", n, mags, my_sin, r_sin);
        sweep1("sin  near multiples of pi/2", n / 2, [] {
            return (double)((long)(rnd() % 2000000) - 1000000) * 1.5707963267948966 + uni(-1e-6, 1e-6); }, my_sin, r_sin);
        sweep1("cos  [-0.9, 0.9]", n, [] { return uni(-0.9, 0.9); }, my_cos, r_cos);
        sweep1("cos  [-2.5, 2.5]", n, [] { return uni(-2.5, 2.5); }, my_cos, r_cos);
        sweep1("cos  [-100, 100]", n, [] { return uni(-100, 100); }, my_cos, r_cos);
        sweep1("cos  [-1e8, 1e8]", n, [] { return uni(-1.05e8, 1.05e8); }, my_cos, r_cos);
        sweep1("cos  all magnitudes", n, mags, my_cos, r_cos);
        sweep1("cos  near multiples of pi/2", n / 2, [] {
            return (double)((long)(rnd() % 2000000) - 1000000) * 1.5707963267948966 + uni(-1e-6, 1e-6); }, my_cos, r_cos);
    }

    sweep2("pow  gate**{3,4}", n, [](double& x, double& y) { x = uni(0, 1); y = 3 + (double)(rnd() & 1); }, my_pow, r_pow);
    sweep2("pow  exp(a)**c (HH rates)", n, [&](double& x, double& y) {
        x = ref_exp(uni(-10, 10)); y = (rnd() & 1) ? 0.025 : 0.05555555555555555; }, my_pow, r_pow);
    sweep2("pow  (0,10)**[-10,10]", n, [](double& x, double& y) { x = uni(0, 10); y = uni(-10, 10); }, my_pow, r_pow);
    sweep2("pow  (0,1e300)**[-3,3]", n, [](double& x, double& y) {
        x = std::ldexp(uni(0.5, 1), (int)(rnd() % 2000) - 1000); y = uni(-3, 3); }, my_pow, r_pow);
    sweep2("pow  near over/underflow", n, [&](double& x, double& y) {
        x = uni(1.5, 30); const double e = (rnd() & 1) ? uni(700, 712) : uni(-750, -700); y = e / std::log(x); }, my_pow, r_pow);
    sweep2("pow  negative ** integer", n / 4, [](double& x, double& y) {
        x = -uni(0, 20); y = (double)((long)(rnd() % 41) - 20); }, my_pow, r_pow);
    sweep2("pow  negative ** any", n / 4, [](double& x, double& y) { x = -uni(0, 20); y = uni(-5, 5); }, my_pow, r_pow);
    sweep2("pow  subnormal base", n / 4, [](double& x, double& y) {
        x = b200g_dbl(rnd() >> 12); y = uni(-1.2, 1.2); }, my_pow, r_pow);
    sweep2("pow  any bits ** any bits", n, [](double& x, double& y) { x = anybits(); y = anybits(); }, my_pow, r_pow);
    sweep2("pow  any bits ** small", n, [](double& x, double& y) {
        x = anybits(); y = uni(-2, 2) * std::ldexp(1.0, -(int)(rnd() % 70)); }, my_pow, r_pow);
    sweep2("pow  special values", n / 4, [](double& x, double& y) {
        static const double v[] = {0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, NAN, 0.5, -0.5, 2.0, -2.0, 3.0, -3.0,
                                   0x1p63, -0x1p63, 0x1p-70, 0x1p53 + 2, 0x1p52 + 1, 1e308, 5e-324, -5e-324};
        const int m = sizeof(v) / sizeof(v[0]);
        x = v[rnd() % m]; y = v[rnd() % m]; }, my_pow, r_pow);

    printf(total_bad ? "FAIL\n" : "OK\n");
    return total_bad ? 1 : 0;
}
