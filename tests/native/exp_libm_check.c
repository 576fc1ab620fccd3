#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
/* Test infrastructure: the operation sequence of hla-la_b200/csrc/exp_libm.cuh (exp_like_host_libm) restated for the host, against this host's libm exp, bit for bit.
   The table is parsed out of that header at run time (argv[1]) so that the two cannot drift apart. Prints "<mismatches> mismatches of <n>". */
static uint64_t T[256];
static int load_table(const char* header) {
    FILE* f = fopen(header, "r"); if (!f) return 0;
    char tok[64]; int n = 0, c, k = 0, in = 0;
    while ((c = fgetc(f)) != EOF) {
        if (!in) { static const char key[] = "k_exp_tab[256] = {"; if (c == key[k]) { if (!key[++k]) in = 1; } else k = (c == key[0]); continue; }
        if (c == '}') break;
        if (c == '0' && n < 256) { int t = 0; tok[t++] = (char)c; while ((c = fgetc(f)) != EOF && ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || c == 'x')) if (t < 60) tok[t++] = (char)c; tok[t] = 0; T[n++] = strtoull(tok, NULL, 16); if (c == '}') break; }
    }
    fclose(f); return n == 256;
}
static inline uint64_t asu(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double asd(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
static double my_exp(double x) {
    const double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p52, NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
    const double C2 = 0x1.ffffffffffdbdp-2, C3 = 0x1.555555555543cp-3, C4 = 0x1.55555cf172b91p-5, C5 = 0x1.1111167a4d017p-7;
    uint32_t abstop = (asu(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9 >= 0x3f) { if (abstop < 0x3c9) return 1.0 + x; return exp(x); }
    double kd = fma(x, InvLn2N, Shift); uint64_t ki = asu(kd); kd -= Shift;
    double r = fma(kd, NegLn2loN, fma(kd, NegLn2hiN, x));
    uint64_t idx = 2 * (ki & 127), top = ki << 45;
    double tail = asd(T[idx]); uint64_t sbits = T[idx + 1] + top;
    double r2 = r * r;
    double p23 = fma(r, C3, C2), p45 = fma(r, C5, C4);
    double t1 = fma(p23, r2, r + tail);
    double tmp = fma(r2 * r2, p45, t1);
    double scale = asd(sbits);
    return fma(scale, tmp, scale);
}
int main(int argc, char** argv) {
    if (argc < 2 || !load_table(argv[1])) { fprintf(stderr, "usage: exp_libm_check <exp_libm.cuh> [samples]\n"); return 2; }
    const long samples = argc > 2 ? atol(argv[2]) : 200000000;
    uint64_t s = 88172645463325252ull; long bad = 0, n = 0;
    for (long i = 0; i < samples; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) / 9007199254740992.0;
        double x;
        switch (i & 3) { case 0: x = -u * 60; break; case 1: x = -u * 2; break; case 2: x = -u * 700; break; default: x = -ldexp(u, -(int)(s & 63)); }
        double a = exp(x), b = my_exp(x); n++;
        if (asu(a) != asu(b)) { if (bad < 5) printf("x=%a libm=%a mine=%a\n", x, a, b); bad++; }
    }
    printf("%ld mismatches of %ld\n", bad, n);
    return 0;
}
