#include <math.h>
#include <stdio.h>
#include <stdint.h>
int main() {
    uint64_t bad = 0;
    for (uint64_t n = 0; n < (1ull << 32); ++n) {
        double a = (double)n;
        double q = a * 0.001;
        double r = fma(-q, 1000.0, a);
        double q1 = fma(r, 0.001, q);
        if (q1 != a / 1000.0) { if (bad < 5) printf("mismatch n=%llu\n", (unsigned long long)n); ++bad; }
    }
    printf("bad=%llu\n", (unsigned long long)bad);
    return 0;
}
