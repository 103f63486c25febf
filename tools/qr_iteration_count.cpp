// Experiment (host): how many QR steps, at which active size, does max_abs_eig<5> take on the lambda box?
//   g++ -O2 -std=c++17 -I sdc_gym_b200/csrc tools/qr_iteration_count.cpp -o /tmp/qrc
//   python -c "import numpy as np; from sdc_gym_b200.collocation import collocation_matrix as C; \
//              from sdc_gym_b200.precond import fixed_preconditioner as F; \
//              print(*map(float, C(5).ravel()), *map(float, np.diag(F('min', 5))))" | /tmp/qrc
#include <cstdio>
#include <cstdint>
static long g_steps[8];
#define SDCGYM_QR_HOOK(hi) (g_steps[(hi)]++)
#include "specrad.cuh"
using namespace sdcgym;
int main() {
    const int M = 5;
    // Radau-right collocation matrix for M = 5 is read from stdin (25 doubles), then the MIN diagonal (5 doubles)
    double Q[25], d[5];
    for (int k = 0; k < 25; k++) if (scanf("%lf", &Q[k]) != 1) return 1;
    for (int k = 0; k < 5; k++) if (scanf("%lf", &d[k]) != 1) return 1;
    const int G = 256;
    long n = 0;
    double sum = 0;
    for (int a = 0; a < G; a++)
        for (int b = 0; b < G; b++) {
            double zr = -100.0 + 100.0 * a / (G - 1), zi = -10.0 + 10.0 * b / (G - 1);
            C2 Qd[25];
            for (int k = 0; k < 25; k++) Qd[k] = C2{0.0, 0.0};
            for (int k = 0; k < 5; k++) Qd[k * 5 + k] = C2{d[k], 0.0};
            sum += spectral_radius_one<M>(Q, zr, zi, Qd);
            n++;
        }
    long tot = 0;
    for (int h = 1; h < 5; h++) { printf("active size %d: %.3f steps per matrix\n", h + 1, (double)g_steps[h] / n); tot += g_steps[h]; }
    printf("total %.3f QR steps per matrix, mean rho %.6f\n", (double)tot / n, sum / n);
}
