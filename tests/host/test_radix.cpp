// Host check of the register butterflies in csrc/radix.cuh against a naive double-precision DFT.
// Built and run by tests/test_host_radix.py (g++, no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../radar_signal_process_b200/csrc/radix.cuh"

template <int R, int SIGN> static double check() {
    float2 v[R];
    double xr[R], xi[R];
    for (int j = 0; j < R; ++j) {
        xr[j] = (double)rand() / RAND_MAX - 0.5;
        xi[j] = (double)rand() / RAND_MAX - 0.5;
        v[j] = make_float2((float)xr[j], (float)xi[j]);
        xr[j] = v[j].x; xi[j] = v[j].y;
    }
    rb::Dft<R, SIGN>::run(v);
    double worst = 0;
    for (int k = 0; k < R; ++k) {
        double ar = 0, ai = 0;
        for (int j = 0; j < R; ++j) {
            double ang = SIGN * 2.0 * M_PI * j * k / R;
            ar += xr[j] * cos(ang) - xi[j] * sin(ang);
            ai += xr[j] * sin(ang) + xi[j] * cos(ang);
        }
        worst = fmax(worst, fmax(fabs(ar - v[k].x), fabs(ai - v[k].y)));
    }
    return worst;
}

int main() {
    double w = 0;
    for (int it = 0; it < 50; ++it) {
        w = fmax(w, check<2, -1>()); w = fmax(w, check<2, 1>());
        w = fmax(w, check<4, -1>()); w = fmax(w, check<4, 1>());
        w = fmax(w, check<8, -1>()); w = fmax(w, check<8, 1>());
        w = fmax(w, check<16, -1>()); w = fmax(w, check<16, 1>());
    }
    printf("worst abs error %.3e\n", w);
    return w < 2e-6 ? 0 : 1;
}
