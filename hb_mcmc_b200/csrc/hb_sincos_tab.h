// hb_sincos_tab.h -- host-side fill of the sin/cos table read by sincos_tab (hb_device.cuh):
// tab[k] = {sin, cos}(2 pi k / 1024), correctly rounded.  Computed in the first octant in extended
// precision (x87 long double, 64-bit mantissa: the argument pi k / 512 carries a 2^-64 relative error,
// far below half an ulp of the double result) and completed by the exact symmetries, so the
// values at the multiples of pi/2 are exactly 0 and +-1.
#pragma once
#include <cmath>

namespace hb {

inline void fill_sincos_table(double* tab /* [2 * 1024]: sin, cos interleaved */)
{
    const int N = 1024;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < N; k++) {
        const int q = k / (N / 4), m = k % (N / 4);  // quadrant and position inside it
        // first-quadrant angle m * (pi/2) / 256, evaluated from the nearer end of the quadrant
        long double s, c;
        if (m <= N / 8) {
            const long double a = pi * (long double)m / 512.0L;
            s = sinl(a);
            c = cosl(a);
        } else {
            const long double a = pi * (long double)(N / 4 - m) / 512.0L;
            s = cosl(a);
            c = sinl(a);
        }
        if (m == 0) { s = 0.0L; c = 1.0L; }
        double sv, cv;
        switch (q) {
            case 0: sv = (double)s; cv = (double)c; break;
            case 1: sv = (double)c; cv = -(double)s; break;
            case 2: sv = -(double)s; cv = -(double)c; break;
            default: sv = -(double)c; cv = (double)s; break;
        }
        tab[2 * k] = sv;
        tab[2 * k + 1] = cv;
    }
}

}  // namespace hb
