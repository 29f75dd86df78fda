// Host build of fastest_lap_b200/csrc/fl_math.cuh for tests/test_flmath.py (the device build is checked by the GPU tests).
#include "../fastest_lap_b200/csrc/fl_math.cuh"
extern "C" void flm_eval(int fn, const double* x, double* y, long n)
{
    for (long i = 0; i < n; ++i)
    {
        double s, c;
        switch (fn)
        {
        case 0: y[i] = fl_rcp(x[i]); break;
        case 1: y[i] = fl_sqrt(x[i]); break;
        case 2: y[i] = fl_sin(x[i]); break;
        case 3: y[i] = fl_cos(x[i]); break;
        case 4: y[i] = fl_atan(x[i]); break;
        case 5: y[i] = fl_tanh(x[i]); break;
        case 6: y[i] = fl_exp(x[i]); break;
        case 7: fl_sincos(x[i], &s, &c); y[i] = s * s + c * c; break;
        default: y[i] = 0.0;
        }
    }
}
