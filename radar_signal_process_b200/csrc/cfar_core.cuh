// cfar_core.cuh -- the single CA-CFAR decision shared by every CFAR kernel.
#pragma once
#include "common.cuh"

namespace rb {

// One CA-CFAR decision for element y of an axis of length N whose element i sits at base[i*stride].
template <typename T>
__device__ __forceinline__ bool cfar_decide(const T* __restrict__ base, ptrdiff_t stride, int y, int N, int ref, int guard,
                                            T thr, int method, int* err_flag) {
    const int l1 = y - guard - ref;
    const int r1 = y + guard + 1;
    const bool okL = l1 >= 0;
    const bool okR = (y + guard + ref) <= N - 1;
    if (!okL && !okR) {            // MATLAB: index exceeds array bounds
        if (err_flag) *err_flag = 1;
        return false;
    }
    T sl = 0, sr = 0;
    if (okL) for (int j = 0; j < ref; ++j) sl += base[(ptrdiff_t)(l1 + j) * stride];
    if (okR) for (int j = 0; j < ref; ++j) sr += base[(ptrdiff_t)(r1 + j) * stride];
    const T mr = sr / (T)ref, ml = sl / (T)ref;
    const T a = okL ? ml : mr;
    const T b = okR ? mr : ml;
    const T mu = method == 0 ? (a > b ? a : b) : (a < b ? a : b);
    return base[(ptrdiff_t)y * stride] >= mu * thr;
}

}  // namespace rb
