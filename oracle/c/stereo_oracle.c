/* Plain-C restatement of the integer/copy rows of the ActiveZero stereo hot
 * path.  TEST INFRASTRUCTURE ONLY -- built by __graft_entry__.build() into
 * oracle/_build/libstereo_oracle.so and called from tests/ via ctypes.
 * Citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* a10: utils/warp_ops.py:22-45.  One (n,c,y) row per outer iteration (one CUDA
 * thread in the reference); disparity row shared by the c channels
 * (dbase = (i/h/c*h + i%h)*w, :27).  positive != 0 replays the `pos` kernel
 * (j descending, guard idx < w), else the `neg` kernel (j ascending, guard
 * idx > -1).  dst must be zero-filled by the caller (warp_ops.py:83). */
void azo_scatter_warp(float* dst, const float* src, const int32_t* disp,
                      int n, int c, int h, int w, int positive) {
    const long rows = (long)n * c * h;
    for (long i = 0; i < rows; ++i) {
        const long dbase = (i / h / c * h + i % h) * (long)w;
        if (positive) {
            for (int j = w - 1; j >= 0; --j) {
                int idx = j + disp[dbase + j];
                if (idx < w) dst[i * w + idx] = src[i * w + j];
            }
        } else {
            for (int j = 0; j < w; ++j) {
                int idx = j + disp[dbase + j];
                if (idx > -1) dst[i * w + idx] = src[i * w + j];
            }
        }
    }
}

/* a1: nets/psmnet/psmnet.py:151-165.  vol is [B,2C,Dq,H,W], zero-filled here. */
void azo_concat_volume(float* vol, const float* L, const float* R,
                       int B, int C, int Dq, int H, int W) {
    const size_t plane = (size_t)H * W;
    memset(vol, 0, sizeof(float) * (size_t)B * 2 * C * Dq * plane);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int i = 0; i < Dq; ++i)
                for (int y = 0; y < H; ++y) {
                    const float* l = L + (((size_t)b * C + c) * H + y) * W;
                    const float* r = R + (((size_t)b * C + c) * H + y) * W;
                    float* ol = vol + ((((size_t)b * 2 * C + c) * Dq + i) * H + y) * W;
                    float* orr = vol + ((((size_t)b * 2 * C + C + c) * Dq + i) * H + y) * W;
                    for (int x = i; x < W; ++x) {
                        ol[x] = l[x];
                        orr[x] = r[x - i];
                    }
                }
}

/* a4: psmnet.py:200-201 + psmnet_submodule.py:83-89, two-pass softmax in
 * double as an accuracy yardstick: out[b,y,x] = sum_d d * softmax_d(cost). */
void azo_soft_argmin_f64(double* out, const float* cost, int B, int D, int H, int W) {
    const size_t plane = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < plane; ++p) {
            const float* c = cost + (size_t)b * D * plane + p;
            double m = -INFINITY, s = 0.0, ws = 0.0;
            for (int d = 0; d < D; ++d) if (c[d * plane] > m) m = c[d * plane];
            for (int d = 0; d < D; ++d) {
                double e = exp((double)c[d * plane] - m);
                s += e;
                ws += e * d;
            }
            out[(size_t)b * plane + p] = ws / s;
        }
}
