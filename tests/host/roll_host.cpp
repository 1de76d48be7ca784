// roll_host.cpp -- lane-by-lane CPU emulation of the register-rolling kernels (k_gauss357_roll,
// k_canny_roll in img2sgf_b200/csrc), built on the SAME arithmetic cores (roll_cores.cuh, compiled
// here as plain C++).  Test infrastructure: lets the packed arithmetic and the strip / rotation
// logic be checked against the oracle in the GPU-less container.  A warp is emulated as 32 lanes
// advanced in lockstep, shuffles become reads of the neighbour lane's value.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../../img2sgf_b200/csrc/roll_cores.cuh"

using namespace i2s::roll;

static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static uint32_t load_word(const uint8_t *row, int x, int w, bool reflect, int stride, int c)
{
    if (x >= w + 8 || x < -8) return 0;
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) {
        int xx = reflect ? reflect101(x + k, w) : clampi(x + k, 0, w - 1);
        v |= (uint32_t)row[(size_t)xx * stride + c] << (8 * k);
    }
    return v;
}

extern "C" void rh_gauss357(const uint8_t *src, int h, int w, uint8_t *d3, uint8_t *d5, uint8_t *d7)
{
    const int TH = 64, OW = 120;
    const int strips_x = (w + OW - 1) / OW, strips_y = (h + TH - 1) / TH;
    for (int sy = 0; sy < strips_y; sy++)
        for (int sx = 0; sx < strips_x; sx++) {
            const int y0 = sy * TH, y1 = std::min(y0 + TH, h);
            uint32_t wl[32][7], wh[32][7];
            for (int lane = 0; lane < 32; lane++)
                for (int k = 0; k < 6; k++) {
                    int x = sx * OW - 4 + 4 * lane;
                    uint32_t v = load_word(src + (size_t)reflect101(y0 - 3 + k, h) * w, x, w, true, 1, 0);
                    wl[lane][k] = pair_lo(v); wh[lane][k] = pair_hi(v);
                }
            for (int yb = y0; yb < y1; yb += 7)
                for (int u = 0; u < 7; u++) {
                    const int y = yb + u;
                    if (y >= y1) continue;
                    uint32_t V[32][6];
                    for (int lane = 0; lane < 32; lane++) {
                        int x = sx * OW - 4 + 4 * lane;
                        uint32_t v = load_word(src + (size_t)reflect101(y + 3, h) * w, x, w, true, 1, 0);
                        wl[lane][(u + 6) % 7] = pair_lo(v); wh[lane][(u + 6) % 7] = pair_hi(v);
                        uint32_t rl[7], rh[7];
                        for (int k = 0; k < 7; k++) { rl[k] = wl[lane][(u + k) % 7]; rh[k] = wh[lane][(u + k) % 7]; }
                        gauss_vertical(rl, rh, V[lane]);
                    }
                    for (int lane = 1; lane <= 30; lane++) {
                        int x = sx * OW - 4 + 4 * lane;
                        if (x >= w) continue;
                        const uint32_t *L = V[lane - 1], *C = V[lane], *R = V[lane + 1];
                        uint32_t o3 = gauss_h3(L[1], C[0], C[1], R[0]);
                        uint32_t o5 = gauss_h5(L[3], C[2], C[3], R[2]);
                        uint32_t o7 = gauss_h7(L[4], L[5], C[4], C[5], R[4], R[5]);
                        for (int k = 0; k < 4 && x + k < w; k++) {
                            size_t o = (size_t)y * w + x + k;
                            d3[o] = (uint8_t)(o3 >> (8 * k)); d5[o] = (uint8_t)(o5 >> (8 * k)); d7[o] = (uint8_t)(o7 >> (8 * k));
                        }
                    }
                }
        }
}

// state map (0 none, 1 weak candidate, 3 strong candidate) of cv.Canny's Sobel + NMS stage.
// always_diag != 0 evaluates the diagonal test on every row (the kernel skips it per warp).
extern "C" void rh_sobel_nms(const uint8_t *src, int ch, int h, int w, int low, int high, int always_diag, uint8_t *state)
{
    const int TH = 128, OW = 120;
    const int strips_x = (w + OW - 1) / OW, strips_y = (h + TH - 1) / TH;
    // thresholds as half2 constants (clamped like canny.cu does), sector table like the kernel's shared-memory copy
    const float lf = (float)std::min(std::max(low, -1), 2047), hf = (float)std::min(std::max(high, -1), 2047);
    const h2 low1 = h2_make(lf, lf), high1 = h2_make(hf, hf);
    static uint16_t tab[SECTOR_TABLE];
    for (int i = 0; i < SECTOR_TABLE; i++) tab[i] = sector_table_entry(i);
    for (int sy = 0; sy < strips_y; sy++)
        for (int sx = 0; sx < strips_x; sx++) {
            const int y0 = sy * TH, y1 = std::min(y0 + TH, h);
            SobelRow R[32][3][3];
            MagRow M[32][3];
            Grad G[32][2];
            memset(R, 0, sizeof(R)); memset(M, 0, sizeof(M)); memset(G, 0, sizeof(G));
            const int iters = (y1 - y0) + 4;
            for (int ib = 0; ib < iters; ib += 6)
                for (int u = 0; u < 6; u++) {
                    const int it = ib + u;
                    if (it >= iters) continue;
                    const int py = y0 - 2 + it;
                    uint32_t word[32][3];
                    for (int lane = 0; lane < 32; lane++) {
                        int x = sx * OW - 4 + 4 * lane;
                        for (int c = 0; c < ch; c++)
                            word[lane][c] = load_word(src + (size_t)clampi(py, 0, h - 1) * w * ch, x, w, false, ch, c);
                    }
                    for (int lane = 0; lane < 32; lane++)
                        for (int c = 0; c < ch; c++) {
                            uint32_t lw = lane > 0 ? word[lane - 1][c] : word[lane][c];      // shfl_up keeps own value at lane 0
                            uint32_t rw = lane < 31 ? word[lane + 1][c] : word[lane][c];
                            R[lane][u % 3][c] = sobel_row(word[lane][c], lw, rw);
                        }
                    uint32_t mA[32], mB[32];
                    if (it >= 2) {
                        const int gy = py - 1;
                        for (int lane = 0; lane < 32; lane++) {
                            int x = sx * OW - 4 + 4 * lane;
                            const uint32_t cmA = ((x >= 0 && x < w) ? 0xffffu : 0u) | ((x + 1 >= 0 && x + 1 < w) ? 0xffff0000u : 0u);
                            const uint32_t cmB = ((x + 2 >= 0 && x + 2 < w) ? 0xffffu : 0u) | ((x + 3 >= 0 && x + 3 < w) ? 0xffff0000u : 0u);
                            Grad g = sobel_grad(R[lane][(u + 1) % 3][0], R[lane][(u + 2) % 3][0], R[lane][u % 3][0]);
                            h2 a, b;
                            grad_mag(g, a, b);
                            for (int c = 1; c < ch; c++) {
                                Grad gc = sobel_grad(R[lane][(u + 1) % 3][c], R[lane][(u + 2) % 3][c], R[lane][u % 3][c]);
                                grad_select(g, a, b, gc);
                            }
                            const bool row_in = gy >= 0 && gy < h;
                            mA[lane] = row_in ? (a & cmA) : 0u;
                            mB[lane] = row_in ? (b & cmB) : 0u;
                            G[lane][u % 2] = g;
                        }
                        for (int lane = 0; lane < 32; lane++) {
                            uint32_t leftB = lane > 0 ? mB[lane - 1] : mB[lane];
                            uint32_t rightA = lane < 31 ? mA[lane + 1] : mA[lane];
                            M[lane][u % 3] = mag_row(mA[lane], mB[lane], leftB, rightA);
                        }
                    }
                    if (it >= 4) {
                        const int ny = py - 2;
                        bool any = always_diag != 0;
                        bool above = always_diag != 0;               // the kernel skips rows with no magnitude above low
                        for (int lane = 0; lane < 32; lane++) above = above || any_above(M[lane][(u + 2) % 3], low1);
                        if (!above) {
                            for (int lane = 1; lane <= 30; lane++) {
                                int x = sx * OW - 4 + 4 * lane;
                                for (int k = 0; k < 4 && x + k < w && x < w; k++) state[(size_t)ny * w + x + k] = 0;
                            }
                            continue;
                        }
                        NmsPartial P[32];
                        for (int lane = 0; lane < 32; lane++) {
                            P[lane] = nms_axis(M[lane][(u + 1) % 3], M[lane][(u + 2) % 3], M[lane][u % 3], G[lane][(u + 1) % 2], low1, tab);
                            any = any || nms_needs_diag(P[lane]);
                        }
                        for (int lane = 1; lane <= 30; lane++) {
                            int x = sx * OW - 4 + 4 * lane;
                            if (x >= w) continue;
                            if (any) nms_diag(P[lane], M[lane][(u + 1) % 3], M[lane][(u + 2) % 3], M[lane][u % 3], G[lane][(u + 1) % 2]);
                            uint32_t st = nms_state(P[lane], M[lane][(u + 2) % 3], high1);
                            for (int k = 0; k < 4 && x + k < w; k++) state[(size_t)ny * w + x + k] = (uint8_t)(st >> (8 * k));
                        }
                    }
                }
        }
}

// hysteresis of the state map (flood from strong through weak, 8-connected) -> 0/255 edges
extern "C" void rh_hysteresis(const uint8_t *state, int h, int w, uint8_t *edges)
{
    std::vector<uint8_t> st(state, state + (size_t)h * w);
    std::vector<long> stack;
    for (long i = 0; i < (long)h * w; i++)
        if (st[i] == 3) stack.push_back(i);
    while (!stack.empty()) {
        long p = stack.back(); stack.pop_back();
        int y = (int)(p / w), x = (int)(p % w);
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                long q = (long)yy * w + xx;
                if (st[q] == 1) { st[q] = 3; stack.push_back(q); }
            }
    }
    for (long i = 0; i < (long)h * w; i++) edges[i] = st[i] == 3 ? 255 : 0;
}
