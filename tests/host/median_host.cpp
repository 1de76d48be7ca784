// median_host.cpp -- runs the median kernel's bit-sliced saturated-window verdicts (median_cores.cuh) on the CPU
// against brute-force window counts.  Test infrastructure.
#include <stdint.h>
#include <stdlib.h>
#include "../../img2sgf_b200/csrc/median_cores.cuh"

using namespace i2s;

template <int B, int RS> static int check(uint32_t seed, int density)
{
    constexpr int HX = 16, SW = MT_W + 2 * HX, SH = MT_H + 2 * RS, GW = (SW + 31) / 32;
    static uint32_t plane[SH][GW + 1];
    static uint8_t bit[SH][SW];
    uint32_t st = seed * 2654435761u + 12345u;
    for (int r = 0; r < SH; r++) {
        for (int g = 0; g <= GW; g++) plane[r][g] = 0;
        for (int c = 0; c < SW; c++) {
            st = st * 1664525u + 1013904223u;
            bit[r][c] = ((st >> 16) % 100) < (uint32_t)density;
            if (bit[r][c]) plane[r][c >> 5] |= 1u << (c & 31);
        }
    }
    constexpr int R = B / 2, KM = (B * B) / 2 + 1;
    int bad = 0;
    for (int ty = 0; ty < MT_H; ty++)
        for (int j = 0; j < MT_W / 32; j++) {
            const uint32_t got = settle_word<B, RS, HX, SH, GW>(plane, ty, j);
            for (int x = 0; x < 32; x++) {
                int cnt = 0;
                for (int dy = -R; dy <= R; dy++)
                    for (int dx = -R; dx <= R; dx++) cnt += bit[ty + RS + dy][HX + 32 * j + x + dx];
                bad += ((got >> x) & 1u) != (uint32_t)(cnt >= KM);
            }
        }
    return bad;
}

extern "C" int mh_check(int seeds)
{
    int bad = 0;
    for (int s = 0; s < seeds; s++)
        for (int d = 10; d <= 90; d += 10) {
            bad += check<3, 1>(s, d) + check<5, 2>(s, d) + check<7, 3>(s, d);
            bad += check<3, 3>(s, d) + check<5, 3>(s, d);          // the fused kernel stages a halo of 3 for every size
        }
    return bad;
}
