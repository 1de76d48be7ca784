// vote_host.cpp -- CPU restatement of the tile scheme of k_vote_peaks2 (img2sgf_b200/csrc/circles.cu):
// 128x128 accumulator tiles with a 1-cell ring and a 2-cell guard band, every edge pixel within reach
// clipped to the tile by a conservative float interval of the signed radius, ONE loop over that
// interval with the vote at t = 0 taken back, column w / row h cleared, 4-neighbour peak scan.
// Test infrastructure: shows on the CPU that the scheme reproduces the reference accumulator's peaks
// exactly (the GPU parity tests show it for the kernel itself).  The kernel's approximate divide is
// replaced by an exact one here; SLACK widens the interval like the kernel does.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {
constexpr int MAX_R = 30, ACC_THR = 30;
constexpr int AT = 128, AG = 2, AS = AT + 2 + 2 * AG, AP = AS + 1;

int cv_round(float v) { return (int)nearbyintf(v); }
}  // namespace

// img: h x w grey, edges: h x w (0/255, the Canny output the votes are cast from).
// out_peaks: linear indices cy * (w + 2) + cx of the accumulator peaks, unsorted; returns their number,
// or -1 if a vote ever left the shared tile (guard band too small), -2 if the kernel's packed address
// arithmetic disagrees with the plain one -- neither must ever happen.
extern "C" int vh_vote_peaks(const uint8_t *img, const uint8_t *edges, int h, int w, float slack, int *out_peaks, int cap,
                             long long *votes_cast)
{
    struct Item { int x, y, sx, sy; };
    std::vector<Item> items;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (!edges[(size_t)y * w + x]) continue;
            const int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            const int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
#define P(yy, xx) ((int)img[(size_t)(yy) * w + (xx)])
            const int dx = (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm));
            const int dy = (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
            if (dx == 0 && dy == 0) continue;
            const float vx = (float)dx, vy = (float)dy;
            const float mag = sqrtf(vx * vx + vy * vy);
            if (mag < 1.0f) continue;
            items.push_back({x, y, cv_round(vx * 1024.0f / mag), cv_round(vy * 1024.0f / mag)});
        }
    int n = 0;
    long long cast = 0;
    std::vector<int> acc(AS * AP);
    for (int ty0 = 0; ty0 < h; ty0 += AT)
        for (int tx0 = 0; tx0 < w; tx0 += AT) {
            std::fill(acc.begin(), acc.end(), 0);
            const int cx0 = tx0 - 1 - AG, cy0 = ty0 - 1 - AG;
            const int X0 = std::max(tx0 - 1, 0), X1 = std::min(tx0 + AT, w - 1);
            const int Y0 = std::max(ty0 - 1, 0), Y1 = std::min(ty0 + AT, h - 1);
            const int rx0 = std::max(tx0 - 1 - MAX_R, 0), rx1 = std::min(tx0 + AT + MAX_R, w - 1);
            const int ry0 = std::max(ty0 - 1 - MAX_R, 0), ry1 = std::min(ty0 + AT + MAX_R, h - 1);
            for (const Item &e : items) {
                const int x = e.x, y = e.y, sx = e.sx, sy = e.sy;
                // the kernel walks whole 32x32 buckets overlapping the region and tests no pixel position:
                // the clipped range of radii of a pixel beyond the reach of the tile comes out empty
                if (x / 32 < rx0 / 32 || x / 32 > rx1 / 32 || y / 32 < ry0 / 32 || y / 32 > ry1 / 32) continue;
                float lo = -(float)MAX_R, hi = (float)MAX_R;
                if (sx != 0) {
                    const float inv = 1024.0f / (float)sx;
                    const float ta = (float)(X0 - x) * inv, tb = (float)(X1 + 1 - x) * inv;
                    lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
                } else if (x < X0 || x > X1) continue;
                if (sy != 0) {
                    const float inv = 1024.0f / (float)sy;
                    const float ta = (float)(Y0 - y) * inv, tb = (float)(Y1 + 1 - y) * inv;
                    lo = fmaxf(lo, fminf(ta, tb)); hi = fminf(hi, fmaxf(ta, tb));
                } else if (y < Y0 || y > Y1) continue;
                const int t_lo = std::max(-MAX_R, (int)floorf(lo - slack)), t_hi = std::min(MAX_R, (int)ceilf(hi + slack));
                if (t_lo > t_hi) continue;
                const int xb = (x - cx0) * 1024, yb = (y - cy0) * 1024;
                int x1 = xb + t_lo * sx, y1 = yb + t_lo * sy;
                for (int t = t_lo; t <= t_hi; t++, x1 += sx, y1 += sy) {
                    const int ly = y1 >> 10, lx = x1 >> 10;
                    if (ly < 0 || ly >= AS || lx < 0 || lx >= AS) return -1;
                    {   // the kernel's packed form of the same cell (circles.cu, "The vote loop"): both offsets in one
                        // register, U(t) = 0x80008000 + t * (sy * 65536 + sx); address by one 16-bit x 8-bit dot product
                        const uint32_t S = (uint32_t)(sy * 65536 + sx);
                        const uint32_t U = 0x80008000u + (uint32_t)t * S;
                        const uint32_t m = (U >> 8) & 0x00FC00FCu;
                        const uint32_t a0 = 4u * (uint32_t)((y - cy0 - 32) * AP + (x - cx0 - 32));
                        const uint32_t addr = a0 + (m & 0xffffu) * 1u + (m >> 16) * (uint32_t)AP;
                        if (addr != 4u * (uint32_t)(ly * AP + lx)) return -2;
                        // the paired form (vote_at2): high bytes of the halves of U(t) and U(t + 1) in one word,
                        // fraction bits cleared, 4-way byte dot products with (1, AP, 0, 0) / (0, 0, 1, AP)
                        const uint32_t V = U + S;
                        const uint32_t pm = (((U >> 8) & 0xffu) | (((U >> 24) & 0xffu) << 8) | (((V >> 8) & 0xffu) << 16) |
                                             (((V >> 24) & 0xffu) << 24)) & 0xFCFCFCFCu;
                        const uint32_t pa0 = a0 + (pm & 0xffu) + ((pm >> 8) & 0xffu) * (uint32_t)AP;
                        const uint32_t pa1 = a0 + ((pm >> 16) & 0xffu) + (pm >> 24) * (uint32_t)AP;
                        if (pa0 != addr) return -2;
                        if (t < t_hi && pa1 != 4u * (uint32_t)(((y1 + sy) >> 10) * AP + ((x1 + sx) >> 10))) return -2;
                    }
                    acc[ly * AP + lx]++;
                    cast++;
                }
                if (t_lo <= 0 && t_hi >= 0) acc[(y - cy0) * AP + (x - cx0)]--;
            }
            if (w - cx0 < AS)
                for (int ly = 0; ly < AS; ly++) acc[ly * AP + (w - cx0)] = 0;
            if (h - cy0 < AS)
                for (int lx = 0; lx < AS; lx++) acc[(h - cy0) * AP + lx] = 0;
            for (int ty = 0; ty < AT; ty++)
                for (int tx = 0; tx < AT; tx++) {
                    const int cx = tx0 + tx, cy = ty0 + ty;
                    if (cx < 1 || cy < 1 || cx >= w || cy >= h) continue;
                    const int *c = acc.data() + (ty + 1 + AG) * AP + tx + 1 + AG;
                    const int v = c[0];
                    if (v > ACC_THR && v > c[-1] && v >= c[1] && v > c[-AP] && v >= c[AP]) {
                        if (n < cap) out_peaks[n] = cy * (w + 2) + cx;
                        n++;
                    }
                }
        }
    if (votes_cast) *votes_cast = cast;
    return n;
}
