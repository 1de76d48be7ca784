// canny.cu -- Sobel 3x3 + L1 magnitude + non-maximum suppression + hysteresis.
// Reference call sites: cv.Canny(rgb,50,200,3,L1) img2sgf.py:162-165 (3-channel variant) and the
// Canny(50,100) that cv.HoughCircles runs on each of its inputs (img2sgf.py:180).
// Arithmetic: SURVEY.md Appendix A.4 (integer, bit-exact).
//
// State map encoding (one byte per pixel): bit0 = NMS candidate (mag > low and local max),
// bit1 = edge (strong, or weak reached from a strong one).  Final edge <=> bit1.
#include "canny.cuh"
#include "profile.cuh"
#include "tma.cuh"
#include "roll_cores.cuh"

namespace i2s {

// ------------------------------------------------------------------ Sobel + NMS
constexpr int NT = 64;                 // output tile (NT x NT), 256 threads
constexpr int NM = NT + 2;             // magnitude rows (tile + 1-px ring)

template <int CH>
__global__ void __launch_bounds__(256) k_sobel_nms(const MapSet ms, uint8_t *__restrict__ state,
                                                   int h, int w, int low, int high, bool al, bool bulk)
{
    // staged source: y halo 2; x halo 16 for the single-channel kernel (rows are then 16-byte aligned
    // bulk copies), 4 for the 3-channel one.  Gradient arrays are indexed [mr][staged column].
    constexpr int HXL = (CH == 1) ? 16 : 4;
    constexpr int SW_ = NT + 2 * HXL, SH_ = NT + 4, MP = SW_;
    __shared__ __align__(128) uint8_t s_src[SH_ * SW_ * CH];
    __shared__ __align__(16) int s_dxy[NM * MP];          // dx | dy << 16 (two's complement halves)
    __shared__ __align__(16) uint16_t s_mag[NM * MP];
    __shared__ uint64_t s_bar;
    const size_t plane = (size_t)h * w;
    const uint8_t *img = ms.plane(blockIdx.z, plane * CH);
    const int x0 = blockIdx.x * NT, y0 = blockIdx.y * NT;

    if (CH == 1) {
        stage_tile_bulk(s_src, img, h, w, x0 - HXL, y0 - 2, SW_, SH_, BORDER_REPLICATE, bulk, al, &s_bar);
    } else {
        for (int idx = threadIdx.x; idx < SH_ * SW_; idx += blockDim.x) {
            int ty = idx / SW_, tx = idx - ty * SW_;
            int y = border_index(y0 - 2 + ty, h, BORDER_REPLICATE);
            int x = border_index(x0 - HXL + tx, w, BORDER_REPLICATE);
            const uint8_t *p = img + ((size_t)y * w + x) * CH;
#pragma unroll
            for (int c = 0; c < CH; c++) s_src[idx * CH + c] = __ldg(p + c);
        }
        __syncthreads();
    }

    // gradients for rows y0-1 .. y0+NT (the tile and its 1-px ring); magnitude is 0 outside the image
    if (CH == 1) {
        // four pixels per thread from three staged rows: column sums v = t + 2m + b and column
        // differences d = b - t give dx_i = v_{i+1} - v_{i-1}, dy_i = d_{i-1} + 2 d_i + d_{i+1}
        constexpr int G = SW_ / 4, G0 = HXL / 4 - 1, GN = NT / 4 + 2;     // groups holding columns x0-4 .. x0+NT+3
        for (int idx = threadIdx.x; idx < NM * GN; idx += blockDim.x) {
            const int mr = idx / GN, g = G0 + idx - mr * GN;
            const int y = y0 - 1 + mr;
            const uint32_t *r0 = reinterpret_cast<const uint32_t *>(s_src + mr * SW_);
            const uint32_t *r1 = r0 + G, *r2 = r1 + G;
            int v[6], d[6];
            {
                const uint32_t ta = r0[g - 1], tb = r0[g], tc = r0[g + 1];
                const uint32_t ma = r1[g - 1], mb = r1[g], mc = r1[g + 1];
                const uint32_t ba = r2[g - 1], bb = r2[g], bc = r2[g + 1];
                int t, m, b;
                t = ta >> 24; m = ma >> 24; b = ba >> 24; v[0] = t + 2 * m + b; d[0] = b - t;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    t = (tb >> (8 * k)) & 0xff; m = (mb >> (8 * k)) & 0xff; b = (bb >> (8 * k)) & 0xff;
                    v[k + 1] = t + 2 * m + b; d[k + 1] = b - t;
                }
                t = tc & 0xff; m = mc & 0xff; b = bc & 0xff; v[5] = t + 2 * m + b; d[5] = b - t;
            }
            const int xg = x0 - HXL + 4 * g;
            const bool row_in = y >= 0 && y < h;
            const bool all_in = row_in && xg >= 0 && xg + 3 < w;
            int4 oxy;
            uint32_t om[2] = {0, 0};
            int *po = &oxy.x;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int dx = v[i + 2] - v[i], dy = d[i] + 2 * d[i + 1] + d[i + 2];
                const bool in = all_in || (row_in && xg + i >= 0 && xg + i < w);
                const int mg = in ? abs(dx) + abs(dy) : 0;
                po[i] = (int)((uint32_t)(dx & 0xffff) | ((uint32_t)dy << 16));
                om[i >> 1] |= (uint32_t)mg << (16 * (i & 1));
            }
            *reinterpret_cast<int4 *>(s_dxy + mr * MP + 4 * g) = oxy;
            *reinterpret_cast<uint2 *>(s_mag + mr * MP + 4 * g) = make_uint2(om[0], om[1]);
        }
    } else {
        for (int idx = threadIdx.x; idx < NM * MP; idx += blockDim.x) {
            const int mr = idx / MP, col = idx - mr * MP;
            const int x = x0 - HXL + col, y = y0 - 1 + mr;
            int bdx = 0, bdy = 0, bm = 0;
            if (col >= 1 && col < SW_ - 1 && x >= 0 && x < w && y >= 0 && y < h) {
                const uint8_t *c = s_src + ((mr + 1) * SW_ + col) * CH;   // centre sample
                constexpr int RS = SW_ * CH;
#pragma unroll
                for (int ch = 0; ch < CH; ch++) {
                    int p00 = c[-RS - CH + ch], p01 = c[-RS + ch], p02 = c[-RS + CH + ch];
                    int p10 = c[-CH + ch], p12 = c[CH + ch];
                    int p20 = c[RS - CH + ch], p21 = c[RS + ch], p22 = c[RS + CH + ch];
                    int dx = (p02 + 2 * p12 + p22) - (p00 + 2 * p10 + p20);
                    int dy = (p20 + 2 * p21 + p22) - (p00 + 2 * p01 + p02);
                    int m = abs(dx) + abs(dy);
                    if (ch == 0 || m > bm) { bm = m; bdx = dx; bdy = dy; }
                }
            }
            s_mag[idx] = (uint16_t)bm;
            s_dxy[idx] = (int)((uint32_t)(bdx & 0xffff) | ((uint32_t)bdy << 16));
        }
    }
    __syncthreads();

    for (int idx = threadIdx.x; idx < NT * (NT / 4); idx += blockDim.x) {
        int ty = idx / (NT / 4), gx = (idx - ty * (NT / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c = (ty + 1) * MP + gx + k + HXL;
            int m = s_mag[c];
            uint32_t st = 0;
            if (m > low) {
                int d = s_dxy[c];
                int xs = (int)(short)(d & 0xffff), ys = d >> 16;
                int ax = abs(xs), ay = abs(ys) << 15;
                int t22 = ax * 13573;
                bool keep;
                if (ay < t22) keep = m > s_mag[c - 1] && m >= s_mag[c + 1];
                else {
                    int t67 = t22 + (ax << 16);
                    if (ay > t67) keep = m > s_mag[c - MP] && m >= s_mag[c + MP];
                    else {
                        int s = (xs ^ ys) < 0 ? -1 : 1;
                        keep = m > s_mag[c - MP - s] && m > s_mag[c + MP + s];
                    }
                }
                if (keep) st = m > high ? 3u : 1u;
            }
            packed |= st << (8 * k);
        }
        size_t o = blockIdx.z * plane + (size_t)y * w + x;
        if (al && x + 3 < w) *reinterpret_cast<uint32_t *>(state + o) = packed;
        else
            for (int k = 0; k < 4 && x + k < w; k++) state[o + k] = (uint8_t)(packed >> (8 * k));
    }
}

// ------------------------------------------------------------------ Sobel + NMS, register rolling
// One warp walks down a strip of 128 loaded / 120 stored columns; each lane owns one word (4 pixels)
// per row.  Rolling state per lane: three pixel rows as shifted pairs + horizontal smoothing, three
// magnitude rows with their shifted pairs, two gradient rows.  Per row: one coalesced load (three
// for RGB), two shuffles of the raw words, the packed Sobel/L1 magnitude (two pixels per
// instruction), two shuffles of the magnitudes and the branch-free packed NMS of the row two
// above (roll_cores.cuh).  The diagonal-sector test runs only when some lane of the warp has a
// diagonal candidate.  No shared memory, no barriers.
constexpr int HT = 128;                           // hysteresis tile edge (see k_hysteresis)
constexpr int CR_TH = HT, CR_OW = 120, CR_WARPS = 4;    // a strip spans exactly one row of hysteresis tiles

// Raw words of one row for one lane: CH words holding the 4 pixels of this lane (CH = 3: the 12
// interleaved RGB bytes).  Kept raw so that the row loaded one iteration ahead is not touched -- not
// even by the de-interleaving permutes -- before the iteration that consumes it.
template <int CH>
__device__ __forceinline__ void canny_load(const uint8_t *__restrict__ row, int x, int w, bool al, uint32_t (&raw)[CH])
{
    if (al && x >= 0 && x + 3 < w) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + (size_t)x * CH);
#pragma unroll
        for (int c = 0; c < CH; c++) raw[c] = __ldg(p + c);
        return;
    }
#pragma unroll
    for (int c = 0; c < CH; c++) raw[c] = 0;
    if (x >= w + 8 || x < -8) return;                          // beyond any halo: value never used
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint8_t *p = row + (size_t)min(max(x + k, 0), w - 1) * CH;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const int b = k * CH + c;                          // byte position inside the CH words
            raw[b >> 2] |= (uint32_t)__ldg(p + c) << (8 * (b & 3));
        }
    }
}

// channel words (4 pixels of one channel each) from the raw words
template <int CH>
__device__ __forceinline__ void canny_channels(const uint32_t (&raw)[CH], uint32_t (&ch)[CH])
{
    if (CH == 1) { ch[0] = raw[0]; return; }
    const uint32_t w0 = raw[0], w1 = raw[1 % CH], w2 = raw[2 % CH];
    ch[0] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
    ch[1 % CH] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
    ch[2 % CH] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
}

template <int CH, int MINB>
__global__ void __launch_bounds__(CR_WARPS * 32, MINB) k_canny_roll(const MapSet ms, uint8_t *__restrict__ state, int h, int w,
                                                              uint32_t low1, uint32_t high1, bool al, int strips_x,
                                                              int strips_y, int total, uint8_t *__restrict__ tile_weak)
{
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * CR_WARPS + (threadIdx.x >> 5);
    if (strip >= total) return;                                // warp-uniform
    const int sx = strip % strips_x, t = strip / strips_x, sy = t % strips_y, map = t / strips_y;
    const size_t plane = (size_t)h * w;
    const uint8_t *img = ms.plane(map, plane * CH);
    uint8_t *out = state + map * plane;
    const int x = sx * CR_OW - 4 + 4 * lane;
    const int y0 = sy * CR_TH, y1 = min(y0 + CR_TH, h);
    const bool store_lane = lane >= 1 && lane <= 30 && x < w;
    // magnitude is zero outside the image: per-half column masks of this lane's 4 pixels
    const uint32_t cmA = ((x >= 0 && x < w) ? 0xffffu : 0u) | ((x + 1 >= 0 && x + 1 < w) ? 0xffff0000u : 0u);
    const uint32_t cmB = ((x + 2 >= 0 && x + 2 < w) ? 0xffffu : 0u) | ((x + 3 >= 0 && x + 3 < w) ? 0xffff0000u : 0u);
    roll::SobelRow R[3][CH];
    roll::MagRow M[3];
    roll::Grad G[2];
    uint32_t weak_seen = 0;                                    // OR of "state byte == 1" over the rows stored
    const int iters = (y1 - y0) + 4;
    uint32_t nxt[CH];                                          // row loaded one iteration ahead of its use
    canny_load<CH>(img + (size_t)min(max(y0 - 2, 0), h - 1) * w * CH, x, w, al, nxt);
#pragma unroll 1
    for (int ib = 0; ib < iters; ib += 6) {
#pragma unroll
        for (int u = 0; u < 6; u++) {
            const int it = ib + u;
            if (it < iters) {                                  // warp-uniform
                const int py = y0 - 2 + it;                    // pixel row consumed in this iteration
                {
                    uint32_t ch[CH];
                    canny_channels<CH>(nxt, ch);
                    canny_load<CH>(img + (size_t)min(max(py + 1, 0), h - 1) * w * CH, x, w, al, nxt);
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        const uint32_t lw = __shfl_up_sync(0xffffffffu, ch[c], 1);
                        const uint32_t rw = __shfl_down_sync(0xffffffffu, ch[c], 1);
                        R[u % 3][c] = roll::sobel_row(ch[c], lw, rw);
                    }
                }
                if (it >= 2) {                                 // gradient + magnitude of row gy = py - 1
                    const int gy = py - 1;
                    roll::Grad g = roll::sobel_grad(R[(u + 1) % 3][0], R[(u + 2) % 3][0], R[u % 3][0]);
                    uint32_t mA = g.axA + g.ayA, mB = g.axB + g.ayB;
#pragma unroll
                    for (int c = 1; c < CH; c++) {
                        const roll::Grad gc = roll::sobel_grad(R[(u + 1) % 3][c], R[(u + 2) % 3][c], R[u % 3][c]);
                        roll::grad_select(g, mA, mB, gc);
                    }
                    const bool row_in = gy >= 0 && gy < h;
                    mA = row_in ? (mA & cmA) : 0u;
                    mB = row_in ? (mB & cmB) : 0u;
                    const uint32_t leftB = __shfl_up_sync(0xffffffffu, mB, 1);
                    const uint32_t rightA = __shfl_down_sync(0xffffffffu, mA, 1);
                    M[u % 3] = roll::mag_row(mA, mB, leftB, rightA);
                    G[u % 2] = g;
                }
                if (it >= 4) {                                 // NMS of row ny = py - 2
                    const int ny = py - 2;
                    const roll::MagRow &up = M[(u + 1) % 3], &c = M[(u + 2) % 3], &dn = M[u % 3];
                    const roll::Grad &g = G[(u + 1) % 2];
                    // rows without a single magnitude above `low` in the whole warp (blank paper, and most
                    // of a median-filtered map, whose thin grid lines are gone) skip the NMS arithmetic
                    uint32_t st = 0;
                    if (__any_sync(0xffffffffu, roll::any_above(c, low1))) {
                        roll::NmsPartial p = roll::nms_axis(up, c, dn, g, low1);
                        if (__any_sync(0xffffffffu, roll::nms_needs_diag(p, c, low1))) roll::nms_diag(p, up, c, dn, g);
                        st = roll::nms_state(p, c, high1);
                    }
                    if (store_lane) {
                        weak_seen |= st & ~(st >> 1);
                        const size_t o = (size_t)ny * w + x;
                        if (al) *reinterpret_cast<uint32_t *>(out + o) = st;
                        else
                            for (int k = 0; k < 4 && x + k < w; k++) out[o + k] = (uint8_t)(st >> (8 * k));
                    }
                }
            }
        }
    }
    // Hysteresis only has work where weak candidates exist: flag the 128x128 tile of this lane's pixels
    // (a 4-pixel group never straddles a tile; the strip is one tile row).  k_hysteresis pass 0 visits
    // flagged tiles only -- crisp diagrams have whole maps without a single weak pixel.
    if (tile_weak && (weak_seen & 0x01010101u))
        tile_weak[((size_t)map * strips_y + sy) * ((w + HT - 1) / HT) + (x / HT)] = 1;
}

// ------------------------------------------------------------------ hysteresis
// One block owns a HT x HT tile.  The tile plus a 1-px ring is staged in shared memory;
// every edge pixel (ring included) seeds a level-synchronous flood over 8-neighbours that
// promotes candidates INSIDE the tile.  Each pixel enters the queue at most once, so the
// queue never exceeds the staged area.  A tile whose outermost interior ring changed marks
// its 8 neighbours dirty for the next pass; passes repeat until no tile is dirty.
constexpr int HX = 16;                            // staged x halo: rows are 16-byte aligned bulk copies
constexpr int HS_W = HT + 2 * HX, HS_H = HT + 2;  // y halo 1
constexpr int HQ = (HT + 2) * (HT + 2);

// One tile.  `tile` = (map * tiles_y + ty) * tiles_x + tx.  Returns with all threads (uniform control flow).
__device__ __forceinline__ void hyst_tile(uint8_t *__restrict__ state, int h, int w, int tiles_x, int tiles_y,
                                          uint8_t *dirty_in, uint8_t *dirty_out, int check_dirty, bool al, bool bulk,
                                          int tile, uint8_t *s_map, uint16_t *s_q, int &s_qn, int &s_changed,
                                          int &s_ring, uint64_t &s_bar, uint32_t *phase)
{
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, bz = tile / (tiles_x * tiles_y);
    if (check_dirty) {
        const int d = dirty_in[tile];
        __syncthreads();                               // everyone has read the flag before it is cleared
        if (!d) return;
        if (threadIdx.x == 0) dirty_in[tile] = 0;      // leave the buffer clean for pass+1's writers
    }
    const size_t plane = (size_t)h * w;
    uint8_t *img = state + bz * plane;
    const int x0 = bx * HT, y0 = by * HT;
    if (threadIdx.x == 0) { s_qn = 0; s_changed = 0; s_ring = 0; }
    if (bulk) {
        // in-image part of every staged row with one bulk copy per row; the rest (tiles on the image
        // border only) is zero-filled with ordinary stores to bytes the copies do not touch
        const int cxa = max(x0 - HX, 0), cxb = min(x0 + HT + HX, w);
        const int ra = max(0, 1 - y0), rb = min(HS_H, h - y0 + 1);          // staged rows [ra, rb) lie in the image
        const int off = cxa - (x0 - HX);
        if (!phase) {                                  // single-tile kernel: the barrier is used once
            if (threadIdx.x == 0) mbar_init(&s_bar, 1);
            __syncthreads();
        }
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)((rb - ra) * (cxb - cxa)));
            for (int r = ra + threadIdx.x; r < rb; r += 32)
                bulk_g2s(s_map + r * HS_W + off, img + (size_t)(y0 - 1 + r) * w + cxa, (uint32_t)(cxb - cxa), &s_bar);
        }
        if (ra > 0 || rb < HS_H || off > 0 || cxb - (x0 - HX) < HS_W) {
            const int wa = off >> 2, wb = (cxb - (x0 - HX)) >> 2;           // staged words [wa, wb) are copied
            for (int idx = threadIdx.x; idx < HS_H * (HS_W / 4); idx += blockDim.x) {
                const int r = idx / (HS_W / 4), c = idx - r * (HS_W / 4);
                if (r < ra || r >= rb || c < wa || c >= wb) reinterpret_cast<uint32_t *>(s_map)[idx] = 0u;
            }
        }
        mbar_wait(&s_bar, phase ? (*phase & 1u) : 0u);
        if (phase) ++*phase;
    } else {
        stage_tile_u8(s_map, HS_W, img, h, w, x0 - HX, y0 - 1, HS_W, HS_H, BORDER_ZERO, al);
    }
    __syncthreads();
    // seeds: interior candidates that already touch an edge pixel (of the tile or of the ring).  Scanning
    // from the weak side keeps the queue tiny: most edge pixels have no weak neighbour at all.
    for (int idx = threadIdx.x; idx < HT * (HT / 4); idx += blockDim.x) {
        const int ly = idx / (HT / 4) + 1, wx = idx % (HT / 4) + HX / 4;     // staged word wx holds lx = 4wx-(HX-1) ..
        const uint32_t v = *reinterpret_cast<const uint32_t *>(s_map + ly * HS_W + 4 * wx);
        uint32_t weak = v & ~(v >> 1) & 0x01010101u;                         // bit0 set, bit1 clear: weak candidate
        while (weak) {
            const int k = (__ffs(weak) - 1) >> 3;
            weak &= ~(1u << (8 * k));
            const int b = ly * HS_W + 4 * wx + k;
            const uint8_t *c = s_map + b;
            const uint32_t nb = c[-HS_W - 1] | c[-HS_W] | c[-HS_W + 1] | c[-1] | c[1] | c[HS_W - 1] | c[HS_W] | c[HS_W + 1];
            if (!(nb & 2)) continue;
            uint32_t *word = reinterpret_cast<uint32_t *>(s_map + (b & ~3));
            const int sh = 8 * (b & 3);
            const uint32_t old = atomicOr(word, 2u << sh);
            if (((old >> sh) & 3u) == 1u) {
                const int lx = 4 * wx + k - (HX - 1);
                s_q[atomicAdd(&s_qn, 1)] = (uint16_t)(ly * (HT + 2) + lx);
                s_changed = 1;
                if (ly == 1 || ly == HT || lx == 1 || lx == HT) s_ring = 1;
            }
        }
    }
    int head = 0;
    while (true) {
        __syncthreads();
        int tail = s_qn;
        __syncthreads();
        if (head >= tail) break;
        for (int i = head + threadIdx.x; i < tail; i += blockDim.x) {
            int p = s_q[i];
            int ly = p / (HT + 2), lx = p - ly * (HT + 2);
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    if (dx == 0 && dy == 0) continue;
                    int ny = ly + dy, nx = lx + dx;
                    if (ny < 1 || ny > HT || nx < 1 || nx > HT) continue;   // promote interior only
                    int b = ny * HS_W + nx + HX - 1;
                    if ((s_map[b] & 3) != 1) continue;
                    uint32_t *word = reinterpret_cast<uint32_t *>(s_map + (b & ~3));
                    int sh = 8 * (b & 3);
                    uint32_t old = atomicOr(word, 2u << sh);
                    if (((old >> sh) & 3u) == 1u) {
                        s_q[atomicAdd(&s_qn, 1)] = (uint16_t)(ny * (HT + 2) + nx);
                        s_changed = 1;
                        if (ny == 1 || ny == HT || nx == 1 || nx == HT) s_ring = 1;
                    }
                }
        }
        head = tail;
    }
    if (!s_changed) return;
    for (int idx = threadIdx.x; idx < HT * (HT / 4); idx += blockDim.x) {
        int ty = idx / (HT / 4), gx = (idx - ty * (HT / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        uint32_t v = *reinterpret_cast<const uint32_t *>(s_map + (ty + 1) * HS_W + gx + HX);
        size_t o = (size_t)y * w + x;
        if (al && x + 3 < w) *reinterpret_cast<uint32_t *>(img + o) = v;
        else
            for (int k = 0; k < 4 && x + k < w; k++) img[o + k] = (uint8_t)(v >> (8 * k));
    }
    if (s_ring && threadIdx.x < 9) {
        int dy = threadIdx.x / 3 - 1, dx = threadIdx.x % 3 - 1;
        int ty = by + dy, tx = bx + dx;
        if ((dx || dy) && ty >= 0 && ty < tiles_y && tx >= 0 && tx < tiles_x)
            dirty_out[(bz * tiles_y + ty) * tiles_x + tx] = 1;
    }
}

// One block per tile.  (Several tiles per block on the later, mostly clean passes saves block
// launches but serialises the dirty tiles, and any work ahead of the dirty check is paid by every
// clean tile -- both measured slower.)
__global__ void __launch_bounds__(256) k_hysteresis(uint8_t *__restrict__ state, int h, int w, int tiles_x,
                                                    int tiles_y, uint8_t *dirty_in, uint8_t *dirty_out, int check_dirty,
                                                    bool al, bool bulk)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint8_t *s_map = s_dyn;                                              // HS_H * HS_W bytes
    uint16_t *s_q = reinterpret_cast<uint16_t *>(s_dyn + HS_H * HS_W);   // HQ entries
    __shared__ int s_qn, s_changed, s_ring;
    __shared__ uint64_t s_bar;
    hyst_tile(state, h, w, tiles_x, tiles_y, dirty_in, dirty_out, check_dirty, al, bulk, blockIdx.x, s_map, s_q, s_qn, s_changed,
              s_ring, s_bar, nullptr);
}

// Sparse passes: k_hyst_list compacts the indices of the dirty tiles, k_hysteresis_list walks that
// list with a grid sized for the machine, not for the tile count -- a pass over mostly clean maps
// (the usual case after pass 0, and in pass 0 too for crisp diagrams) costs two small launches
// instead of one block per tile.
__global__ void __launch_bounds__(256) k_hyst_list(const uint8_t *__restrict__ dirty, int tiles, int *__restrict__ list,
                                                   int *count)
{
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < tiles; base += gridDim.x * blockDim.x) {
        const int t = base + lane;
        const bool d = t < tiles && dirty[t];
        const uint32_t m = __ballot_sync(0xffffffffu, d);
        if (m == 0) continue;
        int at = 0;
        if (lane == 0) at = atomicAdd(count, __popc(m));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (d) list[at + __popc(m & ((1u << lane) - 1u))] = t;
    }
}

__global__ void __launch_bounds__(256) k_hysteresis_list(uint8_t *__restrict__ state, int h, int w, int tiles_x, int tiles_y,
                                                         const int *__restrict__ list, const int *__restrict__ count,
                                                         uint8_t *dirty_in, uint8_t *dirty_out, bool al, bool bulk)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint8_t *s_map = s_dyn;
    uint16_t *s_q = reinterpret_cast<uint16_t *>(s_dyn + HS_H * HS_W);
    __shared__ int s_qn, s_changed, s_ring;
    __shared__ uint64_t s_bar;
    const int n = *count;
    if ((int)blockIdx.x >= n) return;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    uint32_t phase = 0;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        hyst_tile(state, h, w, tiles_x, tiles_y, dirty_in, dirty_out, 1, al, bulk, list[i], s_map, s_q, s_qn, s_changed, s_ring,
                  s_bar, &phase);
        __syncthreads();                                   // shared state is reused by the next tile
    }
}

// after the last pass: any tile still dirty => that map did not converge
__global__ void k_hyst_check(const uint8_t *dirty, int tiles_per_map, int n_images, int32_t *status)
{
    int map = blockIdx.x;
    int any = 0;
    for (int t = threadIdx.x; t < tiles_per_map; t += blockDim.x) any |= dirty[(size_t)map * tiles_per_map + t];
    any = __syncthreads_or(any);
    if (any && threadIdx.x == 0) atomicOr(status + map % n_images, I2S_ST_HYST_NOT_CONVERGED);
}

__global__ void __launch_bounds__(256) k_state_to_edges(const uint8_t *__restrict__ state, uint8_t *__restrict__ edges,
                                                        size_t total)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        edges[i] = (state[i] & 2) ? 255 : 0;
}

__global__ void __launch_bounds__(256) k_state_to_edges4(const uint32_t *__restrict__ state, uint32_t *__restrict__ edges,
                                                         size_t words)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t v = (state[i] >> 1) & 0x01010101u;
        edges[i] = v * 255u;
    }
}

constexpr int HYST_MAX_PASSES = 64;               // size of the per-pass list-counter ring, not a limit on passes

size_t canny_scratch_bytes(int maps, int h, int w)
{
    size_t tiles = (size_t)maps * cdiv(w, HT) * cdiv(h, HT);
    // two dirty-flag buffers, the dirty-tile list, one list counter per pass
    return align_up(tiles, 256) * 2 + align_up(tiles * sizeof(int), 256) + HYST_MAX_PASSES * sizeof(int) + 256;
}

// state: [ms.count * ms.n][h][w], map m = k * ms.n + i belongs to image i (for status)
int canny_states(const MapSet &ms, int channels, uint8_t *state, int h, int w, int low, int high, int passes,
                 int32_t *status, void *scratch, cudaStream_t st)
{
    const int maps = ms.count * ms.n;
    bool al = (w & 3) == 0 && ((uintptr_t)state & 3) == 0 && ms.aligned4();
    const size_t tiles = (size_t)maps * cdiv(w, HT) * cdiv(h, HT);
    uint8_t *flags = (uint8_t *)scratch;                       // = the first dirty buffer of hysteresis()
    bool flagged = false;
    {
    ScopedSection sec(SEC_SOBEL_NMS, st);
    if (legacy_enabled("sobel")) {
        dim3 g1(cdiv(w, NT), cdiv(h, NT), maps);
        bool bulk = (w & 15) == 0 && ms.aligned16();
        if (channels == 1) k_sobel_nms<1><<<g1, 256, 0, st>>>(ms, state, h, w, low, high, al, bulk);
        else k_sobel_nms<3><<<g1, 256, 0, st>>>(ms, state, h, w, low, high, al, false);
    } else {
        const int strips_x = cdiv(w, CR_OW), strips_y = cdiv(h, CR_TH);
        const long long total = (long long)maps * strips_x * strips_y;
        I2S_ARG(total < (1ll << 31));
        const unsigned blocks = (unsigned)((total + CR_WARPS - 1) / CR_WARPS);
        // "m > low" as "m >= low + 1" on 16-bit halves; magnitudes never exceed 2040
        const uint32_t l1 = (uint32_t)min(max(low + 1, 0), 0xffff), h1 = (uint32_t)min(max(high + 1, 0), 0xffff);
        const uint32_t low1 = l1 | (l1 << 16), high1 = h1 | (h1 << 16);
        flagged = !legacy_enabled("hystall");
        if (flagged) I2S_CUDA(cudaMemsetAsync(flags, 0, align_up(tiles, 256) * 2, st));
        uint8_t *tw = flagged ? flags : nullptr;
        if (channels == 1) {
            // 5 resident blocks (no spills) measured 2 % faster than 6 (80 registers, a few spilled words)
            if (legacy_enabled("canny6"))
                k_canny_roll<1, 6><<<blocks, CR_WARPS * 32, 0, st>>>(ms, state, h, w, low1, high1, al, strips_x, strips_y, (int)total, tw);
            else
                k_canny_roll<1, 5><<<blocks, CR_WARPS * 32, 0, st>>>(ms, state, h, w, low1, high1, al, strips_x, strips_y, (int)total, tw);
        } else {
            k_canny_roll<3, 4><<<blocks, CR_WARPS * 32, 0, st>>>(ms, state, h, w, low1, high1, al, strips_x, strips_y, (int)total, tw);
        }
    }
    I2S_CHECK_LAUNCH("k_sobel_nms");
    }
    return hysteresis(state, maps, ms.n, h, w, passes, status, scratch, st, flagged);
}

// `tiles_flagged`: the first dirty buffer already holds the tiles pass 0 has to visit (written by
// k_canny_roll: tiles with weak candidates); otherwise pass 0 visits every tile.
int hysteresis(uint8_t *state, int maps, int n_images, int h, int w, int passes, int32_t *status,
               void *scratch, cudaStream_t st, bool tiles_flagged)
{
    bool al = (w & 3) == 0 && ((uintptr_t)state & 3) == 0;
    bool bulk = (w & 15) == 0 && ((uintptr_t)state & 15) == 0 && !legacy_enabled("hyst");
    ScopedSection sec(SEC_HYSTERESIS, st);
    int tx = cdiv(w, HT), ty = cdiv(h, HT);
    size_t tiles = (size_t)maps * tx * ty;
    uint8_t *d0 = (uint8_t *)scratch, *d1 = d0 + align_up(tiles, 256);
    if (!tiles_flagged) I2S_CUDA(cudaMemsetAsync(d0, 0, align_up(tiles, 256) * 2, st));
    if (passes < 1) passes = 1;
    I2S_ARG(tiles < (1ull << 31));
    constexpr int kSmem = HS_H * HS_W + HQ * 2;
    static bool attr_done = false;
    if (!attr_done) {
        I2S_CUDA(cudaFuncSetAttribute(k_hysteresis, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_done = true;
    }
    int *list = (int *)(d1 + align_up(tiles, 256));
    int *counts = (int *)((uint8_t *)list + align_up(tiles * sizeof(int), 256));
    const bool sparse = !legacy_enabled("hystdense");
    if (sparse) I2S_CUDA(cudaMemsetAsync(counts, 0, HYST_MAX_PASSES * sizeof(int), st));
    static bool attr2_done = false;
    if (!attr2_done) {
        I2S_CUDA(cudaFuncSetAttribute(k_hysteresis_list, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr2_done = true;
    }
    for (int p = 0; p < passes; p++) {
        uint8_t *din = (p & 1) ? d1 : d0, *dout = (p & 1) ? d0 : d1;
        const bool check = p > 0 || tiles_flagged;
        if (check && sparse) {
            int *cnt = counts + p % HYST_MAX_PASSES;           // the counters are a ring: re-zero a slot before reuse
            if (p >= HYST_MAX_PASSES) I2S_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
            k_hyst_list<<<(unsigned)min((size_t)592, (tiles + 255) / 256), 256, 0, st>>>(din, (int)tiles, list, cnt);
            I2S_CHECK_LAUNCH("k_hyst_list");
            k_hysteresis_list<<<(unsigned)min((size_t)(148 * 4), tiles), 256, kSmem, st>>>(state, h, w, tx, ty, list, cnt, din,
                                                                                            dout, al, bulk);
        } else {
            k_hysteresis<<<(unsigned)tiles, 256, kSmem, st>>>(state, h, w, tx, ty, din, dout, check ? 1 : 0, al, bulk);
        }
        I2S_CHECK_LAUNCH("k_hysteresis");
    }
    uint8_t *last = (passes & 1) ? d1 : d0;   // buffer written by the final pass
    k_hyst_check<<<maps, 128, 0, st>>>(last, tx * ty, n_images, status);
    I2S_CHECK_LAUNCH("k_hyst_check");
    return I2S_OK;
}

int states_to_edges(const uint8_t *state, uint8_t *edges, size_t total, cudaStream_t st)
{
    if (total == 0) return I2S_OK;
    ScopedSection sec(SEC_STATE_TO_EDGES, st);
    if (((((uintptr_t)state | (uintptr_t)edges) & 3) == 0) && (total & 3) == 0) {
        size_t words = total / 4;
        int blocks = (int)min((size_t)148 * 16, (words + 255) / 256);
        k_state_to_edges4<<<blocks, 256, 0, st>>>((const uint32_t *)state, (uint32_t *)edges, words);
    } else {
        int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
        k_state_to_edges<<<blocks, 256, 0, st>>>(state, edges, total);
    }
    I2S_CHECK_LAUNCH("k_state_to_edges");
    return I2S_OK;
}

}  // namespace i2s

using namespace i2s;

extern "C" size_t i2s_canny_workspace_bytes(int n, int h, int w)
{
    if (n <= 0 || h <= 0 || w <= 0) return 256;
    return canny_scratch_bytes(n, h, w) + 256;
}

extern "C" int i2s_canny(const uint8_t *img, int channels, uint8_t *edges, int n, int h, int w, int low, int high,
                         int hyst_passes, int32_t *status, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(img && edges && status && ws && n >= 0 && h > 0 && w > 0 && (channels == 1 || channels == 3));
    if (n == 0) return I2S_OK;
    if (ws_bytes < i2s_canny_workspace_bytes(n, h, w)) {
        set_error("i2s_canny: workspace too small (%zu < %zu)", ws_bytes, i2s_canny_workspace_bytes(n, h, w));
        return I2S_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // the state map is built in place in `edges` and converted to 0/255 at the end
    MapSet ms = MapSet::single(img, n);
    int rc = canny_states(ms, channels, edges, h, w, low, high, hyst_passes, status, ws, st);
    if (rc) return rc;
    return states_to_edges(edges, edges, (size_t)n * h * w, st);
}
