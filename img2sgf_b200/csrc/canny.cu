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

// ------------------------------------------------------------------ Sobel + NMS, register rolling
// One warp walks down a strip of 128 loaded / 120 stored columns; each lane owns one word (4 pixels)
// per row.  Rolling state per lane: three pixel rows as shifted pairs + horizontal smoothing, three
// magnitude rows with their shifted pairs, two gradient rows.  Per row: one coalesced load (three
// for RGB), two shuffles of the raw words, the packed Sobel/L1 magnitude (two pixels per
// instruction), two shuffles of the magnitudes and the branch-free packed NMS of the row two
// above (roll_cores.cuh).  The diagonal-sector test runs only when some lane of the warp has a
// diagonal candidate.  No shared memory, no barriers.
//
// The 3-channel variant also writes the greyscale plane (cv.cvtColor(.., BGR2GRAY) on the
// RGB-ordered array, img2sgf.py:153) from the words it has just loaded, so the RGB
// input is read from HBM once.
#ifndef I2S_CANNY1_MINB
#define I2S_CANNY1_MINB 6               // 80 registers, two spilled words: measured 5 % faster than 5 blocks of 86
#endif
#ifndef I2S_CANNY3_MINB
#define I2S_CANNY3_MINB 4
#endif
constexpr int HT = 128;                           // hysteresis tile edge (see hyst_tile)
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

// Q15 luma of 4 interleaved pixels (12 bytes in three words): (3735 c0 + 19235 c1 + 9798 c2 + 16384) >> 15
// with c0 the FIRST channel.  The 16-bit weights are split into two byte weights each
// (W = 256 hi + lo) so that a pixel costs two 4-way byte dot products.
__device__ __forceinline__ uint32_t grey4_from_rgb(uint32_t r0, uint32_t r1, uint32_t r2)
{
    constexpr uint32_t HI = 14u | (75u << 8) | (38u << 16), LO = 151u | (35u << 8) | (70u << 16);
    static_assert(14 * 256 + 151 == 3735 && 75 * 256 + 35 == 19235 && 38 * 256 + 70 == 9798, "luma weights");
    const uint32_t p1 = __funnelshift_r(r0, r1, 24), p2 = __funnelshift_r(r1, r2, 16);
    const uint32_t g0 = ((__dp4a(r0, HI, 0u) << 8) + __dp4a(r0, LO, 16384u)) >> 15;
    const uint32_t g1 = ((__dp4a(p1, HI, 0u) << 8) + __dp4a(p1, LO, 16384u)) >> 15;
    const uint32_t g2 = ((__dp4a(p2, HI, 0u) << 8) + __dp4a(p2, LO, 16384u)) >> 15;
    const uint32_t g3 = ((__dp4a(r2, HI << 8, 0u) << 8) + __dp4a(r2, LO << 8, 16384u)) >> 15;
    return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

template <int CH, int MINB>
__global__ void __launch_bounds__(CR_WARPS * 32, MINB) k_canny_roll(const MapSet ms, const Dims dims, uint8_t *__restrict__ state,
                                                              int spitch, size_t sstride, roll::h2 low, roll::h2 high,
                                                              int strips_x, int strips_y, int total,
                                                              uint8_t *__restrict__ tile_weak, int tiles_x,
                                                              uint8_t *__restrict__ grey, int gpitch, size_t gstride)
{
    __shared__ uint16_t s_tab[roll::SECTOR_TABLE];             // sector boundary per |dx| (roll_cores.cuh)
    for (int i = threadIdx.x; i < roll::SECTOR_TABLE; i += blockDim.x) s_tab[i] = roll::sector_table_entry(i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * CR_WARPS + (threadIdx.x >> 5);
    if (strip >= total) return;                                // warp-uniform
    const int sx = strip % strips_x, t = strip / strips_x, sy = t % strips_y, map = t / strips_y;
    const int2 wh = dims.of(map % ms.n);
    const int w = wh.x, h = wh.y;
    int ipitch;
    const uint8_t *img = ms.plane(map, ipitch);
    const bool al = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)ipitch) & 3) == 0;
    uint8_t *out = state + map * sstride;
    const bool al_out = ((reinterpret_cast<uintptr_t>(out) | (uintptr_t)spitch) & 3) == 0;
    // the state row is written up to the next multiple of 16 columns when the pitch has room: the
    // bulk copies of the hysteresis tiles and the 128-bit loads of the edge compaction then read
    // defined (zero: the magnitude is zero outside the image) bytes
    const int wlim = write_limit(w, spitch, 16);
    if (sx * CR_OW >= wlim || sy * CR_TH >= h) return;         // strip outside this image (ragged batch)
    const int x = sx * CR_OW - 4 + 4 * lane;
    const int y0 = sy * CR_TH, y1 = min(y0 + CR_TH, h);
    const bool store_lane = lane >= 1 && lane <= 30 && x < wlim;
    // magnitude is zero outside the image: per-half column masks of this lane's 4 pixels
    const uint32_t cmA = ((x >= 0 && x < w) ? 0xffffu : 0u) | ((x + 1 >= 0 && x + 1 < w) ? 0xffff0000u : 0u);
    const uint32_t cmB = ((x + 2 >= 0 && x + 2 < w) ? 0xffffu : 0u) | ((x + 3 >= 0 && x + 3 < w) ? 0xffff0000u : 0u);
    // fused greyscale output (RGB variant only)
    uint8_t *gout = (CH == 3 && grey) ? grey + map * gstride : nullptr;
    const bool al_g = ((reinterpret_cast<uintptr_t>(gout) | (uintptr_t)gpitch) & 3) == 0;
    const int glim = gout ? write_limit(w, gpitch, 4) : 0;
    const bool grey_lane = gout && lane >= 1 && lane <= 30 && x < glim;
    roll::SobelRow R[3][CH];
    roll::MagRow M[3];
    roll::Grad G[2];
    uint32_t weak_seen = 0;                                    // OR of "state byte == 1" over the rows stored
    const int iters = (y1 - y0) + 4;
    uint32_t nxt[CH];                                          // row loaded one iteration ahead of its use
    canny_load<CH>(img + (size_t)min(max(y0 - 2, 0), h - 1) * ipitch, x, w, al, nxt);
#pragma unroll 1
    for (int ib = 0; ib < iters; ib += 6) {
#pragma unroll
        for (int u = 0; u < 6; u++) {
            const int it = ib + u;
            if (it < iters) {                                  // warp-uniform
                const int py = y0 - 2 + it;                    // pixel row consumed in this iteration
                {
                    if (CH == 3 && grey_lane && py >= y0 && py < y1)
                        store4(gout + (size_t)py * gpitch, x, w, glim, al_g, grey4_from_rgb(nxt[0], nxt[1 % CH], nxt[2 % CH]));
                    uint32_t ch[CH];
                    canny_channels<CH>(nxt, ch);
                    canny_load<CH>(img + (size_t)min(max(py + 1, 0), h - 1) * ipitch, x, w, al, nxt);
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
                    roll::h2 mA, mB;
                    roll::grad_mag(g, mA, mB);
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
                    if (__any_sync(0xffffffffu, roll::any_above(c, low))) {
                        roll::NmsPartial p = roll::nms_axis(up, c, dn, g, low, s_tab);
                        if (__any_sync(0xffffffffu, roll::nms_needs_diag(p))) roll::nms_diag(p, up, c, dn, g);
                        st = roll::nms_state(p, c, high);
                    }
                    if (store_lane) {
                        weak_seen |= st & ~(st >> 1);
                        store4(out + (size_t)ny * spitch, x, w, wlim, al_out, st);
                    }
                }
            }
        }
    }
    // Hysteresis only has work where weak candidates exist: flag the 128x128 tile of this lane's pixels
    // (a 4-pixel group never straddles a tile; the strip is one tile row).  The first hysteresis pass
    // visits flagged tiles only -- crisp diagrams have whole maps without a single weak pixel.
    if (weak_seen & 0x01010101u) tile_weak[((size_t)map * strips_y + sy) * tiles_x + (x / HT)] = 1;
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

// One tile.  `tile` = (map * tiles_y + ty) * tiles_x + tx on the canvas tile grid.  Returns with all
// threads (uniform control flow).
__device__ __forceinline__ void hyst_tile(uint8_t *__restrict__ state, int spitch, size_t sstride, const Dims &dims,
                                          int n_images, int tiles_x, int tiles_y, uint8_t *dirty_in, uint8_t *dirty_out,
                                          int tile, uint8_t *s_map, uint16_t *s_q, int &s_qn, int &s_changed,
                                          int &s_ring, uint64_t &s_bar, uint32_t &phase)
{
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, bz = tile / (tiles_x * tiles_y);
    {
        const int d = dirty_in[tile];
        __syncthreads();                               // everyone has read the flag before it is cleared
        if (!d) return;
        if (threadIdx.x == 0) dirty_in[tile] = 0;      // leave the buffer clean for pass+1's writers
    }
    const int2 wh = dims.of(bz % n_images);
    const int w = wh.x, h = wh.y;
    uint8_t *img = state + bz * sstride;
    const int x0 = bx * HT, y0 = by * HT;
    const bool al = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)spitch) & 3) == 0;
    // bulk staging needs 16-byte rows and defined bytes up to the next multiple of 16 columns
    const int w16 = (w + 15) & ~15;
    const bool bulk = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)spitch) & 15) == 0 && w16 <= spitch;
    if (threadIdx.x == 0) { s_qn = 0; s_changed = 0; s_ring = 0; }
    if (bulk) {
        // in-image part of every staged row with one bulk copy per row; the rest (tiles on the image
        // border only) is zero-filled with ordinary stores to bytes the copies do not touch
        const int cxa = max(x0 - HX, 0), cxb = min(x0 + HT + HX, w16);
        const int ra = max(0, 1 - y0), rb = min(HS_H, h - y0 + 1);          // staged rows [ra, rb) lie in the image
        const int off = cxa - (x0 - HX);
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)((rb - ra) * (cxb - cxa)));
            for (int r = ra + threadIdx.x; r < rb; r += 32)
                bulk_g2s(s_map + r * HS_W + off, img + (size_t)(y0 - 1 + r) * spitch + cxa, (uint32_t)(cxb - cxa), &s_bar);
        }
        if (ra > 0 || rb < HS_H || off > 0 || cxb - (x0 - HX) < HS_W) {
            const int wa = off >> 2, wb = (cxb - (x0 - HX)) >> 2;           // staged words [wa, wb) are copied
            for (int idx = threadIdx.x; idx < HS_H * (HS_W / 4); idx += blockDim.x) {
                const int r = idx / (HS_W / 4), c = idx - r * (HS_W / 4);
                if (r < ra || r >= rb || c < wa || c >= wb) reinterpret_cast<uint32_t *>(s_map)[idx] = 0u;
            }
        }
        mbar_wait(&s_bar, phase & 1u);
        ++phase;
    } else {
        stage_tile_u8(s_map, HS_W, img, h, w, spitch, x0 - HX, y0 - 1, HS_W, HS_H, BORDER_ZERO, al);
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
    const int wlim = write_limit(w, spitch, 4);
    for (int idx = threadIdx.x; idx < HT * (HT / 4); idx += blockDim.x) {
        int ty = idx / (HT / 4), gx = (idx - ty * (HT / 4)) * 4;
        int y = y0 + ty, x = x0 + gx;
        if (y >= h || x >= w) continue;
        uint32_t v = *reinterpret_cast<const uint32_t *>(s_map + (ty + 1) * HS_W + gx + HX);
        store4(img + (size_t)y * spitch, x, w, wlim, al, v);
    }
    if (s_ring && threadIdx.x < 9) {
        int dy = threadIdx.x / 3 - 1, dx = threadIdx.x % 3 - 1;
        int ty = by + dy, tx = bx + dx;
        // neighbours inside THIS image only (the canvas of a ragged batch may be larger)
        if ((dx || dy) && ty >= 0 && ty * HT < h && tx >= 0 && tx * HT < w)
            dirty_out[(bz * tiles_y + ty) * tiles_x + tx] = 1;
    }
}

// Every pass compacts the indices of the dirty tiles (k_hyst_list) and walks that list with a grid
// sized for the machine, not for the tile count (k_hysteresis_list): a pass over mostly clean maps
// -- the usual case after the first pass, and in the first pass too for crisp diagrams -- costs two
// small launches instead of one block per tile.  (One launch per pass whose blocks walk the flags
// themselves was measured slower: every clean tile then costs a block-wide barrier.)
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

__global__ void __launch_bounds__(256) k_hysteresis_list(uint8_t *__restrict__ state, int spitch, size_t sstride,
                                                         const Dims dims, int n_images, int tiles_x, int tiles_y,
                                                         const int *__restrict__ list, const int *__restrict__ count,
                                                         uint8_t *dirty_in, uint8_t *dirty_out)
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
        hyst_tile(state, spitch, sstride, dims, n_images, tiles_x, tiles_y, dirty_in, dirty_out, list[i], s_map, s_q, s_qn,
                  s_changed, s_ring, s_bar, phase);
        __syncthreads();                                   // shared state is reused by the next tile ...
        fence_proxy_async();                               // ... whose bulk copies must not overtake this tile's accesses
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

__global__ void __launch_bounds__(256) k_state_to_edges16(const uint4 *__restrict__ state, uint4 *__restrict__ edges,
                                                          size_t quads)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = state[i];
        v.x = ((v.x >> 1) & 0x01010101u) * 255u; v.y = ((v.y >> 1) & 0x01010101u) * 255u;
        v.z = ((v.z >> 1) & 0x01010101u) * 255u; v.w = ((v.w >> 1) & 0x01010101u) * 255u;
        edges[i] = v;
    }
}

constexpr int HYST_RING = 64;                     // size of the per-pass list-counter ring, not a limit on passes

size_t canny_scratch_bytes(int maps, int h, int w)
{
    size_t tiles = (size_t)maps * cdiv(w, HT) * cdiv(h, HT);
    // two dirty-flag buffers, the dirty-tile list, one list counter per pass
    return align_up(tiles, 256) * 2 + align_up(tiles * sizeof(int), 256) + HYST_RING * sizeof(int) + 256;
}

// state: [ms.count * ms.n] planes of `spitch` bytes per row, `sstride` bytes apart; map m = k * ms.n + i
// belongs to image i (sizes, status).  `grey` (3-channel input only, may be null): fused greyscale output.
int canny_states(const MapSet &ms, const Dims &dims, int channels, uint8_t *state, int spitch, size_t sstride, int low,
                 int high, int passes, int32_t *status, void *scratch, cudaStream_t st, uint8_t *grey, int gpitch,
                 size_t gstride)
{
    const int maps = ms.count * ms.n;
    const int tiles_x = cdiv(dims.w, HT), tiles_y = cdiv(dims.h, HT);
    const size_t tiles = (size_t)maps * tiles_x * tiles_y;
    I2S_ARG(tiles < (1ull << 31));
    uint8_t *flags = (uint8_t *)scratch;                       // = the first dirty buffer of hysteresis()
    {
        ScopedSection sec(channels == 3 ? SEC_SOBEL_NMS_RGB : SEC_SOBEL_NMS, st);
        const int strips_x = cdiv((dims.w + 15) & ~15, CR_OW), strips_y = cdiv(dims.h, CR_TH);
        const long long total = (long long)maps * strips_x * strips_y;
        I2S_ARG(total < (1ll << 31));
        const unsigned blocks = (unsigned)((total + CR_WARPS - 1) / CR_WARPS);
        // thresholds as half2 constants, clamped to [-1, 2047]: magnitudes are integers in 0 .. 2040
        auto half2_of = [](int v) {
            const __half hv = __float2half((float)min(max(v, -1), 2047));
            const uint32_t b = (uint32_t)__half_as_ushort(hv);
            return (roll::h2)(b | (b << 16));
        };
        const roll::h2 low1 = half2_of(low), high1 = half2_of(high);
        I2S_CUDA(cudaMemsetAsync(flags, 0, align_up(tiles, 256) * 2, st));
        if (channels == 1)      // resident blocks per SM: tuned on the GPU (I2S_CANNY1_MINB at build time)
            k_canny_roll<1, I2S_CANNY1_MINB><<<blocks, CR_WARPS * 32, 0, st>>>(ms, dims, state, spitch, sstride, low1, high1, strips_x, strips_y,
                                                                 (int)total, flags, tiles_x, nullptr, 0, 0);
        else
            k_canny_roll<3, I2S_CANNY3_MINB><<<blocks, CR_WARPS * 32, 0, st>>>(ms, dims, state, spitch, sstride, low1, high1, strips_x, strips_y,
                                                                 (int)total, flags, tiles_x, grey, gpitch, gstride);
        I2S_CHECK_LAUNCH("k_canny_roll");
    }
    return hysteresis(state, spitch, sstride, dims, maps, ms.n, passes, status, scratch, st);
}

// The first dirty buffer already holds the tiles the first pass has to visit (written by
// k_canny_roll: tiles with weak candidates).
int hysteresis(uint8_t *state, int spitch, size_t sstride, const Dims &dims, int maps, int n_images, int passes,
               int32_t *status, void *scratch, cudaStream_t st)
{
    ScopedSection sec(SEC_HYSTERESIS, st);
    const int tx = cdiv(dims.w, HT), ty = cdiv(dims.h, HT);
    const size_t tiles = (size_t)maps * tx * ty;
    uint8_t *d0 = (uint8_t *)scratch, *d1 = d0 + align_up(tiles, 256);
    if (passes < 1) passes = 1;
    constexpr int kSmem = HS_H * HS_W + HQ * 2;
    I2S_CUDA(cudaFuncSetAttribute(k_hysteresis_list, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int *list = (int *)(d1 + align_up(tiles, 256));
    int *counts = (int *)((uint8_t *)list + align_up(tiles * sizeof(int), 256));
    I2S_CUDA(cudaMemsetAsync(counts, 0, HYST_RING * sizeof(int), st));
    const int sms = sm_count();
    for (int p = 0; p < passes; p++) {
        uint8_t *din = (p & 1) ? d1 : d0, *dout = (p & 1) ? d0 : d1;
        int *cnt = counts + p % HYST_RING;                     // the counters are a ring: re-zero a slot before reuse
        if (p >= HYST_RING) I2S_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
        k_hyst_list<<<(unsigned)min((size_t)(4 * sms), (tiles + 255) / 256), 256, 0, st>>>(din, (int)tiles, list, cnt);
        I2S_CHECK_LAUNCH("k_hyst_list");
        k_hysteresis_list<<<(unsigned)min((size_t)(4 * sms), tiles), 256, kSmem, st>>>(state, spitch, sstride, dims, n_images, tx,
                                                                                        ty, list, cnt, din, dout);
        I2S_CHECK_LAUNCH("k_hysteresis_list");
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
    const size_t cap = (size_t)sm_count() * 16;
    if (((((uintptr_t)state | (uintptr_t)edges) & 15) == 0) && (total & 15) == 0) {
        size_t quads = total / 16;
        int blocks = (int)min(cap, (quads + 255) / 256);
        k_state_to_edges16<<<blocks, 256, 0, st>>>((const uint4 *)state, (uint4 *)edges, quads);
    } else {
        int blocks = (int)min(cap, (total + 255) / 256);
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

extern "C" int i2s_canny(const uint8_t *img, int channels, int img_pitch, uint8_t *edges, int pitch, int n, int h, int w,
                         int low, int high, int hyst_passes, int32_t *status, void *ws, size_t ws_bytes, void *stream)
{
    I2S_ARG(img && edges && status && ws && n >= 0 && h > 0 && w > 0 && h < 16384 && w < 16384 &&
            (channels == 1 || channels == 3));
    if (img_pitch == 0) img_pitch = w * channels;
    if (pitch == 0) pitch = w;
    I2S_ARG(img_pitch >= w * channels && pitch >= w);
    if (n == 0) return I2S_OK;
    if (ws_bytes < i2s_canny_workspace_bytes(n, h, w)) {
        set_error("i2s_canny: workspace too small (%zu < %zu)", ws_bytes, i2s_canny_workspace_bytes(n, h, w));
        return I2S_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // the state map is built in place in `edges` and converted to 0/255 at the end
    MapSet ms = MapSet::single(img, img_pitch, h, n);
    const Dims dims = Dims::uniform(h, w);
    int rc = canny_states(ms, dims, channels, edges, pitch, (size_t)h * pitch, low, high, hyst_passes, status, ws, st,
                          nullptr, 0, 0);
    if (rc) return rc;
    return states_to_edges(edges, edges, (size_t)n * h * pitch, st);
}
