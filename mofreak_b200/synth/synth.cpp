// Deterministic synthetic gray video ("moving squares") used by tests, smoke() and bench.py — the
// stand-in for decoded, gray-converted video (decode and BGR->gray sit outside the hot path,
// reference MoFREAKUtilities.cpp:380-410).  Definition:
//   background  96 + 40 sin(0.05 x) + 30 cos(0.07 y), rounded to nearest
//   noise       + (r % 7) - 3 per pixel, r from std::mt19937(seed * 7919 + t) in raster order
//   squares     W*H/2400 opaque squares drawn over it in index order; per square, from
//               std::mt19937(seed): side 3 + r%10, value r%256, vx r%7-3, vy r%7-3, x0 r%W, y0 r%H;
//               top-left at ((x0 + vx t) mod W, (y0 + vy t) mod H), wrapping around the borders.
// std::mt19937's output sequence is fixed by the C++ standard, so clips are identical everywhere.
#include <math.h>
#include <stdint.h>
#include <random>
#include <vector>

extern "C" void mfk_synth_clip(uint8_t* out, int w, int h, int n_frames, uint32_t seed, int t0) {
    struct Sq { int side, value, vx, vy, x0, y0; };
    std::mt19937 layout(seed);
    const int n_sq = (w * h) / 2400;
    std::vector<Sq> sq(n_sq);
    for (int i = 0; i < n_sq; i++) {
        sq[i].side = 3 + (int)(layout() % 10);
        sq[i].value = (int)(layout() % 256);
        sq[i].vx = (int)(layout() % 7) - 3;
        sq[i].vy = (int)(layout() % 7) - 3;
        sq[i].x0 = (int)(layout() % (uint32_t)w);
        sq[i].y0 = (int)(layout() % (uint32_t)h);
    }
    std::vector<int> bgx(w), bgy(h);
    std::vector<uint8_t> bg((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            bg[(size_t)y * w + x] = (uint8_t)lrint(96.0 + 40.0 * sin(0.05 * x) + 30.0 * cos(0.07 * y));
    for (int f = 0; f < n_frames; f++) {
        const int t = t0 + f;
        uint8_t* img = out + (size_t)f * w * h;
        std::mt19937 noise(seed * 7919u + (uint32_t)t);
        for (size_t i = 0; i < (size_t)w * h; i++) img[i] = (uint8_t)((int)bg[i] + (int)(noise() % 7) - 3);
        for (int i = 0; i < n_sq; i++) {
            const int px = (((sq[i].x0 + sq[i].vx * t) % w) + w) % w;
            const int py = (((sq[i].y0 + sq[i].vy * t) % h) + h) % h;
            for (int dy = 0; dy < sq[i].side; dy++) {
                uint8_t* row = img + (size_t)((py + dy) % h) * w;
                for (int dx = 0; dx < sq[i].side; dx++) row[(px + dx) % w] = (uint8_t)sq[i].value;
            }
        }
    }
}
