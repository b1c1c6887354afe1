// Host-side check of the inline helpers kernels and host code share (dreamscene_b200/csrc/common.cuh).
// Built with nvcc and run on the CPU by tests/test_host_cpu.py (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main() {
    // multisplit grid: <= 4096 Gaussians per CTA, every Gaussian covered, one full wave below 296 * 4096
    const long long Ps[] = {1, 255, 256, 257, 1000, 75776, 100000, 1000000, 1212416, 1212417, 2627680, 10510720};
    for (long long P : Ps) {
        const int nblk = gsr_ms_blocks(P);
        const long long per = (P + nblk - 1) / nblk;
        CHECK(nblk >= 1 && per >= 1 && per <= 4096 && per * nblk >= P);
        if (P >= 256LL * GSR_MS_WAVE_CTAS && P <= 4096LL * GSR_MS_WAVE_CTAS) CHECK(nblk == GSR_MS_WAVE_CTAS);
        if (P > 4096LL * GSR_MS_WAVE_CTAS) CHECK(nblk == (P + 4095) / 4096);
        if (P < 256) CHECK(nblk == 1);
    }
    // backward size classes: monotone, two per power of two, clamped to the last class
    int prev = 0;
    for (uint32_t n = 1; n < (1u << 20); ++n) {
        const int k = gsr_bwd_class(n);
        CHECK(k >= prev && k < GSR_BWD_CLASSES);
        prev = k;
    }
    CHECK(gsr_bwd_class(1) == 0 && gsr_bwd_class(2) == 2 && gsr_bwd_class(3) == 3 && gsr_bwd_class(4) == 4 &&
          gsr_bwd_class(6) == 5 && gsr_bwd_class(8) == 6 && gsr_bwd_class(16384) == 28 && gsr_bwd_class(24576) == 29 &&
          gsr_bwd_class(1u << 31) == GSR_BWD_CLASSES - 1);
    // tile grid
    const GsrTileGrid g = gsr_grid(1000, 1030);
    CHECK(g.gx == 65 && g.gy == 63 && g.ntiles == 65 * 63);
    CHECK(gsr_use_multisplit(GSR_MS_MAX_TILES) && !gsr_use_multisplit(GSR_MS_MAX_TILES + 1));
    std::printf(fails ? "%d checks failed\n" : "common.cuh helpers ok\n", fails);
    return fails ? 1 : 0;
}
