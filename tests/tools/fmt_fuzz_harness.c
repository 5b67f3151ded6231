/* fmt_fuzz_harness.c -- TEST TOOL: feeds files to the three format loaders of the batch command line (imcvt_b200/host/fmt_*.c)
 * and reads every pixel they return; built with -fsanitize=address,undefined by tests/test_cli_formats.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
uint8_t *loadPNGImageFile(const char *, int *, uint32_t *, uint32_t *);
uint8_t *loadBMPImageFile(const char *, int *, uint32_t *, uint32_t *);
uint8_t *loadQOIImageFile(const char *, int *, uint32_t *, uint32_t *);
uint8_t *loadPNMImageFile(const char *, int *, uint32_t *, uint32_t *);
int main(int argc, char **argv)
{
    for (int i = 1; i < argc; i++) {
        int rgb = 0; uint32_t h = 0, w = 0;
        uint8_t *p;
        p = loadPNGImageFile(argv[i], &rgb, &h, &w); if (p) { volatile uint8_t s = 0; for (size_t k = 0; k < (size_t)h * w * (rgb ? 3 : 1) && k < (1u << 26); k++) s += p[k]; free(p); }
        p = loadBMPImageFile(argv[i], &rgb, &h, &w); if (p) { volatile uint8_t s = 0; for (size_t k = 0; k < (size_t)h * w * (rgb ? 3 : 1) && k < (1u << 26); k++) s += p[k]; free(p); }
        p = loadQOIImageFile(argv[i], &rgb, &h, &w); if (p) { volatile uint8_t s = 0; for (size_t k = 0; k < (size_t)h * w * (rgb ? 3 : 1) && k < (1u << 26); k++) s += p[k]; free(p); }
    }
    return 0;
}
