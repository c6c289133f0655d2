/* c_abi_demo.c - the C ABI of libhp_b200.so from a plain C99 caller (no torch, no C++).
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -L <pkg dir> -lhp_b200 -Wl,-rpath,<pkg dir> -o c_abi_demo
 *
 * Without arguments it only exercises what needs no GPU: version, workspace size, and that argument errors come back
 * as negative codes with a message instead of touching CUDA (tests/test_c_abi_from_c.py runs it on the CPU box).
 * With "--gpu" it also decodes one synthetic 64x64 map through the CUDA runtime's C API (cudaMalloc / cudaMemcpy
 * declared by hand below, so that no CUDA header is needed to build the demo).                                     */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hp_b200.h"

/* the four CUDA runtime entry points the --gpu part uses (resolved from libcudart, which libhp_b200.so links) */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind); /* 1 = H2D, 2 = D2H */
extern int cudaDeviceSynchronize(void);

static int expect(int got, int want, const char* what) {
    if (got != want) {
        fprintf(stderr, "FAIL %s: got %d, want %d (%s)\n", what, got, want, hp_last_error());
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    int bad = 0;
    float dummy[4];
    printf("hp_version = %d\n", hp_version());
    bad += hp_version() < 100;
    bad += hp_workspace_bytes(5376, 21) < 256;
    /* argument errors: negative code, message, no CUDA call */
    bad += expect(hp_argmax_decode(NULL, 1, 64, 64, dummy, dummy, NULL, NULL), HP_ERR_NULL, "null heat");
    bad += strlen(hp_last_error()) == 0;
    bad += expect(hp_argmax_decode(dummy, 1, 0, 64, dummy, dummy, NULL, NULL), HP_ERR_SHAPE, "zero height");
    bad += expect(hp_pck_finalize(NULL, 21, NULL, NULL), HP_ERR_NULL, "null counts");
    if (argc > 1 && strcmp(argv[1], "--gpu") == 0) {
        const int H = 64, W = 64;
        float* host = (float*)calloc((size_t)H * W, sizeof(float));
        float *d_map = NULL, *d_xy = NULL, *d_max = NULL, xy[2], mx;
        host[17 * W + 42] = 0.75f; /* single peak -> (x, y) = (42, 17) */
        bad += cudaMalloc((void**)&d_map, sizeof(float) * H * W) != 0;
        bad += cudaMalloc((void**)&d_xy, sizeof(float) * 2) != 0;
        bad += cudaMalloc((void**)&d_max, sizeof(float)) != 0;
        bad += cudaMemcpy(d_map, host, sizeof(float) * H * W, 1) != 0;
        bad += expect(hp_argmax_decode(d_map, 1, H, W, d_xy, d_max, NULL, NULL), HP_OK, "decode");
        bad += cudaDeviceSynchronize() != 0;
        bad += cudaMemcpy(xy, d_xy, sizeof(xy), 2) != 0;
        bad += cudaMemcpy(&mx, d_max, sizeof(mx), 2) != 0;
        printf("decoded (%g, %g) max %g\n", xy[0], xy[1], mx);
        bad += !(xy[0] == 42.0f && xy[1] == 17.0f && mx == 0.75f);
        cudaFree(d_map);
        cudaFree(d_xy);
        cudaFree(d_max);
        free(host);
    }
    printf(bad ? "FAILED\n" : "ok\n");
    return bad ? 1 : 0;
}
