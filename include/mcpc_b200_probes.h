/*
 * mcpc_b200_probes.h -- C ABI of libmcpc_b200_probes.so: known-answer tests of the sm_100a primitives the kernels of
 * libmcpc_b200.so are built from (tcgen05.mma / TMEM / bulk copies / TMA tensor maps, csrc/umma.cuh + csrc/tma.cuh).
 * VALIDATION ONLY: a separate library, so that the product library carries no test kernels.  Same
 * conventions as mcpc_b200.h: device pointers, asynchronous on `stream`, 0 = success, text in mcpc_probes_last_error().
 */
#ifndef MCPC_B200_PROBES_H_
#define MCPC_B200_PROBES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* mcpc_probes_last_error(void);

/* Known-answer test of the tcgen05/TMEM/bulk-copy primitives of the bf16 path.
 * Wt [128, Kin], Bx [N, Kin], G [N, 128] -> D1 [128, N] = Wt Bx^T,  D2 [128, N]: D2[m][n] = sum_j Wt[j][m] G[n][j]
 * (rows m >= Kin undefined).  ws: >= 128*Kin*2 bytes of device scratch. */
int mcpc_debug_umma(const float* Wt, const float* Bx, const float* G, int32_t Kin, int32_t N, float* D1, float* D2,
                    void* ws, void* stream);

/* TMA (tensor-map) loads + SWIZZLE_128B operands, D [128, N] = A B^T with K = 64.
 * a_mn = 0: A is [128, 64], 1: A is stored transposed [64, 128]; b_mn likewise for B ([N, 64] / [64, N]);
 * N in {64,128,192,256}; ws: >= (128 + N) * 64 * 2 bytes of device scratch. */
int mcpc_debug_tma(const float* A, const float* B, int32_t N, int32_t a_mn, int32_t b_mn, float* D, void* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCPC_B200_PROBES_H_ */
