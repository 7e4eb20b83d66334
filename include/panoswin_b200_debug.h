/*
 * panoswin_b200_debug.h — profiling / diagnostics entry points of libpanoswin_b200.
 *
 * NOT part of the reference-facing contract (include/panoswin_b200.h) and NOT present in the product library: these
 * symbols exist only when the library is built with -DPSW_DIAGNOSTICS (`python -m
 * panoswintransformerobjectdetection_b200._build --diag` writes libpanoswin_b200_diag.so; tools/microbench.py loads
 * it).  The mode switches are process-wide mutable state, which is exactly why they are kept out of the product
 * build: there kernel selection depends on the call's arguments only.
 */
#ifndef PANOSWIN_B200_DEBUG_H_
#define PANOSWIN_B200_DEBUG_H_

#include "panoswin_b200.h"

#ifdef PSW_DIAGNOSTICS
#ifdef __cplusplus
extern "C" {
#endif

/* psw_window_attn_full_fwd plus: per-phase SM-cycle totals of CTA 0 in phase_cycles[6] (device, int64; NULL allowed):
 * {wait-for-loads, S MMA, softmax, P.V MMA, store, steps}; mode 1 = memory skeleton only (same gathers and stores,
 * no MMA / softmax; output = q rows), 2 = no bias loads, 3 = no q/k/v loads (results are then garbage);
 * variant bits [0,4): force the image pairs per unit of the batch-innermost schedule (15 = window-pair schedule).
 *  */
PSW_API int psw_diag_window_attn_full(const void* qkv, void* out, const void* bias_full, const float* qkv_bias,
                                      int B, int H, int W, int C, int heads, int window, int shift, int pano_mode,
                                      float scale, long long* phase_cycles, int mode, int variant, void* stream);
/* Process-wide switch of the bf16 GEMM kernel (0 = normal operation, returns the previous value): bit0 skip the output
 * stores, bit1 skip the operand loads, bit2 skip the MMAs (results are then garbage); bit3 skip the proxy fence, bit4
 * cycle counters; bits [8,12) cap the pipeline stage count, bits [16,25) force the tile width, bit 25 four epilogue
 * warps per group, bit 26 / 27 forbid / force CTA-pair tiles. */
PSW_API int psw_diag_linear_mode(int mode);
/* With mode bit 4 set, CTA 0 of the last bf16 GEMM launch accumulated SM-cycle totals; copies them to the HOST array
 * host_out16[16] (synchronises): {producer wait-empty, mma wait-tempty, mma wait-full, mma issue, epilogue wait-tfull,
 * tmem-ld, math+stage, store-issue, tiles}. */
PSW_API int psw_diag_linear_cycles(long long* host_out16);
/* Same for the fused MLP kernel: bit1 skips the final epilogue, bit3 the GELU arithmetic (results are then garbage). */
PSW_API int psw_diag_mlp_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* PSW_DIAGNOSTICS */
#endif /* PANOSWIN_B200_DEBUG_H_ */
