/* nvit_b200 tuning and measurement hooks - NOT part of the drop-in boundary (include/nvit_b200.h).
 *
 * Section 1 switches select between EQUIVALENT kernel variants (same results, tested in both settings by
 * tests/test_kernels_gpu.py); they are process-wide and exist so that the tests can pin each variant and the scripts
 * under scripts/ can A/B them.  Section 2 hooks make kernels skip work (outputs are WRONG): they are compiled only into
 * builds made with -DNVIT_BENCH_HOOKS (`python -m nvit_b200.build --hooks` -> nvit_b200/libnvit_b200_hooks.so, used by
 * scripts/gemm_bench.py and scripts/attn_phases.py); the product library does not contain them.
 */
#ifndef NVIT_B200_TUNING_H_
#define NVIT_B200_TUNING_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- 1. variant selection (results do not depend on these) ---------------------------------------------------- */
/* 0 = choose automatically (default), 1 = single-CTA 128-row tiles (cta_group::1), 2 = CTA-pair 256-row tiles
 * (cta_group::2, cluster of two SMs). */
int nvit_gemm_force_cta_group(int mode);
/* Tile order of the persistent GEMM grids: 0 = n fastest over the whole output width (default); G > 0 = bands of G tiles
 * along n (inside a band n fastest, then m), applied where an output has more than G tiles along n; -1 = automatic (bands
 * where they cut the operand rows shared by the tiles in flight by >= 10 %). */
int nvit_gemm_raster_group(int group);
/* CTA-group mode (1 or 2, default 2) of the swiglu gate GEMM under the automatic policy; 11 or 12 set the mode of the fused
 * gate-backward GEMM; 22 or 24 its number of epilogue groups. */
int nvit_gemm_swiglu_cta_group(int mode);

/* Attention forward kernel: 2 = persistent and warp-specialised (one CTA per SM claims heads from a device-wide counter; two
 * softmax warpgroups, one per 128-row q tile, one thread per score row; an MMA warp; one thread for every tile load and store;
 * two (Q, K, V) buffer sets; default; runs as 1 for q / k that are normalised inside the kernel), 1 = one head per CTA, two
 * CTAs per SM (round 1). */
int nvit_attention_fwd_variant(int variant);

/* Attention backward kernel: 2 = persistent and warp-specialised (8 compute warps + one MMA warp per SM, the products of the
 * next (kv tile, q tile) item in flight under the passes of the current one, the next head's tiles loading meanwhile;
 * default), 3 = 2 with the dV / dK / dQ epilogues on a warpgroup of their own (one thread per accumulator row; runs as 2 for
 * raw q / k with sqk and for T > 208), 1 = the single-role, one-head-per-CTA kernel of round 1. */
int nvit_attention_bwd_variant(int variant);

/* ---- 2. measurement only, -DNVIT_BENCH_HOOKS builds (outputs are WRONG while active) -------------------------------- */
#ifdef NVIT_BENCH_HOOKS
/* 1 = the GEMM epilogue returns the accumulator without reading it (main-loop-only time), 2 = it reads and converts but
 * neither stages nor stores, 3 / 4 = finer cuts of the gate-backward epilogue. */
int nvit_gemm_debug(int mode);
/* device buffer of 16384 int64 receiving clock64() / globaltimer marks of the next attention launches (NULL switches it off):
 * phase marks of the roles of every CTA of the persistent backward kernel (4th head), entry / exit time and SM id of every CTA;
 * scripts/attn_bwd_phases_v3.py and scripts/attn_bwd_cta_times.py decode it */
int nvit_attention_debug(void* dev_buf_16384_int64);
#endif

#ifdef __cplusplus
}
#endif
#endif /* NVIT_B200_TUNING_H_ */
