/* b200_rmsd_debug.h -- test hooks, timing experiments and peak probes of the B200 RMSD library.
 *
 * NOT part of the drop-in boundary (that is b200_rmsd.h): nothing in cpptraj calls these.  They exist for
 * tests/ (layer-by-layer checks of the tcgen05 path), bench.py (roofline denominators measured live) and
 * tools/ (kernel experiments).
 */
#ifndef B200_RMSD_DEBUG_H
#define B200_RMSD_DEBUG_H
#include "b200_rmsd.h"

#ifdef __cplusplus
extern "C" {
#endif

/* MMA CTA group of the tcgen05 int8 kernel: 2 (default) = CTA pairs, tcgen05.mma.cta_group::2,
 * 28 x 28 frame-pair tiles; 1 = single-CTA MMAs, 14 x 28 tiles.  Same results; a tuning/test knob. */
int b200_set_i8_cta_group(int ctaGroup);
int b200_get_i8_cta_group(void);
/* Test hook for the tcgen05 path (small inputs, device 0): returns the packed int8
 * operand image, the per-frame G, the raw integer covariances (9 doubles per (i,j),
 * i<j, at (i*nFrames+j)*9) and the triangle.  Any output pointer may be NULL. */
int b200_debug_i8(const float* crd, size_t frameStrideFloats, int nFrames,
                  const int* atomIdx, int nAtoms, const double* mass,
                  unsigned char* imageOut, size_t imageCap, size_t* imageBytes,
                  double* GOut, double* SOut, float* outTri, int* qsOut);

/* Timing experiments: out == NULL arms per-CTA cycle counters (16 per CTA) that the next
 * tcgen05 pair launches fill in; a later call with out != NULL copies them back and disarms. */
int b200_debug_i8_clocks(long long* out, int ctas);

/* Tuning knob: PTX shape used for the FP64 MMAs of the pair kernel
 * (0 m8n8k4, 1 m16n8k4, 2 m16n8k8, 3 m16n8k16; all lower to DMMA.8x8x4 SASS). */
int b200_set_mma_variant(int variant);

/* Measures this device's FP64 tensor (DMMA) issue peak with a register-only
 * mma.sync loop; returns TFLOP/s (<=0 on error).  Used as the roofline
 * denominator for the pair-tile kernel because MEASURED_PEAKS.json holds no
 * FP64 figure. */
double b200_measure_fp64_mma_peak(int variant);
/* Same for the tcgen05 kind::i8 pipe (M128 x N256 x K32 MMAs on operands resident in shared
 * memory, one CTA per SM); returns int8 TOP/s.  Roofline denominator of the tcgen05 pair engine. */
double b200_measure_i8_mma_peak(void);
/* variant 0 independent accumulators (issue peak), 1 one accumulator (dependent K loop), 2 = 1 + a commit every second MMA */
double b200_measure_i8_mma_peak_variant(int variant);

/* Latency probe: average cycles for one thread to issue nMma tcgen05 MMAs (M128 N256 K32, one accumulator),
 * commit to an mbarrier and wake up on it; nMma = 0 is the commit round trip alone.  <0 on error. */
double b200_debug_i8_mma_latency(int nMma, int ctas);

#ifdef __cplusplus
}
#endif
#endif
