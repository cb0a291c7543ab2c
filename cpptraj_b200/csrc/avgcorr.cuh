// avgcorr.cuh -- RMSD of running-averaged coordinates for every window size (Analysis_RmsAvgCorr::Analyze,
// src/Analysis_RmsAvgCorr.cpp:119-316): for window size W the frames t = 0 .. F-W are replaced by the average of frames
// t .. t+W-1 (selected atoms), each average is fitted to a reference -- the first averaged frame of that window size
// ("first") or a fixed, pre-centred reference -- and the mean and standard deviation of the RMSDs are reported.  The
// reference keeps a running sum per window size (add frame t, subtract frame t-W): O(F^2 N) coordinate operations over
// all window sizes.  Here the selected coordinates are prefix-summed once over the frames, in FP64,
//     P[t] = x_0 + ... + x_(t-1),      average(W, t) = (P[t+W] - P[t]) / W,
// and one warp per (W, t) streams the two prefix rows, accumulating the 17 FP64 sums the fit needs (sum m a, sum m r,
// sum m |a|^2, sum m a r^T); the centre of the averaged frame is removed algebraically (S = sum m a r^T - c_a (sum m r)^T,
// G_a = sum m |a|^2 - M |c_a|^2), the per-(W,t) root is the same key-matrix quartic as everywhere else.  HBM-bound:
// 48 bytes per atom and (W, t) (two FP64 prefix rows; the window's reference row stays in L2).
#pragma once
#include "rmsd_kernels.cuh"

namespace b200 {

/// Thread c owns coordinate column c (atom c/3, component c%3): gather float -> double and prefix-sum over the frames.
__global__ void __launch_bounds__(128) avgcorr_prefix_kernel(const float* __restrict__ crd, size_t stride, long srcBase,
                                                             const int* __restrict__ atomIdx, int nAtoms, int nFrames,
                                                             const double* __restrict__ shift, double* __restrict__ P) {
  // output column c = plane * nAtoms + atom (x plane, y plane, z plane): a warp of avgcorr_kernel then reads 32
  // consecutive doubles per coordinate (interleaved xyz rows cost three 24-byte-strided loads per atom: ncu showed the
  // L1 data pipe at 74 % with the FP64 pipe at 43 %)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int ld = 3 * nAtoms;
  if (c >= ld) return;
  const int plane = c / nAtoms, atom = c - plane * nAtoms;
  const size_t col = (size_t)3 * (size_t)(atomIdx ? atomIdx[atom] : atom) + (size_t)plane;
  // A common origin near the molecule (the centre of frame 0's selection) is removed first: the fit removes every
  // frame's own centre anyway, and the sums formed from uncentred coordinates then cancel far less.
  const double sh = shift[plane];
  double acc = 0.0;
  P[c] = 0.0;
  const float* src = crd + col - (size_t)srcBase * stride;
  int t = 0;
  for (; t + 4 <= nFrames; t += 4) {
    const float x0 = src[(size_t)t * stride], x1 = src[(size_t)(t + 1) * stride], x2 = src[(size_t)(t + 2) * stride],
                x3 = src[(size_t)(t + 3) * stride];
    acc += (double)x0 - sh; P[(size_t)(t + 1) * ld + c] = acc;
    acc += (double)x1 - sh; P[(size_t)(t + 2) * ld + c] = acc;
    acc += (double)x2 - sh; P[(size_t)(t + 3) * ld + c] = acc;
    acc += (double)x3 - sh; P[(size_t)(t + 4) * ld + c] = acc;
  }
  for (; t < nFrames; ++t) { acc += (double)src[(size_t)t * stride] - sh; P[(size_t)(t + 1) * ld + c] = acc; }
}

/// shift[0..2] = unweighted centre of frame 0's selected atoms.  One warp.
__global__ void avgcorr_shift_kernel(const float* __restrict__ crd, size_t stride, long srcBase, const int* __restrict__ atomIdx,
                                     int nAtoms, double* shift) {
  const float* f0 = crd - (size_t)srcBase * stride;
  double s[3] = {0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < nAtoms; k += 32) {
    const size_t at = (size_t)3 * (size_t)(atomIdx ? atomIdx[k] : k);
    for (int d = 0; d < 3; ++d) s[d] += (double)f0[at + d];
  }
  for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]);
  if (threadIdx.x == 0) for (int d = 0; d < 3; ++d) shift[d] = s[d] / (double)nAtoms;
}

struct AvgCorrArgs {
  const double* P;         // [nFrames+1][ld] prefix sums, row 0 = 0; a row is three planes x | y | z of nAtoms doubles
  int ld, nAtoms, nFrames;
  const double* mass;      // nullable: per selected atom
  const double* refFixed;  // nullable: fixed reference as the caller centred it, plane-major like a row of P; null = "first"
  const int* win;          // window sizes of this batch
  const long long* itemOff;  // [nW+1]: first RMSD slot of each window
  int nW;
  double* refInfo;         // [nW][8]: centre xyz of the window's reference, G_R = sum m |r|^2, sum m r (xyz)
  double* rms;             // RMSD per (window, t)
  double totalMass;
};

/// Per window: centre, G and mass-weighted sum of its reference.  One CTA per window (fixed reference: one CTA).
__global__ void __launch_bounds__(256) avgcorr_ref_kernel(AvgCorrArgs a) {
  __shared__ double sh[8][4];
  const int w = blockIdx.x;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const double dW = (double)a.win[w];
  const double* R = a.refFixed ? a.refFixed : a.P + (size_t)a.win[w] * a.ld;
  const double div = a.refFixed ? 1.0 : dW;
  double c[3] = {0.0, 0.0, 0.0};
  if (!a.refFixed) {   // centre of the first averaged frame (Frame::CenterOnOrigin with the target masses)
    double s[3] = {0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < a.nAtoms; k += blockDim.x) {
      const double m = a.mass ? a.mass[k] : 1.0;
      for (int d = 0; d < 3; ++d) s[d] += m * (R[d * a.nAtoms + k] / div);
    }
    for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]);
    if (lane == 0) for (int d = 0; d < 3; ++d) sh[wp][d] = s[d];
    __syncthreads();
    for (int d = 0; d < 3; ++d) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += sh[i][d];
      c[d] = t / a.totalMass;
    }
    __syncthreads();
  }
  double g = 0.0, sr[3] = {0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < a.nAtoms; k += blockDim.x) {
    const double m = a.mass ? a.mass[k] : 1.0;
    for (int d = 0; d < 3; ++d) {
      const double r = R[d * a.nAtoms + k] / div - c[d];
      g += m * r * r; sr[d] += m * r;
    }
  }
  g = warp_sum(g);
  for (int d = 0; d < 3; ++d) sr[d] = warp_sum(sr[d]);
  if (lane == 0) { sh[wp][0] = g; sh[wp][1] = sr[0]; sh[wp][2] = sr[1]; sh[wp][3] = sr[2]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) t[j] += sh[i][j];
    double* o = a.refInfo + (size_t)w * 8;
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = t[0]; o[4] = t[1]; o[5] = t[2]; o[6] = t[3];
  }
}

constexpr int AVGCORR_WARPS = 8;
/// grid (ceil(maxItems / AVGCORR_WARPS), ceil(nW / NWIN)): warp -> averaged frame t of NWIN window sizes.  The row P[t]
/// is read once for the NWIN windows, the rows P[t+W] of neighbouring warps and windows overlap (consecutive window
/// sizes: t+W runs over AVGCORR_WARPS+NWIN-1 rows for AVGCORR_WARPS*NWIN averages) and are served by L1; nothing is
/// divided inside the loop: the sums are formed from the raw prefix differences d = P[t+W]-P[t] and scaled once.
template <int NWIN>
__global__ void __launch_bounds__(AVGCORR_WARPS * 32, (NWIN <= 2 ? 2 : 1)) avgcorr_kernel(AvgCorrArgs a) {
  const int t = blockIdx.x * AVGCORR_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int w0 = blockIdx.y * NWIN;
  int W[NWIN]; bool ok[NWIN]; const double* Pu[NWIN]; const double* R[NWIN];
  bool any = false;
#pragma unroll
  for (int j = 0; j < NWIN; ++j) {
    const int wi = (w0 + j < a.nW) ? w0 + j : a.nW - 1;
    W[j] = a.win[wi];
    ok[j] = (w0 + j < a.nW) && t <= a.nFrames - W[j];
    any = any || ok[j];
    Pu[j] = a.P + (size_t)(ok[j] ? t + W[j] : 0) * a.ld;
    R[j] = a.refFixed ? a.refFixed : a.P + (size_t)W[j] * a.ld;
  }
  if (!any) return;
  const double* Pt = a.P + (size_t)(t <= a.nFrames ? t : 0) * a.ld;
  double sa[NWIN][3], saa[NWIN], S[NWIN][9];
#pragma unroll
  for (int j = 0; j < NWIN; ++j) {
    sa[j][0] = sa[j][1] = sa[j][2] = 0.0; saa[j] = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) S[j][i] = 0.0;
  }
  for (int k = lane; k < a.nAtoms; k += 32) {
    const double m = a.mass ? a.mass[k] : 1.0;
    const int ky = a.nAtoms + k, kz = 2 * a.nAtoms + k;
    const double p0 = Pt[k], p1 = Pt[ky], p2 = Pt[kz];
#pragma unroll
    for (int j = 0; j < NWIN; ++j) {
      const double dx = Pu[j][k] - p0, dy = Pu[j][ky] - p1, dz = Pu[j][kz] - p2;
      const double mx = m * dx, my = m * dy, mz = m * dz;
      const double rx = R[j][k], ry = R[j][ky], rz = R[j][kz];
      sa[j][0] += mx; sa[j][1] += my; sa[j][2] += mz;
      saa[j] += mx * dx + my * dy + mz * dz;
      S[j][0] += mx * rx; S[j][1] += mx * ry; S[j][2] += mx * rz;
      S[j][3] += my * rx; S[j][4] += my * ry; S[j][5] += my * rz;
      S[j][6] += mz * rx; S[j][7] += mz * ry; S[j][8] += mz * rz;
    }
  }
#pragma unroll
  for (int j = 0; j < NWIN; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) sa[j][i] = warp_sum(sa[j][i]);
    saa[j] = warp_sum(saa[j]);
#pragma unroll
    for (int i = 0; i < 9; ++i) S[j][i] = warp_sum(S[j][i]);
  }
#pragma unroll
  for (int j = 0; j < NWIN; ++j) {
    if (lane != j || !ok[j]) continue;
    const double* info = a.refInfo + (size_t)(w0 + j) * 8;
    double r;
    if (a.totalMass < 1e-14) r = -1.0;   // src/Frame.cpp:1160-1163
    else {
      const double invM = 1.0 / a.totalMass, invW = 1.0 / (double)W[j];
      const double invWR = a.refFixed ? invW : invW * invW;   // averaged frame / W, "first" reference row / W as well
      // a = d / W, r = R / div - cR:  sum m a r^T = S_raw / (W div) - (sum m a) cR^T ; then the centre of a is removed
      const double sma[3] = {sa[j][0] * invW, sa[j][1] * invW, sa[j][2] * invW};
      const double ca[3] = {sma[0] * invM, sma[1] * invM, sma[2] * invM};
      const double ga = saa[j] * invW * invW - (sma[0] * ca[0] + sma[1] * ca[1] + sma[2] * ca[2]);
      double C[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[3 * i + c] = S[j][3 * i + c] * invWR - sma[i] * info[c] - ca[i] * info[4 + c];
      r = rmsd_fit_from_cov(C, 0.5 * (ga + info[3]), invM);
    }
    a.rms[a.itemOff[w0 + j] + t] = r;
  }
}

/// Mean and standard deviation per window exactly as the reference forms them (:196-205, :289-297), summed in a fixed
/// order (deterministic).  One CTA per window.
__global__ void __launch_bounds__(256) avgcorr_reduce_kernel(const double* rms, const long long* itemOff, int nW, double* avgOut,
                                                             double* sdOut) {
  __shared__ double sh[2][256];
  const int w = blockIdx.x;
  const long long n = itemOff[w + 1] - itemOff[w];
  const double* r = rms + itemOff[w];
  double s = 0.0, s2 = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { const double v = r[i]; s += v; s2 += v * v; }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double d = 1.0 / (double)n;
    const double avg = sh[0][0] * d;
    double sd = sh[1][0] * d - avg * avg;
    sd = sd > 0.0 ? sqrt(sd) : 0.0;
    avgOut[w] = avg; sdOut[w] = sd;
  }
}

}  // namespace b200
