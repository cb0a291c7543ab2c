// ref_driver.cpp -- thin C ABI around the UNMODIFIED reference classes.
//
// TEST INFRASTRUCTURE ONLY (see oracle/rmsd_oracle.c header).  This file is
// our own code; it #includes the reference headers from /root/reference/src
// and is linked against the reference's own Frame.cpp, Matrix_3x3.cpp,
// Vec3.cpp, Box.cpp, CoordinateInfo.cpp, CompactFrameArray.cpp, AtomMask.cpp,
// Atom.cpp, ... compiled where they lie (oracle/build_ref.sh) into
// oracle/_ref/libcpptraj_ref_rmsd.so.  No reference source is copied.
//
// What is real reference code on this path: the float->double mask gather
// (CompactFrameArray::GetToMaskDblPtr), mass setup (Frame::SetupFrameFromMask),
// centring (Frame::CenterOnOrigin), Frame::RMSD / RMSD_CenteredRef /
// RMSD_NoFit incl. Matrix_3x3::Diagonalize_Sort, and the output indexers
// (Matrix<float> TRIANGLE / FULL).  What is restated here: the pair loops of
// Analysis_Rms2d::Calculate_2D (src/Analysis_Rms2d.cpp:247-287),
// MetricArray::calcFrameDistances (src/Cluster/MetricArray.cpp:777-795) and
// the per-frame body of Action_Rmsd::DoAction (src/Action_Rmsd.cpp:361-392),
// because those classes drag in the whole DataSet/ArgList/Topology runtime.
#include <vector>
#include <cmath>
#include <cstring>
#include <cstddef>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "Frame.h"
#include "Atom.h"
#include "AtomMask.h"
#include "CompactFrameArray.h"
#include "CoordinateInfo.h"
#include "Matrix.h"
#include "Matrix_3x3.h"
#include "Vec3.h"
#include "Cluster/DynamicMatrix.h"
#include <limits>

namespace {

struct Coords {
  CompactFrameArray frames;
  std::vector<Atom> atoms;
  AtomMask mask;
};

// Build a real CompactFrameArray holding the caller's float frames.  Velocities
// are present in the array when the caller's stride has room for them
// (stride >= 6*natomTotal), which exercises the "frame stride != 3N" layout
// pinned by Test_2DRMS test 7.
int fill(Coords& C, const float* crd, size_t stride, int nTotalFrames, int natomTotal,
         const int* sel, int n, const double* massSel)
{
  bool hasVel = (stride >= (size_t)6 * natomTotal);
  CoordinateInfo cinfo(Box(), hasVel, false, false);
  if (C.frames.SetupFrameArray(cinfo, natomTotal, nTotalFrames)) return 1;
  std::vector<double> tmp(3 * (size_t)natomTotal), vel(3 * (size_t)natomTotal, 0.0);
  for (int f = 0; f < nTotalFrames; f++) {
    const float* fb = crd + (size_t)f * stride;
    for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (double)fb[i];
    C.frames.SeekAndAllocate(f);
    C.frames.SetFromDblPtr(&tmp[0], CoordinateInfo::POSITION);
    if (hasVel) {
      for (size_t i = 0; i < vel.size(); i++) vel[i] = (double)fb[3 * (size_t)natomTotal + i];
      C.frames.SetFromDblPtr(&vel[0], CoordinateInfo::VELOCITY);
    }
  }
  C.atoms.assign(natomTotal, Atom(NameType("X"), 0.0, 1.0, NameType("X")));
  std::vector<int> s(sel, sel + n);
  for (int j = 0; j < n; j++)
    C.atoms[sel[j]] = Atom(NameType("X"), 0.0, massSel ? massSel[j] : 1.0, NameType("X"));
  C.mask = AtomMask(s, natomTotal);
  return 0;
}

inline void getFrame(Coords const& C, int idx, Frame& f) {
  // DataSet_Coords_CRD::GetFrame(idx, Frame&, mask)  (src/DataSet_Coords_CRD.cpp:142-147)
  C.frames.GetToMaskDblPtr(f.xAddress(), C.mask.Selected(), idx, CoordinateInfo::POSITION);
}

double g_loop_seconds = 0.0;
inline double now() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

} // namespace

extern "C" {

/// Wall-clock seconds of the last call's pair/frame loop only (excludes the
/// CompactFrameArray fill) -- what cpptraj's own TIME: line would cover.
double ref_last_loop_seconds() { return g_loop_seconds; }

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

// rms2d, triangle branch.  frameIdx nullable.
int ref_rms2d_tri(const float* crd, size_t stride, int nTotalFrames, int natomTotal,
                  const int* frameIdx, int nF, const int* sel, int n,
                  const double* mass, int fit, float* out)
{
  Coords C;
  if (fill(C, crd, stride, nTotalFrames, natomTotal, sel, n, mass)) return 1;
  bool useMass = (mass != 0);
  Matrix<float> mat;
  if (nF > 1 && mat.resize(0, nF)) return 1;
  Frame SelectedRef, SelectedTgt;
  SelectedRef.SetupFrameFromMask(C.mask, C.atoms);
  SelectedTgt.SetupFrameFromMask(C.mask, C.atoms);
  int nref, ntgt;
  double t0__ = now();
#pragma omp parallel private(nref, ntgt) firstprivate(SelectedTgt, SelectedRef)
  {
#pragma omp for schedule(dynamic)
    for (nref = 0; nref < nF; nref++) {
      getFrame(C, frameIdx ? frameIdx[nref] : nref, SelectedRef);
      if (fit) SelectedRef.CenterOnOrigin(useMass);
      for (ntgt = nref + 1; ntgt < nF; ntgt++) {
        getFrame(C, frameIdx ? frameIdx[ntgt] : ntgt, SelectedTgt);
        float R = fit ? (float)SelectedTgt.RMSD_CenteredRef(SelectedRef, useMass)
                      : (float)SelectedTgt.RMSD_NoFit(SelectedRef, useMass);
        mat.setElement(nref, ntgt, R);
      }
    }
  }
  g_loop_seconds = now() - t0__;
  if (nF > 1) std::memcpy(out, mat.Ptr(), sizeof(float) * mat.size());
  return 0;
}

// Cluster pairwise cache fill (Metric_RMS::FrameDist per pair).
int ref_cluster_tri(const float* crd, size_t stride, int nTotalFrames, int natomTotal,
                    const int* frameIdx, int nF, const int* sel, int n,
                    const double* mass, int fit, float* out)
{
  Coords C;
  if (fill(C, crd, stride, nTotalFrames, natomTotal, sel, n, mass)) return 1;
  bool useMass = (mass != 0);
  Matrix<float> mat;
  if (nF > 1 && mat.resize(0, nF)) return 1;
  Frame frm1, frm2;
  frm1.SetupFrameFromMask(C.mask, C.atoms);
  frm2 = frm1;
  int f1, f2;
  double t0__ = now();
#pragma omp parallel private(f1, f2) firstprivate(frm1, frm2)
  {
#pragma omp for schedule(dynamic)
    for (f1 = 0; f1 < nF - 1; f1++) {
      for (f2 = f1 + 1; f2 < nF; f2++) {
        getFrame(C, frameIdx ? frameIdx[f1] : f1, frm1);
        getFrame(C, frameIdx ? frameIdx[f2] : f2, frm2);
        double d = fit ? frm1.RMSD(frm2, useMass) : frm1.RMSD_NoFit(frm2, useMass);
        mat.setElement(f1, f2, (float)d);
      }
    }
  }
  g_loop_seconds = now() - t0__;
  if (nF > 1) std::memcpy(out, mat.Ptr(), sizeof(float) * mat.size());
  return 0;
}

// rms2d, full-matrix branch (reftraj and/or differing masks).
int ref_rms2d_full(const float* crdT, size_t strideT, int nT, int natomTotalT, const int* selT,
                   const float* crdR, size_t strideR, int nR, int natomTotalR, const int* selR,
                   int n, const double* massTgt, const double* massRef, int fit, float* out)
{
  Coords CT, CR;
  if (fill(CT, crdT, strideT, nT, natomTotalT, selT, n, massTgt)) return 1;
  if (fill(CR, crdR, strideR, nR, natomTotalR, selR, n, massRef)) return 1;
  bool useMass = (massTgt != 0);
  Matrix<float> mat;
  if (mat.resize(nR, nT)) return 1; // Allocate2D(totalref, totaltgt)
  Frame SelectedRef, SelectedTgt;
  SelectedRef.SetupFrameFromMask(CR.mask, CR.atoms);
  SelectedTgt.SetupFrameFromMask(CT.mask, CT.atoms);
  int nref, ntgt;
  double t0__ = now();
#pragma omp parallel private(nref, ntgt) firstprivate(SelectedTgt, SelectedRef)
  {
#pragma omp for schedule(dynamic)
    for (nref = 0; nref < nR; nref++) {
      getFrame(CR, nref, SelectedRef);
      if (fit) SelectedRef.CenterOnOrigin(useMass);
      for (ntgt = 0; ntgt < nT; ntgt++) {
        getFrame(CT, ntgt, SelectedTgt);
        float R = fit ? (float)SelectedTgt.RMSD_CenteredRef(SelectedRef, useMass)
                      : (float)SelectedTgt.RMSD_NoFit(SelectedRef, useMass);
        mat.setElement(nref, ntgt, R);
      }
    }
  }
  g_loop_seconds = now() - t0__;
  std::memcpy(out, mat.Ptr(), sizeof(float) * mat.size());
  return 0;
}

// rmsd action body, fixed reference.  refSel: 3n doubles, not centred.
int ref_rmsd_1vN(const float* crd, size_t stride, int nF, int natomTotal, const int* sel, int n,
                 const double* refSel, const double* mass, int fit,
                 double* rmsdOut, double* rotOut, double* transOut, double* refTrans)
{
  Coords C;
  if (fill(C, crd, stride, nF, natomTotal, sel, n, mass)) return 1;
  bool useMass = (mass != 0);
  Frame selectedRef;
  selectedRef.SetupFrameFromMask(C.mask, C.atoms);
  std::memcpy(selectedRef.xAddress(), refSel, sizeof(double) * 3 * (size_t)n);
  Vec3 rt(0.0);
  if (fit) rt = selectedRef.CenterOnOrigin(useMass);
  if (refTrans) { refTrans[0] = rt[0]; refTrans[1] = rt[1]; refTrans[2] = rt[2]; }
  Frame tgtFrame;
  tgtFrame.SetupFrameFromMask(C.mask, C.atoms);
  int f;
  double t0__ = now();
#pragma omp parallel private(f) firstprivate(tgtFrame)
  {
#pragma omp for schedule(static)
    for (f = 0; f < nF; f++) {
      getFrame(C, f, tgtFrame);
      double v;
      if (!fit)
        v = tgtFrame.RMSD_NoFit(selectedRef, useMass);
      else {
        Matrix_3x3 rot; Vec3 tgtTrans;
        v = tgtFrame.RMSD_CenteredRef(selectedRef, rot, tgtTrans, useMass);
        if (rotOut) std::memcpy(rotOut + 9 * (size_t)f, rot.Dptr(), 9 * sizeof(double));
        if (transOut) { transOut[3*(size_t)f] = tgtTrans[0]; transOut[3*(size_t)f+1] = tgtTrans[1]; transOut[3*(size_t)f+2] = tgtTrans[2]; }
      }
      rmsdOut[f] = v;
    }
  }
  g_loop_seconds = now() - t0__;
  return 0;
}

// Metric_RMS::CalculateCentroid (src/Cluster/Metric_RMS.cpp:86-113) with the reference's own Frame arithmetic.
int ref_build_centroid(const float* crd, size_t stride, int nTotalFrames, int natomTotal, const int* frames, int nIn,
                       const int* sel, int n, const double* mass, int fit, double* cent)
{
  Coords C;
  if (fill(C, crd, stride, nTotalFrames, natomTotal, sel, n, mass)) return 1;
  bool useMass = (mass != 0);
  Frame frm1, cframe;
  frm1.SetupFrameFromMask(C.mask, C.atoms);
  Matrix_3x3 Rot; Vec3 Trans;
  for (int j = 0; j < nIn; j++) {
    getFrame(C, frames[j], frm1);
    if (cframe.empty()) {
      cframe = frm1;
      if (fit) cframe.CenterOnOrigin(useMass);
    } else {
      if (fit) {
        frm1.RMSD_CenteredRef(cframe, Rot, Trans, useMass);
        frm1.Rotate(Rot);
      }
      cframe += frm1;
    }
  }
  if (nIn > 0) {
    cframe.Divide((double)nIn);
    std::memcpy(cent, cframe.xAddress(), sizeof(double) * 3 * (size_t)n);
  }
  return 0;
}

// Algorithm_HierAgglo::DoClustering / MergeClosest (src/Cluster/Algorithm_HierAgglo.cpp:97-245) driven through the
// reference's OWN Cluster::DynamicMatrix (src/Cluster/DynamicMatrix.{h,cpp}: SetCdist / Ignore / FindMin with their
// closest-index bookkeeping) and Matrix<float>.  Restated here: the cluster list (frame vectors, ascending Num) and the
// frame-pair linkage loops (:248-350) over the cache triangle, because List/Node/MetricArray drag in the DataSet runtime.
int ref_hieragglo(const float* tri, int n, int linkage, int targetClusters, double epsilon,
                  int* mergeInto, int* mergeFrom, float* findMin, int* nCalls, int* nMerges)
{
  *nCalls = 0; *nMerges = 0;
  if (n < 2) return 0;
  Matrix<float> cache;
  cache.resize(0L, (size_t)n);
  std::memcpy(cache.Ptr(), tri, sizeof(float) * cache.size());
  std::vector< std::vector<int> > fl((size_t)n);
  std::vector<bool> alive((size_t)n, true);
  for (int i = 0; i < n; i++) fl[i].assign(1, i);
  struct Link {
    static double dist(Matrix<float> const& cache, int linkage, std::vector<int> const& c1, std::vector<int> const& c2) {
      double acc = (linkage == 0) ? std::numeric_limits<double>::max() : (linkage == 2 ? -1.0 : 0.0);
      for (std::vector<int>::const_iterator a = c1.begin(); a != c1.end(); ++a)
        for (std::vector<int>::const_iterator b = c2.begin(); b != c2.end(); ++b) {
          double Dist = cache.element(*a, *b);
          if (linkage == 0) { if (Dist < acc) acc = Dist; }
          else if (linkage == 2) { if (Dist > acc) acc = Dist; }
          else acc += Dist;
        }
      if (linkage == 1) return acc / (double)(c1.size() * c2.size());
      return acc;
    }
  };
  Cpptraj::Cluster::DynamicMatrix CD;
  CD.SetupMatrix((size_t)n);
  for (int c1 = 0; c1 < n; c1++)
    for (int c2 = c1 + 1; c2 < n; c2++)
      CD.SetCdist(c1, c2, Link::dist(cache, linkage, fl[c1], fl[c2]));
  int nClusters = n;
  for (;;) {
    int C1, C2;
    double min = CD.FindMin(C1, C2);
    findMin[(*nCalls)++] = (float)min;
    if (min > epsilon) break;
    mergeInto[*nMerges] = C1; mergeFrom[*nMerges] = C2; (*nMerges)++;
    fl[C1].insert(fl[C1].end(), fl[C2].begin(), fl[C2].end());
    fl[C2].clear(); alive[C2] = false;
    nClusters--;
    CD.Ignore(C2);
    for (int k = 0; k < n; k++)
      if (alive[k] && k != C1) CD.SetCdist(C1, k, Link::dist(cache, linkage, fl[C1], fl[k]));
    if (nClusters <= targetClusters) break;
    if (nClusters == 1) break;
  }
  return 0;
}

// Analysis_RmsAvgCorr::Analyze (src/Analysis_RmsAvgCorr.cpp:119-316) with the reference's own Frame arithmetic
// (operator+=, operator-=, Divide, CenterOnOrigin, RMSD_CenteredRef); the window loop is restated.
int ref_rmsavgcorr(const float* crd, size_t stride, int nTotalFrames, int natomTotal, const int* sel, int n, const double* mass,
                   const double* refSel, const int* windows, int nW, double* avgOut, double* sdOut)
{
  Coords C;
  if (fill(C, crd, stride, nTotalFrames, natomTotal, sel, n, mass)) return 1;
  bool useMass = (mass != 0);
  int maxFrame = nTotalFrames;
  Frame tgtTemplate;
  tgtTemplate.SetupFrameFromMask(C.mask, C.atoms);
  Frame refFixed = tgtTemplate;
  if (refSel != 0) std::memcpy(refFixed.xAddress(), refSel, sizeof(double) * 3 * (size_t)n);
  int w;
#pragma omp parallel for schedule(dynamic)
  for (w = 0; w < nW; w++) {
    Frame tgtFrame = tgtTemplate, refFrame = refFixed, sumFrame(n);
    int window = windows[w];
    double avg = 0.0, stdev = 0.0, d_Nwindow = (double)window;
    bool first = (refSel == 0);
    int subtractWindow = 0, frameThreshold = window - 2;
    sumFrame.ZeroCoords();
    for (int frame = 0; frame < maxFrame; frame++) {
      getFrame(C, frame, tgtFrame);
      if (window != 1) {
        sumFrame += tgtFrame;
        if (!(frame > frameThreshold)) continue;
        tgtFrame.Divide(sumFrame, d_Nwindow);
      }
      if (first) {
        refFrame.SetCoordinates(tgtFrame);
        refFrame.CenterOnOrigin(useMass);
        first = false;
      }
      double rmsd = tgtFrame.RMSD_CenteredRef(refFrame, useMass);
      avg += rmsd; stdev += rmsd * rmsd;
      if (window != 1) {
        getFrame(C, subtractWindow, tgtFrame);
        sumFrame -= tgtFrame;
        ++subtractWindow;
      }
    }
    d_Nwindow = 1.0 / ((double)maxFrame - (double)window + 1.0);
    avg *= d_Nwindow; stdev *= d_Nwindow; stdev -= (avg * avg);
    stdev = (stdev > 0.0) ? sqrt(stdev) : 0.0;
    avgOut[w] = avg; sdOut[w] = stdev;
  }
  return 0;
}

} // extern "C"
