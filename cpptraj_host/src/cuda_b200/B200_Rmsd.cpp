#include "B200_Rmsd.h"
#include "b200_rmsd.h"
#include <cstdlib>
#include <algorithm>
#include "../AtomMask.h"
#include "../CpptrajStdio.h"
#include "../DataSet_Coords_CRD.h"
#include "../DataSet_MatrixFlt.h"
#include "../DataSet_double.h"
#include "../Frame.h"
#include "../Cluster/Cframes.h"

static int b200_err(const char* what) {
  mprinterr("Error: B200 RMSD (%s): %s\n", what, b200_last_error());
  return 1;
}

int Cpptraj::B200::Init() {
  static bool initialised = false;
  if (initialised) return 0;
  int want = 0, used = 0;
  const char* env = getenv("CPPTRAJ_B200_NGPU");
  if (env != 0) want = atoi(env);
  if (b200_init(want, &used)) return b200_err("device probe");
  mprintf("\tB200 RMSD path: %i device(s).\n", used);
  initialised = true;
  return 0;
}

std::vector<double> Cpptraj::B200::MassesOf(Frame const& frm) {
  std::vector<double> m( (size_t)frm.Natom() );
  for (int i = 0; i < frm.Natom(); i++) m[i] = frm.Mass(i);
  return m;
}

static inline const double* ptr_or_null(std::vector<double> const& v) { return v.empty() ? 0 : &v[0]; }

int Cpptraj::B200::Rms2d(DataSet_Coords_CRD const& tgt, AtomMask const& tgtMask, std::vector<double> const& tgtMass,
                         DataSet_Coords_CRD const& ref, AtomMask const& refMask, std::vector<double> const& refMass,
                         bool fullMatrix, bool fit, DataSet_MatrixFlt& out)
{
  if (Init()) return 1;
  if (tgtMask.Nselected() != refMask.Nselected()) {
    mprinterr("Error: B200 RMSD: # target atoms (%i) != # reference atoms (%i)\n", tgtMask.Nselected(), refMask.Nselected());
    return 1;
  }
  if (tgt.Size() < 1 || tgtMask.Nselected() < 1) return 0;
  float* mat = static_cast<float*>( out.MatrixPtr() );
  int err;
  if (!fullMatrix)
    err = b200_rms2d_tri(tgt.RawFrames(), tgt.FrameStride(), (int)tgt.Size(), 0, (int)tgt.Size(),
                         &tgtMask.Selected()[0], tgtMask.Nselected(), ptr_or_null(tgtMass), fit ? 1 : 0, mat);
  else
    err = b200_rms2d_full(tgt.RawFrames(), tgt.FrameStride(), (int)tgt.Size(), &tgtMask.Selected()[0],
                          ref.RawFrames(), ref.FrameStride(), (int)ref.Size(), &refMask.Selected()[0],
                          tgtMask.Nselected(), ptr_or_null(tgtMass), ptr_or_null(refMass), fit ? 1 : 0, mat);
  return err ? b200_err("rms2d") : 0;
}

int Cpptraj::B200::CacheFill(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
                             bool fit, Cluster::Cframes const& framesToCache, float* triangle)
{
  if (Init()) return 1;
  if (framesToCache.size() < 2) return 0;
  if (b200_rms2d_tri(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &(*framesToCache.begin()),
                     (int)framesToCache.size(), &mask.Selected()[0], mask.Nselected(), ptr_or_null(mass), fit ? 1 : 0,
                     triangle))
    return b200_err("pairwise cache");
  return 0;
}

int Cpptraj::B200::ClosestCentroids(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
                                    bool fit, Cluster::Cframes const& frames,
                                    std::vector<Frame const*> const& centroidFrames,
                                    std::vector<int>& closest, std::vector<double>& closestDist)
{
  if (Init()) return 1;
  closest.assign(frames.size(), -1);
  closestDist.assign(frames.size(), 0.0);
  if (frames.size() < 1 || centroidFrames.empty()) return 0;
  const size_t n3 = (size_t)3 * (size_t)mask.Nselected();
  std::vector<double> cen( centroidFrames.size() * n3 );
  for (size_t k = 0; k != centroidFrames.size(); k++) {
    if (centroidFrames[k]->Natom() != mask.Nselected()) {
      mprinterr("Error: B200 RMSD: centroid %zu has %i atoms, mask selects %i\n", k, centroidFrames[k]->Natom(), mask.Nselected());
      return 1;
    }
    std::copy(centroidFrames[k]->xAddress(), centroidFrames[k]->xAddress() + n3, cen.begin() + k * n3);
  }
  if (b200_rmsd_frames_to_centroids(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &(*frames.begin()), (int)frames.size(),
                                    &mask.Selected()[0], mask.Nselected(), ptr_or_null(mass), fit ? 1 : 0,
                                    &cen[0], (int)centroidFrames.size(), 0, &closest[0], &closestDist[0]))
    return b200_err("frame-centroid distances");
  return 0;
}

// -----------------------------------------------------------------------------
Cpptraj::B200::Rmsd1vN::~Rmsd1vN() { if (handle_ != 0) b200_rmsd_1vN_end(handle_); }

int Cpptraj::B200::Rmsd1vN::Begin(Frame const& selectedRef, Frame const& massFrame, bool fit, bool useMass) {
  if (Init()) return 1;
  if (handle_ != 0) { b200_rmsd_1vN_end(handle_); handle_ = 0; }
  nAtoms_ = selectedRef.Natom();
  nBuffered_ = 0; best_ = -1;
  if (nAtoms_ < 1) { mprinterr("Error: B200 RMSD: empty reference selection.\n"); return 1; }
  buffer_.resize( (size_t)BATCH * 3 * (size_t)nAtoms_ );
  std::vector<int> identity( nAtoms_ );           // frames arrive already gathered (tgtFrame_.SetCoordinates)
  for (int i = 0; i < nAtoms_; i++) identity[i] = i;
  std::vector<double> mass;
  if (useMass) mass = MassesOf(massFrame);
  if (b200_rmsd_1vN_begin(selectedRef.xAddress(), &identity[0], nAtoms_, ptr_or_null(mass), fit ? 1 : 0, 0, &handle_))
    return b200_err("rmsd setup");
  return 0;
}

int Cpptraj::B200::Rmsd1vN::pushBuffer() {
  if (nBuffered_ == 0) return 0;
  if (b200_rmsd_1vN_push_f64(handle_, &buffer_[0], (size_t)3 * (size_t)nAtoms_, (int)nBuffered_)) return b200_err("rmsd push");
  nBuffered_ = 0;
  return 0;
}

int Cpptraj::B200::Rmsd1vN::Push(Frame const& selectedTgt) {
  if (selectedTgt.Natom() != nAtoms_) {
    mprinterr("Error: B200 RMSD: frame has %i selected atoms, reference %i\n", selectedTgt.Natom(), nAtoms_);
    return 1;
  }
  const double* x = selectedTgt.xAddress();
  std::copy(x, x + (size_t)3 * (size_t)nAtoms_, buffer_.begin() + (size_t)nBuffered_ * 3 * (size_t)nAtoms_);
  if (++nBuffered_ == BATCH) return pushBuffer();
  return 0;
}

int Cpptraj::B200::Rmsd1vN::Flush(DataSet_double& rmsd) {
  if (handle_ == 0) return 0;
  if (pushBuffer()) return 1;
  const long n = b200_rmsd_1vN_pending(handle_);
  if (n < 1) return 0;
  std::vector<double> r( (size_t)n );
  if (b200_rmsd_1vN_flush(handle_, &r[0], 0, 0, &best_)) return b200_err("rmsd flush");
  for (long i = 0; i != n; i++) rmsd.AddElement( r[i] );   // append-only, frame order (DataSet_double.cpp:14-20)
  return 0;
}
