#include "B200_Rmsd.h"
#include "b200_rmsd.h"
#include <cstdlib>
#include <algorithm>
#include <thread>
#include "../Timer.h"
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

// Device set-up (CUDA context, streams, host copy pools: ~1 s in a fresh process) may be started early, on a thread of
// its own, by the Setup()/Init() of the commands that will use the path: it then overlaps trajectory reading.
static bool b200_initialised = false, b200_async_started = false;
static int b200_init_rc = 0, b200_init_used = 0;
static double b200_init_seconds = 0.0;
static std::thread b200_init_thread;

static void b200_do_init() {
  int want = 0;
  const char* env = getenv("CPPTRAJ_B200_NGPU");
  if (env != 0) want = atoi(env);
  Timer t_init;
  t_init.Start();
  b200_init_rc = b200_init(want, &b200_init_used);
  // in the background (InitAsync): also pin the staging slots and load the kernels; a synchronous Init() does not wait for that
  if (b200_init_rc == 0 && b200_async_started) b200_warmup();
  t_init.Stop();
  b200_init_seconds = t_init.Total();
}

void Cpptraj::B200::InitAsync() {
  if (b200_initialised || b200_async_started) return;
  b200_async_started = true;
  b200_init_thread = std::thread( b200_do_init );
}

int Cpptraj::B200::Init() {
  if (b200_initialised) return 0;
  if (b200_async_started) {
    if (b200_init_thread.joinable()) b200_init_thread.join();
  } else
    b200_do_init();
  if (b200_init_rc) { b200_async_started = false; return b200_err("device probe"); }
  mprintf("\tB200 RMSD path: %i device(s) (set-up %.4f s%s).\n", b200_init_used, b200_init_seconds,
          b200_async_started ? ", started in the background at command set-up" : "");
  b200_initialised = true;
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
  Timer t_call;
  t_call.Start();
  if (!fullMatrix)
    err = b200_rms2d_tri(tgt.RawFrames(), tgt.FrameStride(), (int)tgt.Size(), 0, (int)tgt.Size(),
                         &tgtMask.Selected()[0], tgtMask.Nselected(), ptr_or_null(tgtMass), fit ? 1 : 0, mat);
  else
    err = b200_rms2d_full(tgt.RawFrames(), tgt.FrameStride(), (int)tgt.Size(), &tgtMask.Selected()[0],
                          ref.RawFrames(), ref.FrameStride(), (int)ref.Size(), &refMask.Selected()[0],
                          tgtMask.Nselected(), ptr_or_null(tgtMass), ptr_or_null(refMass), fit ? 1 : 0, mat);
  t_call.Stop();
  if (!err) mprintf("\tB200 RMSD path: %zu x %zu frames, %i atoms: %.4f s in the library.\n", tgt.Size(), ref.Size(),
                    tgtMask.Nselected(), t_call.Total());
  return err ? b200_err("rms2d") : 0;
}

/// The selected atoms of every frame of \a set, frame after frame, as float.
static void b200_pack_selected(DataSet_Coords& set, AtomMask const& mask, std::vector<float>& buf) {
  Frame frm;
  frm.SetupFrameFromMask( mask, set.Top().Atoms() );
  const size_t n3 = (size_t)3 * (size_t)mask.Nselected();
  buf.resize( set.Size() * n3 );
  for (size_t idx = 0; idx != set.Size(); idx++) {
    set.GetFrame( (int)idx, frm, mask );
    const double* x = frm.xAddress();
    float* dst = &buf[idx * n3];
    for (size_t i = 0; i != n3; i++) dst[i] = (float)x[i];
  }
}

int Cpptraj::B200::Rms2dPacked(DataSet_Coords& tgt, AtomMask const& tgtMask, std::vector<double> const& tgtMass,
                               DataSet_Coords& ref, AtomMask const& refMask, std::vector<double> const& refMass,
                               bool fullMatrix, bool fit, DataSet_MatrixFlt& out)
{
  if (Init()) return 1;
  const int nsel = tgtMask.Nselected();
  if (nsel != refMask.Nselected()) {
    mprinterr("Error: B200 RMSD: # target atoms (%i) != # reference atoms (%i)\n", nsel, refMask.Nselected());
    return 1;
  }
  if (tgt.Size() < 1 || nsel < 1) return 0;
  Timer t_read;
  t_read.Start();
  std::vector<float> tbuf, rbuf;
  b200_pack_selected( tgt, tgtMask, tbuf );
  if (fullMatrix) b200_pack_selected( ref, refMask, rbuf );
  t_read.Stop();
  std::vector<int> ident( (size_t)nsel );
  for (int i = 0; i < nsel; i++) ident[i] = i;
  const size_t n3 = (size_t)3 * (size_t)nsel;
  float* mat = static_cast<float*>( out.MatrixPtr() );
  Timer t_call;
  t_call.Start();
  int err;
  if (!fullMatrix)
    err = b200_rms2d_tri(&tbuf[0], n3, (int)tgt.Size(), 0, (int)tgt.Size(), &ident[0], nsel, ptr_or_null(tgtMass), fit ? 1 : 0, mat);
  else
    err = b200_rms2d_full(&tbuf[0], n3, (int)tgt.Size(), &ident[0], &rbuf[0], n3, (int)ref.Size(), &ident[0],
                          nsel, ptr_or_null(tgtMass), ptr_or_null(refMass), fit ? 1 : 0, mat);
  t_call.Stop();
  if (!err) mprintf("\tB200 RMSD path: %zu x %zu frames, %i atoms: selected atoms read once in %.4f s, %.4f s in the library.\n",
                    tgt.Size(), ref.Size(), nsel, t_read.Total(), t_call.Total());
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

int Cpptraj::B200::CacheFillPacked(DataSet_Coords& crd, AtomMask const& mask, std::vector<double> const& mass,
                                   bool fit, Cluster::Cframes const& framesToCache, float* triangle)
{
  if (Init()) return 1;
  if (framesToCache.size() < 2) return 0;
  const int nsel = mask.Nselected();
  const size_t n3 = (size_t)3 * (size_t)nsel;
  Frame frm;
  frm.SetupFrameFromMask( mask, crd.Top().Atoms() );
  std::vector<float> buf( framesToCache.size() * n3 );
  size_t row = 0;
  for (Cluster::Cframes::const_iterator f = framesToCache.begin(); f != framesToCache.end(); ++f, ++row) {
    crd.GetFrame( *f, frm, mask );
    const double* x = frm.xAddress();
    float* dst = &buf[row * n3];
    for (size_t i = 0; i != n3; i++) dst[i] = (float)x[i];
  }
  std::vector<int> ident( (size_t)nsel );
  for (int i = 0; i < nsel; i++) ident[i] = i;
  if (b200_rms2d_tri(&buf[0], n3, (int)framesToCache.size(), 0, (int)framesToCache.size(), &ident[0], nsel, ptr_or_null(mass),
                     fit ? 1 : 0, triangle))
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

int Cpptraj::B200::CentroidDists(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
                                 bool fit, Cluster::Cframes const& frames, Frame const& centroidFrame,
                                 std::vector<double>& dist)
{
  if (Init()) return 1;
  dist.assign(frames.size(), 0.0);
  if (frames.size() < 1) return 0;
  if (centroidFrame.Natom() != mask.Nselected()) {
    mprinterr("Error: B200 RMSD: centroid has %i atoms, mask selects %i\n", centroidFrame.Natom(), mask.Nselected());
    return 1;
  }
  if (b200_rmsd_frames_to_centroids(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &(*frames.begin()), (int)frames.size(),
                                    &mask.Selected()[0], mask.Nselected(), ptr_or_null(mass), fit ? 1 : 0,
                                    centroidFrame.xAddress(), 1, &dist[0], 0, 0))
    return b200_err("frame-centroid distances");
  return 0;
}

int Cpptraj::B200::BuildCentroids(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
                                  bool fit, std::vector<Cluster::Cframes const*> const& clusterFrames,
                                  std::vector<Frame*> const& centroids)
{
  if (Init()) return 1;
  if (clusterFrames.size() != centroids.size()) {
    mprinterr("Internal Error: B200 RMSD: %zu frame lists, %zu centroids\n", clusterFrames.size(), centroids.size());
    return 1;
  }
  if (clusterFrames.empty()) return 0;
  const size_t n3 = (size_t)3 * (size_t)mask.Nselected();
  std::vector<int> offsets(1, 0), frames;
  for (size_t k = 0; k != clusterFrames.size(); k++) {
    frames.insert(frames.end(), clusterFrames[k]->begin(), clusterFrames[k]->end());
    offsets.push_back( (int)frames.size() );
  }
  if (frames.empty()) return 0;
  std::vector<double> out( clusterFrames.size() * n3 );
  if (b200_rmsd_build_centroids(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &frames[0], &offsets[0],
                                (int)clusterFrames.size(), &mask.Selected()[0], mask.Nselected(), ptr_or_null(mass),
                                fit ? 1 : 0, &out[0]))
    return b200_err("centroids");
  for (size_t k = 0; k != centroids.size(); k++) {
    if (clusterFrames[k]->empty()) continue;
    if (centroids[k]->Natom() != mask.Nselected()) {
      mprinterr("Internal Error: B200 RMSD: centroid frame %zu has %i atoms, mask selects %i\n", k, centroids[k]->Natom(), mask.Nselected());
      return 1;
    }
    std::copy(out.begin() + k * n3, out.begin() + (k + 1) * n3, centroids[k]->xAddress());
  }
  return 0;
}

// -----------------------------------------------------------------------------
Cpptraj::B200::ResidentCoords::~ResidentCoords() { if (crd_ != 0) b200_coords_resident_end(crd_); }

int Cpptraj::B200::ResidentCoords::Begin(DataSet_Coords_CRD const& crd, AtomMask const& mask) {
  if (Init()) return 1;
  if (crd.Size() < 1 || mask.None()) return 0;
  if (b200_coords_resident_begin(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &mask.Selected()[0], mask.Nselected())) {
    mprintf("Warning: B200 RMSD: coordinates not kept on the device (%s).\n", b200_last_error());
    return 1;
  }
  crd_ = crd.RawFrames();
  return 0;
}

// -----------------------------------------------------------------------------
Cpptraj::B200::Rmsd1vN::~Rmsd1vN() { End(); }

void Cpptraj::B200::Rmsd1vN::End() { if (handle_ != 0) { b200_rmsd_1vN_end(handle_); handle_ = 0; } }

int Cpptraj::B200::Rmsd1vN::Begin(Frame const& selectedRef, Frame const& massFrame, const int* atomIdx,
                                  bool fit, bool useMass, bool wantRot)
{
  if (Init()) return 1;
  End();
  nAtoms_ = selectedRef.Natom();
  if (nAtoms_ < 1) { mprinterr("Error: B200 RMSD: empty reference selection.\n"); return 1; }
  std::vector<int> identity;
  if (atomIdx == 0) {                              // frames arrive already gathered (tgtFrame_.SetCoordinates)
    identity.resize( nAtoms_ );
    for (int i = 0; i < nAtoms_; i++) identity[i] = i;
    atomIdx = &identity[0];
  }
  std::vector<double> mass;
  if (useMass) mass = MassesOf(massFrame);
  if (b200_rmsd_1vN_begin(selectedRef.xAddress(), atomIdx, nAtoms_, ptr_or_null(mass), fit ? 1 : 0, wantRot ? 1 : 0, &handle_))
    return b200_err("rmsd setup");
  return 0;
}

int Cpptraj::B200::Rmsd1vN::SetRef(Frame const& selectedRef) {
  if (handle_ == 0 || selectedRef.Natom() != nAtoms_) {
    mprinterr("Error: B200 RMSD: reference has %i selected atoms, handle %i\n", selectedRef.Natom(), nAtoms_);
    return 1;
  }
  if (b200_rmsd_1vN_set_ref(handle_, selectedRef.xAddress())) return b200_err("rmsd reference");
  return 0;
}

int Cpptraj::B200::Rmsd1vN::One(Frame const& selectedTgt, double& rmsd, double* rot, double* trans) {
  if (handle_ == 0 || selectedTgt.Natom() != nAtoms_) {
    mprinterr("Error: B200 RMSD: frame has %i selected atoms, reference %i\n", selectedTgt.Natom(), nAtoms_);
    return 1;
  }
  if (b200_rmsd_1vN_push_f64(handle_, selectedTgt.xAddress(), (size_t)3 * (size_t)nAtoms_, 1)) return b200_err("rmsd push");
  if (b200_rmsd_1vN_flush(handle_, &rmsd, rot, trans, 0)) return b200_err("rmsd flush");
  return 0;
}

int Cpptraj::B200::Rmsd1vN::Coords(const float* base, size_t strideFloats, int nFrames, double* rmsd, double* rot, double* trans) {
  if (handle_ == 0) { mprinterr("Internal Error: B200 RMSD: Coords() before Begin().\n"); return 1; }
  if (nFrames < 1) return 0;
  if (b200_rmsd_1vN_push_f32(handle_, base, strideFloats, nFrames)) return b200_err("rmsd push");
  if (b200_rmsd_1vN_flush(handle_, rmsd, rot, trans, 0)) return b200_err("rmsd flush");
  return 0;
}

int Cpptraj::B200::HierAgglo(const float* triangle, int nCached, int linkage, int targetClusters, double epsilon,
                             std::vector<int>& mergeInto, std::vector<int>& mergeFrom, std::vector<float>& findMin)
{
  if (Init()) return 1;
  mergeInto.assign( (size_t)std::max(nCached, 1), 0 );
  mergeFrom.assign( (size_t)std::max(nCached, 1), 0 );
  findMin.assign( (size_t)std::max(nCached, 1), 0.0f );
  int nCalls = 0, nMerges = 0;
  Timer t_call;
  t_call.Start();
  int err = b200_hieragglo(triangle, nCached, linkage, targetClusters, epsilon, &mergeInto[0], &mergeFrom[0], &findMin[0],
                           &nCalls, &nMerges);
  t_call.Stop();
  if (err) return b200_err("hieragglo");
  mergeInto.resize( (size_t)nMerges );
  mergeFrom.resize( (size_t)nMerges );
  findMin.resize( (size_t)nCalls );
  mprintf("\tB200: %i merges of %i initial clusters on the device in %.4f s.\n", nMerges, nCached, t_call.Total());
  return 0;
}

int Cpptraj::B200::RmsAvgCorr(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
                              Frame const* fixedRef, std::vector<int> const& windows, std::vector<double>& avg,
                              std::vector<double>& sd)
{
  if (Init()) return 1;
  avg.assign( windows.size(), 0.0 );
  sd.assign( windows.size(), 0.0 );
  if (windows.empty() || crd.Size() < 1) return 0;
  if (fixedRef != 0 && fixedRef->Natom() != mask.Nselected()) {
    mprinterr("Error: B200 RMSD: # target atoms (%i) != # reference atoms (%i)\n", mask.Nselected(), fixedRef->Natom());
    return 1;
  }
  Timer t_call;
  t_call.Start();
  int err = b200_rmsavgcorr(crd.RawFrames(), crd.FrameStride(), (int)crd.Size(), &(mask.Selected()[0]), mask.Nselected(),
                            ptr_or_null(mass), (fixedRef != 0 ? fixedRef->xAddress() : 0), &windows[0], (int)windows.size(),
                            &avg[0], &sd[0]);
  t_call.Stop();
  if (err) return b200_err("rmsavgcorr");
  mprintf("\tB200: %zu window sizes over %zu frames on the device in %.4f s.\n", windows.size(), crd.Size(), t_call.Total());
  return 0;
}

int Cpptraj::B200::RmsAvgCorrPacked(DataSet_Coords& crd, AtomMask const& mask, std::vector<double> const& mass,
                                    Frame const* fixedRef, std::vector<int> const& windows, std::vector<double>& avg,
                                    std::vector<double>& sd)
{
  if (Init()) return 1;
  avg.assign( windows.size(), 0.0 );
  sd.assign( windows.size(), 0.0 );
  if (windows.empty() || crd.Size() < 1) return 0;
  const int nsel = mask.Nselected();
  if (fixedRef != 0 && fixedRef->Natom() != nsel) {
    mprinterr("Error: B200 RMSD: # target atoms (%i) != # reference atoms (%i)\n", nsel, fixedRef->Natom());
    return 1;
  }
  std::vector<float> buf;
  b200_pack_selected( crd, mask, buf );
  std::vector<int> ident( (size_t)nsel );
  for (int i = 0; i < nsel; i++) ident[i] = i;
  Timer t_call;
  t_call.Start();
  int err = b200_rmsavgcorr(&buf[0], (size_t)3 * (size_t)nsel, (int)crd.Size(), &ident[0], nsel, ptr_or_null(mass),
                            (fixedRef != 0 ? fixedRef->xAddress() : 0), &windows[0], (int)windows.size(), &avg[0], &sd[0]);
  t_call.Stop();
  if (err) return b200_err("rmsavgcorr");
  mprintf("\tB200: %zu window sizes over %zu frames (selected atoms read once) on the device in %.4f s.\n", windows.size(),
          crd.Size(), t_call.Total());
  return 0;
}

int Cpptraj::B200::CacheClusterSums(const float* triangle, int nCached, std::vector<int> const& members,
                                    std::vector<int> const& offsets, std::vector<double>& cum, std::vector<double>* up,
                                    std::vector<double>* up2)
{
  if (Init()) return 1;
  cum.assign( members.size(), 0.0 );
  if (up != 0) up->assign( members.size(), 0.0 );
  if (up2 != 0) up2->assign( members.size(), 0.0 );
  if (members.empty() || offsets.size() < 2) return 0;
  int err = b200_cache_cluster_sums(triangle, nCached, &members[0], &offsets[0], (int)offsets.size() - 1, &cum[0],
                                    (up != 0 ? &(*up)[0] : 0), (up2 != 0 ? &(*up2)[0] : 0));
  if (err) return b200_err("cluster sums over the cache");
  return 0;
}

int Cpptraj::B200::CacheClusterLinks(const float* triangle, int nCached, std::vector<int> const& label, int K,
                                     std::vector<double>& mn, std::vector<double>& mx, std::vector<double>& sum,
                                     std::vector<long long>& count)
{
  if (Init()) return 1;
  size_t K2 = (size_t)K * (size_t)K;
  mn.assign(K2, 0.0); mx.assign(K2, 0.0); sum.assign(K2, 0.0); count.assign(K2, 0);
  if (K < 1 || (int)label.size() != nCached) return 1;
  int err = b200_cache_cluster_links(triangle, nCached, &label[0], K, &mn[0], &mx[0], &sum[0], &count[0]);
  if (err) return b200_err("cluster linkage over the cache");
  return 0;
}

int Cpptraj::B200::ResidentCache::Begin(const float* triangle, int nCached) {
  if (Init()) return 1;
  if (tri_ != 0) { b200_cache_resident_end( tri_ ); tri_ = 0; }
  if (triangle == 0 || nCached < 2) return 0;
  if (b200_cache_resident_begin(triangle, nCached)) {
    mprintf("Warning: B200: pairwise cache not kept on the device (%s); calls will upload it.\n", b200_last_error());
    return 1;
  }
  tri_ = triangle;
  return 0;
}

Cpptraj::B200::ResidentCache::~ResidentCache() { if (tri_ != 0) b200_cache_resident_end( tri_ ); }
