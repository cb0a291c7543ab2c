#ifndef INC_CUDA_B200_RMSD_H
#define INC_CUDA_B200_RMSD_H
/** Host glue between cpptraj's RMSD plugins and the C ABI of the B200 library (b200_rmsd.h).
  * Compiled only with -DCUDA_B200.  The plugins keep parsing keywords and masks and own the
  * DataSets; these functions hand their raw buffers to the device path and map error codes to
  * mprinterr + a non-zero return (Analysis::ERR / Action::ERR at the call sites).
  * There is no CPU fallback: a failure here is an error of the command.
  */
#include <vector>
#include <cstddef>
class DataSet_Coords;
class DataSet_Coords_CRD;
class DataSet_MatrixFlt;
class DataSet_double;
class AtomMask;
class Frame;
struct b200_1vN;
namespace Cpptraj {
namespace Cluster { class Cframes; }
namespace B200 {
/// One-time device probe (pattern of Cpptraj.cpp:120-135). CPPTRAJ_B200_NGPU limits the device count. \return 0 if OK.
int Init();
/// Start the device probe on a background thread (from a command's Setup()/Init(), before trajectory processing); Init() joins it.
void InitAsync();
/// Per-selected-atom masses of a Frame set up with SetupFrameFromMask (Frame.cpp:502-512).
std::vector<double> MassesOf(Frame const&);
/** rms2d (Analysis_Rms2d::Calculate_2D, Analysis_Rms2d.cpp:196-295). \a out must already be allocated by the caller
  * (AllocateTriangle(totaltgt) or Allocate2D(totalref,totaltgt)). Empty mass vectors mean no mass weighting.
  * \return 0 if OK, 1 after mprinterr.
  */
int Rms2d(DataSet_Coords_CRD const& tgt, AtomMask const& tgtMask, std::vector<double> const& tgtMass,
          DataSet_Coords_CRD const& ref, AtomMask const& refMask, std::vector<double> const& refMass,
          bool fullMatrix, bool fit, DataSet_MatrixFlt& out);
/** The same for COORDS-group sets that are NOT in-memory float arrays -- DataSet::TRAJ (frames stay on disk: `loadtraj`,
  * `2drms ... reftraj <file>`), reference frames, FRAMES sets: the reference reads every target frame once PER REFERENCE
  * FRAME (Analysis_Rms2d.cpp:270-274); here the selected atoms of every frame are read ONCE, stored as float (what
  * createcrd / loadcrd would hold, CompactFrameArray.cpp:200-214) and handed to the same library calls.
  */
int Rms2dPacked(DataSet_Coords& tgt, AtomMask const& tgtMask, std::vector<double> const& tgtMass,
                DataSet_Coords& ref, AtomMask const& refMask, std::vector<double> const& refMass,
                bool fullMatrix, bool fit, DataSet_MatrixFlt& out);
/** Pairwise-cache fill (MetricArray::calcFrameDistances, Cluster/MetricArray.cpp:766-801) for a single Metric_RMS.
  * \a triangle is DataSet_PairwiseCache_MEM::Ptr() after SetupCache sized it for framesToCache.size() frames.
  */
int CacheFill(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
              bool fit, Cluster::Cframes const& framesToCache, float* triangle);
/// The same for a set that is not an in-memory float array (DataSet::TRAJ ...): the selected atoms of the frames to cache are
/// read once (the reference reads two frames from the set per pair, Cluster/Metric_RMS.cpp:53-63).
int CacheFillPacked(DataSet_Coords& crd, AtomMask const& mask, std::vector<double> const& mass,
                    bool fit, Cluster::Cframes const& framesToCache, float* triangle);
/** Frame-to-centroid RMSDs for a single Metric_RMS (Metric_RMS::FrameCentroidDist, Cluster/Metric_RMS.cpp:75-81), all
  * frames of \a frames at once: the body of List::AddFramesByCentroid (Cluster/List.cpp:160-207) and of the k-means
  * assignment step.  \a centroidFrames: Centroid_Coord::Cframe() of every cluster, in cluster order.
  * \a closest receives, per frame, the index of the nearest centroid (first minimum wins).
  */
int ClosestCentroids(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass, bool fit,
                     Cluster::Cframes const& frames, std::vector<Frame const*> const& centroidFrames,
                     std::vector<int>& closest, std::vector<double>& closestDist);
/** Frame-to-centroid RMSDs of many frames against ONE centroid (Metric_RMS::FrameCentroidDist, Cluster/Metric_RMS.cpp:75-81):
  * the loops of BestReps::FindBestRepFrames_Centroid (Cluster/BestReps.cpp:310-316), Node::CalcAvgToCentroid
  * (Cluster/Node.cpp:89-101) and Algorithm_Kmeans::FindSeedsFromClusters (Cluster/Algorithm_Kmeans.cpp:296-303).
  * \a dist receives one value per frame of \a frames, in order.
  */
int CentroidDists(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass, bool fit,
                  Cluster::Cframes const& frames, Frame const& centroidFrame, std::vector<double>& dist);
/** Centroid building (Metric_RMS::CalculateCentroid, Cluster/Metric_RMS.cpp:86-113) for many clusters at once: every
  * cluster's frames are fitted, in order, to the running sum of the frames before them and averaged -- one CTA per
  * cluster on the device.  \a clusterFrames: the frame lists, \a centroids: the Centroid_Coord::Cframe() to fill
  * (set up for mask.Nselected() atoms), same order.
  */
int BuildCentroids(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass, bool fit,
                   std::vector<Cluster::Cframes const*> const& clusterFrames, std::vector<Frame*> const& centroids);
/** rmsavgcorr (Analysis_RmsAvgCorr::Analyze, Analysis_RmsAvgCorr.cpp:119-316): mean and standard deviation of the RMSDs of
  * the running-averaged frames for every window size of \a windows.  \a fixedRef: the pre-centred reference frame
  * (selected atoms) or 0 for the "first" mode.  \a avg / \a sd are sized to windows.size().
  */
int RmsAvgCorr(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass, Frame const* fixedRef,
               std::vector<int> const& windows, std::vector<double>& avg, std::vector<double>& sd);
/// The same for a set that is not an in-memory float array (TRAJ ...): the selected atoms of every frame are read once (the
/// reference reads every frame again for every window size, Analysis_RmsAvgCorr.cpp:246-249).
int RmsAvgCorrPacked(DataSet_Coords& crd, AtomMask const& mask, std::vector<double> const& mass, Frame const* fixedRef,
                     std::vector<int> const& windows, std::vector<double>& avg, std::vector<double>& sd);
/** Hierarchical agglomerative clustering on an in-memory pairwise cache (Algorithm_HierAgglo::DoClustering /
  * MergeClosest, Cluster/Algorithm_HierAgglo.cpp:97-245, with the DynamicMatrix bookkeeping): all merges run on the
  * device in one launch.  \a triangle: DataSet_PairwiseCache_MEM::Ptr() for \a nCached frames; \a linkage in
  * Algorithm_HierAgglo::LINKAGETYPE order.  Per MergeClosest call m: \a findMin[m]; for the calls that merged,
  * \a mergeInto[m] (C1, kept) and \a mergeFrom[m] (C2, removed) as initial-cluster numbers.
  */
int HierAgglo(const float* triangle, int nCached, int linkage, int targetClusters, double epsilon,
              std::vector<int>& mergeInto, std::vector<int>& mergeFrom, std::vector<float>& findMin);
/** Sums of cached distances over cluster members (BestReps cumulative distance, Cluster/BestReps.cpp:157-166,243-257; the
  * within-cluster average of Output::Summary, Cluster/Output.cpp:195-225).  \a members: cache indices of all clusters'
  * frames concatenated in list order, \a offsets: start of every cluster (+ the end).  \a cum[p]: sum of the distances of
  * member p to the other members of its cluster, added in list order; \a up / \a up2 (may be 0): sums of d and d*d over
  * the members after p.
  */
int CacheClusterSums(const float* triangle, int nCached, std::vector<int> const& members, std::vector<int> const& offsets,
                     std::vector<double>& cum, std::vector<double>* up, std::vector<double>* up2);
/** Minimum, maximum, sum and count of the cached distances between every two clusters (Algorithm_HierAgglo::ClusterDistance
  * for all pairs, Cluster/Algorithm_HierAgglo.cpp:353-408).  \a label[f]: cluster index of cached frame f or -1.  Tables
  * are K x K row-major, entries c1 < c2.
  */
int CacheClusterLinks(const float* triangle, int nCached, std::vector<int> const& label, int K, std::vector<double>& mn,
                      std::vector<double>& mx, std::vector<double>& sum, std::vector<long long>& count);
/** Keeps an in-memory pairwise cache on the device while it exists (clustering, best representatives and the summary
  * read the same, unchanging triangle: Cluster/Control.cpp:770-1019).
  */
class ResidentCache {
  public:
    ResidentCache() : tri_(0) {}
    ~ResidentCache();
    int Begin(const float* triangle, int nCached);
  private:
    ResidentCache(ResidentCache const&);
    ResidentCache& operator=(ResidentCache const&);
    const float* tri_;
};
/** Keeps the selected span of a COORDS set on the device while it exists (Cluster::Control::Run, Cluster/Control.cpp:690-830):
  * the centroid / frame-to-centroid calls of a clustering run then skip the upload.  The frames must not change meanwhile.
  */
class ResidentCoords {
  public:
    ResidentCoords() : crd_(0) {}
    ~ResidentCoords();
    /// \return 0 if OK; on failure the calls simply upload what they need.
    int Begin(DataSet_Coords_CRD const&, AtomMask const&);
  private:
    ResidentCoords(ResidentCoords const&);
    ResidentCoords& operator=(ResidentCoords const&);
    const float* crd_;
};
/** One-vs-many RMSD for Action_Rmsd (Action_Rmsd.cpp:321-417).  Two uses:
  *  - frame by frame (One): the selected atoms of the current frame (Action_Rmsd::tgtFrame_ after SetCoordinates) are
  *    pushed and the result is read back at once -- RMSD, rotation matrix, target-to-origin translation -- so that
  *    DoAction() continues exactly as with Frame::RMSD_CenteredRef (savematrices, savevectors, coordinate
  *    modification, DataSet::Add by frame number);
  *  - a whole in-memory COORDS set (Coords): raw float frames, gathered by the mask on the device, one pass
  *    ('crdaction <set> rms ...', Exec_CrdAction.cpp:78-98).
  */
class Rmsd1vN {
  public:
    Rmsd1vN() : handle_(0), nAtoms_(0) {}
    ~Rmsd1vN();
    /// \param selectedRef REF_.SelectedRef(): selected atoms, already centred when fitting (ReferenceAction.cpp:155-169)
    /// \param massFrame Frame whose per-atom masses weight the fit (the TARGET frame's, Frame.cpp:1184-1208)
    /// \param atomIdx Selected atom numbers when raw COORDS frames will be pushed (Coords); 0 when frames arrive gathered (One)
    int Begin(Frame const& selectedRef, Frame const& massFrame, const int* atomIdx, bool fit, bool useMass, bool wantRot);
    bool Active() const { return handle_ != 0; }
    void End();
    /// Replace the reference (reftraj / previous).
    int SetRef(Frame const& selectedRef);
    /// Fit one frame (set up from the target mask). \a rot (9, row-major U) and \a trans (3) are set when fitting with wantRot.
    int One(Frame const& selectedTgt, double& rmsd, double* rot, double* trans);
    /// Fit \a nFrames raw COORDS frames (frame i at base + i*strideFloats). Outputs sized nFrames, 9*nFrames, 3*nFrames (rot/trans may be 0).
    int Coords(const float* base, size_t strideFloats, int nFrames, double* rmsd, double* rot, double* trans);
  private:
    Rmsd1vN(Rmsd1vN const&);
    Rmsd1vN& operator=(Rmsd1vN const&);
    b200_1vN* handle_;
    int nAtoms_;
};
}
}
#endif
