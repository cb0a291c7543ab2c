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
/// Per-selected-atom masses of a Frame set up with SetupFrameFromMask (Frame.cpp:502-512).
std::vector<double> MassesOf(Frame const&);
/** rms2d (Analysis_Rms2d::Calculate_2D, Analysis_Rms2d.cpp:196-295). \a out must already be allocated by the caller
  * (AllocateTriangle(totaltgt) or Allocate2D(totalref,totaltgt)). Empty mass vectors mean no mass weighting.
  * \return 0 if OK, 1 after mprinterr.
  */
int Rms2d(DataSet_Coords_CRD const& tgt, AtomMask const& tgtMask, std::vector<double> const& tgtMass,
          DataSet_Coords_CRD const& ref, AtomMask const& refMask, std::vector<double> const& refMass,
          bool fullMatrix, bool fit, DataSet_MatrixFlt& out);
/** Pairwise-cache fill (MetricArray::calcFrameDistances, Cluster/MetricArray.cpp:766-801) for a single Metric_RMS.
  * \a triangle is DataSet_PairwiseCache_MEM::Ptr() after SetupCache sized it for framesToCache.size() frames.
  */
int CacheFill(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass,
              bool fit, Cluster::Cframes const& framesToCache, float* triangle);
/** Frame-to-centroid RMSDs for a single Metric_RMS (Metric_RMS::FrameCentroidDist, Cluster/Metric_RMS.cpp:75-81), all
  * frames of \a frames at once: the body of List::AddFramesByCentroid (Cluster/List.cpp:160-207) and of the k-means
  * assignment step.  \a centroidFrames: Centroid_Coord::Cframe() of every cluster, in cluster order.
  * \a closest receives, per frame, the index of the nearest centroid (first minimum wins).
  */
int ClosestCentroids(DataSet_Coords_CRD const& crd, AtomMask const& mask, std::vector<double> const& mass, bool fit,
                     Cluster::Cframes const& frames, std::vector<Frame const*> const& centroidFrames,
                     std::vector<int>& closest, std::vector<double>& closestDist);
/** One-vs-many RMSD for Action_Rmsd (Action_Rmsd.cpp:321-417) when coordinates are not modified (nomod / nofit):
  * the selected atoms of every frame (Action_Rmsd::tgtFrame_ after SetCoordinates) are buffered on the host and
  * pushed to the device in batches; results are appended to the DataSet in frame order at Flush().
  */
class Rmsd1vN {
  public:
    Rmsd1vN() : handle_(0), nAtoms_(0), nBuffered_(0), best_(-1) {}
    ~Rmsd1vN();
    /// \param selectedRef REF_.SelectedRef(): selected atoms, already centred when fitting (ReferenceAction.cpp:155-169)
    /// \param massFrame Frame whose per-atom masses weight the fit (the TARGET frame's, Frame.cpp:1184-1208)
    int Begin(Frame const& selectedRef, Frame const& massFrame, bool fit, bool useMass);
    bool Active() const { return handle_ != 0; }
    /// Buffer the selected atoms of one frame (Frame set up from the target mask).
    int Push(Frame const& selectedTgt);
    /// Push what is buffered, wait, append all pending RMSDs to \a rmsd in push order. \return 0 if OK.
    int Flush(DataSet_double& rmsd);
    /// Index (over all frames pushed) of the smallest RMSD so far.
    long BestFrame() const { return best_; }
  private:
    Rmsd1vN(Rmsd1vN const&);
    Rmsd1vN& operator=(Rmsd1vN const&);
    int pushBuffer();
    static const unsigned int BATCH = 512;   ///< frames per device push
    b200_1vN* handle_;
    int nAtoms_;
    unsigned int nBuffered_;
    long best_;
    std::vector<double> buffer_;             ///< BATCH x 3*nAtoms_ selected coordinates
};
}
}
#endif
