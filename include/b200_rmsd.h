/* b200_rmsd.h -- C ABI of the B200-native best-fit RMSD path for cpptraj.
 *
 * This is the drop-in boundary: plain pointers and
 * sizes, no C++/torch types, no exceptions.  cpptraj keeps parsing commands,
 * masks and owning DataSets; inside its three RMSD plugins the new host code
 * (src/cuda_b200/) calls these entry points instead of the
 * OpenMP pair loops.  File:line citations are relative to the cpptraj tree.
 *
 * Conventions
 *   - all `crd` pointers are cpptraj COORDS storage: float32, one frame per
 *     `frameStrideFloats` floats, position component first, xyz interleaved
 *     (CompactFrameArray, src/CompactFrameArray.cpp:110-147,244-262);
 *   - `atomIdx` is AtomMask::Selected() (src/AtomMask.h:32): atom numbers,
 *     the library reads crd[frame*stride + 3*atomIdx[k] + {0,1,2}];
 *   - `mass` is one double per SELECTED atom (Frame::SetupFrameFromMask,
 *     src/Frame.cpp:502-512) or NULL for unit weights ("mass" keyword absent);
 *   - `fit` != 0: best-fit RMSD (Frame::RMSD_CenteredRef, src/Frame.cpp:1137),
 *     `fit` == 0: no-fit RMSD (Frame::RMSD_NoFit, src/Frame.cpp:1279);
 *   - every function returns B200_OK (0) or a B200_ERR_* code; the message is
 *     available from b200_last_error().  The caller maps non-zero to
 *     mprinterr + Analysis::ERR / Action::ERR (src/Analysis.h:81, src/Action.h:32);
 *   - the caller owns every host buffer; the library owns device memory,
 *     streams and pinned staging.  Calls are blocking and must come from
 *     outside any OpenMP region.  There is NO CPU fallback: without a usable
 *     sm_100 device every compute entry point fails with B200_ERR_NO_DEVICE.
 */
#ifndef B200_RMSD_H
#define B200_RMSD_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_NO_DEVICE 1 /* no CUDA device / not initialised            */
#define B200_ERR_CUDA 2      /* a CUDA runtime call failed                   */
#define B200_ERR_ARG 3       /* invalid argument                             */
#define B200_ERR_NOMEM 4     /* host or device allocation failed             */
#define B200_ERR_STATE 5     /* call sequence error on a streaming handle    */

/* ---- lifetime -------------------------------------------------------------
 * Device probe at first use, like src/Cpptraj.cpp:120-135,168-179 does for the
 * existing CUDA build.  ngpu_requested <= 0 means "all visible devices".
 * b200_init may be called again to change the device count. */
int b200_init(int ngpu_requested, int* ngpu_used);
/* Same, with explicit CUDA device ordinals (one process per GPU: pass LOCAL_RANK). */
int b200_init_devices(const int* deviceIds, int n);
/* Optional, after b200_init: reserves the device chunk buffers and loads the main kernels, so that the first call of
 * a process does not pay for it (a cpptraj command does it on a background thread at set-up, while the trajectory is
 * read; the pattern of src/Cpptraj.cpp:120-135 is a synchronous probe). */
int b200_warmup(void);
void b200_shutdown(void);
const char* b200_last_error(void);
int b200_version(void);
int b200_num_devices(void);

/* ---- rms2d / pairwise cache: triangle -------------------------------------
 * Replaces the pair loops of Analysis_Rms2d::Calculate_2D when
 * !calculateFullMatrix (src/Analysis_Rms2d.cpp:247-287) and of
 * MetricArray::calcFrameDistances (src/Cluster/MetricArray.cpp:777-795;
 * frameIdx = framesToCache.Ptr(), src/Cluster/Cframes.h:31).
 * frameIdx NULL means frames 0..nFrames-1.  outTri is Matrix<float>::Ptr() of
 * a TRIANGLE matrix of nFrames columns: element (i<j) at
 * nFrames*i - i*(i+1)/2 + j - i - 1 (src/Matrix.h:110-122).
 * Work is split over all initialised devices (one host thread each; the shards agree on one fixed-point grid).
 * Host buffers may be pageable (cpptraj's std::vector<float> / new float[]): they are staged through pinned ring
 * slots by a pool of host threads (env B200_HOST_THREADS); pinned buffers are read and written by DMA directly. */
int b200_rms2d_tri(const float* crd, size_t frameStrideFloats, int nFramesTotal,
                   const int* frameIdx, int nFrames,
                   const int* atomIdx, int nAtoms,
                   const double* mass, int fit, float* outTri);

/* Same, but computes only shard `shardRank` of `shardCount` on the CURRENT
 * process's device 0 (one process per GPU, e.g. under torchrun / MPI).  The
 * shard is a contiguous band of matrix rows, hence a contiguous range
 * [*firstElt, *firstElt + *nElts) of outTri; elements outside are untouched.
 * outTri may be NULL to query the range only. */
int b200_rms2d_tri_shard(const float* crd, size_t frameStrideFloats, int nFramesTotal,
                         const int* frameIdx, int nFrames,
                         const int* atomIdx, int nAtoms,
                         const double* mass, int fit,
                         int shardRank, int shardCount,
                         float* outTri, size_t* firstElt, size_t* nElts);

/* ---- rms2d: full matrix ------------------------------------------------------
 * Calculate_2D when calculateFullMatrix (reftraj and/or tgt mask != ref mask,
 * src/Analysis_Rms2d.cpp:204-209).  outFull[itgt*nRef + iref]
 * (Allocate2D(totalref,totaltgt), src/Matrix.h:94-96).  massTgt weights the
 * covariance and total mass (target frame's Mass_, src/Frame.cpp:1184-1208);
 * massRefCentering is used only to centre each reference frame
 * (SelectedRef.CenterOnOrigin, src/Analysis_Rms2d.cpp:265-266).  Both NULL
 * when the "mass" keyword is absent.  Target rows are split over all initialised devices. */
int b200_rms2d_full(const float* crdTgt, size_t strideTgt, int nTgt, const int* atomIdxTgt,
                    const float* crdRef, size_t strideRef, int nRef, const int* atomIdxRef,
                    int nAtoms, const double* massTgt, const double* massRefCentering,
                    int fit, float* outFull);

/* ---- rmsd action: one-vs-many, streaming ------------------------------------
 * Action_Rmsd::{Setup,DoAction,Print} (src/Action_Rmsd.cpp:321-417).
 * refSelected: the selected reference atoms (3*nAtoms doubles) exactly as
 * ReferenceAction holds them in SelectedRef(): already centred on the origin
 * when fitting (src/ReferenceAction.cpp:155-169), raw when fit == 0.
 * Frames are pushed in chunks, either as cpptraj Frame coordinates (double,
 * all atoms, xyz interleaved: Frame::xAddress()) or as COORDS storage (float).
 * Results come back in push order (DataSet_double::Add is append-only,
 * src/DataSet_double.cpp:14-20): rmsdOut[n]; rotOut[9n] = the rotation U of
 * RMSD_CenteredRef row-major (NULL if not wanted); transOut[3n] = the
 * target->origin translation "Trans" (NULL if not wanted).
 * A push returns when the caller's buffer may be reused (pageable frames are packed into pinned staging, pinned
 * ones have left by DMA); the kernels run asynchronously until flush.  Chunks go round-robin to all initialised
 * devices; total mass < 1e-14 gives -1 for every frame (src/Frame.cpp:1160-1163). */
typedef struct b200_1vN b200_1vN;
int b200_rmsd_1vN_begin(const double* refSelected, const int* atomIdx, int nAtoms,
                        const double* mass, int fit, int wantRot, b200_1vN** handle);
int b200_rmsd_1vN_push_f64(b200_1vN* h, const double* xyz, size_t frameStrideDoubles, int nFrames);
int b200_rmsd_1vN_push_f32(b200_1vN* h, const float* crd, size_t frameStrideFloats, int nFrames);
/* Replaces the reference of a live handle (reftraj / previous modes, src/ReferenceAction.h:73-88: the reference
 * changes from frame to frame).  Frames already pushed keep the reference they were pushed against. */
int b200_rmsd_1vN_set_ref(b200_1vN* h, const double* refSelected);
/* Number of frames pushed and not yet flushed. */
long b200_rmsd_1vN_pending(const b200_1vN* h);
/* Blocks until all pushed frames are done; argminFrame (nullable) receives the
 * index (over ALL frames pushed since begin) of the smallest RMSD so far,
 * first one wins on ties. */
int b200_rmsd_1vN_flush(b200_1vN* h, double* rmsdOut, double* rotOut, double* transOut,
                        long* argminFrame);
int b200_rmsd_1vN_end(b200_1vN* h);

/* ---- cluster: frame-to-centroid distances ----------------------------------------
 * Metric_RMS::FrameCentroidDist (src/Cluster/Metric_RMS.cpp:75-81) for many frames at once: the consumers are the
 * sieve restore List::AddFramesByCentroid (src/Cluster/List.cpp:160-207), k-means assignment
 * (src/Cluster/Algorithm_Kmeans.cpp:223,298) and the representative-frame search.  centroids: nCentroids x 3*nAtoms
 * doubles, the selected atoms of each Centroid_Coord::Cframe() (already centred on the origin when fitting, as cpptraj
 * keeps them).  frameIdx (nullable) lists the frames, e.g. the sieved-out ones.  Outputs (each nullable):
 * distOut[f*nCentroids + k]; closestOut[f] = nearest centroid, first minimum wins (List.cpp:183-189);
 * closestDistOut[f] its distance.  Three or more centroids with a fitted RMSD run as ONE frames x centroids contraction
 * on the tcgen05 engine (frames read and quantised once); fewer, or nofit, as one streaming pass per centroid.  Large
 * frame lists are split over all initialised devices. */
int b200_rmsd_frames_to_centroids(const float* crd, size_t frameStrideFloats, int nFramesTotal,
                                  const int* frameIdx, int nFrames,
                                  const int* atomIdx, int nAtoms, const double* mass, int fit,
                                  const double* centroids, int nCentroids,
                                  double* distOut, int* closestOut, double* closestDistOut);

/* ---- cluster: COORDS resident on the device -----------------------------------------------
 * A clustering run calls the two functions above (and the one below) many times on the same, unchanging COORDS set
 * (Cluster::Control::Run, src/Cluster/Control.cpp:690-830).  Between begin and end the leading 3*(max atomIdx + 1)
 * floats of every frame stay on device 0 and calls that pass the same `crd` (same stride and frame count, atoms
 * within that span) skip the upload.  The caller must not modify the frames in between.  end(NULL) drops any. */
int b200_coords_resident_begin(const float* crd, size_t frameStrideFloats, int nFramesTotal,
                               const int* atomIdx, int nAtoms);
int b200_coords_resident_end(const float* crd);

/* ---- cluster: centroid building with fit ----------------------------------------------
 * Metric_RMS::CalculateCentroid (src/Cluster/Metric_RMS.cpp:86-113) for nClusters clusters at once: cluster k owns
 * frames[offsets[k] .. offsets[k+1]); its first frame starts a running sum (centred when fit != 0), every further
 * frame is fitted to that sum, rotated and added, and the sum is divided by the frame count.  centroidsOut:
 * nClusters x 3*nAtoms doubles (the selected atoms of each Centroid_Coord::Cframe()); clusters without frames are
 * left as zeros.  The scan over a cluster's frames is sequential by definition; clusters run side by side (one CTA
 * each).  Runs on device 0. */
int b200_rmsd_build_centroids(const float* crd, size_t frameStrideFloats, int nFramesTotal,
                              const int* frames, const int* offsets, int nClusters,
                              const int* atomIdx, int nAtoms, const double* mass, int fit,
                              double* centroidsOut);

/* ---- rmsavgcorr: RMSD of running-averaged coordinates -------------------------------------------------
 * Analysis_RmsAvgCorr::Analyze (src/Analysis_RmsAvgCorr.cpp:119-316).  For every window size windows[w] the frames
 * t = 0 .. nFrames - W are replaced by the average of frames t .. t+W-1 (selected atoms), each averaged frame is fitted
 * (Frame::RMSD_CenteredRef, masses of the selected atoms when mass != NULL) to the reference, and avgOut[w] / sdOut[w]
 * receive the mean and the standard deviation of those RMSDs as the reference forms them (:196-205, :289-297).
 * refSelected NULL: the "first" mode -- the reference of a window size is its own first averaged frame, centred with
 * the same weights (:258-272); else 3*nAtoms doubles, the fixed reference exactly as the caller centred it (:86-90).
 * Window size 1 is the plain trajectory (:176-205).  All frames of crd are used.  64 or more window sizes per device: the
 * window sizes are dealt round-robin to all initialised devices (each builds its own prefix sums). */
int b200_rmsavgcorr(const float* crd, size_t frameStrideFloats, int nFrames, const int* atomIdx, int nAtoms,
                    const double* mass, const double* refSelected, const int* windows, int nWindows,
                    double* avgOut, double* sdOut);

/* ---- cluster: hierarchical agglomerative clustering on the pairwise cache ---------------------------
 * The merge loop of Algorithm_HierAgglo::DoClustering / MergeClosest (src/Cluster/Algorithm_HierAgglo.cpp:97-245) with
 * the Cluster::DynamicMatrix bookkeeping (src/Cluster/DynamicMatrix.h:43-126, DynamicMatrix.cpp:7-33), run on the
 * device by one thread-block cluster.  tri: the pairwise cache, DataSet_PairwiseCache_MEM::Ptr() (triangle of nFrames
 * columns, layout above; pageable or pinned); one initial cluster per cached frame, Num() = cache index
 * (buildInitialClusters, :86-95).  linkage: 0 single, 1 average, 2 complete (LINKAGETYPE,
 * src/Cluster/Algorithm_HierAgglo.h:35); targetClusters / epsilon: nclusters_ / epsilon_ after DoClustering's defaults
 * (-1 -> 1, -1.0 -> DBL_MAX).  Outputs (caller-allocated, nFrames entries each): for MergeClosest call m the value
 * FindMin returned, findMin[m] (what `epsilonplot` prints), and -- for the calls that merged -- the cluster kept,
 * mergeInto[m] (the lower Num), and the cluster merged into it, mergeFrom[m].  *nCalls = MergeClosest calls made,
 * *nMerges = merges (nCalls - 1 when the last call stopped on epsilon).  The caller replays the merges on its cluster
 * list (Node::MergeFrames, List::RemoveCluster).  Which pair merges, ties included, is the reference's choice exactly;
 * single and complete linkage distances are exact, average-linkage sums are kept in double (see hieragglo.cuh).
 * Device memory: the uploaded triangle plus a symmetric nFrames x nFrames working matrix, 4 bytes per entry (12 for
 * average linkage: 50,000 frames = 5 + 10 (+ 20) GB).  Runs on device 0. */
int b200_hieragglo(const float* tri, int nFrames, int linkage, int targetClusters, double epsilon,
                   int* mergeInto, int* mergeFrom, float* findMin, int* nCalls, int* nMerges);

/* ---- cluster: consumers of the pairwise cache in the post-processing --------------------------------------------
 * Between begin and end the cache triangle (DataSet_PairwiseCache_MEM::Ptr(), nCached frames) stays on device 0 and the
 * calls of this section and b200_hieragglo that pass the same `tri` skip the upload (Cluster::Control::Run,
 * src/Cluster/Control.cpp:755-830: clustering, best representatives and the summary read the same, unchanging cache).
 * end(NULL) drops any. */
int b200_cache_resident_begin(const float* tri, int nCached);
int b200_cache_resident_end(const float* tri);
/* Sums of cached distances over the members of every cluster.  members: cache indices of the frames of all clusters,
 * concatenated in list order; cluster c owns members[offsets[c] .. offsets[c+1]).  cumOut[p] = sum over the other members
 * q of its cluster of d(p, q), added in list order in double exactly like the inner loop of
 * BestReps::FindBestRepFrames_CumulativeDist (src/Cluster/BestReps.cpp:157-166, :243-257); upOut[p] / up2Out[p] (nullable)
 * = sums of d and d*d over the members after p: summed over p they are the numerators of the within-cluster average and
 * its standard deviation (Cluster::Output::Summary, src/Cluster/Output.cpp:195-225). */
int b200_cache_cluster_sums(const float* tri, int nCached, const int* members, const int* offsets, int nClusters,
                            double* cumOut, double* upOut, double* up2Out);
/* Linkage between every two clusters in ONE pass over the triangle (Algorithm_HierAgglo::ClusterDistance for all cluster
 * pairs, src/Cluster/Algorithm_HierAgglo.cpp:353-408, as Cluster::Output::Summary needs them, Output.cpp:164-172):
 * label[f] = cluster of cached frame f, or -1 for a frame that takes no part (not clustered, or sieved out and excluded).
 * Outputs are nClusters x nClusters row-major, entries [c1*nClusters + c2] with c1 < c2: minimum, maximum, sum and
 * count of the cached distances between members of c1 and c2 (the sum is accumulated in double, in no fixed order). */
int b200_cache_cluster_links(const float* tri, int nCached, const int* label, int nClusters,
                             double* minOut, double* maxOut, double* sumOut, long long* countOut);

/* ---- device-resident variants (benchmarks, pipelines that keep COORDS in HBM)
 * All pointers are DEVICE pointers on the current device; `stream` is a
 * cudaStream_t (NULL = default stream); asynchronous w.r.t. the host.
 * d_outTri holds the whole triangle (nFrames*(nFrames-1)/2 floats). */
int b200_dev_rms2d_tri(const float* d_crd, size_t frameStrideFloats,
                       const int* d_frameIdx, int nFrames,
                       const int* d_atomIdx, int nAtoms,
                       const double* d_mass, int fit,
                       int shardRank, int shardCount,
                       float* d_outTri, void* stream);
int b200_dev_rmsd_1vN(const float* d_crd, size_t frameStrideFloats, int nFrames,
                      const int* d_atomIdx, int nAtoms,
                      const double* d_refSelected, const double* d_mass, int fit,
                      double* d_rmsdOut, double* d_rotOut, double* d_transOut, void* stream);

/* ---- shard geometry (host only, no device needed) -----------------------------
 * Rows [*row0,*row1) of the nFrames x nFrames triangle owned by a shard:
 * contiguous, 32-row aligned, balanced by pair count. */
int b200_shard_rows(int nFrames, int shardRank, int shardCount, int* row0, int* row1);

/* ---- instrumentation -------------------------------------------------------------
 * When profiling is enabled the library brackets its kernels with CUDA events
 * on the streams it launches on and accumulates per-kernel device time. */
typedef struct b200_stats {
  double pack_ms;        /* centring/packing kernel                           */
  double pair_ms;        /* pair-tile kernel (sum over launches)              */
  double onevn_ms;       /* one-vs-many kernel                                */
  long   pack_launches;
  long   pair_launches;
  long   onevn_launches;
  double pairs;          /* pair RMSDs produced                               */
  double frames_1vN;     /* frames processed by the one-vs-many kernel        */
  double h2d_bytes;
  double d2h_bytes;
  long   kernel_launches; /* every kernel this library launched (all kinds)     */
  double onevn_stream_ms; /* the streaming kernel of the one-vs-many path alone  */
  long   onevn_stream_launches;
} b200_stats;
void b200_set_profiling(int on);
void b200_reset_stats(void);
void b200_get_stats(b200_stats* out);

/* ---- pair engine ---------------------------------------------------------------
 * Two kernels compute the pair covariances:
 *   1 = FP64 tensor cores (mma.sync DMMA), any input, fit and nofit;
 *   2 = tcgen05 int8 tensor cores on coordinates rounded once per frame to a 24-bit
 *       fixed-point grid (exact integer covariance in TMEM), fit only, used when the
 *       selection's extent leaves enough fractional bits for a worst-case RMSD change
 *       < 8.5e-5 A;
 *   0 = automatic (2 when eligible, else 1) -- the default; env B200_PAIR_ENGINE.
 * Forcing 2 on an ineligible call fails with B200_ERR_ARG. */
int b200_set_pair_engine(int engine);
/* Pins the number of fractional bits of engine 2's fixed-point grid for the following calls (0 = automatic: from
 * the extent of the call's own frames).  One process per GPU: every rank computes a shard of one matrix from the
 * frames it holds; the ranks agree on min(bits) (b200_last_pair_engine reports what a call used) and pin it, so
 * that all shards are rounded to the same grid whatever the GPU count.  A single process driving several devices
 * does this by itself.  A pinned scale too fine for the data fails with B200_ERR_ARG. */
int b200_set_fixed_point_bits(int bits);
/* Engine the last rms2d call used (1 or 2; 0 = none yet) and, for 2, the number of
 * fractional bits of its fixed-point grid. */
int b200_last_pair_engine(int* fractionalBits);

#ifdef __cplusplus
}
#endif
#endif
