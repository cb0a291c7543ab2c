"""cpptraj input decks run both by the UNMODIFIED reference (goldens, tools/make_golden_cpptraj.py, here in the build
container) and by cpptraj.B200 (tests/test_gpu_cpptraj_e2e.py, on the GPU box).  They restate the reference's own
test decks (file:line below) with the ASCII trajectories tz2.crd / tz2.truncoct.crd in place of the NetCDF ones
(this build has no NetCDF).  {D} = directory holding tz2.parm7, tz2.crd, tz2.truncoct.parm7,
tz2.truncoct.crd.  Each deck: name -> (input text, [(output file, kind)]); kind 'table' = numeric columns compared
with a tolerance, 'crd' = Amber ASCII coordinates compared number by number, 'text' = compared after numeric parsing of
every token that is a number."""

DECKS = {
    # test/Test_RMSD/RunTest.sh:14-42 "Basic RMSD tests": first, mass + savevectors combined, reftraj from file and from COORDS
    "rmsd_basic": ("""noprogress
parm {D}/tz2.truncoct.parm7
trajin {D}/tz2.truncoct.crd
rms Res2-11 first :2-11 out rmsd.dat
rms Res2-11_mass first :2-11 out rmsd.mass.dat mass savevectors combined vecsout vecs.dat
rms Res2_11_traj reftraj {D}/tz2.truncoct.crd :2-11 out rmsd.reftraj.dat
run
removedata Res2_11_traj
loadtraj {D}/tz2.truncoct.crd name TZ2
rms Res2_11_traj reftraj TZ2 :2-11 out rmsd.refcoords.dat
""", [("rmsd.dat", "table"), ("rmsd.mass.dat", "table"), ("vecs.dat", "table"), ("rmsd.reftraj.dat", "table"),
      ("rmsd.refcoords.dat", "table")]),
    # test/Test_RMSD/RunTest.sh:44-62 "RMS coordinate rotation/rotation matrices test": norotate, rotate + savematrices, outtraj
    "rmsd_rotate": ("""noprogress
parm {D}/tz2.parm7 [NOWAT]
reference {D}/tz2.crd parm [NOWAT] 1 [first]
parm {D}/tz2.truncoct.parm7 [WAT]
trajin {D}/tz2.truncoct.crd parm [WAT]
strip :WAT
rms NOROT ref [first] norotate @CA
outtraj tz2.norotate.crd parm [WAT]
rms ROT ref [first] out rms.dat @CA savematrices matricesout rmatrices.dat
outtraj tz2.rotate.crd parm [WAT]
""", [("rms.dat", "table"), ("rmatrices.dat", "table"), ("tz2.norotate.crd", "crd"), ("tz2.rotate.crd", "crd")]),
    # test/Test_RMSD/RunTest.sh:64-76 "RMS nomod"
    "rmsd_nomod": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
rms First_CA :2-12@CA out NoMod.dat nomod
trajout NoMod.crd
""", [("NoMod.dat", "table"), ("NoMod.crd", "crd")]),
    # test/Test_RMSD/RunTest.sh:78-91 "RMS fit to previous"
    "rmsd_previous": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
rms ToPrevious :2-12@CA previous out Previous.dat
""", [("Previous.dat", "table")]),
    # src/Exec_CrdAction.cpp:78-98: the rms action on an in-memory COORDS set (fit + coordinate modification), a
    # data set that reads the RMSD set while it is being filled (filter), and the modified coordinates written out
    "crdaction_rms": ("""noprogress
parm {D}/tz2.parm7
loadcrd {D}/tz2.crd name CRD
crdaction CRD rms R1 first @CA,C,N out crd_rms.dat savematrices matricesout crd_rmat.dat
crdaction CRD rms R2 first :2-12@CA nofit mass out crd_rms_nofit.dat crdframes 3,90,4
crdout CRD fitted.crd
""", [("crd_rms.dat", "table"), ("crd_rmat.dat", "table"), ("crd_rms_nofit.dat", "table"), ("fitted.crd", "crd")]),
    # src/Action_Align.cpp:96: align (fit + move, no RMSD output) with a move mask and mass weighting
    "align": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
align :2-12@CA first mass move :1-13
trajout aligned.crd
""", [("aligned.crd", "crd")]),
    # src/Exec_CrdTransform.cpp:75-134: iterative RMS refinement (every iteration fits all frames to a new average)
    "crdtransform_rmsrefine": ("""noprogress
parm {D}/tz2.parm7
loadcrd {D}/tz2.crd name CRD
crdtransform CRD name REFINED rmsrefine mask @CA rmstol 0.0005
crdout REFINED refined.crd
""", [("refined.crd", "crd")]),
    # an action that reads the RMSD data set during trajectory processing: filter on the running set
    "rmsd_filter": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
rms R1 first :2-12@CA out filt_rms.dat nomod
filter R1 min 0.0 max 2.0 out filt.dat
""", [("filt_rms.dat", "table"), ("filt.dat", "table")]),
    # src/Analysis_Rms2d.cpp:289-293: rms2d with the pseudo-autocorrelation tail
    "rms2d_corr": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd 1 40
2drms crd1 :3-7 rmsout rms2d.dat corr corr.dat
""", [("rms2d.dat", "table"), ("corr.dat", "table")]),
    # test/Test_Cluster_Kmeans/RunTest.sh:26-34: k-means (centroid updates Metric_RMS::FrameOpCentroid on the CPU, seed
    # search, final centroids, best representatives and the cluster summary through the B200 path)
    "cluster_kmeans": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
cluster means clusters 5 rms @CA summary summary.dat info info.dat out cnumvtime.dat
""", [("summary.dat", "text"), ("info.dat", "text"), ("cnumvtime.dat", "table")]),
    # hierarchical agglomerative, fitted, with sieve restore by epsilon (second List::AddFramesByCentroid overload,
    # src/Cluster/List.cpp:215-298) and centroid best representatives
    "cluster_hier_sieve_eps": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
cluster crd1 @CA hieragglo epsilon 2.5 averagelinkage rms out hs.out summary hs.summary.dat sieve 4 sievetoframe repsilon 2.0 bestrep centroid savenreps 2
""", [("hs.out", "table"), ("hs.summary.dat", "text")]),
    # src/Cluster/Algorithm_HierAgglo.cpp:97-245: the whole merge loop on the device (b200_hieragglo) for the three
    # linkages, stopping on the cluster count, on epsilon, and on whichever comes first; the epsilon-vs-clusters file
    # lists every FindMin value in order, the info file every cluster's frames; the summaries carry what the post-processing
    # reads from the cache: best representatives by cumulative distance (with and without sieved frames), within-cluster
    # averages and standard deviations, average linkage to the other clusters
    "cluster_hier_linkages": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
cluster L1 @CA hieragglo clusters 4 linkage rms out hl.single.dat info hl.single.info epsilonplot hl.single.eps summary hl.single.summary savenreps 3
cluster L2 @CA hieragglo epsilon 1.8 averagelinkage rms out hl.avg.dat summary hl.avg.summary epsilonplot hl.avg.eps
cluster L3 :2-12 hieragglo clusters 6 epsilon 3.0 complete rms mass out hl.complete.dat info hl.complete.info epsilonplot hl.complete.eps summary hl.complete.summary
cluster L4 @CA hieragglo clusters 3 averagelinkage rms nofit sieve 3 out hl.sieve.dat info hl.sieve.info summary hl.sieve.summary
""", [("hl.single.dat", "table"), ("hl.single.info", "text"), ("hl.single.eps", "table"), ("hl.single.summary", "text"),
      ("hl.avg.dat", "table"), ("hl.avg.summary", "text"), ("hl.avg.eps", "table"), ("hl.complete.dat", "table"),
      ("hl.complete.info", "text"), ("hl.complete.eps", "table"), ("hl.complete.summary", "text"), ("hl.sieve.dat", "table"),
      ("hl.sieve.info", "text"), ("hl.sieve.summary", "text")]),
    # src/Cluster/Cmatrix_Binary.cpp:12-21: the cache filled on the device written by the reference's own binary writer
    # (savepairdist), read back (readdata) and clustered on the device again from the loaded cache
    "cluster_cmatrix_roundtrip": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
cluster P1 @CA hieragglo clusters 5 averagelinkage rms savepairdist pairdist cm.bin out p1.dat
run
readdata cm.bin name CM2
cluster P2 @CA hieragglo clusters 4 complete rms pairdist CM2 out p2.dat info p2.info
""", [("cm.bin", "cmatrix"), ("p1.dat", "table"), ("p2.dat", "table"), ("p2.info", "text")]),
    # test/Test_RmsAvgCorr/RunTest.sh:11-34 (fixed reference; offset; first running-averaged frame as reference) plus
    # mass weighting with a window limit: every window size on the device (b200_rmsavgcorr)
    "rmsavgcorr": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
reference {D}/tz2.crd 5
rmsavgcorr RF :2-12@CA out rac.first.dat first
rmsavgcorr RM :2-12@CA,C,N out rac.mass.dat first mass stop 40 offset 3
rmsavgcorr RR :2-12@CA out rac.ref.dat reference
rmsavgcorr R10 :2-12@CA out rac.ref10.dat reference offset 10
""", [("rac.first.dat", "table"), ("rac.mass.dat", "table"), ("rac.ref.dat", "table"), ("rac.ref10.dat", "table")]),
    # src/Cluster/Results_Coords.cpp:331-377: reference structures assigned to clusters by the RMSD of the best
    # representative to every reference (nearest reference, cut-off naming), one frames x references call
    "cluster_assignrefs": ("""noprogress
parm {D}/tz2.parm7
trajin {D}/tz2.crd
reference {D}/tz2.crd 1 [first]
reference {D}/tz2.crd 50 [mid]
reference {D}/tz2.crd 100 [late]
cluster C1 @CA clusters 5 rms out ar.out summary ar.summary.dat assignrefs refcut 2.0 refmask @CA
cluster C2 :2-12 clusters 4 rms mass summary ar.mass.summary.dat assignrefs refcut 1.0 refmask :2-12@CA,C,N
""", [("ar.out", "table"), ("ar.summary.dat", "text"), ("ar.mass.summary.dat", "text")]),
    # src/Analysis_Rms2d.cpp:92-110,247-287 with sets that are NOT in-memory COORDS: a TRAJ set (loadtraj: frames stay on disk),
    # a reference trajectory given as a file name (test/Test_2DRMS/RunTest.sh:33-46, ASCII instead of NetCDF), different
    # target / reference masks inside a TRAJ set, an in-memory target against a TRAJ reference.  (Golden made with ONE
    # thread: the reference reads a TRAJ target from inside its OpenMP loop.)
    "rms2d_traj_sets": ("""noprogress
parm {D}/tz2.parm7
loadtraj name TZ2 {D}/tz2.crd 1 48
2drms crdset TZ2 :3-7 out traj_tri.dat
2drms crdset TZ2 :2-12@CA,C,N mass out traj_mass.dat
2drms crdset TZ2 :3-7 out traj_reftraj.dat reftraj {D}/tz2.crd
2drms crdset TZ2 :2 :11 nofit out traj_nofit_masks.dat
loadcrd {D}/tz2.crd 3 60 2 name CRD
2drms crdset CRD :3-7 out crd_reftraj.dat reftraj TZ2
precision traj_tri.dat 12 6
precision traj_mass.dat 12 6
precision traj_reftraj.dat 12 6
precision traj_nofit_masks.dat 12 6
precision crd_reftraj.dat 12 6
""", [("traj_tri.dat", "table"), ("traj_mass.dat", "table"), ("traj_reftraj.dat", "table"), ("traj_nofit_masks.dat", "table"),
      ("crd_reftraj.dat", "table")]),
    # src/Cluster/MetricArray.cpp:766-801 on a TRAJ set (frames on disk): the cache fill reads the selected atoms of the frames
    # to cache once and runs on the device, the merge loop and the cache consumers too; centroids (single-frame reads) stay
    # on the reference's code.  (Golden made with ONE thread.)
    "cluster_traj_set": ("""noprogress
parm {D}/tz2.parm7
loadtraj name TZ2 {D}/tz2.crd
cluster T1 crdset TZ2 @CA hieragglo clusters 5 averagelinkage rms out ct.dat summary ct.summary.dat info ct.info
cluster T2 crdset TZ2 :2-12 hieragglo clusters 4 complete rms mass sieve 3 out ct2.dat summary ct2.summary.dat
""", [("ct.dat", "table"), ("ct.summary.dat", "text"), ("ct.info", "text"), ("ct2.dat", "table"), ("ct2.summary.dat", "text")]),
    # src/Analysis_RmsAvgCorr.cpp:176-316 on a TRAJ set (frames on disk; the reference reads every frame again for every
    # window size): selected atoms read once, every window size on the device.  (Golden made with ONE thread.)
    "rmsavgcorr_traj_set": ("""noprogress
parm {D}/tz2.parm7
loadtraj name TZ2 {D}/tz2.crd
reference {D}/tz2.crd 5
rmsavgcorr TF crdset TZ2 :2-12@CA out tac.first.dat first
rmsavgcorr TR crdset TZ2 :2-12@CA,C,N out tac.ref.dat reference mass offset 4
""", [("tac.first.dat", "table"), ("tac.ref.dat", "table")]),
}
