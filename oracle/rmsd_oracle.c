/* rmsd_oracle.c -- CPU restatement of cpptraj's best-fit RMSD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the timed CPU baseline; the product (cpptraj_b200/csrc) never calls it and
 * fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the reference's own golden vectors (test/Test_2DRMS/{rmsd,rmsd.mass,
 *       trp,nofit}.dat.save, test/Test_RMSD/NoMod.dat.save; copies of the
 *       numbers under tests/golden/, generator tools/make_golden.py), and
 *   (2) the reference's real Frame/Matrix_3x3/CompactFrameArray/Matrix<float>
 *       code compiled into oracle/_ref/ (oracle/build_ref.sh), to <= 1e-12.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it follows.  All arithmetic is double, inputs float32 AoS,
 * outputs float32 (rms2d / pairwise cache) or double (rmsd action) exactly as
 * in the reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SMALL 0.00000000000001 /* src/Constants.h:35 (Constants::SMALL) */
#define ORC_JACOBI_SWEEPS 50       /* src/Matrix_3x3.cpp:97 */

/* ---- a6: mask gather float32 -> double ----------------------------------
 * src/CompactFrameArray.cpp:244-262 (GetToMaskDblPtr): position component at
 * offset 0 of each frame, frame stride = whole frame incl. vel/box/etc. */
static void orc_gather(const float *crd, size_t stride, long frame,
                       const int *sel, int n, double *X)
{
    const float *fb = crd + (size_t)frame * stride;
    for (int j = 0; j < n; j++) {
        int s = 3 * sel[j];
        X[3 * j + 0] = (double)fb[s + 0];
        X[3 * j + 1] = (double)fb[s + 1];
        X[3 * j + 2] = (double)fb[s + 2];
    }
}

/* ---- a3: centre on origin ------------------------------------------------
 * src/Frame.cpp:1043-1055 (CenterOnOrigin), src/Frame.h:373-391 (VCenterOfMass)
 * and :418-431 (VGeometricCenter), :500-506 (NegTranslate).
 * Returns centre in c[3]. */
static void orc_center_on_origin(double *X, int n, const double *mass,
                                 int useMass, double *c)
{
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, sm = 0.0;
    if (useMass) {
        for (int i = 0; i < n; i++) {
            sm += mass[i];
            s0 += X[3 * i + 0] * mass[i];
            s1 += X[3 * i + 1] * mass[i];
            s2 += X[3 * i + 2] * mass[i];
        }
    } else {
        for (int i = 0; i < n; i++) {
            s0 += X[3 * i + 0];
            s1 += X[3 * i + 1];
            s2 += X[3 * i + 2];
        }
        sm = (double)n;
    }
    if (sm == 0.0) {
        c[0] = c[1] = c[2] = 0.0;
    } else {
        c[0] = s0 / sm; c[1] = s1 / sm; c[2] = s2 / sm;
    }
    for (int i = 0; i < n; i++) {
        X[3 * i + 0] -= c[0];
        X[3 * i + 1] -= c[1];
        X[3 * i + 2] -= c[2];
    }
}

/* ---- a2: cyclic Jacobi on a symmetric 3x3 ----------------------------------
 * src/Matrix_3x3.cpp:110-211 (Diagonalize).  a[9] row-major symmetric input
 * (destroyed); V[9] receives eigenvectors in columns; d[3] eigenvalues.
 * Returns 1 when the sweep limit is hit (reference then returns RMSD 0). */
static void orc_jrot(double *m, int i1, int i2, double s, double tau)
{
    double g = m[i1], h = m[i2];
    m[i1] = g - s * (h + g * tau);
    m[i2] = h + s * (g - h * tau);
}

static int orc_jacobi3(double *a, double *V, double *d)
{
    double b[3], z[3];
    for (int i = 0; i < 9; i++) V[i] = 0.0;
    V[0] = V[4] = V[8] = 1.0;
    for (int i = 0; i < 3; i++) { b[i] = d[i] = a[4 * i]; z[i] = 0.0; }
    for (int sweep = 0; sweep < ORC_JACOBI_SWEEPS; sweep++) {
        double sm = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
        if (sm == 0.0) return 0;
        double tresh = (sweep < 3) ? 0.2 * sm / 9 : 0.0;
        for (int ip = 0; ip < 2; ip++) {
            for (int iq = ip + 1; iq < 3; iq++) {
                int pq = 3 * ip + iq;
                double g = 100.0 * fabs(a[pq]);
                if (sweep > 3 && fabs(d[ip]) + g == fabs(d[ip]) &&
                                 fabs(d[iq]) + g == fabs(d[iq])) {
                    a[pq] = 0.0;
                } else if (fabs(a[pq]) > tresh) {
                    double h = d[iq] - d[ip], t;
                    if (fabs(h) + g == fabs(h)) {
                        t = a[pq] / h;
                    } else {
                        double theta = 0.5 * h / a[pq];
                        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                        if (theta < 0.0) t = -t;
                    }
                    double c = 1.0 / sqrt(1 + t * t);
                    double s = t * c;
                    double tau = s / (1.0 + c);
                    h = t * a[pq];
                    z[ip] -= h; z[iq] += h;
                    d[ip] -= h; d[iq] += h;
                    a[pq] = 0.0;
                    for (int j = 0; j <= ip - 1; j++)  orc_jrot(a, 3 * j + ip, 3 * j + iq, s, tau);
                    for (int j = ip + 1; j <= iq - 1; j++) orc_jrot(a, 3 * ip + j, 3 * j + iq, s, tau);
                    for (int j = iq + 1; j < 3; j++)   orc_jrot(a, 3 * ip + j, 3 * iq + j, s, tau);
                    for (int j = 0; j < 3; j++)        orc_jrot(V, 3 * j + ip, 3 * j + iq, s, tau);
                }
            }
        }
        for (int i = 0; i < 3; i++) { b[i] += z[i]; d[i] = b[i]; z[i] = 0.0; }
    }
    return 1;
}

/* src/Matrix_3x3.cpp:219-268 (Diagonalize_Sort): descending order, vectors in
 * rows of E.  Order selection mirrors the reference's strict comparisons. */
static int orc_diag_sort(double *a, double *E, double *ev)
{
    double V[9], d[3];
    if (orc_jacobi3(a, V, d)) return 1;
    int i1, i2, i3;
    if (d[0] > d[1] && d[0] > d[2]) {
        i1 = 0; if (d[1] > d[2]) { i2 = 1; i3 = 2; } else { i2 = 2; i3 = 1; }
    } else if (d[1] > d[0] && d[1] > d[2]) {
        i1 = 1; if (d[0] > d[2]) { i2 = 0; i3 = 2; } else { i2 = 2; i3 = 0; }
    } else if (d[0] > d[1]) {
        i1 = 2; i2 = 0; i3 = 1;
    } else {
        i1 = 2; i2 = 1; i3 = 0;
    }
    for (int c = 0; c < 3; c++) {
        E[0 + c] = V[i1 + 3 * c];
        E[3 + c] = V[i2 + 3 * c];
        E[6 + c] = V[i3 + 3 * c];
    }
    ev[0] = d[i1]; ev[1] = d[i2]; ev[2] = d[i3];
    return 0;
}

static void orc_unit3(double *v)
{
    double b = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= b; v[1] *= b; v[2] *= b;
}

/* ---- a1: RMSD of target T to a pre-centred reference R --------------------
 * src/Frame.cpp:1137-1273 (RMSD_CenteredRef).  T is translated in place.
 * U[9] (rotation) and trans[3] (target -> origin translation) may be NULL.
 * mass is the TARGET frame's mass array (src/Frame.cpp:1184). */
double orc_rmsd_centered_ref(double *T, const double *R, int n,
                             const double *mass, int useMass,
                             double *U, double *trans)
{
    double tm, tr[3] = {0, 0, 0};
    if (useMass) {
        tm = 0.0;
        for (int i = 0; i < n; i++) {
            tm += mass[i];
            tr[0] += T[3 * i + 0] * mass[i];
            tr[1] += T[3 * i + 1] * mass[i];
            tr[2] += T[3 * i + 2] * mass[i];
        }
    } else {
        tm = (double)n;
        for (int i = 0; i < n; i++) {
            tr[0] += T[3 * i + 0]; tr[1] += T[3 * i + 1]; tr[2] += T[3 * i + 2];
        }
    }
    if (tm < ORC_SMALL) return -1.0;            /* :1160-1163 */
    tr[0] /= tm; tr[1] /= tm; tr[2] /= tm;
    tr[0] = -tr[0]; tr[1] = -tr[1]; tr[2] = -tr[2];
    for (int i = 0; i < n; i++) {               /* Translate(Trans) */
        T[3 * i + 0] += tr[0]; T[3 * i + 1] += tr[1]; T[3 * i + 2] += tr[2];
    }
    if (trans) { trans[0] = tr[0]; trans[1] = tr[1]; trans[2] = tr[2]; }

    double mwss = 0.0, rot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, am = 1.0;
    for (int i = 0; i < n; i++) {               /* :1184-1208 */
        double xt = T[3 * i], yt = T[3 * i + 1], zt = T[3 * i + 2];
        double xr = R[3 * i], yr = R[3 * i + 1], zr = R[3 * i + 2];
        if (useMass) am = mass[i];
        mwss += am * ((xt * xt) + (yt * yt) + (zt * zt) + (xr * xr) + (yr * yr) + (zr * zr));
        rot[0] += am * xt * xr; rot[1] += am * xt * yr; rot[2] += am * xt * zr;
        rot[3] += am * yt * xr; rot[4] += am * yt * yr; rot[5] += am * yt * zr;
        rot[6] += am * zt * xr; rot[7] += am * zt * yr; rot[8] += am * zt * zr;
    }
    mwss *= 0.5;
    /* rot * rot^T (src/Matrix_3x3.cpp:421-433 TransposeMult) */
    double rr[9], E[9], ev[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            rr[3 * i + j] = rot[3 * i] * rot[3 * j] + rot[3 * i + 1] * rot[3 * j + 1] +
                            rot[3 * i + 2] * rot[3 * j + 2];
    if (orc_diag_sort(rr, E, ev)) return 0.0;   /* :1218 */
    /* a3 = a1 x a2 */
    E[6] = E[1] * E[5] - E[2] * E[4];
    E[7] = E[2] * E[3] - E[0] * E[5];
    E[8] = E[0] * E[4] - E[1] * E[3];
    double b[9];
    for (int k = 0; k < 3; k++) {               /* b_k = R . a_k, normalised */
        b[3 * k + 0] = E[3 * k] * rot[0] + E[3 * k + 1] * rot[3] + E[3 * k + 2] * rot[6];
        b[3 * k + 1] = E[3 * k] * rot[1] + E[3 * k + 1] * rot[4] + E[3 * k + 2] * rot[7];
        b[3 * k + 2] = E[3 * k] * rot[2] + E[3 * k + 1] * rot[5] + E[3 * k + 2] * rot[8];
        orc_unit3(b + 3 * k);
    }
    double cp[3];
    cp[0] = b[1] * b[5] - b[2] * b[4];
    cp[1] = b[2] * b[3] - b[0] * b[5];
    cp[2] = b[0] * b[4] - b[1] * b[3];
    double sig3 = ((cp[0] * b[6] + cp[1] * b[7] + cp[2] * b[8]) < 0.0) ? -1.0 : 1.0;
    b[6] = cp[0]; b[7] = cp[1]; b[8] = cp[2];
    if (U) {                                    /* :1246-1256 */
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                U[3 * r + c] = E[c] * b[r] + E[3 + c] * b[3 + r] + E[6 + c] * b[6 + r];
    }
    double e = mwss - sqrt(fabs(ev[0])) - sqrt(fabs(ev[1])) - (sig3 * sqrt(fabs(ev[2])));
    if (e < 0) return 0.0;                      /* :1264-1266 */
    return sqrt((2.0 * e) / tm);
}

/* ---- a4: no-fit RMSD  (src/Frame.cpp:1279-1307) -------------------------- */
double orc_rmsd_nofit(const double *T, const double *R, int n,
                      const double *mass, int useMass)
{
    double acc = 0.0, tm = 0.0, am = 1.0;
    for (int i = 0; i < n; i++) {
        double xx = R[3 * i] - T[3 * i], yy = R[3 * i + 1] - T[3 * i + 1], zz = R[3 * i + 2] - T[3 * i + 2];
        if (useMass) am = mass[i];
        tm += am;
        acc += am * (xx * xx + yy * yy + zz * zz);
    }
    if (tm < ORC_SMALL) return -1.0;
    if (acc < 0) return 0.0;
    return sqrt(acc / tm);
}

/* ---- a11: output indexers  (src/Matrix.h:94-96,110-122) ------------------- */
size_t orc_tri_index(size_t n, size_t x, size_t y)
{
    size_t i = x < y ? x : y, j = x < y ? y : x;
    size_t i1 = i + 1;
    return ((n * i) - ((i1 * i) / 2)) + j - i1;
}
size_t orc_full_index(size_t ncols, size_t x, size_t y) { return y * ncols + x; }

/* ---- a7: rms2d pair enumeration -------------------------------------------
 * src/Analysis_Rms2d.cpp:196-295 (Calculate_2D), triangle branch: reference =
 * frame nref (centred once per row, :263-266), target = ntgt > nref (:272-283).
 * frameIdx (nullable) selects/permutes frames like the cluster path's
 * framesToCache.  mass NULL => unit masses (useMass false).
 * out: float[nF(nF-1)/2] in calcTriIndex order.  Returns 0 / nonzero. */
int orc_rms2d_tri(const float *crd, size_t stride, const int *frameIdx, int nF,
                  const int *sel, int n, const double *mass, int fit, float *out)
{
    int useMass = (mass != NULL);
    int err = 0;
#pragma omp parallel
    {
        double *R = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *T = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double c[3];
        if (!R || !T) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic)
            for (int r = 0; r < nF; r++) {
                orc_gather(crd, stride, frameIdx ? frameIdx[r] : r, sel, n, R);
                if (fit) orc_center_on_origin(R, n, mass, useMass, c);
                for (int t = r + 1; t < nF; t++) {
                    orc_gather(crd, stride, frameIdx ? frameIdx[t] : t, sel, n, T);
                    double v = fit ? orc_rmsd_centered_ref(T, R, n, mass, useMass, NULL, NULL)
                                   : orc_rmsd_nofit(T, R, n, mass, useMass);
                    out[orc_tri_index((size_t)nF, (size_t)r, (size_t)t)] = (float)v;
                }
            }
        }
        free(R); free(T);
    }
    return err;
}

/* Cluster flavour: src/Cluster/Metric_RMS.cpp:44-63 (FrameDist) called from
 * src/Cluster/MetricArray.cpp:766-801: frm1 = frame f1 is "this" (target),
 * frm2 = frame f2 is re-centred per pair (src/Frame.cpp:1102-1109), result is
 * a double stored to a float cache (src/DataSet_PairwiseCache_MEM.h:28). */
int orc_cluster_tri(const float *crd, size_t stride, const int *frameIdx, int nF,
                    const int *sel, int n, const double *mass, int fit, float *out)
{
    int useMass = (mass != NULL);
    int err = 0;
#pragma omp parallel
    {
        double *A = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *B = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double c[3];
        if (!A || !B) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic)
            for (int f1 = 0; f1 < nF - 1; f1++) {
                for (int f2 = f1 + 1; f2 < nF; f2++) {
                    orc_gather(crd, stride, frameIdx ? frameIdx[f1] : f1, sel, n, A);
                    orc_gather(crd, stride, frameIdx ? frameIdx[f2] : f2, sel, n, B);
                    double v;
                    if (fit) {
                        orc_center_on_origin(B, n, mass, useMass, c);
                        v = orc_rmsd_centered_ref(A, B, n, mass, useMass, NULL, NULL);
                    } else
                        v = orc_rmsd_nofit(A, B, n, mass, useMass);
                    out[orc_tri_index((size_t)nF, (size_t)f1, (size_t)f2)] = (float)v;
                }
            }
        }
        free(A); free(B);
    }
    return err;
}

/* Full-matrix branch of Calculate_2D (reftraj or differing masks):
 * Allocate2D(totalref,totaltgt) => ncols = nRef, element (x=nref,y=ntgt) at
 * ntgt*nRef + nref  (src/Analysis_Rms2d.cpp:208-209, src/Matrix.h:94-96).
 * massTgt weights the covariance and total mass (target frame's Mass_);
 * massRef is used only to centre the reference (SelectedRef.CenterOnOrigin). */
int orc_rms2d_full(const float *crdT, size_t strideT, int nT, const int *selT,
                   const float *crdR, size_t strideR, int nR, const int *selR,
                   int n, const double *massTgt, const double *massRef, int fit,
                   float *out)
{
    int useMass = (massTgt != NULL);
    int err = 0;
#pragma omp parallel
    {
        double *R = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *T = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double c[3];
        if (!R || !T) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic)
            for (int r = 0; r < nR; r++) {
                orc_gather(crdR, strideR, r, selR, n, R);
                if (fit) orc_center_on_origin(R, n, massRef, useMass, c);
                for (int t = 0; t < nT; t++) {
                    orc_gather(crdT, strideT, t, selT, n, T);
                    double v = fit ? orc_rmsd_centered_ref(T, R, n, massTgt, useMass, NULL, NULL)
                                   : orc_rmsd_nofit(T, R, n, massTgt, useMass);
                    out[orc_full_index((size_t)nR, (size_t)r, (size_t)t)] = (float)v;
                }
            }
        }
        free(R); free(T);
    }
    return err;
}

/* ---- a10: one-vs-many (rmsd action) ----------------------------------------
 * src/Action_Rmsd.cpp:361-392 (DoAction) with a fixed reference
 * (ReferenceAction FIRST/FRAME modes, src/ReferenceAction.cpp:155-169):
 * ref = selected atoms, centred once when fitting (refTrans returned).
 * Input frames are double (cpptraj Frame) or float (COORDS set); both offered.
 * rmsdOut[nF] double; rotOut (nullable) 9 per frame; transOut (nullable)
 * 3 per frame = target -> origin translation (tgtTrans_). */
int orc_rmsd_1vN(const float *crd, size_t stride, int nF, const int *sel, int n,
                 const double *refSel /*3n, NOT yet centred*/, const double *mass,
                 int fit, double *rmsdOut, double *rotOut, double *transOut,
                 double *refTrans /*3, nullable*/)
{
    int useMass = (mass != NULL);
    double *R = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    if (!R) return 1;
    memcpy(R, refSel, sizeof(double) * 3 * (size_t)n);
    double c[3] = {0, 0, 0};
    if (fit) orc_center_on_origin(R, n, mass, useMass, c);
    if (refTrans) { refTrans[0] = c[0]; refTrans[1] = c[1]; refTrans[2] = c[2]; }
    int err = 0;
#pragma omp parallel
    {
        double *T = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        if (!T) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(static)
            for (int f = 0; f < nF; f++) {
                orc_gather(crd, stride, f, sel, n, T);
                if (fit)
                    rmsdOut[f] = orc_rmsd_centered_ref(T, R, n, mass, useMass,
                                                       rotOut ? rotOut + 9 * (size_t)f : NULL,
                                                       transOut ? transOut + 3 * (size_t)f : NULL);
                else
                    rmsdOut[f] = orc_rmsd_nofit(T, R, n, mass, useMass);
            }
        }
        free(T);
    }
    free(R);
    return err;
}

/* ---- f2: centroid building with fit ------------------------------------------
 * src/Cluster/Metric_RMS.cpp:86-113 (Metric_RMS::CalculateCentroid): the first
 * frame of the cluster is the start of a running SUM (centred when fitting);
 * every further frame is fitted to that sum (RMSD_CenteredRef translates the
 * frame to the origin and returns the rotation, :101), rotated (Frame::Rotate,
 * src/Frame.h:508-517: x' = U x) and added; finally divided by the frame count.
 * frames[nIn]: frame numbers of the cluster; cent: 3n doubles out. */
int orc_build_centroid(const float *crd, size_t stride, const int *frames, int nIn,
                       const int *sel, int n, const double *mass, int fit, double *cent)
{
    int useMass = (mass != NULL);
    if (nIn < 1) return 0;
    double *T = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    if (!T) return 1;
    for (int j = 0; j < nIn; j++) {
        orc_gather(crd, stride, frames[j], sel, n, T);
        if (j == 0) {
            memcpy(cent, T, sizeof(double) * 3 * (size_t)n);
            if (fit) { double c[3]; orc_center_on_origin(cent, n, mass, useMass, c); }   /* :96-97 */
        } else {
            if (fit) {
                double U[9], tr[3];
                orc_rmsd_centered_ref(T, cent, n, mass, useMass, U, tr);                   /* :100 */
                for (int i = 0; i < n; i++) {                                              /* :102, Frame.h:508-517 */
                    double x = T[3 * i], y = T[3 * i + 1], z = T[3 * i + 2];
                    T[3 * i]     = x * U[0] + y * U[1] + z * U[2];
                    T[3 * i + 1] = x * U[3] + y * U[4] + z * U[5];
                    T[3 * i + 2] = x * U[6] + y * U[7] + z * U[8];
                }
            }
            for (int i = 0; i < 3 * n; i++) cent[i] += T[i];                               /* :104 */
        }
    }
    for (int i = 0; i < 3 * n; i++) cent[i] /= (double)nIn;                                /* :108 */
    free(T);
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- (f) rank 3: hierarchical agglomerative clustering on the pairwise cache --------
 * Restates Algorithm_HierAgglo::DoClustering / MergeClosest
 * (src/Cluster/Algorithm_HierAgglo.cpp:97-245) with its frame-pair linkage loops
 * minDist / maxDist / avgDist (:248-298) and calc{Min,Max,Avg}Dist (:303-350), and the
 * DynamicMatrix bookkeeping they go through (src/Cluster/DynamicMatrix.h:43-126,
 * src/Cluster/DynamicMatrix.cpp:7-33): float cluster-distance triangle, ignore flags
 * and the per-cluster "closest" index with its exact update rules (strict '<',
 * rescan when the closest distance grew, rescan when the closest was ignored).
 * tri = the pairwise cache (DataSet_PairwiseCache_MEM::Mat_, src/Matrix.h:110-122),
 * one initial cluster per cached frame, Num() = cache index
 * (buildInitialClusters, :86-95).  linkage: 0 single, 1 average, 2 complete
 * (LINKAGETYPE, src/Cluster/Algorithm_HierAgglo.h:35).  targetClusters / epsilon
 * are nclusters_ / epsilon_ after DoClustering's defaults (-1 -> 1, -1.0 -> DBL_MAX).
 * Outputs: mergeInto[m] (C1, the lower Num), mergeFrom[m] (C2), findMin[m] = the
 * value FindMin returned in MergeClosest call m; *nCalls = number of MergeClosest
 * calls, *nMerges = merges performed (nCalls-1 when the last call stopped on epsilon). */
typedef struct {
    int n;
    float *mat;
    unsigned char *ignore;
    int *closest;
} orc_dynmat;

static size_t orc_tri_idx(int n, int a, int b)
{   /* src/Matrix.h:110-122 calcTriIndex */
    if (a > b) { int t = a; a = b; b = t; }
    return (size_t)n * (size_t)a - ((size_t)a * ((size_t)a + 1)) / 2 + (size_t)b - (size_t)a - 1;
}
static float orc_dm_get(const orc_dynmat *M, int a, int b) { return M->mat[orc_tri_idx(M->n, a, b)]; }

/* DynamicMatrix::updateClosestIdx, src/Cluster/DynamicMatrix.h:43-62 */
static void orc_dm_update_closest(orc_dynmat *M, int idx)
{
    M->closest[idx] = -1;
    float cur = 0;
    for (int jdx = 0; jdx != M->n; jdx++) {
        if (!M->ignore[jdx] && idx != jdx) {
            if (M->closest[idx] == -1) {
                M->closest[idx] = jdx;
                cur = orc_dm_get(M, idx, jdx);
            } else {
                float fdist = orc_dm_get(M, idx, jdx);
                if (fdist < cur) { M->closest[idx] = jdx; cur = fdist; }
            }
        }
    }
}
/* DynamicMatrix::SetCdist, src/Cluster/DynamicMatrix.h:65-113 */
static void orc_dm_set(orc_dynmat *M, int col, int row, float val)
{
    int update_col = -1, update_row = -1;
    if (M->closest[col] < 0) M->closest[col] = row;
    else {
        float c = orc_dm_get(M, col, M->closest[col]);
        if (val < c) M->closest[col] = row;
        else if (row == M->closest[col] && val > c) update_col = col;
    }
    if (M->closest[row] < 0) M->closest[row] = col;
    else {
        float c = orc_dm_get(M, row, M->closest[row]);
        if (val < c) M->closest[row] = col;
        else if (col == M->closest[row] && val > c) update_row = row;
    }
    M->mat[orc_tri_idx(M->n, col, row)] = val;
    if (update_col != -1) orc_dm_update_closest(M, update_col);
    if (update_row != -1) orc_dm_update_closest(M, update_row);
}
/* DynamicMatrix::Ignore, src/Cluster/DynamicMatrix.h:116-126 */
static void orc_dm_ignore(orc_dynmat *M, int row)
{
    M->ignore[row] = 1;
    for (int idx = 0; idx != M->n; idx++)
        if (!M->ignore[idx] && M->closest[idx] == row) orc_dm_update_closest(M, idx);
}
/* DynamicMatrix::FindMin, src/Cluster/DynamicMatrix.cpp:7-33 */
static double orc_dm_findmin(const orc_dynmat *M, int *iOut, int *jOut)
{
    float currentMin = 3.402823466e+38F;
    int minRow = -1, minCol = -1;
    for (int col = 0; col != M->n; col++) {
        int row = M->closest[col];
        if (!M->ignore[col] && row >= 0 && !M->ignore[row]) {
            float mval = orc_dm_get(M, col, row);
            if (mval < currentMin) { currentMin = mval; minRow = row; minCol = col; }
        }
    }
    if (minRow < minCol) { *iOut = minRow; *jOut = minCol; } else { *iOut = minCol; *jOut = minRow; }
    return (double)currentMin;
}
/* minDist / maxDist / avgDist, src/Cluster/Algorithm_HierAgglo.cpp:248-298: frame-pair loops over the cache
 * (MetricArray::Frame_Distance -> CachedDistance, src/Cluster/MetricArray.cpp:599-609), C1 frames outer. */
static double orc_linkage(const float *tri, int n, int linkage, const int *f1, int n1, const int *f2, int n2)
{
    double acc = (linkage == 0) ? 1.7976931348623157e308 : (linkage == 2 ? -1.0 : 0.0);
    for (int a = 0; a < n1; a++)
        for (int b = 0; b < n2; b++) {
            double Dist = (double)tri[orc_tri_idx(n, f1[a], f2[b])];
            if (linkage == 0) { if (Dist < acc) acc = Dist; }
            else if (linkage == 2) { if (Dist > acc) acc = Dist; }
            else acc += Dist;
        }
    if (linkage == 1) return acc / (double)(n1 * n2);
    return acc;
}

int orc_hieragglo(const float *tri, int n, int linkage, int targetClusters, double epsilon,
                  int *mergeInto, int *mergeFrom, float *findMin, int *nCalls, int *nMerges)
{
    *nCalls = 0; *nMerges = 0;
    if (n < 2) return 0;
    orc_dynmat M;
    M.n = n;
    size_t nElt = (size_t)n * (size_t)(n - 1) / 2;
    M.mat = (float *)calloc(nElt, sizeof(float));
    M.ignore = (unsigned char *)calloc((size_t)n, 1);
    M.closest = (int *)malloc(sizeof(int) * (size_t)n);
    /* clusters: frame lists (Cframes::Insert appends, src/Cluster/Cframes.h:36) */
    int **fl = (int **)malloc(sizeof(int *) * (size_t)n);
    int *nf = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
        M.closest[i] = -1;
        fl[i] = (int *)malloc(sizeof(int)); fl[i][0] = i; nf[i] = 1;
    }
    /* initial cluster distance matrix, :119-137 */
    for (int c1 = 0; c1 < n; c1++)
        for (int c2 = c1 + 1; c2 < n; c2++)
            orc_dm_set(&M, c1, c2, (float)orc_linkage(tri, n, linkage, fl[c1], nf[c1], fl[c2], nf[c2]));
    int nClusters = n;
    for (;;) {
        /* MergeClosest, :170-245 */
        int C1, C2;
        double min = orc_dm_findmin(&M, &C1, &C2);
        findMin[*nCalls] = (float)min;
        (*nCalls)++;
        if (min > epsilon) break;
        mergeInto[*nMerges] = C1; mergeFrom[*nMerges] = C2; (*nMerges)++;
        fl[C1] = (int *)realloc(fl[C1], sizeof(int) * (size_t)(nf[C1] + nf[C2]));
        memcpy(fl[C1] + nf[C1], fl[C2], sizeof(int) * (size_t)nf[C2]);
        nf[C1] += nf[C2];
        free(fl[C2]); fl[C2] = 0; nf[C2] = 0;
        nClusters--;
        orc_dm_ignore(&M, C2);
        for (int k = 0; k < n; k++)   /* cluster list order = ascending Num */
            if (fl[k] != 0 && k != C1)
                orc_dm_set(&M, C1, k, (float)orc_linkage(tri, n, linkage, fl[C1], nf[C1], fl[k], nf[k]));
        if (nClusters <= targetClusters) break;
        if (nClusters == 1) break;
    }
    for (int i = 0; i < n; i++) free(fl[i]);
    free(fl); free(nf); free(M.mat); free(M.ignore); free(M.closest);
    return 0;
}

/* ---- (f) rank 4: rmsavgcorr -----------------------------------------------------------
 * Analysis_RmsAvgCorr::Analyze, src/Analysis_RmsAvgCorr.cpp:119-316.  Per window size W: a running SUM of the
 * selected coordinates (sumFrame += frame t, :252; -= frame t-W+1 after use, :283-285), the averaged frame
 * tgt = sum / W (Frame::Divide, src/Frame.cpp:950-963), in "first" mode the first averaged frame -- centred -- is the
 * reference of that window size (:258-272), RMSD_CenteredRef of every averaged frame (:273-277), and
 * avg = sum/n, sd = sqrt(max(0, sum2/n - avg^2)) over the n = F - W + 1 of them (:289-297).  Window size 1 is the
 * initial pass over the plain frames (:176-205; in "first" mode against the centred frame 0, :134-141).
 * refSel NULL = "first" mode, else the fixed reference as the caller centred it.  All frames of crd are used. */
int orc_rmsavgcorr(const float *crd, size_t stride, int nF, const int *sel, int n, const double *mass,
                   const double *refSel, const int *windows, int nW, double *avgOut, double *sdOut)
{
    int useMass = (mass != NULL);
    int err = 0;
#pragma omp parallel
    {
        double *T = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *A = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *R = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        double *Sum = (double *)malloc(sizeof(double) * 3 * (size_t)n);
        if (!T || !A || !R || !Sum) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic)
            for (int w = 0; w < nW; w++) {
                int window = windows[w];
                double dW = (double)window, avg = 0.0, sd = 0.0, c[3];
                int first = (refSel == NULL), sub = 0;
                if (refSel) memcpy(R, refSel, sizeof(double) * 3 * (size_t)n);
                for (int i = 0; i < 3 * n; i++) Sum[i] = 0.0;
                for (int f = 0; f < nF; f++) {
                    orc_gather(crd, stride, f, sel, n, T);
                    if (window == 1) {
                        memcpy(A, T, sizeof(double) * 3 * (size_t)n);
                    } else {
                        for (int i = 0; i < 3 * n; i++) Sum[i] += T[i];
                        if (f <= window - 2) continue;
                        for (int i = 0; i < 3 * n; i++) A[i] = Sum[i] / dW;
                    }
                    if (first) {
                        memcpy(R, A, sizeof(double) * 3 * (size_t)n);
                        orc_center_on_origin(R, n, mass, useMass, c);
                        first = 0;
                    }
                    double r = orc_rmsd_centered_ref(A, R, n, mass, useMass, NULL, NULL);
                    avg += r; sd += r * r;
                    if (window != 1) {
                        orc_gather(crd, stride, sub, sel, n, T);
                        for (int i = 0; i < 3 * n; i++) Sum[i] -= T[i];
                        sub++;
                    }
                }
                double d = 1.0 / ((double)nF - dW + 1.0);
                avg *= d; sd *= d; sd -= avg * avg;
                sd = (sd > 0.0) ? sqrt(sd) : 0.0;
                avgOut[w] = avg; sdOut[w] = sd;
            }
        }
        free(T); free(A); free(R); free(Sum);
    }
    return err;
}
