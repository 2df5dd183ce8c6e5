/*
 * region_oracle.c -- CPU restatement of compute_normals / compute_residuals / segmentation_3D of
 * /root/reference/3D_clustering/region_growing.py (`rg`).  TEST INFRASTRUCTURE ONLY (same rule as
 * gsl_oracle.c).
 *
 * Pinned to the reference run verbatim in the build container (oracle/make_golden.py ->
 * tests/golden/region_*.npz).  What is restated exactly: the neighbour set and order
 * (KDTree.query: k nearest by float64 Euclidean distance, nearest first; scipy's squared distance
 * for three dimensions is ((dx*dx) + dy*dy) + dz*dz), the float32 sequential mean of the
 * neighbours in that order (NumPy mean(axis=0) of a float32 [k][3] array, rg:105), the float32
 * centring (rg:108), the orientation rule (rg:120-121), the normalisation (rg:124) and the residual
 * (rg:161).  What is third-party arithmetic and NOT reproducible bit for bit: the float32 sgemm of
 * rg:111 (OpenBLAS blocking) and scipy.linalg.eigh on float32 (LAPACK ssyevr).  Here the
 * covariance of the float32 centred values is accumulated in float64 and diagonalised with cyclic
 * Jacobi rotations in float64; tests/test_region_oracle.py states the resulting tolerance against
 * the golden vectors.  Exact distance ties at the k-th neighbour: lower index first (scipy: tree
 * layout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double d2; int32_t idx; } Cand;

static int cand_less(const Cand *a, const Cand *b)
{
    return a->d2 < b->d2 || (a->d2 == b->d2 && a->idx < b->idx);
}
static int cand_cmp(const void *a, const void *b)
{
    const Cand *x = (const Cand *)a, *y = (const Cand *)b;
    return cand_less(x, y) ? -1 : cand_less(y, x) ? 1 : 0;
}

/* Rearranges c[0..n) so that c[0..k) are the k smallest (quickselect, median of three). */
static void select_k(Cand *c, int64_t n, int64_t k)
{
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        Cand a = c[lo], b = c[mid], d = c[hi], piv;
        if (cand_less(&a, &b)) piv = cand_less(&b, &d) ? b : (cand_less(&a, &d) ? d : a);
        else piv = cand_less(&a, &d) ? a : (cand_less(&b, &d) ? d : b);
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (cand_less(&c[i], &piv)) ++i;
            while (cand_less(&piv, &c[j])) --j;
            if (i <= j) { Cand t = c[i]; c[i] = c[j]; c[j] = t; ++i; --j; }
        }
        if (k - 1 <= j) hi = j;
        else if (k - 1 >= i) lo = i;
        else break;
    }
}

/* Eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix: cyclic Jacobi, float64. */
static void smallest_eigvec(double A[3][3], double v[3])
{
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
    for (int sweep = 0; sweep < 50; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int r = 0; r < 3; ++r) {
            int p = P[r], q = Q[r];
            if (A[p][q] == 0.0) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {
                double akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {
                double apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; ++k) {
                double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    v[0] = V[0][m]; v[1] = V[1][m]; v[2] = V[2][m];
}

/*
 * compute_normals (rg:78-127) and compute_residuals (rg:130-163) for the same k.
 *   pos f32[n][3]; normals_in optional f64[n][3] (the `normals` argument of compute_residuals; NULL =
 *   the normals computed here); outputs optional: normals f64[n][3], residuals f64[n],
 *   centroids f32[n][3] (the reference's float32 centroid), knn int32[n][k] (query order),
 *   gap f64[n] = (second smallest - smallest eigenvalue) / largest: how well the normal is defined.
 */
void orc_region_knn_pca_range(const float *pos, int64_t n, int k, const double *normals_in, double *normals,
                              double *residuals, float *centroids, int32_t *knn, double *gap, int64_t q0, int64_t q1);

void orc_region_knn_pca(const float *pos, int64_t n, int k, const double *normals_in, double *normals,
                        double *residuals, float *centroids, int32_t *knn, double *gap)
{
    orc_region_knn_pca_range(pos, n, k, normals_in, normals, residuals, centroids, knn, gap, 0, n);
}

/* The same for the query points [q0, q1) only (outputs still indexed by point). */
void orc_region_knn_pca_range(const float *pos, int64_t n, int k, const double *normals_in, double *normals,
                              double *residuals, float *centroids, int32_t *knn, double *gap, int64_t q0, int64_t q1)
{
#pragma omp parallel
    {
        Cand *c = (Cand *)malloc((size_t)(n > 0 ? n : 1) * sizeof(Cand));
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = q0; i < q1; ++i) {
            const double qx = pos[3 * i], qy = pos[3 * i + 1], qz = pos[3 * i + 2];
            for (int64_t j = 0; j < n; ++j) {
                double dx = (double)pos[3 * j] - qx, dy = (double)pos[3 * j + 1] - qy, dz = (double)pos[3 * j + 2] - qz;
                c[j].d2 = ((dx * dx) + dy * dy) + dz * dz;
                c[j].idx = (int32_t)j;
            }
            select_k(c, n, k);
            qsort(c, (size_t)k, sizeof(Cand), cand_cmp);                     /* rg:100: nearest first */
            if (knn) for (int j = 0; j < k; ++j) knn[i * k + j] = c[j].idx;
            float sum[3] = {0.f, 0.f, 0.f};                                  /* rg:105 float32 sequential */
            for (int j = 0; j < k; ++j)
                for (int a = 0; a < 3; ++a) sum[a] += pos[3 * (int64_t)c[j].idx + a];
            float cen[3];
            for (int a = 0; a < 3; ++a) cen[a] = (float)((double)sum[a] / (double)k);
            if (centroids) for (int a = 0; a < 3; ++a) centroids[3 * i + a] = cen[a];
            double nrm[3] = {0, 0, 0};
            if (normals || gap || (residuals && !normals_in)) {
                double A[3][3] = {{0}};
                for (int j = 0; j < k; ++j) {
                    float d[3];
                    for (int a = 0; a < 3; ++a) d[a] = pos[3 * (int64_t)c[j].idx + a] - cen[a];   /* rg:108 */
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) A[a][b] += (double)d[a] * (double)d[b];       /* rg:111 */
                }
                double B[3][3];
                memcpy(B, A, sizeof(B));
                smallest_eigvec(B, nrm);                                     /* rg:114-117 */
                if (gap) {
                    double e[3] = {B[0][0], B[1][1], B[2][2]};
                    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) if (e[b] < e[a]) { double t = e[a]; e[a] = e[b]; e[b] = t; }
                    gap[i] = e[2] > 0 ? (e[1] - e[0]) / e[2] : 0.0;
                }
                float pc[3];
                for (int a = 0; a < 3; ++a) pc[a] = pos[3 * i + a] - cen[a];                     /* rg:120 */
                if (nrm[0] * pc[0] + nrm[1] * pc[1] + nrm[2] * pc[2] > 0)
                    for (int a = 0; a < 3; ++a) nrm[a] = -nrm[a];
                double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);          /* rg:124 */
                for (int a = 0; a < 3; ++a) nrm[a] /= len;
                if (normals) for (int a = 0; a < 3; ++a) normals[3 * i + a] = nrm[a];
            }
            if (residuals) {
                const double *nn = normals_in ? normals_in + 3 * i : nrm;
                float pc[3];
                for (int a = 0; a < 3; ++a) pc[a] = pos[3 * i + a] - cen[a];                     /* rg:161 */
                residuals[i] = fabs(nn[0] * pc[0] + nn[1] * pc[1] + nn[2] * pc[2]);
            }
        }
        free(c);
    }
}

/*
 * segmentation_3D (rg:166-221) with the neighbour lists given (kd_tree.query(points[seed], k)[1],
 * rg:203).  region_of[i] = index of the region (in creation order) point i joined.  Returns the
 * number of regions.  min(A, key=residual) (rg:193): lowest residual, lowest index on a tie.
 */
int64_t orc_region_grow(const int32_t *knn, int k, const double *normals, const double *residuals, int64_t n,
                        double residual_threshold, double angle_threshold, int32_t *region_of)
{
    char *avail = (char *)malloc((size_t)(n > 0 ? n : 1));
    int32_t *queue = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    memset(avail, 1, (size_t)n);
    int64_t left = n, regions = 0;
    const double cos_thr = cos(angle_threshold);
    while (left > 0) {
        int64_t pmin = -1;
        for (int64_t i = 0; i < n; ++i)
            if (avail[i] && (pmin < 0 || residuals[i] < residuals[pmin])) pmin = i;
        int64_t head = 0, tail = 0;
        queue[tail++] = (int32_t)pmin;
        avail[pmin] = 0; --left;
        region_of[pmin] = (int32_t)regions;
        while (head < tail) {
            int32_t seed = queue[head++];
            for (int j = 0; j < k; ++j) {
                int32_t nb = knn[(int64_t)seed * k + j];
                if (!avail[nb]) continue;
                double ca = fabs(normals[3 * (int64_t)seed] * normals[3 * (int64_t)nb] + normals[3 * (int64_t)seed + 1] * normals[3 * (int64_t)nb + 1]
                                 + normals[3 * (int64_t)seed + 2] * normals[3 * (int64_t)nb + 2]);
                if (ca > cos_thr) {
                    region_of[nb] = (int32_t)regions;
                    avail[nb] = 0; --left;
                    if (residuals[nb] < residual_threshold) queue[tail++] = nb;
                }
            }
        }
        ++regions;
    }
    free(avail);
    free(queue);
    return regions;
}
