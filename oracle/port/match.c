/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/port/port.h).
 *
 * Restatement of the RANSAC matchers' pre-processing and hypothesis scoring
 * (reference src/obvision/registration/ransacMatching/{RandomMatching,TSD_PDFMatching,
 *  RandomNormalMatching,PDFMatching}.cpp).  PCA goes through the GSL shim exactly as
 * obvious::Matrix::pcaAnalysis does (src/obcore/math/linalg/gsl/Matrix.cpp:227-327).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gsl/gsl_blas.h"
#include "gsl/gsl_linalg.h"
#include "gsl/gsl_statistics_double.h"
#include "port.h"

int port_interpolate_bilinear_one(const port_grid_t* g, const double coord[2], double* tsd);

/* ---------------------------------------------------------------- RNG (same LCG as oracle/ref_capi.cpp) */
static uint32_t g_lcg = 12345u;

void port_seed(uint32_t seed) { g_lcg = seed; }

int port_rand(void)
{
  g_lcg = g_lcg * 1103515245u + 12345u;
  return (int)((g_lcg >> 8) & 0x7fffffu);
}

/* ---------------------------------------------------------------- small dense helpers */
/* gslcblas dgemm NoTrans x NoTrans for 3x3 * 3xN (gsl/Matrix.cpp:90-95) */
static void mat3_mul_cols(const double A[9], const double* B /*3 x n*/, int n, double* C /*3 x n*/)
{
  for(int i = 0; i < 3 * n; i++) C[i] = 0.0;
  for(int k = 0; k < 3; k++)
    for(int i = 0; i < 3; i++)
    {
      const double temp = 1.0 * A[3 * i + k];
      if(temp != 0.0)
        for(int j = 0; j < n; j++) C[i * n + j] += temp * B[k * n + j];
    }
}

/* TSD_PDFMatching.cpp:206-221 (identical in RandomNormalMatching.cpp:251-263, PDFMatching.cpp:235-250).
 * Returns 0 when the hypothesis is skipped (fabs(phi) >= phiMax). */
static int hypothesis_transform(const double* M, const double* S, const double* phiM, const double* phiS, int idx,
                                int i, double phiMax, double T[9], double* phi_out)
{
  double phi = phiM[idx] - phiS[i];
  if(phi > M_PI) phi -= 2.0 * M_PI;
  else if(phi < -M_PI) phi += 2.0 * M_PI;
  *phi_out = phi;
  if(!(fabs(phi) < phiMax)) return 0;
  /* MatrixFactory.cpp:88-96 TransformationMatrix33(phi, 0, 0) */
  const double cphi = cos(phi);
  const double sphi = sin(phi);
  T[0] = cphi; T[1] = -sphi; T[2] = 0.0;
  T[3] = sphi; T[4] = cphi;  T[5] = 0.0;
  T[6] = 0.0;  T[7] = 0.0;   T[8] = 1.0;
  const double sx = S[2 * i];
  const double sy = S[2 * i + 1];
  T[2] = M[2 * idx] - (T[0] * sx + T[1] * sy);
  T[5] = M[2 * idx + 1] - (T[3] * sx + T[4] * sy);
  return 1;
}

static void set_identity3(double T[9])
{
  for(int i = 0; i < 9; i++) T[i] = (i % 4 == 0) ? 1.0 : 0.0;
}

/* ---------------------------------------------------------------- scorers */

/* TSD_PDFMatching.cpp:222-259 */
int port_match_score_tsd(port_grid_t* grid, int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n,
                         const double* model, const double* scene, const double* phi_m, const double* phi_s,
                         double phi_max, int32_t n_control, const double* control, const double t_sensor[9],
                         double zrand, double* score, int32_t* best, double t_best[9])
{
  (void)n;
  double* STemp = (double*)malloc(sizeof(double) * 3 * (n_control > 0 ? n_control : 1));
  double bestProb = 0.0;
  *best = -1;
  set_identity3(t_best);
  for(int h = 0; h < n_hyp; h++)
  {
    double T[9], phi;
    if(!hypothesis_transform(model, scene, phi_m, phi_s, hyps[h].idx_model, hyps[h].idx_scene, phi_max, T, &phi))
    {
      score[h] = -1.0;
      continue;
    }
    double TMap[9];
    mat3_mul_cols(t_sensor, T, 3, TMap); /* :222 TSensor * T */
    mat3_mul_cols(TMap, control, n_control, STemp); /* :225 */
    double probOfActualMeasurement = 1.0;
    for(int s = 0; s < n_control; s++)
    {
      double coord[2] = {STemp[s], STemp[n_control + s]};
      double tsd;
      if(!port_interpolate_bilinear_one(grid, coord, &tsd))
        probOfActualMeasurement *= (1.0 - (1.0 - zrand) * fabs(tsd));
      else
        probOfActualMeasurement *= zrand;
    }
    score[h] = probOfActualMeasurement;
    if(probOfActualMeasurement > bestProb)
    {
      memcpy(t_best, T, sizeof(T));
      bestProb = probOfActualMeasurement;
      *best = h;
    }
  }
  free(STemp);
  return TSD_OK;
}

/* RandomNormalMatching.cpp:265-359 */
int port_match_score_rnm(int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                         const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                         int32_t n_control, const double* control, const double* phi_control, int32_t n_valid,
                         const double* model_valid, const double* phi_valid, double theta_min, double theta_max,
                         double scale_distance, double scale_orientation, uint32_t cnt_match_thresh,
                         int32_t* cnt_match, int32_t* max_cnt_match, double* err_sum, int32_t* best, double t_best[9])
{
  (void)n;
  double* STemp = (double*)malloc(sizeof(double) * 3 * (n_control > 0 ? n_control : 1));
  uint8_t* maskControl = (uint8_t*)malloc(n_control > 0 ? n_control : 1);
  double bestRatio = 0.0;
  unsigned int bestCnt = 0;
  double bestErr = 1e12;
  *best = -1;
  set_identity3(t_best);
  for(int h = 0; h < n_hyp; h++)
  {
    double T[9], phi;
    if(!hypothesis_transform(model, scene, phi_m, phi_s, hyps[h].idx_model, hyps[h].idx_scene, phi_max, T, &phi))
    {
      cnt_match[h] = -1;
      max_cnt_match[h] = 0;
      err_sum[h] = 0.0;
      continue;
    }
    mat3_mul_cols(T, control, n_control, STemp); /* :266 */
    unsigned int maxCntMatch = 0;
    for(int j = 0; j < n_control; j++)
    {
      const double theta = atan2(STemp[n_control + j], STemp[j]);
      if(theta > theta_max || theta < theta_min) maskControl[j] = 0;
      else { maskControl[j] = 1; maxCntMatch++; }
    }
    unsigned int cntMatch = 0;
    double errSum = 0;
    for(int s = 0; s < n_control; s++)
    {
      if(maskControl[s])
      {
        /* exact 1-NN among the valid model points, lowest index on ties (flann shim) */
        const double q0 = STemp[s], q1 = STemp[n_control + s];
        int bi = -1;
        double bd = INFINITY;
        for(int k = 0; k < n_valid; k++)
        {
          const double d0 = q0 - model_valid[2 * k];
          const double d1 = q1 - model_valid[2 * k + 1];
          double d = 0.0;
          d += d0 * d0;
          d += d1 * d1;
          if(d < bd) { bd = d; bi = k; }
        }
        double distConsensus = bd;
        double normalConsensus = (1.0 - cos(phi_valid[bi] - phi_control[s] - phi)) / 2.0;
        double err = distConsensus * scale_distance + normalConsensus * scale_orientation;
        errSum += err;
        if(err < 1.0) cntMatch++;
      }
    }
    cnt_match[h] = (int32_t)cntMatch;
    max_cnt_match[h] = (int32_t)maxCntMatch;
    err_sum[h] = errSum;
    if(cntMatch <= cnt_match_thresh) continue;
    double ratio = (double)cntMatch / (double)maxCntMatch;
    {
      double equalThres = 1e-5;
      int rateCondition = ((ratio - bestRatio) > equalThres) && (cntMatch > bestCnt);
      /* `fabs( (ratio-bestRatio) < equalThres )` in the reference: fabs of a bool */
      int similarityCondition = fabs((double)((ratio - bestRatio) < equalThres)) && (cntMatch == bestCnt) && errSum < bestErr;
      int goodMatch = rateCondition || similarityCondition;
      if(goodMatch)
      {
        bestRatio = ratio;
        bestCnt = cntMatch;
        bestErr = errSum;
        memcpy(t_best, T, sizeof(T));
        *best = h;
      }
    }
  }
  free(STemp);
  free(maskControl);
  return TSD_OK;
}

/* PDFMatching.cpp:435-487; p = zhit zphi zshort zmax zrand percentagePointsInC rangemax sigphi sighit lamshort
 * maxAngleDiff maxAnglePenalty */
static double probability_of_two_single_scans(const double* p, double m, double s)
{
  const double zhit = p[0], zphi = p[1], zshort = p[2], zmax = p[3], zrand = p[4];
  const double rangemax = p[6], sigphi = p[7], sighit = p[8], lamshort = p[9];
  const double sigphit = 1.0 / (sqrt(2.0 * M_PI) * sighit); /* PDFMatching.cpp:33 */
  double phit = 0, pphi = 0, pshort = 0, pmax = 0, prand = 0;
  if(s < rangemax) phit = sigphit * pow(M_E, ((-0.5 * ((m - s) * (m - s))) / (sighit * sighit)));
  pphi = sigphi * pow(M_E, ((-0.5 * s * s) / (sigphi * sigphi)));
  if(s < m)
  {
    double n = 1.0 / (1.0 - pow(M_E, (-lamshort * m)));
    pshort = n * lamshort * pow(M_E, (-lamshort * s));
  }
  if(s >= rangemax) pmax = 1.0;
  if(s < rangemax) prand = 1.0 / rangemax;
  return zhit * phit + zshort * pshort + zmax * pmax + zrand * prand + zphi * pphi;
}

/* PDFMatching.cpp:252-388 */
int port_match_score_pdf(int32_t n_hyp, const tsd_hypothesis_t* hyps, int32_t n, const double* model,
                         const double* scene, const double* phi_m, const double* phi_s, double phi_max,
                         int32_t n_control, const double* control, int32_t n_valid, const double* model_angles,
                         const double* model_dists, const double params[12], double* prob, int32_t* fov_count,
                         int32_t* best, double t_best[9])
{
  (void)n;
  double* STemp = (double*)malloc(sizeof(double) * 3 * (n_control > 0 ? n_control : 1));
  const double angleThresh = (M_PI / 180.0) * params[10]; /* :227 */
  const double percentagePointsInC = params[5];
  double bestProb = 0.0;
  *best = -1;
  set_identity3(t_best);
  for(int h = 0; h < n_hyp; h++)
  {
    double T[9], phi;
    if(!hypothesis_transform(model, scene, phi_m, phi_s, hyps[h].idx_model, hyps[h].idx_scene, phi_max, T, &phi))
    {
      prob[h] = -1.0;
      fov_count[h] = 0;
      continue;
    }
    mat3_mul_cols(T, control, n_control, STemp); /* :253 */
    const unsigned int pointsInControl = (unsigned int)n_control;
    int fieldOfViewCount = 0;
    double probOfActualMeasurement = 1;
    for(unsigned int s = 0; s < pointsInControl; s++)
    {
      const double x = STemp[s], y = STemp[n_control + s];
      double angle = atan2(y, x);
      double distance = sqrt(x * x + y * y); /* pow(.,2) is folded to a product by the compiler */
      double minAngleDiff = 2 * M_PI;
      int idxMinAngleDiff = 0;
      for(int k = 0; k < n_valid; k++)
      {
        double diff = fabs(angle - model_angles[k]);
        if(diff < minAngleDiff) { minAngleDiff = diff; idxMinAngleDiff = k; }
      }
      if(minAngleDiff < angleThresh) fieldOfViewCount++;
      probOfActualMeasurement *= probability_of_two_single_scans(params, model_dists[idxMinAngleDiff], distance);
    }
    if(pointsInControl == 0) probOfActualMeasurement = 0; /* :359-363 */
    prob[h] = probOfActualMeasurement;
    fov_count[h] = fieldOfViewCount;
    if((probOfActualMeasurement > bestProb) && (fieldOfViewCount > pointsInControl * percentagePointsInC))
    {
      memcpy(t_best, T, sizeof(T));
      bestProb = probOfActualMeasurement;
      *best = h;
    }
  }
  free(STemp);
  return TSD_OK;
}

/* ---------------------------------------------------------------- pre-processing */

/* obvious::Matrix::pcaAnalysis (gsl/Matrix.cpp:227-327) for an n x 2 matrix; axes is 2 x 4 row-major */
static void pca_analysis(const double* A, int rows, double axes[8])
{
  const size_t dim = 2;
  gsl_matrix* Mo = gsl_matrix_alloc(rows, dim); /* the object's own _M */
  gsl_matrix* M = gsl_matrix_alloc(rows, dim);
  for(int i = 0; i < rows; i++) for(size_t j = 0; j < dim; j++) gsl_matrix_set(Mo, i, j, A[i * dim + j]);
  gsl_matrix_memcpy(M, Mo);
  gsl_matrix* V = gsl_matrix_alloc(dim, dim);
  gsl_vector* vcent = gsl_vector_alloc(dim);
  for(size_t i = 0; i < dim; i++)
  {
    gsl_vector_view col = gsl_matrix_column(Mo, i);
    double m = gsl_stats_mean(col.vector.data, dim, rows);
    gsl_vector_set(vcent, i, m);
  }
  double* cent = vcent->data;
  for(size_t i = 0; i < dim; i++)
  {
    gsl_vector_view c = gsl_matrix_column(M, i);
    gsl_vector_add_constant(&c.vector, -cent[i]);
  }
  gsl_matrix* MtM = gsl_matrix_alloc(dim, dim);
  gsl_blas_dgemm(CblasTrans, CblasNoTrans, 1.0, M, M, 0.0, MtM);
  gsl_vector* s = gsl_vector_alloc(dim);
  gsl_linalg_SV_decomp_jacobi(MtM, V, s);
  gsl_matrix* P = gsl_matrix_alloc(dim, rows);
  gsl_blas_dgemm(CblasTrans, CblasTrans, 1.0, V, M, 0.0, P);
  for(size_t i = 0; i < dim; i++)
  {
    gsl_vector_view coord = gsl_matrix_row(P, i);
    double max = gsl_vector_max(&coord.vector);
    double min = gsl_vector_min(&coord.vector);
    double ext = max - min;
    double align = 0.0;
    if(ext > 1e-6) align = (max + min) / 2.0;
    gsl_vector_view eigen = gsl_matrix_column(V, i);
    for(size_t j = 0; j < dim; j++)
    {
      double e = gsl_vector_get(&eigen.vector, j) * align;
      cent[j] += e;
    }
  }
  for(size_t i = 0; i < dim; i++)
  {
    gsl_vector_view coord = gsl_matrix_row(P, i);
    double ext = gsl_vector_max(&coord.vector) - gsl_vector_min(&coord.vector);
    gsl_vector_view eigen = gsl_matrix_column(V, i);
    for(size_t j = 0; j < dim; j++)
    {
      double e = gsl_vector_get(&eigen.vector, j) * ext / 2.0;
      axes[i * 4 + 2 * j] = cent[j] - e;
      axes[i * 4 + 2 * j + 1] = cent[j] + e;
    }
  }
  gsl_matrix_free(P);
  gsl_vector_free(s);
  gsl_matrix_free(MtM);
  gsl_vector_free(vcent);
  gsl_matrix_free(V);
  gsl_matrix_free(M);
  gsl_matrix_free(Mo);
}

/* RandomMatching.cpp:77-146 */
static void calc_normals(const double* M, double* N, int points, const uint8_t* maskIn, uint8_t* maskOut,
                         int searchRadius)
{
  for(int i = 0; i < searchRadius; i++) maskOut[i] = 0;
  for(int i = points - searchRadius; i < points; i++) maskOut[i] = 0;
  double* A = (double*)malloc(sizeof(double) * 2 * 2 * searchRadius);
  for(int i = searchRadius; i < points - searchRadius; i++)
  {
    if(maskIn[i])
    {
      unsigned int cnt = 0;
      for(int j = -searchRadius; j < searchRadius; j++)
        if(maskIn[i + j]) cnt++;
      if(cnt > 3)
      {
        cnt = 0;
        for(int j = -searchRadius; j < searchRadius; j++)
          if(maskIn[i + j])
          {
            A[2 * cnt] = M[2 * (i + j)];
            A[2 * cnt + 1] = M[2 * (i + j) + 1];
            cnt++;
          }
        double Axes[8];
        pca_analysis(A, (int)cnt, Axes);
        double xLong = Axes[1] - Axes[0];
        double yLong = Axes[3] - Axes[2];
        double xShort = Axes[5] - Axes[4];
        double yShort = Axes[7] - Axes[6];
        double lenLongSqr = xLong * xLong + yLong * yLong;
        double lenShortSqr = xShort * xShort + yShort * yShort;
        if(lenShortSqr > 1e-6 && (lenLongSqr / lenShortSqr) < 4.0)
        {
          maskOut[i] = 0;
          continue;
        }
        double len = sqrt(lenShortSqr);
        if((M[2 * i] * xShort + M[2 * i + 1] * yShort) < 0.0)
        {
          N[2 * i] = xShort / len;
          N[2 * i + 1] = yShort / len;
        }
        else
        {
          N[2 * i] = -xShort / len;
          N[2 * i + 1] = -yShort / len;
        }
      }
      else
        maskOut[i] = 0;
    }
  }
  free(A);
}

/* RandomMatching.cpp:148-169 */
static void calc_phi(const double* N, int rows, const uint8_t* mask, double* phi)
{
  for(int i = 0; i < rows; i++)
  {
    if(mask == NULL || mask[i]) phi[i] = atan2(N[2 * i + 1], N[2 * i]);
    else phi[i] = -1e6;
  }
}

void port_match_prep_free(port_match_prep_t* p)
{
  if(!p) return;
  free(p->phi_m); free(p->phi_s); free(p->mask_m_pca); free(p->mask_s_pca); free(p->idx_m_valid);
  free(p->idx_s_valid); free(p->idx_control); free(p->control); free(p->phi_control); free(p->hyps);
  free(p);
}

/* TSD_PDFMatching.cpp:59-205 == RandomNormalMatching.cpp:94-247 == PDFMatching.cpp:67-233 */
port_match_prep_t* port_match_prepare(int32_t n, const double* M, const uint8_t* maskM, const double* S,
                                      const uint8_t* maskS, uint32_t trials_in, uint32_t size_control_set,
                                      double phiMax, double resolution)
{
  const int pcaSearchRange = 10;
  const int pointsInM = n, pointsInS = n;
  if(pointsInM < 3) return NULL;

  port_match_prep_t* p = (port_match_prep_t*)calloc(1, sizeof(*p));
  p->n = n;
  /* ----------------- Model ------------------ */
  double* NMpca = (double*)calloc(2 * (size_t)n, sizeof(double));
  p->phi_m = (double*)malloc(sizeof(double) * n);
  p->mask_m_pca = (uint8_t*)malloc(n);
  memcpy(p->mask_m_pca, maskM, n);
  calc_normals(M, NMpca, n, maskM, p->mask_m_pca, pcaSearchRange / 2);
  calc_phi(NMpca, n, p->mask_m_pca, p->phi_m);
  /* RandomMatching.cpp:41-50 extractSamples */
  p->idx_m_valid = (int32_t*)malloc(sizeof(int32_t) * n);
  for(unsigned int i = pcaSearchRange / 2; i < (unsigned int)pointsInM - pcaSearchRange / 2; i++)
    if(p->mask_m_pca[i]) p->idx_m_valid[p->n_valid_m++] = (int32_t)i;

  /* ----------------- Scene ------------------- */
  double* NSpca = (double*)calloc(2 * (size_t)n, sizeof(double));
  p->phi_s = (double*)malloc(sizeof(double) * n);
  p->mask_s_pca = (uint8_t*)malloc(n);
  memcpy(p->mask_s_pca, maskS, n);
  unsigned int validPoints = 0;
  for(int i = 0; i < pointsInS; i++)
    if(p->mask_s_pca[i]) validPoints++;
  double probability = 180.0 / (double)validPoints;
  if(probability < 0.99)
  {
    /* RandomMatching.cpp:171-183 subsampleMask */
    double pr = probability;
    if(pr > 1.0) pr = 1.0;
    if(pr < 0.0) pr = 0.0;
    int probability_thresh = (int)(1000.0 - pr * 1000.0 + 0.5);
    for(int i = 0; i < pointsInS; i++)
      if((port_rand() % 1000) < probability_thresh) p->mask_s_pca[i] = 0;
  }
  calc_normals(S, NSpca, n, maskS, p->mask_s_pca, pcaSearchRange / 2);
  calc_phi(NSpca, n, p->mask_s_pca, p->phi_s);
  p->idx_s_valid = (int32_t*)malloc(sizeof(int32_t) * n);
  for(unsigned int i = pcaSearchRange / 2; i < (unsigned int)pointsInS - pcaSearchRange / 2; i++)
    if(p->mask_s_pca[i]) p->idx_s_valid[p->n_valid_s++] = (int32_t)i;

  /* --------------- Control set --------------- RandomMatching.cpp:52-75 pickControlSet */
  unsigned int sizeControlSet = size_control_set;
  if((unsigned int)p->n_valid_s < sizeControlSet) sizeControlSet = (unsigned int)p->n_valid_s;
  p->n_control = (int32_t)sizeControlSet;
  p->idx_control = (int32_t*)malloc(sizeof(int32_t) * (sizeControlSet + 1));
  p->control = (double*)malloc(sizeof(double) * 3 * (sizeControlSet + 1));
  p->phi_control = (double*)malloc(sizeof(double) * (sizeControlSet + 1));
  {
    int32_t* idxTemp = (int32_t*)malloc(sizeof(int32_t) * (p->n_valid_s + 1));
    int nTemp = p->n_valid_s;
    memcpy(idxTemp, p->idx_s_valid, sizeof(int32_t) * nTemp);
    unsigned int ctr = 0;
    while(ctr < sizeControlSet)
    {
      unsigned int r = (unsigned int)port_rand() % (unsigned int)nTemp;
      int32_t idx = idxTemp[r];
      p->idx_control[ctr] = idx;
      memmove(&idxTemp[r], &idxTemp[r + 1], sizeof(int32_t) * (nTemp - r - 1));
      nTemp--;
      p->control[0 * sizeControlSet + ctr] = S[2 * idx];
      p->control[1 * sizeControlSet + ctr] = S[2 * idx + 1];
      p->control[2 * sizeControlSet + ctr] = 1.0;
      ctr++;
    }
    free(idxTemp);
  }
  for(unsigned int i = 0; i < sizeControlSet; i++)
  {
    const double nx = NSpca[2 * p->idx_control[i]], ny = NSpca[2 * p->idx_control[i] + 1];
    p->phi_control[i] = atan2(ny, nx);
  }
  free(NMpca);
  free(NSpca);

  if(p->n_valid_s < 3 || p->n_valid_m < 3) { port_match_prep_free(p); return NULL; }

  /* frustum */
  p->theta_min = atan2(M[2 * p->idx_m_valid[0] + 1], M[2 * p->idx_m_valid[0]]);
  p->theta_max = atan2(M[2 * p->idx_m_valid[p->n_valid_m - 1] + 1], M[2 * p->idx_m_valid[p->n_valid_m - 1]]);

  unsigned int trials = trials_in;
  if((unsigned int)p->n_valid_m < trials_in) trials = (unsigned int)p->n_valid_m;

  phiMax = (phiMax <= M_PI * 0.5) ? phiMax : M_PI * 0.5;
  p->phi_max = phiMax;
  int span;
  if(resolution > 1e-6)
  {
    span = (int)floor(phiMax / resolution);
    if(span > (int)pointsInM) span = (int)pointsInM;
  }
  else { port_match_prep_free(p); return NULL; }
  p->span = span;

  /* trial loop: (idx, i) enumeration in single-thread order */
  int32_t* idxTrials = (int32_t*)malloc(sizeof(int32_t) * (p->n_valid_m + 1));
  int nTrials = p->n_valid_m;
  memcpy(idxTrials, p->idx_m_valid, sizeof(int32_t) * nTrials);
  size_t cap = 1024;
  p->hyps = (tsd_hypothesis_t*)malloc(sizeof(tsd_hypothesis_t) * cap);
  for(unsigned int trial = 0; trial < trials; trial++)
  {
    const int randIdx = port_rand() % nTrials;
    const int idx = idxTrials[randIdx];
    memmove(&idxTrials[randIdx], &idxTrials[randIdx + 1], sizeof(int32_t) * (nTrials - randIdx - 1));
    nTrials--;
    const int a = idx - span, b = pcaSearchRange / 2;
    const int iMin = (a >= b) ? a : b;
    const int c = idx + span, d = pointsInS - pcaSearchRange / 2;
    const int iMax = (c <= d) ? c : d;
    for(int i = iMin; i < iMax; i++)
    {
      if(p->mask_s_pca[i])
      {
        if((size_t)p->n_hyp == cap)
        {
          cap *= 2;
          p->hyps = (tsd_hypothesis_t*)realloc(p->hyps, sizeof(tsd_hypothesis_t) * cap);
        }
        p->hyps[p->n_hyp].idx_model = idx;
        p->hyps[p->n_hyp].idx_scene = i;
        p->n_hyp++;
      }
    }
  }
  free(idxTrials);
  return p;
}

/* ---------------------------------------------------------------- full match() */

void port_match_tsd(port_grid_t* grid, uint32_t trials, double eps_thresh, uint32_t size_control_set, double zrand,
                    const double t_sensor[9], int32_t n, const double* model, const uint8_t* mask_m,
                    const double* scene, const uint8_t* mask_s, double phi_max, double trans_max, double resolution,
                    double t_out[9])
{
  (void)eps_thresh; (void)trans_max;
  set_identity3(t_out);
  port_match_prep_t* p = port_match_prepare(n, model, mask_m, scene, mask_s, trials, size_control_set, phi_max, resolution);
  if(!p) return;
  double* score = (double*)malloc(sizeof(double) * (p->n_hyp + 1));
  int32_t best;
  port_match_score_tsd(grid, p->n_hyp, p->hyps, n, model, scene, p->phi_m, p->phi_s, p->phi_max, p->n_control,
                       p->control, t_sensor, zrand, score, &best, t_out);
  free(score);
  port_match_prep_free(p);
}

void port_match_rnm(uint32_t trials, double eps_thresh, uint32_t size_control_set, int32_t n, const double* model,
                    const uint8_t* mask_m, const double* scene, const uint8_t* mask_s, double phi_max,
                    double trans_max, double resolution, double t_out[9])
{
  (void)trans_max;
  set_identity3(t_out);
  port_match_prep_t* p = port_match_prepare(n, model, mask_m, scene, mask_s, trials, size_control_set, phi_max, resolution);
  if(!p) return;
  /* RandomNormalMatching.cpp:41-65 initKDTree over idxMValid */
  double* mv = (double*)malloc(sizeof(double) * 2 * p->n_valid_m);
  double* pv = (double*)malloc(sizeof(double) * p->n_valid_m);
  for(int k = 0; k < p->n_valid_m; k++)
  {
    mv[2 * k] = model[2 * p->idx_m_valid[k]];
    mv[2 * k + 1] = model[2 * p->idx_m_valid[k] + 1];
    pv[k] = p->phi_m[p->idx_m_valid[k]];
  }
  int32_t* cnt = (int32_t*)malloc(sizeof(int32_t) * (p->n_hyp + 1));
  int32_t* maxcnt = (int32_t*)malloc(sizeof(int32_t) * (p->n_hyp + 1));
  double* err = (double*)malloc(sizeof(double) * (p->n_hyp + 1));
  int32_t best;
  const double scaleDistance = 1.0 / (eps_thresh * eps_thresh); /* RandomNormalMatching.cpp:21-22 */
  const double scaleOrientation = 0.33;
  port_match_score_rnm(p->n_hyp, p->hyps, n, model, scene, p->phi_m, p->phi_s, p->phi_max, p->n_control, p->control,
                       p->phi_control, p->n_valid_m, mv, pv, p->theta_min, p->theta_max, scaleDistance,
                       scaleOrientation, (uint32_t)p->n_control / 3, cnt, maxcnt, err, &best, t_out);
  free(mv); free(pv); free(cnt); free(maxcnt); free(err);
  port_match_prep_free(p);
}

void port_match_pdf(uint32_t trials, double eps_thresh, uint32_t size_control_set, const double params[12], int32_t n,
                    const double* model, const uint8_t* mask_m, const double* scene, const uint8_t* mask_s,
                    double phi_max, double trans_max, double resolution, double t_out[9])
{
  (void)eps_thresh; (void)trans_max;
  set_identity3(t_out);
  port_match_prep_t* p = port_match_prepare(n, model, mask_m, scene, mask_s, trials, size_control_set, phi_max, resolution);
  if(!p) return;
  /* PDFMatching.cpp:196-204 */
  double* ang = (double*)malloc(sizeof(double) * p->n_valid_m);
  double* dst = (double*)malloc(sizeof(double) * p->n_valid_m);
  for(int k = 0; k < p->n_valid_m; k++)
  {
    const double x = model[2 * p->idx_m_valid[k]], y = model[2 * p->idx_m_valid[k] + 1];
    ang[k] = atan2(y, x);
    dst[k] = sqrt(x * x + y * y);
  }
  double* prob = (double*)malloc(sizeof(double) * (p->n_hyp + 1));
  int32_t* fov = (int32_t*)malloc(sizeof(int32_t) * (p->n_hyp + 1));
  int32_t best;
  port_match_score_pdf(p->n_hyp, p->hyps, n, model, scene, p->phi_m, p->phi_s, p->phi_max, p->n_control, p->control,
                       p->n_valid_m, ang, dst, params, prob, fov, &best, t_out);
  free(ang); free(dst); free(prob); free(fov);
  port_match_prep_free(p);
}
