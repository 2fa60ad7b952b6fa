/* TEST INFRASTRUCTURE ONLY -- part of oracle/ (see oracle/port/port.h).
 *
 * Restatement of obvious::Icp with FlannPairAssignment, OutOfBoundsFilter2D, DistanceFilter,
 * ReciprocalFilter and ClosedFormEstimator2D as ThreadLocalize wires them
 * (reference src/ThreadLocalize.cpp:210-225, :571-581; src/obvision/registration/icp/...).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "port.h"

struct port_icp
{
  uint32_t max_iterations;
  uint32_t conv_cnt;
  double max_rms;
  /* DistanceFilter.cpp:11-25 */
  double max_dist_sqr, min_dist_sqr, dist_sqr, multiplier;
  /* OutOfBoundsFilter2D.cpp:8-15 */
  double x_min, x_max, y_min, y_max;
  /* trace of the last run */
  int32_t tr_cap, tr_it, tr_maxit;
  uint32_t* tr_model;
  uint32_t* tr_scene;
  int32_t* tr_count;
  double* tr_mse;
  double* tr_T;
};

port_icp_t* port_icp_create(uint32_t max_iterations, double dist_max, double dist_min, uint32_t dist_iterations,
                            const double bounds[4])
{
  port_icp_t* icp = (port_icp_t*)calloc(1, sizeof(*icp));
  icp->max_iterations = max_iterations; /* ThreadLocalize.cpp:224 */
  icp->conv_cnt = max_iterations;       /* :225 */
  icp->max_rms = 0.0;                   /* :223 */
  /* DistanceFilter.cpp:11-20 */
  icp->max_dist_sqr = dist_max * dist_max;
  icp->min_dist_sqr = dist_min * dist_min;
  icp->dist_sqr = icp->max_dist_sqr;
  double it = (double)(dist_iterations - 1);
  if(dist_iterations < 1) it = 1.0;
  icp->multiplier = pow((dist_min / dist_max), 1.0 / it);
  icp->x_min = bounds[0];
  icp->x_max = bounds[1];
  icp->y_min = bounds[2];
  icp->y_max = bounds[3];
  return icp;
}

void port_icp_destroy(port_icp_t* icp)
{
  if(!icp) return;
  free(icp->tr_model); free(icp->tr_scene); free(icp->tr_count); free(icp->tr_mse); free(icp->tr_T);
  free(icp);
}

/* gslcblas dgemm NoTrans x NoTrans 4x4 (Icp.cpp:454 `(*_Tlast) * (*_Tfinal4x4)`, gsl/Matrix.cpp:90-95) */
static void mat4_mul(const double* A, const double* B, double* C)
{
  double out[16];
  for(int i = 0; i < 16; i++) out[i] = 0.0;
  for(int k = 0; k < 4; k++)
    for(int i = 0; i < 4; i++)
    {
      const double temp = 1.0 * A[4 * i + k];
      if(temp != 0.0)
        for(int j = 0; j < 4; j++) out[4 * i + j] += temp * B[4 * k + j];
    }
  memcpy(C, out, sizeof(out));
}

/* Icp.cpp:371-408 applyTransformation for dim 2: data <- data * R^T (dgemm NoTrans x Trans,
 * gsl/Matrix.cpp:489-496), then the translation column is added */
static void apply_transformation(double* data, int size, const double* T44)
{
  const double R[2][2] = {{T44[0], T44[1]}, {T44[4], T44[5]}};
  for(int i = 0; i < size; i++)
  {
    const double x = data[2 * i], y = data[2 * i + 1];
    double out[2];
    for(int j = 0; j < 2; j++)
    {
      double temp = 0.0;
      temp += x * R[j][0];
      temp += y * R[j][1];
      out[j] = 0.0 + 1.0 * temp;
    }
    data[2 * i] = out[0];
    data[2 * i + 1] = out[1];
  }
  for(int i = 0; i < size; i++)
  {
    data[2 * i] += T44[3];
    data[2 * i + 1] += T44[7];
  }
}

typedef struct
{
  unsigned int idx;
  unsigned int model;
  double dist;
} recip_t;

/* ReciprocalFilter.cpp:16-21 operator<, with the original position as the final tie-break so that the
 * order is total (the reference uses an unstable std::sort: exact (model, dist) ties are undefined there) */
static int recip_cmp(const void* a, const void* b)
{
  const recip_t* f = (const recip_t*)a;
  const recip_t* s = (const recip_t*)b;
  if(f->model < s->model) return -1;
  if(f->model > s->model) return 1;
  if(f->dist < s->dist) return -1;
  if(f->dist > s->dist) return 1;
  return (f->idx < s->idx) ? -1 : (f->idx > s->idx);
}

int port_icp_run(port_icp_t* icp, const double* model_in, const double* normals, int32_t n_model,
                 const double* scene_in, int32_t n_scene, const double pose[9], const double* t_init,
                 double t_out[9], double* mse, uint32_t* pairs_out, uint32_t* iterations, int32_t* state)
{
  (void)normals; /* ClosedFormEstimator2D ignores normals (ClosedFormEstimator2D.cpp:26-34) */
  double Tfinal[16], Tlast[16];
  for(int i = 0; i < 16; i++) Tfinal[i] = Tlast[i] = (i % 5 == 0) ? 1.0 : 0.0;

  /* Icp::reset (Icp.cpp:333-339) -> PairAssignment::reset -> DistanceFilter::reset (DistanceFilter.cpp:27-30) */
  icp->dist_sqr = icp->max_dist_sqr;

  *mse = 0.0;
  *pairs_out = 0;
  *iterations = 0;
  for(int i = 0; i < 9; i++) t_out[i] = (i % 4 == 0) ? 1.0 : 0.0;

  /* Icp.cpp:467-471 */
  if(n_model == 0 || n_scene == 0)
  {
    *state = TSD_ICP_NOTMATCHABLE;
    return TSD_OK;
  }

  double* model = (double*)malloc(sizeof(double) * 2 * n_model);
  double* scene = (double*)malloc(sizeof(double) * 2 * n_scene);
  memcpy(model, model_in, sizeof(double) * 2 * n_model);
  memcpy(scene, scene_in, sizeof(double) * 2 * n_scene);

  const int cap = n_model > n_scene ? n_model : n_scene;
  free(icp->tr_model); free(icp->tr_scene); free(icp->tr_count); free(icp->tr_mse); free(icp->tr_T);
  icp->tr_cap = cap;
  icp->tr_maxit = (int32_t)icp->max_iterations;
  icp->tr_it = 0;
  icp->tr_model = (uint32_t*)calloc((size_t)cap * icp->max_iterations + 1, sizeof(uint32_t));
  icp->tr_scene = (uint32_t*)calloc((size_t)cap * icp->max_iterations + 1, sizeof(uint32_t));
  icp->tr_count = (int32_t*)calloc(icp->max_iterations + 1, sizeof(int32_t));
  icp->tr_mse = (double*)calloc(icp->max_iterations + 1, sizeof(double));
  icp->tr_T = (double*)calloc(16 * (size_t)icp->max_iterations + 16, sizeof(double));

  uint8_t* mask = (uint8_t*)malloc(n_scene);
  unsigned int* pm = (unsigned int*)malloc(sizeof(unsigned int) * n_scene); /* pair model idx  */
  unsigned int* ps = (unsigned int*)malloc(sizeof(unsigned int) * n_scene); /* pair scene idx  */
  double* pd = (double*)malloc(sizeof(double) * n_scene);
  unsigned int* fm = (unsigned int*)malloc(sizeof(unsigned int) * n_scene);
  unsigned int* fs = (unsigned int*)malloc(sizeof(unsigned int) * n_scene);
  double* fd = (double*)malloc(sizeof(double) * n_scene);
  recip_t* rp = (recip_t*)malloc(sizeof(recip_t) * n_scene);

  /* Icp.cpp:480-487 */
  if(t_init)
  {
    apply_transformation(scene, n_scene, t_init);
    mat4_mul(t_init, Tfinal, Tfinal);
  }

  int eRetval = TSD_ICP_PROCESSING;
  unsigned int iter = 0;
  double rms_prev = 10e12;
  unsigned int conv_cnt = 0;
  double rms = *mse;
  unsigned int pairs = 0;
  while(eRetval == TSD_ICP_PROCESSING)
  {
    /* ---- Icp::step (Icp.cpp:410-462) ---- */
    /* PairAssignment::determinePairs (PairAssignment.cpp:38-84): pre-filter */
    memset(mask, 1, n_scene);
    {
      /* OutOfBoundsFilter2D.cpp:27-37: S.transform(pose) = S * R^T + t (gsl/Matrix.cpp:403-432) */
      for(int i = 0; i < n_scene; i++)
      {
        const double x = scene[2 * i], y = scene[2 * i + 1];
        double tx = 0.0;
        tx += x * pose[0];
        tx += y * pose[1];
        tx = 0.0 + 1.0 * tx;
        double ty = 0.0;
        ty += x * pose[3];
        ty += y * pose[4];
        ty = 0.0 + 1.0 * ty;
        tx += pose[2];
        ty += pose[5];
        if(tx < icp->x_min || tx > icp->x_max || ty < icp->y_min || ty > icp->y_max) mask[i] = 0;
      }
    }
    /* FlannPairAssignment.cpp:64-92: exact 1-NN, squared L2, lowest model index on ties (flann shim) */
    int np = 0;
    for(int i = 0; i < n_scene; i++)
    {
      if(mask[i] == 1)
      {
        int best = -1;
        double bestD = INFINITY;
        for(int k = 0; k < n_model; k++)
        {
          const double d0 = scene[2 * i] - model[2 * k];
          const double d1 = scene[2 * i + 1] - model[2 * k + 1];
          double d = 0.0;
          d += d0 * d0;
          d += d1 * d1;
          if(d < bestD) { bestD = d; best = k; }
        }
        pm[np] = (unsigned int)best;
        ps[np] = (unsigned int)i;
        pd[np] = bestD;
        np++;
      }
    }
    /* DistanceFilter.cpp:32-64 */
    int nf = 0;
    for(int p = 0; p < np; p++)
    {
      if(pd[p] <= icp->dist_sqr)
      {
        fm[nf] = pm[p]; fs[nf] = ps[p]; fd[nf] = pd[p];
        nf++;
      }
    }
    icp->dist_sqr *= icp->multiplier;
    if(icp->dist_sqr < icp->min_dist_sqr) icp->dist_sqr = icp->min_dist_sqr;
    /* ReciprocalFilter.cpp:32-78 (input = output of the distance filter, PairAssignment.cpp:61-67) */
    np = nf;
    memcpy(pm, fm, sizeof(unsigned int) * np);
    memcpy(ps, fs, sizeof(unsigned int) * np);
    memcpy(pd, fd, sizeof(double) * np);
    nf = 0;
    if(np > 0)
    {
      for(int i = 0; i < np; i++) { rp[i].idx = i; rp[i].model = pm[i]; rp[i].dist = pd[i]; }
      qsort(rp, np, sizeof(recip_t), recip_cmp);
      unsigned int last = rp[0].model;
      fm[nf] = pm[rp[0].idx]; fs[nf] = ps[rp[0].idx]; nf++;
      for(int i = 1; i < np; i++)
      {
        if(rp[i].model == last) continue;
        last = rp[i].model;
        fm[nf] = pm[rp[i].idx]; fs[nf] = ps[rp[i].idx]; nf++;
      }
    }
    pairs = (unsigned int)nf;

    int retval = TSD_ICP_PROCESSING;
    if(pairs > 2)
    {
      /* ClosedFormEstimator2D::setPairs (ClosedFormEstimator2D.cpp:36-67) */
      double cm[2] = {0.0, 0.0}, cs[2] = {0.0, 0.0};
      double r = 0.0;
      for(unsigned int i = 0; i < pairs; i++)
      {
        const double* pointModel = &model[2 * fm[i]];
        const double* pointScene = &scene[2 * fs[i]];
        cm[0] += pointModel[0];
        cm[1] += pointModel[1];
        cs[0] += pointScene[0];
        cs[1] += pointScene[1];
        /* mathbase.h:147-153 distSqr2D(pointModel, pointScene) */
        const double dx = pointScene[0] - pointModel[0];
        const double dy = pointScene[1] - pointModel[1];
        r += dx * dx + dy * dy;
      }
      double sizeInv = 1.0 / (double)pairs;
      r *= sizeInv;
      cm[0] *= sizeInv; cm[1] *= sizeInv; cs[0] *= sizeInv; cs[1] *= sizeInv;
      rms = r;
      /* estimateTransformation (:74-109) */
      double nominator = 0.0, denominator = 0.0;
      for(unsigned int i = 0; i < pairs; i++)
      {
        double xFCm = model[2 * fm[i]] - cm[0];
        double yFCm = model[2 * fm[i] + 1] - cm[1];
        double xSCs = scene[2 * fs[i]] - cs[0];
        double ySCs = scene[2 * fs[i] + 1] - cs[1];
        nominator += yFCm * xSCs - xFCm * ySCs;
        denominator += xFCm * xSCs + yFCm * ySCs;
      }
      double deltaTheta = atan2(nominator, denominator);
      double cosDeltaTheta = cos(deltaTheta);
      double sinDeltaTheta = sin(deltaTheta);
      double deltaX = (cm[0] - (cosDeltaTheta * cs[0] - sinDeltaTheta * cs[1]));
      double deltaY = (cm[1] - (cosDeltaTheta * cs[1] + sinDeltaTheta * cs[0]));
      for(int i = 0; i < 16; i++) Tlast[i] = (i % 5 == 0) ? 1.0 : 0.0;
      Tlast[0] = cosDeltaTheta; Tlast[1] = -sinDeltaTheta; Tlast[3] = deltaX;
      Tlast[4] = sinDeltaTheta; Tlast[5] = cosDeltaTheta;  Tlast[7] = deltaY;
      Tlast[11] = 0;
      /* Icp.cpp:449-454 */
      apply_transformation(scene, n_scene, Tlast);
      mat4_mul(Tlast, Tfinal, Tfinal);
    }
    else
    {
      retval = TSD_ICP_NOTMATCHABLE;
    }
    /* trace */
    if((int)iter < icp->tr_maxit)
    {
      icp->tr_count[iter] = nf;
      for(int k = 0; k < nf; k++)
      {
        icp->tr_model[(size_t)iter * cap + k] = fm[k];
        icp->tr_scene[(size_t)iter * cap + k] = fs[k];
      }
      icp->tr_mse[iter] = rms;
      memcpy(&icp->tr_T[16 * iter], Tfinal, sizeof(Tfinal));
      icp->tr_it = (int32_t)iter + 1;
    }
    eRetval = retval;
    /* ---- Icp::iterate bookkeeping (Icp.cpp:495-508) ---- */
    iter++;
    if(fabs(rms - rms_prev) < 10e-10) conv_cnt++;
    else conv_cnt = 0;
    if((rms <= icp->max_rms || conv_cnt >= icp->conv_cnt)) eRetval = TSD_ICP_SUCCESS;
    else if(iter >= icp->max_iterations) eRetval = TSD_ICP_MAXITERATIONS;
    rms_prev = rms;
  }
  *iterations = iter;
  *mse = rms;
  *pairs_out = pairs;
  *state = eRetval;

  /* Icp.cpp:528-546 getFinalTransformation */
  t_out[0] = Tfinal[0]; t_out[1] = Tfinal[1]; t_out[2] = Tfinal[3];
  t_out[3] = Tfinal[4]; t_out[4] = Tfinal[5]; t_out[5] = Tfinal[7];
  t_out[6] = 0; t_out[7] = 0; t_out[8] = 1;

  free(model); free(scene); free(mask); free(pm); free(ps); free(pd); free(fm); free(fs); free(fd); free(rp);
  return TSD_OK;
}

int port_icp_get_trace(port_icp_t* icp, int32_t max_it, int32_t cap, uint32_t* pair_model, uint32_t* pair_scene,
                       int32_t* pair_count, double* mse, double* t_final16, int32_t* n_it)
{
  const int its = icp->tr_it < max_it ? icp->tr_it : max_it;
  for(int it = 0; it < its; it++)
  {
    pair_count[it] = icp->tr_count[it];
    const int n = icp->tr_count[it] < cap ? icp->tr_count[it] : cap;
    for(int k = 0; k < n; k++)
    {
      pair_model[(size_t)it * cap + k] = icp->tr_model[(size_t)it * icp->tr_cap + k];
      pair_scene[(size_t)it * cap + k] = icp->tr_scene[(size_t)it * icp->tr_cap + k];
    }
    mse[it] = icp->tr_mse[it];
    memcpy(&t_final16[16 * it], &icp->tr_T[16 * it], 16 * sizeof(double));
  }
  *n_it = its;
  return TSD_OK;
}
