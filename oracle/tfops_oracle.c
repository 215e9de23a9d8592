/* CPU restatement of the reference tf_ops -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Plain C, one function per reference kernel.  The GPU-only reference kernels (sampling, grouping) are restated
 * with the arithmetic nvcc generates for them (-fmad=true contracts a*a+b*b+c*c into mul, fma, fma -- checked in the
 * SASS of the unmodified tf_sampling_g.cu / tf_grouping_g.cu built for sm_100a), the CPU-only interpolation ops with
 * plain float arithmetic.  Build with -ffp-contract=off so the compiler adds no contractions of its own.
 *
 * Pinned on the GPU box against the reference kernels themselves (oracle/_ref, oracle/build_ref.py):
 * tests/test_tfops_gpu.py.
 */
#include <math.h>
#include <string.h>

static float sqdist_gpu(float x1, float y1, float z1, float x2, float y2, float z2) {
  /* (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) as compiled by nvcc: FMUL, FFMA, FFMA */
  float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* tf_sampling_g.cu:105-170 farthestpointsamplingKernel.  Tie rule of the 512-thread argmax: each thread keeps the
 * first maximum of its stride (strict '>'), the tree keeps the lower slot on ties => smallest (k mod 512), then k. */
void oracle_farthest_point_sampling(int b, int n, int m, const float* inp, float* temp /* n floats */, int* out) {
  for (int i = 0; i < b; ++i) {
    const float* pts = inp + (size_t)i * n * 3;
    if (m <= 0) continue;
    int old = 0;
    out[(size_t)i * m] = 0;
    for (int k = 0; k < n; ++k) temp[k] = 1e38f;
    for (int j = 1; j < m; ++j) {
      float x1 = pts[old * 3], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
      float slot_best[512];
      int slot_besti[512];
      for (int t = 0; t < 512; ++t) { slot_best[t] = -1.f; slot_besti[t] = 0; }
      for (int k = 0; k < n; ++k) {
        float d = sqdist_gpu(x1, y1, z1, pts[k * 3], pts[k * 3 + 1], pts[k * 3 + 2]);
        float d2 = d < temp[k] ? d : temp[k];
        temp[k] = d2;
        int t = k & 511;
        if (d2 > slot_best[t]) { slot_best[t] = d2; slot_besti[t] = k; }
      }
      int bt = 0;
      for (int t = 1; t < 512; ++t)
        if (slot_best[bt] < slot_best[t]) bt = t;
      old = slot_besti[bt];
      out[(size_t)i * m + j] = old;
    }
  }
}

/* tf_sampling_g.cu:172-181 */
void oracle_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int c = 0; c < 3; ++c) out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
    }
}

/* tf_sampling_g.cu:183-192 (sequential accumulation; the GPU uses atomics in arbitrary order) */
void oracle_scatter_add_point(int b, int n, int m, const float* out_g, const int* idx, float* inp_g) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < m; ++j) {
      int a = idx[(size_t)i * m + j];
      for (int c = 0; c < 3; ++c) inp_g[((size_t)i * n + a) * 3 + c] += out_g[((size_t)i * m + j) * 3 + c];
    }
}

/* tf_grouping_g.cu:3-36 query_ball_point_gpu; rows without a hit are left untouched */
void oracle_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2, int* idx,
                             int* pts_cnt) {
  for (int bi = 0; bi < b; ++bi) {
    const float* p1 = xyz1 + (size_t)bi * n * 3;
    const float* p2 = xyz2 + (size_t)bi * m * 3;
    for (int j = 0; j < m; ++j) {
      int* row = idx + ((size_t)bi * m + j) * nsample;
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        float d2 = sqdist_gpu(p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2], p2[j * 3], p2[j * 3 + 1], p2[j * 3 + 2]);
        float d = sqrtf(d2);
        if (d < 1e-20f) d = 1e-20f;
        if (d < radius) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt++] = k;
        }
      }
      pts_cnt[(size_t)bi * m + j] = cnt;
    }
  }
}

/* tf_grouping_g.cu:40-57 */
void oracle_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out) {
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m * nsample; ++j) {
      int ii = idx[(size_t)bi * m * nsample + j];
      memcpy(out + ((size_t)bi * m * nsample + j) * c, points + ((size_t)bi * n + ii) * c, sizeof(float) * c);
    }
}

/* tf_grouping_g.cu:61-78 */
void oracle_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx, float* grad_points) {
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m * nsample; ++j) {
      int ii = idx[(size_t)bi * m * nsample + j];
      for (int l = 0; l < c; ++l) grad_points[((size_t)bi * n + ii) * c + l] += grad_out[((size_t)bi * m * nsample + j) * c + l];
    }
}

/* tf_grouping_g.cu:83-123 selection_sort_gpu */
void oracle_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out) {
  for (size_t r = 0; r < (size_t)b * m; ++r) {
    float* o = out + r * n;
    int* oi = outi + r * n;
    for (int s = 0; s < n; ++s) { o[s] = dist[r * n + s]; oi[s] = s; }
    for (int s = 0; s < k && s < n; ++s) {
      int mn = s;
      for (int t = s + 1; t < n; ++t)
        if (o[t] < o[mn]) mn = t;
      if (mn != s) {
        float tv = o[mn]; o[mn] = o[s]; o[s] = tv;
        int ti = oi[mn]; oi[mn] = oi[s]; oi[s] = ti;
      }
    }
  }
}

/* tf_interpolate.cpp:60-103 threenn_cpu */
void oracle_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      float x1 = xyz1[j * 3], y1 = xyz1[j * 3 + 1], z1 = xyz1[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int b1 = 0, b2 = 0, b3 = 0;
      for (int k = 0; k < m; ++k) {
        float x2 = xyz2[k * 3], y2 = xyz2[k * 3 + 1], z2 = xyz2[k * 3 + 2];
        float df = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        double d = df;
        if (d < best1) { best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k; }
        else if (d < best2) { best3 = best2; b3 = b2; best2 = d; b2 = k; }
        else if (d < best3) { best3 = d; b3 = k; }
      }
      dist[j * 3] = (float)best1; idx[j * 3] = b1;
      dist[j * 3 + 1] = (float)best2; idx[j * 3 + 1] = b2;
      dist[j * 3 + 2] = (float)best3; idx[j * 3 + 2] = b3;
    }
    xyz1 += n * 3; xyz2 += m * 3; dist += n * 3; idx += n * 3;
  }
}

/* tf_interpolate.cpp:107-127 threeinterpolate_cpu */
void oracle_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight, float* out) {
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      float w1 = weight[j * 3], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
      int i1 = idx[j * 3], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
      for (int l = 0; l < c; ++l) out[j * c + l] = points[i1 * c + l] * w1 + points[i2 * c + l] * w2 + points[i3 * c + l] * w3;
    }
    points += m * c; idx += n * 3; weight += n * 3; out += n * c;
  }
}

/* tf_interpolate.cpp:131-153 threeinterpolate_grad_cpu */
void oracle_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points) {
  for (int i = 0; i < b; ++i) {
    for (int j = 0; j < n; ++j) {
      float w1 = weight[j * 3], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
      int i1 = idx[j * 3], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
      for (int l = 0; l < c; ++l) {
        grad_points[i1 * c + l] += grad_out[j * c + l] * w1;
        grad_points[i2 * c + l] += grad_out[j * c + l] * w2;
        grad_points[i3 * c + l] += grad_out[j * c + l] * w3;
      }
    }
    grad_out += n * c; idx += n * 3; weight += n * 3; grad_points += m * c;
  }
}
