// Rigid-transform estimation shared by the device kernels and the host refit
// (replaces estimateRigidTransform3D / estimateRigidTransform2D, danielsuo/cuSIFT
// extras/rigidTransform.cu:15-290, and the math_utils.cu helpers they call).
//
// 3-D: Horn-style quaternion fit.  For the selected correspondences (ref x_i, mov y_i):
// centre both sets, build the skew-symmetric 4x4  A_i = [[0, (y-x)^T], [-(y-x), [y+x]_x]],
// B = sum A_i A_i^T, the unit quaternion is the singular vector of B's smallest singular value
// (rigidTransform.cu:66-165), R = quat2rot(q) (math_utils.cu:266-280) and
// Rt = T(x_centroid) * R * T(-y_centroid) (rigidTransform.cu:170-196).
// The reference obtains the vector with a Numerical-Recipes SVD in float (dsvd); B is symmetric
// positive semi-definite, so the same vector is its eigenvector of the smallest eigenvalue, computed
// here with cyclic Jacobi rotations in double (no heap allocation, no iteration limits to tune).  The
// result agrees with the reference's to float rounding; the sign of q does not matter (R is even in q).
#ifndef CSB_RIGID_MATH_H
#define CSB_RIGID_MATH_H

#include <math.h>

#ifdef __CUDACC__
#define CSB_HD __host__ __device__ __forceinline__
#else
#define CSB_HD static inline
#endif

// eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix a (destroyed)
CSB_HD void csb_min_eigvec4(double a[4][4], double q[4]) {
  double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0.0;
    for (int i = 0; i < 4; i++)
      for (int j = i + 1; j < 4; j++) off += a[i][j] * a[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; p++)
      for (int r = p + 1; r < 4; r++) {
        if (fabs(a[p][r]) < 1e-300) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; k++) {          // columns p, r of a
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - s * akr;
          a[k][r] = s * akp + c * akr;
        }
        for (int k = 0; k < 4; k++) {          // rows p, r of a
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - s * ark;
          a[r][k] = s * apk + c * ark;
        }
        for (int k = 0; k < 4; k++) {
          const double vkp = v[k][p], vkr = v[k][r];
          v[k][p] = c * vkp - s * vkr;
          v[k][r] = s * vkp + c * vkr;
        }
      }
  }
  int m = 0;
  for (int i = 1; i < 4; i++)
    if (a[i][i] < a[m][m]) m = i;
  for (int i = 0; i < 4; i++) q[i] = v[i][m];
}

// coord: numPts x 6 floats (ref xyz, mov xyz); idx: n point indices; Rt: 3 x 4 row-major
CSB_HD void csb_rigid3d(const float *coord, const int *idx, int n, float *Rt) {
  float xc[3] = {0.f, 0.f, 0.f}, yc[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < n; i++) {
    const float *p = coord + 6 * idx[i];
    for (int d = 0; d < 3; d++) {
      xc[d] += p[d];
      yc[d] += p[3 + d];
    }
  }
  for (int d = 0; d < 3; d++) {
    xc[d] = xc[d] / n;
    yc[d] = yc[d] / n;
  }
  // B = sum A A^T with A = [[0, d^T], [-d, [s]_x]], d = y - x, s = y + x (centred): expanded, in float like
  // the reference's A / multi4by4 / B accumulation
  double B[4][4];
  float Bf[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) Bf[i][j] = 0.f;
  for (int i = 0; i < n; i++) {
    const float *p = coord + 6 * idx[i];
    float x[3], y[3];
    for (int d = 0; d < 3; d++) {
      x[d] = p[d] - xc[d];
      y[d] = p[3 + d] - yc[d];
    }
    const float s0 = y[0] + x[0], s1 = y[1] + x[1], s2 = y[2] + x[2];
    float A[4][4];
    A[0][0] = 0.f;            A[0][1] = y[0] - x[0];  A[0][2] = y[1] - x[1];  A[0][3] = y[2] - x[2];
    A[1][0] = -y[0] + x[0];   A[1][1] = 0.f;          A[1][2] = -s2;          A[1][3] = s1;
    A[2][0] = -y[1] + x[1];   A[2][1] = s2;           A[2][2] = 0.f;          A[2][3] = -s0;
    A[3][0] = -y[2] + x[2];   A[3][1] = -s1;          A[3][2] = s0;           A[3][3] = 0.f;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        float acc = 0.f;
        for (int k = 0; k < 4; k++) acc += A[r][k] * A[c][k];
        Bf[r][c] += acc;
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) B[i][j] = 0.5 * ((double)Bf[i][j] + (double)Bf[j][i]);
  double qd[4];
  csb_min_eigvec4(B, qd);
  const float Q[4] = {(float)qd[0], (float)qd[1], (float)qd[2], (float)qd[3]};
  float R[9];   // math_utils.cu:266-280 (the products are float, the combination double, as written there)
  R[0] = (float)(1.0 - 2.0 * (Q[2] * Q[2] + Q[3] * Q[3]));
  R[1] = (float)(2.0 * (Q[1] * Q[2] - Q[0] * Q[3]));
  R[2] = (float)(2.0 * (Q[1] * Q[3] + Q[0] * Q[2]));
  R[3] = (float)(2.0 * (Q[1] * Q[2] + Q[0] * Q[3]));
  R[4] = (float)(1.0 - 2.0 * (Q[1] * Q[1] + Q[3] * Q[3]));
  R[5] = (float)(2.0 * (Q[2] * Q[3] - Q[0] * Q[1]));
  R[6] = (float)(2.0 * (Q[1] * Q[3] - Q[0] * Q[2]));
  R[7] = (float)(2.0 * (Q[2] * Q[3] + Q[0] * Q[1]));
  R[8] = (float)(1.0 - 2.0 * (Q[1] * Q[1] + Q[2] * Q[2]));
  // T3 * (T2 * T1): rotation R, translation x_c - R y_c (rigidTransform.cu:170-196)
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Rt[4 * r + c] = R[3 * r + c];
    const float ry = R[3 * r + 0] * -yc[0] + R[3 * r + 1] * -yc[1] + R[3 * r + 2] * -yc[2];
    Rt[4 * r + 3] = ry + xc[r];
  }
}

// Planar (x, z) two-point fit, y ignored (rigidTransform.cu:214-290); a, b: the two point indices
CSB_HD void csb_rigid2d(const float *coord, int a, int b, float *Rt) {
  const float *A = coord + 6 * a, *Bp = coord + 6 * b;
  const float wAx = A[0], wAz = A[2], cAx = A[3], cAz = A[5];
  const float wBx = Bp[0], wBz = Bp[2], cBx = Bp[3], cBz = Bp[5];
  const float dxw = wAx - wBx, dzw = wAz - wBz;
  const float lw = sqrtf(dxw * dxw + dzw * dzw);
  const float dxwn = dxw / lw, dzwn = dzw / lw;
  const float dxc = cAx - cBx, dzc = cAz - cBz;
  const float lc = sqrtf(dxc * dxc + dzc * dzc);
  const float dxcn = dxc / lc, dzcn = dzc / lc;
  const float cosA = dxwn * dxcn + dzwn * dzcn;
  const float sinA = dzwn * dxcn - dxwn * dzcn;
  const float sxw = wAx + wBx, szw = wAz + wBz, sxc = cAx + cBx, szc = cAz + cBz;
  Rt[0] = cosA; Rt[1] = 0.f; Rt[2] = -sinA;
  Rt[4] = 0.f;  Rt[5] = 1.f; Rt[6] = 0.f;
  Rt[8] = sinA; Rt[9] = 0.f; Rt[10] = cosA;
  Rt[3] = (sxw - cosA * sxc + sinA * szc) / 2;
  Rt[7] = 0.f;
  Rt[11] = (szw - sinA * sxc - cosA * szc) / 2;
}

// inlier test of one correspondence (rigidTransform.cu:303-318)
CSB_HD bool csb_rigid_inlier(const float *Rt, const float *p, float thresh2) {
  const float x1 = p[0], y1 = p[1], z1 = p[2], x2 = p[3], y2 = p[4], z2 = p[5];
  const float xt = Rt[0] * x2 + Rt[1] * y2 + Rt[2] * z2 + Rt[3];
  const float yt = Rt[4] * x2 + Rt[5] * y2 + Rt[6] * z2 + Rt[7];
  const float zt = Rt[8] * x2 + Rt[9] * y2 + Rt[10] * z2 + Rt[11];
  const float err = (xt - x1) * (xt - x1) + (yt - y1) * (yt - y1) + (zt - z1) * (zt - z1);
  return err < thresh2;
}

#endif
