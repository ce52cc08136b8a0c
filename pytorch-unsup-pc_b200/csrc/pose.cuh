// Pose math shared by the forward (pose + scatter) and backward (gather + pose
// adjoint) kernels.  Restates quaternion.py:110-132 (quaternion_rotate) and
// point_cloud_to.py:118-178 (pc_perspective_transform, quaternion branch).
//
// The forward mirrors the reference's *mixed precision* operation by
// operation: quaternion normalise and the first Hamilton product q^ (x) (0,p)
// in fp32 with every product/sum rounded separately (torch eager does not
// contract into FMAs), then fp64 for the second product and everything after
// it (quaternion_conjugate multiplies by a float64 array, quaternion.py:91-93).
// Keeping the same roundings keeps every point in the same voxel cell as the
// reference (hard part 1).  The ~60 fp64 flops per point
// are negligible next to the grid traffic.
#pragma once
#include <cuda_runtime.h>

namespace dpc {

struct Quat {
  float w, x, y, z;   // normalised, fp32 (as the reference holds it)
  float inv_norm;     // 1/|q| for the normalisation Jacobian
};

// q / ||q||: torch CPU computes the fp32 norm as a sequential fp32 sum of
// squares with the square root taken in double (verified bit-exact against
// torch 2.11), then an IEEE fp32 divide (quaternion.py:119-121).
__device__ __forceinline__ Quat load_quat(const float *__restrict__ q) {
  float q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)),
                                __fmul_rn(q2, q2)), __fmul_rn(q3, q3));
  float n = (float)sqrt((double)s);
  Quat r;
  r.w = __fdiv_rn(q0, n);
  r.x = __fdiv_rn(q1, n);
  r.y = __fdiv_rn(q2, n);
  r.z = __fdiv_rn(q3, n);
  r.inv_norm = 1.f / n;
  return r;
}

struct PosePoint {
  double r0, r1, r2;  // rotated (+translated) point p'
  double zc;          // p'_0 + camera_distance
  double u0, u1, u2;  // tr_pc in (z, y, x) order
};

__device__ __forceinline__ PosePoint pose_point(const Quat &q, float p0, float p1, float p2,
                                                bool has_t, float t0, float t1, float t2,
                                                double f, double cam_dist) {
  // first product, fp32, quaternion.py:80-85 with b = (0, p)
  float aw = __fsub_rn(__fsub_rn(-__fmul_rn(q.x, p0), __fmul_rn(q.y, p1)), __fmul_rn(q.z, p2));
  float ax = __fsub_rn(__fadd_rn(__fmul_rn(q.w, p0), __fmul_rn(q.y, p2)), __fmul_rn(q.z, p1));
  float ay = __fsub_rn(__fadd_rn(__fmul_rn(q.w, p1), __fmul_rn(q.z, p0)), __fmul_rn(q.x, p2));
  float az = __fsub_rn(__fadd_rn(__fmul_rn(q.w, p2), __fmul_rn(q.x, p1)), __fmul_rn(q.y, p0));
  // second product with the fp64 conjugate
  double Aw = aw, Ax = ax, Ay = ay, Az = az;
  double bw = q.w, bx = -(double)q.x, by = -(double)q.y, bz = -(double)q.z;
  PosePoint o;
  o.r0 = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(Aw, bx), __dmul_rn(Ax, bw)), __dmul_rn(Ay, bz)),
                   __dmul_rn(Az, by));
  o.r1 = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(Aw, by), __dmul_rn(Ay, bw)), __dmul_rn(Az, bx)),
                   __dmul_rn(Ax, bz));
  o.r2 = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(Aw, bz), __dmul_rn(Az, bw)), __dmul_rn(Ax, by)),
                   __dmul_rn(Ay, bx));
  if (has_t) {  // point_cloud_to.py:137-139
    o.r0 = __dadd_rn(o.r0, (double)t0);
    o.r1 = __dadd_rn(o.r1, (double)t1);
    o.r2 = __dadd_rn(o.r2, (double)t2);
  }
  // point_cloud_to.py:145-148, 169-175
  o.zc = __dadd_rn(o.r0, cam_dist);
  o.u1 = __ddiv_rn(__dmul_rn(o.r1, f), o.zc);
  o.u2 = __ddiv_rn(__dmul_rn(o.r2, f), o.zc);
  o.u0 = __dsub_rn(o.zc, cam_dist);
  if (has_t) o.u0 = __dsub_rn(o.u0, (double)t0);
  return o;
}

// Trilinear cell of one point (point_cloud_to.py:25-40).
struct Cell {
  bool valid;
  int iz, iy, ix;
  double rz, ry, rx;
};

__device__ __forceinline__ Cell make_cell(double u0, double u1, double u2, int Vz, int V) {
  Cell c;
  c.valid = (u0 >= -0.5) && (u0 <= 0.5) && (u1 >= -0.5) && (u1 <= 0.5) && (u2 >= -0.5) &&
            (u2 <= 0.5);
  double gz = __dmul_rn(__dadd_rn(u0, 0.5), (double)(Vz - 1));
  double gy = __dmul_rn(__dadd_rn(u1, 0.5), (double)(V - 1));
  double gx = __dmul_rn(__dadd_rn(u2, 0.5), (double)(V - 1));
  double fz = floor(gz), fy = floor(gy), fx = floor(gx);
  c.rz = gz - fz;
  c.ry = gy - fy;
  c.rx = gx - fx;
  c.iz = (int)fz;
  c.iy = (int)fy;
  c.ix = (int)fx;
  return c;
}

}  // namespace dpc
