// Backward of the parametric (medium_MANO) tail of the last decoder block — training path, SURVEY §8 rows a16 / f3.
// Forward: mano.cuh (`flat_verts_kernel`, `mano_tail_kernel`; reference pt_metro_transformer.py:139-151,
// lib/utils/transform.py:448-466, manotorch ManoLayer).  One block per sample recomputes the (tiny) forward in shared
// memory and walks it in reverse: output centring -> fingertip / joint selection -> linear-blend skinning -> rest-pose
// removal -> kinematic chain -> pose blend -> joint regressor -> shape blend -> 6-D rotation chain -> mano_linear.
// The 6-D -> matrix -> quaternion -> axis-angle -> (Rodrigues) matrix chain is differentiated with dual numbers
// (96 threads = 16 joints x 6 inputs, one forward-mode pass each): exact, and it follows the same branches (argmax
// quaternion candidate, small-angle series) as the forward.
#pragma once
#include "mano.cuh"
#include "train_simt.cuh"

namespace poem {

struct Dual {
  float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d = 0.f) { return Dual{v, d}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
__device__ __forceinline__ Dual operator+(Dual a, float b) { return {a.v + b, a.d}; }
__device__ __forceinline__ Dual operator-(Dual a, float b) { return {a.v - b, a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, float b) { return {a.v * b, a.d * b}; }
__device__ __forceinline__ Dual operator*(float a, Dual b) { return {a * b.v, a * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, float b) { return {a.v / b, a.d / b}; }
__device__ __forceinline__ Dual operator-(float a, Dual b) { return {a - b.v, -b.d}; }
__device__ __forceinline__ Dual operator+(float a, Dual b) { return {a + b.v, b.d}; }
__device__ __forceinline__ Dual dsqrt(Dual a) {
  const float s = sqrtf(a.v);
  return {s, s > 0.f ? 0.5f * a.d / s : 0.f};
}
__device__ __forceinline__ Dual dmaxc(Dual a, float c) { return a.v >= c ? a : Dual{c, 0.f}; }   // fmaxf(a, c) / clamp_min
__device__ __forceinline__ Dual datan2(Dual y, Dual x) {
  const float den = x.v * x.v + y.v * y.v;
  return {atan2f(y.v, x.v), den > 0.f ? (x.v * y.d - y.v * x.d) / den : 0.f};
}
__device__ __forceinline__ Dual dsin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }
__device__ __forceinline__ Dual dcos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }

// 6-D rotation -> axis-angle (pred_pose) -> rotation matrix as the MANO layer builds it; same arithmetic as mano_tail_kernel.
__device__ __forceinline__ void rot6d_chain_dual(const Dual d6[6], Dual aa[3], Dual R[9]) {
  const Dual a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
  const Dual n1 = dmaxc(dsqrt(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  const Dual b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const Dual dp = b1x * a2x + b1y * a2y + b1z * a2z;
  Dual b2x = a2x - dp * b1x, b2y = a2y - dp * b1y, b2z = a2z - dp * b1z;
  const Dual n2 = dmaxc(dsqrt(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
  b2x = b2x / n2, b2y = b2y / n2, b2z = b2z / n2;
  const Dual b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  const Dual m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z, m20 = b3x, m21 = b3y, m22 = b3z;
  Dual qa[4] = {1.f + m00 + m11 + m22, 1.f + m00 - m11 - m22, 1.f - m00 + m11 - m22, 1.f - m00 - m11 + m22};
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = qa[i].v > 0.f ? dsqrt(qa[i]) : mk(0.f);
  int pick = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (qa[i].v > qa[pick].v) pick = i;
  Dual q[4];
  if (pick == 0) { q[0] = qa[0] * qa[0]; q[1] = m21 - m12; q[2] = m02 - m20; q[3] = m10 - m01; }
  else if (pick == 1) { q[0] = m21 - m12; q[1] = qa[1] * qa[1]; q[2] = m10 + m01; q[3] = m02 + m20; }
  else if (pick == 2) { q[0] = m02 - m20; q[1] = m10 + m01; q[2] = qa[2] * qa[2]; q[3] = m12 + m21; }
  else { q[0] = m10 - m01; q[1] = m20 + m02; q[2] = m21 + m12; q[3] = qa[3] * qa[3]; }
  const Dual den = 2.f * dmaxc(qa[pick], 0.1f);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = q[i] / den;
  const Dual nrm = dsqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const Dual half = datan2(nrm, q[0]);
  const Dual ang = 2.f * half;
  const Dual k = fabsf(ang.v) < 1e-6f ? 0.5f - ang * ang / 48.f : dsin(half) / ang;
  aa[0] = q[1] / k, aa[1] = q[2] / k, aa[2] = q[3] / k;
  const Dual ex = aa[0] + 1e-8f, ey = aa[1] + 1e-8f, ez = aa[2] + 1e-8f;
  const Dual an = dsqrt(ex * ex + ey * ey + ez * ez);
  const Dual ch = dcos(0.5f * an), sh = dsin(0.5f * an);
  Dual w = ch, x = sh * (aa[0] / an), y = sh * (aa[1] / an), z = sh * (aa[2] / an);
  const Dual qn = dsqrt(w * w + x * x + y * y + z * z);
  w = w / qn, x = x / qn, y = y / qn, z = z / qn;
  const Dual w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const Dual wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2, R[1] = 2.f * xy - 2.f * wz, R[2] = 2.f * wy + 2.f * xz;
  R[3] = 2.f * wz + 2.f * xy, R[4] = w2 - x2 + y2 - z2, R[5] = 2.f * yz - 2.f * wx;
  R[6] = 2.f * xz - 2.f * wy, R[7] = 2.f * wx + 2.f * yz, R[8] = w2 - x2 - y2 + z2;
}

struct ManoTailBwdArgs {
  const float *lin_w, *lin_b;                  // mano_linear: [106, D], [106]
  const float *v_template, *shapedirs, *posedirs, *j_regressor, *skin_weights;
  const float* flat;                           // [B, D]   (flat_verts output of the forward)
  const float *dcoords, *dpose, *dshape;       // [B, Q, 3]; [B, 48] / [B, 10] or nullptr
  float *dflat;                                // [B, D]   written
  float *dlin_w, *dlin_b;                      // += (atomics)
  int D, center_idx;
};

__global__ void __launch_bounds__(kManoThreads) mano_tail_bwd_kernel(const ManoTailBwdArgs a) {
  __shared__ float par[106], dpar[106];
  __shared__ float R[kManoJoints][9], dR[kManoJoints][9];
  __shared__ float JacR[kManoJoints][6][9], JacP[kManoJoints][6][3];
  __shared__ float pm[135];
  __shared__ float vp[kManoV3];            // v_shaped, then v_posed
  __shared__ float gv[kManoV3];            // d verts, then d v_posed, then d v_shaped
  __shared__ float J[kManoJoints][3], dJ[kManoJoints][3];
  __shared__ float G[kManoJoints][12], dG[kManoJoints][12];
  __shared__ float A[kManoJoints][12], dA[kManoJoints][12];
  __shared__ float djt[21][3];
  __shared__ float csum[3];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
  const int tips[5] = {745, 317, 444, 556, 673};

  // ================================================================ forward, recomputed (see mano_tail_kernel)
  const float* f = a.flat + (size_t)b * D;
  for (int o = warp; o < 106; o += kManoWarps) {
    const float* w = a.lin_w + (size_t)o * D;
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) acc += f[c] * w[c];
    acc = warp_sum(acc);
    if (lane == 0) par[o] = acc + a.lin_b[o];
  }
  for (int i = tid; i < kManoJoints * 12; i += kManoThreads) (&dA[0][0])[i] = 0.f, (&dG[0][0])[i] = 0.f;
  for (int i = tid; i < kManoJoints * 9; i += kManoThreads) (&dR[0][0])[i] = 0.f;
  for (int i = tid; i < kManoJoints * 3; i += kManoThreads) (&dJ[0][0])[i] = 0.f;
  if (tid < 3) csum[tid] = 0.f;
  __syncthreads();
  if (tid < kManoJoints * 6) {             // joint j, derivative w.r.t. its i-th 6-D input
    const int j = tid / 6, i = tid % 6;
    Dual d6[6], aa[3], Rd[9];
#pragma unroll
    for (int k = 0; k < 6; ++k) d6[k] = mk(par[j * 6 + k], k == i ? 1.f : 0.f);
    rot6d_chain_dual(d6, aa, Rd);
#pragma unroll
    for (int e = 0; e < 9; ++e) JacR[j][i][e] = Rd[e].d;
#pragma unroll
    for (int e = 0; e < 3; ++e) JacP[j][i][e] = aa[e].d;
    if (i == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) R[j][e] = Rd[e].v;
    }
  }
  __syncthreads();
  if (tid < 135) pm[tid] = R[1 + tid / 9][tid % 9] - ((tid % 9) % 4 == 0 ? 1.f : 0.f);
  for (int i = tid; i < kManoV3; i += kManoThreads) {
    float v = a.v_template[i];
#pragma unroll
    for (int k = 0; k < 10; ++k) v += a.shapedirs[k * kManoV3 + i] * par[96 + k];
    vp[i] = v;
  }
  __syncthreads();
  for (int o = warp; o < kManoJoints * 3; o += kManoWarps) {
    const int j = o / 3, c = o % 3;
    float acc = 0.f;
    for (int v = lane; v < kManoVerts; v += 32) acc += a.j_regressor[j * kManoVerts + v] * vp[v * 3 + c];
    acc = warp_sum(acc);
    if (lane == 0) J[j][c] = acc;
  }
  __syncthreads();
  for (int i = tid; i < kManoV3; i += kManoThreads) {
    float v = 0.f;
#pragma unroll 15
    for (int k = 0; k < 135; ++k) v += a.posedirs[k * kManoV3 + i] * pm[k];
    vp[i] += v;                              // v_posed
  }
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      G[0][r * 4 + 0] = R[0][r * 3 + 0], G[0][r * 4 + 1] = R[0][r * 3 + 1], G[0][r * 4 + 2] = R[0][r * 3 + 2];
      G[0][r * 4 + 3] = J[0][r];
    }
  }
  __syncthreads();
  if (tid < 5) {
    for (int s = 0; s < 3; ++s) {
      const int j = 1 + tid * 3 + s, p = s == 0 ? 0 : j - 1;
      const float tx = J[j][0] - J[p][0], ty = J[j][1] - J[p][1], tz = J[j][2] - J[p][2];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float g0 = G[p][r * 4 + 0], g1 = G[p][r * 4 + 1], g2 = G[p][r * 4 + 2], g3 = G[p][r * 4 + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[j][r * 4 + c] = g0 * R[j][c] + g1 * R[j][3 + c] + g2 * R[j][6 + c];
        G[j][r * 4 + 3] = g0 * tx + g1 * ty + g2 * tz + g3;
      }
    }
  }
  __syncthreads();
  if (tid < kManoJoints) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float g0 = G[tid][r * 4 + 0], g1 = G[tid][r * 4 + 1], g2 = G[tid][r * 4 + 2];
      A[tid][r * 4 + 0] = g0, A[tid][r * 4 + 1] = g1, A[tid][r * 4 + 2] = g2;
      A[tid][r * 4 + 3] = G[tid][r * 4 + 3] - (g0 * J[tid][0] + g1 * J[tid][1] + g2 * J[tid][2]);
    }
  }

  // ================================================================ backward
  // ---- outputs: out = (x - jt[centre]) + hand centre  ->  d x = d out, d jt[centre] -= sum of every d out
  const float* dc = a.dcoords + (size_t)b * (21 + kManoVerts) * 3;
  {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int i = tid; i < (21 + kManoVerts) * 3; i += kManoThreads) {
      const float g = dc[i];
      if (i < 63) djt[i / 3][i % 3] = g;
      else gv[i - 63] = g;
      const int c = i % 3;
      s0 += c == 0 ? g : 0.f, s1 += c == 1 ? g : 0.f, s2 += c == 2 ? g : 0.f;
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1), s2 = warp_sum(s2);
    if (lane == 0) atomicAdd(&csum[0], s0), atomicAdd(&csum[1], s1), atomicAdd(&csum[2], s2);
  }
  __syncthreads();
  if (tid < 3 && a.center_idx >= 0) djt[a.center_idx][tid] -= csum[tid];
  __syncthreads();
  // ---- 21-joint selection: joints <- chain translations, tips <- vertices
  if (tid < 21) {
    const int s = order[tid];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (s < 16) dG[s][c * 4 + 3] += djt[tid][c];       // distinct s per thread
      else gv[tips[s - 16] * 3 + c] += djt[tid][c];
    }
  }
  __syncthreads();
  // ---- skinning: verts_v = T_v [p_v; 1], T_v = sum_j w_vj A_j
  for (int v = tid; v < kManoVerts; v += kManoThreads) {
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < kManoJoints; ++j) {
      const float wj = a.skin_weights[v * kManoJoints + j];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] += wj * A[j][e];
    }
    const float px = vp[v * 3], py = vp[v * 3 + 1], pz = vp[v * 3 + 2];
    const float g0 = gv[v * 3], g1 = gv[v * 3 + 1], g2 = gv[v * 3 + 2];
    const float dT[12] = {g0 * px, g0 * py, g0 * pz, g0, g1 * px, g1 * py, g1 * pz, g1, g2 * px, g2 * py, g2 * pz, g2};
    for (int j = 0; j < kManoJoints; ++j) {
      const float wj = a.skin_weights[v * kManoJoints + j];
      if (wj != 0.f) {
#pragma unroll
        for (int e = 0; e < 12; ++e) atomicAdd(&dA[j][e], wj * dT[e]);
      }
    }
    gv[v * 3] = T[0] * g0 + T[4] * g1 + T[8] * g2;          // d v_posed = T^R^T g
    gv[v * 3 + 1] = T[1] * g0 + T[5] * g1 + T[9] * g2;
    gv[v * 3 + 2] = T[2] * g0 + T[6] * g1 + T[10] * g2;
  }
  __syncthreads();
  // ---- A_j = [G_j^R | G_j^t - G_j^R J_j]
  if (tid < kManoJoints) {
    const int j = tid;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float dt = dA[j][r * 4 + 3];
#pragma unroll
      for (int c = 0; c < 3; ++c) dG[j][r * 4 + c] += dA[j][r * 4 + c] - dt * J[j][c];
      dG[j][r * 4 + 3] += dt;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      dJ[j][c] -= G[j][0 * 4 + c] * dA[j][3] + G[j][1 * 4 + c] * dA[j][7] + G[j][2 * 4 + c] * dA[j][11];
  }
  __syncthreads();
  // ---- kinematic chain in reverse: thread per finger, leaf to knuckle; the root collects from the five fingers
  if (tid < 5) {
    for (int s = 2; s >= 0; --s) {
      const int j = 1 + tid * 3 + s, p = s == 0 ? 0 : j - 1;
      const float t[3] = {J[j][0] - J[p][0], J[j][1] - J[p][1], J[j][2] - J[p][2]};
      float dt[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float gt = dG[j][r * 4 + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // d G_p^R[r][c] += sum_k dG_j^R[r][k] R_j[c][k] + dG_j^t[r] t[c]
          const float add = dG[j][r * 4 + 0] * R[j][c * 3 + 0] + dG[j][r * 4 + 1] * R[j][c * 3 + 1] +
                            dG[j][r * 4 + 2] * R[j][c * 3 + 2] + gt * t[c];
          if (p == 0) atomicAdd(&dG[0][r * 4 + c], add);
          else dG[p][r * 4 + c] += add;
          // d R_j[c][k] += sum_r G_p^R[r][c] dG_j^R[r][k]     (accumulated over r below)
          dt[c] += G[p][r * 4 + c] * gt;
        }
        if (p == 0) atomicAdd(&dG[0][r * 4 + 3], gt);
        else dG[p][r * 4 + 3] += gt;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          dR[j][c * 3 + k] += G[p][0 * 4 + c] * dG[j][0 * 4 + k] + G[p][1 * 4 + c] * dG[j][1 * 4 + k] + G[p][2 * 4 + c] * dG[j][2 * 4 + k];
        dJ[j][c] += dt[c];
        if (p == 0) atomicAdd(&dJ[0][c], -dt[c]);
        else dJ[p][c] -= dt[c];
      }
    }
  }
  __syncthreads();
  if (tid < 9) dR[0][tid] += dG[0][(tid / 3) * 4 + (tid % 3)];
  if (tid >= 32 && tid < 35) dJ[0][tid - 32] += dG[0][(tid - 32) * 4 + 3];
  __syncthreads();
  // ---- pose blend: v_posed = v_shaped + posedirs . pm  ->  d pm_k = <posedirs[k], d v_posed>, d R[1 + k/9][k%9] += d pm_k
  for (int k = warp; k < 135; k += kManoWarps) {
    float acc = 0.f;
    for (int i = lane; i < kManoV3; i += 32) acc += a.posedirs[k * kManoV3 + i] * gv[i];
    acc = warp_sum(acc);
    if (lane == 0) dR[1 + k / 9][k % 9] += acc;
  }
  __syncthreads();
  // ---- joints: J = J_regressor . v_shaped  ->  d v_shaped = d v_posed + J_regressor^T dJ
  for (int i = tid; i < kManoV3; i += kManoThreads) {
    const int v = i / 3, c = i % 3;
    float acc = gv[i];
#pragma unroll
    for (int j = 0; j < kManoJoints; ++j) acc += a.j_regressor[j * kManoVerts + v] * dJ[j][c];
    gv[i] = acc;
  }
  __syncthreads();
  // ---- shape blend -> d betas ; rotation chain -> d 6-D inputs
  for (int k = warp; k < 10; k += kManoWarps) {
    float acc = 0.f;
    for (int i = lane; i < kManoV3; i += 32) acc += a.shapedirs[k * kManoV3 + i] * gv[i];
    acc = warp_sum(acc);
    if (lane == 0) dpar[96 + k] = acc + (a.dshape ? a.dshape[(size_t)b * 10 + k] : 0.f);
  }
  if (tid >= 512 && tid < 512 + 96) {
    const int j = (tid - 512) / 6, i = (tid - 512) % 6;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) acc += JacR[j][i][e] * dR[j][e];
    if (a.dpose) {
#pragma unroll
      for (int e = 0; e < 3; ++e) acc += JacP[j][i][e] * a.dpose[(size_t)b * 48 + j * 3 + e];
    }
    dpar[j * 6 + i] = acc;
  }
  __syncthreads();
  // ---- mano_linear: par = W flat + bias
  for (int c = tid; c < D; c += kManoThreads) {
    float acc = 0.f;
    const float fc = f[c];
    for (int o = 0; o < 106; ++o) {
      acc += dpar[o] * a.lin_w[(size_t)o * D + c];
      atomicAdd(a.dlin_w + (size_t)o * D + c, dpar[o] * fc);
    }
    a.dflat[(size_t)b * D + c] = acc;
  }
  if (tid < 106) atomicAdd(a.dlin_b + tid, dpar[tid]);
}

// flat[r] = <x[r, :], w> + b  (x = the features re-interpreted as (rows, Q)):  dx[r, j] = dflat[r] w[j] ;
// dw[j] += sum_r dflat[r] x[r, j] ; db += sum_r dflat[r].     grid (ceil(Q / 32), row slabs), block (32, 8)
__global__ void tr_flat_verts_bwd_kernel(const float* dflat, const float* x, const float* w, float* dx, float* dw, float* db,
                                         int Q, int rows) {
  __shared__ float part[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f, bsum = 0.f;
  const float wj = j < Q ? w[j] : 0.f;
  for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) {
    const float g = dflat[r];
    if (j < Q) {
      acc += g * x[(size_t)r * Q + j];
      dx[(size_t)r * Q + j] = g * wj;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) bsum += g;
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < Q) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
    atomicAdd(dw + j, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(db, bsum);
}

}  // namespace poem
