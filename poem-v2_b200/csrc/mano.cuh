// Parametric (medium_MANO) tail of the last decoder block, SURVEY §8a row a16:
// `point_METRO_block.get_parametric_output` (reference lib/models/bricks/pt_metro_transformer.py:139-151),
// `rot6d_to_aa` (lib/utils/transform.py:448-466 = pytorch3d rotation_6d_to_matrix -> matrix_to_quaternion ->
// quaternion_to_axis_angle) and the manotorch `ManoLayer` forward (axis-angle, no PCA, flat hand mean, root-centred).
// Everything is fp32 like the reference; the work is tiny (0.3 MFLOP + 26 MB of reads per 32 samples), so the design
// goal is simply one pass over the features and one block per sample for the rest.
#pragma once
#include "common.cuh"

namespace poem {

// ------------------------------------------------------------------------------------------------
// flat[r] = <feats_flat[r*Q : (r+1)*Q], w> + b   for r < B*D.
// The reference RE-INTERPRETS the contiguous (B,Q,D) features as (B*D, Q) rows (`reshape(-1, 799)`, no transpose),
// so row r is simply the r-th run of Q consecutive floats.  Warp per row, coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void flat_verts_kernel(const float* __restrict__ feats, const float* __restrict__ w,
                                  const float* __restrict__ b, float* __restrict__ flat, int Q, int rows) {
  pdl_wait();
  pdl_trigger();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = feats + (size_t)r * Q;
  float acc = 0.f;
  for (int j = lane; j < Q; j += 32) acc += x[j] * w[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) flat[r] = acc + b[0];
}

struct ManoTailArgs {
  const float *lin_w, *lin_b;                  // mano_linear: [106, D], [106]
  const float *v_template, *shapedirs, *posedirs, *j_regressor, *skin_weights;
  const float* flat;                           // [B, D]
  const float* ref_joints;                     // [B, 21, 3] or nullptr (normalised / transformer-only call)
  float *coords, *pose_out, *shape_out;        // [B, Q, 3], [B, 48], [B, 10]
  int D, center_idx;
};

__device__ __forceinline__ float nan_to_num_f(float v) {   // torch.nan_to_num (ptEmb_head.py:944)
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

constexpr int kManoVerts = 778, kManoJoints = 16, kManoV3 = kManoVerts * 3;
constexpr int kManoThreads = 1024, kManoWarps = kManoThreads / 32;   // one block per sample: wide, so the blend loops keep many loads in flight

// One block per sample.
__global__ void __launch_bounds__(kManoThreads) mano_tail_kernel(const ManoTailArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float par[106];
  __shared__ float R[kManoJoints][9];
  __shared__ float pm[135];
  __shared__ float vs[kManoV3];
  __shared__ float J[kManoJoints][3];
  __shared__ float G[kManoJoints][12];   // 3x4 rows of the global joint transforms
  __shared__ float A[kManoJoints][12];   // same with the rest pose removed
  __shared__ float jt[21][3];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;

  // ---- mano_linear: (D) -> 106 = 16 x 6-D rotations | 10 betas
  const float* f = a.flat + (size_t)b * D;
  for (int o = warp; o < 106; o += kManoWarps) {
    const float* w = a.lin_w + (size_t)o * D;
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) acc += f[c] * w[c];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) par[o] = acc + a.lin_b[o];
  }
  __syncthreads();

  // ---- 6-D rotation -> matrix -> quaternion -> axis-angle (pred_pose), then axis-angle -> matrix as the MANO layer does
  if (tid < kManoJoints) {
    const float* d6 = par + tid * 6;
    float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
    float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);   // F.normalize eps
    const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
    const float dp = b1x * a2x + b1y * a2y + b1z * a2z;
    float b2x = a2x - dp * b1x, b2y = a2y - dp * b1y, b2z = a2z - dp * b1z;
    const float n2 = fmaxf(sqrtf(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
    b2x /= n2, b2y /= n2, b2z /= n2;
    const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
    // rows (b1, b2, b3)
    const float m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z, m20 = b3x, m21 = b3y, m22 = b3z;
    float qa[4] = {1.f + m00 + m11 + m22, 1.f + m00 - m11 - m22, 1.f - m00 + m11 - m22, 1.f - m00 - m11 + m22};
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
    int pick = 0;   // argmax, first maximum wins like torch.argmax
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (qa[i] > qa[pick]) pick = i;
    float q[4];
    if (pick == 0) { q[0] = qa[0] * qa[0]; q[1] = m21 - m12; q[2] = m02 - m20; q[3] = m10 - m01; }
    else if (pick == 1) { q[0] = m21 - m12; q[1] = qa[1] * qa[1]; q[2] = m10 + m01; q[3] = m02 + m20; }
    else if (pick == 2) { q[0] = m02 - m20; q[1] = m10 + m01; q[2] = qa[2] * qa[2]; q[3] = m12 + m21; }
    else { q[0] = m10 - m01; q[1] = m20 + m02; q[2] = m21 + m12; q[3] = qa[3] * qa[3]; }
    const float den = 2.f * fmaxf(qa[pick], 0.1f);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] /= den;
    const float nrm = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float half = atan2f(nrm, q[0]);
    const float ang = 2.f * half;
    const float k = fabsf(ang) < 1e-6f ? 0.5f - ang * ang / 48.f : sinf(half) / ang;
    const float ax = q[1] / k, ay = q[2] / k, az = q[3] / k;
    float* po = a.pose_out + (size_t)b * 48 + tid * 3;
    po[0] = ax, po[1] = ay, po[2] = az;
    // manotorch: Rodrigues through a unit quaternion
    const float ex = ax + 1e-8f, ey = ay + 1e-8f, ez = az + 1e-8f;
    const float an = sqrtf(ex * ex + ey * ey + ez * ez);
    const float ch = cosf(0.5f * an), sh = sinf(0.5f * an);
    float w = ch, x = sh * (ax / an), y = sh * (ay / an), z = sh * (az / an);
    const float qn = sqrtf(w * w + x * x + y * y + z * z);
    w /= qn, x /= qn, y /= qn, z /= qn;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    float* r = R[tid];
    r[0] = w2 + x2 - y2 - z2, r[1] = 2 * xy - 2 * wz, r[2] = 2 * wy + 2 * xz;
    r[3] = 2 * wz + 2 * xy, r[4] = w2 - x2 + y2 - z2, r[5] = 2 * yz - 2 * wx;
    r[6] = 2 * xz - 2 * wy, r[7] = 2 * wx + 2 * yz, r[8] = w2 - x2 - y2 + z2;
  } else if (tid >= 32 && tid < 42) {
    a.shape_out[(size_t)b * 10 + (tid - 32)] = par[96 + tid - 32];
  }
  __syncthreads();
  if (tid < 135) pm[tid] = R[1 + tid / 9][tid % 9] - ((tid % 9) % 4 == 0 ? 1.f : 0.f);

  // ---- shape blend: v_shaped = v_template + shapedirs . betas   (shapedirs stored [10][778*3])
  for (int i = tid; i < kManoV3; i += kManoThreads) {
    float v = a.v_template[i];
#pragma unroll
    for (int k = 0; k < 10; ++k) v += a.shapedirs[k * kManoV3 + i] * par[96 + k];
    vs[i] = v;
  }
  __syncthreads();
  // ---- joints of the shaped mesh: J = J_regressor . v_shaped
  for (int o = warp; o < kManoJoints * 3; o += kManoWarps) {
    const int j = o / 3, c = o % 3;
    float acc = 0.f;
    for (int v = lane; v < kManoVerts; v += 32) acc += a.j_regressor[j * kManoVerts + v] * vs[v * 3 + c];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) J[j][c] = acc;
  }
  __syncthreads();
  // ---- pose blend: v_posed = v_shaped + posedirs . (R[1:] - I)   (posedirs stored [135][778*3])
  for (int i = tid; i < kManoV3; i += kManoThreads) {
    float v = 0.f;
#pragma unroll 15
    for (int k = 0; k < 135; ++k) v += a.posedirs[k * kManoV3 + i] * pm[k];   // 15 independent L2 loads in flight
    vs[i] += v;
  }
  // ---- kinematic chain (MANO tree: joint j hangs on j-1, except 1,4,7,10,13 which hang on the root): thread per finger
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
  __syncthreads();
  // ---- skinning: thread per vertex, in place
  for (int v = tid; v < kManoVerts; v += kManoThreads) {
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < kManoJoints; ++j) {
      const float wj = a.skin_weights[v * kManoJoints + j];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] += wj * A[j][e];
    }
    const float px = vs[v * 3], py = vs[v * 3 + 1], pz = vs[v * 3 + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) vs[v * 3 + r] = T[r * 4] * px + T[r * 4 + 1] * py + T[r * 4 + 2] * pz + T[r * 4 + 3];
  }
  __syncthreads();
  // ---- 16 joints + 5 fingertip vertices in the 21-joint order
  if (tid < 21) {
    const int order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
    const int tips[5] = {745, 317, 444, 556, 673};
    const int s = order[tid];
#pragma unroll
    for (int c = 0; c < 3; ++c) jt[tid][c] = s < 16 ? G[s][c * 4 + 3] : vs[tips[s - 16] * 3 + c];
  }
  __syncthreads();
  // ---- root-centre on joint `center_idx`, add the sample's hand centre (ptEmb_head.py:957-958), write joints | verts
  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (a.center_idx >= 0) cx = jt[a.center_idx][0], cy = jt[a.center_idx][1], cz = jt[a.center_idx][2];
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (a.ref_joints) {
    const float* c = a.ref_joints + ((size_t)b * 21 + 9) * 3;   // hand centre: always joint 9 (ptEmb_head.py:873,958)
    ox = c[0], oy = c[1], oz = c[2];
  }
  float* out = a.coords + (size_t)b * (21 + kManoVerts) * 3;
  for (int i = tid; i < (21 + kManoVerts) * 3; i += kManoThreads) {
    const int c = i % 3;
    const float v = i < 63 ? jt[i / 3][c] : vs[i - 63];
    out[i] = nan_to_num_f(v - (c == 0 ? cx : c == 1 ? cy : cz)) + (c == 0 ? ox : c == 1 ? oy : oz);
  }
}

}  // namespace poem
