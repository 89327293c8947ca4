// cloud_sdf.cuh -- scene side of the pipeline: signed distance / cost of query points with
// respect to a depth point cloud, the job of the reference's DepthPointCloud.get_sdf / get_sdf_cost
// (mesh_to_sdf/depth_point_cloud.py:57-91,127-142: scikit-learn KD-tree query + camera-visibility sign + CHOMP cost).
// Included by gto_b200.cu.
//
// k_cloud_query: exact nearest neighbour by tiled brute force -- the cloud (<= 307 200 points, 4.9 MB as float4) is streamed
// through shared memory in 1024-point tiles, every thread keeps QPT query points in registers (7 fp32 instructions per pair,
// one 128-bit shared load per QPT pairs); FP32-pipe bound, no data structure to build, no divergence.  The sign comes from
// projecting the query into the depth image in float64 (is_outside, :127-142), the cost transform is fused.
#pragma once

#define CLOUD_TILE 1024
#define CLOUD_QPT 4
#define CLOUD_THREADS 256

struct CloudParams {
  const float4* pts;   // [Mpad] (x, y, z, -), padded with far-away points
  int Mpad;
  const double* query; // [N][3]
  long long N;
  const float* depth;  // [H][W]
  int H, W;
  double K[9];         // intrinsics, row-major
  double RT[12];       // inverse camera pose, rows of [R|t]
  int mode;            // 0: signed distance, 1: cost
  float eps, half_eps, two_eps, w_inside;
  float* out;          // [N]
};

__global__ void __launch_bounds__(CLOUD_THREADS) k_cloud_query(const __grid_constant__ CloudParams p) {
  __shared__ float4 tile[CLOUD_TILE];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * (CLOUD_THREADS * CLOUD_QPT);
  float qx[CLOUD_QPT], qy[CLOUD_QPT], qz[CLOUD_QPT], best[CLOUD_QPT];
#pragma unroll
  for (int k = 0; k < CLOUD_QPT; ++k) {
    const long long i = base + (long long)k * CLOUD_THREADS + tid;
    const long long ii = i < p.N ? i : p.N - 1;
    qx[k] = (float)p.query[3 * ii + 0];
    qy[k] = (float)p.query[3 * ii + 1];
    qz[k] = (float)p.query[3 * ii + 2];
    best[k] = 3.0e38f;
  }
  for (int t0 = 0; t0 < p.Mpad; t0 += CLOUD_TILE) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CLOUD_TILE / CLOUD_THREADS; ++j) tile[j * CLOUD_THREADS + tid] = __ldg(p.pts + t0 + j * CLOUD_THREADS + tid);
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < CLOUD_TILE; ++j) {
      const float4 c = tile[j];
#pragma unroll
      for (int k = 0; k < CLOUD_QPT; ++k) {
        const float dx = qx[k] - c.x, dy = qy[k] - c.y, dz = qz[k] - c.z;
        best[k] = fminf(best[k], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CLOUD_QPT; ++k) {
    const long long i = base + (long long)k * CLOUD_THREADS + tid;
    if (i >= p.N) continue;
    float d = sqrtf(best[k]);
    // is_outside (:127-142): project into the depth image; behind the visible surface (or not in the viewport: outside)
    const double x = p.query[3 * i + 0], y = p.query[3 * i + 1], z = p.query[3 * i + 2];
    const double cx = p.RT[0] * x + p.RT[1] * y + p.RT[2] * z + p.RT[3];
    const double cy = p.RT[4] * x + p.RT[5] * y + p.RT[6] * z + p.RT[7];
    const double cz = p.RT[8] * x + p.RT[9] * y + p.RT[10] * z + p.RT[11];
    const double u0 = p.K[0] * cx + p.K[1] * cy + p.K[2] * cz;
    const double u1 = p.K[3] * cx + p.K[4] * cy + p.K[5] * cz;
    const double u2 = p.K[6] * cx + p.K[7] * cy + p.K[8] * cz;
    const double px = u0 / u2, py = u1 / u2;
    bool outside = true;
    if (isfinite(px) && isfinite(py) && fabs(px) < 2.0e9 && fabs(py) < 2.0e9) {
      const long long ix = (long long)px, iy = (long long)py;  // numpy astype(int): truncation toward zero
      if (ix >= 0 && iy >= 0 && ix < p.W && iy < p.H) outside = cz < (double)p.depth[iy * p.W + ix];
    }
    if (!outside) d = -d;
    if (p.mode == 0) {
      p.out[i] = d;
    } else {  // get_sdf_cost (:84-89), float32 arithmetic as NumPy does it
      float c = 0.f;
      if (d < 0.f) c = p.w_inside * (-d + p.half_eps);
      else if (d > 0.f && d < p.eps) { const float t = d - p.eps; c = (t * t) / p.two_eps; }
      p.out[i] = c;
    }
  }
}
