// cloud_sdf.cuh -- scene side of the pipeline: signed distance / cost of query points with
// respect to a depth point cloud, the job of the reference's DepthPointCloud.get_sdf / get_sdf_cost
// (mesh_to_sdf/depth_point_cloud.py:57-91,127-142: scikit-learn KD-tree query + camera-visibility sign + CHOMP cost).
// Included by gto_b200.cu.
//
// k_cloud_query: exact nearest neighbour by tiled brute force -- the cloud (<= 307 200 points, 4.9 MB as float4) is streamed
// through shared memory in 1024-point tiles, every thread keeps QPT query points in registers (7 fp32 instructions per pair,
// one 128-bit shared load per QPT pairs); FP32-pipe bound, no data structure to build, no divergence.  The sign comes from
// projecting the query into the depth image in float64 (is_outside, :127-142), the cost transform is fused.
//
// k_cloud_query_pruned (default): the same exact search with pruning.  gto_cloud_set sorts the cloud along a Morton curve and
// cuts it into 256-point tiles with bounding boxes; a warp owns 128 consecutive queries (4 per lane: a grid line in the usual
// z-fastest layout), scans the tile boxes in passes of doubling radius and only visits tiles whose box is closer than the
// worst current nearest-neighbour distance of its queries.  Every tile that could hold a nearer point is visited, so the
// result is bit-identical to the brute-force kernel (tests/test_gpu_cloud.py) at a fraction of the pair evaluations.
#pragma once

#define CLOUD_TILE 1024
#define CLOUD_PTILE 256      // points per pruning tile
#define CLOUD_WQ 128         // queries per warp in the pruned kernel
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
  const float* tiles;  // [ntiles][6] bounding boxes (lo xyz, hi xyz) of the CLOUD_PTILE-point tiles (pruned kernel)
  int ntiles;
};

// is_outside (mesh_to_sdf/depth_point_cloud.py:127-142): project into the depth image; hidden behind the visible surface -> inside
__device__ __forceinline__ bool cloud_outside(const CloudParams& p, long long i) {
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
  return outside;
}

__device__ __forceinline__ float cloud_value(const CloudParams& p, float d2, bool outside) {
  float d = sqrtf(d2);
  if (!outside) d = -d;
  if (p.mode == 0) return d;
  float c = 0.f;  // get_sdf_cost (:84-89), float32 arithmetic as NumPy does it
  if (d < 0.f) c = p.w_inside * (-d + p.half_eps);
  else if (d > 0.f && d < p.eps) { const float t = d - p.eps; c = (t * t) / p.two_eps; }
  return c;
}

// mode 2: the visibility test alone (DepthPointCloud.is_outside, :127-142), 1.0 = outside / visible, 0.0 = hidden
__global__ void k_cloud_outside(const __grid_constant__ CloudParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.N) p.out[i] = cloud_outside(p, i) ? 1.f : 0.f;
}

// Back-projection of a depth image into a world-frame point cloud (DepthPointCloud.__init__ / backproject_camera, :15-19,32-52):
// pixel (u, v) with 0 < depth < threshold (and outside the target mask) -> X = depth * Kinv [u v 1]^T -> world = R X + t, float64.
// One thread per pixel; `valid` marks the pixels that pass the test (the caller compacts, keeping the row-major pixel order).
struct BackprojParams {
  const float* depth;          // [H][W]
  const unsigned char* mask;   // [H][W] non-zero: pixel belongs to the target object (excluded), or NULL
  int H, W;
  double Kinv[9], pose[12];    // inverse intrinsics (row-major), camera pose rows of [R|t]
  float threshold;
  double* points;              // [H*W][3]
  unsigned char* valid;        // [H*W]
};
__global__ void k_cloud_backproject(const __grid_constant__ BackprojParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.H * p.W) return;
  const int v = (int)(i / p.W), u = (int)(i - (long long)v * p.W);
  const float d = p.depth[i];
  const bool ok = d > 0.f && d < p.threshold && (p.mask == nullptr || p.mask[i] == 0);
  const double dd = (double)d, uu = (double)u, vv = (double)v;
  const double x = dd * (p.Kinv[0] * uu + p.Kinv[1] * vv + p.Kinv[2]);
  const double y = dd * (p.Kinv[3] * uu + p.Kinv[4] * vv + p.Kinv[5]);
  const double z = dd * (p.Kinv[6] * uu + p.Kinv[7] * vv + p.Kinv[8]);
  p.points[3 * i + 0] = p.pose[0] * x + p.pose[1] * y + p.pose[2] * z + p.pose[3];
  p.points[3 * i + 1] = p.pose[4] * x + p.pose[5] * y + p.pose[6] * z + p.pose[7];
  p.points[3 * i + 2] = p.pose[8] * x + p.pose[9] * y + p.pose[10] * z + p.pose[11];
  p.valid[i] = ok ? 1 : 0;
}

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
    p.out[i] = cloud_value(p, best[k], cloud_outside(p, i));
  }
}

__global__ void __launch_bounds__(128) k_cloud_query_pruned(const __grid_constant__ CloudParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long qbase = warp * CLOUD_WQ;
  if (qbase >= p.N) return;
  float qx[4], qy[4], qz[4], best[4];
  bool outside[4], valid[4];
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = qbase + lane * 4 + k;
    valid[k] = i < p.N;
    const long long ii = valid[k] ? i : p.N - 1;
    qx[k] = (float)p.query[3 * ii + 0];
    qy[k] = (float)p.query[3 * ii + 1];
    qz[k] = (float)p.query[3 * ii + 2];
    outside[k] = cloud_outside(p, ii);
    best[k] = 3.0e38f;
    lo[0] = fminf(lo[0], qx[k]); hi[0] = fmaxf(hi[0], qx[k]);
    lo[1] = fminf(lo[1], qy[k]); hi[1] = fmaxf(hi[1], qy[k]);
    lo[2] = fminf(lo[2], qz[k]); hi[2] = fmaxf(hi[2], qz[k]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  const float eps2 = p.eps * p.eps;
  // the largest distance any query of this warp still has to beat (cost mode: a visible query at >= eps has cost 0 whatever its
  // true distance is)
  auto warp_worst = [&]() {
    float w = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float e = best[k];
      if (p.mode == 1 && outside[k]) e = fminf(e, eps2);
      if (valid[k]) w = fmaxf(w, e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, o));
    return w;
  };
  float worst = 3.0e38f;
  float Rprev2 = -1.f;
  for (int pass = 0;; ++pass) {
    const float R = (pass < 7) ? 0.05f * (float)(1 << pass) : 3.0e18f;
    const float R2 = R * R;
    for (int t0 = 0; t0 < p.ntiles; t0 += 32) {
      const int tl = t0 + lane;
      float lb2 = 3.0e38f;
      if (tl < p.ntiles) {  // squared distance between the query box and the tile box
        const float* tb = p.tiles + 6 * (size_t)tl;
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float g = fmaxf(0.f, fmaxf(__ldg(tb + a) - hi[a], lo[a] - __ldg(tb + 3 + a)));
          s = fmaf(g, g, s);
        }
        lb2 = s * 0.9999f;  // conservative under rounding
      }
      unsigned m = __ballot_sync(0xffffffffu, lb2 <= R2 && lb2 > Rprev2);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float lbt = __shfl_sync(0xffffffffu, lb2, src);
        if (lbt > worst) continue;  // no point of this tile can be nearer than what every query already has
        const float4* tp = p.pts + (size_t)(t0 + src) * CLOUD_PTILE;
#pragma unroll 8
        for (int j = 0; j < CLOUD_PTILE; ++j) {
          const float4 c = __ldg(tp + j);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float dx = qx[k] - c.x, dy = qy[k] - c.y, dz = qz[k] - c.z;
            best[k] = fminf(best[k], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
          }
        }
        worst = warp_worst();
      }
    }
    if (worst <= R2 || pass >= 7) break;  // every tile not visited yet is farther than R >= all current distances
    Rprev2 = R2;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long i = qbase + lane * 4 + k;
    if (valid[k]) p.out[i] = cloud_value(p, best[k], outside[k]);
  }
}
