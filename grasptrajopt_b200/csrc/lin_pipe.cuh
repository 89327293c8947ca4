// lin_pipe.cuh -- k_linearize_pipe: warp-specialised, TMA-pipelined version of the fused linearisation kernel.
// Included by gto_b200.cu (uses its device tables and PTX helpers).
//
// One persistent CTA = 1 producer warp + NC consumer warps, looping over (problem, knot) work items:
//   producer   float64 chain FK of the item -> item context (link frames, joint twists, goal-frame differences, brick
//              placement) in a double-buffered shared-memory slot; then one TMA tile load per link of the SDF brick that
//              encloses the link's point set at this knot into a ring of NSLOT shared-memory slots (full/empty mbarriers).
//              The producer runs ahead of the consumers by up to NSLOT bricks and one item.
//   consumers  deal the item's 32-point chunks round-robin (chunks are ordered by link, so a warp walks the ring in
//              order): point -> base frame -> brick-local voxel coordinates -> 8 shared-memory taps -> trilinear value and
//              gradient -> Jacobian row (twist form) -> staged in shared memory -> TF32 tensor-core J^T J + fp32 J^T r ->
//              coalesced 128-bit row stores to HBM.  One named barrier per item for the cross-warp reduction.
// Box sizes per axis come from {8,12,...,32} (one tensor map per combination and field); the start coordinate of the
// innermost (z) axis is kept a multiple of 4 elements: TMA faults on a tile whose innermost start is not 16-byte aligned.
#pragma once

#define PIPE_NSLOT 4
#define PIPE_NAXC 7  // per-axis box sizes 8,12,...,32
#define PIPE_MAX_CONS 8

struct ItemCtx {
  float frames[GTO_MAX_LINKS][12];  // visual frame of each collision link (robot base frame)
  float tw[GTO_MAX_OPT][8];         // (omega.xyz, -, m.xyz, -) per optimised joint
  float gripf[12];
  float goal[2][12];                // gripper frame minus goal / stand-off frame
  float cl[GTO_MAX_LINKS][4];       // brick-local voxel coordinate = Wb * inv_pitch + cl
  int blo[GTO_MAX_LINKS][4];        // brick lower corner (grid index) per link
  int bdim[GTO_MAX_LINKS][4];       // brick dims (x, y, z) and fast flag
  float basep[4];
  int b, t, fid, obuf;
  int part, l0, l1, last_part;
};

struct LinkMeta {
  int c0, c1, pt_start, pt_end;
  unsigned mask;
};

struct PipeShared {
  double Tm[GTO_MAX_MOV][12];
  double A[GTO_MAX_MOV][12];
  ItemCtx ctx[2];
  LinkMeta links[GTO_MAX_LINKS];
  unsigned long long slot_full[PIPE_NSLOT], slot_empty[PIPE_NSLOT], ctx_full[2], ctx_empty[2];
};

// try_wait with a suspend-time hint: the hardware parks the thread instead of burning issue slots on polling
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int tries = 0; tries < (1 << 20); ++tries) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) return;
  }
  __trap();  // a lost transaction must abort the launch, never hang the device
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic (slow-path) trilinear lookup with a non-cubic brick; same arithmetic as sdf_trilinear
__device__ __forceinline__ void sdf_trilinear_box(const FieldDev& f, const float* __restrict__ brick, const int* bdim, const int* bl,
                                                  float wx, float wy, float wz, float& val, float& gx, float& gy, float& gz) {
  float ux = (wx - f.ox) * f.inv_pitch, uy = (wy - f.oy) * f.inv_pitch, uz = (wz - f.oz) * f.inv_pitch;
  int ix = min(max((int)floorf(ux), 0), f.nx - 2);
  int iy = min(max((int)floorf(uy), 0), f.ny - 2);
  int iz = min(max((int)floorf(uz), 0), f.nz - 2);
  float fx = ux - (float)ix, fy = uy - (float)iy, fz = uz - (float)iz;
  const bool inx = (fx >= 0.f) && (fx <= 1.f), iny = (fy >= 0.f) && (fy <= 1.f), inz = (fz >= 0.f) && (fz <= 1.f);
  fx = fminf(fmaxf(fx, 0.f), 1.f);
  fy = fminf(fmaxf(fy, 0.f), 1.f);
  fz = fminf(fmaxf(fz, 0.f), 1.f);
  float c000, c001, c010, c011, c100, c101, c110, c111;
  const int lx = ix - bl[0], ly = iy - bl[1], lz = iz - bl[2];
  if (lx >= 0 && ly >= 0 && lz >= 0 && lx <= bdim[0] - 2 && ly <= bdim[1] - 2 && lz <= bdim[2] - 2) {
    const int sy = bdim[2], sx = bdim[1] * bdim[2];
    const float* p = brick + lx * sx + ly * sy + lz;
    c000 = p[0]; c001 = p[1]; c010 = p[sy]; c011 = p[sy + 1];
    c100 = p[sx]; c101 = p[sx + 1]; c110 = p[sx + sy]; c111 = p[sx + sy + 1];
  } else {
    const float* p = f.data + ((long long)ix * f.ny + iy) * f.nzp + iz;
    const long long sy = f.nzp, sx = (long long)f.ny * f.nzp;
    c000 = __ldg(p); c001 = __ldg(p + 1); c010 = __ldg(p + sy); c011 = __ldg(p + sy + 1);
    c100 = __ldg(p + sx); c101 = __ldg(p + sx + 1); c110 = __ldg(p + sx + sy); c111 = __ldg(p + sx + sy + 1);
  }
  const float d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
  const float z00 = fmaf(fz, d00, c000), z01 = fmaf(fz, d01, c010), z10 = fmaf(fz, d10, c100), z11 = fmaf(fz, d11, c110);
  const float y0 = fmaf(fy, z01 - z00, z00), y1 = fmaf(fy, z11 - z10, z10);
  val = fmaf(fx, y1 - y0, y0);
  const float dy0 = z01 - z00, dy1 = z11 - z10;
  const float dz0 = fmaf(fy, d01 - d00, d00), dz1 = fmaf(fy, d11 - d10, d10);
  gx = inx ? (y1 - y0) * f.inv_pitch : 0.f;
  gy = iny ? fmaf(fx, dy1 - dy0, dy0) * f.inv_pitch : 0.f;
  gz = inz ? fmaf(fx, dz1 - dz0, dz0) * f.inv_pitch : 0.f;
}

struct PipeParams {
  LinParams lin;
  const int* chunk_link;  // [nchunks] link of each chunk
  int slot_floats;        // capacity of one brick slot
  int ncons;              // consumer warps
};

// NP: padded tensor-core tile width (8 or 16); NOPT_CT: number of optimised joints when known at compile time (0: runtime)
template <int NP, int NOPT_CT>
__global__ void __launch_bounds__((PIPE_MAX_CONS + 1) * 32, 2) k_linearize_pipe(const __grid_constant__ PipeParams pp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LinParams& p = pp.lin;
  PipeShared& S = *reinterpret_cast<PipeShared*>(smem_raw);
  const RobotDev& R = *p.robot;
  const int nopt = NOPT_CT ? NOPT_CT : R.nopt, RS = nopt + 1, NC = pp.ncons;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  size_t off = (sizeof(PipeShared) + 127) & ~(size_t)127;
  float* ring = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)PIPE_NSLOT * pp.slot_floats * sizeof(float);
  const int st_floats = ((32 * RS + 16 + 31) / 32) * 32;
  float* stage_base = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)NC * st_floats * sizeof(float);
  const int red_floats = nopt * nopt + nopt + 2;
  float* red_base = reinterpret_cast<float*>(smem_raw + off);  // [2][NC][red_floats]
  uint64_t* slot_full = reinterpret_cast<uint64_t*>(S.slot_full);
  uint64_t* slot_empty = reinterpret_cast<uint64_t*>(S.slot_empty);
  uint64_t* ctx_full = reinterpret_cast<uint64_t*>(S.ctx_full);
  uint64_t* ctx_empty = reinterpret_cast<uint64_t*>(S.ctx_empty);

  if (threadIdx.x == 0) {
    for (int s = 0; s < PIPE_NSLOT; ++s) {
      mbar_init(slot_full + s, 1);
      mbar_init(slot_empty + s, NC);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(ctx_full + c, 1);
      mbar_init(ctx_empty + c, NC);
    }
    mbar_fence_init();
  }
  if ((int)threadIdx.x < R.nlinks) {
    LinkMeta m;
    m.c0 = R.link_chunk0[threadIdx.x];
    m.c1 = R.link_chunk0[threadIdx.x + 1];
    m.pt_start = R.link_pt_start[threadIdx.x];
    m.pt_end = m.pt_start + R.link_pt_count[threadIdx.x];
    m.mask = R.link_optmask[threadIdx.x];
    S.links[threadIdx.x] = m;
  }
  __syncthreads();

  const int nknots = p.T - p.t_lo;
  const int nprob = p.nactive ? *p.nactive : p.nproblems;
  // tail launches: split every (problem, knot) item into link ranges so that the whole chip works on the few problems left
  const int nsplit = p.allow_split ? split_factor((long long)nprob * nknots, (int)gridDim.x) : 1;
  const int sidx = nsplit == 4 ? 2 : (nsplit == 2 ? 1 : 0);
  const long long nitems = (long long)nprob * nknots * nsplit;

  if (warp == NC) {
    // =============================== PRODUCER WARP ===============================
    unsigned ic = 0, bc = 0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++ic) {
      const int ci = ic & 1;
      if (ic >= 2) mbar_wait_sleep(ctx_empty + ci, ((ic >> 1) - 1) & 1);
      ItemCtx& C = S.ctx[ci];
      const int part = (int)(item % nsplit);
      const long long it2 = item / nsplit;
      const int a = (int)(it2 / nknots);
      const int t = p.t_lo + (int)(it2 - (long long)a * nknots);
      const int b = p.active ? p.active[a] : a;
      const int L0 = R.part_link0[sidx][part], L1 = R.part_link0[sidx][part + 1];
      const double* q = p.q + ((long long)b * p.T + t) * R.ndof;
      const int fid = p.collision ? p.field_ids[2 * b + (t < p.knot_standoff ? 0 : 1)] : -1;
      // ---- chain FK in float64 ----
      if (lane < R.nmov) {
        const double qj = q[R.mov_qidx[lane]];
        const double ax = R.mov_axis_d[lane][0], ay = R.mov_axis_d[lane][1], az = R.mov_axis_d[lane][2];
        double M[12];
        if (R.mov_type[lane] == GTO_JOINT_REVOLUTE) {
          double s, c;
          sincos(qj, &s, &c);
          const double v = 1.0 - c;
          M[0] = 1.0 - v * (ay * ay + az * az); M[1] = -s * az + v * ax * ay;      M[2] = s * ay + v * ax * az;       M[3] = 0.0;
          M[4] = s * az + v * ax * ay;          M[5] = 1.0 - v * (ax * ax + az * az); M[6] = -s * ax + v * ay * az;   M[7] = 0.0;
          M[8] = -s * ay + v * ax * az;         M[9] = s * ax + v * ay * az;       M[10] = 1.0 - v * (ax * ax + ay * ay); M[11] = 0.0;
        } else {
          M[0] = 1.0; M[1] = 0.0; M[2] = 0.0; M[3] = qj * ax;
          M[4] = 0.0; M[5] = 1.0; M[6] = 0.0; M[7] = qj * ay;
          M[8] = 0.0; M[9] = 0.0; M[10] = 1.0; M[11] = qj * az;
        }
        double Cm[12];
        mul34(R.mov_origin_d[lane], M, Cm);
#pragma unroll
        for (int e = 0; e < 12; ++e) S.A[lane][e] = Cm[e];
      }
      __syncwarp();
      for (int j = 0; j < R.nmov; ++j) {
        if (lane < 12) {
          const int r = lane >> 2, c = lane & 3;
          const int pj = R.mov_parent[j];
          double s;
          if (pj < 0) {
            s = S.A[j][lane];
          } else {
            const double* P = S.Tm[pj];
            s = dot3<double>(P[r * 4 + 0], S.A[j][c], P[r * 4 + 1], S.A[j][4 + c], P[r * 4 + 2], S.A[j][8 + c]);
            if (c == 3) s += P[r * 4 + 3];
          }
          S.Tm[j][lane] = s;
        }
        __syncwarp();
      }
      if (lane >= L0 && lane < L1) {  // visual frames + brick placement (links of this part only)
        const int mj = R.link_mov[lane];
        double Fd[12];
        if (mj < 0) {
#pragma unroll
          for (int e = 0; e < 12; ++e) Fd[e] = R.link_tf_d[lane][e];
        } else {
          mul34(S.Tm[mj], R.link_tf_d[lane], Fd);
        }
        float F[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) {
          F[e] = (float)Fd[e];
          C.frames[lane][e] = F[e];
        }
        if (fid >= 0) {
          const FieldDev& f = p.fields[fid];
          const float* cc = R.link_center[lane];
          const float* hh = R.link_half[lane];
          const float bp[3] = {p.base[4 * b + 0], p.base[4 * b + 1], p.base[4 * b + 2]};
          const float org[3] = {f.ox, f.oy, f.oz};
          const int N3[3] = {f.nx, f.ny, f.nz};
          int lo3[3], sz3[3];
          bool fast = true;
#pragma unroll
          for (int a3 = 0; a3 < 3; ++a3) {
            const float cw = F[a3 * 4 + 0] * cc[0] + F[a3 * 4 + 1] * cc[1] + F[a3 * 4 + 2] * cc[2] + F[a3 * 4 + 3] + bp[a3];
            const float hw = fabsf(F[a3 * 4 + 0]) * hh[0] + fabsf(F[a3 * 4 + 1]) * hh[1] + fabsf(F[a3 * 4 + 2]) * hh[2] + 1e-4f;
            int lo = (int)floorf((cw - hw - org[a3]) * f.inv_pitch);
            const int hi = (int)floorf((cw + hw - org[a3]) * f.inv_pitch) + 1;  // highest node touched
            if (lo < 0 || hi > N3[a3] - 1) fast = false;                       // a point may need index clamping
            if (a3 == 2) lo &= ~3;
            const int need = hi - lo + 1;
            int sz = min(32, max(8, (need + 3) & ~3));
            if (need > sz) fast = false;
            lo3[a3] = lo;
            sz3[a3] = sz;
          }
          while (sz3[0] * sz3[1] * sz3[2] > pp.slot_floats) {  // does not fit a slot: shrink the longest axis
            int am = 0;
            if (sz3[1] > sz3[am]) am = 1;
            if (sz3[2] > sz3[am]) am = 2;
            sz3[am] -= 4;
            fast = false;
          }
          C.blo[lane][0] = lo3[0]; C.blo[lane][1] = lo3[1]; C.blo[lane][2] = lo3[2]; C.blo[lane][3] = 0;
          C.bdim[lane][0] = sz3[0]; C.bdim[lane][1] = sz3[1]; C.bdim[lane][2] = sz3[2]; C.bdim[lane][3] = fast ? 1 : 0;
#pragma unroll
          for (int a3 = 0; a3 < 3; ++a3) C.cl[lane][a3] = (bp[a3] - org[a3]) * f.inv_pitch - (float)lo3[a3];
          C.cl[lane][3] = f.inv_pitch;
        }
      }
      if (lane < nopt) {
        const int j = R.opt_mov[lane];
        double om[3] = {0.0, 0.0, 0.0}, mm[3] = {0.0, 0.0, 0.0};
        if (j >= 0) {
          const double* Tj = S.Tm[j];
          const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
          const double zx = Tj[0] * ax + Tj[1] * ay + Tj[2] * az;
          const double zy = Tj[4] * ax + Tj[5] * ay + Tj[6] * az;
          const double zz = Tj[8] * ax + Tj[9] * ay + Tj[10] * az;
          if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
            const double ox = Tj[3], oy = Tj[7], oz = Tj[11];
            om[0] = zx; om[1] = zy; om[2] = zz;
            mm[0] = oy * zz - oz * zy; mm[1] = oz * zx - ox * zz; mm[2] = ox * zy - oy * zx;
          } else {
            mm[0] = zx; mm[1] = zy; mm[2] = zz;
          }
        }
        C.tw[lane][0] = (float)om[0]; C.tw[lane][1] = (float)om[1]; C.tw[lane][2] = (float)om[2]; C.tw[lane][3] = 0.f;
        C.tw[lane][4] = (float)mm[0]; C.tw[lane][5] = (float)mm[1]; C.tw[lane][6] = (float)mm[2]; C.tw[lane][7] = 0.f;
      }
      if (lane == 31) {
        double F[12];
        if (R.grip_mov < 0) {
#pragma unroll
          for (int e = 0; e < 12; ++e) F[e] = R.grip_tf_d[e];
        } else {
          mul34(S.Tm[R.grip_mov], R.grip_tf_d, F);
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) {
          C.gripf[e] = (float)F[e];
          C.goal[0][e] = (float)(F[e] - p.goal_tf[(long long)b * 24 + e]);
          C.goal[1][e] = (float)(F[e] - p.goal_tf[(long long)b * 24 + 12 + e]);
        }
        C.b = b; C.t = t; C.fid = fid;
        C.obuf = p.bufsel ? (1 - p.bufsel[b]) : 0;
        C.part = part; C.l0 = L0; C.l1 = L1; C.last_part = (part == nsplit - 1);
      }
      if (lane >= 24 && lane < 27) C.basep[lane - 24] = p.base[4 * b + (lane - 24)];
      __syncwarp();
      if (lane == 0) mbar_arrive(ctx_full + ci);  // release: the context is visible to the consumers
      // ---- one TMA brick per link into the ring ----
      if (fid >= 0) {
        if (lane == 0) {
          const CUtensorMap* maps = p.fields[fid].maps2;
          for (int l = L0; l < L1; ++l) {
            const unsigned idx = bc + (l - L0), s = idx % PIPE_NSLOT;
            if (idx >= PIPE_NSLOT) mbar_wait_sleep(slot_empty + s, ((idx / PIPE_NSLOT) - 1) & 1);
            const int sx = C.bdim[l][0], sy = C.bdim[l][1], sz = C.bdim[l][2];
            const int mi = ((sx / 4 - 2) * PIPE_NAXC + (sy / 4 - 2)) * PIPE_NAXC + (sz / 4 - 2);
            mbar_expect_tx(slot_full + s, (uint32_t)(sx * sy * sz * sizeof(float)));
            tma_load_3d(ring + (size_t)s * pp.slot_floats, maps + mi, C.blo[l][2], C.blo[l][1], C.blo[l][0], slot_full + s);
          }
        }
        bc += L1 - L0;
        __syncwarp();
      }
    }
    return;
  }

  // =============================== CONSUMER WARPS ===============================
  float* stage = stage_base + warp * st_floats;
  const int gq = lane >> 2, tq = lane & 3;
  unsigned ic = 0, bc = 0;
  for (long long item = blockIdx.x; item < nitems; item += gridDim.x, ++ic) {
    const int ci = ic & 1;
    mbar_wait_sleep(ctx_full + ci, (ic >> 1) & 1);
    const ItemCtx& C = S.ctx[ci];
    const int b = C.b, t = C.t, fid = C.fid, L0 = C.l0, L1 = C.l1;
    const bool is_goal = C.last_part && (t == p.T - 1), is_stand = C.last_part && (p.use_standoff && t == p.knot_standoff);
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    float gacc[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) gacc[k] = 0.f;
    float cacc = 0.f;
    float* rows_b = p.rows ? p.rows + (long long)(b - p.b0) * p.rows_per_problem * RS : nullptr;

    if (p.collision) {
      int ch = S.links[L0].c0 + warp;  // chunks are dealt round-robin over the item's links: first, first+NC, ...
      for (int l = L0; l < L1; ++l) {
        // every consumer warp observes every brick (full) before it releases it (empty), chunks or not: an early release
        // of a later use of the same slot could otherwise complete the empty barrier of the current use
        const float* brick = ring;
        float clx = 0.f, cly = 0.f, clz = 0.f, ipitch = 0.f;
        int dxm2 = 0, dym2 = 0, dzm2 = 0, dy = 0, dz = 0, fast = 0;
        if (fid >= 0) {
          const unsigned idx = bc + (l - L0);
          mbar_wait_sleep(slot_full + (idx % PIPE_NSLOT), (idx / PIPE_NSLOT) & 1);
          brick = ring + (size_t)(idx % PIPE_NSLOT) * pp.slot_floats;
          clx = C.cl[l][0]; cly = C.cl[l][1]; clz = C.cl[l][2]; ipitch = C.cl[l][3];
          dxm2 = C.bdim[l][0] - 2; dy = C.bdim[l][1]; dz = C.bdim[l][2]; fast = C.bdim[l][3];
          dym2 = dy - 2; dzm2 = dz - 2;
        }
        const LinkMeta lm = S.links[l];
        const int c0 = lm.c0, c1 = lm.c1;
        if (ch < c1) {
          const unsigned mask = lm.mask;
          float F[12];
#pragma unroll
          for (int e = 0; e < 12; ++e) F[e] = C.frames[l][e];
          for (; ch < c1; ch += NC) {
            const int p0 = lm.pt_start + 32 * (ch - c0), cnt = min(32, lm.pt_end - p0);
            const bool act = lane < cnt;
            float J[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) J[k] = 0.f;
            float r = 0.f;
            if (act && fid >= 0) {
              const float x = __ldg(p.px + p0 + lane), y = __ldg(p.py + p0 + lane), z = __ldg(p.pz + p0 + lane);
              const float wbx = F[0] * x + F[1] * y + F[2] * z + F[3];
              const float wby = F[4] * x + F[5] * y + F[6] * z + F[7];
              const float wbz = F[8] * x + F[9] * y + F[10] * z + F[11];
              float val, gx, gy, gz;
              if (fast) {
                // the brick encloses every point of this link and lies inside the grid: no grid clamping can occur
                const float ux = fmaf(wbx, ipitch, clx), uy = fmaf(wby, ipitch, cly), uz = fmaf(wbz, ipitch, clz);
                const int ix = min(max((int)floorf(ux), 0), dxm2), iy = min(max((int)floorf(uy), 0), dym2), iz = min(max((int)floorf(uz), 0), dzm2);
                const float fx = ux - (float)ix, fy = uy - (float)iy, fz = uz - (float)iz;
                const float* q8 = brick + (ix * dy + iy) * dz + iz;
                const int sx = dy * dz;
                const float c000 = q8[0], c001 = q8[1], c010 = q8[dz], c011 = q8[dz + 1];
                const float c100 = q8[sx], c101 = q8[sx + 1], c110 = q8[sx + dz], c111 = q8[sx + dz + 1];
                const float d00 = c001 - c000, d01 = c011 - c010, d10 = c101 - c100, d11 = c111 - c110;
                const float z00 = fmaf(fz, d00, c000), z01 = fmaf(fz, d01, c010), z10 = fmaf(fz, d10, c100), z11 = fmaf(fz, d11, c110);
                const float y0 = fmaf(fy, z01 - z00, z00), y1 = fmaf(fy, z11 - z10, z10);
                val = fmaf(fx, y1 - y0, y0);
                const float dy0 = z01 - z00, dy1 = z11 - z10;
                const float dz0 = fmaf(fy, d01 - d00, d00), dz1 = fmaf(fy, d11 - d10, d10);
                gx = (y1 - y0) * ipitch;
                gy = fmaf(fx, dy1 - dy0, dy0) * ipitch;
                gz = fmaf(fx, dz1 - dz0, dz0) * ipitch;
              } else {
                sdf_trilinear_box(p.fields[fid], brick, C.bdim[l], C.blo[l], wbx + C.basep[0], wby + C.basep[1], wbz + C.basep[2], val, gx, gy, gz);
              }
              r = p.sw_obs * val;
              gx *= p.sw_obs; gy *= p.sw_obs; gz *= p.sw_obs;
              const float nx = wby * gz - wbz * gy, ny = wbz * gx - wbx * gz, nz = wbx * gy - wby * gx;
#pragma unroll
              for (int k = 0; k < NP; ++k) {
                if (k < nopt && ((mask >> k) & 1u)) {
                  const float4 o4 = *reinterpret_cast<const float4*>(&C.tw[k][0]);
                  const float4 m4 = *reinterpret_cast<const float4*>(&C.tw[k][4]);
                  J[k] = o4.x * nx + o4.y * ny + o4.z * nz + m4.x * gx + m4.y * gy + m4.z * gz;
                }
              }
            }
            cacc = fmaf(r, r, cacc);
#pragma unroll
            for (int k = 0; k < NP; ++k)
              if (k < nopt) gacc[k] = fmaf(J[k], r, gacc[k]);
            if (NOPT_CT == 7) {  // row = [J0..J6 | r] = 32 bytes: two 128-bit shared stores
              float4* s4 = reinterpret_cast<float4*>(stage + lane * 8);
              s4[0] = make_float4(J[0], J[1], J[2], J[3]);
              s4[1] = make_float4(J[4], J[5], J[6], r);
            } else {
#pragma unroll
              for (int k = 0; k < NP; ++k)
                if (k < nopt) stage[lane * RS + k] = J[k];
              stage[lane * RS + nopt] = r;
            }
            __syncwarp();
            mma_rows<NP>(stage, RS, cnt, acc0, acc1, lane);
            if (rows_b) {
              float* dst = rows_b + ((long long)t * R.npoints + p0) * RS;
              if (NOPT_CT == 7 && cnt == 32) {  // 1 KB tile, 16-byte aligned by construction
                const float4* s4 = reinterpret_cast<const float4*>(stage);
                float4* d4 = reinterpret_cast<float4*>(dst);
                const float4 v0 = s4[lane], v1 = s4[lane + 32];
                __stcs(d4 + lane, v0);
                __stcs(d4 + lane + 32, v1);
              } else {
                store_rows(dst, stage, cnt * RS, lane);
              }
            }
            __syncwarp();
          }
        }
        if (fid >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(slot_empty + ((bc + (l - L0)) % PIPE_NSLOT));
        }
      }
      if (fid >= 0) bc += L1 - L0;
    }

    // ---- goal / stand-off rows ----
    if (is_goal || is_stand) {
      const int Pg = R.grip_pt_count;
      const long long obs_rows = p.collision ? (long long)p.T * R.npoints : 0;
      const float* Fg = C.gripf;
      const unsigned mask = R.grip_optmask;
      for (int which = 0; which < 2; ++which) {
        if (which == 0 && !is_goal) continue;
        if (which == 1 && !is_stand) continue;
        const float* Dg = C.goal[which];
        const long long rbase = obs_rows + (which == 1 ? 3LL * Pg : 0);
        const int nch = (Pg + 31) / 32;
        for (int ch = warp; ch < nch; ch += NC) {
          const int k0 = ch * 32, cnt = min(32, Pg - k0);
          const bool act = lane < cnt;
          float w3[3] = {0.f, 0.f, 0.f}, r3[3] = {0.f, 0.f, 0.f};
          if (act) {
            const int pi = R.grip_pt_start + k0 + lane;
            const float x = __ldg(p.px + pi), y = __ldg(p.py + pi), z = __ldg(p.pz + pi);
#pragma unroll
            for (int a3 = 0; a3 < 3; ++a3) {
              w3[a3] = Fg[a3 * 4 + 0] * x + Fg[a3 * 4 + 1] * y + Fg[a3 * 4 + 2] * z + Fg[a3 * 4 + 3];
              r3[a3] = p.sw_goal * (Dg[a3 * 4 + 0] * x + Dg[a3 * 4 + 1] * y + Dg[a3 * 4 + 2] * z + Dg[a3 * 4 + 3]);
            }
          }
#pragma unroll
          for (int a3 = 0; a3 < 3; ++a3) {
            float J[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              J[k] = 0.f;
              if (act && k < nopt && ((mask >> k) & 1u)) {
                const float4 o4 = *reinterpret_cast<const float4*>(&C.tw[k][0]);
                const float4 m4 = *reinterpret_cast<const float4*>(&C.tw[k][4]);
                float v;
                if (a3 == 0) v = o4.y * w3[2] - o4.z * w3[1] + m4.x;
                else if (a3 == 1) v = o4.z * w3[0] - o4.x * w3[2] + m4.y;
                else v = o4.x * w3[1] - o4.y * w3[0] + m4.z;
                J[k] = p.sw_goal * v;
              }
            }
            const float r = r3[a3];
            cacc = fmaf(r, r, cacc);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              if (k < nopt) {
                gacc[k] = fmaf(J[k], r, gacc[k]);
                stage[lane * RS + k] = J[k];
              }
            }
            stage[lane * RS + nopt] = r;
            __syncwarp();
            mma_rows<NP>(stage, RS, cnt, acc0, acc1, lane);
            if (rows_b) store_rows(rows_b + (rbase + (long long)a3 * Pg + k0) * RS, stage, cnt * RS, lane);
            __syncwarp();
          }
        }
      }
    }

    // ---- reduce over lanes / consumer warps, write the per-knot Gauss-Newton block ----
#pragma unroll
    for (int k = 0; k < NP; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gacc[k] += __shfl_xor_sync(0xffffffffu, gacc[k], o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cacc += __shfl_xor_sync(0xffffffffu, cacc, o);
    {
      float* red = red_base + ((size_t)ci * NC + warp) * red_floats;
      const int c0 = 2 * tq;
      if (gq < nopt) {
        if (c0 < nopt) red[gq * nopt + c0] = acc0[0];
        if (c0 + 1 < nopt) red[gq * nopt + c0 + 1] = acc0[1];
      }
      if (NP == 16) {
        if (gq + 8 < nopt) {
          if (c0 < nopt) red[(gq + 8) * nopt + c0] = acc0[2];
          if (c0 + 1 < nopt) red[(gq + 8) * nopt + c0 + 1] = acc0[3];
        }
        if (gq < nopt) {
          if (c0 + 8 < nopt) red[gq * nopt + c0 + 8] = acc1[0];
          if (c0 + 9 < nopt) red[gq * nopt + c0 + 9] = acc1[1];
        }
        if (gq + 8 < nopt) {
          if (c0 + 8 < nopt) red[(gq + 8) * nopt + c0 + 8] = acc1[2];
          if (c0 + 9 < nopt) red[(gq + 8) * nopt + c0 + 9] = acc1[3];
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NP; ++k)
          if (k < nopt) red[nopt * nopt + k] = gacc[k];
        red[nopt * nopt + nopt] = cacc;
      }
    }
    const int obuf = C.obuf;
    asm volatile("bar.sync 1, %0;" ::"r"(NC * 32) : "memory");  // consumers only; the producer keeps running ahead
    {
      const int nH = nopt * nopt, ntot = nH + nopt + 1;
      for (int i = threadIdx.x; i < ntot; i += NC * 32) {
        float s = 0.f;
        for (int w = 0; w < NC; ++w) s += red_base[((size_t)ci * NC + w) * red_floats + i];
        const long long bt = (long long)b * p.T + t;
        if (i < nH) p.H[obuf * p.buf_stride_H + C.part * p.part_stride_H + bt * nH + i] = s;
        else if (i < nH + nopt) p.g[obuf * p.buf_stride_g + C.part * p.part_stride_g + bt * nopt + (i - nH)] = s;
        else p.costp[obuf * p.buf_stride_c + C.part * p.part_stride_c + bt] = s;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ctx_empty + ci);  // this warp no longer reads ctx[ci] / red[ci]
  }
}
