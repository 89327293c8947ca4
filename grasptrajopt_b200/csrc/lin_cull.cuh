// lin_cull.cuh -- k_item_fk + k_linearize_cull: the fused linearisation of one Gauss-Newton iteration.
// Included by gto_b200.cu (uses its device tables and PTX helpers).
//
// Warp-specialised persistent CTAs (1 producer warp + NC consumer warps + 1 zero-row warp), SDF bricks staged by TMA into a
// shared-memory ring (box sizes per axis from {8,12,...,32}, one tensor map per combination and field; the start coordinate
// of the innermost (z) axis is kept a multiple of 4 elements: TMA faults on a tile whose innermost start is not 16-byte
// aligned), plus three things:
//   * exact culling   the cost field is identically zero away from obstacles (DepthPointCloud.get_sdf_cost is 0 for
//                     d >= epsilon, mesh_to_sdf/depth_point_cloud.py:65-91).  The producer tests, per link, whether the
//                     box of grid nodes its points can touch holds any non-zero node -- 8 reads of a summed-volume table
//                     built at gto_set_field -- and only links that do are handed to the consumer warps.  The Jacobian
//                     rows, J^T J, J^T r and the cost of a culled link are exactly zero, so results are bit-identical.
//   * zero rows by TMA the dense row block of a culled link is still materialised in HBM (it is part of the Jacobian):
//                     one cp.async.bulk shared->global copy from a zeroed shared buffer per link (SASS UBLKCP), issued
//                     by the producer; no consumer instruction is spent on it.
//   * dynamic items   (problem, knot) items are handed out through a global atomic counter, because their cost now
//                     varies with the number of surviving links.
#pragma once

#define CULL_NSLOT_MAX 8  // brick ring slots: run-time (CullParams.nslot), default 4
#define CULL_ZERO_BYTES 8192
#define CULL_NCTX 4        // item records in flight per CTA (the producer runs up to this many items ahead of the consumers)
#define CULL_REC_LINKS 16  // collision links per item record (Panda 12, Fetch 10); gto_set_robot rejects more

struct LinkMeta {
  int c0, c1, pt_start, pt_end;
  unsigned mask;
};

#ifndef GTO_MBAR_HINT_NS
#define GTO_MBAR_HINT_NS 2000  // (100 / 500 / 2000 ns measured equal on C2)
#endif
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
        : "r"(addr), "r"(parity), "r"((unsigned)GTO_MBAR_HINT_NS)
        : "memory");
    if (ok) return;
  }
  __trap();  // a lost transaction must abort the launch, never hang the device
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic (slow-path) trilinear lookup for points whose cell may need index clamping or lies outside
// the staged brick; same arithmetic and association order as the fast path below and as the oracle
// (ulx, uly, ulz): float64 voxel coordinate relative to the brick's lower corner bl
__device__ __forceinline__ void sdf_trilinear_box(const FieldDev& f, const float* __restrict__ brick, const int* bdim, const int* bl,
                                                  double ulx, double uly, double ulz, float& val, float& gx, float& gy, float& gz) {
  const double ux = ulx + (double)bl[0], uy = uly + (double)bl[1], uz = ulz + (double)bl[2];  // grid voxel coordinate
  int ix = min(max((int)floor(ux), 0), f.nx - 2);
  int iy = min(max((int)floor(uy), 0), f.ny - 2);
  int iz = min(max((int)floor(uz), 0), f.nz - 2);
  float fx = (float)(ux - (double)ix), fy = (float)(uy - (double)iy), fz = (float)(uz - (double)iz);
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

struct __align__(16) CullCtx {
  float frames[CULL_REC_LINKS][12];  // visual frame of each collision link (robot base frame)
  float tw[GTO_MAX_OPT][8];         // (omega.xyz, -, m.xyz, -) per optimised joint
  float gripf[12];
  float goal[2][12];                // gripper frame minus goal / stand-off frame
  double vf[CULL_REC_LINKS][12];     // float64 3x4 map: point in the link's visual frame -> brick-local voxel coordinate
  int blo[CULL_REC_LINKS][4];        // brick lower corner (grid index) per link
  int bdim[CULL_REC_LINKS][4];       // brick dims (x, y, z) and fast flag
  float basep[4];
  int b, t, fid, obuf;
  int nact, kind;                   // surviving links; kind bit 0: goal rows, bit 1: stand-off rows
  unsigned amask;                   // bit l: link l survived the culling test
  int pad_;
  int act[CULL_REC_LINKS];           // surviving link ids, ascending
};

#define CULL_ZQ 16  // entries of the producer -> zero-row warp queue
struct CullShared {
  CullCtx ctx[CULL_NCTX];
  LinkMeta links[CULL_REC_LINKS];
  unsigned long long slot_full[CULL_NSLOT_MAX], slot_empty[CULL_NSLOT_MAX], ctx_full[CULL_NCTX], ctx_empty[CULL_NCTX];
  int4 zq[CULL_ZQ];            // (b, t, amask, -) of the items whose culled links still need their zero rows; b < 0: stop
  unsigned long long zq_full[CULL_ZQ], zq_empty[CULL_ZQ];  // producer -> zero-row warp hand-over (mbarriers: waiting threads are parked)
  unsigned long long pts_full;                             // the robot's surface points have arrived in shared memory
};

struct CullParams {
  LinParams lin;
  int slot_floats;     // capacity of one brick slot
  int ncons;           // consumer warps
  int nslot;           // brick ring slots (<= CULL_NSLOT_MAX)
  int count_early;     // the launch before this one is k_item_fk (not the step kernel): *nactive may be read before pdl_wait
  int* work_counter;   // zero before the launch: next (problem, knot) item
  CullCtx* recs;       // [items] per-item records written by k_item_fk, read by the producer warps (bulk copy)
  CullCtx* rec_dummy;  // one record nobody reads
  long long* dbg;              // k_item_fk: clock64() at phase boundaries of CTA 0 (diagnostics, NULL: off)
  unsigned long long* ts_fk;   // launch time stamps of k_item_fk / k_linearize_cull (NULL: off), see stamp_begin
  unsigned long long* ts_lin;
  unsigned long long* stats;  // [0] items, [1] links tested, [2] links that survived the culling test (NULL: off)
};

__device__ __forceinline__ void bulk_store_zero(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// number of non-zero nodes in the inclusive box [x0,x1] x [y0,y1] x [z0,z1] of the field (summed-volume table lookup)
__device__ __forceinline__ unsigned svt_count(const FieldDev& f, int x0, int x1, int y0, int y1, int z0, int z1) {
  const unsigned sy = (unsigned)(f.nz + 1), sx = (unsigned)(f.ny + 1) * sy;
  const unsigned* S = f.svt;
  const unsigned X0 = x0 * sx, X1 = (x1 + 1) * sx, Y0 = y0 * sy, Y1 = (y1 + 1) * sy, Z0 = z0, Z1 = z1 + 1;
  return __ldg(S + X1 + Y1 + Z1) - __ldg(S + X0 + Y1 + Z1) - __ldg(S + X1 + Y0 + Z1) - __ldg(S + X1 + Y1 + Z0) +
         __ldg(S + X0 + Y0 + Z1) + __ldg(S + X0 + Y1 + Z0) + __ldg(S + X1 + Y0 + Z0) - __ldg(S + X0 + Y0 + Z0);
}


// ------------------------------------------------------------------------------------------------------------------
// k_item_fk: everything of a (problem, knot) item that does not depend on the surface points, 16 lanes per item:
// float64 chain FK (optas/models.py:826-868) -> visual frames of the collision links (gto/gto_models.py:83-101), joint
// twists, gripper-minus-goal frames, SDF brick placement per link and the culling test.  One record per item in HBM;
// the linearise kernel's producer warp pulls it into shared memory with one bulk copy.  Items whose Gauss-Newton
// block is identically zero (no surviving link, no goal rows) get their zeros written here.
// ------------------------------------------------------------------------------------------------------------------
// 16 lanes (`hl` = lane within the group, `hshift` = 0 / 16: which half of the warp) compute the record of item `item` =
// (problem b, knot t) at configuration q; A / Tm: 2 x nmov x 12 doubles of shared scratch owned by the group.  The two
// halves of a warp run in lock step (full-warp barriers, one instruction stream for two items): a half without an item
// of its own (`valid` false) recomputes a neighbour's item into the dummy record.
#define FK_MARK(k) do { if (pp.dbg && blockIdx.x == 0 && threadIdx.x == 0) pp.dbg[32 + (k)] = clock64(); } while (0)
__device__ __forceinline__ void item_fk_body(const CullParams& pp, CullCtx* recs, const RobotDev& R, int item, bool valid, int b, int t, int obuf,
                                             const double* q, double* A, double* Tm, int hl, int hshift) {
  const LinParams& p = pp.lin;
  const int nopt = R.nopt, nmov = R.nmov, nlinks = R.nlinks;
  const int fid = p.collision ? p.field_ids[2 * b + (t < p.knot_standoff ? 0 : 1)] : -1;
  const bool cull = !(p.flags & GTO_FLAG_NO_CULL);
  CullCtx& C = valid ? recs[item] : *pp.rec_dummy;
  // issued before the FK chain, consumed after it: field geometry, base offset
  FieldDev fld;
  fld.data = nullptr; fld.nx = fld.ny = fld.nz = fld.nzp = 2; fld.ox = fld.oy = fld.oz = 0.f; fld.inv_pitch = 1.f; fld.inv_pitch_d = 1.0; fld.org_d[0] = fld.org_d[1] = fld.org_d[2] = 0.0;
  fld.maps2 = nullptr; fld.svt = nullptr;
  if (fid >= 0) fld = p.fields[fid];
  const float bpx = p.base[4 * b + 0], bpy = p.base[4 * b + 1], bpz = p.base[4 * b + 2];

  FK_MARK(1);
  for (int j = hl; j < nmov; j += 16) {  // A_j = origin_j * motion_j(q_j)
    const double qj = q[R.mov_qidx[j]];
    const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
    double M[12];
    if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
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
    mul34(R.mov_origin_d[j], M, Cm);
#pragma unroll
    for (int e = 0; e < 12; ++e) A[j * 12 + e] = Cm[e];
  }
  __syncwarp();
  FK_MARK(2);
  for (int j = 0; j < nmov; ++j) {  // sequential along the tree, 12 lanes per product
    if (hl < 12) {
      const int r = hl >> 2, c = hl & 3;
      const int pj = R.mov_parent[j];
      double s;
      if (pj < 0) {
        s = A[j * 12 + hl];
      } else {
        const double* P = Tm + pj * 12;
        s = dot3<double>(P[r * 4 + 0], A[j * 12 + c], P[r * 4 + 1], A[j * 12 + 4 + c], P[r * 4 + 2], A[j * 12 + 8 + c]);
        if (c == 3) s += P[r * 4 + 3];
      }
      Tm[j * 12 + hl] = s;
    }
    __syncwarp();
  }
  FK_MARK(3);
  // ---- visual frames, brick placement and the culling test (one lane per link) ----
  unsigned amask = 0u;
  for (int l0 = 0; l0 < nlinks; l0 += 16) {
    const int l = l0 + hl;
    bool survives = false;
    if (l < nlinks) {
      const int mj = R.link_mov[l];
      double Fd[12];
      if (mj < 0) {
#pragma unroll
        for (int e = 0; e < 12; ++e) Fd[e] = R.link_tf_d[l][e];
      } else {
        mul34(Tm + mj * 12, R.link_tf_d[l], Fd);
      }
      float F[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) F[e] = (float)Fd[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) reinterpret_cast<float4*>(C.frames[l])[e] = make_float4(F[4 * e], F[4 * e + 1], F[4 * e + 2], F[4 * e + 3]);
      if (fid >= 0) {
        const FieldDev& f = fld;
        const float* cc = R.link_center[l];
        const float* hh = R.link_half[l];
        const float bp[3] = {bpx, bpy, bpz};
        const float org[3] = {f.ox, f.oy, f.oz};
        const int N3[3] = {f.nx, f.ny, f.nz};
        int lo3[3], sz3[3], c0[3], c1[3];
        bool fast = true;
#pragma unroll
        for (int a3 = 0; a3 < 3; ++a3) {
          const float cw = F[a3 * 4 + 0] * cc[0] + F[a3 * 4 + 1] * cc[1] + F[a3 * 4 + 2] * cc[2] + F[a3 * 4 + 3] + bp[a3];
          const float hw = fabsf(F[a3 * 4 + 0]) * hh[0] + fabsf(F[a3 * 4 + 1]) * hh[1] + fabsf(F[a3 * 4 + 2]) * hh[2] + 1e-4f;
          int lo = (int)floorf((cw - hw - org[a3]) * f.inv_pitch);
          const int hi = (int)floorf((cw + hw - org[a3]) * f.inv_pitch) + 1;  // highest node touched
          // nodes a (possibly index-clamped) lookup of this link can read, for the culling test
          c0[a3] = min(max(lo, 0), N3[a3] - 1);
          c1[a3] = min(max(hi, 0), N3[a3] - 1);
          if (lo < 0 || hi > N3[a3] - 1) fast = false;  // a point may need index clamping
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
        *reinterpret_cast<int4*>(C.blo[l]) = make_int4(lo3[0], lo3[1], lo3[2], 0);
        *reinterpret_cast<int4*>(C.bdim[l]) = make_int4(sz3[0], sz3[1], sz3[2], fast ? 1 : 0);
        // voxel coordinate of a point x of this link, relative to the brick's lower corner, formed in float64:
        // u = (Fd x + base - origin) / pitch - lo.  The float32 position (frames rounded to float32, |W| ~ 1 m) is only good to
        // 6e-8 m, which puts 1e-6 of pseudo-random noise on the cost of a trajectory -- more than the cost reductions that
        // decide the last accept / reject steps of the solver.
#pragma unroll
        for (int a3 = 0; a3 < 3; ++a3) {
          const double ip = f.inv_pitch_d;
          C.vf[l][a3 * 4 + 0] = Fd[a3 * 4 + 0] * ip;
          C.vf[l][a3 * 4 + 1] = Fd[a3 * 4 + 1] * ip;
          C.vf[l][a3 * 4 + 2] = Fd[a3 * 4 + 2] * ip;
          C.vf[l][a3 * 4 + 3] = (Fd[a3 * 4 + 3] + ((double)bp[a3] - f.org_d[a3])) * ip - (double)lo3[a3];
        }
        survives = (R.link_pt_count[l] > 0) &&
                   (!cull || f.svt == nullptr || svt_count(f, c0[0], c1[0], c0[1], c1[1], c0[2], c1[2]) != 0u);
      }
    }
    const unsigned bal = (__ballot_sync(0xffffffffu, survives) >> hshift) & 0xffffu;
    if (survives) C.act[__popc(amask) + __popc(bal & ((1u << hl) - 1u))] = l;
    amask |= bal << l0;
  }
  FK_MARK(4);
  for (int k = hl; k < nopt; k += 16) {  // joint twists: v(W) = omega x W + m
    const int j = R.opt_mov[k];
    double om[3] = {0.0, 0.0, 0.0}, mm[3] = {0.0, 0.0, 0.0};
    if (j >= 0) {
      const double* Tj = Tm + j * 12;
      const double ax = R.mov_axis_d[j][0], ay = R.mov_axis_d[j][1], az = R.mov_axis_d[j][2];
      const double zx = Tj[0] * ax + Tj[1] * ay + Tj[2] * az;
      const double zy = Tj[4] * ax + Tj[5] * ay + Tj[6] * az;
      const double zz = Tj[8] * ax + Tj[9] * ay + Tj[10] * az;
      if (R.mov_type[j] == GTO_JOINT_REVOLUTE) {
        const double ox = Tj[3], oy = Tj[7], oz = Tj[11];
        om[0] = zx; om[1] = zy; om[2] = zz;
        mm[0] = oy * zz - oz * zy; mm[1] = oz * zx - ox * zz; mm[2] = ox * zy - oy * zx;  // o x z
      } else {
        mm[0] = zx; mm[1] = zy; mm[2] = zz;
      }
    }
    reinterpret_cast<float4*>(C.tw[k])[0] = make_float4((float)om[0], (float)om[1], (float)om[2], 0.f);
    reinterpret_cast<float4*>(C.tw[k])[1] = make_float4((float)mm[0], (float)mm[1], (float)mm[2], 0.f);
  }
  FK_MARK(5);
  const bool is_goal = (t == p.T - 1), is_stand = (p.use_standoff && t == p.knot_standoff);
  if (hl < 12) {  // gripper link frame and its difference to the two goal frames, formed in float64: one entry per lane
    const int r = hl >> 2, c = hl & 3;
    double F;
    if (R.grip_mov < 0) {
      F = R.grip_tf_d[hl];
    } else {
      const double* P = Tm + R.grip_mov * 12;
      F = dot3<double>(P[r * 4 + 0], R.grip_tf_d[c], P[r * 4 + 1], R.grip_tf_d[4 + c], P[r * 4 + 2], R.grip_tf_d[8 + c]);
      if (c == 3) F += P[r * 4 + 3];
    }
    C.gripf[hl] = (float)F;
    C.goal[0][hl] = (float)(F - p.goal_tf[(long long)b * 24 + hl]);
    C.goal[1][hl] = (float)(F - p.goal_tf[(long long)b * 24 + 12 + hl]);
  }
  if (hl == 15) {
    C.b = b; C.t = t; C.fid = fid; C.obuf = obuf;
    C.nact = __popc(amask);
    C.kind = (is_goal ? 1 : 0) | (is_stand ? 2 : 0);
    C.amask = amask;
    C.pad_ = 0;
  }
  if (hl < 3) C.basep[hl] = (hl == 0) ? bpx : (hl == 1 ? bpy : bpz);
  if (valid && amask == 0u && !is_goal && !is_stand) {  // nothing for the point kernel: this knot's Gauss-Newton block is zero
    const int nH = nopt * nopt;
    const long long bt = (long long)b * p.T + t;
    for (int i = hl; i < nH; i += 16) p.H[obuf * p.buf_stride_H + bt * nH + i] = 0.f;
    if (hl < nopt) p.g[obuf * p.buf_stride_g + bt * nopt + hl] = 0.0;
    if (hl == 0) p.costp[obuf * p.buf_stride_c + bt] = 0.0;
  }
  FK_MARK(6);
  if (pp.stats && valid && hl == 0) {
    atomicAdd(pp.stats + 0, 1ull);
    if (p.collision) atomicAdd(pp.stats + 1, (unsigned long long)nlinks);
    atomicAdd(pp.stats + 2, (unsigned long long)__popc(amask));
  }
}

__global__ void __launch_bounds__(128) k_item_fk(const __grid_constant__ CullParams pp) {
  extern __shared__ __align__(16) unsigned char fk_smem[];
  const LinParams& p = pp.lin;
  // the robot table is staged in shared memory with one batch of loads: the serial FK chain then never waits on L2
  RobotDev& R = *reinterpret_cast<RobotDev*>(fk_smem);
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(p.robot);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(fk_smem);
    for (int i = threadIdx.x; i < (int)(sizeof(RobotDev) / 8); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  pdl_wait();  // the trial point / active list of the step kernel before us
  pdl_trigger();  // the successor may be scheduled from here on (it blocks in its own pdl_wait until we are done)
  FK_MARK(0);
  stamp_begin(pp.ts_fk);
  const int nprob = p.nactive ? *p.nactive : p.nproblems;
  __syncthreads();
  const int hl = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const int hshift = threadIdx.x & 16;
  double* A = reinterpret_cast<double*>(fk_smem + ((sizeof(RobotDev) + 15) & ~(size_t)15)) + (size_t)grp * 2 * R.nmov * 12;
  double* Tm = A + (size_t)R.nmov * 12;
  const int nknots = p.T - p.t_lo;
  const int nitems = nprob * nknots;
  int item = blockIdx.x * (blockDim.x >> 4) + grp;
  if ((item & ~1) >= nitems) return;  // the whole warp (two items) leaves together
  const bool valid = item < nitems;
  if (!valid) item = nitems - 1;
  const int a = item / nknots;
  const int t = p.t_lo + (item - a * nknots);
  const int b = p.active ? p.active[a] : a;
  const double* q = p.q + ((long long)b * p.T + t) * R.ndof;
  const int obuf = p.bufsel ? (1 - p.bufsel[b]) : 0;
  item_fk_body(pp, pp.recs, R, item, valid, b, t, obuf, q, A, Tm, hl, hshift);
  FK_MARK(7);
  stamp_end(pp.ts_fk);
}

__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Warp sum of NP values per lane by recursive halving (see the use in cull_body).  On return gacc[0] of lane l holds the total of
// value warp_multi_owner<NP>(l) when that is >= 0.
template <int NP>
__device__ __forceinline__ void warp_reduce_multi(double (&v)[NP], int lane) {
  static_assert(NP == 8 || NP == 16, "8 or 16 values");
  int m = 16;
#pragma unroll
  for (int half = NP / 2; half >= 1; half >>= 1, m >>= 1) {
    const bool upper = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  for (; m >= 1; m >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
}
template <int NP>
__device__ __forceinline__ int warp_multi_owner(int lane) {
  if (NP == 8) return (lane & 3) == 0 ? ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1) : -1;
  return (lane & 1) == 0 ? ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1) : -1;
}

// per-warp reduction record: [nopt*nopt] float (J^T J), then [nopt + 1] double (J^T r, sum r^2), 8-byte aligned
__host__ __device__ inline int cull_red_floats(int nopt) { return ((nopt * nopt + 1) & ~1) + 2 * (nopt + 1); }
__host__ __device__ inline size_t cull_smem_bytes(int nopt, int ncons, int nslot, int slot_floats, int npad) {
  const int RS = nopt + 1;
  size_t sm = (sizeof(CullShared) + 127) & ~(size_t)127;
  sm += CULL_ZERO_BYTES;
  sm += (size_t)nslot * slot_floats * sizeof(float);
  sm += (size_t)ncons * (((32 * RS + 16 + 31) / 32) * 32) * sizeof(float);
  sm += (size_t)2 * ncons * cull_red_floats(nopt) * sizeof(float);
  sm = (sm + 127) & ~(size_t)127;
  sm += (size_t)3 * npad * sizeof(float);  // surface points
  return (sm + 127) & ~(size_t)127;
}

// NP: padded tensor-core tile width (8 or 16); NOPT_CT: number of optimised joints when known at compile time (0: runtime)
// cull_body: the linearisation of a list of (problem, knot) items whose records are recs[0 .. nitems).  Two callers:
//   k_linearize_cull  one launch per iteration for the whole batch: items are handed out through the global counter
//                     pp.work_counter (FUSED = false; nitems is read from the device-side active count)
//   k_solve_fused     (solve_fused.cuh) one CTA per problem runs the whole solver loop: the CTA's own items, counted locally
//                     (FUSED = true; rows_slot = the CTA's slot in the row buffer)
// smem_raw: cull_smem_bytes() bytes, 128-byte aligned; every warp of the CTA must call it ((NC + 2) warps).
template <int NP, int NOPT_CT, bool FUSED>
__device__ __forceinline__ void cull_body(const CullParams& pp, unsigned char* smem_raw, const RobotDev& R, const CullCtx* recs, int nitems_fused,
                                          int rows_slot) {
  const LinParams& p = pp.lin;
  CullShared& S = *reinterpret_cast<CullShared*>(smem_raw);
  const int nopt = NOPT_CT ? NOPT_CT : R.nopt, RS = nopt + 1, NC = pp.ncons;
  const unsigned CULL_NSLOT = (unsigned)pp.nslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  size_t off = (sizeof(CullShared) + 127) & ~(size_t)127;
  float* zero_buf = reinterpret_cast<float*>(smem_raw + off);
  off += CULL_ZERO_BYTES;
  float* ring = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)CULL_NSLOT * pp.slot_floats * sizeof(float);
  const int st_floats = ((32 * RS + 16 + 31) / 32) * 32;
  float* stage_base = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)NC * st_floats * sizeof(float);
  const int red_floats = cull_red_floats(nopt), red_g0 = (nopt * nopt + 1) & ~1;  // doubles start at an even float index
  float* red_base = reinterpret_cast<float*>(smem_raw + off);  // [2][NC][red_floats]
  off += (size_t)2 * NC * red_floats * sizeof(float);
  off = (off + 127) & ~(size_t)127;
  float* spts = reinterpret_cast<float*>(smem_raw + off);      // [3][npad] surface points x | y | z
  const float *spx = spts, *spy = spts + p.npad, *spz = spts + 2 * p.npad;
  uint64_t* slot_full = reinterpret_cast<uint64_t*>(S.slot_full);
  uint64_t* slot_empty = reinterpret_cast<uint64_t*>(S.slot_empty);
  uint64_t* ctx_full = reinterpret_cast<uint64_t*>(S.ctx_full);
  uint64_t* ctx_empty = reinterpret_cast<uint64_t*>(S.ctx_empty);
  uint64_t* zq_full = reinterpret_cast<uint64_t*>(S.zq_full);
  uint64_t* zq_empty = reinterpret_cast<uint64_t*>(S.zq_empty);
  uint64_t* pts_full = reinterpret_cast<uint64_t*>(&S.pts_full);

  if (threadIdx.x == 0) {
    for (int s = 0; s < (int)CULL_NSLOT; ++s) {
      mbar_init(slot_full + s, 1);
      mbar_init(slot_empty + s, NC);
    }
    for (int c = 0; c < CULL_NCTX; ++c) {
      mbar_init(ctx_full + c, 1);
      mbar_init(ctx_empty + c, NC);
    }
    for (int s = 0; s < CULL_ZQ; ++s) {
      mbar_init(reinterpret_cast<uint64_t*>(S.zq_full) + s, 1);
      mbar_init(reinterpret_cast<uint64_t*>(S.zq_empty) + s, 1);
    }
    mbar_init(reinterpret_cast<uint64_t*>(&S.pts_full), 1);
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
  for (int i = threadIdx.x; i < CULL_ZERO_BYTES / 4; i += blockDim.x) zero_buf[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the zeros are read by the async proxy (bulk stores)
  __syncthreads();
  // the active count is written by the step kernel: two launches back (long complete, read it ahead of the wait) when a
  // k_item_fk launch sits in between, directly before us when the step kernel wrote the item records itself
  int nitems = nitems_fused;
  if (!FUSED) {
    int nprob = p.nproblems;
    if (p.nactive && pp.count_early) nprob = *p.nactive;
    pdl_wait();  // the item records and everything before them
    pdl_trigger();  // the successor may be scheduled from here on (it blocks in its own pdl_wait until we are done)
    if (p.nactive && !pp.count_early) nprob = *p.nactive;
    stamp_begin(pp.ts_lin);
    nitems = nprob * (p.T - p.t_lo);
  }
  const int nH = nopt * nopt;

  if (warp == NC) {
    // =============================== PRODUCER WARP ===============================
    // per item: read the header of its record, write the zero rows of the culled links (bulk stores), and -- if anything
    // is left for the point kernel -- pull the record into shared memory (one bulk copy) and the surviving links' bricks
    // into the ring (one TMA tile each).  The next item index is fetched one item ahead.
    unsigned pub = 0, bc = 0, zq_tail = 0;
    int next_item = 0;
    if (lane == 0) {  // the robot's surface point set, staged once per call by ONE bulk copy (x | y | z, 16-byte padded)
      const uint32_t bytes = (uint32_t)(3 * p.npad * sizeof(float));
      mbar_expect_tx(pts_full, bytes);
      bulk_load(spts, p.pts3, bytes, pts_full);
    }
    if (!FUSED && lane == 0) next_item = atomicAdd(pp.work_counter, 1);
    for (;;) {
      const int item = FUSED ? next_item : __shfl_sync(0xffffffffu, next_item, 0);
      if (item >= nitems) break;
      if (FUSED) ++next_item;
      else if (lane == 0) next_item = atomicAdd(pp.work_counter, 1);
      const CullCtx* G = recs + item;  // (records are rewritten every iteration: __ldcg, never the non-coherent read-only path)
      const int4 h0 = __ldcg(reinterpret_cast<const int4*>(&G->b));     // b, t, fid, obuf
      const int4 h1 = __ldcg(reinterpret_cast<const int4*>(&G->nact));  // nact, kind, amask, -
      int4 mydim = make_int4(0, 0, 0, 0), mylo = make_int4(0, 0, 0, 0);
      const int b = h0.x, t = h0.y, fid = h0.z;
      const unsigned amask = (unsigned)h1.z;
      const bool survives = (amask >> lane) & 1u;
      if (survives) {
        mydim = __ldcg(reinterpret_cast<const int4*>(G->bdim[lane]));
        mylo = __ldcg(reinterpret_cast<const int4*>(G->blo[lane]));
      }
      // ---- the zero rows of the culled links are handed to the zero-row warp (nobody waits for them) ----
      if (p.collision && p.rows) {
        if (lane == 0) {
          const unsigned zs = zq_tail % CULL_ZQ;
          if (zq_tail >= CULL_ZQ) mbar_wait_sleep(zq_empty + zs, ((zq_tail / CULL_ZQ) - 1) & 1);  // queue full: the zero-row warp is behind
          S.zq[zs] = make_int4(b, t, (int)amask, 0);
          mbar_arrive(zq_full + zs);  // (release: the entry is visible to the waiter)
        }
        ++zq_tail;
      }
      if (amask == 0u && h1.y == 0) continue;  // k_item_fk wrote the zero Gauss-Newton block
      const int ci = pub % CULL_NCTX;
      if (pub >= CULL_NCTX) mbar_wait_sleep(ctx_empty + ci, ((pub / CULL_NCTX) - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(ctx_full + ci, (uint32_t)sizeof(CullCtx));
        bulk_load(&S.ctx[ci], G, (uint32_t)sizeof(CullCtx), ctx_full + ci);
      }
      ++pub;
      // ---- one TMA brick per surviving link into the ring ----
      if (amask) {
        const CUtensorMap* maps = p.fields[fid].maps2;
        unsigned m = amask, idx = bc;
        while (m) {
          const int l = __ffs(m) - 1;
          m &= m - 1;
          const int sx = __shfl_sync(0xffffffffu, mydim.x, l), sy = __shfl_sync(0xffffffffu, mydim.y, l), sz = __shfl_sync(0xffffffffu, mydim.z, l);
          const int lx = __shfl_sync(0xffffffffu, mylo.x, l), ly = __shfl_sync(0xffffffffu, mylo.y, l), lz = __shfl_sync(0xffffffffu, mylo.z, l);
          if (lane == 0) {
            const unsigned s = idx % CULL_NSLOT;
            if (idx >= CULL_NSLOT) mbar_wait_sleep(slot_empty + s, ((idx / CULL_NSLOT) - 1) & 1);
            const int mi = ((sx / 4 - 2) * CULL_NAXC + (sy / 4 - 2)) * CULL_NAXC + (sz / 4 - 2);
            mbar_expect_tx(slot_full + s, (uint32_t)(sx * sy * sz * sizeof(float)));
            tma_load_3d(ring + (size_t)s * pp.slot_floats, maps + mi, lz, ly, lx, slot_full + s);
          }
          ++idx;
        }
        bc += __popc(amask);
        __syncwarp();
      }
    }
    // ---- tell the consumers to stop, drain the bulk stores ----
    {
      const int ci = pub % CULL_NCTX;
      if (pub >= CULL_NCTX) mbar_wait_sleep(ctx_empty + ci, ((pub / CULL_NCTX) - 1) & 1);
      if (lane == 0) {
        S.ctx[ci].b = -1;
        mbar_arrive(ctx_full + ci);
      }
    }
    if (p.collision && p.rows && lane == 0) {  // stop record for the zero-row warp
      const unsigned zs = zq_tail % CULL_ZQ;
      if (zq_tail >= CULL_ZQ) mbar_wait_sleep(zq_empty + zs, ((zq_tail / CULL_ZQ) - 1) & 1);
      S.zq[zs] = make_int4(-1, 0, 0, 0);
      mbar_arrive(zq_full + zs);
    }
    if (!FUSED) stamp_end(pp.ts_lin);
    return;
  }
  if (warp == NC + 1) {
    // =============================== ZERO-ROW WARP ===============================
    // rows of the culled links: zeros, straight from shared memory by bulk copy, decoupled from the producer so that the HBM
    // write stream keeps running while the producer waits for ring slots / contexts
    if (!(p.collision && p.rows)) return;
    const int nlinks = R.nlinks;
    int my_start = 0, my_cnt = 0;
    if (lane < nlinks) { my_start = S.links[lane].pt_start; my_cnt = S.links[lane].pt_end - my_start; }
    for (unsigned head = 0;; ++head) {
      const unsigned zs = head % CULL_ZQ;
      mbar_wait_sleep(zq_full + zs, (head / CULL_ZQ) & 1);
      const int4 e = S.zq[zs];
      __syncwarp();
      if (lane == 0) mbar_arrive(zq_empty + zs);
      const int b = e.x, t = e.y;
      if (b < 0) break;
      const unsigned amask = (unsigned)e.z;
      const bool survives = (amask >> lane) & 1u;
      {
        float* rows_b = p.rows + (long long)(FUSED ? rows_slot : b - p.b0) * p.rows_per_problem * RS;
        bool slow_zero = false;
        if (lane < nlinks && !survives) {
          char* dst = reinterpret_cast<char*>(rows_b + ((long long)t * R.npoints + my_start) * RS);
          const unsigned bytes = (unsigned)my_cnt * RS * 4u;
          if ((((uintptr_t)dst | bytes) & 15u) == 0u) {
            for (unsigned o = 0; o < bytes; o += CULL_ZERO_BYTES) bulk_store_zero(dst + o, zero_buf, min((unsigned)CULL_ZERO_BYTES, bytes - o));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          } else {
            slow_zero = my_cnt > 0;
          }
        }
        unsigned zm = __ballot_sync(0xffffffffu, slow_zero);  // row blocks that are not 16-byte aligned: plain stores
        while (zm) {
          const int l = __ffs(zm) - 1;
          zm &= zm - 1;
          float* dst = rows_b + ((long long)t * R.npoints + S.links[l].pt_start) * RS;
          const int nfl = (S.links[l].pt_end - S.links[l].pt_start) * RS;
          for (int i = lane; i < nfl; i += 32) __stcs(dst + i, 0.f);
        }
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (!FUSED) stamp_end(pp.ts_lin);
    return;
  }

  // =============================== CONSUMER WARPS ===============================
  float* stage = stage_base + warp * st_floats;
  const int gq = lane >> 2, tq = lane & 3;
  const bool obs_linear = (p.flags & GTO_FLAG_OBS_LINEAR) != 0;
  unsigned ic = 0, bc = 0;
  mbar_wait_sleep(pts_full, 0);
  for (;; ++ic) {
    const int ci = ic % CULL_NCTX, ri = ic & 1;  // record slot, reduction buffer
    mbar_wait_sleep(ctx_full + ci, (ic / CULL_NCTX) & 1);
    const CullCtx& C = S.ctx[ci];
    const int b = C.b;
    if (b < 0) break;
    const int t = C.t, fid = C.fid, nact = C.nact;
    const bool is_goal = (C.kind & 1) != 0, is_stand = (C.kind & 2) != 0;
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    // J^T r and sum r^2 are accumulated in float64: at the optimum J^T r is the small difference of large sums (goal pull against
    // obstacle push), and the float32 rounding of a running sum of magnitude ~10 (6e-7 per add) would move the fixed point
    // by more than the 1e-4 rad parity bar along the weakly determined directions of the redundant arm
    double gacc[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) gacc[k] = 0.0;
    double cacc = 0.0;
    float* rows_b = p.rows ? p.rows + (long long)(FUSED ? rows_slot : b - p.b0) * p.rows_per_problem * RS : nullptr;

    int next = warp, cb = 0;  // chunks of the surviving links are dealt round-robin: warp, warp+NC, ... of the concatenation
    for (int ai = 0; ai < nact; ++ai) {
      const int l = C.act[ai];
      // every consumer warp observes every brick (full) before it releases it (empty), chunks or not: an early release
      // of a later use of the same slot could otherwise complete the empty barrier of the current use
      const unsigned idx = bc + ai;
      mbar_wait_sleep(slot_full + (idx % CULL_NSLOT), (idx / CULL_NSLOT) & 1);
      const float* brick = ring + (size_t)(idx % CULL_NSLOT) * pp.slot_floats;
      const double* vf = C.vf[l];
      const float ipitch = p.fields[fid].inv_pitch;
      const int dy = C.bdim[l][1], dz = C.bdim[l][2], fast = C.bdim[l][3];
      const int dxm2 = C.bdim[l][0] - 2, dym2 = dy - 2, dzm2 = dz - 2;
      const LinkMeta lm = S.links[l];
      const int nch = lm.c1 - lm.c0;
      if (next < cb + nch) {
        const unsigned mask = lm.mask;
        float F[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) F[e] = C.frames[l][e];
        for (; next < cb + nch; next += NC) {
          const int p0 = lm.pt_start + 32 * (next - cb), cnt = min(32, lm.pt_end - p0);
          const bool act = lane < cnt;
          float J[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) J[k] = 0.f;
          float r = 0.f;
          if (act) {
            const float x = spx[p0 + lane], y = spy[p0 + lane], z = spz[p0 + lane];
            const float wbx = F[0] * x + F[1] * y + F[2] * z + F[3];
            const float wby = F[4] * x + F[5] * y + F[6] * z + F[7];
            const float wbz = F[8] * x + F[9] * y + F[10] * z + F[11];
            float val, gx, gy, gz;
            // brick-local voxel coordinate in float64 (see k_item_fk), cell index, float32 fraction inside the cell
            const double xd = (double)x, yd = (double)y, zd = (double)z;
            const double ux = fma(vf[0], xd, fma(vf[1], yd, fma(vf[2], zd, vf[3])));
            const double uy = fma(vf[4], xd, fma(vf[5], yd, fma(vf[6], zd, vf[7])));
            const double uz = fma(vf[8], xd, fma(vf[9], yd, fma(vf[10], zd, vf[11])));
            if (fast) {
              // the brick encloses every point of this link and lies inside the grid: no grid clamping can occur
              const int ix = min(max((int)floor(ux), 0), dxm2), iy = min(max((int)floor(uy), 0), dym2), iz = min(max((int)floor(uz), 0), dzm2);
              const float fx = (float)(ux - (double)ix), fy = (float)(uy - (double)iy), fz = (float)(uz - (double)iz);
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
              sdf_trilinear_box(p.fields[fid], brick, C.bdim[l], C.blo[l], ux, uy, uz, val, gx, gy, gz);
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
          if (act) {  // (inactive lanes hold zeros)
            const double rd = (double)r;
            if (!obs_linear) {
              cacc = fma(rd, rd, cacc);
#pragma unroll
              for (int k = 0; k < NP; ++k)
                if (k < nopt) gacc[k] = fma((double)J[k], rd, gacc[k]);
            } else {  // unsquared term w*c (gto/ik_solver.py:69): value w*c, half gradient (w/2) dc/dq, no curvature
              cacc = fma((double)p.sw_obs, rd, cacc);
#pragma unroll
              for (int k = 0; k < NP; ++k)
                if (k < nopt) gacc[k] = fma(0.5 * (double)p.sw_obs, (double)J[k], gacc[k]);
            }
          }
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
          if (!obs_linear) mma_rows<NP>(stage, RS, cnt, acc0, acc1, lane);
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
      cb += nch;
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_empty + (idx % CULL_NSLOT));
    }
    bc += nact;

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
            const float x = spx[pi], y = spy[pi], z = spz[pi];
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
            cacc = fma((double)r, (double)r, cacc);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
              if (k < nopt) {
                gacc[k] = fma((double)J[k], (double)r, gacc[k]);
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
    // sum over the 32 lanes by recursive halving: at every stage a lane keeps half of its values and hands the other half to its
    // partner (NP/2 + NP/4 + ... + 1 shuffles, then plain butterfly steps; 9 instead of 40 for 8 values).  sum r^2 rides in the
    // first free slot when nopt < NP.  Afterwards value k sits in the lanes selected by warp_multi_owner.
    constexpr bool kCostRides = NOPT_CT > 0 && NOPT_CT < NP;
    if (kCostRides) gacc[kCostRides ? NOPT_CT : 0] = cacc;
    warp_reduce_multi<NP>(gacc, lane);
    if (!kCostRides) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cacc += __shfl_xor_sync(0xffffffffu, cacc, o);
    }
    {
      float* red = red_base + ((size_t)ri * NC + warp) * red_floats;
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
      {
        double* redd = reinterpret_cast<double*>(red + red_g0);
        const int kown = warp_multi_owner<NP>(lane);  // the value this lane holds in gacc[0] (< 0: none)
        if (kCostRides) {
          if (kown >= 0 && kown <= nopt) redd[kown] = gacc[0];
        } else {
          if (kown >= 0 && kown < nopt) redd[kown] = gacc[0];
          if (lane == 0) redd[nopt] = cacc;
        }
      }
    }
    const int obuf = C.obuf;
    asm volatile("bar.sync 1, %0;" ::"r"(NC * 32) : "memory");  // consumers only; the producer keeps running ahead
    {
      const int ntot = nH + nopt + 1;
      const long long bt = (long long)b * p.T + t;
      for (int i = threadIdx.x; i < ntot; i += NC * 32) {
        if (i < nH) {
          float s = 0.f;
          for (int w = 0; w < NC; ++w) s += red_base[((size_t)ri * NC + w) * red_floats + i];
          p.H[obuf * p.buf_stride_H + bt * nH + i] = s;
        } else {
          double s = 0.0;
          for (int w = 0; w < NC; ++w) s += reinterpret_cast<const double*>(red_base + ((size_t)ri * NC + w) * red_floats + red_g0)[i - nH];
          if (i < nH + nopt) p.g[obuf * p.buf_stride_g + bt * nopt + (i - nH)] = s;
          else p.costp[obuf * p.buf_stride_c + bt] = s;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ctx_empty + ci);  // this warp no longer reads ctx[ci] / red[ci]
  }
  if (!FUSED) stamp_end(pp.ts_lin);
}

template <int NP, int NOPT_CT>
__global__ void __launch_bounds__((CULL_MAX_CONS + 2) * 32, 2) k_linearize_cull(const __grid_constant__ CullParams pp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cull_body<NP, NOPT_CT, false>(pp, smem_raw, *pp.lin.robot, pp.recs, 0, 0);
}
