// Advection-path kernels (sm_100a): per-cell density sums, the 1-D Poisson field integrals, the
// DG upwind right-hand side with the SSP-RK3 stage combination, the periodic x halo, and the
// per-step moments.  Reference: advection_1.cpp, FieldCalculations.cpp, MomentCalculations.cpp
// under /root/reference/source.
//
// State layout: plane-major  U[(p*6 + c)*sv + j],  p = local x cell + 1 (planes 0 and ncell+1 are
// the x halos), c = DG coefficient 0..5, j = j1*Nv^2 + j2*Nv + j3.  Every access below is
// unit-stride in j across a warp; an x-plane is one contiguous 6*sv block, so a halo is one copy.
#include "lpgpu_internal.h"
#include <numeric>

#define LP_LAUNCHED(c)                                  \
  do {                                                  \
    (c)->launches++;                                    \
    LP_CUDA(cudaGetLastError());                        \
  } while (0)

// fixed-order block sum of NV values per thread (deterministic run to run)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* NV*32 */)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  #pragma unroll
  for (int m = 0; m < NV; m++) {
    double x = v[m];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) smem[m * 32 + wid] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    #pragma unroll
    for (int m = 0; m < NV; m++) { double x = 0.; for (int w = 0; w < nw; w++) x += smem[m * 32 + w]; v[m] = x; }
  }
}

// ---------------------------------------------------------------------------------------------
// m_i = scalev * sum_j (U0 + U5/4),  s_i = scalev * sum_j U1  -- the only quantities the nested
// loops of computePhi_x_0_Normal / computeC_rho / Int_E* (FieldCalculations.cpp:223-243, 60-73,
// 356-409) depend on.
__global__ void __launch_bounds__(256) k_field_reduce(const double *__restrict__ planes, double *__restrict__ ms, int sv, double scalev)
{
  __shared__ double red[2 * 32];
  const long long cell = blockIdx.x;
  const double *u = planes + ((cell + 1) * 6) * (long long)sv;
  double v[2] = {0., 0.};
  for (int j = threadIdx.x; j < sv; j += blockDim.x) {
    v[0] += u[j] + u[5LL * sv + j] / 4.;
    v[1] += u[1LL * sv + j];
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) { ms[2 * cell] = v[0] * scalev; ms[2 * cell + 1] = v[1] * scalev; }
}
// large velocity grids: LP_FR_CH blocks per x cell (one block per cell would use ncell of 148 SMs), partials
// folded in a fixed order
#define LP_FR_CH 32
__global__ void __launch_bounds__(256) k_field_reduce_part(const double *__restrict__ planes, double *__restrict__ part, int sv)
{
  __shared__ double red[2 * 32];
  const long long cell = blockIdx.x;
  const double *u = planes + ((cell + 1) * 6) * (long long)sv;
  const int per = (sv + LP_FR_CH - 1) / LP_FR_CH, lo = blockIdx.y * per, hi = min(sv, lo + per);
  double v[2] = {0., 0.};
  for (int j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    v[0] += u[j] + u[5LL * sv + j] / 4.;
    v[1] += u[1LL * sv + j];
  }
  block_sum<2>(v, red);
  if (threadIdx.x == 0) { part[2 * (cell * LP_FR_CH + blockIdx.y)] = v[0]; part[2 * (cell * LP_FR_CH + blockIdx.y) + 1] = v[1]; }
}
__global__ void k_field_reduce_fold(const double *__restrict__ part, double *__restrict__ ms, int ncell, double scalev)
{
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncell) return;
  double a = 0., b = 0.;
  for (int k = 0; k < LP_FR_CH; k++) { a += part[2 * (cell * LP_FR_CH + k)]; b += part[2 * (cell * LP_FR_CH + k) + 1]; }
  ms[2 * cell] = a * scalev; ms[2 * cell + 1] = b * scalev;
}
int lp_launch_field_reduce(lpgpu_ctx *c, const double *planes)
{
  if (c->sv < 8192) {
    k_field_reduce<<<c->ncell, 256, 0, c->stream>>>(planes, c->d_ms_local, c->sv, c->tab.scalev);
    LP_LAUNCHED(c);
    return LPGPU_OK;
  }
  k_field_reduce_part<<<dim3(c->ncell, LP_FR_CH), 256, 0, c->stream>>>(planes, c->d_ms_part, c->sv);
  LP_LAUNCHED(c);
  k_field_reduce_fold<<<(c->ncell + 127) / 128, 128, 0, c->stream>>>(c->d_ms_part, c->d_ms_local, c->ncell, c->tab.scalev);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// Closed forms of ce = computePhi_x_0, cp = computeC_rho, intE = Int_E, intE1 = Int_E1st,
// intE2 = Int_E2nd (advection_1.cpp:419-429) from the gathered (m_q, s_q), q = 0..Nx-1.  The scan is
// Nx long (<= a few hundred) and sequential on purpose: ce is a catastrophic cancellation
// (Lx/2 - O(Lx/2)), so every GPU evaluates it in the same fixed order.
// Doping = True: computePhi_x_0_Doping, Int_E_Doping, Int_E1st_Doping, Int_E2nd_Doping (FieldCalculations.cpp:427-450,
// 585-676) with the step profile ND(i) = NH for i <= a_i or i > b_i, NL between (:413-425)
struct DopingParams { int on, a_i, b_i; double NL, NH, eps; };
// body of the scan, shared by k_field_scan and the fused k_field_finish; every thread of the (single) block calls it
__device__ __forceinline__ void field_scan_body(const double *ms_all, double *__restrict__ fld, int Nx, int x_begin, int x_count,
                                                double dx, double Lx, const DopingParams &D, double *sms, double *s_ce)
{
  double *sm = sms, *s12 = sms + Nx, *sP = sms + 2 * Nx;
  for (int q = threadIdx.x; q < Nx; q += blockDim.x) { sm[q] = ms_all[2 * q]; s12[q] = ms_all[2 * q + 1] / 12.; }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the serial part keeps the summation order of the one-thread scan: only additions are left on the chain
    double P = 0., acc = 0.;
    for (int q = 0; q < Nx; q++) { sP[q] = P; acc += P + 0.5 * sm[q] - s12[q]; P += sm[q]; }
    if (D.on) {
      const double a_val = (D.a_i + 1) * dx, b_val = (D.b_i + 1) * dx, Phi_Lx = 1, tmp = acc * dx * dx;
      *s_ce = Phi_Lx / Lx + 0.5 * D.NH * Lx / D.eps + (D.NL - D.NH) * (b_val - a_val) / D.eps
             - (0.5 * (D.NL - D.NH) * (b_val * b_val - a_val * a_val) + tmp) / (Lx * D.eps);
    } else {
      *s_ce = 0.5 * Lx - acc * dx * dx / Lx;
    }
    fld[0] = *s_ce;
  }
  __syncthreads();
  const double ce = *s_ce;
  for (int q = x_begin + threadIdx.x; q < x_begin + x_count; q += blockDim.x) {
    const double m = sm[q], s = ms_all[2 * q + 1], P = sP[q];
    const double xi = (q + 0.5) * dx, xl = ((q - 0.5) + 0.5) * dx, c2 = s * dx / 2., cp = dx * P;
    double *o = fld + 1 + 4 * (q - x_begin);
    o[0] = cp;
    if (D.on) {
      const double ND = (q <= D.a_i || q > D.b_i) ? D.NH : D.NL, a_val = (D.a_i + 1) * dx, b_val = (D.b_i + 1) * dx;
      double r = -(P + 0.5 * m - s12[q]) * dx * dx + ND * xi * dx;
      if (q > D.a_i) r += (D.NH - D.NL) * a_val * dx;
      if (q > D.b_i) r += (D.NL - D.NH) * b_val * dx;
      o[1] = r / D.eps - ce * dx;
      o[2] = (ND - m) * dx * dx / (12. * D.eps);
      r = (-cp + (m * xl + 0.25 * c2)) * dx / 12. + (ND - m) * dx * xi / 12. - c2 * dx / 80.;
      if (q > D.a_i) r += (D.NH - D.NL) * a_val * dx / 12.;
      if (q > D.b_i) r += (D.NL - D.NH) * b_val * dx / 12.;
      o[3] = r / D.eps - ce * dx / 12.;
      continue;
    }
    o[1] = -ce * dx - (P + 0.5 * m - s12[q]) * dx * dx + xi * dx;
    o[2] = (1 - m) * dx * dx / 12.;
    o[3] = (-cp - ce + (m * xl + 0.25 * c2)) * dx / 12. + (1 - m) * dx * xi / 12. - c2 * dx / 80.;
  }
}
__global__ void __launch_bounds__(256) k_field_scan(const double *__restrict__ ms_all, double *__restrict__ fld, int Nx, int x_begin, int x_count,
                                                    double dx, double Lx, DopingParams D)
{
  extern __shared__ double sms[];          // m[Nx] | s/12 [Nx] | prefix P[Nx]
  __shared__ double s_ce;
  field_scan_body(ms_all, fld, Nx, x_begin, x_count, dx, Lx, D, sms, &s_ce);
}
int lp_launch_field_scan(lpgpu_ctx *c)
{
  const double dx = c->p.Lx / c->p.Nx;
  DopingParams D = {c->p.doping, c->p.Nx / 3 - 1, 2 * c->p.Nx / 3 - 1, c->p.NL, c->p.NH, c->p.eps};   // a_i, b_i: LP_ompi.cpp:160-161
  const size_t scan_smem = (size_t)3 * c->p.Nx * sizeof(double);
  if (scan_smem > 48 * 1024 && !c->scan_attr) {       // Nx > 2048 (lpgpu_init bounds Nx so that this fits an SM)
    LP_CUDA(cudaFuncSetAttribute(k_field_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
    c->scan_attr = true;
  }
  k_field_scan<<<1, 256, scan_smem, c->stream>>>(c->d_ms_all, c->d_fld, c->p.Nx, c->p.x_begin, c->ncell, dx, c->p.Lx, D);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// One SSP-RK3 stage (advection_1.cpp:436-455 / 488-507 / 538-557): per DG cell
//   tp_l = I1 - I2 - I3 + I5   (I1,I2: :72-103; I3_Normal: :284-321; I5: :323-390)
//   H = M^-1 tp / (dx scalev)  (:449-452)
//   stage 0: U1 = U + dt H(U);  stage 1: U2 = 3/4 U + 1/4 U1 + 1/4 dt H(U1);
//   stage 2: U  = 1/3 U + 2/3 U2 + 2/3 dt H(U2)
struct DgParams {
  int Nv, sv, ncell;
  double dv, dx, dt, scalev, Lv, inv_dxs;
};
// Two DG cells (j, j + 1: the same j1, j2) per thread, every access a double2: half the load instructions for the same
// bytes and twice the bytes in flight per thread -- the kernel is bound by memory latency, not bandwidth (ncu r02s: long
// scoreboard 62 % of the warp samples at 43 % of the DRAM throughput).  U0 / Uout may be the same buffer in stage 2 (each
// thread reads its old values before it writes them): they are not declared __restrict__.
struct DgCell { double tp[6]; };
template <int STAGE>
__global__ void __launch_bounds__(256) k_dg_stage(const double *__restrict__ Uin, const double *U0,
                                                  double *Uout, const double *__restrict__ fld, DgParams P)
{
  const long long t = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= (long long)P.ncell * P.sv) return;
  const int sv = P.sv, Nv = P.Nv, NN = Nv * Nv;
  const long long cell = t / sv; const int j = (int)(t % sv), j1 = j / NN;
  // the reference's divisions by 12, 720, dx*scalev ... are multiplications by the correctly rounded reciprocals
  // here (1 ulp per operation away from the reference; an FP64 division costs ~20 FP64-pipe slots on this GPU)
  constexpr double K12 = 1. / 12.;
  const double dv = P.dv, dv2 = dv * dv, dv3 = dv2 * dv;
  const double c1 = -P.Lv + (j1 + 0.5) * dv;
  const double *f4 = fld + 1 + 4 * cell;
  const double E = f4[1], E1 = f4[2], E2 = f4[3];
  const long long own = ((cell + 1) * 6) * (long long)sv + j;
  auto ld2 = [&](const double *base, long long off) { return *reinterpret_cast<const double2 *>(base + off); };
  double2 u[6], X[6], V[6];
  #pragma unroll
  for (int c = 0; c < 6; c++) u[c] = ld2(Uin, own + (long long)c * sv);
  // x-upwind neighbour: plane p+1 for v1 < 0 (j1 < Nv/2), plane p-1 otherwise (halo planes at the ends)
  const bool right = j1 < Nv / 2;
  const long long nbx = right ? own + 6LL * sv : own - 6LL * sv;
  #pragma unroll
  for (int c = 0; c < 6; c++) X[c] = ld2(Uin, nbx + (long long)c * sv);
  // v1-upwind neighbour: j1+1 when the cell-integrated field is positive, j1-1 otherwise; no flux through |v1| = Lv
  const bool up = E > 0;
  const bool have_v = up ? (j1 + 1 < Nv) : (j1 > 0);
  const long long nbv = up ? own + NN : own - NN;
  #pragma unroll
  for (int c = 0; c < 6; c++) V[c] = have_v ? ld2(Uin, nbv + (long long)c * sv) : make_double2(0., 0.);
  double2 old[6];
  if (STAGE != 0) {
    #pragma unroll
    for (int c = 0; c < 6; c++) old[c] = ld2(U0, own + (long long)c * sv);
  }
  double2 res[6];
  #pragma unroll
  for (int h = 0; h < 2; h++) {
    double uu[6], xx[6], vv[6];
    #pragma unroll
    for (int c = 0; c < 6; c++) { uu[c] = h ? u[c].y : u[c].x; xx[c] = h ? X[c].y : X[c].x; vv[c] = h ? V[c].y : V[c].x; }
    double tp[6] = {0., 0., 0., 0., 0., 0.};
    // I1: only the phi_x test function sees the v1 f volume term
    tp[1] += dv3 * (c1 * uu[0] + dv * uu[2] * K12 + uu[5] * c1 * 0.25);
    // I2: E f volume term (Int_fE, FieldCalculations.cpp:126-135)
    tp[2] -= ((uu[0] + uu[5] * 0.25) * E + uu[1] * E1) * dv2;   // scalev/dv = dv^2
    tp[5] -= uu[2] * dv2 * E * (1. / 6.);
    // I3: x faces, upwind on the sign of the v1 cell index
    {
      double R[6], L[6], ur, ul;
      if (right) {
        #pragma unroll
        for (int c = 0; c < 6; c++) { R[c] = xx[c]; L[c] = uu[c]; }
        ur = -R[1]; ul = -L[1];
      } else {
        #pragma unroll
        for (int c = 0; c < 6; c++) { L[c] = xx[c]; R[c] = uu[c]; }
        ur = R[1]; ul = L[1];
      }
      tp[0] -= dv3 * ((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * c1 + (R[2] - L[2]) * dv * K12 + (R[5] - L[5]) * c1 * 0.25);
      tp[1] -= 0.5 * dv3 * ((R[0] + 0.5 * ur + L[0] + 0.5 * ul) * c1 + (R[2] + L[2]) * dv * K12 + (R[5] + L[5]) * c1 * 0.25);
      tp[2] -= dv2 * (((R[0] - L[0]) * dv2 + (ur - ul) * 0.5 * dv2 + (R[2] - L[2]) * dv * c1) * K12 + (R[5] - L[5]) * dv2 * (19. / 720.));
      tp[3] -= (R[3] - L[3]) * c1 * dv3 * K12;
      tp[4] -= (R[4] - L[4]) * c1 * dv3 * K12;
      tp[5] -= dv3 * ((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * c1 * 0.25 + (R[2] - L[2]) * dv * (19. / 720.) + (R[5] - L[5]) * c1 * (19. / 240.));
    }
    // I5: v1 faces, upwind on the sign of the cell-integrated field
    {
      double R[6], L[6], ur, ul;
      if (up) {
        #pragma unroll
        for (int c = 0; c < 6; c++) { L[c] = uu[c]; R[c] = vv[c]; }
        ul = -L[2]; ur = have_v ? -R[2] : 0.;
      } else {
        #pragma unroll
        for (int c = 0; c < 6; c++) { R[c] = uu[c]; L[c] = vv[c]; }
        ur = R[2]; ul = have_v ? L[2] : 0.;
      }
      const double gR = R[0] + 0.5 * ur + R[5] * (5. / 12.), gL = L[0] + 0.5 * ul + L[5] * (5. / 12.);
      tp[0] += dv2 * (gR - gL) * E + dv2 * (R[1] - L[1]) * E1;
      tp[1] += dv2 * ((gR - gL) * E1 + (R[1] - L[1]) * E2);
      tp[2] += 0.5 * (dv2 * (gR + gL) * E + dv2 * (R[1] + L[1]) * E1);
      tp[3] += (R[3] - L[3]) * E * dv2 * K12;
      tp[4] += (R[4] - L[4]) * E * dv2 * K12;
      tp[5] += dv2 * (((R[0] + 0.5 * ur - L[0] - 0.5 * ul) * (5. / 12.) + (R[5] - L[5]) * (133. / 720.)) * E + (R[1] - L[1]) * E1 * (5. / 12.));
    }
    const double idxs = P.inv_dxs;   // 1/(dx scalev)
    double H[6];
    H[0] = (19 * tp[0] * 0.25 - 15 * tp[5]) * idxs;
    H[5] = (60 * tp[5] - 15 * tp[0]) * idxs;
    #pragma unroll
    for (int l = 1; l < 5; l++) H[l] = tp[l] * (12. * idxs);
    #pragma unroll
    for (int c = 0; c < 6; c++) {
      const double o0 = (STAGE != 0) ? (h ? old[c].y : old[c].x) : 0.;
      double r;
      if (STAGE == 0) r = uu[c] + P.dt * H[c];
      else if (STAGE == 1) r = 0.75 * o0 + 0.25 * uu[c] + 0.25 * P.dt * H[c];
      else r = o0 * (1. / 3.) + uu[c] * (2. / 3.) + P.dt * H[c] * (2. / 3.);
      if (h) res[c].y = r; else res[c].x = r;
    }
  }
  #pragma unroll
  for (int c = 0; c < 6; c++) *reinterpret_cast<double2 *>(Uout + own + (long long)c * sv) = res[c];
}
int lp_launch_dg_stage(lpgpu_ctx *c, int stage)
{
  DgParams P;
  P.Nv = c->p.Nv; P.sv = c->sv; P.ncell = c->ncell; P.dv = c->tab.dv; P.dx = c->p.Lx / c->p.Nx; P.dt = c->p.dt;
  P.scalev = c->tab.scalev; P.Lv = c->p.Lv; P.inv_dxs = 1. / (P.dx * P.scalev);
  const long long n = (long long)c->ncell * c->sv;      // Nv is even: the two cells of a thread share j1 and j2
  const unsigned grid = (unsigned)((n / 2 + 255) / 256);
  const bool prof3 = c->prof_on == 3 && c->prof_used + 2 <= c->prof_ev.size();   // bench.py: HBM roofline of the DG stage kernels
  if (prof3) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  // buffers: stage 0 reads U -> writes U1; stage 1 reads U1 (+U) -> U2; stage 2 reads U2 (+U) -> U
  if (stage == 0) k_dg_stage<0><<<grid, 256, 0, c->stream>>>(c->d_U[0], c->d_U[0], c->d_U[1], c->d_fld, P);
  else if (stage == 1) k_dg_stage<1><<<grid, 256, 0, c->stream>>>(c->d_U[1], c->d_U[0], c->d_U[2], c->d_fld, P);
  else if (stage == 2) k_dg_stage<2><<<grid, 256, 0, c->stream>>>(c->d_U[2], c->d_U[0], c->d_U[0], c->d_fld, P);
  else { lp_set_error("lp_launch_dg_stage: stage must be 0, 1 or 2"); return LPGPU_EINVAL; }
  LP_LAUNCHED(c);
  if (prof3) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  return LPGPU_OK;
}

// Doping: I3_Doping takes the upwind state beyond a domain wall from DirichletBC (advection_1.cpp:230-233, 254-257);
// the wall planes are precomputed (tables.cpp) and copied into the halo planes that face a wall
int lp_launch_wall_halo(lpgpu_ctx *c, double *planes)
{
  if (!c->p.doping || c->p.homogeneous) return LPGPU_OK;
  const size_t plane = (size_t)6 * c->sv;
  if (c->p.x_begin == 0)
    LP_CUDA(cudaMemcpyAsync(planes, c->d_dirichlet, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (c->p.x_begin + c->ncell == c->p.Nx)
    LP_CUDA(cudaMemcpyAsync(planes + plane * (c->ncell + 1), c->d_dirichlet + plane, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return LPGPU_OK;
}
// periodic halo when one context owns all of x (advection_1.cpp:297, :306)
int lp_launch_local_halo(lpgpu_ctx *c, double *planes)
{
  if (c->p.doping) return lp_launch_wall_halo(c, planes);
  const size_t plane = (size_t)6 * c->sv;
  LP_CUDA(cudaMemcpyAsync(planes, planes + plane * c->ncell, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  LP_CUDA(cudaMemcpyAsync(planes + plane * (c->ncell + 1), planes + plane, plane * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange (replaces the MPI_Bcast / the NCCL all-gather + halo send/recv of the sharded run).
// Flags carry epochs that only grow, so nothing is ever reset: the publisher bumps its own epoch and stores it into the
// peers' flag words after its data (release at system scope); a consumer waits until every flag has reached its own
// epoch -- all ranks execute the same sequence of stages.  A rank can be at most one stage ahead of a peer (its next
// stage needs that peer's next densities), hence two density buffers selected by the epoch's parity; the halo planes
// belong to three different stage buffers.  Waits are bounded (lpgpu_peer_set_timeout, default 60 s): on a timeout the
// kernel counts it in the mailbox and poisons the state with NaN, so that a dead peer gives an error from the next host
// call and NaN, neither a hung GPU nor plausible wrong numbers.
typedef unsigned long long ull;
__device__ __forceinline__ void st_release_sys(ull *p, ull v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ ull ld_acquire_sys(const ull *p) { ull v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ ull global_ns() { ull t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
struct PeerBoxes { ull *box[LP_MAX_PEERS]; };

// first owned plane -> left neighbour's right halo, last owned plane -> right neighbour's left halo; the last block to
// finish raises the two flags
__global__ void __launch_bounds__(256) k_peer_put_halo(const double2 *__restrict__ first, const double2 *__restrict__ last, double2 *__restrict__ left_dst,
                                                       double2 *__restrict__ right_dst, long long n2, ull *mbox, ull *left_flag, ull *right_flag)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n2; i += (long long)gridDim.x * blockDim.x) {
    if (i < n2) left_dst[i] = first[i];
    else right_dst[i - n2] = last[i - n2];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const ull done = atomicAdd(mbox + LP_MB_PUTCNT, 1ULL);
    if (done == gridDim.x - 1) {
      mbox[LP_MB_PUTCNT] = 0;
      const ull e = mbox[LP_MB_EPOCH_H] + 1;
      mbox[LP_MB_EPOCH_H] = e;
      __threadfence_system();
      st_release_sys(left_flag, e);       // the left neighbour's "from the right" flag
      st_release_sys(right_flag, e);      // the right neighbour's "from the left" flag
    }
  }
}
// this rank's (m_i, s_i) into the parity buffer of every mailbox, then the flags
__global__ void __launch_bounds__(256) k_peer_publish_density(const double *__restrict__ ms_local, int n2, int off2, int Nx2, PeerBoxes pb, int world, int rank)
{
  ull *mine = pb.box[rank];
  const ull e = mine[LP_MB_EPOCH_D] + 1;
  for (int t = threadIdx.x; t < n2 * world; t += blockDim.x) {
    const int r = t / n2, i = t % n2;
    reinterpret_cast<double *>(pb.box[r] + LP_MB_MS)[(e & 1) * Nx2 + off2 + i] = ms_local[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) st_release_sys(pb.box[threadIdx.x] + LP_MB_DFLAG + rank, e);
  if (threadIdx.x == 0) mine[LP_MB_EPOCH_D] = e;
}
// wait for every rank's densities and both halo planes of this stage; gather the densities into ms_all.
// A wait that exceeds timeout_ns is counted in the mailbox, and the state is poisoned: ms_all is filled with NaN, so the
// field, this stage and everything after it turn into NaN instead of plausible numbers computed from stale halos; the
// host calls (lpgpu_step, lpgpu_synchronize, lpgpu_download_U ...) return an error from then on (api.cu, peer_check).
__global__ void __launch_bounds__(256) k_peer_wait(ull *mbox, int world, double *__restrict__ ms_all, int Nx2, ull timeout_ns)
{
  const ull ed = mbox[LP_MB_EPOCH_D], eh = mbox[LP_MB_EPOCH_H];
  if (threadIdx.x < world + 2) {
    const ull *flag = threadIdx.x < world ? mbox + LP_MB_DFLAG + threadIdx.x : mbox + LP_MB_HFLAG + (threadIdx.x - world);
    const ull want = threadIdx.x < world ? ed : eh;
    const ull t0 = global_ns();
    while (ld_acquire_sys(flag) < want) {
      __nanosleep(200);
      if (global_ns() - t0 > timeout_ns) { atomicAdd(mbox + LP_MB_ERR, 1ULL); break; }
    }
  }
  __syncthreads();
  const bool poisoned = *reinterpret_cast<volatile ull *>(mbox + LP_MB_ERR) != 0;     // sticky: this wait or an earlier one
  const double *src = reinterpret_cast<const double *>(mbox + LP_MB_MS) + (ed & 1) * Nx2;
  for (int i = threadIdx.x; i < Nx2; i += blockDim.x) ms_all[i] = poisoned ? __longlong_as_double(0x7ff8000000000000LL) : __ldcg(src + i);
}
// ---------------------------------------------------------------------------------------------------------------------
// The serial tail of an SSP-RK3 stage's field solve in ONE single-block kernel (it was four launches and a copy): fold the
// per-cell partial density sums in their fixed order; PEER: store them into every rank's mailbox and raise the flags, wait
// for every rank's densities and for both halo planes of this stage, gather; then the Nx-long scan.  The boundary planes
// travel meanwhile (k_peer_put_halo on the context's side stream, api.cu), so a stage's exchange costs one small kernel on
// the critical path instead of the chain put -> reduce -> fold -> publish -> wait -> scan.
template <bool PEER>
__global__ void __launch_bounds__(256) k_field_finish(const double *__restrict__ part, int nchunk, double scalev, double *ms_local, double *ms_all,
                                                      double *__restrict__ fld, int Nx, int x_begin, int ncell, double dx, double Lx, DopingParams D,
                                                      PeerBoxes pb, int world, int rank, ull timeout_ns)
{
  extern __shared__ double sms[];
  __shared__ double s_ce;
  for (int cell = threadIdx.x; cell < ncell; cell += blockDim.x) {
    double a, b;
    if (nchunk > 0) {
      a = 0.; b = 0.;
      for (int k = 0; k < nchunk; k++) { a += part[2 * (cell * nchunk + k)]; b += part[2 * (cell * nchunk + k) + 1]; }
      a *= scalev; b *= scalev;
      ms_local[2 * cell] = a; ms_local[2 * cell + 1] = b;
    } else {                                   // small velocity grids: k_field_reduce has written ms_local
      a = ms_local[2 * cell]; b = ms_local[2 * cell + 1];
    }
    if (!PEER) { ms_all[2 * (x_begin + cell)] = a; ms_all[2 * (x_begin + cell) + 1] = b; }
  }
  __syncthreads();
  if (PEER) {
    ull *mine = pb.box[rank];
    const ull e = mine[LP_MB_EPOCH_D] + 1, eh = mine[LP_MB_EPOCH_H];
    const int n2 = 2 * ncell, Nx2 = 2 * Nx;
    for (int t = threadIdx.x; t < n2 * world; t += blockDim.x) {
      const int r = t / n2, i = t % n2;
      reinterpret_cast<double *>(pb.box[r] + LP_MB_MS)[(e & 1) * Nx2 + 2 * x_begin + i] = ms_local[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) st_release_sys(pb.box[threadIdx.x] + LP_MB_DFLAG + rank, e);
    if (threadIdx.x == 0) mine[LP_MB_EPOCH_D] = e;
    if (threadIdx.x < world + 2) {
      const ull *flag = threadIdx.x < world ? mine + LP_MB_DFLAG + threadIdx.x : mine + LP_MB_HFLAG + (threadIdx.x - world);
      const ull want = threadIdx.x < world ? e : eh;
      const ull t0 = global_ns();
      while (ld_acquire_sys(flag) < want) {
        __nanosleep(100);
        if (global_ns() - t0 > timeout_ns) { atomicAdd(mine + LP_MB_ERR, 1ULL); break; }
      }
    }
    __syncthreads();
    const bool poisoned = *reinterpret_cast<volatile ull *>(mine + LP_MB_ERR) != 0;
    const double *src = reinterpret_cast<const double *>(mine + LP_MB_MS) + (e & 1) * Nx2;
    for (int i = threadIdx.x; i < Nx2; i += blockDim.x) ms_all[i] = poisoned ? __longlong_as_double(0x7ff8000000000000LL) : __ldcg(src + i);
  }
  __syncthreads();
  field_scan_body(ms_all, fld, Nx, x_begin, ncell, dx, Lx, D, sms, &s_ce);
}
// reduce + the fused tail; peer = false needs a context that owns all of x
int lp_launch_field_stage(lpgpu_ctx *c, const double *planes, bool peer)
{
  int nchunk = 0;
  if (c->sv < 8192) {
    k_field_reduce<<<c->ncell, 256, 0, c->stream>>>(planes, c->d_ms_local, c->sv, c->tab.scalev);
  } else {
    k_field_reduce_part<<<dim3(c->ncell, LP_FR_CH), 256, 0, c->stream>>>(planes, c->d_ms_part, c->sv);
    nchunk = LP_FR_CH;
  }
  LP_LAUNCHED(c);
  const double dx = c->p.Lx / c->p.Nx;
  DopingParams D = {c->p.doping, c->p.Nx / 3 - 1, 2 * c->p.Nx / 3 - 1, c->p.NL, c->p.NH, c->p.eps};
  const size_t smem = (size_t)3 * c->p.Nx * sizeof(double);
  PeerBoxes pb;
  for (int r = 0; r < LP_MAX_PEERS; r++) pb.box[r] = (peer && r < c->peer_world) ? c->peer_mbox[r] : nullptr;
  if (smem > 48 * 1024 && !c->finish_attr) {
    LP_CUDA(cudaFuncSetAttribute(k_field_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LP_CUDA(cudaFuncSetAttribute(k_field_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c->finish_attr = true;
  }
  if (peer)
    k_field_finish<true><<<1, 256, smem, c->stream>>>(c->d_ms_part, nchunk, c->tab.scalev, c->d_ms_local, c->d_ms_all, c->d_fld, c->p.Nx, c->p.x_begin, c->ncell,
                                                      dx, c->p.Lx, D, pb, c->peer_world, c->peer_rank, (ull)(c->peer_timeout_s * 1e9));
  else
    k_field_finish<false><<<1, 256, smem, c->stream>>>(c->d_ms_part, nchunk, c->tab.scalev, c->d_ms_local, c->d_ms_all, c->d_fld, c->p.Nx, c->p.x_begin, c->ncell,
                                                       dx, c->p.Lx, D, pb, 1, 0, 0ULL);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
int lp_launch_peer_put_halo(lpgpu_ctx *c, int stage)
{
  const long long plane = (long long)6 * c->sv;
  const double *in = c->d_U[stage];
  double *left_dst = c->peer_U[0][stage] + plane * (c->ncell + 1), *right_dst = c->peer_U[1][stage];
  const int left = (c->peer_rank + c->peer_world - 1) % c->peer_world, right = (c->peer_rank + 1) % c->peer_world;
  k_peer_put_halo<<<148, 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(in + plane), reinterpret_cast<const double2 *>(in + plane * c->ncell),
                                              reinterpret_cast<double2 *>(left_dst), reinterpret_cast<double2 *>(right_dst), plane / 2, c->d_mbox,
                                              c->peer_mbox[left] + LP_MB_HFLAG + 1, c->peer_mbox[right] + LP_MB_HFLAG + 0);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
int lp_launch_peer_publish_density(lpgpu_ctx *c)
{
  PeerBoxes pb;
  for (int r = 0; r < LP_MAX_PEERS; r++) pb.box[r] = r < c->peer_world ? c->peer_mbox[r] : nullptr;
  k_peer_publish_density<<<1, 256, 0, c->stream>>>(c->d_ms_local, 2 * c->ncell, 2 * c->p.x_begin, 2 * c->p.Nx, pb, c->peer_world, c->peer_rank);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
int lp_launch_peer_wait(lpgpu_ctx *c)
{
  k_peer_wait<<<1, 256, 0, c->stream>>>(c->d_mbox, c->peer_world, c->d_ms_all, 2 * c->p.Nx, (ull)(c->peer_timeout_s * 1e9));
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// computeMass / computeMomentum / computeKiE (MomentCalculations.cpp:23-131) partial sums over the
// local cells: one block per cell, then one block folds the per-cell partials in cell order.
__global__ void __launch_bounds__(256) k_moments_cell(const double *__restrict__ planes, double *__restrict__ part, int Nv, int sv,
                                                      double dv, double Lv)
{
  __shared__ double red[5 * 32];
  const long long cell = blockIdx.x;
  const double *u = planes + ((cell + 1) * 6) * (long long)sv;
  double v[5] = {0., 0., 0., 0., 0.};
  // two DG cells (j, j + 1: Nv is even, so they share j1 and j2) per thread and pass, double2 loads: the reduction is bound
  // by memory latency (five loads per cell), not by its arithmetic
  int per = (sv + gridDim.y - 1) / gridDim.y;
  per += per & 1;
  const int jlo = blockIdx.y * per, jhi = min(sv, jlo + per);
  for (int j = jlo + 2 * threadIdx.x; j < jhi; j += 2 * blockDim.x) {
    const int j3 = j % Nv, j2 = (j / Nv) % Nv, j1 = j / (Nv * Nv);
    const double c1 = -Lv + (j1 + 0.5) * dv, c2 = -Lv + (j2 + 0.5) * dv;
    const double2 A0 = *reinterpret_cast<const double2 *>(u + j), A2 = *reinterpret_cast<const double2 *>(u + 2LL * sv + j),
                  A3 = *reinterpret_cast<const double2 *>(u + 3LL * sv + j), A4 = *reinterpret_cast<const double2 *>(u + 4LL * sv + j),
                  A5 = *reinterpret_cast<const double2 *>(u + 5LL * sv + j);
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const double c3 = -Lv + (j3 + h + 0.5) * dv, r2 = c1 * c1 + c2 * c2 + c3 * c3;
      const double U0 = h ? A0.y : A0.x, U2 = h ? A2.y : A2.x, U3 = h ? A3.y : A3.x, U4 = h ? A4.y : A4.x, U5 = h ? A5.y : A5.x;
      v[0] += U0 + U5 / 4.;
      v[1] += c1 * dv * U0 + U2 * dv * dv / 12. + U5 * c1 * dv / 4.;
      v[2] += c2 * dv * U0 + U3 * dv * dv / 12. + U5 * c2 * dv / 4.;
      v[3] += c3 * dv * U0 + U4 * dv * dv / 12. + U5 * c3 * dv / 4.;
      v[4] += U0 * (r2 + dv * dv / 4.) * dv + (c1 * U2 + c2 * U3 + c3 * U4) * dv * dv / 6. + U5 * (dv * dv * dv * 19. / 240. + r2 * dv / 4.);
    }
  }
  block_sum<5>(v, red);
  if (threadIdx.x == 0)
    for (int m = 0; m < 5; m++) part[5 * (cell * gridDim.y + blockIdx.y) + m] = v[m];
}
// fold of n partial rows of NV values: thread t adds rows t, t + 256, ... in that order, block_sum adds the 256 threads in
// a fixed tree -- the same bits run to run, and not the 0.8 ms (moments, 512 cells) / 3.5 ms (diagnostics) that one
// thread walking all the rows took
template <int NV>
__device__ __forceinline__ void fold_rows(const double *__restrict__ part, int n, double (&v)[NV], double *red)
{
  #pragma unroll
  for (int m = 0; m < NV; m++) v[m] = 0.;
  for (int c = threadIdx.x; c < n; c += blockDim.x)
    #pragma unroll
    for (int m = 0; m < NV; m++) v[m] += part[NV * c + m];
  block_sum<NV>(v, red);
}
__global__ void __launch_bounds__(256) k_moments_fold(const double *__restrict__ part, double *__restrict__ out, int ncell, double xs, double dv, double scalev)
{
  __shared__ double red[5 * 32];
  double v[5];
  fold_rows<5>(part, ncell, v, red);
  if (threadIdx.x != 0) return;
  out[0] = v[0] * xs * scalev;
  out[1] = v[1] * xs * dv * dv; out[2] = v[2] * xs * dv * dv; out[3] = v[3] * xs * dv * dv;
  out[4] = 0.5 * v[4] * xs * dv * dv;
}
// ---------------------------------------------------------------------------------------------
// The other per-step diagnostics of the reference's rank 0 (LP_ompi.cpp:819,829,846):
// computeEntropy (EntropyCalculations.cpp:23-122: 5^4-point Gauss rule of f log f per DG cell, 5^3 when
// homogeneous), FindNegVals (NegativityChecks.cpp:24-160: sign of the cell average by the same rule) and
// computeKiEratio (MomentCalculations.cpp:133-199).  One thread per DG cell; per-cell-block partials
// are folded in order.  out4 = entropy (scaled), KiE terms over non-negative cells, over negative cells,
// number of negative cells.
__constant__ double c_gw[5] = {0.5688888888888889, 0.4786286704993665, 0.4786286704993665, 0.2369268850561891, 0.2369268850561891};
__constant__ double c_gt[5] = {0., -0.5384693101056831, 0.5384693101056831, -0.9061798459386640, 0.9061798459386640};
// log of a positive double for the entropy integrand: 128-interval table on the mantissa (c_i = 1 + (i + 1/2)/128, the
// table holds rc_i = fl(1/c_i) and -log(rc_i)), d = m rc_i - 1 by one FMA (|d| <= 1/256), log1p(d) to degree 6
// (truncation d^7/7 <= 2e-18).  |error| <= 2e-16 (1 + |log x|): the same bound as libm's to within a factor of two, at a fifth of
// its FP64 instructions -- the entropy (5^4 logarithms per DG cell per step) is bound by exactly those.
// x is a positive normal number or +inf where the result is used; anything else gives a finite or NaN value the caller
// discards.  No special cases, no branches.  The 5^4-point rule executes this 10^10 times per step at Nx = 512, so the
// instruction count matters as much as the FP64 count: the coefficients are constant-bank operands (as immediates the
// compiler rebuilt each of them in uniform registers for every call), and the table is read with ld.shared from a
// 32-bit address made once per thread (through a pointer the shared window base was recomputed per call).
__constant__ double c_lp[4] = {-1. / 6., 0.2, 1. / 3., 0.693147180559945309417};
__device__ __forceinline__ double log_pos(double x, unsigned tab_s)
{
  const int hi = __double2hiint(x);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
  double rc, lrc;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rc), "=d"(lrc) : "r"(tab_s + ((hi >> 9) & 0x7f0)));
  const double d = fma(m, rc, -1.0);
  double q = fma(d, c_lp[0], c_lp[1]);
  q = fma(d, q, -0.25); q = fma(d, q, c_lp[2]); q = fma(d, q, -0.5);
  const double p = fma(d * d, q, d);
  return fma((double)((hi >> 20) - 1023), c_lp[3], lrc + p);
}
__global__ void __launch_bounds__(128, 4) k_diag_cell(const double *__restrict__ planes, double *__restrict__ part, int Nv, int sv,
                                                   double dv, double Lv, int homogeneous)
{
  __shared__ double red[4 * 32];
  __shared__ double2 ltab[128];
  {
    const double rc = 1.0 / (1.0 + (threadIdx.x + 0.5) / 128.);   // 128 threads: one entry each
    ltab[threadIdx.x] = make_double2(rc, -log(rc));
  }
  __syncthreads();
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(ltab);
  const long long cell = blockIdx.x;
  const double *u = planes + ((cell + 1) * 6) * (long long)sv;
  double v[4] = {0., 0., 0., 0.};
  const int nxq = homogeneous ? 1 : 5;
  const int per = (sv + gridDim.y - 1) / gridDim.y, jlo = blockIdx.y * per, jhi = min(sv, jlo + per);
  for (int j = jlo + threadIdx.x; j < jhi; j += blockDim.x) {
    const double U0 = u[j], U1 = u[1LL * sv + j], U2 = u[2LL * sv + j], U3 = u[3LL * sv + j], U4 = u[4LL * sv + j], U5 = u[5LL * sv + j];
    double e = 0., avg = 0.;
    for (int a = 0; a < nxq; a++)
      for (int b = 0; b < 5; b++)
        #pragma unroll
        for (int cc = 0; cc < 5; cc++) {
          // the point values f are formed exactly as in the reference's one-line expression; the weights are applied
          // innermost-first (w_d inside, w_a w_b w_c once per line of five points)
          const double xs = homogeneous ? 0. : 0.5 * c_gt[a], x1 = 0.5 * c_gt[b], x2 = 0.5 * c_gt[cc];
          const double head = U0 + (homogeneous ? 0. : U1 * xs) + U2 * x1 + U3 * x2;
          const double wabc = (homogeneous ? 1. : c_gw[a]) * c_gw[b] * c_gw[cc], q12 = x1 * x1 + x2 * x2;
          double ein = 0., ain = 0.;
          #pragma unroll
          for (int d = 0; d < 5; d++) {
            const double x3 = 0.5 * c_gt[d];
            const double f = head + U4 * x3 + U5 * (q12 + x3 * x3);
            const double lf = log_pos(f, tab_s);
            // f > 0 (EntropyCalculations.cpp:71 adds f log f only there), as an integer test on the high word:
            // positive normal numbers and +inf pass; zero, negatives, NaN and denormals below 2^-1042 (whose f log f is
            // below 1e-310) do not
            const bool pos = (unsigned)(__double2hiint(f) - 1) < 0x7ff00000u;
            ein = pos ? fma(c_gw[d], f * lf, ein) : ein;
            ain = fma(c_gw[d], f, ain);
          }
          e = fma(wabc, ein, e);
          avg = fma(wabc, ain, avg);
        }
    const int j3 = j % Nv, j2 = (j / Nv) % Nv, j1 = j / (Nv * Nv);
    const double c1 = -Lv + (j1 + 0.5) * dv, c2 = -Lv + (j2 + 0.5) * dv, c3 = -Lv + (j3 + 0.5) * dv, r2 = c1 * c1 + c2 * c2 + c3 * c3;
    const double ke = U0 * (r2 + dv * dv / 4.) * dv + (c1 * U2 + c2 * U3 + c3 * U4) * dv * dv / 6. + U5 * (dv * dv * dv * 19. / 240. + r2 * dv / 4.);
    v[0] += e;
    if (avg < 0) { v[2] += ke; v[3] += 1.; } else v[1] += ke;
  }
  block_sum<4>(v, red);
  if (threadIdx.x == 0)
    for (int m = 0; m < 4; m++) part[4 * (cell * gridDim.y + blockIdx.y) + m] = v[m];
}
__global__ void __launch_bounds__(256) k_diag_fold(const double *__restrict__ part, double *__restrict__ out, int ncell, double scale)
{
  __shared__ double red[4 * 32];
  double v[4];
  fold_rows<4>(part, ncell, v, red);
  if (threadIdx.x != 0) return;
  out[0] = v[0] * scale; out[1] = v[1]; out[2] = v[2]; out[3] = v[3];
}
int lp_launch_diagnostics(lpgpu_ctx *c, const double *planes, double *out4_dev)
{
  double *part = c->d_B;   // free outside the projection
  // blocks per x cell: four 128-thread blocks fit an SM, so a grid that is a multiple of 4 x 148 runs in full waves
  // (32 cells x 37 chunks = 2 waves); the 5^4-point rule makes this kernel FP64-bound, a ragged last wave costs real time
  int chunks = 148 / std::gcd(c->ncell, 148);
  while (chunks > 1 && c->sv / chunks < 128) chunks = (chunks + 1) / 2;
  k_diag_cell<<<dim3(c->ncell, chunks), 128, 0, c->stream>>>(planes, part, c->p.Nv, c->sv, c->tab.dv, c->p.Lv, c->p.homogeneous);
  LP_LAUNCHED(c);
  const double dv = c->tab.dv, dx = c->p.Lx / c->p.Nx;
  const double scale = 0.5 * dv * 0.5 * dv * 0.5 * dv * (c->p.homogeneous ? 1. : 0.5 * dx);
  k_diag_fold<<<1, 256, 0, c->stream>>>(part, out4_dev, c->ncell * chunks, scale);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// PrintMarginal (MarginalCreation.cpp:16-67, 160-219) needs only four sums per output cell:
//   inhomogeneous: per (x cell, j1)  sum over (j2, j3) of U0, U1, U2, U5   (f_marg_Inhomo)
//   homogeneous:   per (j1, j2)      sum over j3       of U0, U2, U3, U5   (f_marg_Homo)
// One warp per output cell, fixed-order tree; the host evaluates the 4 x 4 sub-grid points from them.
__global__ void __launch_bounds__(256) k_marginal_sums(const double *__restrict__ planes, double *__restrict__ out, int Nv, int sv, int ncell,
                                                       int homogeneous)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nout = homogeneous ? Nv * Nv : ncell * Nv;
  if (warp >= nout) return;
  const int cmap_in[4] = {0, 1, 2, 5}, cmap_ho[4] = {0, 2, 3, 5};
  long long base; int count;
  if (homogeneous) { base = 6LL * sv + (long long)warp * Nv; count = Nv; }                       // (j1, j2) -> j3 run
  else { const int cell = warp / Nv, j1 = warp % Nv; base = ((long long)(cell + 1) * 6) * sv + (long long)j1 * Nv * Nv; count = Nv * Nv; }
  double v[4] = {0., 0., 0., 0.};
  for (int t = lane; t < count; t += 32)
    #pragma unroll
    for (int a = 0; a < 4; a++) v[a] += planes[base + (long long)(homogeneous ? cmap_ho[a] : cmap_in[a]) * sv + t];
  #pragma unroll
  for (int a = 0; a < 4; a++) {
    for (int o = 16; o > 0; o >>= 1) v[a] += __shfl_down_sync(0xffffffffu, v[a], o);
    if (lane == 0) out[4LL * warp + a] = v[a];
  }
}
int lp_launch_marginal_sums(lpgpu_ctx *c, const double *planes, double *out_dev)
{
  const int nout = c->p.homogeneous ? c->p.Nv * c->p.Nv : c->ncell * c->p.Nv;
  k_marginal_sums<<<(nout * 32 + 255) / 256, 256, 0, c->stream>>>(planes, out_dev, c->p.Nv, c->sv, c->ncell, c->p.homogeneous);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}

int lp_launch_moments(lpgpu_ctx *c, const double *planes)
{
  // d_B is free outside the projection: use its head for the per-cell partials
  double *part = c->d_B;
  const int chunks = c->sv >= 4096 ? (c->ncell >= 64 ? 8 : 32) : 1;   // blocks per x cell (one block per cell left most of the GPU idle)
  const bool prof4 = c->prof_on == 4 && c->prof_used + 2 <= c->prof_ev.size();   // bench.py: achieved HBM GB/s of the moment reduction
  if (prof4) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  k_moments_cell<<<dim3(c->ncell, chunks), 256, 0, c->stream>>>(planes, part, c->p.Nv, c->sv, c->tab.dv, c->p.Lv);
  LP_LAUNCHED(c);
  if (prof4) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  const double xs = c->p.homogeneous ? 1. : c->p.Lx / c->p.Nx;
  k_moments_fold<<<1, 256, 0, c->stream>>>(part, c->d_mom, c->ncell * chunks, xs, c->tab.dv, c->tab.scalev);
  LP_LAUNCHED(c);
  return LPGPU_OK;
}
