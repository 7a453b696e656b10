// Internal declarations shared by the CUDA translation units of liblpgpu.so.
#pragma once
#include "../../include/lpgpu.h"
#include <cuda_runtime.h>
#include <string>
#include <vector>

#define LP_ETAB_PAD 8
#define LP_MAX_PEERS 8
// mailbox, in 8-byte words: density flags (epoch published by rank r), halo flags (from the left, from the right
// neighbour), timeout counter, own density epoch, own halo epoch, block counter of the halo put; then ms[2][2*Nx] doubles
#define LP_MB_DFLAG 0
#define LP_MB_HFLAG 8
#define LP_MB_ERR 10
#define LP_MB_EPOCH_D 11
#define LP_MB_EPOCH_H 12
#define LP_MB_PUTCNT 13
#define LP_MB_MS 16

// Host-side tables that depend only on (N, Nv, Lv): built once in lpgpu_init.
struct LpTables {
  int N, Nv;
  double Lv, dv, scalev, scaleL, scale3, L_eta, h_v, h_eta;
  std::vector<double> v, eta, wt;          // N
  std::vector<double> G;                   // 7*N^3  folded kernel symbols, [w*7 + t]
  std::vector<double> Gl;                  // 3*N^3  FullandLinear: h_eta^3 wt scale3 sum_i S_ij(w) w_i, [w*3 + j]
  double CCt_lin[4];                       // (C C^T)^-1 of the mass and energy rows
  std::vector<double> C5;                  // 5*N^3  conservation rows, [q*5 + m] (interleaved)
  double CCt[25];                          // (C C^T)^-1
  std::vector<double> Wfwd, Winv;          // N*N complex: DFT twiddle matrices exp(-/+ 2 pi i jk/N)
  std::vector<double> pre_fwd, pre_inv;    // (3N-2) complex: pre-phases by s = i+j+k
  std::vector<double> post_fwd, post_inv;  // N^3 complex: post-phases
  double c3_fwd;                           // scale3*h_v^3
  std::vector<double> T, M, S;             // N*Nv complex: 1-D IntModes factors [k*Nv + j]
  std::vector<int> node_cell;              // N
  std::vector<double> node_xi;             // N
  std::vector<double> vc;                  // Nv  cell centres Gridv(j)
  std::vector<double> Etab;                // N + 2*LP_ETAB_PAD: eta[z] - eta[N/2] for z = -PAD .. N+PAD-1
  std::vector<double> dirichlet;           // Doping: 2 x-planes (left, right wall) of DirichletBC coefficients, plane-major [wall][c][j]
};
void lp_build_tables(const lpgpu_params &p, LpTables &t);

struct lpgpu_ctx {
  lpgpu_params p;
  LpTables tab;
  int N3, sv, ncell;       // N^3, Nv^3, local cells (x_count or 1)
  int num_sms;             // multiprocessors of the device (grid of the persistent kernels)
  cudaStream_t stream;
  long long launches;
  // ---- device tables
  double *d_Etab, *d_qpart;
  size_t cap_part;         // capacity of d_qpart in spectra (cells x l-splits)
  double *d_eta, *d_G, *d_C5, *d_CCt, *d_Wfwd, *d_Winv, *d_pre_fwd, *d_pre_inv, *d_post_fwd, *d_post_inv, *d_wt, *d_T, *d_M, *d_S, *d_node_xi, *d_vc;
  int project_fold;                        // 1: T, M, S are conjugate-symmetric in k, the projection may fold k1 < 0 onto k1 > 0
  double *d_Tx, *d_Mx, *d_Sx;              // T, M, S with a row N = conj(row 0) appended (wave number +N/2; folded projection)
  int *d_node_cell;
  // ---- DG state, plane-major: buf[((p*6 + c)*sv + j)], p = 0..ncell+1 (planes 0 and ncell+1 are x halos)
  double *d_U[3];
  double *d_aos;           // staging for AoS <-> plane-major (ncell*sv*6)
  // ---- field
  double *d_ms_part;       // per-cell partial density sums (2 * 8 chunks * ncell)
  double *d_ms_local, *d_ms_all, *d_fld;   // 2*ncell, 2*Nx, 1 + 4*ncell (ce | cp,iE,iE1,iE2 interleaved by 4)
  double *d_mom;           // 5 partial moments
  // ---- collision work arrays (ncell cells each)
  double *d_f, *d_f1, *d_Qv;               // real N^3
  double *d_fhat, *d_tmp;                  // complex N^3
  double *d_q[4];                          // complex N^3: qHat, Q1_fft..Q3_fft
  double *d_lam;                           // 5 per cell
  double *d_Gl, *d_ql, *d_CCt_lin;          // FullandLinear: linear symbols, qHat_linear work array, 2x2 inverse
  double *d_GtLin, *d_ones, *d_fl_tmp, *d_fl_g;   // FullandLinear through the FFT-convolution pipeline: 4 symbol tables [4][7][y][z][x], a spectrum of ones, two work spectra
  double *d_dirichlet;                     // Doping: the two wall planes (2 * 6 * sv)
  double *d_mhat;                          // LinearLandau: DFTMaxwell, ncell * N^3 complex
  bool have_mhat;
  double *d_cpart;                         // [cell][N][5] partial conservation dots written by the fused ComputeQ
  double *d_B;                             // projection intermediate: ncell*N*4*Nv^2 complex
  size_t cap_cells;        // capacity (in cells) of the collision work arrays
  // ---- FFT-convolution variant of ComputeQ (allocated on first use)
  double *d_fc1, *d_fc2, *d_fctw, *d_Gt;
  bool fc3_attr;
  int fc_chunk;
  // ---- optional CUDA-event timing of the ComputeQ launches (bench.py roofline)
  // ---- CUDA graphs, captured once on gstream and replayed on `stream`: [0] one whole timestep (lpgpu_step*),
  //      [1] the collision step alone (lpgpu_collide_step*, what the sharded driver calls between its exchanges).
  //      The first execution of either runs eagerly (lazy allocations, function attributes), the second is captured.
  cudaStream_t gstream;
  cudaGraphExec_t gexec[2];
  bool graph_failed[2];
  int eager_runs[2];
  long long graph_launches[2];   // kernels per replay
  // ---- concurrent collision chains: the local cells in contiguous groups, each group a view of this context (same
  //      tables, per-cell arrays offset to its first cell) with its own stream; group 0 runs on `stream` itself
  std::vector<lpgpu_ctx *> groups;
  std::vector<cudaStream_t> group_streams;
  std::vector<cudaEvent_t> group_done;
  cudaEvent_t group_fork;
  // lpgpu_step_host: chunks of cells (views), the two copy streams, per-chunk "upload landed" / "ready to download" events
  std::vector<lpgpu_ctx *> hchunks;
  std::vector<int> hchunk_begin;
  cudaStream_t h2d_stream, d2h_stream;
  cudaEvent_t h_fork, h_join, h_trace0;
  std::vector<cudaEvent_t> h_up, h_down;
  bool is_view;
  // ---- peer-memory exchange of the sharded advection (one process per GPU on one node, CUDA IPC): every rank writes
  //      its densities into all ranks' mailboxes and its boundary planes into its neighbours' halo planes, then raises a
  //      flag; replaces the NCCL all-gather + send/recv, so the whole sharded timestep is one stream of kernels (one graph)
  bool peer_ready;
  double peer_timeout_s;                  // bound of a flag wait (k_peer_wait); a timeout poisons the state and fails the host calls
  bool scan_attr, finish_attr;            // k_field_scan / k_field_finish opted in to more than 48 KB of shared memory (Nx > 2048)
  cudaStream_t halo_stream;               // the boundary planes of a stage travel here while the densities are reduced and exchanged
  cudaEvent_t halo_fork, halo_join;
  int peer_rank, peer_world;
  unsigned long long *d_mbox;            // own mailbox (layout: LP_MB_* below)
  unsigned long long *peer_mbox[LP_MAX_PEERS];   // every rank's mailbox as mapped here ([peer_rank] = d_mbox)
  double *peer_U[2][3];                  // left / right neighbour's three stage buffers as mapped here
  std::vector<void *> peer_opened;       // cudaIpcOpenMemHandle results to close
  // ---- diagnostics of a snapshot on a side stream (lpgpu_diagnostics_begin/_end)
  lpgpu_ctx *diag_view;            // stream = diag_stream, scratch and result arrays of its own
  cudaStream_t diag_stream;
  cudaEvent_t diag_snap, diag_done;
  double *d_snap, *d_diag_scratch, *h_diag;   // state copy (with halo planes); partials; pinned results (5 + 4 + 2*ncell)
  bool diag_pending;
  int prof_on;            // 0 off, 1 events around the whole ComputeQ chain, 2 around its dominant kernel (F2) only
  std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
  size_t prof_used;        // events used so far
};

// error plumbing
void lp_set_error(const std::string &s);
#define LP_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      lp_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
      return LPGPU_ECUDA;                                                                   \
    }                                                                                       \
  } while (0)

#ifdef __CUDACC__
// ---- shifted-transform plumbing shared by collision.cu and fftconv.cu
struct FsEpilogue {
  double scaleL, scale3, dt, nu;
  double inv;        // 1/(scaleL*scale3): the register-line kernels multiply instead of dividing twice
  const double *f;   // stage-0 samples
  double *Qv;        // first-stage Q (written in mode 1, read in 2,3)
  double *f1;        // stage input for the next ComputeQ
};
struct PhaseTabs {
  const double2 *pre;    // [3N-2]  pre-phase by i+j+k            (null: none)
  const double2 *post;   // [N^3]   post-phase by (i,j,k)         (null: none)
  const double *wt;      // [N]     trapezoid weights, forward pre-factor only (null: none)
  double c3;             // scale3*h_v^3 (forward pre-factor)
};
// (c + i s) * (x + i y) exactly as the reference writes it: cos*re - sin*im, cos*im + sin*re
__device__ __forceinline__ double2 phase_mul(double2 cs, double2 x)
{
  return make_double2(__dsub_rn(__dmul_rn(cs.x, x.x), __dmul_rn(cs.y, x.y)), __dadd_rn(__dmul_rn(cs.x, x.y), __dmul_rn(cs.y, x.x)));
}
#endif

// ---- kernel launchers (collision.cu) -- all asynchronous on c->stream
int lp_launch_aos_to_planes(lpgpu_ctx *c, const double *aos, double *planes);
int lp_launch_planes_to_aos(lpgpu_ctx *c, const double *planes, double *aos);
int lp_launch_sample(lpgpu_ctx *c, const double *planes, double *f, int ncell);
int lp_launch_fft3d(lpgpu_ctx *c, const double *in, bool in_real, double *out, int B);
int lp_launch_fft3d_jk(lpgpu_ctx *c, const double *in, bool in_real, int B);
// true when ComputeQ runs as the fused FFT-convolution pipeline that takes the output of lp_launch_fft3d_jk
bool lp_fc3_available(const lpgpu_ctx *c);
// part (nullable, fc3 only): receives the [cell][N][5] partial conservation dot products of the unconserved spectrum
int lp_launch_computeQ_fftconv(lpgpu_ctx *c, const double *fhat, double *q, int B, bool fused_i, double *part);
// the same pipeline with other symbols Gt[7][y][z][x] and another first factor (first_stride = 0: one spectrum for all cells)
int lp_launch_fftconv_with(lpgpu_ctx *c, const double *fhat, double *q, int B, const double *Gt, const double *first, long long first_stride);
// allocates the pipeline's work arrays on first use (idempotent); -1 when the size has no FFT-convolution form
int lp_fc_prepare(lpgpu_ctx *c);
// conservation correction from those partials (in place)
int lp_launch_conserve_from_parts(lpgpu_ctx *c, double *q, const double *part, int B);
// FS whose first pass also applies the conservation correction from `part` to q (in place) before transforming
int lp_launch_fs_conserving(lpgpu_ctx *c, double *q, const double *part, int mode, int B, bool next_fwd);
// FS + RK-stage epilogue.  mode 0: plain (writes complex out, imag 0); 1..3: stage updates of f1
int lp_launch_fs(lpgpu_ctx *c, const double *q, int mode, double *out_complex, int B);
int lp_launch_computeQ(lpgpu_ctx *c, const double *fhat, double *q, int B);
int lp_launch_conserve(lpgpu_ctx *c, double *q, int B);
// FullandLinear: ComputeQ_FandL and conserveAllMoments_FandL followed by qHat += qHat_linear (RK4_FandL's first loop)
int lp_launch_computeQ_fandl(lpgpu_ctx *c, const double *fhat, double *q, double *ql, int B);
int lp_launch_conserve_fandl(lpgpu_ctx *c, double *q, double *ql, int B);
int lp_launch_project(lpgpu_ctx *c, double *planes, int B);
// ---- advection.cu
int lp_launch_field_reduce(lpgpu_ctx *c, const double *planes);
int lp_launch_field_scan(lpgpu_ctx *c);
// density reduction + one single-block kernel for the rest of a stage's field solve (fold, [peer: publish, wait, gather], scan)
int lp_launch_field_stage(lpgpu_ctx *c, const double *planes, bool peer);
int lp_launch_dg_stage(lpgpu_ctx *c, int stage);
int lp_launch_local_halo(lpgpu_ctx *c, double *planes);
// peer-memory exchange of one stage: boundary planes into the neighbours' halos; densities into every mailbox;
// wait for all of them (and copy the gathered densities to d_ms_all)
int lp_launch_peer_put_halo(lpgpu_ctx *c, int stage);
int lp_launch_peer_publish_density(lpgpu_ctx *c);
int lp_launch_peer_wait(lpgpu_ctx *c);
// Doping: Dirichlet wall planes into the halo planes that face a domain wall (no-op otherwise)
int lp_launch_wall_halo(lpgpu_ctx *c, double *planes);
int lp_launch_moments(lpgpu_ctx *c, const double *planes);
int lp_launch_marginal_sums(lpgpu_ctx *c, const double *planes, double *out_dev);
int lp_launch_diagnostics(lpgpu_ctx *c, const double *planes, double *out4_dev);
