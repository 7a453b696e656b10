/* include/lpgpu.h -- C ABI of the B200 Landau-Poisson hot path (liblpgpu.so).
 *
 * The reference (ClarkPennie/landau-poisson-solver) has no plugin/FFI interface: the seam is the
 * set of free functions its main() calls inside the time loop (LP_ompi.cpp:662-813).  Each entry
 * point below names the reference call it replaces.  Plain pointers and sizes only.
 *
 * Conventions kept from the reference: the caller owns every host array; `U` is the single
 * source of truth between phases, host layout AoS U[6*k+l] with k = i*Nv^3 + j1*Nv^2 + j2*Nv + j3
 * (advection_1.cpp:75-80); spectral arrays are C-order N^3 with interleaved (re,im) complex
 * (fftw_complex); calls are synchronous at the boundary unless stated; one caller thread per
 * context.  Where the reference returns void and exits on error, these return 0 on success and
 * a non-zero LPGPU_E* code otherwise (lpgpu_last_error() gives the text).
 *
 * There is no CPU fallback: every compute entry point fails with LPGPU_ENODEV when no CUDA
 * device is usable.
 */
#ifndef LPGPU_H
#define LPGPU_H
#ifdef __cplusplus
extern "C" {
#endif

#define LPGPU_OK 0
#define LPGPU_EINVAL 1   /* bad argument / unsupported option (e.g. gamma outside {-3, 0, 1}) */
#define LPGPU_ENODEV 2   /* no usable CUDA device */
#define LPGPU_ECUDA 3    /* CUDA runtime error */
#define LPGPU_ENOMEM 4

typedef struct lpgpu_ctx lpgpu_ctx;

/* Mirrors the scalars the reference reads from LPsolver-input.txt (InputParsing.cpp:291-470)
 * plus the shard this context owns.  x_begin/x_count select the contiguous block of spatial
 * cells held by this GPU (the reference's chunk_Nx split, LP_ompi.cpp:208-220, :673);
 * single GPU: x_begin = 0, x_count = Nx.  Homogeneous: Nx is ignored, one cell. */
typedef struct lpgpu_params {
  int Nx, Nv, N;
  double Lv, Lx, nu, dt;
  int gamma;        /* collision kernel, as ReadGamma (InputParsing.cpp:202-238): -3 Landau/Coulomb, 0 Maxwell molecules,
                       1 hard spheres (generate_conv_weights(conv_weights, gamma), LP_ompi.cpp:424) */
  int homogeneous;  /* reference flag Homogeneous */
  int x_begin, x_count;
  int device;       /* CUDA device ordinal */
  int computeq_variant; /* which ComputeQ kernel evaluates the weighted spectral convolution (same result to round-off):
                           0 = fastest validated: seven zero-padded FFT convolutions when 3N/2 = 2^a 3^b (N = 8, 16, 24, 32
                               run the register-resident pipeline fused with fft3D/conserveMoments/FS),
                               otherwise the register-tiled direct sum;
                           1 = simple one-thread-per-xi direct kernel (on-device cross-check);
                           2 = FFT convolutions; 3 = register-tiled direct sum (the O(N^6) form of the reference) */
  int full_and_linear;  /* reference flag FullandLinear (test 3): electron-ion term next to Q(f,f) -- ComputeQ_FandL,
                           conserveAllMoments_FandL, RK4_FandL (collisionRoutines_1.cpp:605-689, 800-901, 987-1085;
                           conservationRoutines.cpp:102-129).  With N = 8, 16, 24, 32 and computeq_variant 0 / 2 the linear part runs
                           through the FFT-convolution pipeline too (fixed symbols convolved with monomials of e times fhat);
                           variants 1 / 3 and other N use the direct O(N^6) sum. */
  /* ---- reference test 1 options (appended: older callers that zero the struct keep their behaviour) ---- */
  int doping;           /* reference flag Doping: step doping profile ND(x) (NH outside, NL inside the middle third,
                           FieldCalculations.cpp:413-425), the *_Doping field integrals (:427-676) and Dirichlet walls in
                           I3 (advection_1.cpp:24-69, 214-282) instead of the periodic neighbour */
  double NL, NH, eps, T_L, T_R;   /* [Doping] section: densities, dielectric constant, wall temperatures */
  int linear_landau;    /* reference flag LinearLandau: Q(f, M) -- ComputeQLinear / RK4Linear (collisionRoutines_1.cpp:
                           1185-1350) against the transform of the state captured by lpgpu_set_maxwellian */
  int mass_cons_only;   /* reference flag MassConsOnly: conserveMass_Normal (conservationRoutines.cpp:290-349) */
} lpgpu_params;

const char *lpgpu_last_error(void);
int lpgpu_device_count(void);

/* Replaces the start-up calls createCCtAndPivot() and generate_conv_weights() (LP_ompi.cpp:375,
 * :424; conservationRoutines.cpp:159-216; collisionRoutines_1.cpp:220-237) and the allocation
 * of the collision/advection work arrays (LP_ompi.cpp:227-348). */
int lpgpu_init(const lpgpu_params *p, lpgpu_ctx **out);
int lpgpu_finalize(lpgpu_ctx *c);
/* Run all work of this context on the given cudaStream_t (passed as void*); NULL = default. */
int lpgpu_set_stream(lpgpu_ctx *c, void *cuda_stream);
int lpgpu_synchronize(lpgpu_ctx *c);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
long long lpgpu_launch_count(const lpgpu_ctx *c);

/* ---- state transfer: replaces MPI_Bcast(U) at LP_ompi.cpp:658 and :813 ------------------- */
/* U_host: this shard only, x_count*Nv^3*6 doubles, reference AoS layout. */
int lpgpu_upload_U(lpgpu_ctx *c, const double *U_host);
int lpgpu_download_U(lpgpu_ctx *c, double *U_host);
/* Enqueue-only forms for pipelined callers (several contexts, one stream each: the copies of one overlap the
 * kernels of another).  U_host must be page-locked and stay untouched until lpgpu_synchronize(c) returns. */
int lpgpu_upload_U_async(lpgpu_ctx *c, const double *U_host);
int lpgpu_download_U_async(lpgpu_ctx *c, double *U_host);

/* ComputeDFTofMaxwellian(U, f, DFTMaxwell) (LP_ompi.cpp:516, :541; collisionRoutines_1.cpp:1169-1183): capture the
 * shifted transform of the state currently held on the device, cell by cell, as the fixed second argument M of the
 * linear operator Q(f, M).  Needs linear_landau = 1; call once after uploading the initial condition. */
int lpgpu_set_maxwellian(lpgpu_ctx *c);

/* ---- whole phases ----------------------------------------------------------------------- */
/* RK3(U), LP_ompi.cpp:666 / advection_1.cpp:412-576.  A context that owns all of x, or a shard whose peers are mapped
 * (lpgpu_peer_import below); other sharded runs drive the three per-stage calls below around their own exchange. */
int lpgpu_advect_rk3(lpgpu_ctx *c);
/* setInit_spectral + for every local cell ComputeQ, conserveMoments, RK4 + the scatter of the 5
 * updated coefficients into U: LP_ompi.cpp:671-754. */
int lpgpu_collide_step(lpgpu_ctx *c);
/* Same work, enqueued only: the sharded driver queues the next stage's exchange behind it instead of stalling the
 * host once per phase.  lpgpu_synchronize (or any synchronous call) before reading results. */
int lpgpu_collide_step_async(lpgpu_ctx *c);
/* nsteps passes of the while(t<nT) body without diagnostics (all of x, or a shard whose peers are mapped).  From the
 * second execution on a timestep is a CUDA graph replay. */
int lpgpu_step(lpgpu_ctx *c, int nsteps);
/* Same, enqueued only (lpgpu_synchronize before reading results). */
int lpgpu_step_async(lpgpu_ctx *c, int nsteps);
/* One pass of the while(t<nT) body (LP_ompi.cpp:662-813) on a state that stays in HOST memory, as the reference's U does:
 * U_in (this context's shard, the reference's layout U[6k+l]) -> RK3 -> collision step -> U_out.  Equivalent to
 * lpgpu_upload_U(U_in); lpgpu_step(1); lpgpu_download_U(U_out), but pipelined over chunks of x cells: a chunk's layout
 * kernel runs while the next chunk is still being copied, and the collided chunks travel back to U_out (the other PCIe
 * direction) while the following chunks are collided.  Pass page-locked buffers for the copies to be asynchronous;
 * U_out may be U_in.  Synchronous: the result is in U_out on return. */
int lpgpu_step_host(lpgpu_ctx *c, const double *U_in, double *U_out);

/* ---- sharded advection: one exchange per SSP-RK3 stage (stage = 0,1,2) ------------------- */
/* Device addresses (in this context's memory) for the exchange of stage `stage`:
 *   ms_local  : 2*x_count doubles written by lpgpu_advect_reduce: (m_i, s_i) per local cell
 *   ms_all    : 2*Nx doubles the caller must fill (all-gather of every shard's ms_local)
 *   send_left / send_right : first / last owned x-plane of the stage input (6*Nv^3 doubles each)
 *   recv_left / recv_right : halo planes to fill from the left / right neighbour's
 *                            send_right / send_left (periodic in x, advection_1.cpp:297,306)  */
typedef struct lpgpu_exchange {
  double *ms_local, *ms_all;
  double *send_left, *send_right, *recv_left, *recv_right;
  long long plane_doubles;
} lpgpu_exchange;
int lpgpu_advect_exchange_info(lpgpu_ctx *c, int stage, lpgpu_exchange *out);
/* per-cell density sums of the stage input (asynchronous on the context's stream) */
int lpgpu_advect_reduce(lpgpu_ctx *c, int stage);
/* after the caller's all-gather + halo exchange: field integrals, DG right-hand side and the
 * SSP combination of this stage (asynchronous on the context's stream) */
int lpgpu_advect_apply(lpgpu_ctx *c, int stage);

/* Peer-memory exchange: with one process per GPU on one node, the ranks map each other's stage buffers and a small
 * mailbox (CUDA IPC); from then on every rank writes its boundary planes straight into its neighbours' halo planes and its
 * densities into every rank's mailbox, followed by a flag, from kernels on the context's stream.  This replaces the
 * reference's MPI_Bcast(U) / the all-gather + send/recv above: lpgpu_advect_rk3 and lpgpu_step* then work on a shard,
 * the host makes no call per stage, and the whole timestep replays as one CUDA graph.
 *   every rank:  lpgpu_peer_export(ctx, blob)  ->  all-gather the blobs (MPI_Allgather / torch.distributed)  ->
 *                lpgpu_peer_import(ctx, rank, world, blobs)     (rank r owns cells [r Nx/world, (r+1) Nx/world), world <= 8)
 * All ranks must then execute the same sequence of steps, with one peer-mapped context in flight per process (the
 * waits of two independent contexts could block each other across hardware queues); synchronise the ranks before
 * lpgpu_finalize.
 * Waits for a peer are bounded (default 60 s, or LPGPU_PEER_TIMEOUT_S; lpgpu_peer_set_timeout before the first timestep).
 * Fail-stop: a wait that times out poisons the state with NaN on the device, and every later lpgpu_step / _synchronize /
 * _advect_rk3 / _download_U of this context returns LPGPU_ECUDA; lpgpu_peer_status reports the count at any time. */
#define LPGPU_PEER_HANDLE_BYTES 256
int lpgpu_peer_export(lpgpu_ctx *c, void *blob);
int lpgpu_peer_import(lpgpu_ctx *c, int rank, int world, const void *blobs);
int lpgpu_peer_set_timeout(lpgpu_ctx *c, double seconds);
int lpgpu_peer_status(lpgpu_ctx *c, long long *timeouts);

/* ---- fine-grained entry points (host buffers, B cells per call) -------------------------- */
/* void setInit_spectral(double *U, double **f)            SetInit_1.h:48 ; f: x_count*N^3 */
int lpgpu_setInit_spectral(lpgpu_ctx *c, double *f_host);
/* void fft3D(fftw_complex *in, fftw_complex *out)         collisionRoutines_1.cpp:285 */
int lpgpu_fft3D(lpgpu_ctx *c, const double *in, double *out, int B);
/* void FS(fftw_complex *in, fftw_complex *out)            collisionRoutines_1.cpp:363 (real part in out[.][0], imag set to 0) */
int lpgpu_FS(lpgpu_ctx *c, const double *in, double *out, int B);
/* void ComputeQ(double *f, fftw_complex *qHat, double **conv_weights)   collisionRoutines_1.h:72 */
int lpgpu_ComputeQ(lpgpu_ctx *c, const double *f, double *qHat, int B);
/* void conserveMoments(fftw_complex *qHat, ...)           conservationRoutines.h:28 */
int lpgpu_conserveMoments(lpgpu_ctx *c, double *qHat, int B);
/* ComputeQ + conserveMoments with everything resident on the device: evaluates the B cells
 * currently sampled by lpgpu_sample_device() -- the unit of the "collision-cell evals/s" metric. */
int lpgpu_sample_device(lpgpu_ctx *c);
int lpgpu_eval_device(lpgpu_ctx *c, int B);
/* read back stage spectra of the last lpgpu_collide_step: which = 0..3 (qHat, Q1_fft..Q3_fft) */
int lpgpu_get_stage_spectrum(lpgpu_ctx *c, int which, double *out /* x_count*N^3*2 */);
/* field integrals of the current U (single shard): out = ce, cp[Nx], intE[Nx], intE1[Nx], intE2[Nx]
 * -- computePhi_x_0, computeC_rho, Int_E, Int_E1st, Int_E2nd (advection_1.cpp:419-429) */
int lpgpu_field(lpgpu_ctx *c, double *out);

/* ---- measurement helpers (bench.py) -------------------------------------------------------- */
/* Bracket ComputeQ with CUDA events on the context's stream: enable = 1 around the whole ComputeQ chain of
 * kernels, 2 around its dominant kernel only (the y/x-transform + product kernel of the FFT-convolution
 * pipeline), 3 around the DG stage kernels of the advection instead, 4 around the moment-reduction kernel, 0 off; read returns the summed device time and
 * the number of bracketed launches since enable (synchronises the stream).  While on, the library launches eagerly. */
int lpgpu_profile_computeQ(lpgpu_ctx *c, int enable);
int lpgpu_profile_read(lpgpu_ctx *c, double *total_ms, long long *launches);
/* DFMA micro-benchmark: sustained FP64 FMA rate of `device` in TFLOP/s (roofline denominator). */
int lpgpu_fp64_peak(int device, double *tflops);

/* ---- diagnostics ------------------------------------------------------------------------ */
/* Partial sums over this shard: out5 = mass, P1, P2, P3, KiE (computeMass/Momentum/KiE,
 * MomentCalculations.cpp:23-131); sum over shards for the global value.  ms_local_host (may be
 * NULL) receives the 2*x_count (m_i, s_i) pairs needed by lpgpu_eleE_from_ms. */
int lpgpu_moments_partial(lpgpu_ctx *c, double *out5, double *ms_local_host);
/* The remaining per-step diagnostics of the reference's rank 0 (LP_ompi.cpp:819,829,846), partial over
 * this shard: out4 = entropy (computeEntropy, EntropyCalculations.cpp:23-122), sum of the KiE terms over
 * cells with non-negative average, the same over cells with negative average (computeKiEratio =
 * out4[2]/out4[1] after summing over shards, MomentCalculations.cpp:133-199), number of cells FindNegVals
 * flags (NegativityChecks.cpp:24-160). */
int lpgpu_diagnostics_partial(lpgpu_ctx *c, double *out4);
/* The same numbers without stalling the time loop (the reference computes them on rank 0 between two timesteps,
 * LP_ompi.cpp:817-849; its 5^4-point entropy rule costs about as much FP64 work as a timestep).  _begin snapshots the
 * state on the context's stream and enqueues the moment, density and entropy/negativity reductions of the snapshot on a
 * side stream; the caller goes on to enqueue the next timestep; _end waits for the reductions and returns what
 * lpgpu_moments_partial (out5, ms_local_host: 2*x_count doubles, may be NULL) and lpgpu_diagnostics_partial (out4)
 * return for the snapshotted state.  One snapshot in flight per context: _end before the next _begin. */
int lpgpu_diagnostics_begin(lpgpu_ctx *c);
int lpgpu_diagnostics_end(lpgpu_ctx *c, double *out5, double *ms_local_host, double *out4);
/* PrintMarginal (LP_ompi.cpp:649, :870; MarginalCreation.cpp:16-67): the sums over the velocity directions that are
 * integrated out, per output cell: x_count*Nv rows of (sum U0, U1, U2, U5 over j2, j3) or, homogeneous, Nv*Nv rows of
 * (sum U0, U2, U3, U5 over j3).  The host evaluates the marginal at its 4 x 4 sub-grid points from them. */
int lpgpu_marginal_sums(lpgpu_ctx *c, double *out);
/* computeEleE (MomentCalculations.cpp:201-230) from the gathered per-cell sums; host-only. */
int lpgpu_eleE_from_ms(const lpgpu_params *p, const double *ms_all, double *EleE);

#ifdef __cplusplus
}
#endif
#endif
