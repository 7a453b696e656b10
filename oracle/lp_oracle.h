/* oracle/lp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's hot path: the per-cell conservative spectral
 * Landau collision step and the DG/SSP-RK3 advection step of ClarkPennie/landau-poisson-solver.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (landau-poisson-solver_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement element-wise against the
 * unmodified reference compiled from /root/reference (oracle/_ref/libref.so) and against the
 * reference's golden files tests/Moments_Test0.dc / _Test1 / _Test3 / _Test4 (copied values in
 * tests/golden/).
 *
 * Layouts are the reference's: U is AoS, U[6*k+l], k = i*Nv^3 + j1*Nv^2 + j2*Nv + j3
 * (advection_1.cpp:75-80); spectral arrays are C-order N^3, complex = interleaved (re,im).
 */
#ifndef LP_ORACLE_H
#define LP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpo_ctx lpo_ctx;

lpo_ctx *lpo_create(int Nx, int Nv, int N, double Lv, double Lx, double nu, double dt,
                    int homogeneous, int gamma);
void lpo_destroy(lpo_ctx *c);
/* 0: Fourier->DG projection through separable 1-D tables (default); 1: literal IntModes loop */
void lpo_set_direct_intmodes(lpo_ctx *c, int on);

/* tables */
void lpo_get_grids(const lpo_ctx *c, double *v, double *eta, double *wt);
double lpo_gHat3(const lpo_ctx *c, double z1, double z2, double z3, double k1, double k2, double k3);
void lpo_weight_row(const lpo_ctx *c, int xi, double *row);
void lpo_get_conservation(const lpo_ctx *c, double *C5, double *CCt25);
void lpo_IntModes(const lpo_ctx *c, int k1, int k2, int k3, int j1, int j2, int j3, double *out10);

/* collision path */
void lpo_setInit_spectral(const lpo_ctx *c, const double *U, double *f /* ncell*N^3 */);
void lpo_fft3D(const lpo_ctx *c, const double *in, double *out);
void lpo_FS(const lpo_ctx *c, const double *in, double *out);
void lpo_ComputeQ(const lpo_ctx *c, const double *f, double *qHat);
void lpo_conserveMoments(const lpo_ctx *c, double *qHat);
/* one cell: qHat must hold conserveMoments(ComputeQ(f)); writes dU[5*Nv^3] for that cell and,
 * when q123 != NULL, the three later stage spectra (3 * N^3 complex). */
void lpo_RK4(const lpo_ctx *c, const double *f, int cell, const double *qHat, const double *U,
             double *dU, double *q123);
void lpo_collide_step(const lpo_ctx *c, double *U);
/* FullandLinear variant (reference test 3): switches lpo_collide_step / lpo_step to ComputeQ_FandL,
 * conserveAllMoments_FandL and RK4_FandL (collisionRoutines_1.cpp:605-689, 800-901, 987-1085;
 * conservationRoutines.cpp:102-129) */
void lpo_set_fandl(lpo_ctx *c, int on);
void lpo_ComputeQ_FandL(const lpo_ctx *c, const double *f, double *qHat, double *qLin);
void lpo_conserveMoments_FandL(const lpo_ctx *c, double *qHat, double *qLin);

/* reference test 1 flags.  Doping = True: SetInit_ND, the *_Doping field integrals, I3_Doping with DirichletBC
 * (SetInit_1.cpp:125-173, FieldCalculations.cpp:413-676, advection_1.cpp:24-69, 214-282).  LinearLandau = True:
 * ComputeQLinear / RK4Linear with the transform of the state passed here as the fixed Maxwellian
 * (collisionRoutines_1.cpp:1169-1350); U = NULL switches back.  MassConsOnly = True: conserveMass_Normal
 * (conservationRoutines.cpp:290-349). */
void lpo_set_doping(lpo_ctx *c, double NL, double NH, double eps, double T_L, double T_R);
void lpo_SetInit_ND(const lpo_ctx *c, double *U);
void lpo_set_linear_landau(lpo_ctx *c, const double *U);
void lpo_set_mass_cons_only(lpo_ctx *c, int on);
void lpo_ComputeQLinear(const lpo_ctx *c, const double *f, const double *mhat, double *qHat);

/* advection path */
void lpo_field(const lpo_ctx *c, const double *U, double *out /* 1 + 4*Nx */);
void lpo_RK3(const lpo_ctx *c, double *U);

/* initial conditions and diagnostics */
void lpo_SetInit_LD(const lpo_ctx *c, double *U, double A_amp, double k_wave, int twostream);
void lpo_SetInit_4H(const lpo_ctx *c, double *U);
void lpo_SetInit_4H_Homo(const lpo_ctx *c, double *U);
void lpo_moments(const lpo_ctx *c, const double *U, double *out6);
void lpo_diagnostics(const lpo_ctx *c, const double *U, double *out4);

/* whole time step (advection then collision), LP_ompi.cpp:662-813 */
void lpo_step(const lpo_ctx *c, double *U);
int lpo_num_threads(void);
/* launchers such as torchrun export OMP_NUM_THREADS=1: bench.py sets the thread count explicitly */
void lpo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
