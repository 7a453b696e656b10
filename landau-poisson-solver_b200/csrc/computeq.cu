// Register-tiled ComputeQ for sm_100a -- the dominant kernel of the collision step.
//
//   Qhat[xi] = sum_{omega in win(xi)} Wt(xi,omega) fhat[omega] fhat[xi + N/2 - omega]
//   (collisionRoutines_1.cpp:691-774; Wt = h_eta^3 wt_l wt_m wt_n gHat3(xi,omega), :98-161)
//
// FP64-pipe bound by construction (27 N^6/64 pairs, no dense-contraction structure because Wt
// depends on both xi and omega), so the design minimises DFMA-pipe instructions per pair and keeps
// every lane busy on valid pairs only:
//
//  * the N^6 weight table of the reference (8.6 GB at N=32) is never materialised: with
//    e = eta[xi] - eta[omega],  Wt = c0 + c1 e3 + c2 e3^2  where c0,c1,c2 depend on (omega,e1,e2);
//    per l-slab the CTA folds e1 into 6 coefficients per (m,n) in shared memory, a thread folds e2
//    once per n-step (3 DFMA) and evaluates Wt per pair with 2 DFMA; the complex product costs 6.
//  * one CTA owns (cell, i = xi_1) and walks l = omega_1 over its window; per l it stages the
//    fhat slabs a = fhat[l,:,:], b = fhat[i+N/2-l,:,:] (zero padded in z) and the coefficients.
//  * a thread owns the outputs k in two R-wide tiles {k0, k0+N/2} of two rows {j, j+N/2}: the
//    window sizes of paired tiles / paired rows add up to a constant, so every thread executes
//    exactly the same number of (m, n) steps although the per-xi windows differ (no divergence, no
//    masked lanes; waste is only the R-1 padded pairs at the tile edges, ~6 % at N=32).  The b
//    values slide through registers (one 16-byte shared load per R pairs).
//  * the (row, m) items of a row pair are split over CS "slices" (lanes and warps); slices are
//    summed in a fixed order through shared memory at the end, so results are deterministic.
//
// Summation order differs from the reference's lexicographic omega loop -> ~1e-15 relative
// differences in Qhat (tests allow 1e-12).
#include "lpgpu_internal.h"

#define LP_LAUNCHED(c)                                  \
  do {                                                  \
    (c)->launches++;                                    \
    LP_CUDA(cudaGetLastError());                        \
  } while (0)

template <int N, int R, int JL, int CSL, int CSW>
struct CqCfg {
  static constexpr int H = N / 2;
  static constexpr int KG = N / (2 * R);          // thread groups along k (each owns tiles g*R and g*R + H)
  static constexpr int JG = H / JL;               // warp groups along j
  static constexpr int CS = CSL * CSW;            // item slices
  static constexpr int NITEM = N + H;             // (row, m) items of a row pair: |win(j)| + |win(j+H)|
  static constexpr int NI = (NITEM + CS - 1) / CS;
  static constexpr int ZP = (N + 2 * (R - 1)) | 1; // padded, odd row length of the b slab (bank-conflict free)
  static constexpr int NT = 32 * KG * JG * CSW;   // threads per CTA
  static constexpr int NE = N + 2 * LP_ETAB_PAD;   // eta-difference table, also staged in shared memory
  // a-slab and coefficient rows are padded, and rows m >= H shifted by 32 bytes, so that the (up to four) distinct rows a warp
  // touches in one load -- m, m+1 (two slices) and m-H, m+1-H (row B lanes) -- fall into different banks: one wavefront.
  static constexpr int SA = N + 1;                 // a-slab row stride (complex)
  static constexpr int SC = 6 * N + 2;             // coefficient row stride (doubles)
  static constexpr int A_ELEMS = N * SA + 2, C_ELEMS = N * SC + 4;
  static constexpr size_t SLAB_BYTES = ((size_t)A_ELEMS + (size_t)N * ZP) * 16 + (size_t)C_ELEMS * 8;
  static constexpr size_t RED_BYTES = (size_t)CS * N * N * 16;
  static constexpr size_t SMEM = (SLAB_BYTES > RED_BYTES ? SLAB_BYTES : RED_BYTES) + (size_t)NE * 8;
  static_assert(JL * CSL == 32, "a warp is JL row pairs x CSL slices");
  static_assert(H % JL == 0 && N % (2 * R) == 0, "tiling must divide N");
};


// One n-step of a tile: R pairs  acc[r] += Wt(e3_r) * a * b_r  against the z-window held in rotated
// registers.  U is the rotation phase: logical window slot r lives in physical slot (r - U) mod R, so
// sliding the window by one z costs no register moves -- the slot that drops out receives the next
// element.  Wt(e3) = c0 + e3 (c1 + e3 c2) with c0 = r0 + e2 (r1 + e2 r2), c1 = p1 + p2 e2.
template <int R, int U>
__device__ __forceinline__ void cq_step(double2 (&acc)[R], double2 (&bw)[R], double (&ew)[R], const double2 *ap,
                                        const double *cp, const double2 *bnew, const double *enew, double e2)
{
  const double2 a = *ap;
  const double2 c01 = reinterpret_cast<const double2 *>(cp)[0], c23 = reinterpret_cast<const double2 *>(cp)[1],
                c45 = reinterpret_cast<const double2 *>(cp)[2];
  const double c0 = fma(e2, fma(e2, c23.x, c01.y), c01.x);
  const double c1 = fma(c45.x, e2, c23.y);
  const double c2 = c45.y;
  #pragma unroll
  for (int r = 0; r < R; r++) {
    const int p = (r - U + R) % R;
    const double W = fma(ew[p], fma(ew[p], c2, c1), c0);
    const double war = W * a.x, wai = W * a.y;
    acc[r].x = fma(war, bw[p].x, acc[r].x);
    acc[r].x = fma(-wai, bw[p].y, acc[r].x);
    acc[r].y = fma(war, bw[p].y, acc[r].y);
    acc[r].y = fma(wai, bw[p].x, acc[r].y);
  }
  constexpr int pn = (R - 1 - U + R) % R;
  bw[pn] = *bnew; ew[pn] = *enew;
}
// steps U .. R-1 of a chunk; FULL = false stops after `cnt` steps (tail of the n loop)
template <int R, int U, bool FULL>
struct CqChunk {
  static __device__ __forceinline__ void run(double2 (&acc)[R], double2 (&bw)[R], double (&ew)[R], const double2 *ap,
                                             const double *cp, const double2 *bp, const double *ep, double e2, int cnt)
  {
    if (FULL || U < cnt) {
      cq_step<R, U>(acc, bw, ew, ap + U, cp + 6 * U, bp - U, ep - U, e2);
      CqChunk<R, U + 1, FULL>::run(acc, bw, ew, ap, cp, bp, ep, e2, cnt);
    }
  }
};
template <int R, bool FULL>
struct CqChunk<R, R, FULL> {
  static __device__ __forceinline__ void run(double2 (&)[R], double2 (&)[R], double (&)[R], const double2 *, const double *,
                                             const double2 *, const double *, double, int) {}
};

template <int N, int R, int JL, int CSL, int CSW, int MINB>
__global__ void __launch_bounds__((CqCfg<N, R, JL, CSL, CSW>::NT), MINB)
k_computeQ_tiled(const double2 *__restrict__ fhat, double2 *__restrict__ out, const double *__restrict__ G,
                 const double *__restrict__ eta, const double *__restrict__ Etab, int SL)
{
  using C = CqCfg<N, R, JL, CSL, CSW>;
  constexpr int H = C::H, KG = C::KG, JG = C::JG, CS = C::CS, NITEM = C::NITEM, NI = C::NI, ZP = C::ZP, NT = C::NT;
  extern __shared__ double2 sm2[];
  double2 *As = sm2;                                        // [N][N]      a = fhat[l, m, n]
  double2 *Bs = As + C::A_ELEMS;                            // [N][ZP]     b = fhat[x, y, z] at zz = z + R - 1
  double *Cf = reinterpret_cast<double *>(Bs + N * ZP);     // [N][N][6]   r0 r1 r2 p1 p2 c2
  double *Es = reinterpret_cast<double *>(reinterpret_cast<char *>(sm2) + (C::SMEM - (size_t)C::NE * 8));   // eta[z] - eta[N/2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jj = lane % JL, csl = lane / JL;
  const int g = warp % KG, jg = (warp / KG) % JG, csw = warp / (KG * JG);
  const int cs = csl + CSL * csw;
  const int j = jg * JL + jj;                               // row A = j, row B = j + H
  const int wA = j + H + 1;                                 // items c < wA: row A, m = c; else row B, m = c - H

  // heavy CTAs (large l-window) first
  const int nB = gridDim.x / N;                             // cells in this launch
  const int b = blockIdx.x / nB;                            // rank of i by window size: all cells of the heaviest i first
  const int i = (b & 1) ? H + (b >> 1) : H - 1 - (b >> 1);
  const int sl = blockIdx.y;
  const long long cell = blockIdx.x % nB;
  const double2 *fh = fhat + cell * (long long)(N * N * N);

  int ls, le;
  if (i < H) { ls = 0; le = i + H + 1; } else { ls = i - H + 1; le = N; }
  const int lo = ls + ((le - ls) * sl) / SL, hi = ls + ((le - ls) * (sl + 1)) / SL;

  const int k0a = g * R, k0b = g * R + H;
  // d = z of output r = 0; n = k0 + H - d
  const int dmaxA = k0a + H, dminA = (k0a + H - N + 1 > -(R - 1)) ? k0a + H - N + 1 : -(R - 1);
  const int dmaxB = N - 1, dminB = k0b + H - N + 1;        // k0b >= H: dmax clips at N-1, dmin = k0b-H+1 >= 1

  double2 live[2][R], saved[2][R];
  #pragma unroll
  for (int t = 0; t < 2; t++)
    #pragma unroll
    for (int r = 0; r < R; r++) { live[t][r] = make_double2(0., 0.); saved[t][r] = make_double2(0., 0.); }
  bool liveB = false;
  for (int t = tid; t < C::NE; t += NT) Es[t] = Etab[t];

  for (int l = lo; l < hi; l++) {
    const int x = i + H - l;
    const double e1 = eta[i] - eta[l];
    __syncthreads();
    for (int t = tid; t < N * N; t += NT) {
      const int m = t / N, n = t % N, up = (m >= H);
      As[m * C::SA + 2 * up + n] = fh[l * N * N + t];
      const double *gg = G + 7LL * (l * N * N + t);
      double *cf = Cf + m * C::SC + 4 * up + 6 * n;
      cf[0] = gg[0] - gg[1] * e1 * e1; cf[1] = -gg[4] * e1; cf[2] = -gg[2];
      cf[3] = -gg[5] * e1; cf[4] = -gg[6]; cf[5] = -gg[3];
    }
    for (int t = tid; t < N * ZP; t += NT) {
      const int y = t / ZP, z = t % ZP - (R - 1);
      Bs[t] = (z >= 0 && z < N) ? fh[(x * N + y) * N + z] : make_double2(0., 0.);
    }
    __syncthreads();

    const bool fwd = ((l - lo) & 1) == 0;                   // alternate item order so only one row switch per slab
    for (int it = 0; it < NI; it++) {
      const int c = cs + CS * (fwd ? it : NI - 1 - it);
      if (NITEM % CS != 0 && c >= NITEM) continue;
      const bool rowB = c >= wA;
      if (rowB != liveB) {
        #pragma unroll
        for (int t = 0; t < 2; t++)
          #pragma unroll
          for (int r = 0; r < R; r++) { const double2 tmp = live[t][r]; live[t][r] = saved[t][r]; saved[t][r] = tmp; }
        liveB = rowB;
      }
      const int row = rowB ? j + H : j, m = rowB ? c - H : c;
      const int y = row + H - m;
      const double e2 = eta[row] - eta[m];
      const double2 *brow = Bs + y * ZP + (R - 1);
      const double2 *arow = As + m * C::SA + (m >= H ? 2 : 0);
      const double *crow = Cf + m * C::SC + (m >= H ? 4 : 0);
      const double *et = Es + LP_ETAB_PAD;                   // et[z] = eta[z] - eta[N/2], z in [-PAD, N+PAD)

      #pragma unroll
      for (int t = 0; t < 2; t++) {
        const int k0 = t ? k0b : k0a, dmax = t ? dmaxB : dmaxA, dmin = t ? dminB : dminA;
        double2 bw[R]; double ew[R];
        #pragma unroll
        for (int r = 0; r < R; r++) { bw[r] = brow[dmax + r]; ew[r] = et[dmax + r]; }
        // d runs dmax .. dmin (n = k0 + H - d ascending); pointers to a[n], coefficients, next b and next e3
        const double2 *ap = arow + (k0 + H - dmax);
        const double *cp = crow + 6 * (k0 + H - dmax);
        const double2 *bp = brow + (dmax - 1);
        const double *ep = et + (dmax - 1);
        int cnt = dmax - dmin + 1;
        for (; cnt >= R; cnt -= R) {
          CqChunk<R, 0, true>::run(live[t], bw, ew, ap, cp, bp, ep, e2, R);
          ap += R; cp += 6 * R; bp -= R; ep -= R;
        }
        CqChunk<R, 0, false>::run(live[t], bw, ew, ap, cp, bp, ep, e2, cnt);
      }
    }
  }

  // fold the slices in a fixed order and write the CTA's N x N outputs
  __syncthreads();
  double2 *red = sm2;                                       // [CS][N][N]
  #pragma unroll
  for (int t = 0; t < 2; t++)
    #pragma unroll
    for (int r = 0; r < R; r++) {
      const int k = (t ? k0b : k0a) + r;
      const double2 vA = liveB ? saved[t][r] : live[t][r], vB = liveB ? live[t][r] : saved[t][r];
      red[(cs * N + j) * N + k] = vA;
      red[(cs * N + j + H) * N + k] = vB;
    }
  __syncthreads();
  double2 *o = out + ((cell * SL + sl) * N + i) * (long long)(N * N);
  for (int t = tid; t < N * N; t += NT) {
    double2 s = red[t];
    #pragma unroll
    for (int q = 1; q < CS; q++) { const double2 v = red[q * N * N + t]; s.x += v.x; s.y += v.y; }
    o[t] = s;
  }
}

// sum the SL partial spectra of each cell in a fixed order
__global__ void k_sum_partials(const double2 *__restrict__ part, double2 *__restrict__ q, int N3, int SL, long long total)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long cell = t / N3; const int e = (int)(t % N3);
  double2 s = part[(cell * SL) * N3 + e];
  for (int p = 1; p < SL; p++) { const double2 v = part[(cell * SL + p) * N3 + e]; s.x += v.x; s.y += v.y; }
  q[t] = s;
}

template <int N, int R, int JL, int CSL, int CSW, int MINB>
static int launch_tiled(lpgpu_ctx *c, const double *fhat, double *q, int B)
{
  using C = CqCfg<N, R, JL, CSL, CSW>;
  auto kern = k_computeQ_tiled<N, R, JL, CSL, CSW, MINB>;
  LP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  // split the l-window when there are too few (cell, i) CTAs to fill 148 SMs
  int SL = 1;
  if ((long long)B * N < 2 * 148) {
    SL = (int)((2 * 148) / ((long long)B * N));      // keep all CTAs in one resident wave (2 per SM)
    if (SL < 1) SL = 1;
    if (SL > N / 2) SL = N / 2;
    while (SL > 1 && (size_t)B * SL > c->cap_part) SL--;
  }
  double2 *out = reinterpret_cast<double2 *>(SL > 1 ? c->d_qpart : q);
  dim3 grid(N * B, SL, 1);
  const bool prof = c->prof_on == 1 && c->prof_used + 2 <= c->prof_ev.size();
  if (prof) LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used], c->stream));
  kern<<<grid, C::NT, C::SMEM, c->stream>>>(reinterpret_cast<const double2 *>(fhat), out, c->d_G, c->d_eta, c->d_Etab, SL);
  LP_LAUNCHED(c);
  if (prof) { LP_CUDA(cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream)); c->prof_used += 2; }
  if (SL > 1) {
    const long long total = (long long)B * c->N3;
    k_sum_partials<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(out, reinterpret_cast<double2 *>(q), c->N3, SL, total);
    LP_LAUNCHED(c);
  }
  return LPGPU_OK;
}

// returns -1 when N has no tiled instantiation (caller falls back to the simple kernel)
int lp_launch_computeQ_tiled(lpgpu_ctx *c, const double *fhat, double *q, int B)
{
  // developer knob for tile-shape experiments (not part of the ABI): LPGPU_CQ_SHAPE=1 -> 512-thread CTAs
  static const int shape = getenv("LPGPU_CQ_SHAPE") ? atoi(getenv("LPGPU_CQ_SHAPE")) : 0;
  if (c->p.N == 32 && shape == 1) return launch_tiled<32, 4, 16, 2, 4, 1>(c, fhat, q, B);
  switch (c->p.N) {
    case 32: return launch_tiled<32, 4, 16, 2, 2, 2>(c, fhat, q, B);
    case 24: return launch_tiled<24, 3, 4, 8, 1, 1>(c, fhat, q, B);
    case 16: return launch_tiled<16, 4, 8, 4, 1, 4>(c, fhat, q, B);
    case 8: return launch_tiled<8, 2, 4, 8, 1, 8>(c, fhat, q, B);
    default: return -1;
  }
}
