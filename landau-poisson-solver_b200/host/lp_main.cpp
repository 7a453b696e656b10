// lpsolver -- host driver above the C ABI (include/lpgpu.h), the counterpart of the reference's
// `solver` binary (LP_ompi.cpp:76-952) for the in-scope physics: it reads ./LPsolver-input.txt in the
// reference's GRVY syntax, sets the initial condition on the host (SetInit_1.cpp:68-325), runs the
// time loop with every hot-path call going to the GPU, and writes Data/Moments_*.dc (one row per step,
// row 1 = initial state; format LP_ompi.cpp:622-630, 836-844), Data/EntropyVals_*.dc (:632, :846) and the final Data/U_*.dc checkpoint
// (raw doubles, LP_ompi.cpp:896).  `Second = True` restarts from the last U in Data/<Second/Name>
// (LP_ompi.cpp:529-571).  All five decks of the reference's test suite run (Damping / TwoStream / FourHump / Doping ICs,
// Homogeneous, FullandLinear, LinearLandau, MassConsOnly; gamma = -3, 0, 1); TwoHump stops with an error, as the
// reference does for bad input (exit(1)).
//
// usage: lpsolver [input-file] [--device k] [--quiet] [--ranks R]
//
// --ranks R: R processes, one GPU each (rank r on device (k + r) mod #devices), forked before any CUDA call; rank r owns
// the x cells [r Nx/R, (r+1) Nx/R) -- the reference's chunk_Nx split (LP_ompi.cpp:208-220) without MPI: the ranks exchange
// their CUDA IPC handles once over socket pairs (lpgpu_peer_export / _import), after which every timestep's halo planes
// and densities travel between the GPUs inside the kernels; per step only the diagnostics' partial sums (a few doubles
// per cell) go to rank 0, which writes every file.
#include "../../include/lpgpu.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

namespace {

struct Deck {
  std::map<std::string, std::string> kv;
  bool load(const std::string &path)
  {
    std::ifstream in(path);
    if (!in.good()) return false;
    std::string line, section;
    while (std::getline(in, line)) {
      bool quoted = false;
      size_t cut = std::string::npos;
      for (size_t i = 0; i < line.size(); i++) {
        if (line[i] == '\'' || line[i] == '"') quoted = !quoted;
        if (line[i] == '#' && !quoted) { cut = i; break; }
      }
      if (cut != std::string::npos) line.erase(cut);
      auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      line = trim(line);
      if (line.empty()) continue;
      if (line[0] == '[') { section = trim(line.substr(1, line.find(']') - 1)); continue; }
      size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string k = trim(line.substr(0, eq)), v = trim(line.substr(eq + 1));
      if (v.size() >= 2 && (v[0] == '\'' || v[0] == '"') && v.back() == v[0]) v = v.substr(1, v.size() - 2);
      kv[section.empty() ? k : section + "/" + k] = v;
    }
    return true;
  }
  bool has(const std::string &k) const { return kv.count(k) != 0; }
  std::string str(const std::string &k, const std::string &d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
  double num(const std::string &k, double d) const { return has(k) ? atof(kv.at(k).c_str()) : d; }
  int integer(const std::string &k, int d) const { return has(k) ? atoi(kv.at(k).c_str()) : d; }
  bool flag(const std::string &k) const
  {
    std::string s = str(k, "false");
    for (auto &c : s) c = (char)tolower(c);
    return s == "true" || s == "1" || s == "yes";
  }
};

[[noreturn]] void die(const std::string &msg)
{
  fprintf(stderr, "Program cannot run... %s\n", msg.c_str());
  exit(1);
}

// 5-point Gauss-Legendre rule of the reference's projections (advection_1.cpp:12-13)
const double GW[5] = {0.5688888888888889, 0.4786286704993665, 0.4786286704993665, 0.2369268850561891, 0.2369268850561891};
const double GT[5] = {0., -0.5384693101056831, 0.5384693101056831, -0.9061798459386640, 0.9061798459386640};

struct Grid { int Nx, Nv; double Lv, Lx, dv, dx; };
double vcentre(const Grid &g, double m) { return -g.Lv + (m + 0.5) * g.dv; }
double xcentre(const Grid &g, double m) { return (m + 0.5) * g.dx; }
double maxwell3(double a, double b, double c, double T) { return exp(-(a * a + b * b + c * c) / (2 * T)) / (2 * M_PI * T * sqrt(2 * T * M_PI)); }
double twogauss(double a, double b, double c)
{
  const double s = M_PI / 10;
  return 0.5 * (exp(-((a - 2 * s) * (a - 2 * s) + b * b + c * c) / (2 * s * s)) + exp(-((a + 2 * s) * (a + 2 * s) + b * b + c * c) / (2 * s * s)))
         / (2 * M_PI * s * s * sqrt(2 * M_PI * s * s));
}
// moments of a velocity profile over DG cell (j1,j2,j3) against {1, xi1, xi2, xi3, |xi|^2}
void cell_moments(const Grid &g, bool two, const double sh[3], int j1, int j2, int j3, double t[5], double T = 0.4)
{
  for (int l = 0; l < 5; l++) t[l] = 0.;
  for (int a = 0; a < 5; a++) for (int b = 0; b < 5; b++) for (int c = 0; c < 5; c++) {
    const double v1 = vcentre(g, j1) + 0.5 * g.dv * GT[a] + sh[0], v2 = vcentre(g, j2) + 0.5 * g.dv * GT[b] + sh[1],
                 v3 = vcentre(g, j3) + 0.5 * g.dv * GT[c] + sh[2];
    const double w = GW[a] * GW[b] * GW[c] * (two ? twogauss(v1, v2, v3) : maxwell3(v1, v2, v3, T));
    t[0] += w; t[1] += w * 0.5 * GT[a]; t[2] += w * 0.5 * GT[b]; t[3] += w * 0.5 * GT[c];
    t[4] += w * 0.25 * (GT[a] * GT[a] + GT[b] * GT[b] + GT[c] * GT[c]);
  }
  for (int l = 0; l < 5; l++) t[l] *= 0.125;
}
// SetInit_LD (SetInit_1.cpp:68-123): Damping (Maxwellian) or TwoStream (two Gaussians) times 1 + A cos(kx)
void ic_perturbed(const Grid &g, bool two, double A, double kw, std::vector<double> &U)
{
  const double zero[3] = {0, 0, 0};
  const int sv = g.Nv * g.Nv * g.Nv;
  for (int j1 = 0; j1 < g.Nv; j1++) for (int j2 = 0; j2 < g.Nv; j2++) for (int j3 = 0; j3 < g.Nv; j3++) {
    double t[5]; cell_moments(g, two, zero, j1, j2, j3, t);
    for (int i = 0; i < g.Nx; i++) {
      double *u = &U[6 * ((size_t)i * sv + (j1 * g.Nv + j2) * g.Nv + j3)];
      const double xp = xcentre(g, i + 0.5), xm = xcentre(g, i - 0.5);
      const double xf = g.dx + (sin(kw * xp) - sin(kw * xm)) * A / kw, tp0 = xf * t[0] / g.dx, tp5 = xf * t[4] / g.dx;
      u[0] = 19 * tp0 / 4. - 15 * tp5;
      u[5] = 60 * tp5 - 15 * tp0;
      u[1] = (0.5 * (sin(kw * xp) + sin(kw * xm)) + (cos(kw * xp) - cos(kw * xm)) / (kw * g.dx)) * (A / kw) * t[0] * 12. / g.dx;
      u[2] = xf * t[1] * 12 / g.dx; u[3] = xf * t[2] * 12 / g.dx; u[4] = xf * t[3] * 12 / g.dx;
    }
  }
}
// DopingProfile (FieldCalculations.cpp:413-425) with a_i = Nx/3 - 1, b_i = 2Nx/3 - 1 (LP_ompi.cpp:160-161)
double doping_profile(int Nx, int i, double NL, double NH) { return (i <= Nx / 3 - 1 || i > 2 * Nx / 3 - 1) ? NH : NL; }
// SetInit_ND (SetInit_1.cpp:125-173): ND(x_i) times a Maxwellian of temperature T_R
void ic_doping(const Grid &g, double NL, double NH, double T0, std::vector<double> &U)
{
  const double zero[3] = {0, 0, 0};
  const int sv = g.Nv * g.Nv * g.Nv;
  for (int j1 = 0; j1 < g.Nv; j1++) for (int j2 = 0; j2 < g.Nv; j2++) for (int j3 = 0; j3 < g.Nv; j3++) {
    double t[5]; cell_moments(g, false, zero, j1, j2, j3, t, T0);
    for (int i = 0; i < g.Nx; i++) {
      double *u = &U[6 * ((size_t)i * sv + (j1 * g.Nv + j2) * g.Nv + j3)];
      const double ND = doping_profile(g.Nx, i, NL, NH), tp0 = ND * t[0], tp5 = ND * t[4];
      u[0] = 19 * tp0 / 4. - 15 * tp5;
      u[5] = 60 * tp5 - 15 * tp0;
      u[1] = 0;
      u[2] = ND * t[1] * 12; u[3] = ND * t[2] * 12; u[4] = ND * t[3] * 12;
    }
  }
}
// SetInit_4H (SetInit_1.cpp:175-258) and SetInit_4H_Homo (:261-325)
void ic_four_hump(const Grid &g, bool homogeneous, std::vector<double> &U)
{
  const int sv = g.Nv * g.Nv * g.Nv, ncell = homogeneous ? 1 : g.Nx;
  std::fill(U.begin(), U.end(), 0.);
  const double C = homogeneous ? 0.02 : 1.;
  for (int p = 0; p < 4; p++) {
    const double sa = C * pow(-1, p), sb = C * pow(-1, p / 2);
    const double sh[3] = {homogeneous ? sb : sa, sa, sa};
    for (int j1 = 0; j1 < g.Nv; j1++) for (int j2 = 0; j2 < g.Nv; j2++) for (int j3 = 0; j3 < g.Nv; j3++) {
      double t[5]; cell_moments(g, false, sh, j1, j2, j3, t);
      for (int i = 0; i < ncell; i++) {
        double x0 = 1., x1 = 0.;
        if (!homogeneous) {
          x0 = 0.;
          for (int m = 0; m < 5; m++) {
            const double x = xcentre(g, i) + 0.5 * g.dx * GT[m] - g.Lx / 2 + sb, w = GW[m] * exp(-x * x / 0.8) / sqrt(0.8 * M_PI);
            x0 += w; x1 += w * 0.5 * GT[m];
          }
          x0 *= 0.5; x1 *= 0.5;
        }
        double *u = &U[6 * ((size_t)i * sv + (j1 * g.Nv + j2) * g.Nv + j3)];
        const double tp0 = x0 * t[0], tp5 = x0 * t[4];
        u[0] += 19 * tp0 / 4. - 15 * tp5; u[5] += 60 * tp5 - 15 * tp0;
        u[1] += x1 * t[0] * 12; u[2] += x0 * t[1] * 12; u[3] += x0 * t[2] * 12; u[4] += x0 * t[3] * 12;
      }
    }
  }
  for (auto &v : U) v /= 4;
}

void make_parent_dir(const std::string &path)
{
  for (size_t pos = path.find('/'); pos != std::string::npos; pos = path.find('/', pos + 1))
    if (pos > 0) mkdir(path.substr(0, pos).c_str(), 0755);
}
// ---- the ranks of a --ranks R run: rank 0 is the parent and talks to every child over a socket pair ----
struct Team {
  int rank = 0, world = 1;
  std::vector<int> fd;                       // rank 0: fd[r] to child r; child: fd[0] to the parent
  static void wr(int f, const void *b, size_t n)
  {
    const char *p = (const char *)b;
    while (n) { ssize_t k = write(f, p, n); if (k <= 0) die("a rank of the run has died (write)"); p += k; n -= (size_t)k; }
  }
  static void rd(int f, void *b, size_t n)
  {
    char *p = (char *)b;
    while (n) { ssize_t k = read(f, p, n); if (k <= 0) die("a rank of the run has died (read)"); p += k; n -= (size_t)k; }
  }
  void start(int R)
  {
    world = R; fd.assign(R, -1);
    for (int r = 1; r < R; r++) {
      int sp[2];
      if (socketpair(AF_UNIX, SOCK_STREAM, 0, sp) != 0) die("socketpair failed");
      const pid_t pid = fork();
      if (pid < 0) die("fork failed");
      if (pid == 0) {                        // child r
        for (int q = 1; q < r; q++) close(fd[q]);
        close(sp[0]);
        rank = r; fd.assign(1, sp[1]);
        return;
      }
      close(sp[1]); fd[r] = sp[0];
    }
  }
  // concatenation of every rank's n doubles in rank order (on rank 0; children only send)
  void gather(const double *local, size_t n, std::vector<double> &all)
  {
    if (rank) { wr(fd[0], local, n * sizeof(double)); return; }
    all.resize(n * world);
    memcpy(all.data(), local, n * sizeof(double));
    for (int r = 1; r < world; r++) rd(fd[r], all.data() + n * r, n * sizeof(double));
  }
  void bcast(void *buf, size_t bytes)
  {
    if (rank) { rd(fd[0], buf, bytes); return; }
    for (int r = 1; r < world; r++) wr(fd[r], buf, bytes);
  }
  void barrier() { double x = 0; std::vector<double> all; gather(&x, 1, all); bcast(&x, sizeof x); }
  void finish()
  {
    if (rank) { _exit(0); }
    for (int r = 1; r < world; r++) { int st = 0; wait(&st); }
  }
};

#define CHECK(call)                                                                          \
  do { int rc_ = (call); if (rc_ != LPGPU_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, lpgpu_last_error()); exit(1); } } while (0)

} // namespace

int main(int argc, char **argv)
{
  std::string input = "./LPsolver-input.txt";
  int device = 0, nranks = 1; bool quiet = false;
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--device") && a + 1 < argc) device = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--ranks") && a + 1 < argc) nranks = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--quiet")) quiet = true;
    else input = argv[a];
  }
  Deck d;
  if (!d.load(input)) die("The file " + input + " cannot be found. Please create an appropriate input file before running again.");

  const char *ics[] = {"Damping", "TwoStream", "FourHump", "TwoHump", "Doping"};
  std::string ic; int nic = 0;
  for (const char *n : ics) if (d.flag(n)) { ic = n; nic++; }
  if (nic == 0) die("No initial condition has been chosen.");
  if (nic > 1) die("Please ONLY set ONE of Damping, TwoStream, FourHump or TwoHump to true.");
  if (d.flag("First") == d.flag("Second")) die("Need to choose if this is a first run or a subsequent one (First / Second).");
  if (ic == "TwoHump") die("TwoHump initial conditions are not part of the GPU hot path.");
  if (!d.has("flag")) die("Please set the name of 'flag' in the input file.");
  for (const char *n : {"nT", "Nx", "Nv", "N", "nu", "dt"}) if (!d.has(n)) die(std::string("Please set ") + n + " in the input file.");

  lpgpu_params p;
  memset(&p, 0, sizeof(p));
  const int nT = d.integer("nT", 0);
  p.Nx = d.integer("Nx", 0); p.Nv = d.integer("Nv", 0); p.N = d.integer("N", 0);
  p.nu = d.num("nu", 0.); p.dt = d.num("dt", 0.); p.gamma = d.integer("gamma", -3);
  p.homogeneous = d.flag("Homogeneous");
  if (!d.has(ic + "/Lv")) die("Please set " + ic + "/Lv in the input file.");
  p.Lv = d.num(ic + "/Lv", 0.);
  double A_amp = d.num(ic + "/A_amp", 0.), k_wave = d.num(ic + "/k_wave", 0.5);
  if (ic == "TwoStream") { if (!d.has("TwoStream/Lx")) die("Please set TwoStream/Lx."); p.Lx = d.num("TwoStream/Lx", 0.); k_wave = 2 * M_PI / 4.; }
  else p.Lx = d.num(ic + "/Lx", 2 * M_PI / k_wave);
  if (p.homogeneous && ic != "FourHump") die("Trying to run the space homogeneous code, but current IC is not available (only FourHump).");
  if (nranks < 1 || nranks > 8) die("--ranks must be between 1 and 8");
  if (nranks > 1 && (p.homogeneous || p.Nx % nranks != 0)) die("--ranks needs an inhomogeneous run with Nx divisible by the number of ranks");
  Team team;
  if (nranks > 1) team.start(nranks);                                 // forks: no CUDA call has been made yet
  if (team.rank) quiet = true;
  const int ndev = lpgpu_device_count();
  p.x_count = p.homogeneous ? 1 : p.Nx / team.world;
  p.x_begin = p.homogeneous ? 0 : team.rank * p.x_count;
  p.device = ndev > 0 ? (device + team.rank) % ndev : device; p.computeq_variant = 0;
  p.full_and_linear = d.flag("FullandLinear");
  p.linear_landau = d.flag("LinearLandau");
  p.mass_cons_only = d.flag("MassConsOnly");
  if (ic == "Doping") {                                              // ReadDopingParameters, InputParsing.cpp:512-570
    for (const char *n : {"Doping/NL", "Doping/NH", "Doping/eps"}) if (!d.has(n)) die(std::string("Please set ") + n + " in the input file.");
    p.doping = 1;
    p.NL = d.num("Doping/NL", 0.); p.NH = d.num("Doping/NH", 0.); p.eps = d.num("Doping/eps", 1.);
    p.T_L = d.num("Doping/T_L", 0.4); p.T_R = d.num("Doping/T_R", 0.4);
  }

  char name[512], tail[400];
  const std::string flag = d.str("flag");
  if (p.homogeneous) snprintf(tail, sizeof tail, "nu%gA%gk%gNv%dLv%gSpectralN%ddt%gnT%d_%s.dc", p.nu, A_amp, k_wave, p.Nv, p.Lv, p.N, p.dt, nT, flag.c_str());
  else snprintf(tail, sizeof tail, "nu%gA%gk%gNx%dLx%gNv%dLv%gSpectralN%ddt%gnT%d_%s.dc", p.nu, A_amp, k_wave, p.Nx, p.Lx, p.Nv, p.Lv, p.N, p.dt, nT, flag.c_str());
  snprintf(name, sizeof name, "Data/Moments_%s", tail);
  const std::string fmom_name = name;
  snprintf(name, sizeof name, "Data/U_%s", tail);
  const std::string fu_name = name;
  snprintf(name, sizeof name, "Data/EntropyVals_%s", tail);
  const std::string fent_name = name;
  snprintf(name, sizeof name, "Data/Marginals_%s", tail);
  const std::string fmarg_name = name;
  snprintf(name, sizeof name, "Data/PhiVals_%s", tail);
  const std::string fphi_name = name;
  snprintf(name, sizeof name, "Data/FieldVals_%s", tail);
  const std::string fE_name = name;

  const int sv = p.Nv * p.Nv * p.Nv, ncell = p.x_count, ncell_all = p.homogeneous ? 1 : p.Nx;
  std::vector<double> U((size_t)6 * sv * ncell_all);                   // every rank sets the whole initial state and keeps its shard
  Grid g = {p.Nx, p.Nv, p.Lv, p.Lx, 2. * p.Lv / p.Nv, p.homogeneous ? 1. : p.Lx / p.Nx};
  if (d.flag("Second")) {
    if (!d.has("Second/Name")) die("Please set the name of the file from the previous run under Second/Name.");
    FILE *f = fopen(("Data/" + d.str("Second/Name")).c_str(), "rb");
    if (!f) die("cannot open Data/" + d.str("Second/Name"));
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f), need = (long)(U.size() * sizeof(double));
    if (bytes < need || bytes % need != 0) die("Error reading file (size does not match 6*size doubles)");
    fseek(f, bytes - need, SEEK_SET);
    if (fread(U.data(), sizeof(double), U.size(), f) != U.size()) die("Error reading file");
    fclose(f);
  } else if (ic == "Doping") ic_doping(g, p.NL, p.NH, p.T_R, U);
  else if (ic == "FourHump") ic_four_hump(g, p.homogeneous, U);
  else ic_perturbed(g, ic == "TwoStream", A_amp, k_wave, U);

  const size_t shard = (size_t)6 * sv * ncell;
  if (team.world > 1) { std::vector<double> mine(U.begin() + (size_t)team.rank * shard, U.begin() + (size_t)(team.rank + 1) * shard); U.swap(mine); }
  lpgpu_ctx *ctx = nullptr;
  CHECK(lpgpu_init(&p, &ctx));
  CHECK(lpgpu_upload_U(ctx, U.data()));
  if (team.world > 1) {
    // the ranks map each other's stage buffers and mailboxes once; from then on lpgpu_step works on the shard
    std::vector<double> blob(LPGPU_PEER_HANDLE_BYTES / sizeof(double)), blobs;
    CHECK(lpgpu_peer_export(ctx, blob.data()));
    team.gather(blob.data(), blob.size(), blobs);
    blobs.resize(blob.size() * team.world);
    team.bcast(blobs.data(), blobs.size() * sizeof(double));
    CHECK(lpgpu_peer_import(ctx, team.rank, team.world, blobs.data()));
  }
  if (p.linear_landau && p.nu > 0.) CHECK(lpgpu_set_maxwellian(ctx));   // ComputeDFTofMaxwellian(U, f, DFTMaxwell), LP_ompi.cpp:516, :541
  const bool root = team.rank == 0;
  if (root) make_parent_dir(fmom_name);
  FILE *fmom = fopen(root ? fmom_name.c_str() : "/dev/null", "w");
  if (!fmom) die("cannot open " + fmom_name);
  FILE *fent = fopen(root ? fent_name.c_str() : "/dev/null", "w");
  if (!fent) die("cannot open " + fent_name);

  // the rows of one step from the numbers lpgpu_diagnostics_end (or the synchronous pair) returns
  auto report = [&](int step, const double *m5_loc, const std::vector<double> &ms_loc, const double *d4_loc) {
    // the partial sums of every rank (MPI_Reduce / MPI_Gather in the reference's terms): moments and diagnostics add up,
    // the per-cell densities concatenate in rank order = x order
    double m5[5], d4[4];
    std::vector<double> ms;
    if (team.world > 1) {
      double loc[9]; memcpy(loc, m5_loc, sizeof m5); memcpy(loc + 5, d4_loc, sizeof d4);
      std::vector<double> all;
      team.gather(loc, 9, all);
      team.gather(ms_loc.data(), ms_loc.size(), ms);
      if (!root) return;
      for (int k = 0; k < 5; k++) { m5[k] = 0.; for (int r = 0; r < team.world; r++) m5[k] += all[9 * r + k]; }
      for (int k = 0; k < 4; k++) { d4[k] = 0.; for (int r = 0; r < team.world; r++) d4[k] += all[9 * r + 5 + k]; }
    } else { memcpy(m5, m5_loc, sizeof m5); memcpy(d4, d4_loc, sizeof d4); ms = ms_loc; }
    double ele = 0.;                                // d4: entropy, KiE over positive / negative cells, #negative cells
    if (!p.homogeneous) CHECK(lpgpu_eleE_from_ms(&p, ms.data(), &ele));
    const double ent = d4[0], lent = log(fabs(ent));
    fprintf(fent, "%11.8g %11.8g %11.8g \n", ent, lent, log(fabs(lent)));   // LP_ompi.cpp:632, 846
    if (!quiet) printf("entropy = %11.8g, Kinetic Energy Ratio = %g\n", ent, d4[2] / d4[1]);
    if (p.homogeneous) {
      if (!quiet) printf("step %d: %11.8g  %11.8g  %11.8g  %11.8g  %11.8g \n", step, m5[0], m5[1], m5[2], m5[3], m5[4]);
      fprintf(fmom, "%11.8g %11.8g %11.8g %11.8g %11.8g %11.8g %11.8g %11.8g \n", m5[0], m5[1], m5[2], m5[3], 0.0, 0.0, 0.0, m5[4]);
    } else {
      const double t = sqrt(ele);
      if (!quiet) printf("step %d: %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g %11.8g \n", step, m5[0], m5[1], m5[2], m5[3], m5[4], ele, t, log(t), m5[4] + ele);
      fprintf(fmom, "%11.8g %11.8g %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g \n", m5[0], m5[1], m5[2], m5[3], m5[4], ele, t, log(t), m5[4] + ele);
    }
  };
  auto diagnostics = [&](int step) {                // synchronous: the state now on the device
    double m5[5], d4[4];
    std::vector<double> ms((size_t)2 * ncell);
    CHECK(lpgpu_moments_partial(ctx, m5, ms.data()));
    CHECK(lpgpu_diagnostics_partial(ctx, d4));
    report(step, m5, ms, d4);
  };
  auto collect = [&](int step) {                    // the snapshot lpgpu_diagnostics_begin took after `step`
    double m5[5], d4[4];
    std::vector<double> ms((size_t)2 * ncell);
    CHECK(lpgpu_diagnostics_end(ctx, m5, ms.data(), d4));
    report(step, m5, ms, d4);
  };
  // ---- Marginals_*.dc / PhiVals_*.dc / FieldVals_*.dc: first the evaluation points, then one row for the initial state
  // and one every 20 steps (LP_ompi.cpp:648-655, :868-875).  The GPU reduces over the integrated-out velocity
  // directions (lpgpu_marginal_sums) and over whole velocity space ((m_i, s_i) of lpgpu_moments_partial); the host
  // evaluates the reference's closed forms at its 4 sub-points per cell.
  FILE *fmarg = fopen(root ? fmarg_name.c_str() : "/dev/null", "w"), *fphi = fopen(root ? fphi_name.c_str() : "/dev/null", "w"),
       *fE = fopen(root ? fE_name.c_str() : "/dev/null", "w");
  if (!fmarg || !fphi || !fE) die("cannot open the Marginals/PhiVals/FieldVals files under Data/");
  const int np = 4, Nv = p.Nv;
  const double dv = g.dv, dx = g.dx, ddv = dv / np, ddx = dx / np;
  auto gridv = [&](double j) { return -p.Lv + (j + 0.5) * dv; };     // Gridv / Gridx, advection_1.cpp:15-21
  auto gridx = [&](double i) { return (i + 0.5) * dx; };
  if (p.homogeneous) {                                               // PrintMarginalLoc_Homo, MarginalCreation.cpp:118-158
    for (int row = 0; row < 2; row++) {
      for (int j1 = 0; j1 < Nv; j1++) for (int n1 = 0; n1 < np; n1++) for (int j2 = 0; j2 < Nv; j2++) for (int n2 = 0; n2 < np; n2++)
        fprintf(fmarg, "%11.8g  ", row == 0 ? gridv(j1 - 0.5) + n1 * ddv : gridv(j2 - 0.5) + n2 * ddv);
      fprintf(fmarg, "\n");
    }
  } else {                                                           // PrintMarginalLoc_Inhomo :81-116, PrintFieldLoc FieldCalculations.cpp:40-58
    for (int row = 0; row < 2; row++) {
      for (int i = 0; i < p.Nx; i++) for (int nx = 0; nx < np; nx++) for (int j1 = 0; j1 < Nv; j1++) for (int nv = 0; nv < np; nv++)
        fprintf(fmarg, "%11.8g  ", row == 0 ? gridx(i - 0.5) + nx * ddx : gridv(j1 - 0.5) + nv * ddv);
      fprintf(fmarg, "\n");
    }
    for (int i = 0; i < p.Nx; i++) for (int nx = 0; nx < np; nx++) {
      fprintf(fphi, "%11.8g  ", gridx(i - 0.5) + nx * ddx);
      fprintf(fE, "%11.8g  ", gridx(i - 0.5) + nx * ddx);
    }
    fprintf(fphi, "\n"); fprintf(fE, "\n");
  }
  std::vector<double> msum_loc((size_t)4 * (p.homogeneous ? Nv * Nv : ncell * Nv)), msum;
  auto print_marginal_and_field = [&]() {
    CHECK(lpgpu_marginal_sums(ctx, msum_loc.data()));
    double m5[5];
    std::vector<double> ms_loc((size_t)2 * ncell), ms;
    if (!p.homogeneous) CHECK(lpgpu_moments_partial(ctx, m5, ms_loc.data()));
    if (team.world > 1) {                                            // every rank's sums to rank 0, in x order
      team.gather(msum_loc.data(), msum_loc.size(), msum);
      team.gather(ms_loc.data(), ms_loc.size(), ms);
      if (!root) return;
    } else { msum = msum_loc; ms = ms_loc; }
    if (p.homogeneous) {                                             // PrintMarginal_Homo :196-219 with f_marg_Homo :43-59
      for (int j1 = 0; j1 < Nv; j1++) for (int n1 = 0; n1 < np; n1++) for (int j2 = 0; j2 < Nv; j2++) for (int n2 = 0; n2 < np; n2++) {
        const double d1 = gridv(j1 - 0.5) + n1 * ddv - gridv(j1), d2 = gridv(j2 - 0.5) + n2 * ddv - gridv(j2);
        const double *q = &msum[(size_t)4 * (j1 * Nv + j2)];
        fprintf(fmarg, "%11.8g  ", dv * q[0] + q[1] * d1 + q[2] * d2 + q[3] * ((d1 * d1 + d2 * d2) / dv + dv / 12));
      }
      fprintf(fmarg, "\n");
      return;
    }
    for (int i = 0; i < p.Nx; i++) for (int nx = 0; nx < np; nx++) for (int j1 = 0; j1 < Nv; j1++) for (int nv = 0; nv < np; nv++) {
      const double xd = gridx(i - 0.5) + nx * ddx - gridx(i), vd = gridv(j1 - 0.5) + nv * ddv - gridv(j1);   // PrintMarginal_Inhomo :162-194
      const double *q = &msum[(size_t)4 * (i * Nv + j1)];
      fprintf(fmarg, "%11.8g  ", dv * dv * q[0] + dv * dv * q[1] * xd / dx + dv * q[2] * vd + q[3] * (vd * vd + dv * dv / 6));
    }
    fprintf(fmarg, "\n");
    // PrintFieldData_Normal (FieldCalculations.cpp:331-355): phi at 4 points per cell; computePhi_Normal (:244-329) in
    // terms of m_q = scalev sum(U0 + U5/4), s_q = scalev sum U1 and computePhi_x_0 (:223-243)
    double P = 0., acc = 0.;
    std::vector<double> Pq(p.Nx), Sq(p.Nx);                           // P_q = sum_{q' < q} m_q',  S_q = sum_{q' < q} (P_q' + m_q'/2 - s_q'/12)
    for (int q = 0; q < p.Nx; q++) { Pq[q] = P; Sq[q] = acc; acc += P + 0.5 * ms[2 * q] - ms[2 * q + 1] / 12.; P += ms[2 * q]; }
    if (p.doping) {
      // PrintFieldData_Doping (FieldCalculations.cpp:562-583): phi and E at 4 points per cell; computePhi_Doping (:452-522),
      // computeE_Doping (:524-560) and computePhi_x_0_Doping (:427-450) in terms of the same per-cell sums
      const int a_i = p.Nx / 3 - 1, b_i = 2 * p.Nx / 3 - 1;
      const double a_val = (a_i + 1) * dx, b_val = (b_i + 1) * dx, NL = p.NL, NH = p.NH, eps = p.eps;
      const double ce = 1. / p.Lx + 0.5 * NH * p.Lx / eps + (NL - NH) * (b_val - a_val) / eps
                        - (0.5 * (NL - NH) * (b_val * b_val - a_val * a_val) + acc * dx * dx) / (p.Lx * eps);
      for (int i = 0; i < p.Nx; i++) for (int nx = 0; nx < np; nx++) {
        const double x = gridx(i - 0.5) + nx * ddx, xd = x - gridx(i - 0.5), xm = x - gridx(i), ND = doping_profile(p.Nx, i, NL, NH);
        const double xe = xm * xm * xm / (6. * dx) - dx * xm / 8. - dx * dx / 24., xev = xm * xm / (2. * dx) - dx / 8.;
        double phi = Sq[i] * dx * dx + Pq[i] * dx * xd + ms[2 * i] * xd * xd / 2. + ms[2 * i + 1] * xe - ND * x * x / 2.;
        double E = ND * x - (ms[2 * i] * xd + ms[2 * i + 1] * xev + Pq[i] * dx);
        if (i > a_i) { phi -= (NH - NL) * a_val * (x - 0.5 * a_val); E += (NH - NL) * a_val; }
        if (i > b_i) { phi -= (NL - NH) * b_val * (x - 0.5 * b_val); E += (NL - NH) * b_val; }
        fprintf(fphi, "%11.8g ", phi / eps + ce * x);
        fprintf(fE, "%11.8g ", E / eps - ce);
      }
      fprintf(fphi, "\n"); fprintf(fE, "\n");
      return;
    }
    const double ce = 0.5 * p.Lx - acc * dx * dx / p.Lx;
    for (int i = 0; i < p.Nx; i++) for (int nx = 0; nx < np; nx++) {
      const double x = gridx(i - 0.5) + nx * ddx, xd = x - gridx(i - 0.5), xm = x - gridx(i);
      const double xe = xm * xm * xm / (6. * dx) - dx * xm / 8. - dx * dx / 24.;
      const double phi = Sq[i] * dx * dx + Pq[i] * dx * xd + ms[2 * i] * xd * xd / 2. + ms[2 * i + 1] * xe - x * x / 2 - ce * x;
      fprintf(fphi, "%11.8g ", phi);
    }
    fprintf(fphi, "\n");
  };
  diagnostics(0);
  print_marginal_and_field();
  const auto t0 = std::chrono::steady_clock::now();
  // The reference computes the diagnostics of step t between steps t and t+1 (LP_ompi.cpp:817-849); here they run on a side
  // stream over a snapshot while step t+1 runs, and their rows are written one step later, in the same order.
  for (int t = 0; t < nT; t++) {
    CHECK(lpgpu_step_async(ctx, 1));
    if (t > 0) collect(t);                                           // rows of step t while step t+1 runs
    CHECK(lpgpu_diagnostics_begin(ctx));                             // snapshot of the state after step t+1
    if (t % 20 == 0) print_marginal_and_field();                     // after steps 1, 21, 41, ...: the reference tests its 0-based counter (LP_ompi.cpp:79, :868)
  }
  if (nT > 0) collect(nT);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (root) printf("\nTime duration for %d time steps is %gs\n\n", nT, secs);
  fclose(fmom);
  fclose(fent);
  fclose(fmarg); fclose(fphi); fclose(fE);
  CHECK(lpgpu_download_U(ctx, U.data()));
  std::vector<double> Uall;
  if (team.world > 1) team.gather(U.data(), U.size(), Uall);          // the shards in rank order are the reference's U
  if (root) {
    const std::vector<double> &Uw = team.world > 1 ? Uall : U;
    FILE *fu = fopen(fu_name.c_str(), "wb");
    if (fu) { fwrite(Uw.data(), sizeof(double), Uw.size(), fu); fclose(fu); }
  }
  if (team.world > 1) team.barrier();                                 // nobody unmaps or frees while a peer may still write
  CHECK(lpgpu_finalize(ctx));
  team.finish();
  return 0;
}
