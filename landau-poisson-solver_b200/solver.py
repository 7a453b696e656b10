"""Host-side mirror of the reference driver for the hot path (LP_ompi.cpp:76-952), above the C ABI.

What stays on the host, as in the reference: reading ``LPsolver-input.txt`` (InputParsing.cpp),
the initial conditions (SetInit_1.cpp:68-325), the time loop, and writing ``Data/Moments_*.dc``
(LP_ompi.cpp:442-472, 622-630, 836-844).  What runs on the GPU through include/lpgpu.h: RK3,
setInit_spectral, ComputeQ, conserveMoments, RK4 and the moment reductions.

Sharding (replaces the reference's MPI split, LP_ompi.cpp:205-220): one process per GPU, each owning
a contiguous block of x cells.  Collisions need no communication.  Each SSP-RK3 stage needs one
all-gather of the per-cell sums (m_i, s_i) and one periodic halo plane from each x neighbour
(advection_1.cpp:297,306); both go through ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).
"""
import math
import os
import sys

import numpy as np

from . import lpgpu

# 5-point Gauss-Legendre rule used by every SetInit_* (advection_1.cpp:12-13)
_GW = np.array([0.5688888888888889, 0.4786286704993665, 0.4786286704993665, 0.2369268850561891, 0.2369268850561891])
_GT = np.array([0., -0.5384693101056831, 0.5384693101056831, -0.9061798459386640, 0.9061798459386640])


# ------------------------------------------------------------------------------------------------
# input file (GRVY syntax subset the reference uses: key = value  # comment, [Section], True/False, 'str')
def parse_input_file(path):
    kv, section = {}, ""
    with open(path) as fh:
        for raw in fh:
            line, inq = "", False
            for ch in raw:
                if ch in "'\"":
                    inq = not inq
                if ch == "#" and not inq:
                    break
                line += ch
            line = line.strip()
            if not line:
                continue
            if line.startswith("["):
                section = line[1:line.index("]")].strip()
                continue
            if "=" not in line:
                continue
            k, v = (s.strip() for s in line.split("=", 1))
            if len(v) >= 2 and v[0] in "'\"" and v[-1] == v[0]:
                v = v[1:-1]
            kv[(section + "/" if section else "") + k] = v
    return kv


def _bool(kv, key, default=False):
    return kv.get(key, str(default)).strip().lower() in ("true", "1", "yes", ".true.")


class RunConfig:
    """The scalars main() derives from the input file (LP_ompi.cpp:142-184; InputParsing.cpp:291-470)."""

    def __init__(self, kv):
        self.flag = kv["flag"]
        self.nT, self.Nx, self.Nv, self.N = int(kv["nT"]), int(kv["Nx"]), int(kv["Nv"]), int(kv["N"])
        self.nu, self.dt = float(kv["nu"]), float(kv["dt"])
        self.gamma = int(kv.get("gamma", -3))
        ics = [n for n in ("Damping", "TwoStream", "FourHump", "TwoHump", "Doping") if _bool(kv, n)]
        if len(ics) != 1:
            raise ValueError("exactly one of Damping, TwoStream, FourHump, TwoHump, Doping must be True (got %s)" % ics)
        self.ic = ics[0]
        if _bool(kv, "First") == _bool(kv, "Second"):
            raise ValueError("exactly one of First / Second must be True")
        self.second = _bool(kv, "Second")
        self.second_name = kv.get("Second/Name")
        self.homogeneous = _bool(kv, "Homogeneous")
        self.full_and_linear = _bool(kv, "FullandLinear")
        self.linear_landau = _bool(kv, "LinearLandau")
        self.mass_cons_only = _bool(kv, "MassConsOnly")
        if self.ic == "TwoHump":
            raise NotImplementedError("TwoHump initial conditions are outside the GPU hot path")
        if self.gamma not in (-3, 0, 1):              # ReadGamma (InputParsing.cpp:202-238)
            raise ValueError("Currently the code can only use gamma = -3 (True Landau), 0 (Maxwell Molecules) or 1 (Hard Spheres)")
        if self.gamma != -3 and self.full_and_linear:
            raise NotImplementedError("FullandLinear is implemented for gamma = -3 only")
        self.doping = None
        if self.ic == "Doping":                      # ReadDopingParameters (InputParsing.cpp:512-570)
            if self.homogeneous:
                raise ValueError("Doping needs an inhomogeneous run")
            for k in ("NL", "NH", "eps"):
                if "Doping/" + k not in kv:
                    raise ValueError("Please set Doping/%s in the input file" % k)
            self.doping = dict(NL=float(kv["Doping/NL"]), NH=float(kv["Doping/NH"]), eps=float(kv["Doping/eps"]),
                               T_L=float(kv.get("Doping/T_L", 0.4)), T_R=float(kv.get("Doping/T_R", 0.4)))
        sec = self.ic
        self.A_amp = float(kv.get(sec + "/A_amp", 0.))
        self.k_wave = float(kv.get(sec + "/k_wave", 0.5))
        self.Lv = float(kv[sec + "/Lv"])
        if sec == "TwoStream":
            self.Lx = float(kv[sec + "/Lx"])
            self.k_wave = 2 * math.pi / 4.           # forced by the reference (InputParsing.cpp:456)
        else:
            self.Lx = float(kv.get(sec + "/Lx", 2 * math.pi / self.k_wave))
        if self.homogeneous and self.ic != "FourHump":
            raise ValueError("the homogeneous code only supports the FourHump IC (LP_ompi.cpp:477-492)")

    @classmethod
    def from_file(cls, path):
        return cls(parse_input_file(path))

    def moments_filename(self):
        """Data/Moments_... name, byte-for-byte the reference's sprintf (LP_ompi.cpp:442-443, 459-460)."""
        g = lambda x: "%g" % x
        if self.homogeneous:
            return "Data/Moments_nu%sA%sk%sNv%dLv%sSpectralN%ddt%snT%d_%s.dc" % (
                g(self.nu), g(self.A_amp), g(self.k_wave), self.Nv, g(self.Lv), self.N, g(self.dt), self.nT, self.flag)
        return "Data/Moments_nu%sA%sk%sNx%dLx%sNv%dLv%sSpectralN%ddt%snT%d_%s.dc" % (
            g(self.nu), g(self.A_amp), g(self.k_wave), self.Nx, g(self.Lx), self.Nv, g(self.Lv), self.N, g(self.dt), self.nT, self.flag)

    def initial_condition(self, x_begin=0, x_count=None):
        if self.homogeneous:
            return set_init_4h_homo(self.Nv, self.Lv)
        if self.ic in ("Damping", "TwoStream"):
            return set_init_ld(self.Nx, self.Nv, self.Lv, self.Lx, self.A_amp, self.k_wave, self.ic == "TwoStream", x_begin, x_count)
        if self.ic == "Doping":
            return set_init_nd(self.Nx, self.Nv, self.Lv, self.doping["NL"], self.doping["NH"], self.doping["T_R"], x_begin, x_count)
        return set_init_4h(self.Nx, self.Nv, self.Lv, self.Lx, x_begin, x_count)


def format_moments_row(m, homogeneous):
    """One line of Moments_*.dc (LP_ompi.cpp:622-630 / 836-844); m = mass, P1..3, KiE, EleE."""
    if homogeneous:
        return "%11.8g %11.8g %11.8g %11.8g %11.8g %11.8g %11.8g %11.8g \n" % (m[0], m[1], m[2], m[3], 0.0, 0.0, 0.0, m[4])
    t = math.sqrt(m[5])
    return "%11.8g %11.8g %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g  %11.8g \n" % (
        m[0], m[1], m[2], m[3], m[4], m[5], t, math.log(t), m[4] + m[5])


# ------------------------------------------------------------------------------------------------
# initial conditions (host, run once) -- vectorised restatement of SetInit_1.cpp
def _maxwellian(v1, v2, v3, T):
    return np.exp(-(v1 * v1 + v2 * v2 + v3 * v3) / (2 * T)) / (2 * np.pi * T * np.sqrt(2 * T * np.pi))


def _two_gauss(v1, v2, v3):
    s = np.pi / 10
    return 0.5 * (np.exp(-((v1 - 2 * s) ** 2 + v2 * v2 + v3 * v3) / (2 * s * s)) + np.exp(-((v1 + 2 * s) ** 2 + v2 * v2 + v3 * v3) / (2 * s * s))) \
        / (2 * np.pi * s * s * np.sqrt(2 * np.pi * s * s))


def _velocity_cell_moments(Nv, Lv, profile, shift=(0., 0., 0.)):
    """tmp0..tmp4 of SetInit_LD / SetInit_4H (SetInit_1.cpp:80-99): cell moments of `profile`
    against 1, xi1, xi2, xi3, |xi|^2 with the 5^3-point Gauss rule.  Returns 5 arrays [Nv,Nv,Nv]."""
    dv = 2. * Lv / Nv
    c = -Lv + (np.arange(Nv) + 0.5) * dv
    q = c[:, None] + 0.5 * dv * _GT[None, :]                          # [Nv,5] quadrature points per dimension
    V1 = q[:, None, None, :, None, None] + shift[0]
    V2 = q[None, :, None, None, :, None] + shift[1]
    V3 = q[None, None, :, None, None, :] + shift[2]
    W = _GW[:, None, None] * _GW[None, :, None] * _GW[None, None, :]
    tp = W[None, None, None] * profile(V1, V2, V3)
    t1 = 0.5 * _GT[:, None, None] * np.ones((5, 5, 5))
    t2 = 0.5 * _GT[None, :, None] * np.ones((5, 5, 5))
    t3 = 0.5 * _GT[None, None, :] * np.ones((5, 5, 5))
    t4 = 0.25 * (_GT[:, None, None] ** 2 + _GT[None, :, None] ** 2 + _GT[None, None, :] ** 2)
    out = [np.sum(tp, axis=(3, 4, 5))]
    for wgt in (t1, t2, t3, t4):
        out.append(np.sum(tp * wgt[None, None, None], axis=(3, 4, 5)))
    return [o * 0.125 for o in out]


def set_init_ld(Nx, Nv, Lv, Lx, A_amp, k_wave, twostream=False, x_begin=0, x_count=None):
    """SetInit_LD (SetInit_1.cpp:68-123): Maxwellian (T=0.4) or two-Gaussian profile in v times
    1 + A cos(k x) in x, L2-projected on the DG basis.  Returns AoS U for cells [x_begin, x_begin+x_count)."""
    x_count = Nx if x_count is None else x_count
    prof = _two_gauss if twostream else (lambda a, b, c: _maxwellian(a, b, c, 0.4))
    t0, t1, t2, t3, t4 = _velocity_cell_moments(Nv, Lv, prof)
    dx = Lx / Nx
    i = np.arange(x_begin, x_begin + x_count)
    xp, xm = (i + 0.5 + 0.5) * dx, (i - 0.5 + 0.5) * dx
    a, c = A_amp, k_wave
    xf = dx + (np.sin(c * xp) - np.sin(c * xm)) * a / c
    u1f = (0.5 * (np.sin(c * xp) + np.sin(c * xm)) + (np.cos(c * xp) - np.cos(c * xm)) / (c * dx)) * (a / c)
    U = np.empty((x_count, Nv, Nv, Nv, 6))
    X = xf[:, None, None, None]
    tp0, tp5 = X * t0[None] / dx, X * t4[None] / dx
    U[..., 0] = 19 * tp0 / 4. - 15 * tp5
    U[..., 5] = 60 * tp5 - 15 * tp0
    U[..., 1] = u1f[:, None, None, None] * t0[None] * 12. / dx
    U[..., 2] = X * t1[None] * 12 / dx
    U[..., 3] = X * t2[None] * 12 / dx
    U[..., 4] = X * t3[None] * 12 / dx
    return U.reshape(-1)


def doping_profile(Nx, NL, NH):
    """DopingProfile (FieldCalculations.cpp:413-425; a_i, b_i from LP_ompi.cpp:160-161): NH in the outer thirds, NL between."""
    i = np.arange(Nx)
    a_i, b_i = Nx // 3 - 1, 2 * Nx // 3 - 1
    return np.where((i <= a_i) | (i > b_i), NH, NL).astype(np.float64)


def set_init_nd(Nx, Nv, Lv, NL, NH, T0, x_begin=0, x_count=None):
    """SetInit_ND (SetInit_1.cpp:125-173): ND(x_i) times a Maxwellian of temperature T0 = T_R, L2-projected on the DG basis."""
    x_count = Nx if x_count is None else x_count
    t0, t1, t2, t3, t4 = _velocity_cell_moments(Nv, Lv, lambda a, b, c: _maxwellian(a, b, c, T0))
    ND = doping_profile(Nx, NL, NH)[x_begin:x_begin + x_count][:, None, None, None]
    U = np.empty((x_count, Nv, Nv, Nv, 6))
    tp0, tp5 = ND * t0[None], ND * t4[None]
    U[..., 0] = 19 * tp0 / 4. - 15 * tp5
    U[..., 5] = 60 * tp5 - 15 * tp0
    U[..., 1] = 0.
    U[..., 2] = ND * t1[None] * 12
    U[..., 3] = ND * t2[None] * 12
    U[..., 4] = ND * t3[None] * 12
    return U.reshape(-1)


def set_init_4h(Nx, Nv, Lv, Lx, x_begin=0, x_count=None):
    """SetInit_4H (SetInit_1.cpp:175-258)."""
    x_count = Nx if x_count is None else x_count
    dx, C = Lx / Nx, 1.
    i = np.arange(x_begin, x_begin + x_count)
    xq = (i[:, None] + 0.5) * dx + 0.5 * dx * _GT[None, :]
    U = np.zeros((x_count, Nv, Nv, Nv, 6))
    for p in range(4):
        sv, sx = C * (-1.) ** p, C * (-1.) ** (p // 2)
        t0, t1, t2, t3, t4 = _velocity_cell_moments(Nv, Lv, lambda a, b, c: _maxwellian(a, b, c, 0.4), (sv, sv, sv))
        T = 0.4
        tpx = _GW[None, :] * np.exp(-(xq - Lx / 2 + sx) ** 2 / (2 * T)) / np.sqrt(2 * T * np.pi)
        x0 = (0.5 * np.sum(tpx, axis=1))[:, None, None, None]
        x1 = (0.5 * np.sum(tpx * 0.5 * _GT[None, :], axis=1))[:, None, None, None]
        tp0, tp5 = x0 * t0[None], x0 * t4[None]
        U[..., 0] += 19 * tp0 / 4. - 15 * tp5
        U[..., 5] += 60 * tp5 - 15 * tp0
        U[..., 1] += x1 * t0[None] * 12
        U[..., 2] += x0 * t1[None] * 12
        U[..., 3] += x0 * t2[None] * 12
        U[..., 4] += x0 * t3[None] * 12
    return (U / 4).reshape(-1)


def set_init_4h_homo(Nv, Lv):
    """SetInit_4H_Homo (SetInit_1.cpp:261-325); U1 is 0 (the reference never sets it)."""
    C = 0.02
    U = np.zeros((Nv, Nv, Nv, 6))
    for p in range(4):
        s1, s23 = C * (-1.) ** (p // 2), C * (-1.) ** p
        t0, t1, t2, t3, t4 = _velocity_cell_moments(Nv, Lv, lambda a, b, c: _maxwellian(a, b, c, 0.4), (s1, s23, s23))
        U[..., 0] += 19 * t0 / 4. - 15 * t4
        U[..., 5] += 60 * t4 - 15 * t0
        U[..., 2] += t1 * 12
        U[..., 3] += t2 * 12
        U[..., 4] += t3 * 12
    return (U / 4).reshape(-1)


# ------------------------------------------------------------------------------------------------
# sharding helpers (pure host logic; unit-tested on CPU with gloo)
def shard_range(Nx, world, rank):
    """Contiguous block of x cells owned by `rank` (the reference's chunk_Nx, LP_ompi.cpp:208-220,
    with the constraint Nx % world == 0 instead of its Nv^3 % nprocs == 0)."""
    if Nx % world != 0:
        raise ValueError("Nx (%d) must be divisible by the number of GPUs (%d)" % (Nx, world))
    n = Nx // world
    return rank * n, n


def exchange_density(dist, world, ms_local, ms_all, group=None):
    """All-gather of the local (m_i, s_i) block: every rank then evaluates the same Nx-long field scan."""
    if ms_all.is_cuda:
        dist.all_gather_into_tensor(ms_all, ms_local, group=group)
    else:
        dist.all_gather(list(ms_all.chunk(world)), ms_local, group=group)


def exchange_halo(dist, rank, world, send_left, send_right, recv_left, recv_right, group=None):
    """The periodic halo planes: recv_left <- left neighbour's send_right, recv_right <- right neighbour's send_left."""
    left, right = (rank - 1) % world, (rank + 1) % world
    # One grouped launch (ncclGroupStart/End): separately issued NCCL sends would deadlock, each rank's
    # send waiting for a receive queued behind the peer's own send.  Order matters when left == right
    # (world == 2): the first send to a peer pairs with the first receive from it on the other side.
    ops = [dist.P2POp(dist.isend, send_right, right, group), dist.P2POp(dist.isend, send_left, left, group),
           dist.P2POp(dist.irecv, recv_left, left, group), dist.P2POp(dist.irecv, recv_right, right, group)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()


def exchange_stage(dist, rank, world, ms_local, ms_all, send_left, send_right, recv_left, recv_right):
    """The one exchange step of an SSP-RK3 stage: all-gather of (m_i, s_i) and the periodic halo
    planes.  Arguments are torch tensors (CUDA for NCCL, CPU for gloo) aliasing the buffers of
    lpgpu_exchange."""
    exchange_density(dist, world, ms_local, ms_all)
    exchange_halo(dist, rank, world, send_left, send_right, recv_left, recv_right)


class _DeviceBuffer:
    """Expose a raw device address as a CUDA array so torch can alias it without copying."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class ShardedSolver:
    """One rank of the x-sharded solver.  world == 1 needs no process group."""

    def __init__(self, Nx, Nv, N, Lv, Lx, nu, dt, homogeneous=False, rank=0, world=1, device=0, dist=None, full_and_linear=False,
                 stream=None, doping=None, linear_landau=False, mass_cons_only=False, gamma=-3, exchange=None):
        """stream: a torch.cuda.Stream all work of this solver (kernels, copies, the NCCL exchange) is ordered on;
        None = torch's current stream for sharded runs, the library's default otherwise.  Two solvers on two streams
        pipeline: the host<->device copies of one overlap the kernels of the other.
        exchange: "peer" (default on NCCL process groups) maps the ranks' buffers into each other (CUDA IPC): the halo
        planes and densities of every SSP-RK3 stage are written straight into the peers' memory by kernels and the whole
        timestep replays as one CUDA graph; "nccl" keeps the all-gather + send/recv per stage through torch.distributed."""
        self.rank, self.world, self.dist = rank, world, dist
        self.stream = stream
        self.homogeneous = bool(homogeneous)
        if homogeneous:
            self.x_begin, self.x_count = 0, 1
        else:
            self.x_begin, self.x_count = shard_range(Nx, world, rank)
        self.g = lpgpu.LPGpu(Nx, Nv, N, Lv, Lx, nu, dt, homogeneous=homogeneous, x_begin=self.x_begin,
                             x_count=self.x_count, device=device, full_and_linear=full_and_linear, doping=doping,
                             linear_landau=linear_landau, mass_cons_only=mass_cons_only, gamma=gamma)
        self.linear_landau = bool(linear_landau)
        self.nu = nu
        self._ex = None
        self.exchange = None
        if stream is not None:
            self.g.set_stream(stream.cuda_stream)
        if world > 1 and not homogeneous:
            import torch
            self.torch = torch
            if stream is None:
                self.g.set_stream(torch.cuda.current_stream().cuda_stream)
            if exchange is None:
                exchange = os.environ.get("LPGPU_EXCHANGE", "peer")
            self.exchange = exchange
            if exchange == "peer":
                # every rank must end up on the same exchange: a rank that cannot map its peers (no CUDA IPC between the
                # processes, GPUs without peer access) sends everybody to the NCCL path
                try:
                    blob, err = self.g.peer_export().tobytes(), None
                except lpgpu.LPGpuError as e:
                    blob, err = b"", str(e)
                blobs = [None] * world
                dist.all_gather_object(blobs, blob)     # also orders every rank's mailbox initialisation before any put
                if all(len(b) == lpgpu.PEER_HANDLE_BYTES for b in blobs):
                    try:
                        self.g.peer_import(rank, world, np.frombuffer(b"".join(blobs), dtype=np.uint8))
                    except lpgpu.LPGpuError as e:
                        err = str(e)
                oks = [None] * world
                dist.all_gather_object(oks, err is None)
                if not all(oks):
                    if rank == 0:
                        print("lpsolver_b200: peer-memory exchange unavailable (%s); using NCCL" % (err or "on another rank"), file=sys.stderr)
                    exchange = self.exchange = "nccl"
            if exchange != "peer":
                self.halo_group = dist.new_group(backend="nccl")   # collective: every rank constructs its solver
                self.halo_stream = torch.cuda.Stream(device=device)
            self._ex = []
            for stage in range(3):
                e = self.g.exchange_info(stage)
                n = e.plane_doubles
                wrap = lambda p, cnt: torch.as_tensor(_DeviceBuffer(p, cnt), device="cuda:%d" % device)
                self._ex.append(dict(ms_local=wrap(e.ms_local, 2 * self.x_count), ms_all=wrap(e.ms_all, 2 * Nx),
                                     send_left=wrap(e.send_left, n), send_right=wrap(e.send_right, n),
                                     recv_left=wrap(e.recv_left, n), recv_right=wrap(e.recv_right, n)))

    def upload(self, U_shard, wait=True):
        self.g.upload_U(U_shard, wait=wait)

    def set_maxwellian(self):
        """LinearLandau: the state now on the device becomes M of Q(f, M) (ComputeDFTofMaxwellian, LP_ompi.cpp:516)."""
        self.g.set_maxwellian()

    def download(self, out=None, wait=True):
        return self.g.download_U(out, wait=wait)

    def synchronize(self):
        self.g.synchronize()

    def advect(self):
        if self.homogeneous:
            return
        if self.world == 1 or self.exchange == "peer":
            self.g.advect_rk3()
            return
        torch = self.torch
        main = self.stream if self.stream is not None else torch.cuda.current_stream()
        for stage in range(3):
            ex = self._ex[stage]
            # the halo planes travel on their own stream and communicator while the density reduction, its all-gather
            # and the field scan run: one NCCL latency per stage on the critical path instead of two
            self.halo_stream.wait_stream(main)                 # this stage's input planes are final
            with torch.cuda.stream(self.halo_stream):
                exchange_halo(self.dist, self.rank, self.world, ex["send_left"], ex["send_right"], ex["recv_left"], ex["recv_right"],
                              group=self.halo_group)
            self.g.advect_reduce(stage)
            with torch.cuda.stream(main):
                exchange_density(self.dist, self.world, ex["ms_local"], ex["ms_all"])
            main.wait_stream(self.halo_stream)
            self.g.advect_apply(stage)

    def step_host(self, U_in, U_out=None):
        """One timestep on this rank's shard held in host memory (upload, RK3, collisions, download pipelined over chunks
        of cells inside the library).  NCCL-exchange runs fall back to the three calls: their advection is driven from here."""
        if self.world == 1 or self.homogeneous or self.exchange == "peer":
            return self.g.step_host(U_in, U_out)
        self.upload(U_in)
        self.step(1)
        return self.download(U_out)

    def step(self, nsteps=1, wait=True):
        """nsteps passes of the while(t<nT) body (LP_ompi.cpp:662-813) without diagnostics.  wait=False only
        enqueues (pipelined callers: synchronize() before touching host buffers)."""
        if self.world == 1 or self.homogeneous or self.exchange == "peer":
            self.g.step(nsteps, wait=wait)
            return
        for _ in range(nsteps):
            self.advect()
            if self.nu > 0:
                self.g.collide_step(wait=False)      # enqueue only: the host runs ahead into the next exchange
        if wait:
            self.g.synchronize()

    def diagnostics_begin(self):
        """Snapshot the state now on the device and start its per-step diagnostics (LP_ompi.cpp:817-849) on a side
        stream; enqueue the next timestep, then call diagnostics_end()."""
        self.g.diagnostics_begin()

    def diagnostics_end(self):
        """(moments[6], entropy, KiE ratio, number of negative cells) of the snapshot, global over the ranks."""
        m5, ms, d4 = self.g.diagnostics_end()
        if self.world > 1 and not self.homogeneous:
            torch = self.torch
            t = torch.from_numpy(np.concatenate([m5, d4])).cuda()
            self.dist.all_reduce(t)
            red = t.cpu().numpy()
            m5, d4 = red[:5], red[5:]
            loc = torch.from_numpy(ms).cuda()
            allms = torch.empty(2 * self.g.Nx, dtype=torch.float64, device=loc.device)
            self.dist.all_gather_into_tensor(allms, loc)
            ms = allms.cpu().numpy()
        ele = 0. if self.homogeneous else self.g.eleE_from_ms(ms)
        return np.concatenate([m5, [ele]]), d4[0], (d4[2] / d4[1] if d4[1] != 0 else 0.), d4[3]

    def moments(self):
        """Global mass, P1..3, KiE, EleE (LP_ompi.cpp:820-827) on every rank."""
        m5, ms = self.g.moments_partial()
        if self.world > 1 and not self.homogeneous:
            torch = self.torch
            t = torch.from_numpy(m5).cuda()
            self.dist.all_reduce(t)
            m5 = t.cpu().numpy()
            loc = torch.from_numpy(ms).cuda()
            allms = torch.empty(2 * self.g.Nx, dtype=torch.float64, device=loc.device)
            self.dist.all_gather_into_tensor(allms, loc)
            ms = allms.cpu().numpy()
        ele = 0. if self.homogeneous else self.g.eleE_from_ms(ms)
        return np.concatenate([m5, [ele]])

    def close(self):
        if self.world > 1 and not self.homogeneous and self.exchange == "peer":
            # every rank learns whether any rank's wait for a peer timed out BEFORE anybody raises: a rank that raised
            # alone would leave the others hanging in the barrier
            try:
                self.g.synchronize()
                bad = 0
            except lpgpu.LPGpuError:
                bad = 1
            t = self.torch.tensor([bad], dtype=self.torch.int32, device="cuda" if self.dist.get_backend() == "nccl" else "cpu")
            self.dist.all_reduce(t)                # also the barrier: nobody unmaps or frees while a peer may still write
            self.g.close()
            if int(t.item()):
                raise lpgpu.LPGpuError("peer exchange: a wait for a peer's flag timed out on %d rank(s); the run's results are poisoned (NaN)" % int(t.item()))
            return
        self.g.close()


def run_from_input_file(path="LPsolver-input.txt", outdir=".", device=0, quiet=False):
    """Single-GPU equivalent of running the reference's `solver` in a directory holding
    LPsolver-input.txt: writes Data/Moments_*.dc with one row per step (row 1 = initial state)."""
    cfg = RunConfig.from_file(path)
    s = ShardedSolver(cfg.Nx, cfg.Nv, cfg.N, cfg.Lv, cfg.Lx, cfg.nu, cfg.dt, homogeneous=cfg.homogeneous, device=device,
                      full_and_linear=cfg.full_and_linear, doping=cfg.doping, linear_landau=cfg.linear_landau,
                      mass_cons_only=cfg.mass_cons_only, gamma=cfg.gamma)
    if cfg.second:
        # LP_ompi.cpp:529-571: pick up from the last record of Data/<Second/Name> (raw doubles, 6*size per record)
        if not cfg.second_name:
            raise ValueError("Please set the name of the file from the previous run under Second/Name")
        need = 6 * (1 if cfg.homogeneous else cfg.Nx) * cfg.Nv ** 3
        path = os.path.join(outdir, "Data", cfg.second_name)
        nbytes = os.path.getsize(path)
        if nbytes < 8 * need or nbytes % (8 * need):
            raise ValueError("Error reading file %s: its size does not match records of 6*size doubles" % path)
        U0 = np.fromfile(path, dtype=np.float64, count=need, offset=nbytes - 8 * need)
    else:
        U0 = cfg.initial_condition()
    s.upload(U0)
    if cfg.linear_landau and cfg.nu > 0:
        s.set_maxwellian()
    out = os.path.join(outdir, cfg.moments_filename())
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        for t in range(cfg.nT + 1):
            m = s.moments()
            fh.write(format_moments_row(m, cfg.homogeneous))
            if not quiet:
                print("step %d: " % t + " ".join("%11.8g" % x for x in m))
            if t < cfg.nT:
                s.step(1)
    s.close()
    return out
