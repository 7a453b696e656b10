"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the two CPU checkers:

* ``PortOracle``  -> oracle/liblp_oracle.so, the plain-C restatement (oracle/lp_oracle.c).
* ``RefOracle``   -> oracle/_ref/libref.so, the UNMODIFIED reference compiled from
  /root/reference/source behind the shims in oracle/shim/ (oracle/ref_driver.cpp only sets the
  reference's globals and forwards calls).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "liblp_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref.so")
REF_SOLVER = os.path.join(_HERE, "_ref", "solver")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(verbose=False):
    """Compile the C restatement and, when /root/reference exists, the reference itself."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_ref():
    return os.path.exists(REF_LIB)


def host_cores():
    """Cores this process may run on (the affinity mask, not the machine's total)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_num_threads(n=None):
    """OpenMP thread count of both checkers (default: every core this process may use).  torchrun exports
    OMP_NUM_THREADS=1, which would silently time a one-thread reference; callers that time the CPU path set it."""
    n = int(n or host_cores())
    os.environ["OMP_NUM_THREADS"] = str(n)
    if os.path.exists(PORT_LIB):
        L = C.CDLL(PORT_LIB)
        if hasattr(L, "lpo_set_num_threads"):
            L.lpo_set_num_threads(n)
    if os.path.exists(REF_LIB):
        L = C.CDLL(REF_LIB)
        if hasattr(L, "ref_set_num_threads"):
            L.ref_set_num_threads(n)
    return n


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PortOracle:
    """The C restatement.  Method names follow the reference's function names."""

    def __init__(self, Nx, Nv, N, Lv, Lx, nu, dt, homogeneous=False, gamma=-3):
        if not os.path.exists(PORT_LIB):
            build()
        L = C.CDLL(PORT_LIB)
        L.lpo_create.restype = C.c_void_p
        L.lpo_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.lpo_gHat3.restype = C.c_double
        L.lpo_gHat3.argtypes = [C.c_void_p] + [C.c_double] * 6
        L.lpo_num_threads.restype = C.c_int
        self.L = L
        self.Nx, self.Nv, self.N = Nx, Nv, N
        self.Lv, self.Lx, self.nu, self.dt = Lv, Lx, nu, dt
        self.homogeneous = bool(homogeneous)
        self.ncell = 1 if homogeneous else Nx
        self.N3, self.sv = N ** 3, Nv ** 3
        self.h = C.c_void_p(L.lpo_create(Nx, Nv, N, Lv, Lx, nu, dt, int(bool(homogeneous)), gamma))
        if not self.h:
            raise ValueError("lpo_create failed (only gamma=-3 is supported)")

    def close(self):
        if self.h:
            self.L.lpo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        fn = getattr(self.L, name)
        conv = []
        for a in args:
            if isinstance(a, np.ndarray):
                conv.append(a.ctypes.data_as(C.c_void_p))
            elif isinstance(a, float):
                conv.append(C.c_double(a))
            elif a is None:
                conv.append(C.c_void_p(None))
            else:
                conv.append(C.c_int(a))
        fn.restype = None
        fn(self.h, *conv)

    num_threads = property(lambda self: self.L.lpo_num_threads())

    def set_direct_intmodes(self, on):
        self._call("lpo_set_direct_intmodes", int(on))

    def set_fandl(self, on=True):
        """FullandLinear = True (reference test 3): collide_step / step use the *_FandL routines."""
        self._call("lpo_set_fandl", int(on))

    def set_doping(self, NL, NH, eps, T_L=0.4, T_R=0.4):
        """Doping = True (reference test 1): non-uniform background, Dirichlet walls."""
        self._call("lpo_set_doping", float(NL), float(NH), float(eps), float(T_L), float(T_R))

    def SetInit_ND(self):
        U = np.zeros(self.ncell * self.sv * 6)
        self._call("lpo_SetInit_ND", U)
        return U

    def set_linear_landau(self, U):
        """LinearLandau = True: Q(f, M) with M the state U (ComputeDFTofMaxwellian); None switches back."""
        self._call("lpo_set_linear_landau", None if U is None else _f64(U))

    def set_mass_cons_only(self, on=True):
        self._call("lpo_set_mass_cons_only", int(on))

    def ComputeQLinear(self, f, mhat):
        q = np.empty((self.N3, 2))
        self._call("lpo_ComputeQLinear", _f64(f), _f64(mhat), q)
        return q

    def ComputeQ_FandL(self, f):
        q, ql = np.empty((self.N3, 2)), np.empty((self.N3, 2))
        self._call("lpo_ComputeQ_FandL", _f64(f), q, ql)
        return q, ql

    def conserveMoments_FandL(self, q, ql):
        q, ql = _f64(q).copy(), _f64(ql).copy()
        self._call("lpo_conserveMoments_FandL", q, ql)
        return q, ql

    def grids(self):
        v, e, w = (np.empty(self.N) for _ in range(3))
        self._call("lpo_get_grids", v, e, w)
        return v, e, w

    def gHat3(self, z, k):
        return self.L.lpo_gHat3(self.h, *[float(x) for x in z], *[float(x) for x in k])

    def weight_row(self, xi):
        row = np.empty(self.N3)
        self._call("lpo_weight_row", int(xi), row)
        return row

    def conservation(self):
        C5, CCt = np.empty((5, self.N3)), np.empty(25)
        self._call("lpo_get_conservation", C5, CCt)
        return C5, CCt.reshape(5, 5)

    def IntModes(self, k1, k2, k3, j1, j2, j3):
        out = np.empty(10)
        self._call("lpo_IntModes", k1, k2, k3, j1, j2, j3, out)
        return out

    def setInit_spectral(self, U):
        f = np.empty((self.ncell, self.N3))
        self._call("lpo_setInit_spectral", _f64(U), f)
        return f

    def fft3D(self, x):
        x = _f64(x); out = np.empty_like(x)
        self._call("lpo_fft3D", x, out)
        return out

    def FS(self, x):
        x = _f64(x); out = np.empty_like(x)
        self._call("lpo_FS", x, out)
        return out

    def ComputeQ(self, f):
        q = np.empty((self.N3, 2))
        self._call("lpo_ComputeQ", _f64(f), q)
        return q

    def conserveMoments(self, q):
        q = _f64(q).copy()
        self._call("lpo_conserveMoments", q)
        return q

    def RK4(self, f, cell, qHat, U):
        dU = np.empty((self.sv, 5)); q123 = np.empty((3, self.N3, 2))
        self._call("lpo_RK4", _f64(f), int(cell), _f64(qHat), _f64(U), dU, q123)
        return dU, q123

    def collide_step(self, U):
        U = _f64(U).copy()
        self._call("lpo_collide_step", U)
        return U

    def field(self, U):
        out = np.empty(1 + 4 * self.Nx)
        self._call("lpo_field", _f64(U), out)
        return out

    def RK3(self, U):
        U = _f64(U).copy()
        self._call("lpo_RK3", U)
        return U

    def step(self, U):
        U = _f64(U).copy()
        self._call("lpo_step", U)
        return U

    def SetInit_LD(self, A_amp, k_wave, twostream=False):
        U = np.zeros(self.ncell * self.sv * 6)
        self._call("lpo_SetInit_LD", U, float(A_amp), float(k_wave), int(twostream))
        return U

    def SetInit_4H(self):
        U = np.zeros(self.ncell * self.sv * 6)
        self._call("lpo_SetInit_4H", U)
        return U

    def SetInit_4H_Homo(self):
        U = np.zeros(self.sv * 6)
        self._call("lpo_SetInit_4H_Homo", U)
        return U

    def moments(self, U):
        out = np.empty(6)
        self._call("lpo_moments", _f64(U), out)
        return out

    def diagnostics(self, U):
        """entropy, KiE ratio (negative / positive cells), number of negative cells"""
        out = np.empty(4)
        self._call("lpo_diagnostics", _f64(U), out)
        return np.array([out[0], out[2] / out[1], out[3]])


class RefOracle:
    """The unmodified reference (global state: one live configuration per process)."""

    def __init__(self, Nx, Nv, N, Lv, Lx, nu, dt, homogeneous=False, gamma=-3, build_weights=True):
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(REF_LIB + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(REF_LIB)
        L.ref_setup.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_gHat3.restype = C.c_double
        L.ref_gHat3.argtypes = [C.c_double] * 6 + [C.c_int]
        L.ref_entropy.restype = C.c_double
        self.L = L
        self.Nx, self.Nv, self.N = Nx, Nv, N
        self.homogeneous = bool(homogeneous)
        self.ncell = 1 if homogeneous else Nx
        self.N3, self.sv = N ** 3, Nv ** 3
        self.gamma = gamma
        L.ref_setup(Nx, Nv, N, Lv, Lx, nu, dt, int(bool(homogeneous)), gamma, int(bool(build_weights)))

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    num_threads = property(lambda self: self.L.ref_num_threads())

    def grids(self):
        v, e, w = (np.empty(self.N) for _ in range(3))
        self.L.ref_get_grids(self._p(v), self._p(e), self._p(w))
        return v, e, w

    def gHat3(self, z, k):
        return self.L.ref_gHat3(*[float(x) for x in z], *[float(x) for x in k], self.gamma)

    def weight_row(self, xi):
        row = np.empty(self.N3)
        assert self.L.ref_get_weight_row(int(xi), self._p(row)) == 0
        return row

    def conservation(self):
        C5, CCt = np.empty((5, self.N3)), np.empty(25)
        self.L.ref_get_conservation(self._p(C5), self._p(CCt))
        return C5, CCt.reshape(5, 5)

    def IntModes(self, k1, k2, k3, j1, j2, j3):
        out = np.empty(10)
        self.L.ref_IntModes(k1, k2, k3, j1, j2, j3, self._p(out))
        return out

    def setInit_spectral(self, U):
        U = _f64(U); f = np.empty((self.ncell, self.N3))
        self.L.ref_setInit_spectral(self._p(U), self._p(f))
        return f

    def fft3D(self, x):
        x = _f64(x); out = np.empty_like(x)
        self.L.ref_fft3D(self._p(x), self._p(out))
        return out

    def FS(self, x):
        x = _f64(x); out = np.empty_like(x)
        self.L.ref_FS(self._p(x), self._p(out))
        return out

    def ComputeQ(self, f):
        f = _f64(f); q = np.empty((self.N3, 2))
        assert self.L.ref_ComputeQ(self._p(f), self._p(q)) == 0
        return q

    def conserveMoments(self, q):
        q = _f64(q).copy()
        self.L.ref_conserveMoments(self._p(q))
        return q

    def collide_step(self, U):
        U = _f64(U).copy()
        assert self.L.ref_collide_step(self._p(U)) == 0
        return U

    def stage_spectra(self):
        q = np.empty((3, self.N3, 2))
        self.L.ref_get_stage_spectra(self._p(q[0]), self._p(q[1]), self._p(q[2]))
        return q

    def field(self, U):
        U = _f64(U); out = np.empty(1 + 4 * self.Nx)
        self.L.ref_field(self._p(U), self._p(out))
        return out

    def RK3(self, U):
        U = _f64(U).copy()
        self.L.ref_RK3(self._p(U))
        return U

    def step(self, U):
        U = U if self.homogeneous else self.RK3(U)
        return self.collide_step(U)

    def SetInit_LD(self, A_amp, k_wave, twostream=False):
        U = np.zeros(self.ncell * self.sv * 6)
        self.L.ref_SetInit_LD(self._p(U), C.c_double(A_amp), C.c_double(k_wave), int(twostream))
        return U

    def SetInit_4H(self):
        U = np.zeros(self.ncell * self.sv * 6)
        self.L.ref_SetInit_4H(self._p(U))
        return U

    def SetInit_4H_Homo(self):
        U = np.zeros(self.sv * 6)
        self.L.ref_SetInit_4H_Homo(self._p(U))
        return U

    def moments(self, U):
        U = _f64(U); out = np.empty(6)
        self.L.ref_moments(self._p(U), self._p(out))
        return out

    def diagnostics(self, U):
        U = _f64(U); out = np.empty(4)
        self.L.ref_diagnostics(self._p(U), self._p(out))
        return out[:3].copy()
