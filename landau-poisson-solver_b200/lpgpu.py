"""ctypes binding of include/lpgpu.h.  No torch types cross this boundary: host arrays are numpy
(float64, C-contiguous), device buffers are raw addresses.  There is no CPU fallback: if
liblpgpu.so is missing or no CUDA device is usable, construction raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    return os.path.join(_HERE, "liblpgpu.so")


class LPGpuError(RuntimeError):
    pass


class Params(C.Structure):
    """struct lpgpu_params (include/lpgpu.h)."""
    _fields_ = [("Nx", C.c_int), ("Nv", C.c_int), ("N", C.c_int),
                ("Lv", C.c_double), ("Lx", C.c_double), ("nu", C.c_double), ("dt", C.c_double),
                ("gamma", C.c_int), ("homogeneous", C.c_int),
                ("x_begin", C.c_int), ("x_count", C.c_int), ("device", C.c_int),
                ("computeq_variant", C.c_int), ("full_and_linear", C.c_int),
                ("doping", C.c_int), ("NL", C.c_double), ("NH", C.c_double), ("eps", C.c_double),
                ("T_L", C.c_double), ("T_R", C.c_double), ("linear_landau", C.c_int), ("mass_cons_only", C.c_int)]


class Exchange(C.Structure):
    """struct lpgpu_exchange (include/lpgpu.h)."""
    _fields_ = [("ms_local", C.c_void_p), ("ms_all", C.c_void_p),
                ("send_left", C.c_void_p), ("send_right", C.c_void_p),
                ("recv_left", C.c_void_p), ("recv_right", C.c_void_p),
                ("plane_doubles", C.c_longlong)]


# every symbol include/lpgpu.h declares (tests check the library exports exactly these)
EXPORTS = [
    "lpgpu_last_error", "lpgpu_device_count", "lpgpu_init", "lpgpu_finalize", "lpgpu_set_stream",
    "lpgpu_synchronize", "lpgpu_launch_count", "lpgpu_upload_U", "lpgpu_download_U",
    "lpgpu_upload_U_async", "lpgpu_download_U_async", "lpgpu_step_async", "lpgpu_step_host", "lpgpu_set_maxwellian",
    "lpgpu_advect_rk3", "lpgpu_collide_step", "lpgpu_collide_step_async", "lpgpu_step", "lpgpu_advect_exchange_info",
    "lpgpu_advect_reduce", "lpgpu_advect_apply", "lpgpu_setInit_spectral", "lpgpu_fft3D", "lpgpu_FS",
    "lpgpu_ComputeQ", "lpgpu_conserveMoments", "lpgpu_sample_device", "lpgpu_eval_device",
    "lpgpu_get_stage_spectrum", "lpgpu_field", "lpgpu_moments_partial", "lpgpu_eleE_from_ms",
    "lpgpu_profile_computeQ", "lpgpu_profile_read", "lpgpu_fp64_peak", "lpgpu_diagnostics_partial", "lpgpu_marginal_sums",
    "lpgpu_diagnostics_begin", "lpgpu_diagnostics_end",
    "lpgpu_peer_export", "lpgpu_peer_import", "lpgpu_peer_status", "lpgpu_peer_set_timeout",
]
PEER_HANDLE_BYTES = 256

_lib = None


def load_library():
    """dlopen liblpgpu.so (needs libcudart, not a GPU) and declare signatures."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LPGpuError("liblpgpu.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "or `make -C landau-poisson-solver_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    L.lpgpu_last_error.restype = C.c_char_p
    L.lpgpu_launch_count.restype = C.c_longlong
    L.lpgpu_launch_count.argtypes = [C.c_void_p]
    L.lpgpu_init.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    for name in ("lpgpu_finalize", "lpgpu_synchronize", "lpgpu_advect_rk3", "lpgpu_collide_step", "lpgpu_collide_step_async", "lpgpu_sample_device", "lpgpu_set_maxwellian"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.lpgpu_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    for name in ("lpgpu_upload_U", "lpgpu_download_U", "lpgpu_upload_U_async", "lpgpu_download_U_async", "lpgpu_setInit_spectral", "lpgpu_field", "lpgpu_diagnostics_partial", "lpgpu_marginal_sums"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
    for name in ("lpgpu_step", "lpgpu_step_async", "lpgpu_advect_reduce", "lpgpu_advect_apply", "lpgpu_eval_device"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
    L.lpgpu_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpgpu_advect_exchange_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(Exchange)]
    for name in ("lpgpu_fft3D", "lpgpu_FS", "lpgpu_ComputeQ"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.lpgpu_conserveMoments.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.lpgpu_get_stage_spectrum.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.lpgpu_moments_partial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpgpu_diagnostics_begin.argtypes = [C.c_void_p]
    L.lpgpu_peer_export.argtypes = [C.c_void_p, C.c_void_p]
    L.lpgpu_peer_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.lpgpu_peer_status.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    L.lpgpu_peer_set_timeout.argtypes = [C.c_void_p, C.c_double]
    L.lpgpu_diagnostics_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lpgpu_profile_computeQ.argtypes = [C.c_void_p, C.c_int]
    L.lpgpu_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.lpgpu_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.lpgpu_eleE_from_ms.argtypes = [C.POINTER(Params), C.c_void_p, C.POINTER(C.c_double)]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def fp64_peak_tflops(device=0):
    """DFMA micro-benchmark on `device` (TFLOP/s)."""
    L = load_library()
    v = C.c_double()
    rc = L.lpgpu_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise LPGpuError("lpgpu error %d: %s" % (rc, (L.lpgpu_last_error() or b"").decode()))
    return v.value


class LPGpu:
    """One context = one GPU's shard of spatial cells.  Method names follow the reference's
    function names (RK3 -> advect_rk3, ComputeQ, conserveMoments, FS, fft3D, setInit_spectral)."""

    def __init__(self, Nx, Nv, N, Lv, Lx, nu, dt, homogeneous=False, gamma=-3, x_begin=0, x_count=None,
                 device=0, computeq_variant=0, full_and_linear=False, doping=None, linear_landau=False, mass_cons_only=False):
        """doping: None or a dict(NL=, NH=, eps=, T_L=, T_R=) -- the [Doping] section of the input deck."""
        self.L = load_library()
        if x_count is None:
            x_count = 1 if homogeneous else Nx
        dp = doping or {}
        self.params = Params(Nx, Nv, N, Lv, Lx, nu, dt, gamma, int(bool(homogeneous)), x_begin, x_count, device,
                             computeq_variant, int(bool(full_and_linear)), int(doping is not None),
                             float(dp.get("NL", 0.)), float(dp.get("NH", 0.)), float(dp.get("eps", 1.)),
                             float(dp.get("T_L", 0.4)), float(dp.get("T_R", 0.4)), int(bool(linear_landau)), int(bool(mass_cons_only)))
        self.Nx, self.Nv, self.N = Nx, Nv, N
        self.homogeneous = bool(homogeneous)
        self.ncell = 1 if homogeneous else x_count
        self.N3, self.sv = N ** 3, Nv ** 3
        self.h = C.c_void_p()
        self._check(self.L.lpgpu_init(C.byref(self.params), C.byref(self.h)))

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise LPGpuError("lpgpu error %d: %s" % (rc, (self.L.lpgpu_last_error() or b"").decode()))

    def close(self):
        if getattr(self, "h", None):
            self.L.lpgpu_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        self._check(self.L.lpgpu_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        self._check(self.L.lpgpu_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.L.lpgpu_launch_count(self.h))

    # -- state
    def upload_U(self, U, wait=True):
        """wait=False: enqueue only; U must then be page-locked and stay alive until synchronize()."""
        U = _f64(U)
        assert U.size == self.ncell * self.sv * 6, "U must hold this shard: x_count*Nv^3*6 doubles"
        self._check((self.L.lpgpu_upload_U if wait else self.L.lpgpu_upload_U_async)(self.h, _ptr(U)))

    def download_U(self, out=None, wait=True):
        U = np.empty(self.ncell * self.sv * 6) if out is None else out
        assert wait or out is not None, "an enqueue-only download needs the caller's page-locked buffer"
        self._check((self.L.lpgpu_download_U if wait else self.L.lpgpu_download_U_async)(self.h, _ptr(U)))
        return U

    def set_maxwellian(self):
        """ComputeDFTofMaxwellian: the state now on the device becomes M of the linear operator Q(f, M)."""
        self._check(self.L.lpgpu_set_maxwellian(self.h))

    # -- phases
    def advect_rk3(self):
        self._check(self.L.lpgpu_advect_rk3(self.h))

    def collide_step(self, wait=True):
        self._check(self.L.lpgpu_collide_step(self.h) if wait else self.L.lpgpu_collide_step_async(self.h))

    def step(self, nsteps=1, wait=True):
        self._check((self.L.lpgpu_step if wait else self.L.lpgpu_step_async)(self.h, int(nsteps)))

    def step_host(self, U_in, U_out=None):
        """One timestep on a host-resident state (the reference's time loop keeps U on the host): upload, RK3, collision
        step and download pipelined over chunks of cells.  U_out defaults to a new array; it may be U_in.  Page-locked
        buffers make the copies asynchronous."""
        assert isinstance(U_in, np.ndarray) and U_in.dtype == np.float64 and U_in.flags.c_contiguous, "U_in: contiguous float64 array"
        assert U_in.size == self.ncell * self.sv * 6, "U must hold this shard: x_count*Nv^3*6 doubles"
        if U_out is None:
            U_out = np.empty(self.ncell * self.sv * 6)
        assert isinstance(U_out, np.ndarray) and U_out.dtype == np.float64 and U_out.flags.c_contiguous and U_out.size == U_in.size
        self._check(self.L.lpgpu_step_host(self.h, _ptr(U_in), _ptr(U_out)))
        return U_out

    def exchange_info(self, stage):
        ex = Exchange()
        self._check(self.L.lpgpu_advect_exchange_info(self.h, int(stage), C.byref(ex)))
        return ex

    def advect_reduce(self, stage):
        self._check(self.L.lpgpu_advect_reduce(self.h, int(stage)))

    def advect_apply(self, stage):
        self._check(self.L.lpgpu_advect_apply(self.h, int(stage)))

    # -- fine-grained (host buffers)
    def setInit_spectral(self):
        f = np.empty((self.ncell, self.N3))
        self._check(self.L.lpgpu_setInit_spectral(self.h, _ptr(f)))
        return f

    def _batched(self, fn, x, in_w, out_w):
        x = _f64(x)
        B = x.size // (self.N3 * in_w)
        assert B * self.N3 * in_w == x.size and B >= 1
        out = np.empty((B, self.N3, out_w))
        self._check(fn(self.h, _ptr(x), _ptr(out), B))
        return out

    def fft3D(self, x):
        return self._batched(self.L.lpgpu_fft3D, x, 2, 2)

    def FS(self, x):
        return self._batched(self.L.lpgpu_FS, x, 2, 2)

    def ComputeQ(self, f):
        return self._batched(self.L.lpgpu_ComputeQ, f, 1, 2)

    def conserveMoments(self, q):
        q = _f64(q).copy()
        B = q.size // (2 * self.N3)
        self._check(self.L.lpgpu_conserveMoments(self.h, _ptr(q), B))
        return q.reshape(B, self.N3, 2)

    def sample_device(self):
        self._check(self.L.lpgpu_sample_device(self.h))

    def eval_device(self, B=None):
        self._check(self.L.lpgpu_eval_device(self.h, int(self.ncell if B is None else B)))

    def stage_spectrum(self, which):
        out = np.empty((self.ncell, self.N3, 2))
        self._check(self.L.lpgpu_get_stage_spectrum(self.h, int(which), _ptr(out)))
        return out

    def field(self):
        out = np.empty(1 + 4 * self.Nx)
        self._check(self.L.lpgpu_field(self.h, _ptr(out)))
        return out

    # -- measurement helpers
    def profile_computeQ(self, enable=True):
        self._check(self.L.lpgpu_profile_computeQ(self.h, int(enable)))   # 0 off, 1 whole ComputeQ chain, 2 its dominant kernel

    def profile_read(self):
        ms, n = C.c_double(), C.c_longlong()
        self._check(self.L.lpgpu_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- diagnostics
    def moments_partial(self):
        out = np.empty(5)
        ms = np.zeros(2 * self.ncell)
        self._check(self.L.lpgpu_moments_partial(self.h, _ptr(out), _ptr(ms)))
        return out, ms

    def marginal_sums(self):
        """[rows, 4] sums PrintMarginal needs: rows = (x cell, j1) or, homogeneous, (j1, j2)."""
        rows = self.Nv * self.Nv if self.homogeneous else self.ncell * self.Nv
        out = np.empty((rows, 4))
        self._check(self.L.lpgpu_marginal_sums(self.h, _ptr(out)))
        return out

    def eleE_from_ms(self, ms_all):
        ms_all = _f64(ms_all)
        assert ms_all.size == 2 * self.Nx
        v = C.c_double()
        self._check(self.L.lpgpu_eleE_from_ms(C.byref(self.params), _ptr(ms_all), C.byref(v)))
        return v.value

    def diagnostics_partial(self):
        """entropy, KiE sum over non-negative cells, over negative cells, number of negative cells (this shard)."""
        out = np.empty(4)
        self._check(self.L.lpgpu_diagnostics_partial(self.h, _ptr(out)))
        return out

    # -- peer-memory exchange of a sharded run (CUDA IPC)
    def peer_export(self):
        blob = np.zeros(PEER_HANDLE_BYTES, dtype=np.uint8)
        self._check(self.L.lpgpu_peer_export(self.h, _ptr(blob)))
        return blob

    def peer_import(self, rank, world, blobs):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        assert blobs.size == world * PEER_HANDLE_BYTES
        self._check(self.L.lpgpu_peer_import(self.h, int(rank), int(world), _ptr(blobs)))

    def peer_set_timeout(self, seconds):
        """Bound of a wait for a peer's flag (default 60 s); call before the first timestep."""
        self._check(self.L.lpgpu_peer_set_timeout(self.h, float(seconds)))

    def peer_status(self):
        """Raises if a bounded wait for a peer's flag ever timed out."""
        n = C.c_longlong()
        self._check(self.L.lpgpu_peer_status(self.h, C.byref(n)))
        return n.value

    def diagnostics_begin(self):
        """Snapshot the state and enqueue its moment / density / entropy / negativity reductions on a side stream; the
        caller enqueues the next timestep, then collects with diagnostics_end()."""
        self._check(self.L.lpgpu_diagnostics_begin(self.h))

    def diagnostics_end(self):
        """(m5, ms_local, d4) of the snapshot: what moments_partial() and diagnostics_partial() return for that state."""
        m5, ms, d4 = np.empty(5), np.zeros(2 * self.ncell), np.empty(4)
        self._check(self.L.lpgpu_diagnostics_end(self.h, _ptr(m5), _ptr(ms), _ptr(d4)))
        return m5, ms, d4

    def moments(self):
        """mass, P1, P2, P3, KiE, EleE for a single-shard context."""
        m5, ms = self.moments_partial()
        ele = 0.0 if self.homogeneous else self.eleE_from_ms(ms)
        return np.concatenate([m5, [ele]])
