"""ctypes binding of libgsmvi_b200.so (the C ABI in include/gsmvi_b200.h).

The library is the product path: there is no CPU or PyTorch fallback. If the shared object is missing
or fails to load, importing any operator raises immediately."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsmvi_b200.so")

_lib = None


class GsmviError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GsmviError(
                "libgsmvi_b200.so not found at %s: build it with `python gsm-vi_b200/build.py` "
                "(no CPU fallback exists)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


c_f = ctypes.c_float
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_p = ctypes.c_void_p


def _declare(L):
    L.gsmvi_abi_version.restype = c_i
    L.gsmvi_gemm_tf32.restype = c_i
    L.gsmvi_gemm_tf32.argtypes = [c_p, c_ll, c_ll, c_ll, c_i, c_p, c_ll, c_ll, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i,
                                  c_f, c_f, c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]


def check(rc, what):
    if rc != 0:
        raise GsmviError("%s failed with status %d%s" % (what, rc, " (CUDA error)" if rc > 0 else ""))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


KR_FULL, KR_A_LOWER, KR_B_LOWER, KR_A_UPPER, KR_B_UPPER = 0, 1, 2, 4, 8


def gemm_tf32(A, B, C, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, Cin=None, bias_n=None, npass=3,
              tri=False, mirror=False, krange=0, neg_from=0x7fffffff, A_lo=None, B_lo=None):
    """C[M,N] = alpha * op(A) op(B)^T + beta*Cin + bias_n. A, B, C are 2-D fp32 CUDA tensors (row-major views with
    stride(1) == 1). K-major operand: [rows, K]; MN-major (a_mn/b_mn): [K, rows]."""
    assert A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    rc = lib().gsmvi_gemm_tf32(ptr(A), A.shape[0], A.shape[1], A.stride(0), int(a_mn), ptr(B), B.shape[0], B.shape[1],
                               B.stride(0), int(b_mn), ptr(C), C.stride(0), M, N, K, alpha, beta, ptr(Cin),
                               Cin.stride(0) if Cin is not None else 0, ptr(bias_n), npass, int(tri), int(mirror),
                               krange, neg_from, ptr(A_lo), ptr(B_lo), stream_ptr())
    check(rc, "gsmvi_gemm_tf32")
    return C


# ------------------------------------------------------------------------------------------------ GSM path
c_ull = ctypes.c_ulonglong
WS_POTRF, WS_GSM_UPDATE = 1, 2


def _declare_gsm(L):
    L.gsmvi_workspace_bytes.restype = c_ll
    L.gsmvi_workspace_bytes.argtypes = [c_i, c_i, c_i]
    L.gsmvi_potrf_check.restype = c_i
    L.gsmvi_potrf_check.argtypes = [c_p, c_ll, c_p, c_ll, c_i, c_p, c_p, c_i, c_p]
    L.gsmvi_philox_normal.restype = c_i
    L.gsmvi_philox_normal.argtypes = [c_p, c_ll, c_i, c_i, c_ull, c_ull, c_p]
    L.gsmvi_sample.restype = c_i
    L.gsmvi_sample.argtypes = [c_p, c_p, c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_i, c_p]
    L.gsmvi_tf32_split.restype = c_i
    L.gsmvi_tf32_split.argtypes = [c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_p]
    L.gsmvi_gauss_score.restype = c_i
    L.gsmvi_gauss_score.argtypes = [c_p, c_ll, c_p, c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_i, c_p]
    L.gsmvi_gsm_update.restype = c_i
    L.gsmvi_gsm_update.argtypes = [c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_p,
                                   c_i, c_p]
    L.gsmvi_gsm_apply_stats.restype = c_i
    L.gsmvi_gsm_apply_stats.argtypes = [c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_ll, c_p, c_i, c_p]


_declare_base = _declare


def _declare(L):  # noqa: F811  (extends the base declarations)
    _declare_base(L)
    _declare_gsm(L)


def workspace_bytes(kind, B, D):
    n = lib().gsmvi_workspace_bytes(kind, B, D)
    if n < 0:
        raise GsmviError("unknown workspace kind %d" % kind)
    return n


def potrf_check(Sigma, L_out, D, bad_flag, ws, npass=3):
    """L_out <- chol(Sigma[:D,:D]); bad_flag (int32[1] device tensor) <- 0 if PD else 1.  No sync."""
    check(lib().gsmvi_potrf_check(ptr(Sigma), Sigma.stride(0), ptr(L_out), L_out.stride(0), D, ptr(bad_flag), ptr(ws),
                                  npass, stream_ptr()), "gsmvi_potrf_check")


def philox_normal(Z, B, D, seed, offset):
    check(lib().gsmvi_philox_normal(ptr(Z), Z.stride(0), B, D, seed & (2**64 - 1), offset & (2**64 - 1), stream_ptr()),
          "gsmvi_philox_normal")


def tf32_split(A, A_hi, A_lo, rows, cols):
    """A_hi <- tf32_rn(A), A_lo <- tf32_rn(A - A_hi): pre-split form of a reused GEMM operand (pass hi as the operand)."""
    assert A_hi.stride(0) == A_lo.stride(0)
    check(lib().gsmvi_tf32_split(ptr(A), A.stride(0), ptr(A_hi), ptr(A_lo), A_lo.stride(0), rows, cols, stream_ptr()),
          "gsmvi_tf32_split")


def sample(mu, L_, Z, X, B, D, npass=3, L_lo=None):
    check(lib().gsmvi_sample(ptr(mu), ptr(L_), ptr(L_lo), L_.stride(0), ptr(Z), Z.stride(0), ptr(X), X.stride(0), B, D,
                             npass, stream_ptr()), "gsmvi_sample")


def gauss_score(X, P, c, G, B, D, npass=3, P_lo=None):
    check(lib().gsmvi_gauss_score(ptr(X), X.stride(0), ptr(P), ptr(P_lo), P.stride(0), ptr(c), ptr(G), G.stride(0), B, D,
                                  npass, stream_ptr()), "gsmvi_gauss_score")


def gsm_update_raw(X, G, mu, Sigma, mu_out, Sigma_out, B, D, B_total, mode, ws, npass=3, Sigma_hi=None, Sigma_lo=None):
    check(lib().gsmvi_gsm_update(ptr(X), X.stride(0), ptr(G), G.stride(0), ptr(mu), ptr(Sigma), ptr(Sigma_hi), ptr(Sigma_lo),
                                 Sigma.stride(0), ptr(mu_out), ptr(Sigma_out), Sigma_out.stride(0), B, D, B_total, mode,
                                 ptr(ws), npass, stream_ptr()), "gsmvi_gsm_update")


def gsm_apply_stats(Sigma, dSigma, mu, dmu, Sigma_out, mu_out, D):
    check(lib().gsmvi_gsm_apply_stats(ptr(Sigma), Sigma.stride(0), ptr(dSigma), dSigma.stride(0), ptr(mu), ptr(dmu),
                                      ptr(Sigma_out), Sigma_out.stride(0), ptr(mu_out), D, stream_ptr()),
          "gsmvi_gsm_apply_stats")


# ------------------------------------------------------------------------------------------------ BaM path
c_d = ctypes.c_double
WS_BAM_STATS, WS_BAM_SOLVE, WS_BAM_SOLVE_LOWRANK = 3, 4, 5


def _declare_bam(L):
    L.gsmvi_dgemm.restype = c_i
    L.gsmvi_dgemm.argtypes = [c_p, c_ll, c_i, c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_d, c_d, c_p, c_ll, c_d, c_i,
                              c_i, c_i, c_p]
    L.gsmvi_bam_stats.restype = c_i
    L.gsmvi_bam_stats.argtypes = [c_p, c_ll, c_p, c_ll, c_i, c_i, c_i, c_p, c_i, c_i, c_p]
    L.gsmvi_bam_solve.restype = c_i
    L.gsmvi_bam_solve.argtypes = [c_p, c_i, c_i, c_i, c_p, c_p, c_ll, c_d, c_d, c_p, c_p, c_ll, c_p, c_i,
                                  ctypes.POINTER(c_i), c_p, c_i, c_i, c_p]
    L.gsmvi_bam_solve_lowrank.restype = c_i
    L.gsmvi_bam_solve_lowrank.argtypes = [c_p, c_i, c_i, c_i, c_p, c_p, c_ll, c_d, c_d, c_p, c_p, c_ll, c_p, c_i,
                                          ctypes.POINTER(c_i), c_p, c_p]


_declare_gsm_level = _declare


def _declare(L):  # noqa: F811
    _declare_gsm_level(L)
    _declare_bam(L)


def dgemm(A, B, C, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, Cin=None, diag_add=0.0, tri=False,
          mirror=False, krange=0):
    check(lib().gsmvi_dgemm(ptr(A), A.stride(0), int(a_mn), ptr(B), B.stride(0), int(b_mn), ptr(C), C.stride(0), M, N, K,
                            alpha, beta, ptr(Cin), Cin.stride(0) if Cin is not None else 0, diag_add, int(tri),
                            int(mirror), krange, stream_ptr()), "gsmvi_dgemm")
    return C


def bam_stats(X, G, B, D, B_total, stats_ws, stage, npass=3):
    check(lib().gsmvi_bam_stats(ptr(X), X.stride(0), ptr(G), G.stride(0), B, D, B_total, ptr(stats_ws), npass, stage,
                                stream_ptr()), "gsmvi_bam_stats")


def bam_solve_sharded(stats_ws, B, D, B_total, mu0, Sigma0, reg, jitter, mu_out, Sigma_out, solve_ws, bad_flag, shard,
                      max_ns=200):
    """Tensor-parallel phase 2 of the solve (gsmvi_bam_solve_sharded); `shard` is a _comm.BamShardC.  Returns the number
    of Newton-Schulz iterations (identical on every rank).  Synchronises the current stream."""
    f = lib().gsmvi_bam_solve_sharded
    if not getattr(f, "_declared", False):
        f.restype = c_i
        f.argtypes = [c_p, c_i, c_i, c_i, c_p, c_p, c_ll, c_d, c_d, c_p, c_p, c_ll, c_p, c_i, ctypes.POINTER(c_i), c_p, c_p,
                      c_p]
        f._declared = True
    it = c_i(0)
    check(f(ptr(stats_ws), B, D, B_total, ptr(mu0), ptr(Sigma0), Sigma0.stride(0), float(reg), float(jitter), ptr(mu_out),
            ptr(Sigma_out), Sigma_out.stride(0), ptr(solve_ws), max_ns, ctypes.byref(it), ptr(bad_flag), ctypes.byref(shard),
            stream_ptr()), "gsmvi_bam_solve_sharded")
    return it.value


def bam_solve(stats_ws, B, D, B_total, mu0, Sigma0, reg, jitter, mu_out, Sigma_out, solve_ws, bad_flag, lowrank=False,
              max_ns=200, world=1, phase=0):
    """Returns the number of Newton-Schulz iterations run.  Synchronises the current stream."""
    it = c_i(0)
    if lowrank:
        rc = lib().gsmvi_bam_solve_lowrank(ptr(stats_ws), B, D, B_total, ptr(mu0), ptr(Sigma0), Sigma0.stride(0),
                                           float(reg), float(jitter), ptr(mu_out), ptr(Sigma_out), Sigma_out.stride(0),
                                           ptr(solve_ws), max_ns, ctypes.byref(it), ptr(bad_flag), stream_ptr())
    else:
        rc = lib().gsmvi_bam_solve(ptr(stats_ws), B, D, B_total, ptr(mu0), ptr(Sigma0), Sigma0.stride(0), float(reg),
                                   float(jitter), ptr(mu_out), ptr(Sigma_out), Sigma_out.stride(0), ptr(solve_ws),
                                   max_ns, ctypes.byref(it), ptr(bad_flag), world, phase, stream_ptr())
    check(rc, "gsmvi_bam_solve_lowrank" if lowrank else "gsmvi_bam_solve")
    return it.value


# ------------------------------------------------------------------------------------------------ monitor
def _declare_mon(L):
    L.gsmvi_gauss_logq_reduce.restype = c_i
    L.gsmvi_gauss_logq_reduce.argtypes = [c_p, c_ll, c_i, c_i, c_p, c_p, c_ll, c_i, c_p, c_p]


_declare_bam_level = _declare


def _declare(L):  # noqa: F811
    _declare_bam_level(L)
    _declare_mon(L)


def gauss_logq_reduce(Z_or_X, N, D, mu, L_, out, from_z):
    check(lib().gsmvi_gauss_logq_reduce(ptr(Z_or_X), Z_or_X.stride(0), N, D, ptr(mu), ptr(L_), L_.stride(0), int(from_z),
                                        ptr(out), stream_ptr()), "gsmvi_gauss_logq_reduce")


# ------------------------------------------------------------------------------------------------ ensemble
def _declare_ens(L):
    L.gsmvi_gsm_ensemble_fit.restype = c_i
    L.gsmvi_gsm_ensemble_fit.argtypes = [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_ull, c_p, c_p, c_i, c_p]


_declare_mon_level = _declare


def _declare(L):  # noqa: F811
    _declare_mon_level(L)
    _declare_ens(L)


def gsm_ensemble_fit_raw(P, c, mu, Sigma, F, D, B, niter, seed, z_tape, reverts, first_fit=0):
    check(lib().gsmvi_gsm_ensemble_fit(ptr(P), ptr(c), ptr(mu), ptr(Sigma), F, D, B, niter, seed & (2**64 - 1), ptr(z_tape),
                                       ptr(reverts), first_fit, stream_ptr()), "gsmvi_gsm_ensemble_fit")


# ------------------------------------------------------------------------------------------------ scaled 3xFP16 engine
c_u = ctypes.c_uint


class H3OperandC(ctypes.Structure):
    """gsmvi_h3_operand of include/gsmvi_b200.h"""
    _fields_ = [("hi", c_p), ("lo", c_p), ("scale", c_p), ("ld", c_ll)]


WS_GSM_UPDATE_H3 = 6
WS_POTRF_H3 = 7


def _declare_h3(L):
    L.gsmvi_gemm_h3.restype = c_i
    L.gsmvi_gemm_h3.argtypes = [c_p, c_p, c_p, c_ll, c_ll, c_ll, c_i, c_p, c_p, c_p, c_ll, c_ll, c_ll, c_i, c_p, c_ll,
                                c_i, c_i, c_i, c_f, c_f, c_p, c_ll, c_p, c_i, c_i, c_i, c_p, c_i, c_ll, c_p]
    L.gsmvi_h3_pair_kernel.restype = c_i
    L.gsmvi_h3_pair_kernel.argtypes = [c_i]
    L.gsmvi_h3_absmax.restype = c_i
    L.gsmvi_h3_absmax.argtypes = [c_p, c_ll, c_i, c_i, c_p, c_p]
    L.gsmvi_h3_split.restype = c_i
    L.gsmvi_h3_split.argtypes = [c_p, c_ll, c_i, c_i, c_p, c_i, c_p, c_p, c_p, c_ll, c_p]
    hp = ctypes.POINTER(H3OperandC)
    L.gsmvi_philox_normal_h3.restype = c_i
    L.gsmvi_philox_normal_h3.argtypes = [hp, c_i, c_i, c_ull, c_ull, c_p, c_p]
    L.gsmvi_sample_h3.restype = c_i
    L.gsmvi_sample_h3.argtypes = [c_p, hp, hp, c_p, c_ll, c_p, hp, c_i, c_i, c_p]
    L.gsmvi_gauss_score_h3.restype = c_i
    L.gsmvi_gauss_score_h3.argtypes = [hp, hp, c_p, c_p, c_ll, c_p, hp, c_i, c_i, c_p]
    L.gsmvi_h3_bound_scales.restype = c_i
    L.gsmvi_h3_bound_scales.argtypes = [c_p, c_i, c_p, c_p, c_f, c_f, c_f, c_p, c_p, c_p]
    L.gsmvi_potrf_h3_plan.restype = c_i
    L.gsmvi_potrf_h3_plan.argtypes = [c_i, c_i, c_p, c_i]
    L.gsmvi_potrf_h3.restype = c_i
    L.gsmvi_potrf_h3.argtypes = [c_p, c_ll, c_p, c_ll, hp, c_i, c_p, c_p, c_i, c_p]
    L.gsmvi_gsm_update_h3.restype = c_i
    L.gsmvi_gsm_update_h3.argtypes = [c_p, c_ll, c_p, c_ll, hp, c_p, c_p, c_ll, hp, c_p, c_p, c_ll, c_p, c_i, c_i, c_i,
                                      c_i, c_p, c_p]


_declare_ens_level = _declare


def _declare(L):  # noqa: F811
    _declare_ens_level(L)
    _declare_h3(L)


class HOperand:
    """fp16 (hi, lo) pair + device scale of one GEMM operand (gsmvi_h3_split)."""

    def __init__(self, rows, cols, dev, ld=None):
        import torch
        self.rows, self.cols = rows, cols
        self.ld = ld if ld is not None else (cols + 31) // 32 * 32
        self.hi = torch.zeros(rows, self.ld, dtype=torch.float16, device=dev)
        self.lo = torch.zeros(rows, self.ld, dtype=torch.float16, device=dev)
        self.scale = torch.ones(1, dtype=torch.float32, device=dev)
        self.absmax = torch.zeros(1, dtype=torch.int32, device=dev)
        self.c = H3OperandC(self.hi.data_ptr(), self.lo.data_ptr(), self.scale.data_ptr(), self.ld)
        self.ref = ctypes.byref(self.c)

    @classmethod
    def from_tensors(cls, hi, lo, scale, rows, cols):
        """Operand over existing storage (2-D fp16 views with the same row stride, 1-element fp32 scale tensor)."""
        o = cls.__new__(cls)
        o.rows, o.cols, o.ld = rows, cols, hi.stride(0)
        o.hi, o.lo, o.scale, o.absmax = hi, lo, scale, None
        o.c = H3OperandC(hi.data_ptr(), lo.data_ptr(), scale.data_ptr(), o.ld)
        o.ref = ctypes.byref(o.c)
        return o

    def split_from(self, A, sqrt_mode=False, absmax=None):
        """Split the fp32 matrix A ([rows, cols] view) into this operand; absmax: device word already holding the bit
        pattern of max |A| (else it is reduced here)."""
        if absmax is None:
            self.absmax.zero_()
            check(lib().gsmvi_h3_absmax(ptr(A), A.stride(0), self.rows, self.cols, ptr(self.absmax), stream_ptr()),
                  "gsmvi_h3_absmax")
            absmax = self.absmax
        check(lib().gsmvi_h3_split(ptr(A), A.stride(0), self.rows, self.cols, ptr(absmax), int(sqrt_mode),
                                   ptr(self.scale), ptr(self.hi), ptr(self.lo), self.ld, stream_ptr()), "gsmvi_h3_split")
        return self

    def dequant(self):
        return (self.hi[:, :self.cols].float() + self.lo[:, :self.cols].float() / 2048.0) / self.scale


def gemm_h3(A, B, C, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, Cin=None, bias_n=None, tri=False,
            mirror=False, krange=0, absmax_out=None, splits=1, split_stride=0):
    """C[M,N] = alpha * op(A) op(B)^T + beta*Cin + bias_n with A, B HOperand (K-major [rows, K] or MN-major [K, rows])."""
    rc = lib().gsmvi_gemm_h3(ptr(A.hi), ptr(A.lo), ptr(A.scale), A.rows, A.cols, A.ld, int(a_mn), ptr(B.hi), ptr(B.lo),
                             ptr(B.scale), B.rows, B.cols, B.ld, int(b_mn), ptr(C), C.stride(0), M, N, K, alpha, beta,
                             ptr(Cin), Cin.stride(0) if Cin is not None else 0, ptr(bias_n), int(tri), int(mirror),
                             krange, ptr(absmax_out), splits, split_stride, stream_ptr())
    check(rc, "gsmvi_gemm_h3")
    return C


def philox_normal_h3(Zh, B, D, seed, offset, offset_dev=None):
    check(lib().gsmvi_philox_normal_h3(Zh.ref, B, D, seed & (2**64 - 1), offset & (2**64 - 1), ptr(offset_dev),
                                       stream_ptr()), "gsmvi_philox_normal_h3")


def sample_h3(mu, Lh, Zh, X, absmax_x, B, D, X_split=None):
    check(lib().gsmvi_sample_h3(ptr(mu), Lh.ref, Zh.ref, ptr(X), X.stride(0), ptr(absmax_x),
                                X_split.ref if X_split is not None else None, B, D, stream_ptr()), "gsmvi_sample_h3")


def gauss_score_h3(Xh, Ph, c, G, absmax_g, B, D, G_split=None):
    check(lib().gsmvi_gauss_score_h3(Xh.ref, Ph.ref, ptr(c), ptr(G), G.stride(0), ptr(absmax_g),
                                     G_split.ref if G_split is not None else None, B, D, stream_ptr()),
          "gsmvi_gauss_score_h3")


def h3_bound_scales(mu, D, sigma_absmax, zmax_bits, zmax_const, pnorm, cmax, scale_x, scale_g):
    check(lib().gsmvi_h3_bound_scales(ptr(mu), D, ptr(sigma_absmax), ptr(zmax_bits), zmax_const, pnorm, cmax, ptr(scale_x),
                                      ptr(scale_g), stream_ptr()), "gsmvi_h3_bound_scales")


def gsm_update_h3(X, G, Gh, mu, Sigma, Sh, mu_out, Sigma_out, absmax_sout, B, D, B_total, mode, ws):
    check(lib().gsmvi_gsm_update_h3(ptr(X), X.stride(0), ptr(G), G.stride(0), Gh.ref, ptr(mu), ptr(Sigma), Sigma.stride(0),
                                    Sh.ref, ptr(mu_out), ptr(Sigma_out), Sigma_out.stride(0), ptr(absmax_sout), B, D,
                                    B_total, mode, ptr(ws), stream_ptr()), "gsmvi_gsm_update_h3")


def h3_pair_kernel(enable=-1):
    """Select (1 / 0) or query (-1) the persistent 2-CTA kernel behind the scaled-3xFP16 GEMM; returns the previous setting."""
    return lib().gsmvi_h3_pair_kernel(int(enable))


def h3_absmax(A, rows, cols, absmax):
    check(lib().gsmvi_h3_absmax(ptr(A), A.stride(0), rows, cols, ptr(absmax), stream_ptr()), "gsmvi_h3_absmax")


def potrf_h3_plan(D, sms=148):
    """Launch plan of potrf_h3 for a D x D matrix on `sms` SMs (host-only dry run): list of dicts, one per panel."""
    import ctypes
    n = (D + 127) // 128
    buf = (ctypes.c_int * (8 * n))()
    got = lib().gsmvi_potrf_h3_plan(D, sms, ctypes.cast(buf, ctypes.c_void_p), n)
    if got != n:
        raise GsmviError("gsmvi_potrf_h3_plan returned %d for D=%d" % (got, D))
    keys = ("j0", "fused", "panel_ctas", "gemm_ctas", "gemm_tiles", "gemm_splits", "splits_in", "helpers")
    return [dict(zip(keys, buf[8 * i: 8 * i + 8])) for i in range(n)]


def potrf_h3(Sigma, L_out, Lh, D, bad_flag, ws, zero_upper=True):
    """L_out <- chol(Sigma[:D,:D]) (fp32) and Lh <- its fp16 split; bad_flag <- 0 if PD else 1.  No sync."""
    check(lib().gsmvi_potrf_h3(ptr(Sigma), Sigma.stride(0), ptr(L_out), L_out.stride(0), Lh.ref, D, ptr(bad_flag), ptr(ws),
                               int(zero_upper), stream_ptr()), "gsmvi_potrf_h3")


# ------------------------------------------------------------------------------------------------ fp64 on int8 tensor cores
def _declare_oz(L):
    L.gsmvi_dgemm_oz_workspace_bytes.restype = c_ll
    L.gsmvi_dgemm_oz_workspace_bytes.argtypes = [c_i, c_i, c_i, c_i]
    L.gsmvi_dgemm_oz.restype = c_i
    L.gsmvi_dgemm_oz.argtypes = [c_p, c_ll, c_i, c_p, c_ll, c_i, c_p, c_ll, c_i, c_i, c_i, c_d, c_d, c_p, c_ll, c_d, c_i,
                                 c_i, c_p, c_i, c_p]


_declare_h3_level = _declare


def _declare(L):  # noqa: F811
    _declare_h3_level(L)
    _declare_oz(L)


def dgemm_oz(A, B, C, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, Cin=None, diag_add=0.0, tri=False,
             mirror=False, slices=8, ws=None):
    """fp64 C = alpha op(A) op(B)^T + beta Cin + diag_add I on the int8 tensor cores (gsmvi_dgemm_oz)."""
    import torch
    if ws is None:
        ws = torch.empty(lib().gsmvi_dgemm_oz_workspace_bytes(M, N, K, slices) + 1024, dtype=torch.uint8, device=A.device)
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    check(lib().gsmvi_dgemm_oz(ptr(A), A.stride(0), int(a_mn), ptr(B), B.stride(0), int(b_mn), ptr(C), C.stride(0), M, N, K,
                               alpha, beta, ptr(Cin), Cin.stride(0) if Cin is not None else 0, diag_add, int(tri),
                               int(mirror), ctypes.c_void_p(base), slices, stream_ptr()), "gsmvi_dgemm_oz")
    return C


# ------------------------------------------------------------------------------------------------ fp64 small-D GSM
SMALL64_INIT, SMALL64_FULL, SMALL64_SAMPLE, SMALL64_UPDATE = 0, 1, 2, 3


def _declare_small64(L):
    L.gsmvi_gsm_commit.restype = c_i
    L.gsmvi_gsm_commit.argtypes = [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p]
    L.gsmvi_gsm_small64_workspace_bytes.restype = c_ll
    L.gsmvi_gsm_small64_workspace_bytes.argtypes = [c_i, c_i]
    L.gsmvi_gsm_small64.restype = c_i
    L.gsmvi_gsm_small64.argtypes = [c_i, c_p, c_p, c_p, c_p, c_ull, c_ull, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]


_declare_oz_level = _declare


def _declare(L):  # noqa: F811
    _declare_oz_level(L)
    _declare_small64(L)


COMMIT_MAX = 12


class CommitPlan:
    """Host-side argument block of gsmvi_gsm_commit: the regions (previous state -> proposal buffers) that are copied on
    the device when an update is rejected.  Built once per buffer parity and reused every iteration."""

    def __init__(self, pairs):
        pairs = [(s, d) for (s, d) in pairs if s is not None]
        assert len(pairs) <= COMMIT_MAX
        self.n = len(pairs)
        self.keep = pairs  # the tensors own the memory the raw pointers below name
        for s_, d_ in pairs:
            assert s_.is_contiguous() and d_.is_contiguous() and s_.numel() * s_.element_size() == d_.numel() * d_.element_size()
        self.src = (c_p * max(self.n, 1))(*[s_.data_ptr() for s_, _ in pairs])
        self.dst = (c_p * max(self.n, 1))(*[d_.data_ptr() for _, d_ in pairs])
        self.nbytes = (c_ll * max(self.n, 1))(*[s_.numel() * s_.element_size() for s_, _ in pairs])


def gsm_commit(bad, plan, status, bad2=None):
    check(lib().gsmvi_gsm_commit(ptr(bad), ptr(bad2), plan.n, plan.src, plan.dst, plan.nbytes, ptr(status), stream_ptr()),
          "gsmvi_gsm_commit")


def gsm_small64_workspace_bytes(B, D):
    n = lib().gsmvi_gsm_small64_workspace_bytes(B, D)
    if n < 0:
        raise GsmviError("gsmvi_gsm_small64 supports 1 <= D <= 64 (got D=%d, B=%d)" % (D, B))
    return n


def gsm_small64(mode, mu, Sigma, L_, z_tape, seed, iter0, X, G, P, c, B, D, iters, status, ws):
    check(lib().gsmvi_gsm_small64(mode, ptr(mu), ptr(Sigma), ptr(L_), ptr(z_tape), seed & (2**64 - 1), iter0 & (2**64 - 1),
                                  ptr(X), ptr(G), ptr(P), ptr(c), B, D, iters, ptr(status), ptr(ws), stream_ptr()),
          "gsmvi_gsm_small64")


# ------------------------------------------------------------------------------------------------ ADVI
def advi_step(L_, mu, G, Z, GtZ, gsum, mL, vL, m_mu, v_mu, B, D, lr, b1, b2, eps, t, npass=3):
    f = lib().gsmvi_advi_step
    if not getattr(f, "_declared", False):
        f.restype = c_i
        f.argtypes = [c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_f, c_f, c_f,
                      c_i, c_i, c_p]
        f._declared = True
    check(f(ptr(L_), L_.stride(0), ptr(mu), ptr(G), G.stride(0), ptr(Z), Z.stride(0), ptr(GtZ), GtZ.stride(0), ptr(gsum),
            ptr(mL), ptr(vL), ptr(m_mu), ptr(v_mu), B, D, lr, b1, b2, eps, t, npass, stream_ptr()), "gsmvi_advi_step")


def potrf64(A, n, bad_flag):
    """In-place fp64 Cholesky (gsmvi_potrf64): A [n, ld] float64 CUDA tensor."""
    f = lib().gsmvi_potrf64
    if not getattr(f, "_declared", False):
        f.restype = c_i
        f.argtypes = [c_p, c_ll, c_i, c_p, c_p]
        f._declared = True
    check(f(ptr(A), A.stride(0), n, ptr(bad_flag), stream_ptr()), "gsmvi_potrf64")
