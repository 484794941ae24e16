"""BaM: batch-and-match VI on B200 - drop-in for gsmvi/bam.py of modichirag/GSM-VI.

Same public surface: `bam_update(samples, vs, mu0, S0, reg)` (gsmvi/bam.py:31-69), `bam_lowrank_update(...)`
(gsmvi/bam.py:72-114), `BaM(D, lp, lp_g, use_lowrank, jit_compile).fit(key, regf, ...) -> (mean, cov)`
(gsmvi/bam.py:117-216) and `Regularizers` (gsmvi/bam.py:237-274).  Sampling and scores run on the 3xTF32
tensor-core path; the batch statistics and the D x D quadratic-matrix-equation solve (scaled Newton-Schulz square root)
are fp64, all inside libgsmvi_b200.so.
"""
import torch

from . import _lib as L
from ._util import device, key_to_seed, new_mat, new_vec, to_dev
from .gsm import _monitor_params


def _update(samples, vs, mu0, S0, reg, lowrank, npass=3, jitter=0.0):
    dev = device()
    samples, vs, mu0, S0 = (to_dev(a, dev) for a in (samples, vs, mu0, S0))
    assert samples.dim() == 2 and vs.dim() == 2  # bam.py:46-47
    B, D = samples.shape
    if lowrank and B + 1 >= D:
        raise ValueError("bam_lowrank_update needs batch_size + 1 < D (the reference's svds(U, k=B) needs B < D)")
    Xb, X = new_mat(B, D, dev)
    Gb, G = new_mat(B, D, dev)
    Sb, S = new_mat(D, D, dev)
    Ob, O = new_mat(D, D, dev)
    X.copy_(samples)
    G.copy_(vs)
    S.copy_(S0)
    mu = new_vec(D, dev)
    mu[:D].copy_(mu0)
    mu_out = new_vec(D, dev)
    ws_s = torch.empty(L.workspace_bytes(L.WS_BAM_STATS, B, D) // 8, dtype=torch.float64, device=dev)
    ws_v = torch.empty(L.workspace_bytes(L.WS_BAM_SOLVE_LOWRANK if lowrank else L.WS_BAM_SOLVE, B, D) // 8,
                       dtype=torch.float64, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    L.bam_stats(Xb, Gb, B, D, B, ws_s, 0, npass)
    L.bam_stats(Xb, Gb, B, D, B, ws_s, 1, npass)
    L.bam_solve(ws_s, B, D, B, mu, Sb, reg, jitter, mu_out, Ob, ws_v, bad, lowrank=lowrank)
    if int(bad.item()) != 0:
        raise FloatingPointError("BaM update: V or I + (I + 4 L^T U L)^(1/2) is not positive definite")
    return mu_out[:D].clone(), O.clone()


def bam_update(samples, vs, mu0, S0, reg, npass=3):
    """Drop-in for gsmvi/bam.py:31-69: returns (mu, S) with S solving S U S + S = V.  CUDA tensors."""
    return _update(samples, vs, mu0, S0, reg, lowrank=False, npass=npass)


def bam_lowrank_update(samples, vs, mu0, S0, reg, npass=3):
    """Drop-in for gsmvi/bam.py:72-114 (requires batch_size + 1 < D)."""
    return _update(samples, vs, mu0, S0, reg, lowrank=True, npass=npass)


class BaMEngine:
    """Device-resident state and workspaces of one BaM fit; `step(i, reg)` is one loop body of gsmvi/bam.py:188-212."""

    def __init__(self, D, batch_size, lp_g, key, mean=None, cov=None, use_lowrank=False, jitter=1e-6, z_tape=None,
                 npass=3, process_group=None, score_input="torch", max_ns=200):
        dev = self.dev = device()
        self.D, self.batch_size, self.lp_g, self.npass = D, batch_size, lp_g, npass
        self.use_lowrank, self.jitter, self.max_ns = use_lowrank, jitter, max_ns
        self.group, self.dist, self.rank, self.world = process_group, None, 0, 1
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        if batch_size % self.world != 0:
            raise ValueError("batch_size must be divisible by the number of ranks")
        B = self.B = batch_size // self.world
        if use_lowrank and batch_size + 1 >= D:
            raise ValueError("the low-rank update needs batch_size + 1 < D (the reference's svds(U, k=B) needs B < D)")
        self.Sb, self.S = new_mat(D, D, dev)
        self.Snb, self.Sn = new_mat(D, D, dev)
        self.Lb, _ = new_mat(D, D, dev)
        self.Lnb, _ = new_mat(D, D, dev)
        self.mu, self.mun = new_vec(D, dev), new_vec(D, dev)
        self.Zb, self.Z = new_mat(B, D, dev)
        self.Xb, self.X = new_mat(B, D, dev)
        self.Gb, self.G = new_mat(B, D, dev)
        self.ws_p = torch.empty(L.workspace_bytes(L.WS_POTRF, B, D) // 4, dtype=torch.float32, device=dev)
        self.ws_s = torch.empty(L.workspace_bytes(L.WS_BAM_STATS, B, D) // 8, dtype=torch.float64, device=dev)
        kind = L.WS_BAM_SOLVE_LOWRANK if use_lowrank else L.WS_BAM_SOLVE
        nbytes_v = L.workspace_bytes(kind, batch_size if use_lowrank else B, D)
        self.peer = None
        if self.world > 1 and not use_lowrank and self.world <= 8 and __import__("os").environ.get("GSMVI_BAM_TP", "1") != "0":
            # tensor-parallel solve: the solve workspace of every rank is mapped into every other
            # rank, the Newton-Schulz / T T^T products are split by rows and their epilogues store through NVLink
            import ctypes
            from ._comm import BamShardC, PeerAlloc
            self.peer = PeerAlloc(nbytes_v, process_group, self.dist)
            self.ws_v = self.peer.flat.view(torch.float64)
            self._epoch = ctypes.c_uint(0)
            self.shard = BamShardC()
            self.shard.rank, self.shard.world = self.rank, self.world
            for r in range(self.world):
                self.shard.peer_ws[r] = self.peer.ptrs[r]
            self.shard.epoch_host = ctypes.pointer(self._epoch)
        else:
            self.ws_v = torch.empty(nbytes_v // 8, dtype=torch.float64, device=dev)
        if use_lowrank and self.world > 1:
            # sharded low-rank update: the K = batch_size + 1 columns of U's exact factor Q come from every rank's centred
            # scores, gathered into a full-batch statistics block; the O(D^2 K) solve is then replicated (bit-identical)
            self.ws_full = torch.zeros(L.workspace_bytes(L.WS_BAM_STATS, batch_size, D) // 8, dtype=torch.float64, device=dev)
        self.bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bad2 = torch.zeros(1, dtype=torch.int32, device=dev)
        try:
            self.reset(lp_g, key, mean, cov, z_tape, score_input, jitter)
        except Exception:
            self.close(collective=False)  # every rank sees the same start
            raise

    def reset(self, lp_g, key, mean=None, cov=None, z_tape=None, score_input="torch", jitter=1e-6):
        """Start a new fit on this engine's buffers (bam.py:163-168).  BaM.fit keeps engines - workspaces and, on a sharded
        fit, the peer-mapped solve workspace - across calls, as GSM.fit does."""
        D, dev = self.D, self.dev
        self.lp_g, self.score_input, self.jitter = lp_g, score_input, jitter
        self.seed = key_to_seed(key)
        self.mu.zero_()
        if mean is not None:
            self.mu[:D].copy_(to_dev(mean, dev))  # bam.py:163-164
        if cov is None:
            self.S.copy_(torch.eye(D, device=dev))  # bam.py:165-166
        else:
            self.S.copy_(to_dev(cov, dev))
        if z_tape is not None and not isinstance(z_tape, torch.Tensor):
            z_tape = torch.as_tensor(z_tape, dtype=torch.float32)
        self.z_tape = z_tape
        self.target = getattr(getattr(lp_g, "__self__", None), "_gsmvi_builtin_target", None)
        self.n_reverts = 0
        self.ns_iters = []
        self.draws = 0
        L.potrf_check(self.Sb, self.Lb, D, self.bad, self.ws_p, self.npass)
        if int(self.bad.item()) != 0:
            raise ValueError("initial covariance is not positive definite")

    def _stats_views(self):
        ld = (self.D + 7) // 8 * 8
        off = 2 * self.B * ld
        c = self.ws_s[off: off + self.D * ld]
        means = self.ws_s[off + self.D * ld: off + self.D * ld + 2 * ld]
        return c, means

    def draw_and_score(self, i):
        """sample -> score (bam.py:191-194)."""
        D, B, npass = self.D, self.B, self.npass
        if self.z_tape is not None:
            self.Z.copy_(self.z_tape[i, self.rank * B:(self.rank + 1) * B], non_blocking=True)
        else:
            L.philox_normal(self.Zb, B, D, self.seed, self.draws * self.world + self.rank)
        self.draws += 1
        L.sample(self.mu, self.Lb, self.Zb, self.Xb, B, D, npass)
        if self.target is not None:
            L.gauss_score(self.Xb, self.target.Phib, self.target.c, self.Gb, B, D, npass, P_lo=self.target.Plob)
        elif self.score_input == "numpy":
            self.G.copy_(to_dev(self.lp_g(self.X.cpu().numpy()), self.dev))
        else:
            self.G.copy_(to_dev(self.lp_g(self.X), self.dev))

    def update(self, reg):
        """bam_update / bam_lowrank_update + jitter + symmetrise (bam.py:197-199); the proposal is left in (mun, Snb).
        Returns the number of Newton-Schulz iterations."""
        D, B, npass = self.D, self.B, self.npass
        L.bam_stats(self.Xb, self.Gb, B, D, self.batch_size, self.ws_s, 0, npass)
        if self.world > 1:
            self.dist.all_reduce(self._stats_views()[1], group=self.group)
        L.bam_stats(self.Xb, self.Gb, B, D, self.batch_size, self.ws_s, 1, npass)
        if self.world > 1:
            self.dist.all_reduce(self._stats_views()[0], group=self.group)
        args = (self.ws_s, B, D, self.batch_size, self.mu, self.Sb, reg, self.jitter, self.mun, self.Snb, self.ws_v,
                self.bad2)
        if self.world == 1:
            it = L.bam_solve(*args, lowrank=self.use_lowrank, max_ns=self.max_ns)
        elif self.use_lowrank:
            # bam.py:72-114 on a sharded batch: C, xbar, gbar are already global; gather the centred score rows (rank order =
            # sample order) so every rank holds Q = [sqrt(reg/B) Gc^T, sqrt(reg/(1+reg)) gbar] in full, then solve as one rank
            ld = (D + 7) // 8 * 8
            Bt = self.batch_size
            gc_local = self.ws_s[B * ld: 2 * B * ld]
            self.dist.all_gather_into_tensor(self.ws_full[Bt * ld: 2 * Bt * ld], gc_local, group=self.group)
            self.ws_full[2 * Bt * ld: 2 * Bt * ld + (D + 2) * ld].copy_(self.ws_s[2 * B * ld: 2 * B * ld + (D + 2) * ld])
            it = L.bam_solve(self.ws_full, Bt, D, Bt, self.mu, self.Sb, reg, self.jitter, self.mun, self.Snb, self.ws_v,
                             self.bad2, lowrank=True, max_ns=self.max_ns)
        else:  # shard partials of M = I + 4 W W^T are summed between the two phases of the solve
            L.bam_solve(*args, max_ns=self.max_ns, world=self.world, phase=1)
            ld = (D + 7) // 8 * 8
            self.dist.all_reduce(self.ws_v[3 * D * ld: 4 * D * ld], group=self.group)
            if self.peer is not None:
                it = L.bam_solve_sharded(*args, self.shard, max_ns=self.max_ns)
            else:
                it = L.bam_solve(*args, max_ns=self.max_ns, world=self.world, phase=2)
        self.ns_iters.append(it)
        # bit 1 of the solve's flag: the Newton-Schulz square root did not converge (the reference's scipy sqrtm raises
        # there, and BaM.fit's retry loop - bam.py:188-206 - draws fresh samples).  With a fixed draw tape a retry would see
        # the same samples, so the proposal is left to the goodness check, which rejects it (flag != 0).
        if self.z_tape is None and (int(self.bad2.item()) & 2):
            raise FloatingPointError("BaM update: the Newton-Schulz square root did not converge")
        return it

    def propose(self, i, reg):
        self.draw_and_score(i)
        return self.update(reg)

    def accept_or_revert(self):
        """_check_goodness (bam.py:208-212, 219-233) = Cholesky of the proposal, reused as the next sampling factor."""
        L.potrf_check(self.Snb, self.Lnb, self.D, self.bad, self.ws_p, self.npass)
        ok = int(self.bad.item()) == 0 and int(self.bad2.item()) == 0
        if ok:
            self.Sb, self.Snb, self.S, self.Sn = self.Snb, self.Sb, self.Sn, self.S
            self.Lb, self.Lnb = self.Lnb, self.Lb
            self.mu, self.mun = self.mun, self.mu
        else:
            self.n_reverts += 1
        return ok

    def step(self, i, reg):
        self.propose(i, reg)
        return self.accept_or_revert()

    def mean(self):
        return self.mu[: self.D]

    def cov(self):
        return self.S

    def chol_buffer(self):
        """Padded fp32 buffer of the lower Cholesky factor of cov(): the goodness check's factor of the accepted state."""
        return self.Lb

    def close(self, collective=True):
        """Release the peer-mapped solve workspace of a tensor-parallel fit (collective: every rank calls it)."""
        if self.peer is not None:
            self.ws_v = None
            self.peer.close(collective=collective)
            self.peer = None


class BaM:
    """Wrapper class for using BaM updates to fit a distribution (gsmvi/bam.py:117-137)."""

    def __init__(self, D, lp, lp_g, use_lowrank=False, jit_compile=True):
        self.D = D
        self.lp = lp
        self.lp_g = lp_g
        self.use_lowrank = use_lowrank
        if use_lowrank:
            print("Using lowrank update")  # bam.py:133-134
        self.jit_compile = jit_compile  # accepted for API compatibility; there is nothing to jit

    def fit(self, key, regf, mean=None, cov=None, batch_size=2, niter=5000, nprint=10, verbose=True,
            check_goodness=True, monitor=None, retries=10, jitter=1e-6, *, z_tape=None, npass=3, process_group=None,
            score_input="torch"):
        """Main function to fit a multivariate Gaussian to the target (gsmvi/bam.py:140-216).  Reference arguments keep
        their meaning; keyword-only extras as in gsm.GSM.fit.  Returns (mean[D], cov[D, D]) as CUDA tensors."""
        from . import gsm as _gsm
        ekey = _gsm._engine_key("bam-lowrank" if self.use_lowrank else "bam", self.D, batch_size, npass, process_group)
        eng = _gsm._ENGINES.pop(ekey, None)
        if eng is None:
            while len(_gsm._ENGINES) >= _gsm._ENGINES_MAX:  # same call sequence on every rank: collective evictions
                _gsm._evict(next(iter(_gsm._ENGINES)), collective=True)
            eng = BaMEngine(self.D, batch_size, self.lp_g, key, mean, cov, self.use_lowrank, jitter, z_tape, npass,
                            process_group, score_input)
        else:
            try:
                eng.reset(self.lp_g, key, mean, cov, z_tape, score_input, jitter)
            except Exception:
                _gsm._ENGINES[ekey] = eng
                raise
        try:
            out = self._fit_loop(eng, key, regf, batch_size, niter, nprint, verbose, monitor, retries)
        except BaseException:
            eng.close(collective=False)  # a failed rank must not wait for its peers in a barrier
            raise
        eng.z_tape = None
        _gsm._ENGINES[ekey] = eng
        return out

    def _fit_loop(self, eng, key, regf, batch_size, niter, nprint, verbose, monitor, retries):
        nevals = 1  # bam.py:168
        if nprint > niter:
            nprint = niter  # bam.py:177
        every = max(niter // max(nprint, 1), 1)
        i = 0
        for i in range(niter + 1):  # bam.py:178
            if verbose and (i % every == 0):
                print(f"Iteration {i} of {niter}")
            if monitor is not None and (i % monitor.checkpoint) == 0:  # bam.py:182-185
                monitor(i, _monitor_params(eng), self.lp, key, nevals=nevals)
                nevals = 0
            j = 0
            while True:  # bam.py:188-206: retry on ANY exception (bad sample, failed callback, ...)
                try:
                    eng.draw_and_score(i)  # bam.py:191-194
                    nevals += batch_size  # bam.py:195
                    reg = regf(i)  # bam.py:196 (the schedule advances on every call, retries included)
                    eng.update(reg)  # bam.py:197-199
                    break
                except Exception as e:
                    if j < retries:
                        j += 1
                        print(f"Failed with exception {e}")
                        print(f"Trying again {j} of {retries}")
                    else:
                        raise e
            if not eng.accept_or_revert() and verbose:  # bam.py:208-212
                print("Bad update for covariance matrix. Revert")
        if monitor is not None:  # bam.py:214-215
            monitor(i, _monitor_params(eng), self.lp, key, nevals=nevals)
        self.n_reverts = eng.n_reverts
        self.ns_iters = eng.ns_iters
        return eng.mean().clone(), eng.cov().clone()


class Regularizers:
    """Class for regularizers used in BaM (gsmvi/bam.py:237-274).  As in the reference, the `iteration` argument of
    the returned callables is ignored and an internal counter advances on every call."""

    def __init__(self):
        self.counter = 0

    def reset(self):
        self.counter = 0

    def constant(self, reg0):
        def reg_iter(iteration):
            self.counter += 1
            return reg0
        return reg_iter

    def linear(self, reg0):
        def reg_iter(iteration):
            self.counter += 1
            return reg0 / self.counter
        return reg_iter

    def custom(self, func):
        def reg_iter(iteration):
            self.counter += 1
            return func(self.counter)
        return reg_iter
