"""GSM: Gaussian score matching VI on B200 - drop-in for gsmvi/gsm.py (and gsmvi/gsm_numpy.py) of modichirag/GSM-VI.

Same public surface: `gsm_update(samples, vs, mu0, S0) -> (mu, S)` (gsmvi/gsm.py:31-58) and
`GSM(D, lp, lp_g).fit(key, mean, cov, batch_size, niter, nprint, verbose, check_goodness, monitor) -> (mean, cov)`
(gsmvi/gsm.py:62-133).  The loop body runs as sm_100a kernels through libgsmvi_b200.so: blocked Cholesky (PD check +
sampling factor), X = mu + Z L^T, score, fused GSM update.  Tensors are torch.float32 on the current CUDA device.
"""
import numpy as np
import torch

from . import _lib as L
from ._util import device, key_to_seed, ld_of, new_mat, new_vec, to_dev


def gsm_update(samples, vs, mu0, S0, npass=4):
    """Drop-in for gsmvi/gsm.py:31-58.  samples, vs: [B, D]; mu0: [D]; S0: [D, D] (symmetric).  Returns (mu, S) as
    new CUDA tensors.  npass = 4 (default) runs the scaled 3xFP16 engine GSM.fit uses (gsmvi_gsm_update_h3), npass <= 3
    the 3xTF32 engine (gsmvi_gsm_update)."""
    dev = device()
    samples, vs, mu0, S0 = (to_dev(a, dev) for a in (samples, vs, mu0, S0))
    assert samples.dim() == 2 and vs.dim() == 2  # gsm.py:48-49
    B, D = samples.shape
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
    if npass == 4:
        Gh = L.HOperand(B, D, dev).split_from(G)
        Sh = L.HOperand(D, D, dev).split_from(S)
        amax = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE_H3, B, D) // 4, dtype=torch.float32, device=dev)
        L.gsm_update_h3(Xb, Gb, Gh, mu, Sb, Sh, mu_out, Ob, amax, B, D, B, 0, ws)
    else:
        ws = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE, B, D) // 4, dtype=torch.float32, device=dev)
        L.gsm_update_raw(Xb, Gb, mu, Sb, mu_out, Ob, B, D, B, 0, ws, npass)
    return mu_out[:D].clone(), O.clone()


def launch_count(D, h3=True, tape=False, builtin_target=True, world=1, lookahead=True):
    """Library kernels launched by one GSM iteration (GSMEngine.step), by engine and options."""
    panels = (D + 127) // 128
    if h3:
        potrf = 1 + panels + (panels - 1)  # prepare, panel kernels, left-looking update GEMMs
        if lookahead and 512 <= D <= 128 * 133:
            # look-ahead (csrc/potrf_h3.cu): the update GEMMs run inside the fused panel launches; only a ragged last
            # panel still has a GEMM launch of its own
            potrf = 1 + panels + (1 if D % 128 else 0)
        draw = 2 if tape else 1  # |Z| max + split of a tape slice, or Philox written split
        score = 4 if builtin_target else 3  # sample, split X, score GEMM, split G | sample, |G| max, split G
        upd = 5 + 2  # W GEMM, row pass, split T, covariance GEMM, axpy + split Sigma_new, its max|.| word copied
        return draw + score + upd + potrf + 1 + (3 if world > 1 else 0)  # + the device-side accept / revert
    potrf = 1 + panels + (panels - 1)  # tril copy, panel kernels, SYRK GEMMs
    upd = 4 + 2  # W GEMM, row pass, covariance GEMM, axpy + the two tf32_split launches (Sigma_new, L_new)
    return (0 if tape else 1) + 1 + (1 if builtin_target else 0) + upd + potrf + 1 + (2 if world > 1 else 0)


def _monitor_params(eng):
    """[mean, cov] for a monitor call (gsmvi/gsm.py:113), carrying the engine's Cholesky factor of that covariance when the
    engine keeps one as a plain fp32 matrix (monitors.MonitorParams.chol): KLMonitor then does not factor it again."""
    from .monitors import MonitorParams
    params = MonitorParams([eng.mean(), eng.cov()])
    params.chol = eng.chol_buffer() if hasattr(eng, "chol_buffer") else None
    return params


class GSMEngine:
    """Device-resident state and workspaces of one GSM fit; `step(i)` is one loop body of gsmvi/gsm.py:107-129
    (sample -> score -> update -> goodness check -> accept/revert).  GSM.fit drives it; bench.py times it.

    The accept / revert of gsm.py:125-129 happens on the device (gsmvi_gsm_commit): the host flips its buffer pointers
    every iteration and a rejected proposal is overwritten with the previous state by a predicated copy, so no iteration
    waits for a flag to come back (`sync_accept=True` - the verbose path - reads the 4-byte status word per step to print
    the reference's "Revert" message in place).  Engines are reusable: `reset()` starts a new fit on the same buffers
    (GSM.fit keeps the last few engines, peer-mapped exchange buffers included, across calls)."""

    def __init__(self, D, batch_size, lp_g, key, mean=None, cov=None, z_tape=None, npass=4, process_group=None,
                 score_input="torch", sync_accept=False):
        dev = self.dev = device()
        self.D, self.batch_size, self.npass = D, batch_size, npass
        self.group, self.dist, self.rank, self.world = process_group, None, 0, 1
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        if batch_size % self.world != 0:
            raise ValueError("batch_size must be divisible by the number of ranks")
        B = self.B = batch_size // self.world
        self._phase_on = bool(__import__("os").environ.get("GSMVI_PHASE_TIMING"))
        self._phase_events = []
        self.h3 = npass == 4
        # state, double-buffered: the proposal is written beside the current state and the roles are exchanged
        self.comm = None
        if self.h3 and self.world > 1:
            # batch-sharded fit on the h3 engine: the two Sigma buffers live in this rank's peer-mapped exchange buffer and
            # the statistics travel through NVLink peer memory, fused with the covariance GEMM (csrc/comm.cu)
            from ._comm import CommBuffer
            self.comm = CommBuffer(D, process_group, self.dist)
            self.Sb, self.Snb = self.comm.S[0], self.comm.S[1]
            self.S, self.Sn = self.Sb[:, :D], self.Snb[:, :D]
        else:
            self.Sb, self.S = new_mat(D, D, dev)
            self.Snb, self.Sn = new_mat(D, D, dev)
        self.cur = 0
        self.Lb, _ = new_mat(D, D, dev)
        self.Lnb, _ = new_mat(D, D, dev)
        if not self.h3:
            # pre-split low parts of the reused GEMM operands (Sigma for W = G Sigma, L for the sampler), kept per buffer
            self.Shi, self.Slo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Snhi, self.Snlo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Lhi, self.Llo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Lnhi, self.Lnlo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
        self.mu, self.mun = new_vec(D, dev), new_vec(D, dev)
        self.Zb, self.Z = new_mat(B, D, dev)
        self.Xb, self.X = new_mat(B, D, dev)
        self.Gb, self.G = new_mat(B, D, dev)
        self.bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)  # [0] rejected updates, [1] last update accepted
        self.status_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        if self.world > 1 and self.comm is None:
            self.dSb, _ = new_mat(D, D, dev)
            self.dmu = new_vec(D, dev)
        self.copy_stream = None
        self._side_stream = None
        self._side_ok = __import__("os").environ.get("GSMVI_SIDE_STREAM", "1")[:1] != "0"
        self._philox_early = __import__("os").environ.get("GSMVI_SIDE_PHILOX", "early") != "chol"
        if self.h3:
            # scaled 3xFP16 engine: every GEMM operand lives as an fp16 (hi, lo) pair + power-of-two scale
            H = L.HOperand
            self.Sh, self.Snh, self.Lh, self.Lnh = H(D, D, dev), H(D, D, dev), H(D, D, dev), H(D, D, dev)
            self.Zh, self.Xh, self.Gh = H(B, D, dev), H(B, D, dev), H(B, D, dev)
            self.slots = torch.zeros(8, dtype=torch.int32, device=dev)  # |X|, |G|, |Sigma_new| maxima (bit patterns)
            self.ctr = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox counter of the next draw (graph mode)
            self.ws_u = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE_H3, B, D) // 4, dtype=torch.float32, device=dev)
            self.ws_p = torch.empty(L.workspace_bytes(L.WS_POTRF_H3, B, D) // 4, dtype=torch.float32, device=dev)
        else:
            self.ws_p = torch.empty(L.workspace_bytes(L.WS_POTRF, B, D) // 4, dtype=torch.float32, device=dev)
            self.ws_u = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE, B, D) // 4, dtype=torch.float32, device=dev)
        self._plans, self._graphs = {}, {}
        self.parity = 0
        try:
            self.reset(lp_g, key, mean, cov, z_tape, score_input, sync_accept)
        except Exception:
            self.close(collective=False)  # every rank sees the same start, so every rank lands here
            raise

    # ------------------------------------------------------------------------------------------------ (re)start a fit
    def reset(self, lp_g, key, mean=None, cov=None, z_tape=None, score_input="torch", sync_accept=False):
        """Start a new fit on this engine's buffers: gsm.py:100-105 (defaults mu = 0, Sigma = I) plus the first
        factorisation.  A start that is not numerically positive definite - lbfgs_init's dense inverse-Hessian estimate
        (gsmvi/initializers.py:5-17) can be - is symmetrised and given the smallest diagonal shift eps * mean(diag) that
        lets the Cholesky through, with a warning; the reference gets past such a start because numpy's SVD-based sampler
        only warns (gsm.py:119)."""
        D, B, dev = self.D, self.B, self.dev
        self.lp_g, self.score_input, self.sync_accept = lp_g, score_input, sync_accept
        seed = key_to_seed(key)
        if getattr(self, "seed", None) != seed:
            self._graphs = {}  # the Philox seed is a captured kernel argument
        self.seed = seed
        self.mu.zero_()
        if mean is not None:
            self.mu[:D].copy_(to_dev(mean, dev))  # gsm.py:100-101 (default zeros)
        if cov is None:
            self.S.copy_(torch.eye(D, device=dev))  # gsm.py:102-103
        else:
            self.S.copy_(to_dev(cov, dev))
        if z_tape is not None and not isinstance(z_tape, torch.Tensor):
            z_tape = torch.as_tensor(z_tape, dtype=torch.float32)
        self._tape_off = self.rank * B
        if z_tape is not None:
            # [n, batch_size, D]: the global draws, every rank reads its slice; [n, batch_size / world, D] on a sharded
            # fit: this rank's own draws (a rank then only holds - and pins - what it copies)
            if self.world > 1 and z_tape.shape[1] == B:
                self._tape_off = 0
            else:
                assert z_tape.shape[1] == self.batch_size
            assert z_tape.shape[2] == D
        self.z_tape = z_tape
        self._tape_async = bool(z_tape is not None and self.npass == 4 and not z_tape.is_cuda and z_tape.is_pinned())
        if self._tape_async:
            if self.copy_stream is None:
                self.copy_stream = torch.cuda.Stream()
                self.Zbufs = [self.Z, new_mat(B, D, dev)[1]]
            self.z_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self.z_free = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self.z_free:
                e.record()
            self._tape_copied, self._z_buf_in_use = -1, 0
        self.target = getattr(getattr(lp_g, "__self__", None), "_gsmvi_builtin_target", None)
        self.status.zero_()
        self.z_drawn_for, self._steps_done, self._ctr_next = -1, 0, -1
        if self.h3:
            # launch-bound sizes only: at D = 512 the replayed step is 29% faster (4378 vs 3393 it/s), at D = 4096 the
            # graph's kernel nodes lose the programmatic-dependent-launch overlap of the Cholesky chain and it is slower
            genv = __import__("os").environ.get("GSMVI_GRAPH", "auto")
            self._graph_ok = (self.world == 1 and z_tape is None and genv != "0" and (D <= 1024 or genv == "1")
                              and self.target is not None and not sync_accept)
        self._factor_start(cov is not None)

    def _factor_current(self):
        D = self.D
        if self.h3:
            L.potrf_h3(self.Sb, self.Lb, self.Lh, D, self.bad, self.ws_p, zero_upper=False)  # upper blocks stay zero
        else:
            L.potrf_check(self.Sb, self.Lb, D, self.bad, self.ws_p, self.npass)
        return int(self.bad.item()) == 0

    def _factor_start(self, user_cov):
        D = self.D
        ok = self._factor_current()
        if not ok and user_cov:
            import warnings
            self.S.copy_((self.S + self.S.t()) * 0.5)
            diag = self.S.diagonal()
            base = float(diag.abs().mean().item())
            if not (base > 0.0) or base != base or base == float("inf"):
                raise ValueError("initial covariance is not positive definite")
            shift_total = 0.0
            for eps in (0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0):
                diag.add_(eps * base - shift_total)
                shift_total = eps * base
                if self._factor_current():
                    warnings.warn("gsmvi_b200: initial covariance is not numerically positive definite; symmetrised and "
                                  "shifted by %.1e * mean|diag| to start the fit" % eps)
                    ok = True
                    break
        if not ok:
            raise ValueError("initial covariance is not positive definite")
        if self.h3:
            self.Sh.split_from(self.S)
        else:
            L.tf32_split(self.Sb, self.Shi, self.Slo, D, D)
            L.tf32_split(self.Lb, self.Lhi, self.Llo, D, D)

    def launches_per_step(self):
        """Kernels of libgsmvi_b200.so launched by one step (bench.py reports it as gpu_launches)."""
        look = __import__("os").environ.get("GSMVI_POTRF_LOOKAHEAD", "1")[:1] != "0"
        return launch_count(self.D, self.h3, self.z_tape is not None, self.target is not None, self.world, look)

    def gemm_calls(self):
        """The four batch-sized tensor-core launches of one step (sampler, score, W = G Sigma, covariance update) as
        callables on the engine's current buffers; bench.py brackets each with CUDA events for the roofline figure."""
        D, B, npass = self.D, self.B, self.npass
        ldw = ld_of(D)
        ws = self.ws_u
        tgt = self.target
        if self.h3:
            off = (4 * B + 1) * ldw + 32
            th = ws[off:].view(torch.float16)
            Thi = th[: 3 * B * ldw].view(3 * B, ldw)
            Tlo = th[3 * B * ldw: 6 * B * ldw].view(3 * B, ldw)
            tscale = ws[(4 * B + 1) * ldw + 1: (4 * B + 1) * ldw + 2]
            W = ws[: B * ldw].view(B, ldw)
            Ta = L.HOperand.from_tensors(Thi[: 2 * B], Tlo[: 2 * B], tscale, 2 * B, D)
            Tb = L.HOperand.from_tensors(Thi[B:], Tlo[B:], tscale, 2 * B, D)
            return [
                lambda: L.sample_h3(self.mu, self.Lh, self.Zh, self.Xb, None, B, D),
                lambda: L.gauss_score_h3(self.Xh, tgt.Ph, tgt.c, self.Gb, None, B, D),
                lambda: L.gemm_h3(self.Gh, self.Sh, W, B, D, D),
                lambda: L.gemm_h3(Ta, Tb, self.Snb, D, D, 2 * B, a_mn=True, b_mn=True, alpha=-1.0 / self.batch_size,
                                  beta=1.0, Cin=self.Sb, tri=True, mirror=True),
            ]
        W = ws[: B * ldw].view(B, ldw)
        T = ws[B * ldw: 4 * B * ldw].view(3 * B, ldw)
        Tlo = ws[4 * B * ldw: 7 * B * ldw].view(3 * B, ldw)
        return [
            lambda: L.sample(self.mu, self.Lhi, self.Zb, self.Xb, B, D, npass, L_lo=self.Llo),
            lambda: L.gauss_score(self.Xb, tgt.Phib, tgt.c, self.Gb, B, D, npass, P_lo=tgt.Plob),
            lambda: L.gemm_tf32(self.Gb[:, :D], self.Shi[:, :D], W[:, :D], B, D, D, npass=npass, B_lo=self.Slo[:, :D]),
            lambda: L.gemm_tf32(T[: 2 * B, :D], T[B:, :D], self.Snb[:, :D], D, D, 2 * B, a_mn=True, b_mn=True,
                                alpha=-1.0 / self.batch_size, beta=1.0, Cin=self.Sb[:, :D], tri=True, mirror=True,
                                npass=npass, A_lo=Tlo[: 2 * B, :D], B_lo=Tlo[B:, :D]),
        ]

    def _tape_copy(self, j):
        """Queue the host -> device copy of this rank's draws of iteration j on the copy stream (pinned tape only)."""
        B, buf = self.B, j & 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.z_free[buf])  # the split that last read this buffer has run
            self.Zbufs[buf].copy_(self.z_tape[j, self._tape_off:self._tape_off + B], non_blocking=True)
            self.z_ready[buf].record()
        self._tape_copied = j

    def _tape_slice(self, i):
        """This rank's [B, D] draws of iteration i as a device matrix.  A pinned host tape is streamed: the slice of
        iteration i + 1 is copied on a second stream while iteration i computes (the reference draws its samples on the
        host every iteration, gsmvi/gsm.py:117-119; this is that hand-over without the stall)."""
        B = self.B
        if not self._tape_async:
            self.Z.copy_(self.z_tape[i, self._tape_off:self._tape_off + B], non_blocking=True)
            return self.Z
        if self._tape_copied < i:
            self._tape_copy(i)
        buf = i & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.z_ready[buf])
        if i + 1 < self.z_tape.shape[0]:
            self._tape_copy(i + 1)
        self._z_buf_in_use = buf
        return self.Zbufs[buf]

    def _commit_plan(self):
        """(previous state -> proposal buffers) regions of gsmvi_gsm_commit for the current buffer roles."""
        plan = self._plans.get(self.parity)
        if plan is None:
            if self.h3:
                pairs = [(self.Sb, self.Snb), (self.Sh.hi, self.Snh.hi), (self.Sh.lo, self.Snh.lo),
                         (self.Sh.scale, self.Snh.scale), (self.Lb, self.Lnb), (self.Lh.hi, self.Lnh.hi),
                         (self.Lh.lo, self.Lnh.lo), (self.Lh.scale, self.Lnh.scale), (self.mu, self.mun)]
            else:
                pairs = [(self.Sb, self.Snb), (self.Shi, self.Snhi), (self.Slo, self.Snlo), (self.Lb, self.Lnb),
                         (self.Lhi, self.Lnhi), (self.Llo, self.Lnlo), (self.mu, self.mun)]
            plan = self._plans[self.parity] = L.CommitPlan(pairs)
        return plan

    def _flip(self):
        """Exchange the roles of the (current, proposal) buffers - unconditionally: a rejected proposal has already been
        overwritten with the previous state on the device (gsm.py:125-129)."""
        self.Sb, self.Snb, self.S, self.Sn = self.Snb, self.Sb, self.Sn, self.S
        self.Lb, self.Lnb = self.Lnb, self.Lb
        if self.h3:
            self.Sh, self.Snh, self.Lh, self.Lnh = self.Snh, self.Sh, self.Lnh, self.Lh
        else:
            self.Shi, self.Snhi, self.Slo, self.Snlo = self.Snhi, self.Shi, self.Snlo, self.Slo
            self.Lhi, self.Lnhi, self.Llo, self.Lnlo = self.Lnhi, self.Lhi, self.Lnlo, self.Llo
        self.mu, self.mun = self.mun, self.mu
        self.parity = 1 - self.parity
        self.cur = 1 - self.cur

    def _accepted(self):
        """sync_accept only: read the status word of the step just queued (4 bytes; the step's only host read)."""
        if not self.sync_accept:
            return True
        self.status_host.copy_(self.status, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(self.status_host[1]) == 1

    @property
    def n_reverts(self):
        """Rejected updates so far (device counter; reading it synchronises)."""
        return int(self.status[0].item())

    def _launch_body_h3(self, i, graph):
        """Every launch of one h3 iteration, the device-side accept / revert and the NEXT iteration's Philox draws
        included.  graph=True: the form that is captured into a CUDA graph (Philox counter read from device memory and
        advanced there, no phase stamps)."""
        D, B = self.D, self.B
        sl = self.slots
        tm = (lambda name: None) if graph else self._phase_mark
        tm("start")
        sl.zero_()
        # ---- sample (gsm.py:117-119)
        if self.z_tape is not None:
            self.Zh.split_from(self._tape_slice(i))
            if self._tape_async:
                self.z_free[self._z_buf_in_use].record()
        elif not graph and self.z_drawn_for != i:
            L.philox_normal_h3(self.Zh, B, D, self.seed, i * self.world + self.rank)
        tgt = self.target
        tm("draw")
        L.sample_h3(self.mu, self.Lh, self.Zh, self.Xb, sl[0:1], B, D)
        tm("sample")
        # The NEXT iteration's Philox draws depend on nothing but the counter and on the sampler having read this
        # iteration's: they run on the second stream beside the score / W GEMMs (ALU work under tensor-core work, a few
        # small CTAs next to the GEMM's one big CTA per SM) instead of beside the Cholesky, whose latency chain they slowed
        # by more than half of what they hid (GSMVI_SIDE_PHILOX=chol keeps the old placement)
        early = None
        if not graph and self._side_ok and self._philox_early and self.z_tape is None:
            early = self._side()
            self._ev_fork2.record()
            with torch.cuda.stream(early):
                early.wait_event(self._ev_fork2)
                L.philox_normal_h3(self.Zh, B, D, self.seed, (i + 1) * self.world + self.rank)
                self.z_drawn_for = i + 1
        if tgt is not None:
            # (the GEMMs can also write their result's fp16 split themselves with an a-priori bound as scale -
            # gsmvi_h3_bound_scales / X_split, G_split - but the strided 8-byte stores lengthen the un-overlapped epilogue
            # by 60-75 us per launch, twice what the separate HBM-bound split pass costs: measured, not used)
            self.Xh.split_from(self.X, absmax=sl[0:1])
            L.gauss_score_h3(self.Xh, tgt.Ph, tgt.c, self.Gb, sl[1:2], B, D)
            self.Gh.split_from(self.G, absmax=sl[1:2])
        else:
            if self.score_input == "numpy":
                self.G.copy_(to_dev(self.lp_g(self.X.cpu().numpy()), self.dev))
            else:
                self.G.copy_(to_dev(self.lp_g(self.X), self.dev))
            self.Gh.split_from(self.G)
        tm("score+splits")
        # ---- update (gsm.py:122)
        if self.world == 1:
            L.gsm_update_h3(self.Xb, self.Gb, self.Gh, self.mu, self.Sb, self.Sh, self.mun, self.Snb, sl[2:3], B, D, B, 0,
                            self.ws_u)
        else:
            # the exchange is fused into the update: partial tiles pushed to their owner rank from the GEMM epilogue,
            # reduced there in rank order and stored into every rank's next Sigma buffer (bit-identical replicas)
            self.comm.update_fused(self.Xb, self.Gb, self.Gh, self.mu, self.Sh, self.mun, self.cur, B, D, self.batch_size,
                                   self.ws_u)
            L.h3_absmax(self.Sn, D, D, sl[2:3])
        tm("update(+exchange)")
        # ---- goodness check = Cholesky of the new covariance, reused as the next sampling factor (gsm.py:125).  The
        # factorisation is a latency chain that leaves most of every SM idle, so the two HBM-bound passes that depend on
        # nothing it produces - the fp16 split of the new covariance and the NEXT iteration's Philox draws (this iteration's
        # were consumed by the sampler) - run beside it on a second stream (eager launches only; ~70 us of a 2.26 ms step)
        side = None if (graph or not self._side_ok) else self._side()
        if side is not None:
            self._ev_fork.record()
            with torch.cuda.stream(side):
                side.wait_event(self._ev_fork)
                self.Snh.split_from(self.Sn, absmax=sl[2:3])
                if self.z_tape is None and early is None:
                    L.philox_normal_h3(self.Zh, B, D, self.seed, (i + 1) * self.world + self.rank)
                    self.z_drawn_for = i + 1
                self._ev_join.record()
        L.potrf_h3(self.Snb, self.Lnb, self.Lnh, D, self.bad, self.ws_p, zero_upper=False)
        tm("cholesky")
        if side is not None:
            torch.cuda.current_stream().wait_event(self._ev_join)
        else:
            self.Snh.split_from(self.Sn, absmax=sl[2:3])
        tm("split Sigma")
        # ---- accept / revert on the device (gsm.py:125-129); then (single-stream form) the NEXT iteration's draws: they
        # depend on nothing but the counter, so the GPU never idles between iterations
        L.gsm_commit(self.bad, self._commit_plan(), self.status)
        if graph:
            L.philox_normal_h3(self.Zh, B, D, self.seed, 0, offset_dev=self.ctr)
            self.ctr.add_(self.world)
        elif self.z_tape is None and side is None:
            L.philox_normal_h3(self.Zh, B, D, self.seed, (i + 1) * self.world + self.rank)
            self.z_drawn_for = i + 1
        tm("commit+draw")

    def _side(self):
        """Second stream (and its fork / join events) for the passes that run beside the Cholesky."""
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
            self._ev_fork, self._ev_join, self._ev_fork2 = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        return self._side_stream

    def step_h3(self, i):
        """One iteration on the scaled 3xFP16 engine (same sequence as `step`).  With the built-in target, Philox draws and
        one GPU and a launch-bound size (D <= 1024), the launches of a step are captured once per buffer parity into a CUDA
        graph and replayed (GSMVI_GRAPH=0 keeps the eager path, =1 forces the graph at any size)."""
        use_graph = self._graph_ok and self._steps_done >= 3 and self.z_drawn_for == i
        if use_graph:
            g = self._graphs.get(self.parity)
            if self._ctr_next != i + 1:
                self.ctr.fill_((i + 1) * self.world + self.rank)
            if g is None:
                try:
                    self._commit_plan()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch_body_h3(i, True)
                    self._graphs[self.parity] = g
                except Exception as exc:  # capture not possible on this driver: stay eager
                    import warnings
                    warnings.warn("gsmvi_b200: CUDA graph capture failed (%s); continuing with eager launches" % exc)
                    self._graph_ok, use_graph, g = False, False, None
                    torch.cuda.synchronize()
            if g is not None:
                g.replay()
                self._ctr_next = i + 2
                self.z_drawn_for = i + 1
        if not use_graph:
            self._launch_body_h3(i, False)
        self._steps_done += 1
        self._flip()
        return self._accepted()

    def _phase_mark(self, name):
        """GSMVI_PHASE_TIMING=1: CUDA-event stamps between the phases of a step; averages printed by close()."""
        if not self._phase_on:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._phase_events.append((name, ev))

    def _phase_report(self):
        if not self._phase_on or not self._phase_events:
            return
        torch.cuda.synchronize()
        tot, cnt, order = {}, {}, []
        prev = None
        for name, ev in self._phase_events:
            if name != "start" and prev is not None:
                tot[name] = tot.get(name, 0.0) + prev.elapsed_time(ev)
                cnt[name] = cnt.get(name, 0) + 1
                if name not in order:
                    order.append(name)
            prev = ev
        if self.rank == 0:
            print("[gsmvi phase timing, rank 0, ms per step] " + "  ".join("%s %.3f" % (n, tot[n] / cnt[n]) for n in order),
                  file=__import__("sys").stderr, flush=True)

    def end_fit(self):
        """End of one fit on this engine (the engine stays usable): phase report, tape released."""
        self._phase_report()
        self._phase_events = []
        self.z_tape = None

    def close(self, collective=True):
        """Release the peer-mapped exchange buffer.  collective=True: every rank must call it (barriers order the unmapping
        and the free); collective=False is the emergency / interpreter-exit form that touches nothing but this rank."""
        self.end_fit()
        self._graphs = {}
        if self.comm is not None:
            self.comm.close(collective=collective)
            self.comm = None

    def step(self, i):
        if self.h3:
            return self.step_h3(i)
        D, B, npass = self.D, self.B, self.npass
        # ---- sample (gsm.py:117-119)
        if self.z_tape is not None:
            self.Z.copy_(self.z_tape[i, self._tape_off:self._tape_off + B], non_blocking=True)
        else:
            L.philox_normal(self.Zb, B, D, self.seed, i * self.world + self.rank)
        L.sample(self.mu, self.Lhi, self.Zb, self.Xb, B, D, npass, L_lo=self.Llo)
        # ---- score (gsm.py:121)
        if self.target is not None:
            L.gauss_score(self.Xb, self.target.Phib, self.target.c, self.Gb, B, D, npass, P_lo=self.target.Plob)
        elif self.score_input == "numpy":
            self.G.copy_(to_dev(self.lp_g(self.X.cpu().numpy()), self.dev))
        else:
            self.G.copy_(to_dev(self.lp_g(self.X), self.dev))
        # ---- update (gsm.py:122)
        if self.world == 1:
            L.gsm_update_raw(self.Xb, self.Gb, self.mu, self.Sb, self.mun, self.Snb, B, D, B, 0, self.ws_u, npass,
                             Sigma_hi=self.Shi, Sigma_lo=self.Slo)
        else:
            L.gsm_update_raw(self.Xb, self.Gb, self.mu, self.Sb, self.dmu, self.dSb, B, D, self.batch_size, 1,
                             self.ws_u, npass, Sigma_hi=self.Shi, Sigma_lo=self.Slo)
            self.dist.all_reduce(self.dSb, group=self.group)
            self.dist.all_reduce(self.dmu, group=self.group)
            L.gsm_apply_stats(self.Sb, self.dSb, self.mu, self.dmu, self.Snb, self.mun, D)
        # ---- goodness check = Cholesky of the new covariance, reused as the next sampling factor (gsm.py:125)
        L.potrf_check(self.Snb, self.Lnb, D, self.bad, self.ws_p, npass)
        L.tf32_split(self.Snb, self.Snhi, self.Snlo, D, D)
        L.tf32_split(self.Lnb, self.Lnhi, self.Lnlo, D, D)
        L.gsm_commit(self.bad, self._commit_plan(), self.status)  # gsm.py:125-129 on the device
        self._flip()
        return self._accepted()

    def mean(self):
        return self.mu[: self.D]

    def cov(self):
        return self.S

    def chol_buffer(self):
        """Padded fp32 buffer of the lower Cholesky factor of cov() (zeros above the diagonal): after a step the proposal's
        factor if it was accepted, the previous state's if not - the device-side commit keeps (mu, Sigma, L) together."""
        return self.Lb


class GSMSmall64Engine:
    """fp64 engine for D <= 64 (csrc/gsm_small64.cu): the reference's numpy path (gsmvi/gsm_numpy.py:60-129, BASELINE
    configs[0]) is fp64 end to end, and its example target (condition number ~3e4) is out of fp32's reach at the 1e-4
    bar.  State (mu, Sigma, L) lives on the device as doubles; sample, score, update, Cholesky check and the accept /
    revert all run inside one CTA, the commit predicated on the device, so `run(i0, n)` is a single launch for n
    iterations with the built-in target and two launches per iteration around a user callable."""

    def __init__(self, D, batch_size, lp_g, key, mean=None, cov=None, z_tape=None, score_input="torch"):
        dev = self.dev = device()
        self.D, self.B, self.batch_size, self.lp_g = D, batch_size, batch_size, lp_g
        self.seed = key_to_seed(key)
        self.score_input = score_input
        f64 = dict(dtype=torch.float64, device=dev)
        self.mu = torch.zeros(D, **f64)
        if mean is not None:
            self.mu.copy_(_to_dev64(mean, dev))  # gsm.py:100-101
        self.S = torch.eye(D, **f64) if cov is None else _to_dev64(cov, dev).clone().contiguous()  # gsm.py:102-103
        self.Lf = torch.zeros(D, D, **f64)
        self.X = torch.zeros(batch_size, D, **f64)
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)
        self.ws = torch.empty(L.gsm_small64_workspace_bytes(batch_size, D) // 8, **f64)
        if z_tape is not None:
            z_tape = torch.as_tensor(np.asarray(z_tape) if not isinstance(z_tape, torch.Tensor) else z_tape)
            assert z_tape.shape[1] == batch_size and z_tape.shape[2] == D
            z_tape = z_tape.to(device=dev, dtype=torch.float32).contiguous()
        self.z_tape = z_tape
        self.target = getattr(getattr(lp_g, "__self__", None), "_gsmvi_builtin_target", None)
        if self.target is not None:
            self.P64, self.c64 = self.target.device_fp64()
        self._reverts_seen = 0
        L.gsm_small64(L.SMALL64_INIT, self.mu, self.S, self.Lf, None, 0, 0, None, None, None, None, batch_size, D, 0,
                      self.status, self.ws)
        if int(self.status[1].item()) != 0:
            raise ValueError("initial covariance is not positive definite")

    def _tape(self, i0, n):
        return None if self.z_tape is None else self.z_tape[i0:i0 + n]

    def run(self, i0, n):
        """Iterations i0 .. i0 + n - 1 (gsm.py:116-129); nothing is read back."""
        B, D = self.B, self.D
        if self.target is not None:
            L.gsm_small64(L.SMALL64_FULL, self.mu, self.S, self.Lf, self._tape(i0, n), self.seed, i0, self.X, None,
                          self.P64, self.c64, B, D, n, self.status, self.ws)
            return
        for i in range(i0, i0 + n):
            L.gsm_small64(L.SMALL64_SAMPLE, self.mu, self.S, self.Lf, self._tape(i, 1), self.seed, i, self.X, None, None,
                          None, B, D, 1, self.status, self.ws)
            if self.score_input == "numpy":  # reference-style callable on host fp64 arrays (gsm_numpy.py:120)
                G = torch.as_tensor(np.asarray(self.lp_g(self.X.cpu().numpy()), dtype=np.float64)).to(self.dev)
            elif self.score_input == "torch64":
                G = self.lp_g(self.X).detach().to(device=self.dev, dtype=torch.float64)
            else:
                G = self.lp_g(self.X.to(torch.float32)).detach().to(device=self.dev, dtype=torch.float64)
            G = G.reshape(B, D).contiguous()
            L.gsm_small64(L.SMALL64_UPDATE, self.mu, self.S, self.Lf, None, self.seed, i, self.X, G, None, None, B, D, 1,
                          self.status, self.ws)

    def step(self, i):
        """One iteration, returning whether the update was accepted (reads the device status word: the verbose path)."""
        self.run(i, 1)
        return int(self.status[2].item()) == 1

    @property
    def n_reverts(self):
        return int(self.status[0].item())

    def mean(self):
        return self.mu

    def cov(self):
        return self.S

    def close(self):
        pass


def _to_dev64(a, dev):
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=dev, dtype=torch.float64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dev)


# ---- engines (workspaces, peer-mapped exchange buffers, captured graphs) are kept across fit calls: setting one up costs
# ~170 MB of cudaMalloc + zeroing at D = 4096, and on a sharded fit an IPC handle exchange, 7 cudaIpcOpenMemHandle and three
# barriers (~95 ms at 8 ranks) - more than the 20-iteration fits of the benchmark themselves.
_ENGINES = __import__("collections").OrderedDict()
_ENGINES_MAX = 2


def _engine_key(kind, D, batch_size, npass, group):
    return (kind, D, batch_size, npass, None if group is None else id(group), torch.cuda.current_device())


def _evict(key, collective):
    eng = _ENGINES.pop(key, None)
    if eng is not None:
        eng.close(collective=collective)


def release_engines(collective=True):
    """Free every cached engine (collective=True: call it on every rank of a sharded fit's group, before the process group
    is destroyed; the interpreter-exit hook uses the non-collective form)."""
    for key in list(_ENGINES):
        _evict(key, collective)


# (no interpreter-exit hook: process teardown releases device memory and IPC mappings, and CUDA calls from an atexit
# handler can outlive the context)


class GSM:
    """Wrapper class for using GSM updates to fit a distribution (gsmvi/gsm.py:62-76)."""

    def __init__(self, D, lp, lp_g):
        """D: number of parameters; lp: target log-probability (used only by the monitor); lp_g: score function,
        called as lp_g(samples[B, D]) -> [B, D] on CUDA tensors.  A `targets.DenseGaussianTarget.lp_g` bound method is
        recognised and evaluated by the built-in score GEMM instead of being called."""
        self.D = D
        self.lp = lp
        self.lp_g = lp_g

    def fit(self, key, mean=None, cov=None, batch_size=2, niter=5000, nprint=10, verbose=True, check_goodness=True,
            monitor=None, *, z_tape=None, npass=None, process_group=None, score_input="torch"):
        """Main function to fit a multivariate Gaussian to the target (gsmvi/gsm.py:79-133).

        Reference arguments keep their meaning (check_goodness is accepted and, as in the reference, the covariance
        is always checked; unlike the reference, niter < nprint does not raise ZeroDivisionError).  Extra keyword-only
        arguments:
          z_tape: optional [niter+1, batch_size, D] standard-normal draws used instead of the Philox stream
                  (parity runs: the same tape is fed to the oracle); with a process_group a
                  rank may pass only its own [niter+1, batch_size / world, D] slice
          npass: arithmetic: None (default) = 0 for D <= 64 on one GPU, else 4;
                  0 = fp64 single-CTA path (D <= 64; what the reference's numpy example computes in, gsm_numpy.py);
                  4 = scaled 3xFP16 tensor-core split (22-bit significands like 3xTF32 at twice the pipe rate),
                  3 = 3xTF32 round-to-nearest split, 2 = 3xTF32 truncation split, 1 = single TF32 pass
          process_group: torch.distributed group; the batch is sharded across its ranks and the D x D statistics are
                  exchanged over NVLink peer memory (one process per GPU)
          score_input: "torch" passes CUDA fp32 tensors to lp_g; "numpy" passes host arrays (reference-style callables;
                  fp64 on the npass = 0 path); "torch64" passes CUDA fp64 tensors (npass = 0 path)
        Returns (mean[D], cov[D, D]) as CUDA tensors (fp32; fp64 on the npass = 0 path)."""
        D = self.D
        if npass is None:
            npass = 0 if (D <= 64 and process_group is None) else 4
        if npass == 0:
            if D > 64 or process_group is not None:
                raise ValueError("npass = 0 (fp64 single-CTA path) needs D <= 64 and a single GPU")
            eng = GSMSmall64Engine(D, batch_size, self.lp_g, key, mean, cov, z_tape, score_input)
            ekey = None
        else:
            ekey = _engine_key("gsm", D, batch_size, npass, process_group)
            eng = _ENGINES.pop(ekey, None)
            if eng is None:
                while len(_ENGINES) >= _ENGINES_MAX:  # same call sequence on every rank => same evictions: collective
                    _evict(next(iter(_ENGINES)), collective=True)
                eng = GSMEngine(D, batch_size, self.lp_g, key, mean, cov, z_tape, npass, process_group, score_input,
                                sync_accept=bool(verbose))
            else:
                try:
                    eng.reset(self.lp_g, key, mean, cov, z_tape, score_input, sync_accept=bool(verbose))
                except Exception:
                    _ENGINES[ekey] = eng  # the buffers are intact; only this start was rejected
                    raise
        if z_tape is not None:
            assert eng.z_tape.shape[0] >= niter + 1
        chunked = npass == 0 and eng.target is not None and not verbose
        try:
            nevals = 1  # gsm.py:105
            every = max(niter // max(nprint, 1), 1)
            i = 0
            while i <= niter:  # gsm.py:107: niter + 1 updates
                if verbose and (i % every == 0):  # gsm.py:108-109
                    print(f"Iteration {i} of {niter}")
                if monitor is not None and (i % monitor.checkpoint) == 0:  # gsm.py:111-114
                    monitor(i, _monitor_params(eng), self.lp, key, nevals=nevals)
                    nevals = 0
                if chunked:
                    # fp64 path with the built-in target: every iteration up to the next monitor checkpoint in ONE launch
                    n = niter + 1 - i
                    if monitor is not None:
                        n = min(n, monitor.checkpoint - (i % monitor.checkpoint))
                    eng.run(i, n)
                    nevals += batch_size * n
                    i += n
                    continue
                ok = eng.step(i)
                nevals += batch_size  # gsm.py:123
                if not ok and verbose:
                    print("Bad update for covariance matrix. Revert")  # gsm.py:128-129
                i += 1
            i = niter
            if monitor is not None:  # gsm.py:131-132
                monitor(i, _monitor_params(eng), self.lp, key, nevals=nevals)
            self.n_reverts = eng.n_reverts
            mean, cov = eng.mean().clone(), eng.cov().clone()
        except BaseException:
            # a rank that failed mid-fit must not leave its peers' buffers mapped into a dead fit, nor wait in a barrier
            # for them: tear this rank's engine down without any collective and let the exception travel
            eng.close(collective=False) if ekey is not None else None
            raise
        if ekey is not None:
            eng.end_fit()
            _ENGINES[ekey] = eng
        return mean, cov
