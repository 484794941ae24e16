"""GSM: Gaussian score matching VI on B200 - drop-in for gsmvi/gsm.py (and gsmvi/gsm_numpy.py) of modichirag/GSM-VI.

Same public surface: `gsm_update(samples, vs, mu0, S0) -> (mu, S)` (gsmvi/gsm.py:31-58) and
`GSM(D, lp, lp_g).fit(key, mean, cov, batch_size, niter, nprint, verbose, check_goodness, monitor) -> (mean, cov)`
(gsmvi/gsm.py:62-133).  The loop body runs as sm_100a kernels through libgsmvi_b200.so: blocked Cholesky (PD check +
sampling factor), X = mu + Z L^T, score, fused GSM update.  Tensors are torch.float32 on the current CUDA device.
"""
import torch

from . import _lib as L
from ._util import device, key_to_seed, ld_of, new_mat, new_vec, to_dev


def gsm_update(samples, vs, mu0, S0, npass=3):
    """Drop-in for gsmvi/gsm.py:31-58.  samples, vs: [B, D]; mu0: [D]; S0: [D, D] (symmetric).  Returns (mu, S) as
    new CUDA tensors."""
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
        return draw + score + upd + potrf + (3 if world > 1 else 0)
    potrf = 1 + panels + (panels - 1)  # tril copy, panel kernels, SYRK GEMMs
    upd = 4 + 2  # W GEMM, row pass, covariance GEMM, axpy + the two tf32_split launches (Sigma_new, L_new)
    return (0 if tape else 1) + 1 + (1 if builtin_target else 0) + upd + potrf + (2 if world > 1 else 0)


class GSMEngine:
    """Device-resident state and workspaces of one GSM fit; `step(i)` is one loop body of gsmvi/gsm.py:107-129
    (sample -> score -> update -> goodness check -> accept/revert).  GSM.fit drives it; bench.py times it."""

    def __init__(self, D, batch_size, lp_g, key, mean=None, cov=None, z_tape=None, npass=4, process_group=None,
                 score_input="torch"):
        dev = self.dev = device()
        self.D, self.batch_size, self.lp_g, self.npass = D, batch_size, lp_g, npass
        self.group, self.dist, self.rank, self.world = process_group, None, 0, 1
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        if batch_size % self.world != 0:
            raise ValueError("batch_size must be divisible by the number of ranks")
        B = self.B = batch_size // self.world
        self.seed = key_to_seed(key)
        self.score_input = score_input
        self._phase_on = bool(__import__("os").environ.get("GSMVI_PHASE_TIMING"))
        self._phase_events = []
        self.h3 = npass == 4
        # state, double-buffered so a rejected update is simply not swapped in
        self.comm = None
        if self.h3 and self.world > 1:
            # batch-sharded fit on the h3 engine: the two Sigma buffers live in this rank's peer-mapped exchange buffer and
            # the statistics travel through NVLink peer memory, fused with the covariance GEMM (csrc/comm.cu)
            from ._comm import CommBuffer
            self.comm = CommBuffer(D, process_group, self.dist)
            self.Sb, self.Snb = self.comm.S[0], self.comm.S[1]
            self.S, self.Sn = self.Sb[:, :D], self.Snb[:, :D]
            self.cur = 0
        else:
            self.Sb, self.S = new_mat(D, D, dev)
            self.Snb, self.Sn = new_mat(D, D, dev)
        self.Lb, _ = new_mat(D, D, dev)
        self.Lnb, _ = new_mat(D, D, dev)
        if not self.h3:
            # pre-split low parts of the reused GEMM operands (Sigma for W = G Sigma, L for the sampler), kept per buffer
            self.Shi, self.Slo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Snhi, self.Snlo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Lhi, self.Llo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
            self.Lnhi, self.Lnlo = new_mat(D, D, dev)[0], new_mat(D, D, dev)[0]
        self.mu, self.mun = new_vec(D, dev), new_vec(D, dev)
        if mean is not None:
            self.mu[:D].copy_(to_dev(mean, dev))  # gsm.py:100-101 (default zeros)
        if cov is None:
            self.S.copy_(torch.eye(D, device=dev))  # gsm.py:102-103
        else:
            self.S.copy_(to_dev(cov, dev))
        self.Zb, self.Z = new_mat(B, D, dev)
        self.Xb, self.X = new_mat(B, D, dev)
        self.Gb, self.G = new_mat(B, D, dev)
        if not self.h3:
            self.ws_p = torch.empty(L.workspace_bytes(L.WS_POTRF, B, D) // 4, dtype=torch.float32, device=dev)
            self.ws_u = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE, B, D) // 4, dtype=torch.float32, device=dev)
        self.bad = torch.zeros(1, dtype=torch.int32, device=dev)
        if self.world > 1 and self.comm is None:
            self.dSb, _ = new_mat(D, D, dev)
            self.dmu = new_vec(D, dev)
        if z_tape is not None and not isinstance(z_tape, torch.Tensor):
            z_tape = torch.as_tensor(z_tape, dtype=torch.float32)
        self._tape_off = self.rank * B
        if z_tape is not None:
            # [n, batch_size, D]: the global draws, every rank reads its slice; [n, batch_size / world, D] on a sharded
            # fit: this rank's own draws (a rank then only holds - and pins - what it copies)
            if self.world > 1 and z_tape.shape[1] == B:
                self._tape_off = 0
            else:
                assert z_tape.shape[1] == batch_size
            assert z_tape.shape[2] == D
        self.z_tape = z_tape
        self._tape_async = bool(z_tape is not None and npass == 4 and not z_tape.is_cuda and z_tape.is_pinned())
        if self._tape_async:
            self.copy_stream = torch.cuda.Stream()
            self.Zbufs = [self.Z, new_mat(B, D, dev)[1]]
            self.z_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self.z_free = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self.z_free:
                e.record()
            self._tape_copied, self._z_buf_in_use = -1, 0
        self.target = getattr(getattr(lp_g, "__self__", None), "_gsmvi_builtin_target", None)
        self.n_reverts = 0
        if self.h3:
            # scaled 3xFP16 engine: every GEMM operand lives as an fp16 (hi, lo) pair + power-of-two scale
            H = L.HOperand
            self.Sh, self.Snh, self.Lh, self.Lnh = H(D, D, dev), H(D, D, dev), H(D, D, dev), H(D, D, dev)
            self.Zh, self.Xh, self.Gh = H(B, D, dev), H(B, D, dev), H(B, D, dev)
            self.slots = torch.zeros(8, dtype=torch.int32, device=dev)  # |X|, |G|, |Sigma_new| maxima (bit patterns)
            self.bad_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self.flag_event = torch.cuda.Event()
            self.z_drawn_for = -1
            self.parity, self._steps_done, self._graphs, self._ctr_next = 0, 0, {}, -1
            self.ctr = torch.zeros(1, dtype=torch.int64, device=dev)  # Philox counter of the next draw (graph mode)
            # launch-bound sizes only: at D = 512 the replayed step is 29% faster (4378 vs 3393 it/s), at D = 4096 the
            # graph's kernel nodes lose the programmatic-dependent-launch overlap of the Cholesky chain and it is slower
            genv = __import__("os").environ.get("GSMVI_GRAPH", "auto")
            self._graph_ok = (self.world == 1 and z_tape is None and genv != "0" and (D <= 1024 or genv == "1")
                              and getattr(getattr(lp_g, "__self__", None), "_gsmvi_builtin_target", None) is not None)
            self.ws_u = torch.empty(L.workspace_bytes(L.WS_GSM_UPDATE_H3, B, D) // 4, dtype=torch.float32, device=dev)
            self.ws_p = torch.empty(L.workspace_bytes(L.WS_POTRF_H3, B, D) // 4, dtype=torch.float32, device=dev)
            L.potrf_h3(self.Sb, self.Lb, self.Lh, D, self.bad, self.ws_p, zero_upper=False)  # buffers start zeroed
            self.Sh.split_from(self.S)
        else:
            L.potrf_check(self.Sb, self.Lb, D, self.bad, self.ws_p, npass)
            L.tf32_split(self.Sb, self.Shi, self.Slo, D, D)
            L.tf32_split(self.Lb, self.Lhi, self.Llo, D, D)
        if int(self.bad.item()) != 0:
            raise ValueError("initial covariance is not positive definite")

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

    def _launch_body_h3(self, i, graph):
        """Every launch of one h3 iteration up to (and including) the copy of the accept flag to pinned host memory and the
        NEXT iteration's Philox draws.  graph=True: the form that is captured into a CUDA graph (Philox counter read from
        device memory and advanced there, no phase stamps)."""
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
        # ---- goodness check = Cholesky of the new covariance, reused as the next sampling factor (gsm.py:125)
        L.potrf_h3(self.Snb, self.Lnb, self.Lnh, D, self.bad, self.ws_p, zero_upper=False)
        tm("cholesky")
        self.Snh.split_from(self.Sn, absmax=sl[2:3])
        tm("split Sigma")
        # the step's only device->host read (4 bytes): copy the flag, then queue the NEXT iteration's draws (they depend
        # on nothing but the counter) so the GPU has work while the host waits for the flag and issues the next launches
        self.bad_host.copy_(self.bad, non_blocking=True)
        if graph:
            L.philox_normal_h3(self.Zh, B, D, self.seed, 0, offset_dev=self.ctr)
            self.ctr.add_(self.world)
        else:
            self.flag_event.record()
            if self.z_tape is None:
                L.philox_normal_h3(self.Zh, B, D, self.seed, (i + 1) * self.world + self.rank)
                self.z_drawn_for = i + 1

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
                self.flag_event.record()
                self._ctr_next = i + 2
                self.z_drawn_for = i + 1
        if not use_graph:
            self._launch_body_h3(i, False)
        self._steps_done += 1
        self.flag_event.synchronize()
        ok = int(self.bad_host[0]) == 0
        if ok:  # gsm.py:126-127
            self.Sb, self.Snb, self.S, self.Sn = self.Snb, self.Sb, self.Sn, self.S
            self.Lb, self.Lnb = self.Lnb, self.Lb
            self.Sh, self.Snh, self.Lh, self.Lnh = self.Snh, self.Sh, self.Lnh, self.Lh
            self.mu, self.mun = self.mun, self.mu
            self.parity = 1 - self.parity
            if self.comm is not None:
                self.cur = 1 - self.cur
        else:
            self.n_reverts += 1
        return ok

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

    def close(self):
        """Release the peer-mapped exchange buffer (collective: every rank must call it)."""
        self._phase_report()
        self._phase_events = []
        if self.comm is not None:
            self.comm.close()
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
        L.tf32_split(self.Snb, self.Snhi, self.Snlo, D, D)  # queued before the flag is read: overlap the host round trip
        L.tf32_split(self.Lnb, self.Lnhi, self.Lnlo, D, D)
        ok = int(self.bad.item()) == 0  # the step's only device->host read (4 bytes)
        if ok:  # gsm.py:126-127
            self.Sb, self.Snb, self.S, self.Sn = self.Snb, self.Sb, self.Sn, self.S
            self.Lb, self.Lnb = self.Lnb, self.Lb
            self.Shi, self.Snhi, self.Slo, self.Snlo = self.Snhi, self.Shi, self.Snlo, self.Slo
            self.Lhi, self.Lnhi, self.Llo, self.Lnlo = self.Lnhi, self.Lhi, self.Lnlo, self.Llo
            self.mu, self.mun = self.mun, self.mu
        else:
            self.n_reverts += 1
        return ok

    def mean(self):
        return self.mu[: self.D]

    def cov(self):
        return self.S


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
            monitor=None, *, z_tape=None, npass=4, process_group=None, score_input="torch"):
        """Main function to fit a multivariate Gaussian to the target (gsmvi/gsm.py:79-133).

        Reference arguments keep their meaning (check_goodness is accepted and, as in the reference, the covariance
        is always checked; unlike the reference, niter < nprint does not raise ZeroDivisionError).  Extra keyword-only
        arguments:
          z_tape: optional [niter+1, batch_size, D] standard-normal draws used instead of the Philox stream
                  (parity runs: the same tape is fed to the oracle); with a process_group a
                  rank may pass only its own [niter+1, batch_size / world, D] slice
          npass: tensor-core precision: 4 = scaled 3xFP16 split (default; 22-bit significands like 3xTF32 at twice the
                  pipe rate), 3 = 3xTF32 round-to-nearest split, 2 = 3xTF32 truncation split, 1 = single TF32 pass
          process_group: torch.distributed group; the batch is sharded across its ranks and the D x D statistics are
                  all-reduced (one process per GPU)
          score_input: "torch" passes CUDA tensors to lp_g; "numpy" passes host arrays (reference-style callables)
        Returns (mean[D], cov[D, D]) as CUDA tensors."""
        eng = GSMEngine(self.D, batch_size, self.lp_g, key, mean, cov, z_tape, npass, process_group, score_input)
        if z_tape is not None:
            assert eng.z_tape.shape[0] >= niter + 1
        nevals = 1  # gsm.py:105
        every = max(niter // max(nprint, 1), 1)
        i = 0
        for i in range(niter + 1):  # gsm.py:107
            if verbose and (i % every == 0):  # gsm.py:108-109
                print(f"Iteration {i} of {niter}")
            if monitor is not None and (i % monitor.checkpoint) == 0:  # gsm.py:111-114
                monitor(i, [eng.mean(), eng.cov()], self.lp, key, nevals=nevals)
                nevals = 0
            ok = eng.step(i)
            nevals += batch_size  # gsm.py:123
            if not ok and verbose:
                print("Bad update for covariance matrix. Revert")  # gsm.py:128-129
        if monitor is not None:  # gsm.py:131-132
            monitor(i, [eng.mean(), eng.cov()], self.lp, key, nevals=nevals)
        self.n_reverts = eng.n_reverts
        mean, cov = eng.mean().clone(), eng.cov().clone()
        eng.close()
        return mean, cov
