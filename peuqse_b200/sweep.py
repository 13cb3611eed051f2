"""Host driver of the multistart sweeps: three kernel passes + (multi-GPU only) three tiny collectives.

pass 1  logP of every point of this rank's index range (optionally kept in HBM) + first strict maximum
        -> all-gather (max, index)            [doListOfPermutationsSearch loop, IP.py:844-956]
pass 2  weights int(exp(logP - thr)), moments about the MAP point, min/max of the retained points
        -> all-reduce                         [convertPermutationsToSamples + calculatePostBurnInStatistics,
                                               IP.py:3463-3491, 1923-1931]
pass 3  weighted marginal histograms          [plotting_functions.py:135]
The best point, the threshold and the bin edges stay on the device between passes, so a whole multistart
is one stream-ordered sequence with a single device->host read of the packed result at the end.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import distributed as D


def grid_size(radii):
    n = 1
    for r in radii:
        n *= 2 * int(r) + 1
    return n


def make_spec(kind, k, center, interval=None, radii=None, spread=0.0, std_prior=None, seed=0):
    s = _lib.SweepSpec()
    s.kind = 0 if kind == "grid" else 1
    s.k = k
    for d in range(k):
        s.center[d] = float(center[d])
        s.interval[d] = float(interval[d]) if interval is not None else 0.0
        s.radii[d] = int(radii[d]) if radii is not None else 0
        s.std_prior[d] = float(std_prior[d]) if std_prior is not None else 0.0
    s.spread = float(spread)
    s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return s


class SweepResult:
    pass


def interleaved_ranges(total, rank, size, pieces):
    """Index ranges of one rank when [0, total) is cut into size*pieces contiguous pieces dealt round-robin.
    combinationGenerator lists a grid from the centre outwards, so the retained (near-MAP) points of a grid sweep sit in
    a prefix of the index space: one contiguous range per rank would leave all of pass 2-3 to rank 0."""
    P = size * pieces
    out = []
    for j in range(rank, P, size):
        a, b = (total * j) // P, (total * (j + 1)) // P
        if b > a:
            out.append((a, b - a))
    return out or [(0, 0)]


class Sweep:
    def __init__(self, dp, spec, total, store_logp=True, nbins=10, shard=None, pieces=None, local=False):
        """shard = (rank, size): take that rank's range instead of this process's; local = True: no exchange with the other
        ranks (with shard = (0, 1): the whole sweep on this GPU alone, e.g. to check a sharded run against)."""
        self.dp = dp
        self.spec = spec
        self.k = k = spec.k
        self.total = int(total)
        self.nbins = int(nbins)
        self.L = _lib.load()
        rank, size = D.world()
        if shard is not None:
            self.ranges = [D.shard_range(self.total, *shard)]
        elif size > 1 and (pieces or 1) > 1:
            self.ranges = interleaved_ranges(self.total, rank, size, int(pieces))
        else:
            self.ranges = [D.shard_range(self.total)]
        self.first = self.ranges[0][0]
        self.count = sum(c for _, c in self.ranges)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.works = [torch.zeros(int(self.L.peuqse_sweep_work_bytes(dp.handle if store_logp else None,
                                                                     C.byref(spec) if store_logp else None, f, c)),
                                  dtype=torch.uint8, device=dev) for f, c in self.ranges]
        self.work = self.works[0]
        self.logp = torch.empty(self.count, dtype=torch.float64, device=dev) if store_logp else None
        offs = np.cumsum([0] + [c for _, c in self.ranges])
        self.logps = [self.logp[offs[i]:offs[i + 1]] if store_logp else None for i in range(len(self.ranges))]
        self.bests = torch.zeros((len(self.ranges), 2), dtype=torch.float64, device=dev) if len(self.ranges) > 1 else None
        # packed result: [best(2) | acc | hist (int64 bits) | edges]
        self.lay = lay = _lib.acc_layout(k)
        nacc = lay["SIZE"]
        self.res = torch.zeros(2 + nacc + k * nbins + k * (nbins + 1) + k, dtype=torch.float64, device=dev)
        self.best = self.res[0:2]                                        # peuqse_best {double logp, int64 index}
        self.acc = self.res[2:2 + nacc]
        self.hist = self.res[2 + nacc:2 + nacc + k * nbins].view(torch.int64)
        self.edges = self.res[2 + nacc + k * nbins:2 + nacc + k * nbins + k * (nbins + 1)]
        self.map_theta = self.res[2 + nacc + k * nbins + k * (nbins + 1):]
        self.acc_init = torch.zeros(nacc, dtype=torch.float64, device=dev)
        self.acc_init[lay["MIN"]:lay["MIN"] + k] = float("inf")
        self.acc_init[lay["MAX"]:lay["MAX"] + k] = float("-inf")
        self.launches = 0
        self.use_graph = True
        self.world = 1 if local else D.world()[1]
        self.peer = D.native_peer() if self.world > 1 else None      # push-wait-merge kernels over peer memory
        self.comm = D.native_comm() if (self.world > 1 and self.peer is None) else None
        self.peer_max = int(self.L.peuqse_peer_max_bytes())

    # -- individual passes: device-resident in and out (bench.py times these for the kernel-only number)
    def pass1(self):
        self.pass1_kernels()
        self.pass1_merge()

    def pass1_kernels(self):
        """The evaluation kernel(s) of pass 1 alone (what bench.py brackets with CUDA events for the roofline)."""
        multi = len(self.ranges) > 1
        for i, (f, c) in enumerate(self.ranges):
            _lib.check(self.L.peuqse_sweep_pass1(self.dp.handle, C.byref(self.spec), f, c, _lib.ptr(self.logps[i]),
                                                 _lib.ptr(self.bests[i]) if multi else _lib.ptr(self.best),
                                                 _lib.ptr(self.works[i]), _lib.current_stream()))
            self.launches += 1

    def pass1_merge(self):
        """This rank's pieces -> one record, the exchange of the best record over the ranks, the MAP point."""
        multi = len(self.ranges) > 1
        if multi:                                           # this rank's pieces -> one record
            _lib.check(self.L.peuqse_merge_best(_lib.ptr(self.bests), len(self.ranges), _lib.ptr(self.best), _lib.current_stream()))
            self.launches += 1
        if self.world > 1:
            self._merge_best()
        _lib.check(self.L.peuqse_sweep_best_point(C.byref(self.spec), _lib.ptr(self.best), _lib.ptr(self.map_theta),
                                                  _lib.current_stream()))
        self.launches += 1

    def _merge_best(self):
        """all-gather of the 16-byte (maxLogP, index) records + one merge kernel (highest value, lowest index among equals)."""
        import torch.distributed as dist
        if self.peer is not None:
            _lib.check(self.L.peuqse_peer_merge_best(self.peer, _lib.ptr(self.best), _lib.current_stream()))
            self.launches += 1
            return
        if not hasattr(self, "_g_best"):
            self._g_best = torch.empty((self.world, 2), dtype=torch.float64, device=self.dev)
            self._g_acc = torch.empty((self.world, self.lay["SIZE"]), dtype=torch.float64, device=self.dev)
        if self.comm is not None:
            _lib.check(self.L.peuqse_comm_allgather(self.comm, _lib.ptr(self.best), _lib.ptr(self._g_best), 16, _lib.current_stream()))
        else:
            dist.all_gather_into_tensor(self._g_best, self.best.reshape(1, 2))
        _lib.check(self.L.peuqse_merge_best(_lib.ptr(self._g_best), self.world, _lib.ptr(self.best), _lib.current_stream()))
        self.launches += 1

    def pass2(self, log_rel, keep=None):
        self.acc.copy_(self.acc_init)
        kidx = klogp = kcount = None
        cap = 0
        if keep is not None:
            kidx, klogp, kcount = keep
            cap = kidx.numel()
            kcount.zero_()
        for i, (f, c) in enumerate(self.ranges):            # sums accumulate into self.acc over the pieces
            _lib.check(self.L.peuqse_sweep_pass2(self.dp.handle, C.byref(self.spec), f, c,
                                                 _lib.ptr(self.logps[i]), _lib.ptr(self.best), float(log_rel), _lib.ptr(self.acc),
                                                 _lib.ptr(kidx), _lib.ptr(klogp), cap, _lib.ptr(kcount), _lib.ptr(self.works[i]),
                                                 _lib.current_stream()))
            self.launches += 2                              # main kernel + finish kernel
        if self.world > 1 and self.peer is not None:
            _lib.check(self.L.peuqse_peer_merge_acc(self.peer, _lib.ptr(self.acc), self.k, _lib.current_stream()))
            self.launches += 1
        elif self.world > 1:                                # one all-gather + one merge kernel (sum / min / max by slot)
            import torch.distributed as dist
            if not hasattr(self, "_g_acc"):
                self._g_acc = torch.empty((self.world, self.lay["SIZE"]), dtype=torch.float64, device=self.dev)
            if self.comm is not None:
                _lib.check(self.L.peuqse_comm_allgather(self.comm, _lib.ptr(self.acc), _lib.ptr(self._g_acc),
                                                        self.lay["SIZE"] * 8, _lib.current_stream()))
            else:
                dist.all_gather_into_tensor(self._g_acc, self.acc.reshape(1, -1))
            _lib.check(self.L.peuqse_merge_acc(_lib.ptr(self._g_acc), self.world, self.k, _lib.ptr(self.acc),
                                               _lib.current_stream()))
            self.launches += 1

    def pass3(self, log_rel, reuse_pass2_sums=True):
        _lib.check(self.L.peuqse_sweep_edges(_lib.ptr(self.acc), self.k, self.nbins, _lib.ptr(self.edges),
                                             _lib.ptr(self.hist), _lib.current_stream()))
        self.launches += 1
        for i, (f, c) in enumerate(self.ranges):
            _lib.check(self.L.peuqse_sweep_hist(self.dp.handle, C.byref(self.spec), f, c,
                                                _lib.ptr(self.logps[i]), _lib.ptr(self.best), float(log_rel), _lib.ptr(self.edges),
                                                self.nbins, _lib.ptr(self.hist), _lib.ptr(self.works[i]), 1 if reuse_pass2_sums else 0,
                                                _lib.current_stream()))
            self.launches += 1
        if self.world > 1 and self.peer is not None and self.hist.numel() * 8 <= self.peer_max:
            _lib.check(self.L.peuqse_peer_sum_i64(self.peer, _lib.ptr(self.hist), self.hist.numel(), _lib.current_stream()))
            self.launches += 1
        elif self.world > 1:
            if self.comm is None and self.peer is not None:
                self.comm = D.native_comm()
            if self.comm is not None:
                _lib.check(self.L.peuqse_comm_allreduce_sum_i64(self.comm, _lib.ptr(self.hist), self.hist.numel(),
                                                                _lib.current_stream()))
            else:
                import torch.distributed as dist
                dist.all_reduce(self.hist, op=dist.ReduceOp.SUM)

    def step(self, log_rel, statistics=True):
        """One whole multistart on the device (no host synchronisation)."""
        self.pass1()
        if statistics:
            self.pass2(log_rel)
            self.pass3(log_rel)

    def graphable(self):
        """Every launch of a step can be captured in a CUDA graph: one rank, or every exchange is a peer-memory kernel."""
        return self.world == 1 or (self.peer is not None and self.hist.numel() * 8 <= self.peer_max)

    def capture(self, fn):
        """fn() captured on a side stream -> torch.cuda.CUDAGraph (None when capture is not possible)."""
        if not (self.use_graph and self.graphable()):
            return None
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    fn()
            cur.wait_stream(side)
            return g
        except Exception as e:                                            # noqa: BLE001
            self._graph_error = repr(e)                                   # kept for diagnosis; the eager path is used instead
            torch.cuda.synchronize()
            return None

    def rest_of_step(self, log_rel):
        """Everything of a step behind the pass-1 kernels (merge / exchange of the best record, pass 2, pass 3 and their
        exchanges) as ONE graph launch when possible: bench.py times  pass1_kernels(); rest_of_step()  per step."""
        key = ("rest", float(log_rel))
        if getattr(self, "_rest_key", None) != key:
            self._rest_key = key
            self.pass1_merge(); self.pass2(log_rel); self.pass3(log_rel)       # eager once (also warms every kernel up)
            self._rest_launches = 0
            l0 = self.launches
            self._rest_graph = self.capture(lambda: (self.pass1_merge(), self.pass2(log_rel), self.pass3(log_rel)))
            self._rest_launches = self.launches - l0 if self._rest_graph is not None else 0
            self.launches = l0
            return
        if self._rest_graph is not None:
            self._rest_graph.replay()
            self.launches += self._rest_launches                          # kernels inside the replayed graph
        else:
            self.pass1_merge(); self.pass2(log_rel); self.pass3(log_rel)

    def points(self, first, count):
        th = torch.empty((count, self.k), dtype=torch.float64, device=self.dev)
        if count:
            _lib.check(self.L.peuqse_sweep_points(C.byref(self.spec), int(first), int(count), _lib.ptr(th),
                                                  _lib.current_stream()))
            self.launches += 1
        return th

    def result_bytes(self):
        return self.res.numel() * 8

    # -- the whole doMultiStart('getLogP') computation, host-facing
    def run(self, coef=2.0, statistics=True):
        k, nbins, lay = self.k, self.nbins, self.lay
        log_rel = float(np.log(10 ** (-1 * coef)))                       # IP.py:1016, 3477
        key = (log_rel, bool(statistics))
        if getattr(self, "_res_host", None) is None:                     # pinned: the packed record comes back by an async copy
            self._res_host = torch.empty(self.res.shape, dtype=torch.float64, pin_memory=True)
        whole = lambda: (self.step(log_rel, statistics), self._res_host.copy_(self.res, non_blocking=True))
        if getattr(self, "_graph_key", None) == key and self._graph is not None:
            self._graph.replay()                                          # the same sweep again: one graph launch
        else:
            whole()
            # The launch sequence is captured when the SAME sweep is asked for a second time (a one-off sweep never pays for
            # a capture; repeated ones replay one graph from their third call on); with several ranks only when every
            # exchange of the step is a peer-memory kernel (their exchange counter lives in device memory, csrc/peer.cu;
            # NCCL calls are not captured).  A failed capture is not retried.
            seen = getattr(self, "_seen_key", None) == key
            self._seen_key = key
            if getattr(self, "_graph_key", None) != key:
                self._graph, self._graph_key = None, None
            if seen and self._graph is None and not getattr(self, "_graph_failed", False):
                g = self.capture(whole)
                if g is not None:
                    self._graph, self._graph_key = g, key
                else:
                    self._graph_failed = True
        torch.cuda.current_stream().synchronize()                        # the single device->host read has landed
        res = self._res_host
        if self.peer is not None and self.L.peuqse_peer_status(self.peer) != 0:
            raise _lib.PeuqseError("multi-GPU exchange timed out: a rank did not reach the exchange within PEUQSE_PEER_TIMEOUT_S "
                                   "seconds (default 60); the merged result is invalid and the peer mailboxes must not be reused")
        R = self.unpack(res.numpy(), coef, statistics)
        return R

    def unpack(self, res, coef, statistics=True):
        k, nbins, lay = self.k, self.nbins, self.lay
        R = SweepResult()
        R.total = self.total
        R.map_logP = float(res[0])
        R.map_index = int(res[1:2].view(np.int64)[0])
        if R.map_index < 0:
            R.map_logP = float("-inf")
            R.map_parameter_set = np.full(k, np.nan)
            R.statistics = False
            return R
        R.map_parameter_set = res[res.shape[0] - k:].copy()
        R.statistics = bool(statistics)
        if not statistics:
            return R
        R.threshold = R.map_logP + np.log(10 ** (-1 * coef))
        acc = res[2:2 + lay["SIZE"]]
        W = acc[lay["W"]]
        c = R.map_parameter_set                                           # moments were taken about the MAP point
        S1 = acc[lay["S1"]:lay["S1"] + k]
        S2 = acc[lay["S2"]:lay["S2"] + k * k].reshape(k, k)
        R.sum_weights = W
        R.n_retained = int(round(acc[lay["NKEPT"]]))
        m1 = S1 / W
        R.mu_AP_parameter_set = c + m1                                    # np.mean of the expanded rows, IP.py:1923
        cov = S2 / W - np.outer(m1, m1)
        R.posterior_covariance = cov
        R.stdap_parameter_set = np.sqrt(np.maximum(np.diag(cov), 0.0))    # np.std (ddof 0), IP.py:1924
        sumsq = np.diag(S2) + 2.0 * c * S1 + c * c * W                    # sum w*theta_d^2
        R.evidence = (acc[lay["WEXP"]] / W) / np.sqrt(np.sum(sumsq))      # IP.py:1931
        R.retained_min = acc[lay["MIN"]:lay["MIN"] + k].copy()
        R.retained_max = acc[lay["MAX"]:lay["MAX"] + k].copy()
        o = 2 + lay["SIZE"]
        R.histogram_counts = res[o:o + k * nbins].view(np.int64).reshape(k, nbins).copy()
        R.histogram_edges = res[o + k * nbins:o + k * nbins + k * (nbins + 1)].reshape(k, nbins + 1).copy()
        return R
