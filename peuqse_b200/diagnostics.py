"""Convergence diagnostics on device-resident chains: the Geweke half of
getConvergenceDiagnostics / calculateAndPlotConvergenceDiagnostics (PEUQSE/InverseProblem.py:2121-2161, 3960-4005, "IP.py")
with geweke_diagnostic (IP.py:4012-4062), for thousands of chains at once (csrc/diag.cu).

The integrated autocorrelation time (IP.py:3916-3948) is zeus.autocorr.AutoCorrTime, a third-party routine that is neither
vendored by the reference nor installed here: autocorr_times() follows its published algorithm (method 'mk': the walkers'
chains laid end to end, FFT autocorrelation function, Sokal's automated window with c = 5) -- PARITY UNPINNED, checked
against the oracle's numpy restatement only (oracle/peuqse_oracle.py:autocorr_time).
The spectral variance inside the Geweke score is statsmodels' yule_walker (order 2, 'adjusted', demeaned), restated from
its published algorithm (oracle/peuqse_oracle.py:yule_walker): parity pinned for everything but that one call.
"""
import numpy as np
import torch

from . import _lib


def geweke_slices(n, initial_window=0.1, comparison_window=0.5, intervals=20):
    """[(start, (a1, b1), (a2, b2))] exactly as geweke_diagnostic forms them (IP.py:4030-4054)."""
    for interval in (initial_window, comparison_window):
        if interval <= 0 or interval >= 1:
            raise ValueError("Invalid intervals for Geweke convergence analysis:", (initial_window, comparison_window), "Must be between 0 and 1")
    if initial_window + comparison_window >= 1:
        raise ValueError("Invalid intervals for Geweke convergence analysis:", (initial_window, comparison_window),
                         "first and last intervals should not overlap.")
    last_index = n - 1
    last_start_index = (1 - comparison_window) * last_index
    out = []
    for start in np.linspace(0, last_start_index, num=intervals, endpoint=True, dtype=int):
        out.append((int(start), (int(start), int(start) + int(initial_window * (last_index - start))),
                    (int(last_index - comparison_window * (last_index - start)), n)))
    return out


_TABLES = {}


def _slice_tables(R, windows, initial_window, comparison_window, intervals):
    """Rows at which prefix sums are needed and, per z-score, the indices of its two slices' six rows (cached per shape)."""
    key = (int(R), tuple(int(w) for w in windows), initial_window, comparison_window, intervals)
    if key in _TABLES:
        return _TABLES[key]
    slices, starts = [], None
    for n in windows:
        sl = geweke_slices(int(n), initial_window, comparison_window, intervals)
        slices.append(sl)
        starts = np.array([s[0] for s in sl])
    need = set()
    for sl in slices:
        for _, s1, s2 in sl:
            for a, b in (s1, s2):
                b = max(b, a)
                need.update(min(max(p, a), b) for p in (a, a + 1, a + 2, b - 2, b - 1, b))
    pos = np.array(sorted(need), dtype=np.int64)
    index = {int(p): i for i, p in enumerate(pos)}
    table = np.empty((len(windows), intervals, 2, 6), dtype=np.int32)
    for w, sl in enumerate(slices):
        for i, (_, s1, s2) in enumerate(sl):
            for h, (a, b) in enumerate((s1, s2)):
                b = max(b, a)
                table[w, i, h] = [index[min(max(p, a), b)] for p in (a, a + 1, a + 2, b - 2, b - 1, b)]
    _TABLES[key] = (pos, table, starts)
    return _TABLES[key]


def geweke_z_scores(chains, windows=None, initial_window=0.1, comparison_window=0.5, intervals=20):
    """chains[R, C, k] (CUDA float64 tensor or array) -> (z[k, C, W, intervals] CUDA tensor, windows, start indices of the
    last window).  windows default to the reference's 20 growing windows np.linspace(0, R, 21).astype(int)[1:]."""
    L = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = chains if torch.is_tensor(chains) else torch.as_tensor(np.ascontiguousarray(np.asarray(chains, dtype=float)))
    x = x.to(dev, dtype=torch.float64).contiguous()
    R, C, k = x.shape
    if windows is None:
        windows = np.linspace(0, R, 21).astype(int)[1:]                              # IP.py:3963
    windows = np.asarray(windows, dtype=np.int64)
    pos, table, starts = _slice_tables(R, windows, initial_window, comparison_window, intervals)
    nz = len(windows) * intervals
    d_pos = torch.as_tensor(pos).to(dev)
    d_tab = torch.as_tensor(table.reshape(-1)).to(dev)
    z = torch.empty((nz, C * k), dtype=torch.float64, device=dev)
    work = torch.empty(int(L.peuqse_geweke_work_bytes(R, C * k, len(pos))), dtype=torch.uint8, device=dev)
    _lib.check(L.peuqse_geweke(_lib.ptr(x), R, C * k, pos.ctypes.data_as(_lib.C.c_void_p), _lib.ptr(d_pos), len(pos), _lib.ptr(d_tab), nz,
                               _lib.ptr(z), _lib.ptr(work), _lib.current_stream()))
    return z.reshape(len(windows), intervals, C, k).permute(3, 2, 0, 1), windows, starts


def autocorr_times(chains, windows=None, c=5.0):
    """chains[R, C, k] (CUDA float64 tensor or array) -> (taus[W, k] CUDA tensor, windows): zeus.autocorr.AutoCorrTime of
    chains[:n] for every growing window n (default: the reference's np.exp(np.linspace(0, log R, 21)).astype(int)[1:],
    IP.py:3918).  Per window: csrc/diag.cu k_act_concat writes, for all parameters in one read, the demeaned walker-major
    series zero-padded to twice the next power of two; the transform pair is cuFFT (torch.fft -- a plain library FFT, like
    numpy's in the routine this replaces); the normalisation, running sum and automated window stay on the device (no host
    synchronisation per window).  Device memory: k series of 2 * next_pow2(R * C) doubles for the largest window plus three
    transform buffers of that length (C3 shape: ~80 GB beside the 17.7 GB of samples)."""
    L = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    x = chains if torch.is_tensor(chains) else torch.as_tensor(np.ascontiguousarray(np.asarray(chains, dtype=float)))
    x = x.to(dev, dtype=torch.float64).contiguous()
    R, C, k = x.shape
    if windows is None:
        with np.errstate(all='ignore'):
            windows = np.exp(np.linspace(0, np.log(R), 21)).astype(int)[1:]          # IP.py:3918
    windows = np.asarray(windows, dtype=np.int64)
    taus = torch.empty((len(windows), k), dtype=torch.float64, device=dev)
    launches = 0
    for i, n in enumerate(int(w) for w in windows):
        if n < 1 or n > R:
            raise ValueError("window sizes must lie in [1, number of samples]")
        length = n * C
        n2 = 2
        while n2 < 2 * length:
            n2 <<= 1                                                                 # 2 * next power of two >= length
        # a true division by a device tensor (torch turns division by a host scalar into a multiplication by its reciprocal):
        # the mean of a constant column is exact, its series is 0 and its tau 0/0 = NaN as in numpy
        mean = (x[:n].sum(dim=(0, 1)) / torch.full((k,), float(length), dtype=torch.float64, device=dev)).contiguous()
        series = torch.empty((k, n2), dtype=torch.float64, device=dev)
        _lib.check(L.peuqse_act_concat(_lib.ptr(x), n, C, k, _lib.ptr(mean), _lib.ptr(series), n2, _lib.current_stream()))
        launches += 1
        lag = torch.arange(n, device=dev, dtype=torch.float64)
        for d in range(k):
            f = torch.fft.rfft(series[d])
            power = f.real * f.real + f.imag * f.imag
            del f
            acf = torch.fft.irfft(power, n=n2)[:n]
            del power
            acf = acf / acf[0]
            t = 2.0 * torch.cumsum(acf, dim=0) - 1.0
            m = lag < c * t
            first_false = torch.where(m, torch.full_like(lag, float(n)), lag).min()  # np.argmin(m): the first False ...
            idx = torch.where(first_false < n, first_false,                          # ... 0 if all True, n-1 if none is
                              torch.where(m.any(), torch.zeros_like(first_false), torch.full_like(first_false, float(n - 1))))
            taus[i, d] = t[idx.long()]
        del series
    return taus, windows, launches


def convergence_diagnostics(pe, discrete_chains_post_burn_in_samples=None):
    """getConvergenceDiagnostics (IP.py:2121-2161): fills PE_object.convergence; returns the reference's output tuple."""
    x = discrete_chains_post_burn_in_samples
    if x is None or (not torch.is_tensor(x) and len(x) == 0):
        out = getattr(pe, '_mh_out', None)
        if out is not None and out.get('samples') is not None and hasattr(pe, '_mh_burn_rows'):
            x = out['samples'][pe._mh_burn_rows:]                                  # still on the device
        else:
            x = pe.discrete_chains_post_burn_in_samples
    if not torch.is_tensor(x) and type(x).__module__ != np.__name__:
        print('The input array needs to a numpy array. If a list was inputted, wrap the variable in np.array(var)'); raise SystemExit
    if len(x.shape) != 3:
        print('The input array needs to take the shape (numSamples, numChains, numParameters). The current inputs does not have the appropriate dimensions.')
        raise SystemExit
    z, windows, starts = geweke_z_scores(x)
    pe.gpu_launches += 3
    a = z.abs()
    comb = a.mean(dim=0).mean(dim=0)                                                # over parameters, then chains (IP.py:3993-3996)
    final = comb[:, -1].cpu().numpy()                                               # IP.py:3998
    nan_safe = torch.where(torch.isnan(comb), torch.zeros_like(comb), comb)
    outlier = ((nan_safe > 1).sum(dim=1).double() / comb.shape[1]).cpu().numpy()    # IP.py:4000
    per_param = a.mean(dim=1)                                                       # [k, W, I]: the per-parameter plots' data
    taus, act_windows, n_act = autocorr_times(x)                                    # IP.py:3916-3948
    pe.gpu_launches += n_act
    taus = taus.cpu().numpy()
    names = list(getattr(pe.UserInput, 'parameterNamesList', pe.UserInput.model['parameterNamesAndMathTypeExpressionsDict'])) if hasattr(pe, 'UserInput') else []
    if len(names) != taus.shape[1]:
        names = ['parameter_%d' % d for d in range(taus.shape[1])]
    act_each = {name: taus[:, d] for d, name in enumerate(names)}                   # IP.py:3938-3942
    pe.convergence = {'AutoCorrTime': {'window_indices': act_windows, 'final_parameter_values': taus[-1, :],
                                       'parameter_act_for_each_window': act_each,
                                       'combined_parameter_act_for_each_window': np.prod(taus, axis=1)},
                      'Geweke': {'window_indices': windows, 'final_combined_parameter_values': final,
                                 'final_combined_parameter_percent_outlier': outlier,
                                 'interval_start_indices': starts, 'z_scores_each_parameter': per_param.cpu().numpy()}}
    pe.geweke_z_scores = z
    return (act_windows, taus[-1, :], act_each, windows, final, outlier)
