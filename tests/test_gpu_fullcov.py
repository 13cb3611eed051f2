"""Dense-covariance likelihood (csrc/fullcov.cu) at the size config C4 is benchmarked on: two responses of 4096 observations
with full 4096 x 4096 covariances (the reference's route to a full covariance: >= 2 response dimensions, IP.py:2938-2960,
SURVEY Q2).  Every kernel shape -- GEMV-shaped (1, 2, 4, 8 row fragments), split-K tiles, fused tiles -- against scipy's
multivariate_normal.logpdf arithmetic (x = observed, mean = simulated) at 1e-12 relative, and against each other bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K, N = 12, 4096


@pytest.fixture(scope="module")
def c4():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import peuqse_b200 as PEUQSE
    from oracle import peuqse_oracle as orc
    rng = np.random.default_rng(1)
    x = np.linspace(0, 1, N)
    Phi = orc.cos_basis(2 * N, K)
    Sig = 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * 0.02 ** 2)) + 0.05 ** 2 * np.eye(N)
    truth = rng.standard_normal(K)
    y = Phi @ truth + 0.1 * np.random.default_rng(2).standard_normal(2 * N)

    def make(whitening):
        UI = PEUQSE.UserInput.fresh()
        UI.responses['responses_observed'] = [list(y[:N]), list(y[N:])]
        UI.responses['responses_observed_uncertainties'] = [Sig, Sig]
        UI.responses['responses_observed_max_covmat_size'] = N
        UI.responses['responses_observed_covmat_whitening'] = whitening
        UI.model['InputParameterPriorValues'] = [0.0] * K
        UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * K
        UI.model['simulateByInputParametersOnlyFunction'] = PEUQSE.LinearModel(Phi)
        return PEUQSE.parameter_estimation(UI)
    wh = orc.psd_whitening(Sig)                                    # scipy's _PSD, once
    th = truth + 0.05 * np.random.default_rng(3).standard_normal((1024, K))
    sim = th @ Phi.T
    c = wh.rank * orc.LOG2PI + wh.log_pdet
    ll = np.zeros(len(th))
    for r in range(2):                                             # scipy _multivariate.py:590-591 per response, summed in order
        t = (y[r * N:(r + 1) * N] - sim[:, r * N:(r + 1) * N]) @ wh.U
        ll = ll + (-0.5 * (c + np.sum(t * t, axis=1)))
    return dict(torch=torch, orc=orc, make=make, th=th, ll=ll, Sig=Sig, y=y, Phi=Phi)


def _ll(PE, th):
    return PE.getLogP_batch(th, return_parts=True)[2].cpu().numpy()


@pytest.mark.parametrize("whitening", ["eigh", "cholesky"])
def test_c4_size_every_kernel_shape_vs_scipy_arithmetic(c4, whitening):
    PE = c4['make'](whitening)
    th, want = c4['th'], c4['ll']
    big = _ll(PE, th)                                              # 1024 rows: fused tiles
    np.testing.assert_allclose(big, want, rtol=1e-12)
    # B = 2, 9, 20, 40 (GEMV shape with 1, 2, 4, 8 row fragments), 130 and 200 (split-K tiles): the rows' values must not depend
    # on the batch they are evaluated in -- not in the last bit
    for lo, B in ((0, 2), (5, 9), (100, 20), (300, 40), (17, 64), (500, 130), (700, 200), (0, 65)):
        got = _ll(PE, th[lo:lo + B])
        assert np.array_equal(got, big[lo:lo + B]), (whitening, B, np.max(np.abs(got - big[lo:lo + B])))


def test_c4_size_against_scipy_logpdf_itself(c4):
    """Two points through scipy.stats.multivariate_normal.logpdf (an eigendecomposition of the 4096 x 4096 matrix per call:
    what the reference executes at IP.py:2960), both whitenings."""
    from scipy.stats import multivariate_normal as mvn
    y, Phi, Sig, th = c4['y'], c4['Phi'], c4['Sig'], c4['th'][:2]
    want = np.array([mvn.logpdf(x=y[:N], mean=(Phi @ t)[:N], cov=Sig) + mvn.logpdf(x=y[N:], mean=(Phi @ t)[N:], cov=Sig) for t in th])
    for whitening in ("eigh", "cholesky"):
        np.testing.assert_allclose(_ll(c4['make'](whitening), th), want, rtol=1e-12)


def test_ragged_sizes_and_mixed_responses(c4):
    """Sizes that are not multiples of any tile dimension (n = 520 and 77, B = 1, 63, 131, 700), a diagonal response between the
    dense ones, a torch-callable simulator whose values arrive in natural order, NaN rows -> -inf (IP.py:2880-2882)."""
    torch, orc = c4['torch'], c4['orc']
    import peuqse_b200 as PEUQSE
    from scipy.stats import multivariate_normal as mvn
    rng = np.random.default_rng(11)
    n1, n2, n3, k = 520, 33, 77, 3
    def cov(n, ell):
        x = np.linspace(0, 1, n)
        return 0.1 ** 2 * np.exp(-(x[:, None] - x[None, :]) ** 2 / (2 * ell ** 2)) + 0.05 ** 2 * np.eye(n)
    S1, S3 = cov(n1, 0.05), cov(n3, 0.1)
    s2 = np.full(n2, 0.2)
    Phi = orc.cos_basis(n1 + n2 + n3, k)
    truth = rng.standard_normal(k)
    y = Phi @ truth + 0.1 * rng.standard_normal(n1 + n2 + n3)
    obs = [list(y[:n1]), list(y[n1:n1 + n2]), list(y[n1 + n2:])]
    for whitening in ("eigh", "cholesky"):
        def make(model):
            UI = PEUQSE.UserInput.fresh()
            UI.responses['responses_observed'] = obs
            UI.responses['responses_observed_uncertainties'] = [S1, s2, S3]
            UI.responses['responses_observed_max_covmat_size'] = 600
            UI.responses['responses_observed_covmat_whitening'] = whitening
            UI.model['InputParameterPriorValues'] = [0.0] * k
            UI.model['InputParametersPriorValuesUncertainties'] = [2.0] * k
            UI.model['simulateByInputParametersOnlyFunction'] = model
            return PEUQSE.parameter_estimation(UI)
        PhiT = torch.from_numpy(Phi.T.copy()).cuda()
        PE, PEx = make(PEUQSE.LinearModel(Phi)), make(PEUQSE.TorchSimulator(lambda t: t @ PhiT))
        th = truth + 0.1 * rng.standard_normal((700, k))
        sim = th @ Phi.T
        # with >= 2 responses the reference uses a vector of n > 1 sigmas AS VARIANCES (IP.py:3599-3605)
        want = np.array([mvn.logpdf(x=y[:n1], mean=s[:n1], cov=S1) + mvn.logpdf(x=y[n1:n1 + n2], mean=s[n1:n1 + n2], cov=np.diag(s2))
                         + mvn.logpdf(x=y[n1 + n2:], mean=s[n1 + n2:], cov=S3) for s in sim[:40]])
        big = _ll(PE, th)
        np.testing.assert_allclose(big[:40], want, rtol=1e-12)
        np.testing.assert_allclose(_ll(PEx, th[:131]), big[:131], rtol=1e-12)
        for lo, B in ((0, 1), (10, 63), (200, 131), (3, 64), (0, 65)):
            assert np.array_equal(_ll(PE, th[lo:lo + B]), big[lo:lo + B]), (whitening, B)
        bad = th[:70].copy()
        bad[3, 0] = np.nan
        out = _ll(PE, bad)
        assert out[3] == -np.inf and np.array_equal(np.delete(out, 3), np.delete(big[:70], 3))
