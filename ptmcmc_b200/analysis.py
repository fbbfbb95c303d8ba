"""Host-side sample analysis used by the statistical parity tests and by bench.py's ESS/s line (not on the hot path).

* `knn_kl(p, q, k)`: k-nearest-neighbour estimator of KL(P||Q) from samples (Perez-Cruz 2008), the estimator behind the
  reference's `test_proposal::KL_divergence` (test_proposal.hh:350-423) and `python/ptmcmc_analysis.py:186-216`, with k > 1
  for lower variance.
* `integrated_act(x)`: integrated autocorrelation time with Sokal's self-consistent window; ESS = N / tau.  The reference's
  own estimator (chain.cc:126-643, windowed autocovariance with log-spaced lags) is a different finite-sample recipe for the
  same quantity; parity compares reference and engine chains with the SAME estimator, which is what "ESS per sample within
  10 %" needs.
"""
import numpy as np
from scipy.spatial import cKDTree


def knn_kl(p, q, k=4):
    """KL(P||Q) in nats from samples p[n,d] ~ P and q[m,d] ~ Q"""
    p = np.asarray(p, dtype=float); q = np.asarray(q, dtype=float)
    n, d = p.shape; m = q.shape[0]
    rho = cKDTree(p).query(p, k=k + 1)[0][:, k]   # k-th neighbour within p, excluding the point itself
    nu = cKDTree(q).query(p, k=k)[0]
    nu = nu[:, k - 1] if k > 1 else nu
    ok = (rho > 0) & (nu > 0)
    return d * np.mean(np.log(nu[ok] / rho[ok])) + np.log(m / (n - 1.0))


def autocorr_fft(x):
    """normalised autocorrelation function of the 1-D series x"""
    x = np.asarray(x, dtype=float)
    n = len(x)
    f = np.fft.rfft(x - x.mean(), n=2 * n)
    acf = np.fft.irfft(f * np.conjugate(f))[:n].real
    return acf / acf[0] if acf[0] > 0 else np.ones(n)


def integrated_act(x, c=5.0):
    """integrated autocorrelation time tau = 1 + 2 sum_{t<=M} rho_t, with the smallest window M >= c * tau(M).
    x: [n] or [n_chains, n] (the autocorrelation functions are averaged over chains before windowing)"""
    x = np.atleast_2d(np.asarray(x, dtype=float))
    rho = np.mean([autocorr_fft(row) for row in x], axis=0)
    taus = 2.0 * np.cumsum(rho) - 1.0
    window = np.arange(len(taus)) >= c * taus
    m = int(np.argmax(window)) if window.any() else len(taus) - 1
    return float(max(taus[m], 1.0))


def ess_per_sample(chains):
    """chains: [n_chains, n_steps, d] cold-chain samples; returns (min over parameters of 1/tau, taus)"""
    chains = np.asarray(chains, dtype=float)
    taus = np.array([integrated_act(chains[:, :, j]) for j in range(chains.shape[2])])
    return float(1.0 / taus.max()), taus
