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


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own effective-sample-size recipe (chain.cc:126-545), restated over a cold-chain history array so that
# ESS numbers quoted by this engine and by the reference's run loop (ptmcmc.cc:645) are the same estimator.
# `hist` is the raw history of one chain, [Nsize, n_features]: n_init start-up entries followed by one entry per
# `add_every` steps; `nstep` is the chain's step count (MH_chain::getStep() = Nhist).  Step i lives at history row
# n_init + i // add_every (MH_chain::get_state_idx, chain.cc:1041-1050).

def _lag_grid(swidth, nevery, max_lag, dlag):
    """Approximately logarithmic lag grid in steps (chain.cc:207-222)."""
    lags = [0]
    fac, idx = 1.0, 1
    while idx < max_lag * swidth:
        lags.append(nevery * idx)
        last = idx
        while last == idx:
            fac *= dlag
            idx = int(fac)
    return lags


def autocovar_windows(hist, nstep, n_init, add_every, width, nevery, burn_windows=1, max_lag=0, dlag=np.sqrt(2.0)):
    """chain::compute_autocovar_windows with loglag=true (chain.cc:126-289), all features at once.

    Returns covar[f][win][lag], means[f][win][lag], counts[win][lag], windows[Nwin+1], lags[Nlag], width (rounded)."""
    width = max(int(width), 2)
    nevery = max(int(nevery), 1)
    swidth = width // nevery
    width = swidth * nevery
    burn_windows = max(int(burn_windows), 1)
    if max_lag == 0 or max_lag > burn_windows:
        max_lag = burn_windows
    dlag = max(dlag, 1.01)
    nwin = max(nstep // width - burn_windows, 0)
    istart = nstep - nwin * width
    windows = [istart + i * width for i in range(nwin + 1)]
    lags = _lag_grid(swidth, nevery, max_lag, dlag)
    nlag, nf = len(lags), hist.shape[1]
    covar = np.zeros((nf, nwin, nlag))
    means = np.zeros((nf, nwin, nlag))
    counts = np.full((nwin, nlag), swidth, dtype=np.int64)
    off = np.arange(swidth, dtype=np.int64) * nevery
    for k in range(nwin):
        idx = windows[k] + off
        f = hist[n_init + idx // add_every]                     # [swidth, nf]
        fsum = f.sum(axis=0)
        means[:, k, 0] = fsum / swidth
        covar[:, k, 0] = (f * f).sum(axis=0) / swidth - means[:, k, 0] ** 2
        for j in range(1, nlag):
            fl = hist[n_init + (idx - lags[j]) // add_every]
            means[:, k, j] = (fl + f).sum(axis=0) / swidth / 2       # mean of window and lagged window
            covar[:, k, j] = (fl * f).sum(axis=0) / swidth - means[:, k, j] ** 2
    return covar, means, counts, windows, lags, width


def compute_effective_samples(hist, nstep, n_init, add_every, width, nevery, burn_windows=1, max_lag=0, dlag=np.sqrt(2.0)):
    """chain::compute_effective_samples (chain.cc:292-449): (ess, best_nwin), min over features, max over tail lengths."""
    oversmall_aclen_fac = 3.0
    covar, means, counts, windows, lags, _ = autocovar_windows(hist, nstep, n_init, add_every, width, nevery, burn_windows, max_lag, dlag)
    nevery = max(int(nevery), 1)
    nf, nwin_tot, nlag = covar.shape
    ess_max, nwin_max = 0.0, 0
    for nwin in range(1, nwin_tot + 1):
        ess = 1e100
        sl = slice(nwin_tot - nwin, nwin_tot)
        for ifeat in range(nf):
            mean = means[ifeat, sl, 0].sum() / nwin
            c = counts[sl]
            dm = mean - means[ifeat, sl, :]
            dm0 = (mean - means[ifeat, sl, 0])[:, None]
            num = ((covar[ifeat, sl, :] + dm * dm) * c).sum(axis=0)
            den = ((covar[ifeat, sl, 0][:, None] + dm0 * dm0) * c).sum(axis=0)
            last_lag, ac_len, lastcorr, dacl = 0, 1.0, 1.0, 0.0
            for ilag in range(1, nlag):
                corr = num[ilag] / den[ilag]
                if lastcorr < 0 and corr < 0:                    # initially-positive-sequence cut
                    ac_len -= dacl
                    break
                lastcorr = corr
                dacl = 2.0 * (lags[ilag] - last_lag) * corr
                ac_len += dacl
                last_lag = lags[ilag]
            essi = nwin * width / ac_len
            if ac_len < nevery:
                essi = nwin * width / oversmall_aclen_fac / nevery
            ess = min(ess, essi)
        if ess > ess_max:
            ess_max, nwin_max = ess, nwin
    return ess_max, nwin_max


def report_effective_samples(hist, nstep, n_init=0, add_every=1, width=40000, every=100, esslimit=-1, imax=-1):
    """chain::report_effective_samples(imax, width, every, esslimit) (chain.cc:457-643): returns (ess, useful length).

    The run loop calls it as report_effective_samples(-1, save_every*1000, save_every, esslimit) (ptmcmc.cc:645)."""
    hist = np.asarray(hist, dtype=np.float64)
    if hist.ndim == 1:
        hist = hist[:, None]
    dim = hist.shape[1]
    if imax < 0 or imax > dim:
        imax = dim
    hist = hist[:, :min(imax, 20)]                                # the reference supports the first 20 parameters
    width = int(width)
    while width < nstep * 0.05:
        width *= 2
    minburn, minbin, maxbins, scalestep, oversmall = 2, 1000, 20, 2, 3.0
    nsize = hist.shape[0]
    if every < 0:
        every = int(0.5 + (float(nstep) - n_init) / (nsize - n_init))
    every = max(int(every), 1)
    ess, nwin, bestwid = 0.0, 0, 0
    if esslimit < 0:
        if width < 0:
            width = every * minbin
        while width * (maxbins + minburn) < nstep:
            width *= 2
        ess, nwin = compute_effective_samples(hist, nstep, n_init, add_every, width, every, minburn, 0, 1.1)
        bestwid = width
    else:
        full = float(nstep)
        slimit = esslimit * oversmall
        done = False
        while True:
            burnwidth = full / (maxbins + minburn)
            bins = min(int(full / (minbin * every)), maxbins)
            if bins < 1:
                break
            width = int(full / bins)
            if width * (bins - 1) > slimit * every:
                bins = min(int(slimit / minbin + 1), maxbins)
                if bins > 1:
                    width = int((slimit * every) / (bins - 1))
                else:
                    bins, width = 1, minbin * every
            else:
                done = True
            if (full - burnwidth) * 0.5 < bins * width:
                burn = int(full / width - bins)
                essc, nwinc = compute_effective_samples(hist, nstep, n_init, add_every, width, every, burn, 0, 1.1)
                if essc > ess:
                    ess, nwin, bestwid = essc, nwinc, width
            if done:
                break
            every *= scalestep
    return ess, int(bestwid * nwin)


def effective_samples_from_windows(covar, means, lags, width, nevery, swidth):
    """chain::compute_effective_samples (chain.cc:340-416) for a BATCH of chains at once, from window statistics
    covar / means [n_chains][n_feat][n_win][n_lag] (e.g. ptg_get_autocovar_windows).  -> ess[n_chains], best_nwin[n_chains]"""
    oversmall_aclen_fac = 3.0
    covar, means = np.asarray(covar), np.asarray(means)
    n, nf, nwin_tot, nlag = covar.shape
    ess_max, nwin_max = np.zeros(n), np.zeros(n, dtype=np.int64)
    for nwin in range(1, nwin_tot + 1):
        sl = slice(nwin_tot - nwin, nwin_tot)
        ess = np.full(n, 1e100)
        for f in range(nf):
            mean = np.zeros(n)
            for i in range(nwin_tot - nwin, nwin_tot):
                mean = mean + means[:, f, i, 0]
            mean = mean / nwin
            dm = mean[:, None, None] - means[:, f, sl, :]
            dm0 = (mean[:, None] - means[:, f, sl, 0])[:, :, None]
            num = np.zeros((n, nlag)); den = np.zeros((n, nlag))
            for i in range(nwin):                                   # window by window, the reference's summation order
                num = num + (covar[:, f, sl, :][:, i, :] + dm[:, i, :] * dm[:, i, :]) * swidth
                den = den + (covar[:, f, sl, 0][:, i, None] + dm0[:, i, :] * dm0[:, i, :]) * swidth
            ac_len, lastcorr, dacl = np.ones(n), np.ones(n), np.zeros(n)
            done = np.zeros(n, dtype=bool)
            last_lag = 0
            with np.errstate(divide="ignore", invalid="ignore"):
                for il in range(1, nlag):
                    corr = num[:, il] / den[:, il]
                    brk = ~done & (lastcorr < 0) & (corr < 0)          # initially-positive-sequence cut
                    ac_len[brk] -= dacl[brk]
                    done |= brk
                    upd = ~done
                    lastcorr[upd] = corr[upd]
                    dacl[upd] = 2.0 * (lags[il] - last_lag) * corr[upd]
                    ac_len[upd] += dacl[upd]
                    last_lag = lags[il]
                essi = nwin * width / ac_len
            small = ac_len < nevery
            essi[small] = nwin * width / oversmall_aclen_fac / nevery
            ess = np.where(essi < ess, essi, ess)
        better = ess > ess_max
        ess_max[better] = ess[better]
        nwin_max[better] = nwin
    return ess_max, nwin_max


def recipe_geometry(nstep, width, every):
    """the window geometry report_effective_samples settles on for a chain of nstep steps with esslimit < 0 (chain.cc:475-481,
    549-552 and compute_autocovar_windows :179-222): -> (width, swidth, n_win, lags in steps)"""
    minburn, maxbins = 2, 20
    width, every = int(width), max(int(every), 1)
    while width < nstep * 0.05:
        width *= 2
    while width * (maxbins + minburn) < nstep:
        width *= 2
    swidth = width // every
    n_win = max(nstep // (swidth * every) - minburn, 0)
    return width, swidth, n_win, _lag_grid(swidth, every, minburn, 1.1)
