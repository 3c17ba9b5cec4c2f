"""Figure helpers — host-side mirror of reference ``plotlib.py`` (six matplotlib / corner
methods, pure consumers of ``get_chain`` / ``get_model_percentile`` / ``data``).

matplotlib and corner are imported lazily (neither is a dependency of the hot path, and
neither is installed in the build image); a missing package raises ``ImportError`` at call
time, which is what reference ``tests/test_module.py:92-98`` expects for ``plot_corner``.
"""
import numpy as np


def _plt():
    import matplotlib.pyplot as plt
    return plt


class plotlib(object):

    def plot_traces(self, chain=None, **kwargs):
        """Walker traces per parameter (reference ``plotlib.py:17-54``)."""
        self._check_if_fitted()
        plt = _plt()
        if chain is None:
            chain = self.get_chain(**kwargs)
        names = self.param_names
        bounds = self.param_bounds
        fig, axes = plt.subplots(self.ndim, figsize=(8, 6), sharex=True)
        for i, ax in enumerate(np.atleast_1d(axes)):
            ax.plot(chain[:, :, i], 'k', alpha=0.3)
            ax.set_xlim(0, len(chain))
            ax.set_ylim(bounds[:, i])
            ax.set_ylabel(names[i])
            ax.yaxis.set_label_coords(-0.1, 0.5)
        np.atleast_1d(axes)[-1].set_xlabel('Steps')
        fig.tight_layout()
        return fig

    def plot_histograms(self, chain=None, bins=25, **kwargs):
        """Marginal histograms per parameter (reference ``plotlib.py:56-90``)."""
        self._check_if_fitted()
        plt = _plt()
        chain = self.parse_chain(chain, **kwargs)
        names = self.param_names
        fig, axes = plt.subplots(self.ndim, figsize=(5, 1.5 * chain.shape[1]))
        for i, ax in enumerate(np.atleast_1d(axes)):
            ax.hist(chain[:, i], bins=bins, fc='w', ec='k')
            ax.set_xlabel(names[i])
            ax.ticklabel_format(axis='x', scilimits=[-2, 2])
        fig.tight_layout()
        return fig

    def plot_fit(self, chain=None, p=[2.5, 50, 97.5], **kwargs):
        """Data with best fit and credible band, real and imaginary parts
        (reference ``plotlib.py:92-135``)."""
        self._check_if_fitted()
        plt = _plt()
        data = self.data
        lo, mid, hi = self.get_model_percentile(p, chain, **kwargs)
        fig, ax = plt.subplots(1, 2, figsize=(8, 3))
        for i in range(2):
            ax[i].errorbar(data['freq'], data['zn'][i], yerr=data['zn_err'][i], markersize=3, fmt=".k", capsize=0)
            ax[i].plot(data['freq'], lo[i], ls=':', c='0.5')
            ax[i].plot(data['freq'], mid[i], c='C3')
            ax[i].plot(data['freq'], hi[i], ls=':', c='0.5')
            ax[i].set_ylabel(r'$\rho${} (normalized)'.format((i + 1) * "'"))
            ax[i].set_xscale('log')
            ax[i].set_xlabel('$f$ (Hz)')
        fig.tight_layout()
        return fig

    def plot_data(self, feature='phase', **kwargs):
        """Raw data: 'phase', 'amplitude', 'real' or 'imaginary' (reference ``plotlib.py:137-175``).
        'real'/'imaginary' use the actual real/imaginary parts (the reference indexes the
        complex vector by position there)."""
        plt = _plt()
        kwargs.setdefault('fmt', '.k')
        kwargs.setdefault('capsize', 0)
        kwargs.setdefault('markersize', 3)
        d = self.data
        nf = d['norm_factor']
        series = {
            'phase': (-d['pha'], d['pha_err'], '-Phase (rad)'),
            'amplitude': (d['amp'] / nf, d['amp_err'] / nf, 'Amplitude (normalized)'),
            'real': (d['Z'].real / nf, d['Z_err'].real / nf, 'Real part (normalized)'),
            'imaginary': (-d['Z'].imag / nf, d['Z_err'].imag / nf, '-Imaginary part (normalized)'),
        }
        yv, ye, label = series[feature]
        fig, ax = plt.subplots()
        ax.errorbar(d['freq'], yv, yerr=ye, **kwargs)
        ax.set_xlabel('Frequency (Hz)')
        ax.set_ylabel(label)
        ax.set_xscale('log')
        fig.tight_layout()
        return fig

    def plot_fit_pa(self, chain=None, p=[2.5, 50, 97.5], **kwargs):
        """Data with best fit and credible band as amplitude and phase
        (reference ``plotlib.py:177-232``)."""
        self._check_if_fitted()
        plt = _plt()
        d = self.data
        nf = d['norm_factor']
        lines = self.get_model_percentile(p, chain, **kwargs)
        styles = [dict(ls=':', c='0.5'), dict(c='C3'), dict(ls=':', c='0.5')]
        fig, ax = plt.subplots(1, 2, figsize=(8, 3))
        ax[0].errorbar(d['freq'], d['amp'] / nf, yerr=d['amp_err'] / nf, markersize=3, fmt=".k", capsize=0)
        ax[1].errorbar(d['freq'], -d['pha'], yerr=d['pha_err'], markersize=3, fmt=".k", capsize=0)
        for line, st in zip(lines, styles):
            ax[0].plot(d['freq'], np.linalg.norm(line, axis=0), **st)
            ax[1].plot(d['freq'], -np.arctan2(line[1], line[0]), **st)
        ax[0].set_ylabel('Amplitude (normalized)')
        ax[1].set_ylabel('-Phase (rad)')
        ax[1].set_yscale('log')
        for a in ax:
            a.set_xscale('log')
            a.set_xlabel('$f$ (Hz)')
        fig.tight_layout()
        return fig

    def plot_corner(self, chain=None, **kwargs):
        """Corner plot of the posterior (reference ``plotlib.py:234-259``)."""
        self._check_if_fitted()
        from corner import corner
        chain = self.parse_chain(chain, **kwargs)
        return corner(chain, labels=self.param_names)
