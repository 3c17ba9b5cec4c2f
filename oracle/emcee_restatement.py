"""TEST INFRASTRUCTURE — restatement of the slice of ``emcee`` (3.x) that BISIP calls.

emcee is a third-party, UNPINNED dependency of the reference (``requirements.txt:2``); its
source is neither under /root/reference nor installed in this image (no network), so this
module restates its published algorithm (Goodman & Weare 2010 stretch move; Foreman-Mackey
et al. 2013) following SURVEY.md App. B.  **Parity unpinned**: no emcee golden vectors exist
in the reference; this restatement is anchored on the reference's call sites
(``models.py:111-118`` construct + run, ``models.py:137`` get_chain) and on the stochastic
notebook outputs to Monte-Carlo error.

``oracle/refload.py`` registers this module as ``sys.modules['emcee']`` so that the
UNMODIFIED reference ``Inversion.fit`` runs end to end on the CPU — that is the
``--impl reference`` arm of bench.py and the source of the MC-error golden summaries.

Draw order per step (NumPy legacy MT19937 ``RandomState`` cloned from the global state at
construction): move choice; ``shuffle(inds)``; then per split ``rand(Ns)``,
``randint(Nc, size=Ns)``, and one ``rand()`` per walker in index order.
"""
import numpy as np

__version__ = "3-restatement"


def walkers_independent(coords):
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C /= C_colmax
    C_colsum = np.sqrt(np.sum(C ** 2, axis=0))
    C /= C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


class StretchMove:
    def __init__(self, a=2.0, nsplits=2, randomize_split=True, live_dangerously=False):
        self.a = a
        self.nsplits = nsplits
        self.randomize_split = randomize_split
        self.live_dangerously = live_dangerously

    def get_proposal(self, s, c, random):
        Ns, Nc = len(s), len(c)
        ndim = s.shape[1]
        zz = ((self.a - 1.0) * random.rand(Ns) + 1) ** 2.0 / self.a
        factors = (ndim - 1.0) * np.log(zz)
        rint = random.randint(Nc, size=(Ns,))
        return c[rint] - (c[rint] - s) * zz[:, None], factors

    def propose(self, sampler, coords, log_prob):
        nwalkers, ndim = coords.shape
        if nwalkers < 2 * ndim and not self.live_dangerously:
            raise RuntimeError("It is unadvisable to use a red-blue move with fewer walkers "
                               "than twice the number of dimensions.")
        random = sampler._random
        accepted = np.zeros(nwalkers, dtype=bool)
        all_inds = np.arange(nwalkers)
        inds = all_inds % self.nsplits
        if self.randomize_split:
            random.shuffle(inds)
        for split in range(self.nsplits):
            S1 = inds == split
            sets = [coords[inds == j] for j in range(self.nsplits)]
            s = sets[split]
            c = np.concatenate(sets[:split] + sets[split + 1:], axis=0)
            q, factors = self.get_proposal(s, c, random)
            new_log_probs = sampler.compute_log_prob(q)
            for i, (j, f, nlp) in enumerate(zip(all_inds[S1], factors, new_log_probs)):
                lnpdiff = f + nlp - log_prob[j]
                if lnpdiff > np.log(random.rand()):
                    accepted[j] = True
            m1 = accepted & S1
            m2 = accepted[S1]
            coords[m1] = q[m2]
            log_prob[m1] = new_log_probs[m2]
        return coords, log_prob, accepted


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, moves=None, args=None,
                 kwargs=None, **_ignored):
        self.nwalkers = nwalkers
        self.ndim = ndim
        self.log_prob_fn = log_prob_fn
        self.args = args or ()
        self.kwargs = kwargs or {}
        self.pool = pool
        self._move = moves if moves is not None else StretchMove()
        self._random = np.random.mtrand.RandomState()
        self._random.set_state(np.random.get_state())
        self.reset()

    def reset(self):
        self.iteration = 0
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._log_prob = np.empty((0, self.nwalkers))
        self.accepted = np.zeros(self.nwalkers)

    def compute_log_prob(self, coords):
        p = coords
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN")
        mapf = self.pool.map if self.pool is not None else map
        lp = np.array([float(v) for v in mapf(self._call, (p[i] for i in range(len(p))))])
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def _call(self, x):
        return self.log_prob_fn(x, *self.args, **self.kwargs)

    def run_mcmc(self, initial_state, nsteps, progress=False, **_ignored):
        coords = np.array(initial_state, dtype=np.float64, copy=True)
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(coords.shape))
        if not walkers_independent(coords):
            raise ValueError("Initial state has a large condition number. Make sure that your "
                             "walkers are linearly independent for the best performance")
        log_prob = self.compute_log_prob(coords)
        if np.any(np.isnan(log_prob)):
            raise ValueError("The initial log_prob was NaN")
        i0 = self.iteration
        self._chain = np.concatenate((self._chain, np.empty((nsteps, self.nwalkers, self.ndim))))
        self._log_prob = np.concatenate((self._log_prob, np.empty((nsteps, self.nwalkers))))
        for it in range(nsteps):
            self._random.choice(1)                      # "choose a random move": one draw
            coords, log_prob, acc = self._move.propose(self, coords, log_prob)
            self._chain[i0 + it] = coords
            self._log_prob[i0 + it] = log_prob
            self.accepted += acc
            self.iteration += 1
        return coords, log_prob

    # backend.get_value slicing
    def _get(self, arr, flat=False, thin=1, discard=0):
        v = arr[discard + thin - 1:self.iteration:thin]
        if flat:
            s = list(v.shape[1:])
            s[0] = np.prod(v.shape[:2])
            return v.reshape(s)
        return v

    def get_chain(self, **kwargs):
        return self._get(self._chain, **kwargs)

    def get_log_prob(self, **kwargs):
        return self._get(self._log_prob, **kwargs)

    @property
    def acceptance_fraction(self):
        return self.accepted / float(self.iteration)

    @property
    def chain(self):
        return np.swapaxes(self._chain[:self.iteration], 0, 1)
