"""Parameter-space update steps either side of the data pass: the VMP global
update of a Gaussian mixture, the SVI natural-parameter blend, reparameterised draws / gradient
assembly and an Adam step -- each one device kernel behind the C-ABI, float64, never syncing
with the host, so a whole iteration is a stream of kernel launches.

The reference names these algorithms in prose only (``README.md:30-37`` VMP, ``:47-51``
reparameterised gradients, ``:69-80`` minibatch SVI); the formulas are the standard ones
(Bishop PRML 10.58-10.77, Hoffman et al. 2013, Kingma & Ba 2015); the parity tests hold float64
restatements of each.
"""

from . import stats
from .backend import library as L

__all__ = ['gmm_global_update', 'svi_natural_blend', 'reparam_draws', 'reparam_gradient', 'adam_step',
           'GmmVmp', 'LogisticReparamSgd']


def _f64(t, shape, what):
    torch = stats._torch()
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float64:
        raise TypeError("%s must be a float64 CUDA torch.Tensor" % what)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (what, tuple(shape), tuple(t.shape)))
    return t.contiguous()


def gmm_global_update(nk, rx, rxx, alpha0, beta0, nu0, m0, W0_inv):
    """VMP global step of a Gaussian mixture (``bb_gmm_global_update``): returns a dict with the
    posterior ``alpha, beta, nu [K]``, ``m [K, D]``, ``W_inv [K, D, D]`` (float64), the whitened
    logit parameters ``U [K, D, D]`` (upper triangular), ``t [K, D]``, ``c [K]`` (float32) that
    ``stats.mixture_logits(..., upper_triangular=True)`` consumes, ``kl [K + 1]`` (per-component
    Gaussian-Wishart KL, then the Dirichlet KL) and ``status`` (device int32; 0, or 1 + index of a
    component whose scale matrix is not positive definite).  Nothing is copied to the host."""
    torch = stats._torch()
    lib = L.load()
    k, d = rx.shape
    nk = _f64(nk, (k,), 'nk')
    rx = _f64(rx, (k, d), 'rx')
    rxx = _f64(rxx, (k, d, d), 'rxx')
    dev = rx.device
    m0 = _f64(m0, (d,), 'm0')
    W0_inv = _f64(W0_inv, (d, d), 'W0_inv')
    with torch.cuda.device(dev):
        f64 = dict(dtype=torch.float64, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        out = {'alpha': torch.empty(k, **f64), 'beta': torch.empty(k, **f64), 'nu': torch.empty(k, **f64),
               'm': torch.empty((k, d), **f64), 'W_inv': torch.empty((k, d, d), **f64),
               'U': torch.empty((k, d, d), **f32), 't': torch.empty((k, d), **f32), 'c': torch.empty(k, **f32)}
        kl = torch.empty(k + 2, **f64)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        L.check(lib.bb_gmm_global_update(nk.data_ptr(), rx.data_ptr(), rxx.data_ptr(), k, d, float(alpha0),
                                         float(beta0), float(nu0), m0.data_ptr(), W0_inv.data_ptr(),
                                         out['alpha'].data_ptr(), out['beta'].data_ptr(), out['nu'].data_ptr(),
                                         out['m'].data_ptr(), out['W_inv'].data_ptr(), out['U'].data_ptr(),
                                         out['t'].data_ptr(), out['c'].data_ptr(), kl.data_ptr(),
                                         status.data_ptr(), stats._stream(dev)), 'bb_gmm_global_update')
    out['kl'] = kl[:k + 1]
    out['status'] = status
    return out


def svi_natural_blend(eta, eta_prior, stat, scale, rho):
    """``eta <- (1 - rho) eta + rho (eta_prior + scale * stat)`` in place (float64 CUDA tensors of
    one shape); returns ``eta``."""
    torch = stats._torch()
    lib = L.load()
    if not eta.is_contiguous():
        raise ValueError("eta must be contiguous (it is updated in place)")
    eta = _f64(eta, None, 'eta')
    eta_prior = _f64(eta_prior, eta.shape, 'eta_prior')
    stat = _f64(stat, eta.shape, 'stat')
    with torch.cuda.device(eta.device):
        L.check(lib.bb_svi_natural_blend(eta.data_ptr(), eta_prior.data_ptr(), stat.data_ptr(), float(scale),
                                         float(rho), eta.numel(), stats._stream(eta.device)), 'bb_svi_natural_blend')
    return eta


def reparam_draws(mu, log_sigma, eps):
    """``W[s, :] = mu + exp(log_sigma) * eps[s, :]`` as float32 ``[S, D]`` (the operand layout of
    ``stats.logistic_reparam_stats``)."""
    torch = stats._torch()
    lib = L.load()
    s, d = eps.shape
    mu, log_sigma, eps = _f64(mu, (d,), 'mu'), _f64(log_sigma, (d,), 'log_sigma'), _f64(eps, (s, d), 'eps')
    with torch.cuda.device(mu.device):
        W = torch.empty((s, d), dtype=torch.float32, device=mu.device)
        L.check(lib.bb_reparam_draws(mu.data_ptr(), log_sigma.data_ptr(), eps.data_ptr(), d, s, W.data_ptr(),
                                     stats._stream(mu.device)), 'bb_reparam_draws')
    return W


def reparam_gradient(G, loglik, eps, mu, log_sigma):
    """``(elbo[1], grad_mu[D], grad_log_sigma[D])`` from ``G[D, S]`` and ``loglik[S]`` of the data
    pass, for q(w) = N(mu, diag sigma^2) and prior N(0, I)."""
    torch = stats._torch()
    lib = L.load()
    s, d = eps.shape
    G, loglik = _f64(G, (d, s), 'G'), _f64(loglik, (s,), 'loglik')
    mu, log_sigma, eps = _f64(mu, (d,), 'mu'), _f64(log_sigma, (d,), 'log_sigma'), _f64(eps, (s, d), 'eps')
    dev = G.device
    with torch.cuda.device(dev):
        grad_mu = torch.empty(d, dtype=torch.float64, device=dev)
        grad_ls = torch.empty(d, dtype=torch.float64, device=dev)
        elbo = torch.empty(1, dtype=torch.float64, device=dev)
        L.check(lib.bb_reparam_gradient(G.data_ptr(), loglik.data_ptr(), eps.data_ptr(), mu.data_ptr(),
                                        log_sigma.data_ptr(), d, s, grad_mu.data_ptr(), grad_ls.data_ptr(),
                                        elbo.data_ptr(), stats._stream(dev)), 'bb_reparam_gradient')
    return elbo, grad_mu, grad_ls


def adam_step(param, grad, m, v, step, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8, maximize=False):
    """One Adam step in place on float64 CUDA tensors (``step`` counts from 1)."""
    torch = stats._torch()
    lib = L.load()
    for name, t in (('param', param), ('m', m), ('v', v)):
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (it is updated in place)" % name)
    param = _f64(param, None, 'param')
    grad, m, v = _f64(grad, param.shape, 'grad'), _f64(m, param.shape, 'm'), _f64(v, param.shape, 'v')
    with torch.cuda.device(param.device):
        L.check(lib.bb_adam_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(),
                                 float(lr), float(beta1), float(beta2), float(eps), int(step), 1 if maximize else 0,
                                 stats._stream(param.device)), 'bb_adam_step')
    return param


class GmmVmp(object):
    """Full-batch VMP for a Gaussian mixture (cfg3 as a loop): local step on the tcgen05 kernels
    (logits, responsibilities in place, statistics), one all-reduce of the
    packed statistics when a process group is up, global step in one kernel.  State lives on the
    device; ``step`` enqueues kernels only and returns device tensors."""

    def __init__(self, k, d, alpha0, beta0, nu0, m0, W0_inv, group=None):
        from .parallel import PackedStats
        self.k, self.d = int(k), int(d)
        self.alpha0, self.beta0, self.nu0 = float(alpha0), float(beta0), float(nu0)
        self.m0, self.W0_inv = m0, W0_inv
        self.group = group
        self.layout = PackedStats.mixture(self.k, self.d)
        self.packed = self.layout.allocate(m0.device)
        self.state = None

    def initialise(self, nk, rx, rxx):
        """Global step from initial statistics (e.g. of a random soft assignment)."""
        self.state = gmm_global_update(nk, rx, rxx, self.alpha0, self.beta0, self.nu0, self.m0, self.W0_inv)
        return self.state

    def step(self, X_local):
        """One VMP iteration over this rank's rows.  Returns the new state dict with the ELBO
        (``sum_n lse - sum kl``, float64[1]) under key ``'elbo'``; the ELBO refers to the
        parameters the local step used."""
        from .parallel import allreduce_packed
        if self.state is None:
            raise RuntimeError("GmmVmp.initialise(...) first")
        prev = self.state
        from .passes import GmmStep
        out = GmmStep.local_step(X_local, prev['U'], prev['t'], prev['c'])
        nk, rx, rxx, sum_lse = out['nk'], out['rx'], out['rxx'], out['sum_lse']
        views = self.layout.views(self.packed)
        views['nk'].copy_(nk)
        views['rx'].copy_(rx)
        views['rxx'].copy_(rxx)
        views['sum_lse'].copy_(sum_lse)
        views['count'].fill_(float(X_local.shape[0]))
        allreduce_packed(self.packed, self.group)
        elbo = views['sum_lse'] - prev['kl'].sum()
        self.state = gmm_global_update(views['nk'], views['rx'], views['rxx'], self.alpha0, self.beta0, self.nu0,
                                       self.m0, self.W0_inv)
        self.state['elbo'] = elbo
        return self.state


class LogisticReparamSgd(object):
    """cfg5 as a loop: draws -> fused data pass -> gradient assembly -> Adam ascent on the ELBO,
    all on the device.  ``eps`` is fixed (common random numbers), as in BASELINE.json cfg5."""

    def __init__(self, mu, log_sigma, eps, lr=1e-2, group=None):
        torch = stats._torch()
        self.mu, self.log_sigma, self.eps = mu, log_sigma, eps
        self.lr, self.group = lr, group
        self.state = [torch.zeros_like(mu) for _ in range(4)]      # Adam moments of mu, log_sigma
        self.t = 0

    def step(self, X_local, y_local):
        from .parallel import PackedStats, allreduce_packed
        d, s = self.mu.shape[0], self.eps.shape[0]
        W = reparam_draws(self.mu, self.log_sigma, self.eps)
        loglik, G = stats.logistic_reparam_stats(X_local, y_local, W)
        layout = PackedStats.logistic(d, s)
        packed = layout.allocate(self.mu.device)
        views = layout.views(packed)
        views['G'].copy_(G)
        views['loglik'].copy_(loglik)
        views['count'].fill_(float(X_local.shape[0]))
        allreduce_packed(packed, self.group)
        elbo, grad_mu, grad_ls = reparam_gradient(views['G'], views['loglik'], self.eps, self.mu, self.log_sigma)
        self.t += 1
        adam_step(self.mu, grad_mu, self.state[0], self.state[1], self.t, lr=self.lr, maximize=True)
        adam_step(self.log_sigma, grad_ls, self.state[2], self.state[3], self.t, lr=self.lr, maximize=True)
        return {'elbo': elbo, 'grad_mu': grad_mu, 'grad_log_sigma': grad_ls}
