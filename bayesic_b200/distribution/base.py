"""Distribution contract, the independent-observations wrapper and exponential families.

Follows ``bayesic/distribution/base.py`` of the reference, which is an interface sketch that
does not parse (``{..)`` at :202-205, a comment after a line continuation at :217, a stray
``:`` at :223, ``raise`` for ``return`` at :98, undefined ``normalizers`` :231-233,
``iid_draw_dims`` :335, ``T`` :290).  This module implements what those lines clearly intend,
with tensor expressions from ``bayesic_b200.algebra`` instead of raw Theano so that the
statistics that come out are einsums the planner turns into data-axis contractions:

    log_likelihood = data_term + interaction_term - log_normalizer        base.py:25-100
    interaction_term = sum_i < s_i(data), eta_i(params) >                 base.py:271-291
    iid statistics  = per-point statistics summed over the iid axes       base.py:328-332
    iid normaliser  = per-draw normaliser x number of draws               base.py:226-244

The sum over the iid axes is THE data-parallel reduction of the hot path: for
``MultivariateNormal`` it canonicalises to ``einsum(out_uv = sum_i X_iu X_iv)``, i.e. the plan
``_tensordot(_dimshuffle(X,1,0), X, [1],[0])`` that the executor serves with the tcgen05
sufficient-statistics kernel.
"""
from .. import algebra as A

_DISCRETE_DTYPES = ('int8', 'int16', 'int32', 'int64')

__all__ = ['ConditionalDistribution', 'IndependentObservations', 'ExponentialFamily',
           'ExpFamIndependentObservations']


class ConditionalDistribution(object):
    """A parameterised family of distributions (``base.py:3-172``)."""

    @property
    def parameter_types(self):
        """``{parameter name: (dtype, ndim)}``"""
        raise NotImplementedError

    @property
    def data_type(self):
        """``(dtype, ndim)`` of one datum"""
        raise NotImplementedError

    def is_discrete(self):
        # the reference compares the (dtype, ndim) pair with the dtype list (base.py:23-24)
        return self.data_type[0] in _DISCRETE_DTYPES

    def log_likelihood(self, data, **params):
        """Expression for the normalised log-likelihood (``base.py:25-100``; the reference
        ``raise``s the sum at :98 where it means ``return``)."""
        return self.log_likelihood_data_term(data) \
            + self.log_likelihood_interaction_term(data, **params) \
            - self.log_normalizer(data_shape=data.shape, **params)

    def log_normalizer(self, data_shape, **params):
        raise NotImplementedError

    def log_likelihood_interaction_term(self, data, **params):
        raise NotImplementedError

    def log_likelihood_data_term(self, data):
        raise NotImplementedError

    def independent_observations(self, param_copy_ndim=1, iid_draw_ndim=0):
        return IndependentObservations(self, param_copy_ndim, iid_draw_ndim)

    def iid(self, extra_ndim=1):
        """Distribution of a tensor of iid draws (``base.py:166-172``)."""
        return self.independent_observations(param_copy_ndim=0, iid_draw_ndim=extra_ndim)


def _sum_all(expr):
    expr = A.wrap_if_literal(expr)
    return A.sum(expr) if expr.ndim > 0 else expr


class IndependentObservations(ConditionalDistribution):
    """Independent observations from ``distribution``: ``param_copy_ndim`` leading axes index
    copies of the parameters, the next ``iid_draw_ndim`` axes index iid draws from each copy
    (``base.py:176-258``)."""

    def __init__(self, distribution, param_copy_ndim=1, iid_draw_ndim=0):
        self.param_copy_ndim = param_copy_ndim
        self.iid_draw_ndim = iid_draw_ndim
        self.underlying = distribution

    @property
    def parameter_types(self):
        return {name: (dtype, self.param_copy_ndim + ndim)
                for name, (dtype, ndim) in self.underlying.parameter_types.items()}

    @property
    def data_type(self):
        dtype, ndim = self.underlying.data_type
        return dtype, self.param_copy_ndim + self.iid_draw_ndim + ndim

    @property
    def iid_draw_dims(self):
        return tuple(range(self.param_copy_ndim, self.param_copy_ndim + self.iid_draw_ndim))

    def _broadcast_params_over_iid_draws(self, params):
        """copies x param  ->  copies x (1,)*iid x param   (``base.py:208-224``)."""
        out = {}
        for name, (_, ndim) in self.underlying.parameter_types.items():
            param = A.wrap_if_literal(params[name])
            if self.iid_draw_ndim == 0 or param.ndim == 0:
                out[name] = param
                continue
            dims = list(range(self.param_copy_ndim)) + ['x'] * self.iid_draw_ndim + \
                [self.param_copy_ndim + d for d in range(ndim)]
            out[name] = A.dimshuffle(param, *dims)
        return out

    def num_draws(self, data_shape):
        """Number of iid draws per parameter copy, as an expression."""
        extents = [data_shape[a] for a in self.iid_draw_dims]
        return A.mul(*extents) if extents else A.constant(1)

    def log_normalizer(self, data_shape, **params):
        single_datum_shape = tuple(data_shape[self.param_copy_ndim + self.iid_draw_ndim:])
        per_copy = self.underlying.log_normalizer(data_shape=single_datum_shape, **params)
        total = _sum_all(per_copy) if self.param_copy_ndim > 0 else A.wrap_if_literal(per_copy)
        if self.iid_draw_ndim > 0:
            # computed once, multiplied by the number of draws (base.py:235-242)
            return total * self.num_draws(data_shape)
        return total

    def log_likelihood_interaction_term(self, data, **params):
        broadcast = self._broadcast_params_over_iid_draws(params)
        return _sum_all(self.underlying.log_likelihood_interaction_term(data, **broadcast))

    def log_likelihood_data_term(self, data):
        return _sum_all(self.underlying.log_likelihood_data_term(data))


class ExponentialFamily(ConditionalDistribution):
    """``interaction_term = <sufficient_statistics(data), natural_parameters(params)>``
    (``base.py:263-325``)."""

    def log_likelihood_interaction_term(self, data, **params):
        stats = self.sufficient_statistics(data)
        naturals = self.natural_parameters(**params)
        terms = [self._pair(s, eta) for s, eta in zip(stats, naturals)]
        return A.add(*terms) if len(terms) > 1 else terms[0]

    def _pair(self, stat, natural):
        """<s, eta> over the datum axes, keeping any leading observation axes: the flattened dot
        product of ``base.py:279-291`` written as an einsum."""
        stat, natural = A.wrap_if_literal(stat), A.wrap_if_literal(natural)
        lead = stat.ndim - natural.ndim if natural.ndim <= stat.ndim else 0
        if stat.ndim == natural.ndim:
            # same rank: leading observation axes (if any) are explicit on both sides
            core = self._core_ndim(stat)
            lead = stat.ndim - core
            s_idx = [('out', i) for i in range(lead)] + [('sum', i) for i in range(core)]
            return A.einsum([(stat, s_idx), (natural, s_idx)], lead)
        s_idx = [('out', i) for i in range(lead)] + [('sum', i) for i in range(natural.ndim)]
        n_idx = [('sum', i) for i in range(natural.ndim)]
        return A.einsum([(stat, s_idx), (natural, n_idx)], lead)

    def _core_ndim(self, stat):
        return stat.ndim

    def sufficient_statistics(self, data):
        raise NotImplementedError

    def natural_parameters(self, **params):
        raise NotImplementedError

    def independent_observations(self, param_copy_ndim=1, iid_draw_ndim=0):
        return ExpFamIndependentObservations(self, param_copy_ndim, iid_draw_ndim)


class ExpFamIndependentObservations(IndependentObservations):
    """iid exponential-family observations are again an exponential family whose statistics are
    the per-point statistics summed over the iid axes (``base.py:328-335``)."""

    def sufficient_statistics(self, data):
        per_point = self.underlying.sufficient_statistics(data)
        if not self.iid_draw_dims:
            return tuple(per_point)
        return tuple(A.sum(s, axis=self.iid_draw_dims) for s in per_point)

    def natural_parameters(self, **params):
        # one natural parameter per copy; nothing to sum (the reference's :334-335 sums over an
        # undefined name)
        return tuple(self.underlying.natural_parameters(**params))

    def log_likelihood_interaction_term(self, data, **params):
        """<summed statistics, natural parameters>: the sufficient-statistic form, so the
        contraction over the data axis happens once and first."""
        if self.param_copy_ndim > 0:
            return IndependentObservations.log_likelihood_interaction_term(self, data, **params)
        stats = self.sufficient_statistics(data)
        naturals = self.natural_parameters(**params)
        terms = []
        for s, eta in zip(stats, naturals):
            s, eta = A.wrap_if_literal(s), A.wrap_if_literal(eta)
            idx = [('sum', i) for i in range(s.ndim)]
            terms.append(A.einsum([(s, idx), (eta, idx)], 0))
        return A.add(*terms) if len(terms) > 1 else terms[0]

    def compile_sufficient_statistics(self, data_var, **options):
        """One plan computing every summed statistic in a single call: ``f(**inputs)`` returns
        the tuple of statistics (device-resident when the data is)."""
        from ..backend.compiled import compile_many
        return compile_many(self.sufficient_statistics(data_var), **options)
