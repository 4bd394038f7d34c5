"""Exponential-family distribution interface over ``bayesic_b200.algebra`` expressions.

Working restatement of the reference's (non-importable) sketch ``bayesic/distribution/``:
``base.py`` (contract, iid wrapper, exponential families) and ``core.py`` (Normal,
MultivariateNormal); ``conjugate.py`` adds the families the BASELINE configurations name."""
from .base import (ConditionalDistribution, IndependentObservations, ExponentialFamily,  # noqa: F401
                   ExpFamIndependentObservations)
from .core import Normal, MultivariateNormal  # noqa: F401
from .conjugate import (BernoulliLogit, Exponential, Gamma, Categorical, Dirichlet,  # noqa: F401
                        Wishart, GaussianWishart)

__all__ = ['ConditionalDistribution', 'IndependentObservations', 'ExponentialFamily',
           'ExpFamIndependentObservations', 'Normal', 'MultivariateNormal',
           'BernoulliLogit', 'Exponential', 'Gamma', 'Categorical', 'Dirichlet', 'Wishart', 'GaussianWishart']
