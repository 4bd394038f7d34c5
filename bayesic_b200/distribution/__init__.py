"""Exponential-family distribution interface over ``bayesic_b200.algebra`` expressions.

Working restatement of the reference's (non-importable) sketch ``bayesic/distribution/``:
``base.py`` (contract, iid wrapper, exponential families) and ``core.py`` (Normal,
MultivariateNormal)."""
from .base import (ConditionalDistribution, IndependentObservations, ExponentialFamily,  # noqa: F401
                   ExpFamIndependentObservations)
from .core import Normal, MultivariateNormal  # noqa: F401

__all__ = ['ConditionalDistribution', 'IndependentObservations', 'ExponentialFamily',
           'ExpFamIndependentObservations', 'Normal', 'MultivariateNormal']
