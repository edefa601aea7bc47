# flake8: noqa  -- mirrors waveflow/flows/__init__.py and flows/bijections/__init__.py
from .bijections import Reverse, Serial
from .made import IMADE, BoxTransformLayer
from .neural_splines import unconstrained_RQS
from .distributions import Flow, MFlow, Normal, Uniform
