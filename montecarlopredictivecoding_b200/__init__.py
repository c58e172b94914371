"""B200-native (sm_100a) implementation of the Monte-Carlo predictive-coding hot path.

    from montecarlopredictivecoding_b200 import predictive_coding as pc      # PCLayer / PCTrainer
    from montecarlopredictivecoding_b200 import mcpc_utils                   # get_model, random_step, factories
"""
from . import predictive_coding  # noqa: F401
from .predictive_coding import PCLayer, PCTrainer  # noqa: F401

__version__ = "0.1.0"
