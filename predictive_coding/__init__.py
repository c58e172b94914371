"""``import predictive_coding as pc`` resolves here when this repository is first on sys.path, so the
reference's figure_*.py / table_1.py / utils/*.py pick up the B200 implementation unmodified."""
from montecarlopredictivecoding_b200.predictive_coding import PCLayer, PCTrainer  # noqa: F401
from montecarlopredictivecoding_b200.predictive_coding import layer as pc_layer  # noqa: F401
from montecarlopredictivecoding_b200.predictive_coding import trainer as pc_trainer  # noqa: F401

__all__ = ["PCLayer", "PCTrainer"]
