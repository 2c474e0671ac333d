"""lightgbm_b200 — B200-native histogram tree learner behind LightGBM's TreeLearner interface.

Only what the hot path needs lives here: csrc/ (sm_100a kernels + C-ABI, include/lgbm_b200.h) and the
host-side mirror of the reference interface (tree_learner.py, booster.py)."""
from .tree_learner import B200TreeLearner, Config, Layout, Tree  # noqa: F401
from .booster import B200Booster, train  # noqa: F401
from .model import Model, ModelTree  # noqa: F401
from .dataset import Binner, Dataset  # noqa: F401
from .distributed import make_row_sharded_learner, make_sharded_learner, shard_columns, shard_rows  # noqa: F401

__all__ = ["B200TreeLearner", "B200Booster", "Binner", "Config", "Dataset", "Layout", "Model", "ModelTree", "Tree", "train"]
